"""BASELINE.json's configurations at their FULL sizes (configs[1..3]: model 101 / 513x513 / OS16 / batch 64, model 50 /
1280x720 frames -> 721x1281 / OS8 / batch 32, model 75 / 257x257 / OS32 / batch 512), checked through size-independent
properties -- the oracle cannot run these sizes in test time:

* batch invariance: every image of the full batch gets bit-for-bit the head tensors and pose records it gets when it is
  processed alone (the full-size launch partitions its tiles / chunks / items differently: persistent CTAs wrap around
  the work list several times, index arithmetic reaches its largest values);
* permutation: reversing the batch reverses the results;
* anchoring: three images of the full batch (incl. the last) are compared with the oracle (reference semantics) at the bf16 tolerance, and the
  decoder's answer on the full-size head tensors with the oracle decoder, bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import posenet  # noqa: E402
from oracle import decode as odec  # noqa: E402
from oracle import net as onet  # noqa: E402
from oracle import preprocess as opre  # noqa: E402

DEV = "cuda"
KW = dict(max_pose_detections=10, score_threshold=0.5, nms_radius=20, min_pose_score=0.25)      # benchmark.py:37-44

CONFIGS = [  # name, model, output stride, frame (h, w), batch
    ("configs[1]", 101, 16, (513, 513), 64),
    ("configs[2]", 50, 8, (720, 1280), 32),
    ("configs[3]", 75, 32, (257, 257), 512),
]


def run(model, frames, os_):
    """uint8 frames [n, h, w, 3] on the device -> (heads tuple, pose records tuple), resized on the GPU when needed."""
    x, _ = posenet.resize_u8_gpu(frames, 1.0, os_)
    heads = [t.clone() for t in model.forward_u8(x)]
    rec = [t.clone() for t in posenet.decode_multiple_poses_batch(*heads, output_stride=os_, **KW)]
    return x, heads, rec


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_full_size_configuration(cfg):
    name, mid, os_, (h, w), batch = cfg
    sd = onet.init_params(mid, seed=0)
    model = posenet.MobileNetV1(mid, output_stride=os_)
    model.load_state_dict(sd)
    model = model.cuda().set_compute_dtype("bf16")
    rng = np.random.default_rng(batch)
    frames = torch.from_numpy(rng.integers(0, 256, (batch, h, w, 3), dtype=np.uint8)).to(DEV)
    x, heads, rec = run(model, frames, os_)
    assert tuple(heads[0].shape[:2]) == (batch, 17) and int((rec[4] >= 0).sum()) == batch
    assert all(torch.isfinite(t).all() for t in heads)

    # batch invariance on the first, a middle and the last image
    for i in (0, batch // 2 + 1, batch - 1):
        _, h1, r1 = run(model, frames[i:i + 1], os_)
        for a, b in zip(heads, h1):
            assert torch.equal(a[i:i + 1], b), (name, i)
        for a, b in zip(rec, r1):
            assert torch.equal(a[i:i + 1], b), (name, i)

    # permutation
    _, heads_r, rec_r = run(model, torch.flip(frames, dims=[0]), os_)
    for a, b in zip(heads + rec, heads_r + rec_r):
        assert torch.equal(a, torch.flip(b, dims=[0])), name

    # anchoring on the oracle: three images of the batch (an early one, a middle one and the LAST), reference semantics on the
    # same pixels -- head tensors at the bf16 tolerance, the decoder's answer on the full-size head tensors bit for bit
    for i in (3, batch // 2, batch - 1):
        img = frames[i].cpu().numpy()
        x_ref, _, _ = opre.process_input(img, 1.0, os_)
        assert np.array_equal(x[i].cpu().numpy(), opre.resize_linear_u8(img, x.shape[2], x.shape[1])) or (x.shape[1], x.shape[2]) == (h, w)
        ref = onet.forward(sd, mid, os_, torch.from_numpy(x_ref))
        for g, r in zip(heads, ref):
            err = float((g[i:i + 1].cpu().double() - r.double()).abs().max() / r.double().abs().max())
            assert err < 2e-2, (name, i, err)
        want = odec.decode_multiple_poses(*[t[i].cpu().numpy() for t in heads], os_, **KW)
        for a, b in zip(rec[:4], want):
            assert np.array_equal(a[i].cpu().numpy(), b), (name, i)
