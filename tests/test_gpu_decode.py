"""Candidate extraction and greedy decode on the B200 must be BIT-EXACT against the oracle
(which is pinned to the reference by tests/golden/decode.npz) when fed the same head tensors."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import abi  # noqa: E402
import posenet  # noqa: E402
from golden.make_golden_cases import DEC_CASES, POSE_CASES, heads_for  # noqa: E402
from oracle import decode as odec  # noqa: E402
from oracle import synth  # noqa: E402

DEV = "cuda"


def gpu_decode(heads, stride, **kw):
    return posenet.decode_multiple_poses(*[torch.from_numpy(t).to(DEV) for t in heads], output_stride=stride, **kw)


def assert_same(res, ref, what=""):
    for nm, a, b in zip(("pose_scores", "keypoint_scores", "keypoint_coords", "pose_offsets"), res, ref):
        assert a.dtype == np.float64 and a.shape == b.shape, nm
        assert np.array_equal(a, b), "%s %s: max diff %g" % (what, nm, np.abs(a - b).max())


@pytest.mark.parametrize("i", range(len(DEC_CASES)))
def test_candidates_exact(i):
    kind, h, w, stride, people, seed, P, thr, rad, minp, _patch, extra = DEC_CASES[i]
    heat = heads_for(kind, h, w, stride, people, seed, extra)[0]
    cs, ci = odec.part_candidates(heat, thr)
    keys, counts = abi.candidates(torch.from_numpy(heat).unsqueeze(0).to(DEV), thr)
    n = int(counts[0])
    assert n == len(cs)
    k = np.sort(keys[0, :n].cpu().numpy().view(np.uint64))
    flat = (k & np.uint64(0xFFFFFFFF)).astype(np.int64)
    got = np.stack([flat // (h * w), (flat % (h * w)) // w, flat % w], axis=1).reshape(-1, 3)
    assert np.array_equal(got, ci.reshape(-1, 3))          # same cells, same (score desc, index asc) order


@pytest.mark.parametrize("i", range(len(DEC_CASES)))
def test_decode_golden(golden_dir, i):
    kind, h, w, stride, people, seed, P, thr, rad, minp, _patch, extra = DEC_CASES[i]
    g = np.load(os.path.join(golden_dir, "decode.npz"))
    heads = heads_for(kind, h, w, stride, people, seed, extra)
    res = gpu_decode(heads, stride, max_pose_detections=P, score_threshold=thr, nms_radius=rad, min_pose_score=minp)
    assert_same(res, [g["%s_%d" % (nm, i)] for nm in ("ps", "ks", "kc", "ko")], "golden case %d" % i)
    assert all(a.flags.writeable for a in res)


@pytest.mark.parametrize("seed", range(12))
def test_decode_random_vs_oracle(seed):
    rng = np.random.default_rng(1000 + seed)
    h, w = int(rng.integers(3, 40)), int(rng.integers(3, 40))
    stride = int(rng.choice([8, 16, 32]))
    heads = synth.random_heads(h, w, seed=seed, disp_scale=float(rng.uniform(5, 120)), off_scale=float(rng.uniform(1, 30)),
                               zero_frac=float(rng.choice([0.0, 0.1])), tie_levels=int(rng.choice([0, 0, 8, 64])))
    kw = dict(max_pose_detections=int(rng.integers(1, 40)), score_threshold=float(rng.uniform(0.0, 0.95)),
              nms_radius=float(rng.choice([5, 20, 33.3])), min_pose_score=float(rng.choice([0.0, 0.1, 0.3])))
    assert_same(gpu_decode(heads, stride, **kw), odec.decode_multiple_poses(*heads, stride, **kw), "seed %d %r" % (seed, kw))


def test_decode_many_candidates_radix_path():
    # plateaus: > 4096 candidates per image forces the multi-round radix selection
    heads = synth.random_heads(61, 83, seed=3, tie_levels=6, disp_scale=60.0)
    kw = dict(max_pose_detections=25, score_threshold=0.3, nms_radius=20, min_pose_score=0.45)
    cs, _ = odec.part_candidates(heads[0], 0.3)
    assert len(cs) > 3 * 4096
    assert_same(gpu_decode(heads, 8, **kw), odec.decode_multiple_poses(*heads, 8, **kw))


def test_decode_stress_config5():
    # BASELINE config 5: 91x161 map, OS8, 50 people, max_pose_detections=50
    heads = synth.people_heads(91, 161, 8, 50, seed=4)[:4]
    kw = dict(max_pose_detections=50, score_threshold=0.5, nms_radius=20, min_pose_score=0.25)
    assert_same(gpu_decode(heads, 8, **kw), odec.decode_multiple_poses(*heads, 8, **kw))


def test_decode_channels_last_views_and_batch():
    # the model returns NCHW tensors; callers may also hand over channels-last strided views
    sets = [synth.people_heads(33, 33, 16, 4, seed=s)[:4] for s in range(3)]
    batched = [torch.from_numpy(np.stack([s[j] for s in sets])).to(DEV) for j in range(4)]
    strided = [t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2) for t in batched]
    assert not strided[0].is_contiguous()
    for variant in (batched, strided):
        ps, ks, kc, ko, cnt = posenet.decode_multiple_poses_batch(*variant, output_stride=16, min_pose_score=0.25)
        for b in range(3):
            ref = odec.decode_multiple_poses(*sets[b], 16, min_pose_score=0.25)
            assert_same([ps[b].cpu().numpy(), ks[b].cpu().numpy(), kc[b].cpu().numpy(), ko[b].cpu().numpy()], ref)
            assert int(cnt[b]) == int((ref[0] != 0).sum())


def test_decode_accepts_cpu_tensors_and_image_demo_usage():
    heads = synth.people_heads(33, 33, 16, 3, seed=0)[:4]
    res = posenet.decode_multi.decode_multiple_poses(*[torch.from_numpy(t) for t in heads], output_stride=16,
                                                     max_pose_detections=10, min_pose_score=0.25)
    assert len(res) == 4                                  # this fork returns a 4-tuple
    res[2][...] *= np.array([1.5, 0.5])                   # image_demo.py:50 scales the coords in place
    assert_same(res[:2], odec.decode_multiple_poses(*heads, 16, max_pose_detections=10, min_pose_score=0.25)[:2])


def test_decode_zero_pads_an_uninitialised_record_buffer():
    # decode_multi.py:94-100: the reference's outputs are np.zeros; pn_decode_greedy pads the rows it does not fill itself,
    # whatever the buffer held before -- including images with no candidate at all (threshold above every score)
    sets = [synth.people_heads(33, 33, 16, p, seed=10 + p)[:4] for p in (1, 2, 5)]
    batched = [torch.from_numpy(np.stack([s[j] for s in sets])).to(DEV) for j in range(4)]
    P = 7
    for thr in (0.5, 2.0):
        out = torch.full((3 * P * 86,), float("nan"), dtype=torch.float64, device=DEV)
        ps, ks, kc, ko, cnt = posenet.decode_multiple_poses_batch(*batched, output_stride=16, max_pose_detections=P,
                                                                  score_threshold=thr, min_pose_score=0.25, out=out)
        assert not torch.isnan(out).any()
        for b in range(3):
            ref = odec.decode_multiple_poses(*sets[b], 16, max_pose_detections=P, score_threshold=thr, min_pose_score=0.25)
            assert_same([ps[b].cpu().numpy(), ks[b].cpu().numpy(), kc[b].cpu().numpy(), ko[b].cpu().numpy()], ref)
            assert int(cnt[b]) == int((ref[0] != 0).sum())


def test_decode_more_poses_than_the_shared_memory_cache():
    """DEC_ACC = 64 accepted poses are cached in shared memory (csrc/decode.cu); later ones are re-read from the output
    arrays by the screen and the commit loop.  Golden cases 14 / 15 (100 and 86 poses) pin that path to the reference;
    here it is also checked inside a batch next to images with few poses."""
    sets = [synth.people_heads(91, 161, 8, p, seed=s)[:4] for p, s in ((120, 20), (3, 1), (90, 21), (70, 5))]
    batched = [torch.from_numpy(np.stack([s[j] for s in sets])).to(DEV) for j in range(4)]
    kw = dict(max_pose_detections=100, score_threshold=0.5, nms_radius=20, min_pose_score=0.25)
    ps, ks, kc, ko, cnt = posenet.decode_multiple_poses_batch(*batched, output_stride=8, **kw)
    counts = []
    for b in range(len(sets)):
        ref = odec.decode_multiple_poses(*sets[b], 8, **kw)
        assert_same([ps[b].cpu().numpy(), ks[b].cpu().numpy(), kc[b].cpu().numpy(), ko[b].cpu().numpy()], ref, "image %d" % b)
        counts.append(int(cnt[b]))
        assert counts[-1] == int((ref[0] != 0).sum())
    assert counts[0] == 100 and counts[2] > 80 and counts[3] > 64 and counts[1] <= 3


@pytest.mark.parametrize("ci", range(len(POSE_CASES)))
def test_decode_pose_and_traverse_golden(golden_dir, ci):
    """posenet.decode.decode_pose / traverse_to_targ_keypoint (decode.py:131-182, :9-63) with the reference's argument
    layouts, bit-exact against the reference's own outputs (tests/golden/decode_pose.npz)."""
    g = np.load(os.path.join(golden_dir, "decode_pose.npz"))
    di, _ = POSE_CASES[ci]
    kind, h, w, stride, people, seed, P, thr, rad, minp, _patch, extra = DEC_CASES[di]
    heat, off, fwd, bwd = heads_for(kind, h, w, stride, people, seed, extra)
    split = lambda a: np.ascontiguousarray(a.reshape(2, -1, h, w).transpose(1, 2, 3, 0))     # decode_multi.py:89-97
    offs, fwd_t, bwd_t = split(off), split(fwd), split(bwd)
    dev = [torch.from_numpy(a).to(DEV) for a in (heat, offs, fwd_t, bwd_t)]
    roots, t = g["roots_%d" % ci], 0
    for r, (rs, rid, ry, rx) in enumerate(roots):
        args = (np.float32(rs), int(rid), np.array([ry, rx]))
        # numpy arrays in the reference's transposed layout, CUDA tensors, and the network's planar [34|32,h,w] tensors
        for maps in ((heat, offs, fwd_t, bwd_t), dev, (heat, off, fwd, bwd)) if r < 2 else ((heat, offs, fwd_t, bwd_t),):
            ks, kc, ko = posenet.decode.decode_pose(*args, maps[0], maps[1], stride, maps[2], maps[3])
            assert ks.dtype == kc.dtype == ko.dtype == np.float64 and kc.shape == ko.shape == (17, 2)
            assert np.array_equal(ks, g["ks_%d" % ci][r]) and np.array_equal(kc, g["kc_%d" % ci][r]) and np.array_equal(ko, g["ko_%d" % ci][r])
        for e, (parent, child) in enumerate(posenet.PARENT_CHILD_TUPLES):
            for tgt, disp in ((child, fwd_t), (parent, bwd_t)):
                if r < 3 or (e + r) % 5 == 0:                  # every hop for the first roots, a sample for the rest
                    sc, xy, dv, ov = posenet.decode.traverse_to_targ_keypoint(e, np.array([ry, rx]), tgt, heat, offs, stride, disp)
                    assert sc.dtype == np.float32 and xy.dtype == np.float64 and dv.dtype == np.float32 and ov.dtype == np.float32
                    assert np.array_equal(np.concatenate([[sc], xy, dv, ov]).astype(np.float64), g["tr_%d" % ci][t]), (r, e, tgt)
                t += 1


def test_decode_pose_agrees_with_decode_multiple_poses():
    """The first accepted pose of decode_multiple_poses IS decode_pose of the best candidate (decode_multi.py:104-121)."""
    heat, off, fwd, bwd = synth.people_heads(33, 33, 16, 4, seed=3)[:4]
    res = gpu_decode((heat, off, fwd, bwd), 16, min_pose_score=0.0)
    cs, ci = odec.part_candidates(heat, 0.5)
    k, y, x = [int(v) for v in ci[0]]
    root = np.array([y, x]) * 16 + np.array([off[k, y, x], off[17 + k, y, x]])
    ks, kc, ko = posenet.decode.decode_pose(cs[0], k, root, heat, off, 16, fwd, bwd)
    assert np.array_equal(ks, res[1][0]) and np.array_equal(kc, res[2][0]) and np.array_equal(ko, res[3][0])
