"""The oracle must reproduce the reference's own outputs (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference).  CPU only."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import decode as odec
from oracle import net as onet
from oracle import preprocess as opre
from oracle import synth

from golden.make_golden_cases import POSE_CASES, DEC_CASES, NET_CASES, PRE_CASES, heads_for


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def g_pre(golden_dir):
    return np.load(os.path.join(golden_dir, "preprocess.npz"))


@pytest.fixture(scope="module")
def g_net(golden_dir):
    return np.load(os.path.join(golden_dir, "net.npz"))


@pytest.fixture(scope="module")
def g_dec(golden_dir):
    return np.load(os.path.join(golden_dir, "decode.npz"))


@pytest.mark.parametrize("i", range(len(PRE_CASES)))
def test_preprocess_bit_exact(g_pre, i):
    h, w, sf, os_, seed, full = PRE_CASES[i]
    img = synth.noise_image(h, w, seed)
    assert sha(img) == str(g_pre["in_sha_%d" % i]), "synthetic input drifted"
    x, src, scale = opre.process_input(img, sf, os_)
    assert src is img
    assert x.dtype == np.float32 and tuple(x.shape) == tuple(g_pre["shape_%d" % i])
    assert tuple(opre.valid_resolution(w * sf, h * sf, os_)) == tuple(g_pre["vres_%d" % i])
    assert np.array_equal(scale, g_pre["scale_%d" % i])
    assert sha(x) == str(g_pre["sha_%d" % i])
    if full:
        assert np.array_equal(x, g_pre["x_%d" % i])


@pytest.mark.parametrize("i", range(len(NET_CASES)))
def test_net_matches_reference(g_net, i):
    mid, os_, H, W, N, scheme, gain, seed = NET_CASES[i]
    sd = onet.init_params(mid, seed, scheme, gain)
    assert sha(np.concatenate([v.numpy().ravel() for v in sd.values()])) == str(g_net["w_sha_%d" % i]), \
        "seeded weights drifted (torch RNG changed?)"
    tab = [(L["stride"], L["dilation"], L["padding"]) for L in onet.layer_table(mid, os_)]
    assert np.array_equal(np.array(tab), g_net["table_%d" % i])
    x = torch.from_numpy(np.stack([
        opre.process_input(synth.smooth_image(H, W, 100 * seed + b), 1.0, os_)[0][0] for b in range(N)]))
    assert sha(x.numpy()) == str(g_net["x_sha_%d" % i])
    heads = onet.forward(sd, mid, os_, x)
    assert tuple(heads[0].shape[2:]) == onet.out_hw(mid, os_, H, W)
    for nm, t in zip(("heat", "off", "fwd", "bwd"), heads):
        ref = g_net["%s_%d" % (nm, i)]
        assert t.shape == ref.shape
        # Same torch ops in the same order: identical up to oneDNN's thread-count dependent blocking.
        # gain >= 1 nets are chaotic: last-bit differences grow ~100x through 27 layers.
        tol = 5e-4 if (scheme == "scaled" and gain >= 1.0) else 1e-5
        np.testing.assert_allclose(t.numpy(), ref, rtol=0, atol=tol * max(1.0, np.abs(ref).max()))


def test_param_shapes_match_reference_keys():
    # 62 tensors; parameter counts for 101 / 75 / 50
    for mid, count in ((101, 3313907), (75, 1258195), (50, 577459)):
        shp = onet.param_shapes(mid)
        assert len(shp) == 62
        assert sum(int(np.prod(s)) for s in shp.values()) == count


@pytest.mark.parametrize("i", range(len(DEC_CASES)))
def test_decode_bit_exact(g_dec, i):
    kind, h, w, stride, people, seed, P, thr, rad, minp, _patch, extra = DEC_CASES[i]
    heat, off, fwd, bwd = heads_for(kind, h, w, stride, people, seed, extra)
    assert sha(np.concatenate([heat.ravel(), off.ravel(), fwd.ravel(), bwd.ravel()])) == str(g_dec["in_sha_%d" % i])
    cs, ci = odec.part_candidates(heat, thr)
    assert np.array_equal(cs, g_dec["cand_s_%d" % i])
    assert np.array_equal(ci, g_dec["cand_i_%d" % i].reshape(-1, 3))
    res = odec.decode_multiple_poses(heat, off, fwd, bwd, stride, max_pose_detections=P,
                                     score_threshold=thr, nms_radius=rad, min_pose_score=minp)
    for nm, a in zip(("ps", "ks", "kc", "ko"), res):
        ref = g_dec["%s_%d" % (nm, i)]
        assert a.dtype == np.float64 and a.shape == ref.shape
        assert np.array_equal(a, ref), "%s differs (case %d)" % (nm, i)


def test_people_generator_is_decodable():
    # the synthetic people decode back to their ground truth
    heat, off, fwd, bwd, kps = synth.people_heads(91, 161, 8, 10, seed=2)
    ps, ks, kc, ko = odec.decode_multiple_poses(heat, off, fwd, bwd, 8, max_pose_detections=50,
                                                min_pose_score=0.25)
    found = int((ps > 0).sum())
    assert found >= 8
    for p in range(found):
        d = np.abs(kps - kc[p][None]).reshape(len(kps), -1).max(axis=1)
        assert d.min() < 1e-3


@pytest.mark.parametrize("ci", range(len(POSE_CASES)))
def test_decode_pose_and_traverse_bit_exact(golden_dir, ci):
    """decode.py:9-63,131-182 called on their own: the oracle against the reference's outputs (decode_pose.npz)."""
    g = np.load(os.path.join(golden_dir, "decode_pose.npz"))
    di, _ = POSE_CASES[ci]
    kind, h, w, stride, people, seed, P, thr, rad, minp, _patch, extra = DEC_CASES[di]
    heat, off, fwd, bwd = heads_for(kind, h, w, stride, people, seed, extra)
    assert sha(np.concatenate([heat.ravel(), off.ravel(), fwd.ravel(), bwd.ravel()])) == str(g["in_sha_%d" % ci])
    split = lambda a: a.reshape(2, -1, h, w).transpose(1, 2, 3, 0)
    offs, fwd_t, bwd_t = split(off), split(fwd), split(bwd)
    roots, t = g["roots_%d" % ci], 0
    assert len(roots) >= 3 and (roots[:, 0] == 0.0).any() and (roots[:, 0] < 0.0).any()
    for r, (rs, rid, ry, rx) in enumerate(roots):
        ks, kc, ko = odec.decode_pose(np.float32(rs), int(rid), np.array([ry, rx]), heat, offs, stride, fwd_t, bwd_t)
        assert np.array_equal(ks, g["ks_%d" % ci][r]) and np.array_equal(kc, g["kc_%d" % ci][r]) and np.array_equal(ko, g["ko_%d" % ci][r])
        for e, (parent, child) in enumerate(odec.EDGES):
            for tgt, disp in ((child, fwd_t), (parent, bwd_t)):
                sc, xy, dv, ov = odec.traverse_to_targ_keypoint(e, np.array([ry, rx]), tgt, heat, offs, stride, disp)
                assert np.array_equal(np.concatenate([[sc], xy, dv, ov]).astype(np.float64), g["tr_%d" % ci][t]), (r, e, tgt)
                t += 1
    assert t == len(g["tr_%d" % ci])
