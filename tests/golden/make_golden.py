"""Generate the golden fixtures that pin ``oracle/`` to the reference implementation.

Run ONLY in the build container (needs the read-only reference checkout):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports the unmodified reference package from /root/reference, runs the reference
functions of the hot path on seeded synthetic inputs and stores their outputs as small
``.npz`` files next to this script.  The inputs are NOT stored: tests regenerate them from
the same seeds through ``oracle.synth`` / ``oracle.net.init_params`` (a checksum of each
regenerated input is stored so a drifting RNG is detected rather than mis-reported).
Nothing here is read at test time except the ``.npz`` files.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("POSENET_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import posenet as ref                      # noqa: E402  (the reference, NOT the product package)
import posenet.decode_multi as ref_dm      # noqa: E402

from oracle import net as onet             # noqa: E402
from oracle import synth                   # noqa: E402
sys.path.insert(0, HERE)
from make_golden_cases import DEC_CASES, NET_CASES, POSE_CASES, PRE_CASES, heads_for as _heads  # noqa: E402
import posenet.decode as ref_dec           # noqa: E402

assert os.path.realpath(ref.__file__).startswith(os.path.realpath(REF)), ref.__file__


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ----------------------------------------------------------------------------- preprocess
def make_preprocess():
    out = {"cases": np.array([(h, w, sf, os_, seed) for h, w, sf, os_, seed, _ in PRE_CASES], dtype=np.float64)}
    for i, (h, w, sf, os_, seed, full) in enumerate(PRE_CASES):
        img = synth.noise_image(h, w, seed)
        x, src, scale = ref.utils._process_input(img, sf, os_)
        assert src is img
        out["in_sha_%d" % i] = np.array(sha(img))
        out["shape_%d" % i] = np.array(x.shape)
        out["sha_%d" % i] = np.array(sha(x))
        out["scale_%d" % i] = scale
        out["vres_%d" % i] = np.array(ref.valid_resolution(w * sf, h * sf, os_))
        if full:
            out["x_%d" % i] = x
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)
    print("preprocess: %d cases" % len(PRE_CASES))


# ----------------------------------------------------------------------------- network
def make_net():
    out = {"cases": np.array([(m, o, h, w, n, 0 if s == "default" else 1, g, sd)
                              for m, o, h, w, n, s, g, sd in NET_CASES], dtype=np.float64)}
    for i, (mid, os_, H, W, N, scheme, gain, seed) in enumerate(NET_CASES):
        sd = onet.init_params(mid, seed, scheme, gain)
        model = ref.MobileNetV1(mid, output_stride=os_)
        model.load_state_dict(sd, strict=True)      # same key names/shapes as the reference
        model.eval()
        x = torch.from_numpy(np.stack([
            ref.utils._process_input(synth.smooth_image(H, W, 100 * seed + b), 1.0, os_)[0][0]
            for b in range(N)]))
        assert tuple(x.shape) == (N, 3, H, W), x.shape
        with torch.no_grad():
            heads = model(x)
        # layer table as the reference module realised it
        tab = []
        for name, m in model.features.named_children():
            conv = m.conv if hasattr(m, "conv") else m.depthwise
            tab.append((conv.stride[0], conv.dilation[0], conv.padding[0]))
        out["table_%d" % i] = np.array(tab)
        out["w_sha_%d" % i] = np.array(sha(np.concatenate([v.numpy().ravel() for v in sd.values()])))
        out["x_sha_%d" % i] = np.array(sha(x.numpy()))
        for nm, t in zip(("heat", "off", "fwd", "bwd"), heads):
            out["%s_%d" % (nm, i)] = t.numpy()
    np.savez_compressed(os.path.join(HERE, "net.npz"), **out)
    print("net: %d cases" % len(NET_CASES))


# ----------------------------------------------------------------------------- decode
def make_decode():
    out = {"n": np.array(len(DEC_CASES))}
    real_argsort = torch.argsort
    for i, (kind, h, w, stride, people, seed, P, thr, rad, minp, patch, extra) in enumerate(DEC_CASES):
        heat, off, fwd, bwd = _heads(kind, h, w, stride, people, seed, extra)
        # The reference's argsort is unstable (decode_multi.py:33): where scores tie, the fixture is
        # produced with the sort forced stable (ties -> ascending flat index), the order the oracle defines.
        cs, _ = ref_dm.build_part_with_score_torch(thr, 1, torch.from_numpy(heat))
        ties = len(np.unique(cs.numpy())) != len(cs)
        assert ties or not patch, "case %d was expected to contain ties" % i
        if ties:
            torch.argsort = lambda t, descending=False, **kw: real_argsort(t, descending=descending, stable=True)
        try:
            cs, ci = ref_dm.build_part_with_score_torch(thr, 1, torch.from_numpy(heat))
            res = ref_dm.decode_multiple_poses(
                torch.from_numpy(heat), torch.from_numpy(off), torch.from_numpy(fwd), torch.from_numpy(bwd),
                output_stride=stride, max_pose_detections=P, score_threshold=thr, nms_radius=rad,
                min_pose_score=minp)
        finally:
            torch.argsort = real_argsort
        out["in_sha_%d" % i] = np.array(sha(np.concatenate([heat.ravel(), off.ravel(), fwd.ravel(), bwd.ravel()])))
        out["cand_s_%d" % i] = cs.numpy()
        out["cand_i_%d" % i] = ci.numpy()
        for nm, a in zip(("ps", "ks", "kc", "ko"), res):
            assert a.dtype == np.float64
            out["%s_%d" % (nm, i)] = a
        out["ties_%d" % i] = np.array(ties)
        print("decode case %2d: %4d candidates, %2d poses, ties=%s" % (i, len(cs), int((res[0] != 0).sum()), ties))
    np.savez_compressed(os.path.join(HERE, "decode.npz"), **out)


# ----------------------------------------------------------------------------- decode_pose / traverse on their own
def pose_roots(heat, off, stride, thr, n_roots):
    """Roots for the stand-alone decode_pose fixtures: the n best candidates in the oracle's order (score desc, flat index
    asc -- no reference sort involved), then one root with score 0.0 and one with a negative score."""
    from oracle import decode as odec
    cs, ci = odec.part_candidates(heat, thr)
    h, w = heat.shape[1:]
    offs = off.reshape(2, -1, h, w).transpose(1, 2, 3, 0)
    roots = []
    for s_, (k, y, x) in list(zip(cs, ci))[:n_roots]:
        roots.append((np.float32(s_), int(k), np.array([y, x]) * stride + offs[k, y, x]))
    if roots:
        roots.append((np.float32(0.0), roots[0][1], roots[0][2]))
        roots.append((np.float32(-0.25), roots[-1][1], roots[-1][2]))
    return roots


def make_decode_pose():
    out = {"n": np.array(len(POSE_CASES))}
    for ci_, (di, n_roots) in enumerate(POSE_CASES):
        kind, h, w, stride, people, seed, P, thr, rad, minp, patch, extra = DEC_CASES[di]
        heat, off, fwd, bwd = _heads(kind, h, w, stride, people, seed, extra)
        split = lambda a: a.reshape(2, -1, h, w).transpose((1, 2, 3, 0))            # decode_multi.py:89-97
        offs, fwd_t, bwd_t = split(off), split(fwd), split(bwd)
        roots = pose_roots(heat, off, stride, thr, n_roots)
        ks, kc, ko, tr = [], [], [], []
        for rs, rid, rxy in roots:
            a, b, c = ref_dec.decode_pose(rs, rid, rxy, heat, offs, stride, fwd_t, bwd_t)
            ks.append(a); kc.append(b); ko.append(c)
            # one hop along every edge in both directions from this root's coordinates (decode.py:9-63)
            for e, (parent, child) in enumerate(ref.PARENT_CHILD_TUPLES):
                for tgt, disp in ((child, fwd_t), (parent, bwd_t)):
                    sc, xy, dv, ov = ref_dec.traverse_to_targ_keypoint(e, rxy, tgt, heat, offs, stride, disp)
                    assert sc.dtype == np.float32 and xy.dtype == np.float64 and dv.dtype == np.float32 and ov.dtype == np.float32
                    tr.append(np.concatenate([[sc], xy, dv, ov]).astype(np.float64))
        out["in_sha_%d" % ci_] = np.array(sha(np.concatenate([heat.ravel(), off.ravel(), fwd.ravel(), bwd.ravel()])))
        out["roots_%d" % ci_] = np.array([[r[0], r[1], r[2][0], r[2][1]] for r in roots], dtype=np.float64).reshape(-1, 4)
        out["ks_%d" % ci_] = np.array(ks).reshape(-1, 17)
        out["kc_%d" % ci_] = np.array(kc).reshape(-1, 17, 2)
        out["ko_%d" % ci_] = np.array(ko).reshape(-1, 17, 2)
        out["tr_%d" % ci_] = np.array(tr).reshape(-1, 7)
        print("decode_pose case %d (decode case %d): %d roots, %d hops" % (ci_, di, len(roots), len(tr)))
    np.savez_compressed(os.path.join(HERE, "decode_pose.npz"), **out)


def make_converter_names():
    """The reference's TF.js -> torch variable-name mapping (converter/tfjs2pytorch.py:15-43) on every variable name of a
    PoseNet MobileNetV1 manifest (14 conv layers, 4 heads, plus names the converter must skip) -> converter_names.json.
    Also runs the reference's load_variables on a synthetic checkpoint and records a checksum of every converted tensor
    (the layout transposes, tfjs2pytorch.py:46-72) -> converter_tensors.json."""
    import json
    import tempfile
    from posenet.converter import tfjs2pytorch as ref_conv
    names = ["MobilenetV1/Conv2d_0/weights", "MobilenetV1/Conv2d_0/biases"]
    for i in range(1, 14):
        names += ["MobilenetV1/Conv2d_%d_depthwise/depthwise_weights" % i, "MobilenetV1/Conv2d_%d_depthwise/biases" % i,
                  "MobilenetV1/Conv2d_%d_pointwise/weights" % i, "MobilenetV1/Conv2d_%d_pointwise/biases" % i]
    for head in ("heatmap", "offset", "displacement_fwd", "displacement_bwd"):
        names += ["MobilenetV1/%s_2/weights" % head, "MobilenetV1/%s_2/biases" % head]
    names += ["MobilenetV1/heatmap_1/weights", "MobilenetV1/segment_2/weights", "MobilenetV1/partheat_2/biases",
              "MobilenetV1/Conv2d_3_pointwise/Relu6", "MobilenetV1/displacement_fwd_1/biases", "MobilenetV1/Logits/weights"]
    table = {n: ref_conv.to_torch_name(n) for n in names}
    with open(os.path.join(HERE, "converter_names.json"), "w") as f:
        json.dump(table, f, indent=0, sort_keys=True)
    # synthetic checkpoint for model 50: values = a seeded ramp per variable, written in TF layouts
    rng = np.random.default_rng(7)
    sd_shapes = {k: tuple(v.shape) for k, v in ref.MobileNetV1(50).state_dict().items()}
    with tempfile.TemporaryDirectory() as tmp:
        ck = os.path.join(tmp, "mobilenet_v1_050")
        os.makedirs(ck)
        manifest, inv = {}, {v: k for k, v in table.items() if v}
        for key, shp in sd_shapes.items():
            tfn = inv[key]
            if len(shp) == 4:
                tf_shape = [shp[2], shp[3], shp[0], shp[1]] if "depthwise" in tfn else [shp[2], shp[3], shp[1], shp[0]]
            else:
                tf_shape = list(shp)
            fn = tfn.replace("/", "_")
            rng.standard_normal(tf_shape).astype("<f4").tofile(os.path.join(ck, fn))
            manifest[tfn] = {"filename": fn, "shape": tf_shape}
        json.dump(manifest, open(os.path.join(ck, "manifest.json"), "w"))
        out = ref_conv.load_variables("mobilenet_v1_050", tmp)
    sums = {k: [list(v.shape), sha(v.numpy())] for k, v in out.items()}
    with open(os.path.join(HERE, "converter_tensors.json"), "w") as f:
        json.dump(sums, f, indent=0, sort_keys=True)
    print("converter: %d names, %d tensors" % (len(table), len(sums)))


if __name__ == "__main__":
    torch.manual_seed(0)
    if "--converter-only" in sys.argv:
        make_converter_names()
        sys.exit(0)
    make_preprocess()
    make_net()
    make_decode()
    make_decode_pose()
    make_converter_names()
