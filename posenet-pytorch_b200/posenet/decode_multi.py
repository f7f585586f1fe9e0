"""Multi-pose decoding on the GPU, behind the reference's ``decode_multiple_poses`` signature.

Mirror of ``posenet/decode_multi.py:61-148`` of the reference: same arguments and defaults, same
4-tuple of writable float64 numpy arrays zero-padded to ``max_pose_detections``.  The candidate
search (:27-34), the six ``.cpu().numpy()`` round trips (:78-97) and the greedy Python loop
(:104-139, with posenet/decode.py) all run inside two CUDA kernels (csrc/decode.cu); the only host
transfer is the final read-back of the pose records.  ``decode_multiple_poses_batch`` is the
batched, device-resident form.
"""
import ctypes as C

import numpy as np
import torch

from posenet import _native as nat
from posenet.constants import *  # noqa: F401,F403  (the reference's module re-exports these)
from posenet.constants import NUM_KEYPOINTS


def _as_maps(t, channels, device):
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.asarray(t))
    if t.dim() == 3:
        t = t.unsqueeze(0)
    assert t.dim() == 4 and t.shape[1] == channels, "expected [%d,h,w] or [N,%d,h,w], got %s" % (
        channels, channels, tuple(t.shape))
    if t.dtype != torch.float32:
        t = t.float()
    return t.to(device) if t.device != device else t


def split_pose_records(flat, n, P):
    """Views of the packed record buffer ``flat`` (float64 [n*P*86], torch or numpy): pose_scores [n,P],
    keypoint_scores [n,P,17], keypoint_coords [n,P,17,2], pose_offsets [n,P,17,2]."""
    K = NUM_KEYPOINTS
    o1, o2, o3 = n * P, n * P * (1 + K), n * P * (1 + K * 3)
    return (flat[:o1].reshape(n, P), flat[o1:o2].reshape(n, P, K), flat[o2:o3].reshape(n, P, K, 2),
            flat[o3:].reshape(n, P, K, 2))


def decode_multiple_poses_batch(scores, offsets, displacements_fwd, displacements_bwd, output_stride,
                                max_pose_detections=10, score_threshold=0.5, nms_radius=20, min_pose_score=0.5,
                                workspace=None, out=None):
    """Decode N images at once, everything staying on the device (no synchronisation).

    Inputs are [N,17|34|32|32,h,w] fp32 CUDA tensors with any strides.  Returns
    ``(pose_scores [N,P], keypoint_scores [N,P,17], keypoint_coords [N,P,17,2], pose_offsets [N,P,17,2],
    pose_counts [N] int32)`` as CUDA tensors (float64 like the reference's numpy outputs); the first four
    are views of one packed buffer (``out``: optional preallocated float64 [N*P*86], fully overwritten).
    """
    nat.require_device()
    dev = scores.device if torch.is_tensor(scores) and scores.is_cuda else torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(dev):                               # buffers, launches and the stream belong to the tensors' GPU
        return _decode_batch_on(dev, scores, offsets, displacements_fwd, displacements_bwd, output_stride, max_pose_detections,
                                score_threshold, nms_radius, min_pose_score, workspace, out)


def _decode_batch_on(dev, scores, offsets, displacements_fwd, displacements_bwd, output_stride, max_pose_detections,
                     score_threshold, nms_radius, min_pose_score, workspace, out):
    lib = nat.load()
    heat = _as_maps(scores, 17, dev)
    off = _as_maps(offsets, 34, dev)
    fwd = _as_maps(displacements_fwd, 32, dev)
    bwd = _as_maps(displacements_bwd, 32, dev)
    n, _, h, w = heat.shape
    assert off.shape[0] == n and tuple(off.shape[2:]) == (h, w) and tuple(fwd.shape[2:]) == (h, w) and tuple(bwd.shape[2:]) == (h, w)
    P = int(max_pose_detections)
    cap = 17 * h * w                                   # every cell of every part can be a candidate
    if workspace is None:
        workspace = {}
    key = (n, h, w, P, dev)
    ws = workspace.get(key)
    if ws is None:
        ws = workspace[key] = dict(keys=torch.empty((n, cap), dtype=torch.int64, device=dev),
                                   counts=torch.empty(n, dtype=torch.int32, device=dev))
    if out is None:                                    # pn_decode_greedy zero-pads the rows it does not fill
        out = torch.empty(n * P * (1 + NUM_KEYPOINTS * 5), dtype=torch.float64, device=dev)
    else:
        assert out.dtype == torch.float64 and out.numel() == n * P * (1 + NUM_KEYPOINTS * 5) and out.is_contiguous()
    ps, ks, kc, ko = split_pose_records(out, n, P)
    pose_counts = torch.empty(n, dtype=torch.int32, device=dev)
    maps = [nat.make_map(t) for t in (heat, off, fwd, bwd)]
    prm = nat.DecodeParams(int(output_stride), P, float(nms_radius ** 2), float(min_pose_score))
    st = nat.stream_ptr()
    nat.check(lib.pn_candidates(C.byref(maps[0]), n, h, w, C.c_float(score_threshold), C.c_void_p(ws["keys"].data_ptr()),
                                cap, C.c_void_p(ws["counts"].data_ptr()), st), "pn_candidates")
    nat.check(lib.pn_decode_greedy(C.byref(maps[0]), C.byref(maps[1]), C.byref(maps[2]), C.byref(maps[3]), n, h, w,
                                   C.c_void_p(ws["keys"].data_ptr()), cap, C.c_void_p(ws["counts"].data_ptr()),
                                   C.byref(prm), C.c_void_p(ps.data_ptr()), C.c_void_p(ks.data_ptr()),
                                   C.c_void_p(kc.data_ptr()), C.c_void_p(ko.data_ptr()),
                                   C.c_void_p(pose_counts.data_ptr()), st), "pn_decode_greedy")
    return ps, ks, kc, ko, pose_counts


def decode_multiple_poses(
        scores, offsets, displacements_fwd, displacements_bwd, output_stride,
        max_pose_detections=10, score_threshold=0.5, nms_radius=20, min_pose_score=0.5):
    """One image ([17|34|32|32,h,w] tensors on any device) -> the reference's 4-tuple of numpy float64
    arrays ``(pose_scores, pose_keypoint_scores, pose_keypoint_coords, pose_offsets)``."""
    ps, ks, kc, ko, _ = decode_multiple_poses_batch(
        scores, offsets, displacements_fwd, displacements_bwd, output_stride,
        max_pose_detections=max_pose_detections, score_threshold=score_threshold,
        nms_radius=nms_radius, min_pose_score=min_pose_score)
    # one D2H copy of the packed records; the arrays are fresh and writable (image_demo.py scales coords in place)
    host = torch.cat([ps.reshape(-1), ks.reshape(-1), kc.reshape(-1), ko.reshape(-1)]).cpu().numpy()
    P, K = int(max_pose_detections), NUM_KEYPOINTS
    a, b, c = P, P * (1 + K), P * (1 + 3 * K)
    return (host[:a].copy(), host[a:b].reshape(P, K).copy(), host[b:c].reshape(P, K, 2).copy(),
            host[c:].reshape(P, K, 2).copy())
