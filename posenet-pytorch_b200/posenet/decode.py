"""``posenet.decode`` -- the single-pose steps of the decoder, callable on their own like the reference's.

Mirror of ``posenet/decode.py`` of the reference: ``traverse_to_targ_keypoint`` (:9-63) and ``decode_pose`` (:131-182) with
the reference's positional arguments, array layouts and return values.  Both run on the GPU through the same device code the
fused decoder uses (``csrc/decode.cu``: ``hop_full`` and the per-part path walk of ``decode_kernel``'s speculation phase), one
hop / one root per call, so they are bit-identical with what ``decode_multiple_poses`` computes internally.  There is no
host implementation: without the library or a CUDA device they raise.

Layouts (what ``decode_multi.py:78-97`` hands to these functions): ``scores`` f32 ``[17,h,w]``; ``offsets`` f32
``[17,h,w,2]`` (last axis y, x); ``displacements*`` f32 ``[16,h,w,2]``.  The network's planar ``[34,h,w]`` / ``[32,h,w]``
tensors are accepted as well.  The helpers the reference also keeps in that file (``find_root``,
``print_decoded_heatmap``, ``build_part_with_score_torch_single_pose``) are dead code there and out of scope.
"""
import ctypes as C

import numpy as np
import torch

from posenet import _native as nat
from posenet.constants import *  # noqa: F401,F403  (the reference's module re-exports these)
from posenet.constants import NUM_KEYPOINTS, PARENT_CHILD_TUPLES


def _device_of(*tensors):
    for t in tensors:
        if torch.is_tensor(t) and t.is_cuda:
            return t.device
    return torch.device("cuda", torch.cuda.current_device())


def _planar(t, channels, dev):
    """f32 CUDA tensor [1, 2*channels | channels, h, w] from ``[channels,h,w,2]`` (the reference's transposed layout),
    ``[2*channels,h,w]`` (the network's) or, for the score map, ``[channels,h,w]``."""
    if not torch.is_tensor(t):
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(t, dtype=np.float32)))
    t = t.to(dev, dtype=torch.float32)
    if t.dim() == 4 and t.shape[3] == 2:
        assert t.shape[0] == channels, "expected [%d,h,w,2], got %s" % (channels, tuple(t.shape))
        t = t.permute(3, 0, 1, 2).reshape(2 * channels, t.shape[1], t.shape[2])
    assert t.dim() == 3, "expected a [C,h,w] or [C,h,w,2] array, got %s" % (tuple(t.shape),)
    return t.contiguous().unsqueeze(0)


def traverse_to_targ_keypoint(edge_id, source_keypoint, target_keypoint_id, scores, offsets, output_stride, displacements):
    """decode.py:9-63.  Returns ``(score f32, image_coord f64[2], displacement_vector f32[2], offset f32[2])``."""
    nat.require_device()
    dev = _device_of(scores, offsets, displacements)
    with torch.cuda.device(dev):
        heat, off, disp = _planar(scores, NUM_KEYPOINTS, dev), _planar(offsets, NUM_KEYPOINTS, dev), \
            _planar(displacements, len(PARENT_CHILD_TUPLES), dev)
        assert heat.shape[1] == NUM_KEYPOINTS and off.shape[1] == 2 * NUM_KEYPOINTS and disp.shape[1] == 2 * len(PARENT_CHILD_TUPLES)
        h, w = heat.shape[2], heat.shape[3]
        src = (C.c_double * 2)(*[float(v) for v in np.asarray(source_keypoint, dtype=np.float64).reshape(2)])
        out = torch.empty(7, dtype=torch.float64, device=dev)
        maps = [nat.make_map(t) for t in (heat, off, disp)]
        nat.check(nat.load().pn_traverse_to_targ_keypoint(int(edge_id), src, int(target_keypoint_id), C.byref(maps[0]),
                                                          C.byref(maps[1]), C.byref(maps[2]), h, w, int(output_stride),
                                                          C.c_void_p(out.data_ptr()), nat.stream_ptr()),
                  "pn_traverse_to_targ_keypoint")
        o = out.cpu().numpy()
    return np.float32(o[0]), o[1:3].copy(), o[3:5].astype(np.float32), o[5:7].astype(np.float32)


def decode_pose(root_score, root_id, root_image_coord, scores, offsets, output_stride, displacements_fwd, displacements_bwd):
    """decode.py:131-182.  Returns ``(instance_keypoint_scores f64[17], instance_keypoint_coords f64[17,2],
    instance_offsets f64[17,2])`` -- the third value is what this fork added (``pose_offsets`` of ``decode_multiple_poses``)."""
    nat.require_device()
    dev = _device_of(scores, offsets, displacements_fwd, displacements_bwd)
    with torch.cuda.device(dev):
        edges = len(PARENT_CHILD_TUPLES)
        heat, off = _planar(scores, NUM_KEYPOINTS, dev), _planar(offsets, NUM_KEYPOINTS, dev)
        fwd, bwd = _planar(displacements_fwd, edges, dev), _planar(displacements_bwd, edges, dev)
        assert heat.shape[1] == NUM_KEYPOINTS and off.shape[1] == 2 * NUM_KEYPOINTS and fwd.shape[1] == 2 * edges and bwd.shape[1] == 2 * edges
        h, w = heat.shape[2], heat.shape[3]
        root = (C.c_double * 2)(*[float(v) for v in np.asarray(root_image_coord, dtype=np.float64).reshape(2)])
        K = NUM_KEYPOINTS
        out = torch.empty(5 * K, dtype=torch.float64, device=dev)
        maps = [nat.make_map(t) for t in (heat, off, fwd, bwd)]
        nat.check(nat.load().pn_decode_pose(float(root_score), int(root_id), root, C.byref(maps[0]), C.byref(maps[1]),
                                            C.byref(maps[2]), C.byref(maps[3]), h, w, int(output_stride),
                                            C.c_void_p(out.data_ptr()), nat.stream_ptr()), "pn_decode_pose")
        o = out.cpu().numpy()
    return o[:K].copy(), o[K:3 * K].reshape(K, 2).copy(), o[3 * K:].reshape(K, 2).copy()
