"""Oracle for part-candidate extraction and greedy multi-pose decoding (TEST INFRASTRUCTURE).

numpy float64 restatement of ``posenet/decode_multi.py`` and ``posenet/decode.py`` of the
reference.  The one deliberate difference: the reference orders candidates with an
*unstable* ``torch.argsort(descending=True)`` (decode_multi.py:33), so the order of tied
scores is implementation-defined there.  This oracle -- and the CUDA decoder it
checks -- define the order as (score descending, flat index (part, y, x) ascending).
On tie-free inputs the two agree exactly; ``tests/golden/decode_*.npz`` pins that.
"""
import numpy as np

PARTS = 17
# (parent, child) per edge -- constants.py:25-36 resolved through PART_IDS (constants.py:10)
EDGES = ((0, 1), (1, 3), (0, 2), (2, 4), (0, 5), (5, 7), (7, 9), (5, 11), (11, 13), (13, 15),
         (0, 6), (6, 8), (8, 10), (6, 12), (12, 14), (14, 16))


def part_candidates(scores, score_threshold):
    """decode_multi.py:27-34.  ``scores`` f32 [17,h,w].  3x3 window, stride 1, -inf padding;
    a cell qualifies when it equals its window max (plateaus all qualify) and is
    >= fp32(threshold).  Returns (scores f32[n], idx int64[n,3]) ordered (score desc, flat idx asc)."""
    s = np.asarray(scores, dtype=np.float32)
    k, h, w = s.shape
    pad = np.full((k, h + 2, w + 2), -np.inf, dtype=np.float32)
    pad[:, 1:-1, 1:-1] = s
    m = s.copy()
    for dy in range(3):
        for dx in range(3):
            np.maximum(m, pad[:, dy:dy + h, dx:dx + w], out=m)
    keep = (s == m) & (s >= np.float32(score_threshold))
    idx = np.argwhere(keep)                       # row-major (part, y, x), like torch.nonzero
    vals = s[keep]
    order = np.argsort(-vals.astype(np.float64), kind="stable")
    return vals[order], idx[order].astype(np.int64)


def _cell(p, stride, h, w):
    """decode.py:15-16 / :50-51 -- f64 divide, round-half-even, clip, int32."""
    return np.clip(np.round(p / stride), a_min=0, a_max=[h - 1, w - 1]).astype(np.int32)


def traverse_to_targ_keypoint(edge, src_xy, target, scores, offsets, stride, disp):
    """decode.py:9-63 (single step, no refinement).  ``offsets`` [17,h,w,2], ``disp`` [16,h,w,2] (decode_multi.py:89-97).
    Returns the reference's 4-tuple ``(score, image_coord, displacement_vector, offset)``."""
    h, w = scores.shape[1], scores.shape[2]
    si = _cell(src_xy, stride, h, w)
    d = disp[edge, si[0], si[1]]                  # f32[2] (dy, dx)
    p = src_xy + d                                # f64
    ti = _cell(p, stride, h, w)
    sc = scores[target, ti[0], ti[1]]
    off = offsets[target, ti[0], ti[1]]           # f32[2]
    return sc, ti * stride + off, d, off


def _hop(edge, src_xy, target, scores, offsets, stride, disp):
    sc, xy, _, off = traverse_to_targ_keypoint(edge, src_xy, target, scores, offsets, stride, disp)
    return sc, xy, off


def decode_pose(root_score, root_id, root_xy, scores, offsets, stride, fwd, bwd):
    """decode.py:131-182.  Backward edges 15..0 (child -> parent via ``bwd``) then forward
    edges 0..15 (parent -> child via ``fwd``); a keypoint is "decoded" iff its score != 0."""
    ks = np.zeros(PARTS)
    kc = np.zeros((PARTS, 2))
    ko = np.zeros((PARTS, 2))
    ks[root_id] = root_score
    kc[root_id] = root_xy
    for e in range(len(EDGES) - 1, -1, -1):
        tgt, src = EDGES[e]
        if ks[src] > 0.0 and ks[tgt] == 0.0:
            ks[tgt], kc[tgt], ko[tgt] = _hop(e, kc[src], tgt, scores, offsets, stride, bwd)
    for e in range(len(EDGES)):
        src, tgt = EDGES[e]
        if ks[src] > 0.0 and ks[tgt] == 0.0:
            ks[tgt], kc[tgt], ko[tgt] = _hop(e, kc[src], tgt, scores, offsets, stride, fwd)
    return ks, kc, ko


def decode_multiple_poses(scores, offsets, displacements_fwd, displacements_bwd, output_stride,
                          max_pose_detections=10, score_threshold=0.5, nms_radius=20,
                          min_pose_score=0.5):
    """decode_multi.py:61-148.  Inputs f32 arrays [17|34|32|32, h, w] of ONE image.
    Returns the 4-tuple (pose_scores[P], keypoint_scores[P,17], keypoint_coords[P,17,2],
    pose_offsets[P,17,2]) in float64, zero padded."""
    scores = np.asarray(scores, dtype=np.float32)
    h, w = scores.shape[1], scores.shape[2]
    cand_s, cand_i = part_candidates(scores, score_threshold)
    split = lambda a: np.asarray(a, dtype=np.float32).reshape(2, -1, h, w).transpose(1, 2, 3, 0)
    offs, fwd, bwd = split(offsets), split(displacements_fwd), split(displacements_bwd)  # :89-97

    r2 = nms_radius ** 2
    P = max_pose_detections
    n = 0
    pose_scores = np.zeros(P)
    pose_ks = np.zeros((P, PARTS))
    pose_kc = np.zeros((P, PARTS, 2))
    pose_ko = np.zeros((P, PARTS, 2))
    for rs, (rid, ry, rx) in zip(cand_s, cand_i):
        root = np.array([ry, rx]) * output_stride + offs[rid, ry, rx]          # :106-109 (f64)
        if n and np.any(np.sum((pose_kc[:n, rid, :] - root) ** 2, axis=1) <= r2):  # :8-11
            continue
        ks, kc, ko = decode_pose(rs, rid, root, scores, offs, output_stride, fwd, bwd)
        if n:                                                                   # :14-24
            far = np.sum((pose_kc[:n] - kc) ** 2, axis=2) > r2
            score = np.sum(ks[np.all(far, axis=0)]) / len(ks)
        else:
            score = np.sum(ks) / len(ks)
        if min_pose_score == 0. or score >= min_pose_score:                     # :128
            pose_scores[n] = score
            pose_ks[n] = ks
            pose_kc[n] = kc
            pose_ko[n] = ko
            n += 1
        if n >= P:                                                              # :138
            break
    return pose_scores, pose_ks, pose_kc, pose_ko
