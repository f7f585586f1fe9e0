"""Seeded synthetic inputs shared by tests and bench (TEST INFRASTRUCTURE).

There is no network for datasets or TF.js checkpoints, so every input is synthetic:
uint8 noise / smooth images for the backbone, and hand-built multi-person head tensors for
the decoder (peaky heatmaps + consistent offsets / displacements), which exercise NMS,
multi-cell traversal and ``min_pose_score`` rejection the way trained weights would.
"""
import numpy as np

from .decode import EDGES, PARTS

# A crude upright 17-part skeleton in unit-height coordinates (y down, x right).
_TEMPLATE = np.array([
    (0.08, 0.50), (0.06, 0.47), (0.06, 0.53), (0.07, 0.43), (0.07, 0.57),      # nose eyes ears
    (0.22, 0.38), (0.22, 0.62), (0.38, 0.33), (0.38, 0.67), (0.52, 0.30), (0.52, 0.70),  # shoulders elbows wrists
    (0.55, 0.42), (0.55, 0.58), (0.75, 0.41), (0.75, 0.59), (0.95, 0.40), (0.95, 0.60),  # hips knees ankles
], dtype=np.float64)


def noise_image(h, w, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def smooth_image(h, w, seed=0):
    """Low-frequency colour gradients + blobs + mild pixel noise (numerically gentler)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 3))
    for c in range(3):
        fy, fx, ph = rng.uniform(0.5, 3.0), rng.uniform(0.5, 3.0), rng.uniform(0, 6.28)
        img[:, :, c] = 128 + 90 * np.sin(fy * yy / h * 6.28 + ph) * np.cos(fx * xx / w * 6.28)
    for _ in range(6):
        cy, cx, r = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(0.05, 0.2) * min(h, w)
        amp = rng.uniform(-80, 80, 3)
        img += amp[None, None, :] * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r))[:, :, None]
    img += rng.normal(0, 4, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def people_heads(h, w, stride, n_people, seed=0, noise=1e-3):
    """Head tensors (f32: heat [17,h,w], offsets [34,h,w], fwd [32,h,w], bwd [32,h,w]) for
    ``n_people`` template skeletons, plus the ground-truth keypoints [n,17,2] (y, x) px."""
    rng = np.random.default_rng(seed)
    H, W = (h - 1) * stride + 1, (w - 1) * stride + 1
    kps = np.zeros((n_people, PARTS, 2))
    for p in range(n_people):
        size = rng.uniform(0.35, 0.60) * H
        top = rng.uniform(0, max(H - size, 1.0))
        left = rng.uniform(-0.2 * size, W - 0.8 * size)
        kps[p, :, 0] = top + _TEMPLATE[:, 0] * size
        kps[p, :, 1] = left + _TEMPLATE[:, 1] * size * 0.8
    kps[..., 0] = np.clip(kps[..., 0], 0, H - 1)
    kps[..., 1] = np.clip(kps[..., 1], 0, W - 1)
    amp = rng.uniform(0.7, 0.99, (n_people, PARTS))
    gy = (np.arange(h) * stride)[:, None]
    gx = (np.arange(w) * stride)[None, :]
    heat = np.zeros((PARTS, h, w))
    offs = np.zeros((2 * PARTS, h, w))
    nearest = np.zeros((PARTS, h, w), dtype=np.int64)
    for k in range(PARTS):
        d2 = (gy[None] - kps[:, k, 0, None, None]) ** 2 + (gx[None] - kps[:, k, 1, None, None]) ** 2
        heat[k] = (amp[:, k, None, None] * np.exp(-d2 / (2.0 * stride * stride))).max(axis=0)
        nearest[k] = d2.argmin(axis=0)
        offs[k] = kps[nearest[k], k, 0] - gy
        offs[PARTS + k] = kps[nearest[k], k, 1] - gx
    ne = len(EDGES)
    fwd = np.zeros((2 * ne, h, w))
    bwd = np.zeros((2 * ne, h, w))
    for e, (par, chi) in enumerate(EDGES):
        who = nearest[par]                       # person owning the parent's cell
        fwd[e] = kps[who, chi, 0] - kps[who, par, 0]
        fwd[ne + e] = kps[who, chi, 1] - kps[who, par, 1]
        who = nearest[chi]
        bwd[e] = kps[who, par, 0] - kps[who, chi, 0]
        bwd[ne + e] = kps[who, par, 1] - kps[who, chi, 1]
    heat = heat + rng.uniform(0, noise, heat.shape)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return f32(heat), f32(offs), f32(fwd), f32(bwd), kps


def random_heads(h, w, seed=0, disp_scale=40.0, off_scale=12.0, zero_frac=0.0, tie_levels=0):
    """Unstructured head tensors: uniform heat in (0,1), wide displacements -> traversals that
    hit borders.  ``zero_frac`` plants exact 0.0 scores (the "undecoded" flag, Appendix A.7);
    ``tie_levels`` > 0 quantises the heat to that many levels (plateaus / ties, F5-F6)."""
    rng = np.random.default_rng(seed)
    heat = rng.uniform(0, 1, (PARTS, h, w))
    if tie_levels:
        heat = np.floor(heat * tie_levels) / tie_levels
    if zero_frac:
        heat[rng.uniform(0, 1, heat.shape) < zero_frac] = 0.0
    offs = rng.normal(0, off_scale, (2 * PARTS, h, w))
    fwd = rng.normal(0, disp_scale, (2 * len(EDGES), h, w))
    bwd = rng.normal(0, disp_scale, (2 * len(EDGES), h, w))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return f32(heat), f32(offs), f32(fwd), f32(bwd)
