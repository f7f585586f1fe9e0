"""Oracle for the preprocessing stage (TEST INFRASTRUCTURE).

Restates ``posenet/utils.py:7-26`` of the reference (``valid_resolution`` and
``_process_input``).  The resize arithmetic is OpenCV's ``cv2.resize(...,
INTER_LINEAR)`` on uint8, which is third-party code that is not vendored in the
reference (README.md:23 names opencv-python 4.6.0; the build container has
4.13.0).  Its published fixed-point algorithm is restated here in numpy and is
pinned bit-exactly against cv2 by ``tests/golden/preprocess_*.npz``.
"""
import numpy as np

INTER_BITS = 11                      # cv2 INTER_RESIZE_COEF_BITS
INTER_SCALE = 1 << INTER_BITS        # 2048


def valid_resolution(width, height, output_stride=16):
    """utils.py:7-10 -- note the (width, height) return order."""
    tw = (int(width) // output_stride) * output_stride + 1
    th = (int(height) // output_stride) * output_stride + 1
    return tw, th


def _axis_table(dst, src, clamp_weights):
    """Source tap index and the two 11-bit integer weights for every output coordinate.

    cv2: ``scale = 1.0 / (dst / src)`` in f64; ``f = (float)((d + 0.5) * scale - 0.5)``;
    ``s = floor(f)``; ``f -= s`` in fp32.  On the x axis out-of-range taps reset the
    weight to 0 (``clamp_weights``); on the y axis only the row index is clipped.
    Weights are ``saturate_cast<short>(w * 2048)`` with round-half-even.
    """
    scale = 1.0 / (np.float64(dst) / np.float64(src))
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_weights:
        lo = s < 0
        s = np.where(lo, 0, s)
        f = np.where(lo, np.float32(0), f)
        hi = s >= src - 1
        s = np.where(hi, src - 1, s)
        f = np.where(hi, np.float32(0), f)
        s0 = s
        s1 = np.minimum(s + 1, src - 1)
    else:
        s0 = np.clip(s, 0, src - 1)
        s1 = np.clip(s + 1, 0, src - 1)
    f = f.astype(np.float32)
    w1 = np.rint(f * np.float32(INTER_SCALE)).astype(np.int32)
    w0 = np.rint((np.float32(1.0) - f) * np.float32(INTER_SCALE)).astype(np.int32)
    return s0, s1, w0, w1


def resize_linear_u8(src, tw, th):
    """``cv2.resize(src, (tw, th), interpolation=cv2.INTER_LINEAR)`` for uint8 HWC."""
    h, w = src.shape[:2]
    if (th, tw) == (h, w):
        return src.copy()
    if h == 2 * th and w == 2 * tw:
        # cv2 routes exact 2x decimation through the INTER_AREA 2x2 box path
        a = src.astype(np.int32)
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    x0, x1, a0, a1 = _axis_table(tw, w, True)
    y0, y1, b0, b1 = _axis_table(th, h, False)
    s = src.astype(np.int32)
    # horizontal pass: int32 rows, scale 2^11
    hrow = s[:, x0, :] * a0[None, :, None] + s[:, x1, :] * a1[None, :, None]
    r0 = hrow[y0]
    r1 = hrow[y1]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def process_input(source_img, scale_factor=1.0, output_stride=16):
    """utils.py:13-26 -> (input f32 [1,3,th,tw], source_img, scale f64[2])."""
    tw, th = valid_resolution(source_img.shape[1] * scale_factor,
                              source_img.shape[0] * scale_factor, output_stride)
    scale = np.array([source_img.shape[0] / th, source_img.shape[1] / tw])
    img = resize_linear_u8(source_img, tw, th)
    rgb = img[:, :, ::-1].astype(np.float32)                      # utils.py:22 BGR -> RGB
    x = rgb * np.float32(2.0 / 255.0) - np.float32(1.0)           # utils.py:23: two rounded fp32 ops
    x = np.ascontiguousarray(x.transpose(2, 0, 1)).reshape(1, 3, th, tw)
    return x, source_img, scale
