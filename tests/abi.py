"""Thin torch-tensor wrappers over the C ABI for the GPU parity tests (every call goes through
libposenet_b200.so via posenet._native -- nothing here computes)."""
import ctypes as C

import torch

from posenet import _native as nat

P = lambda t: C.c_void_p(t.data_ptr())

_diag = None


def load_diag():
    """libposenet_b200_diag.so (include/posenet_b200_diag.h): hardware probes, not part of the product library."""
    global _diag
    if _diag is None:
        import os
        lib = C.CDLL(os.path.join(os.path.dirname(nat.LIB_PATH), "libposenet_b200_diag.so"))
        lib.pn_diag_last_error_string.restype = C.c_char_p
        lib.pn_dwtc_probe.restype = C.c_int
        lib.pn_dwtc_probe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.pn_debug_umma_cost.restype = C.c_int
        lib.pn_debug_umma_cost.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
        _diag = lib
    return _diag


def check_diag(rc, what):
    if rc != 0:
        raise nat.NativeError("%s failed (%d): %s" % (what, rc, (load_diag().pn_diag_last_error_string() or b"?").decode()))

TORCH_DT = {nat.PN_F32: torch.float32, nat.PN_BF16: torch.bfloat16}


def conv_out(n, stride, dil):
    pad = ((stride - 1) + 2 * dil) // 2
    return (n + 2 * pad - 2 * dil - 1) // stride + 1


def preprocess(img_u8, th, tw):
    n, h, w, _ = img_u8.shape
    out = torch.empty((n, 3, th, tw), dtype=torch.float32, device=img_u8.device)
    nat.check(nat.load().pn_preprocess_u8(P(img_u8), n, h, w, th, tw, P(out), nat.stream_ptr()), "pn_preprocess_u8")
    return out


def resize_u8(img_u8, th, tw):
    n, h, w, _ = img_u8.shape
    out = torch.empty((n, th, tw, 3), dtype=torch.uint8, device=img_u8.device)
    nat.check(nat.load().pn_resize_u8(P(img_u8), n, h, w, th, tw, P(out), nat.stream_ptr()), "pn_resize_u8")
    return out


def stem(x, w27, b, stride, dtype, u8=False):
    lib = nat.load()
    if u8:
        n, h, wd, _ = x.shape
    else:
        n, _, h, wd = x.shape
    cout = w27.shape[1]
    y = torch.empty((n, conv_out(h, stride, 1), conv_out(wd, stride, 1), cout), dtype=TORCH_DT[dtype], device=x.device)
    fn = lib.pn_stem_conv_u8 if u8 else lib.pn_stem_conv
    nat.check(fn(P(x), P(w27), P(b), P(y), n, h, wd, cout, stride, dtype, nat.stream_ptr()), "pn_stem_conv")
    return y


def dwconv(x, w9, b, stride, dil, dtype):
    n, h, wd, c = x.shape
    y = torch.empty((n, conv_out(h, stride, dil), conv_out(wd, stride, dil), c), dtype=TORCH_DT[dtype], device=x.device)
    nat.check(nat.load().pn_dwconv3x3(P(x), P(w9), P(b), P(y), n, h, wd, c, stride, dil, dtype, nat.stream_ptr()),
              "pn_dwconv3x3")
    return y


def pwconv(a, w, b, dtype):
    m, k = a.shape
    n = w.shape[0]
    y = torch.empty((m, n), dtype=TORCH_DT[dtype], device=a.device)
    nat.check(nat.load().pn_pwconv_gemm(P(a), P(w), P(b), P(y), m, k, n, dtype, nat.stream_ptr()), "pn_pwconv_gemm")
    return y


def sepconv(x, w9, b_dw, w_pw, b_pw, stride, dil):
    """Fused SeperableConv block (bf16): x NHWC bf16, w9 f32 [9,cin], w_pw bf16 [cout,cin]."""
    n, h, wd, c = x.shape
    cout = w_pw.shape[0]
    y = torch.full((n, conv_out(h, stride, dil), conv_out(wd, stride, dil), cout), float("nan"), dtype=torch.bfloat16, device=x.device)
    nat.check(nat.load().pn_sepconv_block(P(x), P(w9), P(b_dw), P(w_pw), P(b_pw), P(y), n, h, wd, c, cout, stride, dil,
                                          nat.stream_ptr()), "pn_sepconv_block")
    return y


def heads(a, w128, b128, n_img, hw, dtype):
    k = a.shape[1]
    outs = [torch.full((n_img, ch, hw), float("nan"), dtype=torch.float32, device=a.device) for ch in (17, 34, 32, 32)]
    nat.check(nat.load().pn_heads_gemm(P(a), P(w128), P(b128), *[P(o) for o in outs], n_img, hw, k, dtype,
                                       nat.stream_ptr()), "pn_heads_gemm")
    return outs


def candidates(heat4, thr):
    """heat4: [n,17,h,w] f32 cuda (any strides) -> list of sorted uint64 key arrays (numpy), counts."""
    n, _, h, w = heat4.shape
    cap = 17 * h * w
    keys = torch.zeros((n, cap), dtype=torch.int64, device=heat4.device)
    counts = torch.zeros(n, dtype=torch.int32, device=heat4.device)
    m = nat.make_map(heat4)
    nat.check(nat.load().pn_candidates(C.byref(m), n, h, w, C.c_float(thr), P(keys), cap, P(counts), nat.stream_ptr()),
              "pn_candidates")
    return keys, counts
