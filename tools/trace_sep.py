"""Timeline of one fused-block CTA (debug build: PN_EXTRA_NVCC_FLAGS=-DPN_SEP_TRACE python posenet-pytorch_b200/build.py --force).
Prints, for block 0, per k-block: depthwise wait-for-A / wait-for-patch / compute cycles, MMA issue stamps, epilogue spans."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "posenet-pytorch_b200"), os.path.join(ROOT, "tests"), ROOT]
import torch
import abi
from posenet import _native as nat

def run(n, h, w, cin, cout, stride, dil, cap=256, show=24):
    lib = nat.load()
    lib.pn_debug_sep_trace.argtypes = [C.c_void_p, C.c_int]
    g = torch.Generator().manual_seed(0)
    x = (torch.rand((n, h, w, cin), generator=g) * 6).to(torch.bfloat16).cuda()
    w9 = (torch.randn((9, cin), generator=g) * 0.3).cuda()
    bd = torch.zeros(cin).cuda()
    wp = (torch.randn((cout, cin), generator=g) / cin ** 0.5).to(torch.bfloat16).cuda()
    bp = torch.zeros(cout).cuda()
    for _ in range(2):
        abi.sepconv(x, w9, bd, wp, bp, stride, dil)
    torch.cuda.synchronize()
    buf = torch.zeros((4, cap, 4), dtype=torch.int64, device="cuda")
    assert lib.pn_debug_sep_trace(C.c_void_p(buf.data_ptr()), cap) == 0
    abi.sepconv(x, w9, bd, wp, bp, stride, dil)
    torch.cuda.synchronize()
    lib.pn_debug_sep_trace(None, 0)
    t = buf.cpu().numpy()
    t0 = t[t > 0].min()
    dw, mma, epi, prod = t[0], t[1], t[2], t[3]
    desc = C.create_string_buffer(256)
    lib.pn_sepconv_describe(n, h, w, cin, cout, stride, dil, desc, 256)
    print("== %s: %s" % ((n, h, w, cin, cout, stride, dil), desc.value.decode()))
    print("k-block item: dw warp6 start | wait_A wait_patch work || producer issue@ | mma a_full@ issued@ (tempty@)")
    for i in range(show):
        if dw[i, 3] == 0: break
        m = mma[i]
        print("%3d %8d | %6d %6d %6d || prod@%8d | mma a_full@%8d commit@%8d tempty@%8d" % (
            i, dw[i, 0] - t0, dw[i, 1] - dw[i, 0], dw[i, 2] - dw[i, 1], dw[i, 3] - dw[i, 2], prod[i, 0] - t0 if prod[i, 0] else -1,
            m[1] - t0 if m[1] else -1, m[2] - t0 if m[2] else -1, m[0] - t0 if m[0] else -1))
    for i in range(min(show, cap)):
        if epi[i, 1] == 0: break
        print("epilogue tile %d: start@%8d  dur %6d" % (i, epi[i, 0] - t0, epi[i, 1] - epi[i, 0]))
    n_it = int((dw[:, 3] > 0).sum())
    if n_it > 8:
        print("steady state: %.0f cycles per k-block item over items 4..%d" % ((dw[n_it - 1, 0] - dw[4, 0]) / (n_it - 5), n_it - 1))


if __name__ == "__main__":
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(64, 33, 33, 512, 512, 1, 1)]
    for shp in shapes:
        run(*shp, show=40)
