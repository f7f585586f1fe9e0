"""pn_sepconv_block on the tensor-pipe depthwise path (septc.cu) vs a torch fp32 reference with the kernel's rounding points
(bf16 input, bf16 depthwise weights, bf16 depthwise output, bf16 output), plus timings at bench batch sizes.
    python tools/check_septc.py [--time]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "posenet-pytorch_b200"), os.path.join(ROOT, "tests"), ROOT]
import torch

os.environ.setdefault("PN_SEP_TC", "1")          # the path under test is opt-in
import torch.nn.functional as F
import abi
from posenet import _native as nat

DEV = "cuda"
SHAPES = [(2, 5, 3, 64, 64, 1, 1), (1, 12, 33, 64, 64, 1, 1), (3, 33, 33, 512, 512, 1, 1), (2, 65, 65, 256, 256, 1, 1), (2, 129, 129, 128, 128, 1, 1),
          (1, 46, 81, 64, 64, 1, 1), (1, 46, 81, 128, 256, 1, 1), (1, 46, 81, 256, 256, 1, 2), (5, 17, 17, 384, 384, 1, 1),
          (1, 91, 161, 256, 256, 1, 2), (2, 33, 33, 192, 192, 1, 1), (1, 1, 1, 64, 128, 1, 1)]


def make(n, h, w, cin, cout, seed):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand((n, h, w, cin), generator=g) * 6).to(torch.bfloat16)
    wd = torch.randn((cin, 1, 3, 3), generator=g) * 0.35
    bd = torch.randn(cin, generator=g) * 0.3
    wp = (torch.randn((cout, cin), generator=g) * (1.5 / cin ** 0.5)).to(torch.bfloat16)
    bp = torch.randn(cout, generator=g) * 0.5
    return x, wd, bd, wp, bp


def check(shape):
    n, h, w, cin, cout, stride, dil = shape
    x, wd, bd, wp, bp = make(n, h, w, cin, cout, seed=h * 7 + cin + stride + dil)
    buf = C.create_string_buffer(512)
    nat.load().pn_sepconv_describe(n, h, w, cin, cout, stride, dil, buf, 512)
    wdr = wd.to(torch.bfloat16).float()
    t = F.relu6(F.conv2d(x.float().permute(0, 3, 1, 2), wdr, bd, stride=stride, padding=dil, dilation=dil, groups=cin))
    t = t.to(torch.bfloat16).float()
    ref = F.relu6(F.conv2d(t, wp.float().reshape(cout, cin, 1, 1), bp)).permute(0, 2, 3, 1)
    w9 = wd.reshape(cin, 9).t().contiguous().to(DEV)
    xd, bdd, wpd, bpd = x.to(DEV), bd.to(DEV), wp.to(DEV), bp.to(DEV)
    y = abi.sepconv(xd, w9, bdd, wpd, bpd, stride, dil)
    torch.cuda.synchronize()
    yf = y.float().cpu()
    nan = int(torch.isnan(yf).sum())
    err = float((yf - ref).abs().max() / ref.abs().max().clamp_min(1e-6))
    frac = float(((yf - ref).abs() > 0.07).float().mean())
    y2 = abi.sepconv(xd, w9, bdd, wpd, bpd, stride, dil)
    print("%-34s err %.4g  cells off by > 1 ulp %.2g  nan %d  deterministic %s | %s" % (shape, err, frac, nan, torch.equal(y, y2), buf.value.decode()), flush=True)
    return err < 1.2e-2 and nan == 0


def timeit(shape, reps=20):
    n, h, w, cin, cout, stride, dil = shape
    x, wd, bd, wp, bp = make(n, h, w, cin, cout, seed=1)
    w9 = wd.reshape(cin, 9).t().contiguous().to(DEV)
    xs = [x.to(DEV).clone() for _ in range(3)]
    bdd, wpd, bpd = bd.to(DEV), wp.to(DEV), bp.to(DEV)
    for i in range(3):
        abi.sepconv(xs[i % 3], w9, bdd, wpd, bpd, stride, dil)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        abi.sepconv(xs[i % 3], w9, bdd, wpd, bpd, stride, dil)
    e1.record()
    torch.cuda.synchronize()
    print("time %-34s %.1f us  (PN_SEP_TC=%s)" % (shape, e0.elapsed_time(e1) / reps * 1e3, os.environ.get("PN_SEP_TC", "1")), flush=True)


torch.cuda.set_device(0)
if "--time" in sys.argv:
    for shp in [(64, 129, 129, 128, 128, 1, 1), (64, 65, 65, 256, 256, 1, 1), (64, 33, 33, 512, 512, 1, 1), (32, 91, 161, 256, 256, 1, 2),
                (512, 17, 17, 384, 384, 1, 1)]:
        timeit(shp)
else:
    ok = all([check(s) for s in SHAPES])
    print("ALL OK" if ok else "FAILURES")
