"""pn_stem_conv_u8 device time at the three bench geometries, per kernel variant (env switches of csrc/stem.cu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "posenet-pytorch_b200"), os.path.join(ROOT, "tests")]
import torch
import abi
from posenet import _native as nat
torch.cuda.set_device(0)
GEO = {"c2": (64, 513, 513, 32), "c3": (32, 721, 1281, 16), "c4": (512, 257, 257, 24)}
VARIANTS = [("default", {}), ("rows nbuf2", {"PN_STEM_SEGMENTS": "0", "PN_STEM_NBUF": "2"}), ("rows nbuf3", {"PN_STEM_SEGMENTS": "0", "PN_STEM_NBUF": "3"}),
            ("segments nbuf3", {"PN_STEM_SEGMENTS": "1", "PN_STEM_NBUF": "3"}), ("segments nbuf4", {"PN_STEM_SEGMENTS": "1", "PN_STEM_NBUF": "4"})]
for name, (n, h, w, cout) in GEO.items():
    xs = [torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda") for _ in range(4)]   # rotate inputs: > L2 with the outputs
    w27 = torch.randn(27, cout, device="cuda") * 0.25
    b = torch.randn(cout, device="cuda") * 0.3
    ref = None
    for vname, env in VARIANTS:
        os.environ.update(env)
        try:
            for i in range(3): y = abi.stem(xs[i % 4], w27, b, 2, nat.PN_BF16, u8=True)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(8): y = abi.stem(xs[i % 4], w27, b, 2, nat.PN_BF16, u8=True)
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(5): g.replay()
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1000 / 40
            y0 = abi.stem(xs[0], w27, b, 2, nat.PN_BF16, u8=True)
            same = "" if ref is None else (" same bits" if torch.equal(y0, ref) else " BITS DIFFER")
            if ref is None: ref = y0.clone()
            byts = n * h * w * 3 + y0.numel() * 2
            print("%s %-18s %7.1f us  %6.0f GB/s%s" % (name, vname, us, byts / us / 1e3, same), flush=True)
        finally:
            for k in env: del os.environ[k]
