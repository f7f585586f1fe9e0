"""Cycles per tcgen05.mma (M128 x N x K16, bf16) on one SM for small N and operand layouts (csrc/diag/dwtc_probe.cu: umma_cost_kernel)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "posenet-pytorch_b200"), os.path.join(ROOT, "tests")]
import torch
from posenet import _native as nat  # noqa: F401
import abi

torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
lib = abi.load_diag()
out = (C.c_longlong * 2)()
reps = 64
for layout, name in ((0, "SW128, one thread"), (3, "SW128, uniform issue"), (2, "SW64 rows, one thread"), (1, "SW32 rows, one thread")):
    for n in (16, 32, 64, 128):
        for step in (0, 8):
            for _ in range(2):
                rc = lib.pn_debug_umma_cost(n, layout, reps, step, out)
            assert rc == 0, rc
            print("%-20s N %3d  A advance %3d B : issue %6.1f cyc/MMA   complete %6.1f cyc/MMA" % (
                name, n, step * 16, out[0] / (9.0 * reps), out[1] / (9.0 * reps)), flush=True)
