"""Runs pn_dwtc_probe (csrc/diag/dwtc_probe.cu) on a B200 and compares with numpy: shifted-descriptor depthwise on tcgen05,
A-from-TMEM pointwise.  Prints max errors for both base-offset variants and several chunk geometries."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "posenet-pytorch_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from posenet import _native as nat
import abi

lib = nat.load()
P = lambda t: C.c_void_p(t.data_ptr())


def bf16_round(a):
    return torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()


def run(h, w, dil, wp, x_org, band_x0, tw, r0, chunk, flags, seed=0):
    rng = np.random.default_rng(seed)
    x = bf16_round(rng.uniform(0, 6, (h, w, 64)).astype(np.float32))
    wdw = bf16_round(rng.normal(0, 0.4, (9, 64)).astype(np.float32))
    bias = rng.normal(0, 0.3, 64).astype(np.float32)
    pww = bf16_round(rng.normal(0, 0.2, (64, 64)).astype(np.float32))
    diag = np.zeros((9, 16, 64), np.float32)
    for t in range(9):
        for g in range(4):
            for n in range(16):
                diag[t, n, 16 * g + n] = wdw[t, 16 * g + n]
    q0 = chunk * 128
    row0 = q0 // wp                      # first output row (band flat space) the chunk touches
    rows_box = (q0 + 127) // wp - row0 + 1 + 2 * dil
    qoff = q0 - row0 * wp
    y_org = row0 - dil
    xd = torch.from_numpy(x).to(torch.bfloat16).cuda()
    dd = torch.from_numpy(diag.reshape(144, 64)).to(torch.bfloat16).cuda()
    wd_ = torch.from_numpy(pww).to(torch.bfloat16).cuda()
    bd = torch.from_numpy(bias).cuda()
    out_dw = torch.zeros((128, 64), dtype=torch.float32, device="cuda")
    out_pw = torch.zeros((128, 64), dtype=torch.float32, device="cuda")
    abi.check_diag(abi.load_diag().pn_dwtc_probe(P(xd), h, w, P(dd), P(wd_), P(bd), P(out_dw), P(out_pw), wp, dil, qoff, rows_box, x_org, y_org,
                                flags, nat.stream_ptr()), "pn_dwtc_probe")
    torch.cuda.synchronize()
    got_dw, got_pw = out_dw.cpu().numpy(), out_pw.cpu().numpy()
    # reference on the valid positions
    xp = np.zeros((h + 4 * dil + 8, w + 4 * dil + 8, 64), np.float32)
    o = 2 * dil + 4
    xp[o:o + h, o:o + w] = x
    err_dw, err_pw, nvalid = 0.0, 0.0, 0
    for r in range(128):
        q = q0 + r
        ty, tx = q // wp, q % wp
        gx = band_x0 + tx
        if tx >= tw or gx >= w or ty >= h:
            continue
        acc = np.zeros(64, np.float32)
        for t in range(9):
            dy, dx = t // 3 - 1, t % 3 - 1
            acc += xp[o + ty + dy * dil, o + gx + dx * dil] * wdw[t]
        ref_dw = bf16_round(np.clip(acc + bias, 0, 6).astype(np.float32))
        e = np.abs(got_dw[r] - ref_dw).max()
        err_dw = max(err_dw, e)
        ref_pw = got_dw[r].astype(np.float64) @ pww.T.astype(np.float64)
        err_pw = max(err_pw, np.abs(got_pw[r] - ref_pw).max())
        nvalid += 1
    print("h %d w %d dil %d wp %d chunk %d flags %d: valid rows %d  max|dw err| %.4g  max|pw err| %.4g" % (
        h, w, dil, wp, chunk, flags, nvalid, err_dw, err_pw), flush=True)


torch.cuda.set_device(0)
for flags in (0, 1):
    # full-width band, pitch W + D (shared zero gap): 33 wide, dilation 1
    for chunk in (0, 1, 3):
        run(20, 33, 1, 34, -1, 0, 33, 0, chunk, flags)
    # interior band with real neighbours either side: band of 14 columns starting at x = 8, pitch 16
    run(24, 40, 1, 16, 7, 8, 14, 0, 1, flags)
    # dilation 2, full width, pitch W + 2
    run(20, 29, 2, 31, -2, 0, 29, 0, 2, flags)
