"""Debug aid: run one fused block shape under forced tile shapes and report where it differs from the two-kernel path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "posenet-pytorch_b200"), os.path.join(ROOT, "tests"), ROOT]
import torch
import abi
from posenet import _native as nat

def run(shape, tile):
    if tile: os.environ["PN_SEP_TILE"] = tile
    else: os.environ.pop("PN_SEP_TILE", None)
    n, h, w, cin, cout, stride, dil = shape
    g = torch.Generator().manual_seed(1)
    x = (torch.rand((n, h, w, cin), generator=g) * 6).to(torch.bfloat16).cuda()
    w9 = (torch.randn((9, cin), generator=g) * 0.35).cuda()
    bd = (torch.randn(cin, generator=g) * 0.3).cuda()
    wp = (torch.randn((cout, cin), generator=g) * (1.5 / cin ** 0.5)).to(torch.bfloat16).cuda()
    bp = (torch.randn(cout, generator=g) * 0.5).cuda()
    t = abi.dwconv(x, w9, bd, stride, dil, nat.PN_BF16)
    ref = abi.pwconv(t.reshape(-1, cin), wp, bp, nat.PN_BF16).reshape(t.shape[0], t.shape[1], t.shape[2], cout).float()
    try:
        y = abi.sepconv(x, w9, bd, wp, bp, stride, dil).float()
        torch.cuda.synchronize()
    except Exception as e:
        print(shape, tile, "EXC", str(e)[:200]); return
    bad = ((y - ref).abs() > 0.05 * ref.abs().max()) | torch.isnan(y)
    print(shape, tile, "bad cells %d / %d" % (int(bad.sum()), bad.numel()))
    if bad.any():
        idx = bad.nonzero()
        print("  imgs", idx[:, 0].unique().tolist()[:8], "rows", idx[:, 1].unique().tolist()[:40], "cols", idx[:, 2].unique().tolist()[:40],
              "chan range", int(idx[:, 3].min()), int(idx[:, 3].max()))

if __name__ == "__main__":
    shp = (2, 65, 65, 256, 512, 2, 1)
    for tile in sys.argv[1:] or ["", "6,11,2", "17,4,1", "8,4,2", "6,11,1"]:
        run(shp, tile)
