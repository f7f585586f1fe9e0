"""tcgen05 / TMEM / TMA GEMM parity (bf16 in, fp32 accumulate) through the C ABI.  The reference for
each case is the same product computed in fp64 from the bf16-rounded operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import abi  # noqa: E402
from posenet import _native as nat  # noqa: E402

DEV = "cuda"

# (M, K, N): ragged M, every K / N the three architectures produce, multi-tile persistent schedules
SHAPES = [(128, 64, 64), (1089, 512, 512), (1000, 32, 64), (333, 16, 32), (289, 24, 48), (4225, 48, 96),
          (1089, 96, 96), (1100, 96, 192), (700, 192, 192), (578, 192, 384), (289, 384, 384), (2178, 256, 512),
          (1089, 512, 1024), (1089, 1024, 1024), (66049, 32, 64), (16641, 64, 128), (16641, 128, 128),
          (14651, 256, 256), (40000, 128, 256), (1, 64, 16), (129, 1024, 128)]


@pytest.mark.parametrize("m,k,n", SHAPES)
def test_gemm_bf16(m, k, n):
    torch.manual_seed(m + k + n)
    a = (torch.rand(m, k) * 6).to(torch.bfloat16)
    w = (torch.randn(n, k) / k ** 0.5).to(torch.bfloat16)
    b = torch.randn(n)
    ref = (a.double() @ w.double().t() + b.double()).clamp(0, 6)
    y = abi.pwconv(a.to(DEV), w.to(DEV), b.to(DEV), nat.PN_BF16).double().cpu()
    assert (ref == 6).any() and (ref == 0).any()
    err = (y - ref).abs()
    # one bf16 rounding of the result (rel 2^-8) on top of an fp32 accumulation
    bad = err > (ref.abs() * 2.0 ** -8 + 1e-3)
    assert not bad.any(), "%d / %d wrong, max err %g (first bad at %s)" % (
        int(bad.sum()), bad.numel(), float(err.max()), tuple(bad.nonzero()[0].tolist()))


def test_gemm_bf16_is_deterministic_and_reentrant():
    torch.manual_seed(1)
    a = torch.randn(5000, 256).to(torch.bfloat16).to(DEV)
    w = torch.randn(256, 256).to(torch.bfloat16).to(DEV)
    b = torch.randn(256).to(DEV)
    y0 = abi.pwconv(a, w, b, nat.PN_BF16)
    for _ in range(3):
        assert torch.equal(abi.pwconv(a, w, b, nat.PN_BF16), y0)


@pytest.mark.parametrize("n_img,hw,k", [(2, 33 * 33, 1024), (3, 17 * 17, 384), (1, 91 * 161, 256), (1, 5, 256)])
def test_heads_bf16(n_img, hw, k):
    torch.manual_seed(hw)
    a = (torch.rand(n_img * hw, k) * 3).to(torch.bfloat16)
    w = torch.zeros(128, k)
    w[:115] = torch.randn(115, k) / k ** 0.5 * 3
    w = w.to(torch.bfloat16)
    b = torch.zeros(128)
    b[:115] = torch.randn(115)
    outs = abi.heads(a.to(DEV), w.to(DEV), b.to(DEV), n_img, hw, nat.PN_BF16)
    z = (a.double() @ w.double().t() + b.double()).reshape(n_img, hw, 128).permute(0, 2, 1)
    refs = [torch.sigmoid(z[:, :17]), z[:, 17:51], z[:, 51:83], z[:, 83:115]]
    for o, r in zip(outs, refs):
        o = o.double().cpu()
        assert not torch.isnan(o).any()
        assert float((o - r).abs().max()) < 2e-4 * max(1.0, float(r.abs().max()))
