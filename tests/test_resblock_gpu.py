"""vb_resblock_pair (fused HiFi-GAN ResBlock1 conv pair) and vb_mrf_combine vs fp32 torch math on the same
fp16-rounded operands (reference vocoder/hifigan/modules/hifigan.py:51-58,132-138)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from versband_b200 import ops  # noqa: E402
from versband_b200.dit import pack_conv_weight  # noqa: E402

torch.set_grad_enabled(False)
DEV = "cuda"
SLOPE = 0.1


def lrelu(x, s=SLOPE):
    return torch.where(x > 0, x, x * s)


@pytest.mark.parametrize("C,k,d", [(64, 3, 1), (64, 3, 5), (64, 7, 3), (32, 3, 3), (32, 7, 5), (32, 11, 1), (32, 11, 5)])
@pytest.mark.parametrize("B,T", [(1, 1000), (2, 4111), (1, 60000)])
def test_pair_vs_torch(C, k, d, B, T):
    if not ops.resblock_pair_fits(C, k, d):
        pytest.skip("configuration does not fit shared memory")
    g = torch.Generator(device=DEV).manual_seed(C * 100 + k * 10 + d)
    x = torch.randn(B, T, C, device=DEV, generator=g)
    y = lrelu(x).half()                                   # activated representation in HBM
    w1 = (torch.randn(C, C, k, device=DEV, generator=g) / (C * k) ** 0.5).half()
    w2 = (torch.randn(C, C, k, device=DEV, generator=g) / (C * k) ** 0.5).half()
    b1 = torch.randn(C, device=DEV, generator=g) * 0.1
    b2 = torch.randn(C, device=DEV, generator=g) * 0.1
    out = torch.zeros(B, T, C, device=DEV, dtype=torch.float16)
    ops.resblock_pair(y, out, pack_conv_weight(w1.float(), dtype=torch.float16), pack_conv_weight(w2.float(), dtype=torch.float16),
                      b1, b2, B=B, T=T, C=C, k=k, dil=d, slope=SLOPE)
    torch.cuda.synchronize()
    yf = y.float()
    xr = torch.where(yf < 0, yf / SLOPE, yf).permute(0, 2, 1)                     # recovered pre-activation [B, C, T]
    t1 = lrelu(F.conv1d(yf.permute(0, 2, 1), w1.float(), b1, dilation=d, padding=(k - 1) // 2 * d)).half().float()
    ref = lrelu(xr + F.conv1d(t1, w2.float(), b2, padding=(k - 1) // 2)).permute(0, 2, 1)
    diff = (out.float() - ref).abs()
    err = float(diff.max() / ref.abs().max())
    bad_rows = (diff.amax(-1) > 2e-3 * float(ref.abs().max())).nonzero()
    print(f"C={C} k={k} d={d} B={B} T={T}: max/peak {err:.3e}; bad rows {bad_rows.shape[0]} first {bad_rows[:6].tolist()}")
    assert err < 2e-3


def test_mrf_combine():
    g = torch.Generator(device=DEV).manual_seed(3)
    xs = [torch.randn(2, 777, 64, device=DEV, generator=g) for _ in range(3)]
    ys = [lrelu(x).half() for x in xs]
    out = torch.zeros(2, 777, 64, device=DEV, dtype=torch.float16)
    ops.mrf_combine(ys, out, neg=1.0 / SLOPE, scale=1.0 / 3, slope_out=0.01)
    torch.cuda.synchronize()
    rec = [torch.where(y.float() < 0, y.float() / SLOPE, y.float()) for y in ys]
    ref = lrelu(sum(rec) / 3, 0.01)
    assert float((out.float() - ref).abs().max()) < 2e-3
