"""vb_gemm_chain (several dependent GEMM problems in one persistent launch) vs the same problems as
separate vb_gemm launches, and the DiT evaluation with the chained MoE launch vs the unchained one."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import synth  # noqa: E402
from versband_b200 import ops  # noqa: E402
from versband_b200.dit import DiTEngine  # noqa: E402

torch.set_grad_enabled(False)
DEV = "cuda"


@pytest.mark.parametrize("bn", [64, 128])
@pytest.mark.parametrize("M", [200, 1500])
def test_two_layer_ffn_chain_equals_separate_launches(bn, M):
    """SwiGLU FFN: h = silu(x W1^T) * (x W3^T); y = h W2^T + residual  (flag_large_dit_moe.py:470-485)."""
    g = torch.Generator(device=DEV).manual_seed(M + bn)
    D, Fh = 768, 512
    x = (torch.randn(M, D, device=DEV, generator=g)).half()
    w13 = (torch.randn(2 * Fh, D, device=DEV, generator=g) / D ** 0.5).half()
    w2 = (torch.randn(D, Fh, device=DEV, generator=g) / Fh ** 0.5).half()
    res = torch.randn(M, D, device=DEV, generator=g)

    def problems(h, y):
        return [ops.gemm_args(x, w13, swiglu=True, out_fp16=h, outb_ld=Fh, bn=bn),
                ops.gemm_args(h, w2, residual=res, out_f32=y, out_ld=D, bn=bn)]

    h_a, y_a = torch.zeros(M, Fh, device=DEV, dtype=torch.float16), torch.zeros(M, D, device=DEV)
    for pr in problems(h_a, y_a):  # separate launches
        ops._lib.check(ops._lib.lib().vb_gemm(ops.C.byref(pr), ops._stream()), "vb_gemm")
    h_b, y_b = torch.zeros_like(h_a), torch.zeros_like(y_a)
    cnt = torch.zeros(64, device=DEV, dtype=torch.int32)
    ops.gemm_chain(problems(h_b, y_b), [0, 1], cnt, bn=bn)
    torch.cuda.synchronize()
    assert int(cnt[0]) == 0, "dependency wait timed out"
    assert torch.equal(h_a, h_b) and torch.equal(y_a, y_b)
    # and against fp32 math on the same fp16-rounded operands
    hh = x.float() @ w13.float().t()
    href = (torch.nn.functional.silu(hh[:, 0::2]) * hh[:, 1::2])
    yref = href.half().float() @ w2.float().t() + res
    assert float((y_b - yref).abs().max() / yref.abs().max()) < 2e-3


def test_global_dependency_chain():
    """dep_mode 2: the second problem reads rows the first one scatters through a row map."""
    g = torch.Generator(device=DEV).manual_seed(7)
    M, D = 640, 256
    x = torch.randn(M, D, device=DEV, generator=g).half()
    w1 = (torch.randn(D, D, device=DEV, generator=g) / D ** 0.5).half()
    w2 = (torch.randn(D, D, device=DEV, generator=g) / D ** 0.5).half()
    perm = torch.randperm(M, device=DEV, generator=g).int()
    mid = torch.zeros(M, D, device=DEV, dtype=torch.float16)
    out = torch.zeros(M, D, device=DEV)
    cnt = torch.zeros(64, device=DEV, dtype=torch.int32)
    ops.gemm_chain([ops.gemm_args(x, w1, row_map=perm, rows_out=M, out_fp16=mid, outb_ld=D),
                    ops.gemm_args(mid, w2, out_f32=out, out_ld=D)], [0, 2], cnt, bn=64)
    torch.cuda.synchronize()
    assert int(cnt[0]) == 0
    mid_ref = torch.zeros(M, D, device=DEV)
    mid_ref[perm.long()] = x.float() @ w1.float().t()
    ref = mid_ref.half().float() @ w2.float().t()
    assert float((out - ref).abs().max() / ref.abs().max()) < 2e-3


def test_dit_forward_chained_equals_unchained():
    sd = synth.synth_state_dict(synth.load_manifest("dit"), seed=1234)
    eng = DiTEngine(sd, device=DEV)
    B, T = 2, 200
    x, midi, beats, cap, _ = synth.dit_inputs(B, T, seed=5)
    cond = eng.prepare(midi.to(DEV), beats.to(DEV), cap.to(DEV), T)
    expo = [tuple(e.to(DEV) for e in blk) for blk in synth.exp_noise(4, B * T, 4, seed=6)]
    t = torch.tensor([100, 700], device=DEV)
    outs = []
    for chain in (True, False):
        eng.use_chain = chain
        v, _ = eng.forward(x.to(DEV), t, cond, expo)
        torch.cuda.synchronize()
        outs.append(v.clone())
    assert int(eng._workspace(B, T)["chain_cnt"][:, 0].sum()) == 0
    d = float((outs[0] - outs[1]).abs().max() / outs[1].abs().max())
    print("chained vs unchained DiT forward: max rel diff", d)
    assert d < 1e-5
