"""vb_gemm (tcgen05/TMA) against a plain PyTorch fp32 reference of the same op.

Inputs are rounded to fp16 first so products are exact in fp32; the only difference left is
the summation order, hence the tight tolerance."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from versband_b200 import ops  # noqa: E402


def bf(x):
    return x.to(torch.float16)


def close(got, ref, tol=2e-5):
    scale = ref.abs().max().clamp_min(1e-6)
    err = (got.float() - ref).abs().max() / scale
    assert err < tol, f"max rel-to-peak err {err.item():.3e}"


@pytest.mark.parametrize("M,N,K", [(750, 768, 768), (128, 32, 64), (1500, 2304, 768), (80, 768, 1024),
                                   (750, 20, 768), (333, 750, 96), (750, 192, 512)])
@pytest.mark.parametrize("bn", [0, 32, 64, 128])
def test_plain(M, N, K, bn):
    torch.manual_seed(0)
    a = bf(torch.randn(M, K, device="cuda"))
    b = bf(torch.randn(N, K, device="cuda"))
    ld = (N + 7) // 8 * 8
    out = torch.full((M, ld), 7.0, device="cuda")
    ops.gemm(a, b, out_f32=out, out_ld=ld, bn=bn)
    ref = a.float() @ b.float().t()
    close(out[:, :N], ref)
    if ld > N:
        assert (out[:, N:] == 7.0).all()


def test_bn96():
    torch.manual_seed(0)
    a = bf(torch.randn(750, 512, device="cuda"))
    b = bf(torch.randn(192, 512, device="cuda"))
    out = torch.zeros(750, 192, device="cuda")
    ops.gemm(a, b, out_f32=out, out_ld=192, bn=96)
    close(out, a.float() @ b.float().t())


def test_ragged_k_oob_fill():
    # K=750 stored with pitch 752: the TMA zero-fills k in [750, 768)
    torch.manual_seed(1)
    M, N, K, P = 300, 96, 750, 752
    abuf = bf(torch.randn(M, P, device="cuda"))
    bbuf = bf(torch.randn(N, P, device="cuda"))
    a, b = abuf[:, :K], bbuf[:, :K]
    out = torch.zeros(M, N, device="cuda")
    ops.gemm(a, b, K=K, out_f32=out, out_ld=N)
    close(out, a.float() @ b.float().t())


def test_epilogue_bias_act_gate_residual_fp16():
    torch.manual_seed(2)
    B, T, N, K = 2, 375, 768, 768
    M = B * T
    a = bf(torch.randn(M, K, device="cuda"))
    w = bf(torch.randn(N, K, device="cuda") * 0.05)
    bias = torch.randn(N, device="cuda")
    gate = torch.randn(B, N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    out = torch.zeros(M, N, device="cuda")
    outb = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    ops.gemm(a, w, bias=bias, act=ops.ACT_SILU, gate=gate, gate_rows=T, gate_ld=N, residual=res,
             out_f32=out, out_ld=N, out_fp16=outb, outb_ld=N, act_out=ops.ACT_LRELU,
             act_out_slope=0.1, out_scale=0.5, alpha=0.25)
    v = F.silu((a.float() @ w.float().t()) * 0.25 + bias)
    v = v * gate.repeat_interleave(T, 0) + res
    close(out, v, 2e-5)
    close(outb, F.leaky_relu(v * 0.5, 0.1), 5e-3)


@pytest.mark.parametrize("act", [ops.ACT_GELU, ops.ACT_TANH, ops.ACT_LRELU])
def test_acts(act):
    torch.manual_seed(3)
    a = bf(torch.randn(200, 128, device="cuda"))
    w = bf(torch.randn(64, 128, device="cuda") * 0.1)
    out = torch.zeros(200, 64, device="cuda")
    ops.gemm(a, w, act=act, act_slope=0.01, out_f32=out, out_ld=64)
    z = a.float() @ w.float().t()
    ref = {ops.ACT_GELU: F.gelu(z), ops.ACT_TANH: torch.tanh(z), ops.ACT_LRELU: F.leaky_relu(z, 0.01)}[act]
    close(out, ref, 1e-5)


def test_swiglu_pairs():
    torch.manual_seed(4)
    M, K, H = 750, 768, 512
    a = bf(torch.randn(M, K, device="cuda"))
    w1 = bf(torch.randn(H, K, device="cuda") * 0.05)
    w3 = bf(torch.randn(H, K, device="cuda") * 0.05)
    w13 = torch.stack([w1, w3], 1).reshape(2 * H, K).contiguous()
    h = torch.zeros(M, H, device="cuda", dtype=torch.float16)
    ops.gemm(a, w13, swiglu=True, out_fp16=h, outb_ld=H)
    ref = F.silu(a.float() @ w1.float().t()) * (a.float() @ w3.float().t())
    close(h, ref, 6e-3)


@pytest.mark.parametrize("Cin,Cout,k,dil,T", [(128, 96, 3, 1, 750), (64, 64, 7, 3, 500), (20, 64, 5, 1, 750),
                                              (80, 128, 7, 1, 300), (256, 256, 11, 5, 1000)])
def test_conv1d_implicit(Cin, Cout, k, dil, T):
    """channels-last implicit-GEMM conv with 'same' padding from TMA out-of-bounds zero fill."""
    torch.manual_seed(5)
    B = 2
    Cp = (Cin + 7) // 8 * 8          # row pitch must be a multiple of 16 bytes
    xbuf = bf(torch.randn(B, T, Cp, device="cuda"))
    x = xbuf[:, :, :Cin]
    w = bf(torch.randn(Cout, Cin, k, device="cuda") * 0.1)
    kpt = (Cin + 63) // 64
    wp = torch.zeros(Cout, k, kpt * 64, device="cuda", dtype=torch.float16)
    wp[:, :, :Cin] = w.permute(0, 2, 1)
    wp = wp.reshape(Cout, k * kpt * 64)
    pad = (k * dil - dil) // 2
    out = torch.zeros(B, T, Cout, device="cuda")
    ops.gemm(x, wp, M=T, G1=B, kpt=kpt, dil=dil, pad=pad, a_c2=(1, 0),
             out_f32=out, out_ld=Cout, out_g=(T * Cout, 0))
    ref = F.conv1d(x.float().transpose(1, 2), w.float(), padding=pad, dilation=dil).transpose(1, 2)
    close(out, ref)


def test_batched_attention_shapes():
    """S = q k^T per (b, h) and O = P V via the V^T layout."""
    torch.manual_seed(6)
    B, H, T, hd = 2, 8, 750, 96
    Tp = 752
    q = bf(torch.randn(B, H, T, hd, device="cuda"))
    k = bf(torch.randn(B, H, T, hd, device="cuda"))
    s = torch.zeros(B, H, T, Tp, device="cuda")
    ops.gemm(q.view(B * H, T, hd), k.view(B * H, T, hd), M=T, N=T, K=hd, G1=H, G2=B,
             a_c2=(1, H), b_c2=(1, H, 0), alpha=hd ** -0.5,
             out_f32=s, out_ld=Tp, out_g=(T * Tp, H * T * Tp))
    ref = torch.einsum("bhtd,bhsd->bhts", q.float(), k.float()) * hd ** -0.5
    close(s[..., :T], ref)
    p = bf(torch.softmax(ref, -1))
    pbuf = torch.zeros(B, H, T, Tp, device="cuda", dtype=torch.float16)
    pbuf[..., :T] = p
    v = bf(torch.randn(B, H, T, hd, device="cuda"))
    vt = torch.zeros(B, H, hd, Tp, device="cuda", dtype=torch.float16)
    vt[..., :T] = v.transpose(2, 3)
    o = torch.zeros(B * T, H * hd, device="cuda")
    ops.gemm(pbuf.view(B * H, T, Tp)[:, :, :T], vt.view(B * H, hd, Tp)[:, :, :T], M=T, N=hd, K=T,
             G1=H, G2=B, a_c2=(1, H), b_c2=(1, H, 0),
             out_f32=o, out_ld=H * hd, out_g=(hd, T * H * hd))
    ref_o = torch.einsum("bhts,bhsd->bthd", p.float(), v.float()).reshape(B * T, H * hd)
    close(o, ref_o)


def test_qkv_rope_epilogue():
    torch.manual_seed(7)
    B, T, H, hd, D = 2, 750, 8, 96, 768
    Tp = 752
    x = bf(torch.randn(B * T, D, device="cuda"))
    w = bf(torch.randn(3 * D, D, device="cuda") * 0.05)
    freqs = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, device="cuda").float() / hd))
    ang = torch.outer(torch.arange(T, device="cuda").float(), freqs)
    cs = torch.stack([ang.cos(), ang.sin()], -1).contiguous()
    q = torch.zeros(B, H, T, hd, device="cuda", dtype=torch.float16)
    k = torch.zeros_like(q)
    vt = torch.zeros(B, H, hd, Tp, device="cuda", dtype=torch.float16)
    ops.gemm(x, w, mode=ops.EPI_QKV_ROPE, rope_cs=cs, T=T, heads=H, head_dim=hd, Tp=Tp,
             q_out=q, k_out=k, vt_out=vt)
    y = (x.float() @ w.float().t()).view(B, T, 3, H, hd)

    def rope(z):
        zc = torch.view_as_complex(z.reshape(B, T, H, hd // 2, 2))
        return torch.view_as_real(zc * torch.polar(torch.ones_like(ang), ang)[None, :, None]).flatten(3)

    close(q, rope(y[:, :, 0]).permute(0, 2, 1, 3), 6e-3)
    close(k, rope(y[:, :, 1]).permute(0, 2, 1, 3), 6e-3)
    close(vt[..., :T], y[:, :, 2].permute(0, 2, 3, 1), 6e-3)


def test_grouped_tiles_and_row_map():
    """MoE-style: rows sorted by expert in 128-row tiles, scattered back through row_map."""
    torch.manual_seed(8)
    E, K, N, ntok = 4, 512, 768, 750
    expert = torch.randint(0, E, (ntok,), device="cuda")
    order = torch.argsort(expert, stable=True)
    counts = torch.bincount(expert, minlength=E)
    padded = (counts + 127) // 128 * 128
    starts = torch.cumsum(padded, 0) - padded
    max_rows = (ntok + 127) // 128 * 128 + E * 128
    slot_of = torch.empty(ntok, dtype=torch.long, device="cuda")
    row_map = torch.full((max_rows,), -1, dtype=torch.int32, device="cuda")
    tile_group = torch.zeros(max_rows // 128, dtype=torch.int32, device="cuda")
    off = 0
    for e in range(E):
        toks = order[off:off + counts[e]]
        slots = starts[e] + torch.arange(counts[e], device="cuda")
        slot_of[toks] = slots
        row_map[slots] = toks.int()
        tile_group[starts[e] // 128:(starts[e] + padded[e]) // 128] = e
        off += int(counts[e])
    used = torch.tensor([int(padded.sum()) // 128], dtype=torch.int32, device="cuda")
    x = bf(torch.randn(ntok, K, device="cuda"))
    xs = torch.zeros(max_rows, K, device="cuda", dtype=torch.float16)
    xs[slot_of] = x
    w = bf(torch.randn(E, N, K, device="cuda") * 0.05)
    scale = torch.rand(max_rows, device="cuda")
    out = torch.full((ntok, N), 3.0, device="cuda")
    ops.gemm(xs, w, M=max_rows, N=N, K=K, b_c2=(0, 0, 1), tile_group=tile_group, m_tiles_used=used,
             row_map=row_map, rows_out=ntok, row_scale=scale, out_f32=out, out_ld=N)
    ref = torch.einsum("tk,tnk->tn", x.float(), w.float()[expert]) * scale[slot_of][:, None]
    close(out, ref)


def test_transposed_conv_polyphase_addressing():
    """row_mul/row_off/row_off_g1: phase r of a stride-u transposed conv writes rows s*u + r - p."""
    torch.manual_seed(9)
    Cin, Cout, k, u, T = 64, 32, 8, 4, 100
    p = (k - u) // 2
    x = bf(torch.randn(1, T, Cin, device="cuda"))
    w = bf(torch.randn(Cin, Cout, k, device="cuda") * 0.1)
    Q = k // u
    # phase r, tap q (time s - q) uses w[:, :, r + u*q]; taps stored reversed so time = s - (Q-1) + j
    wp = torch.zeros(u, Cout, Q, 64, device="cuda", dtype=torch.float16)
    for r in range(u):
        for j in range(Q):
            qq = Q - 1 - j
            wp[r, :, j, :Cin] = w[:, :, r + u * qq].t()
    wp = wp.reshape(u, Cout, Q * 64)
    Tout = T * u
    out = torch.zeros(Tout, Cout, device="cuda")
    S = T + Q  # s ranges over [0, T + Q - 1)
    ops.gemm(x.view(T, Cin), wp, M=S, N=Cout, K=Q * 64, G1=u, kpt=1, dil=1, pad=Q - 1,
             b_c2=(1, 0, 0), row_mul=u, row_off=-p, row_off_g1=1, rows_out=Tout,
             out_f32=out, out_ld=Cout)
    ref = F.conv_transpose1d(x.float().transpose(1, 2), w.float(), stride=u, padding=p)[0].t()
    close(out, ref)


# ---------------------------------------------------------------------------------------------
# fp32 operands multiplied as TF32 (kind::tf32): the accuracy mode
# ---------------------------------------------------------------------------------------------
def tf32_round(x):
    xi = x.contiguous().view(torch.int32)
    return ((xi + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M,N,K", [(750, 768, 768), (1500, 2304, 768), (300, 20, 96), (257, 192, 512)])
def test_tf32_plain(M, N, K):
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda")
    b = torch.randn(N, K, device="cuda")
    ld = (N + 3) // 4 * 4
    out = torch.zeros(M, ld, device="cuda")
    ops.gemm(a, b, out_f32=out, out_ld=ld)
    ref = a @ b.t()
    err = float((out[:, :N] - ref).abs().max() / ref.abs().max())
    assert err < 1.5e-3, err              # tf32 has a 10-bit mantissa
    # with operands already on the tf32 grid the product is exact up to fp32 summation order
    at, bt = tf32_round(a), tf32_round(b)
    ops.gemm(at, bt, out_f32=out, out_ld=ld)
    close(out[:, :N], at @ bt.t(), 2e-5)


def test_tf32_conv_and_fp32_operand_output():
    torch.manual_seed(1)
    B, T, Cin, Cout, k = 2, 300, 20, 64, 5
    x = tf32_round(torch.randn(B, T, Cin, device="cuda"))
    w = tf32_round(torch.randn(Cout, Cin, k, device="cuda") * 0.2)
    kpt = (Cin + 31) // 32
    wp = torch.zeros(Cout, k, kpt * 32, device="cuda")
    wp[:, :, :Cin] = w.permute(0, 2, 1)
    wp = wp.reshape(Cout, k * kpt * 32)
    nxt = torch.zeros(B, T, Cout, device="cuda")  # fp32 "operand output" slot
    ops.gemm(x, wp, M=T, G1=B, kpt=kpt, dil=1, pad=k // 2, a_c2=(1, 0), out_fp16=nxt, outb_ld=Cout,
             outb_g=(T * Cout, 0), act_out=ops.ACT_LRELU, act_out_slope=0.1)
    ref = F.leaky_relu(F.conv1d(x.transpose(1, 2), w, padding=k // 2), 0.1).transpose(1, 2)
    close(nxt, ref, 2e-5)
