"""Op-level parity of the SIMT kernels (through the C-ABI) against plain PyTorch fp32 references
and, for the router, against the oracle's Gumbel-softmax on identical inputs (bit-exact indices)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import accompband_ref as R  # noqa: E402
from versband_b200 import ops  # noqa: E402

H16 = torch.float16
torch.set_grad_enabled(False)


def close(a, b, tol):
    a, b = a.float().cpu(), b.float().cpu()
    err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-9))
    assert err < tol, err


@pytest.mark.parametrize("kind", [0, 1])
def test_norm_mod(kind):
    torch.manual_seed(0)
    B, T, D = 2, 333, 768
    x = torch.randn(B * T, D, device="cuda") * 3 + 0.5
    w = torch.randn(D, device="cuda")
    mod = torch.randn(B, 4 * D, device="cuda")
    out = torch.empty(B * T, D, device="cuda")
    outb = torch.empty(B * T, D, device="cuda", dtype=H16)
    if kind == 0:
        ops.norm_mod(x, kind=0, w=w, shift=mod[:, D:], scale=mod[:, 2 * D:], mod_ld=4 * D, rows_per_batch=T,
                     eps=1e-5, out_f32=out, out_fp16=outb)
        ref = R.rms_norm(x.view(B, T, D), w, 1e-5)
    else:
        ops.norm_mod(x, kind=1, shift=mod[:, D:], scale=mod[:, 2 * D:], mod_ld=4 * D, rows_per_batch=T, eps=1e-6,
                     out_f32=out, out_fp16=outb)
        ref = F.layer_norm(x.view(B, T, D), (D,), None, None, 1e-6)
    ref = R.modulate(ref, mod[:, D:2 * D], mod[:, 2 * D:3 * D]).reshape(B * T, D)
    close(out, ref, 2e-6)
    close(outb, ref, 5e-3)


def test_softmax_rows():
    torch.manual_seed(1)
    S = torch.randn(500, 752, device="cuda") * 4
    P = torch.full((500, 752), 7.0, device="cuda", dtype=H16)
    ops.softmax_rows(S, 752, P, 752, 500, 750)
    close(P[:, :750], torch.softmax(S[:, :750], -1), 5e-3)
    assert (P[:, 750:] == 0).all()


def test_attn_smallkv():
    torch.manual_seed(2)
    B, T, H, hd, L = 2, 200, 8, 96, 80
    q = torch.randn(B, T, H * hd, device="cuda").to(H16)
    kv = torch.randn(B * L, 2 * H * hd, device="cuda")
    gate = torch.randn(H, device="cuda")
    base = torch.randn(B * T, H * hd, device="cuda")
    out = torch.empty(B * T, H * hd, device="cuda")
    ops.attn_smallkv(q, (T * H * hd, H * hd, hd), kv, kv[:, H * hd:], 2 * H * hd, B=B, T=T, H=H, hd=hd, L=L,
                     scale=hd ** -0.5, gate=gate, base=base, out_f32=out)
    k = kv[:, :H * hd].view(B, L, H, hd).permute(0, 2, 1, 3)
    v = kv[:, H * hd:].view(B, L, H, hd).permute(0, 2, 1, 3)
    o = R.sdpa(q.float().view(B, T, H, hd).permute(0, 2, 1, 3), k, v).permute(0, 2, 1, 3)
    ref = base + (o * gate.tanh().view(1, 1, H, 1)).reshape(B * T, H * hd)
    close(out, ref, 1e-5)


def test_router_bit_exact_indices_vs_oracle():
    """Identical fp32 router inputs + identical Exp(1) noise => identical hard top-1 indices."""
    torch.manual_seed(3)
    B, T, D, E = 2, 750, 768, 4
    N = B * T
    feat = torch.randn(N, D)
    Wc, bc = torch.randn(E, D) * 0.05, torch.randn(E) * 0.1
    ac_logits = torch.randn(N, E)
    hl = torch.randn(B, 2)
    e0, e1, e2 = torch.empty(N, 2).exponential_(), torch.empty(N, E).exponential_(), torch.empty(N, E).exponential_()
    # oracle (vocal2music_moe.py:132-151 restated)
    hl_p, _ = R.gumbel_softmax(hl.repeat_interleave(T, 0), 1.0, False, e0)
    cap_logits = F.linear(feat, Wc, bc)
    cap_p, cap_idx = R.gumbel_softmax(cap_logits, 2.0, True, e1)
    ac_p, ac_idx = R.gumbel_softmax(ac_logits, 2.0, True, e2)
    d = "cuda"
    idx_c = torch.empty(N, dtype=torch.int32, device=d)
    idx_a = torch.empty_like(idx_c)
    sc, sa = torch.empty(N, device=d), torch.empty(N, device=d)
    lg = torch.empty(N, E, device=d)
    ops.router(feat.to(d), Wc.to(d), bc.to(d), ac_logits.to(d), E, hl.to(d), 2, e0.to(d), e1.to(d), e2.to(d), 2.0,
               N, T, idx_c, idx_a, sc, sa, lg)
    torch.cuda.synchronize()
    # the acoustic gate sees bit-identical logits: indices must be exactly equal
    assert torch.equal(idx_a.cpu().long(), ac_idx)
    # the caption gate's logits are a 768-term fp32 dot product (different summation order):
    # indices must agree wherever the oracle's decision margin exceeds that rounding noise
    z = (cap_logits - e1.log()) / 2.0
    top2 = z.topk(2, -1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert torch.equal(idx_c.cpu().long()[safe], cap_idx[safe]) and float(safe.float().mean()) > 0.999
    close(lg, cap_logits, 1e-5)
    ref_sc = cap_p.gather(1, cap_idx[:, None])[:, 0] * hl_p[:, 0]
    ref_sa = ac_p.gather(1, ac_idx[:, None])[:, 0] * hl_p[:, 1]
    close(sc.cpu()[safe], ref_sc[safe], 1e-5)
    close(sa, ref_sa, 1e-5)


def test_moe_sort_and_gather():
    torch.manual_seed(4)
    N, E, D = 1500, 4, 768
    d = "cuda"
    idx_c = torch.randint(0, E, (N,), dtype=torch.int32, device=d)
    idx_a = torch.randint(0, E, (N,), dtype=torch.int32, device=d)
    idx_a[:700] = 2  # unbalanced
    sc, sa = torch.rand(N, device=d), torch.rand(N, device=d)
    max_rows = ((2 * N + 127) // 128 + 2 * E) * 128
    row_map = torch.full((max_rows,), 12345, dtype=torch.int32, device=d)
    row_scale = torch.empty(max_rows, device=d)
    tile_group = torch.full((max_rows // 128,), -7, dtype=torch.int32, device=d)
    used = torch.zeros(1, dtype=torch.int32, device=d)
    lb = torch.zeros(1, device=d)
    ops.moe_sort(idx_c, idx_a, sc, sa, N, E, row_map, row_scale, tile_group, used, max_rows, lb)
    x = torch.randn(N, D, device=d).to(H16)
    xs = torch.full((max_rows, D), 9.0, device=d, dtype=H16)
    ops.gather_rows(x, row_map, used, N, xs, max_rows)
    torch.cuda.synchronize()
    rm, tg, nt = row_map.cpu(), tile_group.cpu(), int(used)
    valid = rm[rm >= 0]
    assert sorted(valid.tolist()) == list(range(2 * N))           # every (token, group) exactly once
    assert (rm[nt * 128:] == -1).all()
    bucket = torch.cat([idx_c.cpu(), idx_a.cpu() + E]).long()      # entry g*N+n -> bucket g*E+idx
    for s in range(nt * 128):
        if rm[s] >= 0:
            assert bucket[rm[s]] == tg[s // 128]
    # deterministic order: inside a bucket tokens keep ascending order
    for t in range(nt):
        seg = rm[t * 128:(t + 1) * 128]
        seg = seg[seg >= 0]
        assert (seg[1:] > seg[:-1]).all()
    scale_all = torch.cat([sc, sa]).cpu()
    ok = rm >= 0
    assert torch.equal(row_scale.cpu()[ok], scale_all[rm[ok].long()])
    assert torch.equal(xs.cpu()[ok], x.cpu()[(rm[ok] % N).long()])
    # load-balancing loss of vocal2music_moe.py:95-115 with hard probabilities
    usage = torch.zeros(2 * E).index_add_(0, bucket, scale_all)
    u = usage / (usage.sum() * E + 1e-10)
    ref_lb = torch.mean(u * torch.log(u + 1e-10))
    assert abs(float(lb) - float(ref_lb)) < 1e-5


def test_layout_embed_pool_ew():
    torch.manual_seed(5)
    d = "cuda"
    x = torch.randn(2, 20, 75, device=d)
    cl = torch.full((2, 75, 32), 5.0, device=d, dtype=H16)
    ops.nct_to_ntc(x, cl, scale=0.5)
    close(cl[:, :, :20], (x * 0.5).transpose(1, 2), 5e-3)
    assert (cl[:, :, 20:] == 0).all()
    e_c, e_u = torch.randn(2 * 75, 32, device=d), torch.randn(2 * 75, 32, device=d)
    x0 = x.clone()
    xcl = torch.zeros(4, 75, 32, device=d)  # two channels-last replicas of the updated latent
    ops.ntc_to_nct(e_c, x, e_u=e_u, ld=32, s=3.0, dt=0.04, x_cl=xcl)
    mix = (e_u + 3.0 * (e_c - e_u))[:, :20].view(2, 75, 20).transpose(1, 2)
    close(x, x0 + 0.04 * mix, 1e-6)
    for r in range(2):
        assert torch.equal(xcl[2 * r:2 * r + 2, :, :20], x.transpose(1, 2))
    assert (xcl[:, :, 20:] == 0).all()
    idx = torch.randint(0, 130, (3, 50), device=d)
    tab = torch.randn(130, 64, device=d)
    out = torch.empty(3, 50, 64, device=d, dtype=H16)
    ops.embed_rows(idx, tab, out)
    close(out, tab[idx], 5e-3)
    a, b = torch.randn(2, 20, 64, device=d), torch.randn(2, 20, 64, device=d)
    pooled = torch.empty(2, 10, 64, device=d)
    ops.pool2_sum(a, b, B=2, Tin=20, Tout=10, D=64, out_f32=pooled)
    ref = F.avg_pool1d(a.transpose(1, 2), 2).transpose(1, 2) + F.avg_pool1d(b.transpose(1, 2), 2).transpose(1, 2)
    close(pooled, ref, 1e-6)
    up = torch.empty(2, 40, 64, device=d)
    ops.ew(a, b, alpha=2.0, beta=-1.0, act=ops.ACT_LRELU, slope=0.1, out_f32=up, rep=2)
    close(up, F.leaky_relu(2 * a - b, 0.1).repeat_interleave(2, 1), 1e-6)


def test_groupnorm_and_conv_post():
    torch.manual_seed(6)
    d = "cuda"
    B, T, C = 2, 300, 384
    x = torch.randn(B, T, C, device=d) * 2 + 1
    w, b = torch.randn(C, device=d), torch.randn(C, device=d)
    mean, rstd = torch.empty(B * 32, device=d), torch.empty(B * 32, device=d)
    out = torch.empty(B, T, C, device=d, dtype=H16)
    ops.gn_stats(x, B, T, C, 32, 1e-6, mean, rstd)
    ops.gn_apply(x, mean, rstd, w, b, B, T, C, 32, ops.ACT_SILU, out)
    ref = F.silu(F.group_norm(x.transpose(1, 2), 32, w, b, 1e-6)).transpose(1, 2)
    close(out, ref, 8e-3)
    # the single-launch cluster kernel: same statistics (fp64 partial sums), same fp16 output up to rounding
    for Bf, Tf, Cf in ((2, 300, 384), (1, 750, 1536), (3, 1501, 768), (1, 9, 128)):
        xf = torch.randn(Bf, Tf, Cf, device=d) * 2 + 1
        wf, bf = torch.randn(Cf, device=d), torch.randn(Cf, device=d)
        assert ops.gn_fused_fits(Tf, Cf, 32)
        of = torch.zeros(Bf, Tf, Cf, device=d, dtype=H16)
        ops.gn_fused(xf, wf, bf, Bf, Tf, Cf, 32, 1e-6, ops.ACT_SILU, of)
        rf = F.silu(F.group_norm(xf.double().transpose(1, 2), 32, wf.double(), bf.double(), 1e-6)).transpose(1, 2)
        close(of, rf.float(), 2e-3)
        m2, r2, o2 = torch.empty(Bf * 32, device=d), torch.empty(Bf * 32, device=d), torch.empty_like(of)
        ops.gn_stats(xf, Bf, Tf, Cf, 32, 1e-6, m2, r2)
        ops.gn_apply(xf, m2, r2, wf, bf, Bf, Tf, Cf, 32, ops.ACT_SILU, o2)
        assert float((of.float() - o2.float()).abs().max()) <= 4e-3 * float(o2.float().abs().max())
    assert not ops.gn_fused_fits(750, 384 + 2, 32) and not ops.gn_fused_fits(4, 384, 32)
    xb = torch.randn(B, 500, 32, device=d).to(H16)
    wp = torch.randn(7, 32, device=d) * 0.1
    wav = torch.empty(B, 500, device=d)
    ops.conv_post(xb, wp, 0.05, B, 500, 32, 7, wav)
    ref = torch.tanh(F.conv1d(xb.float().transpose(1, 2), wp.t()[None].contiguous(), torch.tensor([0.05], device=d), padding=3))[:, 0]
    close(wav, ref, 2e-3)  # fp16 inputs, tanh output; fp32 accumulation order differs


def test_transpose_f32():
    d = "cuda"
    for shape in ((6000, 768), (37, 5), (3, 750, 96), (2, 33, 130)):
        x = torch.randn(*shape, device=d)
        y = ops.transpose_f32(x)
        R, Cc = shape[-2], shape[-1]
        Rp = (R + 3) // 4 * 4
        assert y.shape == (*shape[:-2], Cc, Rp)
        assert torch.equal(y[..., :R], x.transpose(-1, -2)) and float(y[..., R:].abs().sum()) == 0.0
    big = torch.randn(4, 100, 64, device=d)
    v = big[:, 10:50, 8:40]  # strided view: ld 64, batch stride 6400
    assert torch.equal(ops.transpose_f32(v), v.transpose(1, 2).contiguous())
