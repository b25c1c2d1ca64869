"""Thin Python wrappers over the C-ABI ops (one call = one kernel launch on torch's current
stream). No arithmetic happens here: these functions only describe tensors to the library."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import AttnArgs, GemmArgs

ACT_NONE, ACT_SILU, ACT_LRELU, ACT_GELU, ACT_TANH, ACT_LOG10 = 0, 1, 2, 3, 4, 5
EPI_GENERIC, EPI_QKV_ROPE, EPI_ROUTE = 0, 1, 2


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _rev_desc(t: torch.Tensor, rank: int):
    """(dims, strides) of a torch view, innermost first, left-padded to `rank` dims."""
    assert t.dim() <= rank and t.stride(-1) == 1, (t.shape, t.stride())
    dims = list(reversed(t.shape)) + [1] * (rank - t.dim())
    strides = list(reversed(t.stride())) + [0] * (rank - t.dim())
    ext = 1
    for i in range(rank):
        if i > 0 and (dims[i] == 1 or strides[i] == 0):
            strides[i] = max(8, (ext + 7) // 8 * 8)  # any 16-byte multiple: the coordinate is always 0
        ext = max(ext, dims[i] * max(strides[i], 1))
    return dims, strides


def _route(out_fp16, out_f32):
    """An fp32 tensor passed in the operand-output slot is routed to the fp32 output pointer."""
    if out_fp16 is not None and out_fp16.dtype == torch.float32:
        assert out_f32 is None
        return None, out_fp16
    return out_fp16, out_f32


def _check_f32(t: Optional[torch.Tensor], name: str):
    if t is not None:
        assert t.dtype == torch.float32 and t.is_cuda, name


def gemm(a: torch.Tensor, b: torch.Tensor, **kw) -> None:
    """D = epilogue(A @ B^T) on the tcgen05 GEMM: one launch (see `gemm_args` for the arguments)."""
    args = gemm_args(a, b, **kw)
    _lib.check(_lib.lib().vb_gemm(C.byref(args), _stream()), "vb_gemm")


def gemm_chain(problems, dep_mode, counters: torch.Tensor, *, bn: int = 128) -> None:
    """Several dependent GEMM problems (`gemm_args` objects) in ONE persistent launch (include/versband.h:
    vb_gemm_chain). `counters`: zeroed int32 device scratch."""
    n = len(problems)
    arr = (GemmArgs * n)(*problems)
    deps = (C.c_int32 * n)(*dep_mode)
    assert counters.dtype == torch.int32 and counters.is_cuda
    _lib.check(_lib.lib().vb_gemm_chain(arr, deps, n, bn, _ptr(counters), counters.numel(), _stream()), "vb_gemm_chain")


def gemm_args(a: torch.Tensor, b: torch.Tensor, *, M: Optional[int] = None, N: Optional[int] = None,
         K: Optional[int] = None, G1: int = 1, G2: int = 1, kpt: int = 0, dil: int = 0, pad: int = 0,
         a_c2=(0, 0), a_c3=(0, 0), b_c2=(0, 0, 0), bn: int = 0,
         tile_group: Optional[torch.Tensor] = None, m_tiles_used: Optional[torch.Tensor] = None,
         tile_row0: Optional[torch.Tensor] = None,
         mode: int = EPI_GENERIC, alpha: float = 1.0, swiglu: bool = False,
         bias: Optional[torch.Tensor] = None, bias_g1: int = 0, bias_tile: int = 0,
         act: int = ACT_NONE, act_slope: float = 0.0,
         row_scale: Optional[torch.Tensor] = None,
         gate: Optional[torch.Tensor] = None, gate_rows: int = 1, gate_ld: int = 0, gate_g1: int = 0,
         row_map: Optional[torch.Tensor] = None, row_mul: int = 1, row_off: int = 0,
         row_off_g1: int = 0, rows_out: Optional[int] = None,
         residual: Optional[torch.Tensor] = None, res16: Optional[torch.Tensor] = None, res16_neg: float = 1.0,
         accumulate: bool = False,
         out_f32: Optional[torch.Tensor] = None, out_ld: int = 0, out_g=(0, 0),
         out_fp16: Optional[torch.Tensor] = None, outb_ld: int = 0, outb_g=(0, 0),
         act_out: int = ACT_NONE, act_out_slope: float = 0.0, out_scale: float = 1.0,
         rope_cs: Optional[torch.Tensor] = None, T: int = 0, heads: int = 0, head_dim: int = 0,
         Tp: int = 0, q_out=None, k_out=None, vt_out=None, qkv_sec0: int = 0,
         route_vg: Optional[torch.Tensor] = None, route_vg_g1: int = 0) -> GemmArgs:
    """Describe D = epilogue(A @ B^T) (see include/versband.h: vb_gemm_args).

    `a` is a fp16 (or fp32 = TF32 mode) view of up to 4 dims [c3, c2, rows, k]; `b` a view [group, n, k].
    """
    assert a.dtype == b.dtype and a.dtype in (torch.float16, torch.float32) and a.is_cuda and b.is_cuda
    args = GemmArgs()
    args.operand_f32 = int(a.dtype == torch.float32)
    kq = 32 if args.operand_f32 else 64
    ad, as_ = _rev_desc(a, 4)
    bd, bs = _rev_desc(b, 3)
    args.a = _ptr(a)
    args.b = _ptr(b)
    for i in range(4):
        args.a_dims[i], args.a_strides[i] = ad[i], as_[i]
    for i in range(3):
        args.b_dims[i], args.b_strides[i] = bd[i], bs[i]
    args.M = a.shape[-2] if M is None else M
    args.N = b.shape[-2] if N is None else N
    kk = b.shape[-1] if K is None else K
    args.K = (kk + kq - 1) // kq * kq
    args.G1, args.G2 = G1, G2
    args.kpt, args.dil, args.pad = kpt, dil, pad
    args.a_c2_g1, args.a_c2_g2 = a_c2
    args.a_c3_g1, args.a_c3_g2 = a_c3
    args.b_c2_g1, args.b_c2_g2, args.b_c2_tile = b_c2
    args.bn = bn
    args.tile_group, args.m_tiles_used, args.tile_row0 = _ptr(tile_group), _ptr(m_tiles_used), _ptr(tile_row0)
    args.mode, args.alpha, args.swiglu = mode, alpha, int(swiglu)  # 1 = SwiGLU, 2 = GEGLU (gelu_new)
    _check_f32(bias, "bias"); _check_f32(row_scale, "row_scale"); _check_f32(gate, "gate")
    _check_f32(residual, "residual"); _check_f32(out_f32, "out_f32")
    args.bias, args.bias_g1, args.bias_tile = _ptr(bias), bias_g1, bias_tile
    args.act, args.act_slope = act, act_slope
    args.row_scale = _ptr(row_scale)
    args.gate, args.gate_rows, args.gate_ld, args.gate_g1 = _ptr(gate), gate_rows, gate_ld, gate_g1
    args.row_map = _ptr(row_map)
    args.row_mul, args.row_off, args.row_off_g1 = row_mul, row_off, row_off_g1
    args.rows_out = args.M if rows_out is None else rows_out
    args.residual, args.accumulate = _ptr(residual), int(accumulate)
    args.res16, args.res16_neg = _ptr(res16), res16_neg
    args.out_f32, args.out_ld = _ptr(out_f32), out_ld
    args.out_g1, args.out_g2 = out_g
    if out_fp16 is not None:
        assert out_fp16.dtype in (torch.float16, torch.float32)
        args.outb_f32 = int(out_fp16.dtype == torch.float32)
    args.out_fp16, args.outb_ld = _ptr(out_fp16), outb_ld
    args.outb_g1, args.outb_g2 = outb_g
    args.act_out, args.act_out_slope, args.out_scale = act_out, act_out_slope, out_scale
    _check_f32(rope_cs, "rope_cs")
    args.rope_cs, args.T, args.heads, args.head_dim, args.Tp = _ptr(rope_cs), T, heads, head_dim, Tp
    args.q_out, args.k_out, args.vt_out = _ptr(q_out), _ptr(k_out), _ptr(vt_out)
    args.qkv_sec0 = qkv_sec0
    _check_f32(route_vg, "route_vg")
    args.route_vg, args.route_vg_g1 = _ptr(route_vg), route_vg_g1
    return args


# ------------------------------------------------------------------------------------------
# plain-argument ops
# ------------------------------------------------------------------------------------------
def _call(name: str, *args) -> None:
    _lib.check(getattr(_lib.lib(), name)(*args, _stream()), name)


def norm_mod(x, *, kind: int, w=None, b=None, shift=None, scale=None, mod_ld: int = 0,
             rows_per_batch: int = 0, eps: float, out_fp16=None, out_f32=None) -> None:
    """kind 0 RMSNorm / 1 LayerNorm (+ optional affine, + optional adaLN modulate)."""
    rows, D = x.numel() // x.shape[-1], x.shape[-1]
    assert x.dtype == torch.float32 and x.is_contiguous()
    out_fp16, out_f32 = _route(out_fp16, out_f32)
    _call("vb_norm_mod", kind, _ptr(x), _ptr(w), _ptr(b), _ptr(shift), _ptr(scale), mod_ld,
          rows_per_batch or rows, eps, _ptr(out_fp16), _ptr(out_f32), rows, D)


def softmax_rows(S, lds: int, P, ldp: int, rows: int, n: int) -> None:
    _call("vb_softmax_rows", _ptr(S), lds, _ptr(P), ldp, rows, n)


def attn_smallkv(q, q_strides, K, V, kv_ld, *, B, T, H, hd, L, scale, gate=None, base=None, out_fp16=None,
                 out_f32=None) -> None:
    _call("vb_attn_smallkv", _ptr(q), q_strides[0], q_strides[1], q_strides[2], _ptr(K), _ptr(V), kv_ld, B, T, H,
          hd, L, scale, _ptr(gate), _ptr(base), _ptr(out_fp16), _ptr(out_f32))


def timestep_embed(t, freqs, *, out_f32=None, out_fp16=None, dim: int = 256) -> None:
    assert t.dtype == torch.int64
    out_fp16, out_f32 = _route(out_fp16, out_f32)
    _call("vb_timestep_embed", _ptr(t), _ptr(freqs), _ptr(out_f32), _ptr(out_fp16), t.numel(), dim)


def linear_small(x, W, bias, out) -> None:
    rows, D = x.numel() // x.shape[-1], x.shape[-1]
    _call("vb_linear_small", _ptr(x), x.stride(-2) if x.dim() > 1 else D, _ptr(W), _ptr(bias), _ptr(out),
          rows, D, W.shape[0])


def router(cap_feat, Wc, bc, ac_logits, ac_ld, hl_logits, hl_ld, e0, e1, e2, tau, N, T, idx_c, idx_a,
           scale_c, scale_a, cap_logits_out=None, bucket_count=None, usage=None, slot_c=None, slot_a=None,
           cap_partial=None) -> None:
    """cap_feat: fp32/fp16 caption-attention features [N, D] (gate = Wc . feat + bc), or None with
    cap_partial [N, heads, E] = per-head gate contributions from a VB_EPI_ROUTE GEMM."""
    f32 = cap_feat if cap_feat is not None and cap_feat.dtype == torch.float32 else None
    b16 = cap_feat if cap_feat is not None and cap_feat.dtype == torch.float16 else None
    D = cap_feat.shape[-1] if cap_feat is not None else 0
    _check_f32(cap_partial, "cap_partial")
    _call("vb_router", _ptr(f32), _ptr(b16), _ptr(Wc), _ptr(bc), _ptr(ac_logits), ac_ld, _ptr(hl_logits), hl_ld,
          _ptr(e0), _ptr(e1), _ptr(e2), tau, N, T, D, bc.shape[0], _ptr(idx_c), _ptr(idx_a),
          _ptr(scale_c), _ptr(scale_a), _ptr(cap_logits_out), _ptr(bucket_count), _ptr(usage), _ptr(slot_c),
          _ptr(slot_a), _ptr(cap_partial), cap_partial.shape[1] if cap_partial is not None else 0)


def route_dispatch(cap_partial, bc, ac_logits, ac_ld, hl_logits, hl_ld, e0, e1, e2, tau, N, T, E, x, xs, cap_rows,
                   idx_c, idx_a, bucket_count, usage, done_ticket, row_map, row_scale, tile_group, tile_row0,
                   m_tiles_used, lb_out=None, cap_logits_out=None) -> None:
    """Router + dispatch of one MoE block in one launch (include/versband.h: vb_route_dispatch)."""
    _check_f32(cap_partial, "cap_partial")
    _call("vb_route_dispatch", _ptr(cap_partial), cap_partial.shape[1], _ptr(bc), _ptr(ac_logits), ac_ld,
          _ptr(hl_logits), hl_ld, _ptr(e0), _ptr(e1), _ptr(e2), tau, N, T, E, x.shape[-1] * x.element_size(),
          _ptr(x), _ptr(xs), cap_rows, _ptr(idx_c), _ptr(idx_a), _ptr(cap_logits_out), _ptr(bucket_count),
          _ptr(usage), _ptr(done_ticket), None, None, _ptr(row_map), _ptr(row_scale), _ptr(tile_group),
          _ptr(tile_row0), _ptr(m_tiles_used), _ptr(lb_out))


def route_prepare(kv, B, L, H, hd, bq, capg_w, bias_s, vg) -> None:
    _call("vb_route_prepare", _ptr(kv), B, L, H, hd, _ptr(bq), _ptr(capg_w), _ptr(bias_s), _ptr(vg))


def moe_finalize(bucket_count, usage, E, bucket_start, row_map, row_scale, tile_group, m_tiles_used, max_rows,
                 lb_out=None) -> None:
    _call("vb_moe_finalize", _ptr(bucket_count), _ptr(usage), E, _ptr(bucket_start), _ptr(row_map), _ptr(row_scale),
          _ptr(tile_group), _ptr(m_tiles_used), max_rows, _ptr(lb_out))


def moe_dispatch(idx_c, idx_a, slot_c, slot_a, scale_c, scale_a, bucket_count, usage, next_count, next_usage, N, E,
                 x, xs, row_map, row_scale, tile_group, m_tiles_used, max_rows, lb_out=None) -> None:
    _call("vb_moe_dispatch", _ptr(idx_c), _ptr(idx_a), _ptr(slot_c), _ptr(slot_a), _ptr(scale_c), _ptr(scale_a),
          _ptr(bucket_count), _ptr(usage), _ptr(next_count), _ptr(next_usage), N, E, x.shape[-1] * x.element_size(),
          _ptr(x), _ptr(xs), _ptr(row_map), _ptr(row_scale), _ptr(tile_group), _ptr(m_tiles_used), max_rows, _ptr(lb_out))


def moe_scatter(idx_c, idx_a, slot_c, slot_a, scale_c, scale_a, bucket_start, N, E, x, xs, row_map, row_scale) -> None:
    _call("vb_moe_scatter", _ptr(idx_c), _ptr(idx_a), _ptr(slot_c), _ptr(slot_a), _ptr(scale_c), _ptr(scale_a),
          _ptr(bucket_start), N, E, x.shape[-1] * x.element_size(), _ptr(x), _ptr(xs), _ptr(row_map), _ptr(row_scale))


def moe_sort(idx_c, idx_a, scale_c, scale_a, N, E, row_map, row_scale, tile_group, m_tiles_used, max_rows,
             lb_out=None) -> None:
    _call("vb_moe_sort", _ptr(idx_c), _ptr(idx_a), _ptr(scale_c), _ptr(scale_a), N, E, _ptr(row_map),
          _ptr(row_scale), _ptr(tile_group), _ptr(m_tiles_used), max_rows, _ptr(lb_out))


def gather_rows(x, row_map, m_tiles_used, n_tokens, out, max_rows) -> None:
    _call("vb_gather_rows", _ptr(x), _ptr(row_map), _ptr(m_tiles_used), n_tokens, x.shape[-1], _ptr(out),
          max_rows)


def ew(a, b=None, *, alpha=1.0, beta=1.0, act=ACT_NONE, slope=0.0, out_f32=None, out_fp16=None, rep=1):
    rows, D = a.numel() // a.shape[-1], a.shape[-1]
    out_fp16, out_f32 = _route(out_fp16, out_f32)
    _call("vb_ew", _ptr(a), _ptr(b), alpha, beta, act, slope, _ptr(out_f32), _ptr(out_fp16), rows, D, rep)


def nct_to_ntc(x, out_fp16, *, scale=1.0) -> None:
    B, Cc, T = x.shape
    _call("vb_nct_to_ntc", _ptr(x), scale, _ptr(out_fp16), B, Cc, T, out_fp16.shape[-1],
          int(out_fp16.dtype == torch.float32))


def ntc_to_nct(e_c, x, *, e_u=None, ld, s=0.0, dt=0.0, x_cl=None) -> None:
    """x_cl: optional [rep*B, T, 32] channels-last operand copies of x to refresh in the same launch."""
    B, Cc, T = x.shape
    rep = 0
    if x_cl is not None:
        assert x_cl.is_contiguous() and x_cl.shape[-1] == 32 and x_cl.shape[0] % B == 0 and x_cl.shape[1] == T
        rep = x_cl.shape[0] // B
    _call("vb_ntc_to_nct", _ptr(e_c), _ptr(e_u), ld, s, dt, _ptr(x), B, Cc, T, _ptr(x_cl),
          int(x_cl is not None and x_cl.dtype == torch.float32), rep)


def embed_rows(idx, table, out_fp16) -> None:
    _call("vb_embed_rows", _ptr(idx), _ptr(table), table.shape[0], table.shape[1], _ptr(out_fp16), idx.numel(),
          int(out_fp16.dtype == torch.float32))


def pool2_sum(a, b, *, B, Tin, Tout, D, out_fp16=None, out_f32=None) -> None:
    out_fp16, out_f32 = _route(out_fp16, out_f32)
    _call("vb_pool2_sum", _ptr(a), _ptr(b), B, Tin, Tout, D, _ptr(out_fp16), _ptr(out_f32))


GN_SPLIT = 8


def gn_stats(x, B, T, Cc, G, eps, mean, rstd, scratch=None) -> None:
    if scratch is None:
        scratch = torch.empty(B * G * GN_SPLIT * 2, device=x.device, dtype=torch.float64)
    _call("vb_gn_stats", _ptr(x), B, T, Cc, G, eps, _ptr(mean), _ptr(rstd), _ptr(scratch))


def gn_apply(x, mean, rstd, w, b, B, T, Cc, G, act, out_fp16) -> None:
    _call("vb_gn_apply", _ptr(x), _ptr(mean), _ptr(rstd), _ptr(w), _ptr(b), B, T, Cc, G, act, _ptr(out_fp16))


def gn_fused_fits(T: int, Cc: int, G: int) -> bool:
    return bool(_lib.lib().vb_gn_fused_fits(T, Cc, G))


def gn_fused(x, w, b, B, T, Cc, G, eps, act, out_fp16) -> None:
    """GroupNorm + activation -> fp16 in one cluster launch (include/versband.h: vb_gn_fused)."""
    _call("vb_gn_fused", _ptr(x), _ptr(w), _ptr(b), B, T, Cc, G, eps, act, _ptr(out_fp16))


def conv_post(x_fp16, w, bias: float, B, T, Cc, k, wav) -> None:
    _call("vb_conv_post", _ptr(x_fp16), _ptr(w), bias, B, T, Cc, k, _ptr(wav))


def stem_gather(midi, beats, tab_m, tab_b, bias_m, bias_b, *, k, Tpool, slope, out) -> None:
    B, _, Tm = midi.shape
    D = tab_m.shape[-1]
    assert out.shape[-1] == 3 * D and midi.dtype == torch.int64 and beats.dtype == torch.int64
    _call("vb_stem_gather", _ptr(midi), _ptr(beats), _ptr(tab_m), _ptr(tab_b), _ptr(bias_m), _ptr(bias_b),
          tab_m.shape[1], tab_b.shape[1], k, B, Tm, Tpool, D, slope, _ptr(out), int(out.dtype == torch.float32))


def resblock_pair_fits(C: int, k: int, dil: int) -> bool:
    return bool(_lib.lib().vb_resblock_pair_fits(C, k, dil))


def resblock_pair(y_in, y_out, w1, w2, b1, b2, *, B, T, C, k, dil, slope) -> None:
    """One HiFi-GAN ResBlock1 conv pair in one kernel (include/versband.h: vb_resblock_pair)."""
    assert y_in.dtype == torch.float16 and y_out.dtype == torch.float16 and y_in.is_contiguous() and y_out.is_contiguous()
    _call("vb_resblock_pair", _ptr(y_in), _ptr(y_out), _ptr(w1), _ptr(w2), _ptr(b1), _ptr(b2), B, T, C, k, dil, slope)


def mrf_combine(rs, out, *, neg: float, scale: float, slope_out: float) -> None:
    rs = list(rs) + [None] * (3 - len(rs))
    _call("vb_mrf_combine", _ptr(rs[0]), _ptr(rs[1]), _ptr(rs[2]), neg, scale, slope_out, _ptr(out), out.numel())


def adamw(p, g, m, v, *, lr, b1, b2, eps, wd, c1, c2) -> None:
    assert p.dtype == g.dtype == m.dtype == v.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous()
    _call("vb_adamw", _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), lr, b1, b2, eps, wd, c1, c2)


def transpose_f32(x: torch.Tensor) -> torch.Tensor:
    """[..., R, C] fp32 (last two dims dense: stride(-1) == 1) -> contiguous [..., C, ceil4(R)] with zero padding: the
    K-major operand of a GEMM that reduces over R (include/versband.h: vb_transpose_f32)."""
    assert x.dtype == torch.float32 and x.is_cuda and x.stride(-1) == 1 and x.dim() in (2, 3)
    R, Cc = x.shape[-2], x.shape[-1]
    G = x.shape[0] if x.dim() == 3 else 1
    Rp = (R + 3) // 4 * 4
    out = torch.empty((*x.shape[:-2], Cc, Rp), device=x.device, dtype=torch.float32)
    _call("vb_transpose_f32", _ptr(x), R, Cc, x.stride(-2), x.stride(0) if x.dim() == 3 else 0, _ptr(out), Rp, Cc * Rp, G)
    return out


def wave_prepare(wav, out_fp16, *, pad: int, plane: int) -> None:
    B, n = wav.shape
    _call("vb_wave_prepare", _ptr(wav), B, n, pad, plane, _ptr(out_fp16))


def loudness_pcm16(wav, out_i16, *, target_db: float) -> None:
    B, n = wav.shape
    scratch = torch.empty(B * 64, device=wav.device, dtype=torch.float32)
    _call("vb_loudness_pcm16", _ptr(wav), B, n, target_db, _ptr(scratch), _ptr(out_i16))


def mean_tokens(x, out) -> None:
    B, L, D = x.shape
    _call("vb_mean_tokens", _ptr(x), B, L, D, _ptr(out))


def attn_fused(q, q_strides, *, B, T, H, hd, scale, k_self=None, vt_self=None, Tp=0, k_text=None, vt_text=None,
               L=0, Lp=0, gate=None, out_fp16=None, out_f32=None) -> None:
    """Fused tcgen05 attention (self segment and/or one short "text" segment), see vb_attn_fused."""
    out_fp16, out_f32 = _route(out_fp16, out_f32)
    a = AttnArgs()
    a.q = _ptr(q)
    a.q_st, a.q_sh, a.q_sb = q_strides
    a.k_self, a.vt_self, a.k_text, a.vt_text = _ptr(k_self), _ptr(vt_self), _ptr(k_text), _ptr(vt_text)
    a.gate, a.out_fp16, a.out_f32 = _ptr(gate), _ptr(out_fp16), _ptr(out_f32)
    a.B, a.T, a.H, a.hd, a.Tp, a.L, a.Lp, a.scale = B, T, H, hd, Tp, L, Lp, scale
    _lib.check(_lib.lib().vb_attn_fused(C.byref(a), _stream()), "vb_attn_fused")


# ---- debug timeline: ops.TIMELINE = (int64 cuda tensor, [names]) enables a timestamp after each op
TIMELINE = None


def mark(name: str) -> None:
    if TIMELINE is None:
        return
    buf, names = TIMELINE
    i = len(names)
    if i < buf.numel():
        names.append(name)
        _lib.check(_lib.lib().vb_timestamp(C.c_void_p(buf.data_ptr() + 8 * i), _stream()), "vb_timestamp")


def router_noise(seed: int, offset: int, N: int, E: int, depth: int, steps: int, e0, e1, e2) -> None:
    _call("vb_router_noise", seed, offset, N, E, depth, steps, _ptr(e0), _ptr(e1), _ptr(e2))
