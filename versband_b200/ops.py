"""Thin Python wrappers over the C-ABI ops (one call = one kernel launch on torch's current
stream). No arithmetic happens here: these functions only describe tensors to the library."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GemmArgs

ACT_NONE, ACT_SILU, ACT_LRELU, ACT_GELU, ACT_TANH = 0, 1, 2, 3, 4
EPI_GENERIC, EPI_QKV_ROPE = 0, 1


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
            strides[i] = max(8, (ext + 7) // 8 * 8)
        ext = max(ext, dims[i] * max(strides[i], 1))
    return dims, strides


def _check_f32(t: Optional[torch.Tensor], name: str):
    if t is not None:
        assert t.dtype == torch.float32 and t.is_cuda, name


def gemm(a: torch.Tensor, b: torch.Tensor, *, M: Optional[int] = None, N: Optional[int] = None,
         K: Optional[int] = None, G1: int = 1, G2: int = 1, kpt: int = 0, dil: int = 0, pad: int = 0,
         a_c2=(0, 0), a_c3=(0, 0), b_c2=(0, 0, 0), bn: int = 0,
         tile_group: Optional[torch.Tensor] = None, m_tiles_used: Optional[torch.Tensor] = None,
         mode: int = EPI_GENERIC, alpha: float = 1.0, swiglu: bool = False,
         bias: Optional[torch.Tensor] = None, bias_g1: int = 0, bias_tile: int = 0,
         act: int = ACT_NONE, act_slope: float = 0.0,
         row_scale: Optional[torch.Tensor] = None,
         gate: Optional[torch.Tensor] = None, gate_rows: int = 1, gate_ld: int = 0,
         row_map: Optional[torch.Tensor] = None, row_mul: int = 1, row_off: int = 0,
         row_off_g1: int = 0, rows_out: Optional[int] = None,
         residual: Optional[torch.Tensor] = None, accumulate: bool = False,
         out_f32: Optional[torch.Tensor] = None, out_ld: int = 0, out_g=(0, 0),
         out_bf16: Optional[torch.Tensor] = None, outb_ld: int = 0, outb_g=(0, 0),
         act_out: int = ACT_NONE, act_out_slope: float = 0.0, out_scale: float = 1.0,
         rope_cs: Optional[torch.Tensor] = None, T: int = 0, heads: int = 0, head_dim: int = 0,
         Tp: int = 0, q_out=None, k_out=None, vt_out=None) -> None:
    """D = epilogue(A @ B^T) on the tcgen05 GEMM (see include/versband.h: vb_gemm).

    `a` is a bf16 view of up to 4 dims [c3, c2, rows, k]; `b` a bf16 view [group, n, k].
    """
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_cuda and b.is_cuda
    args = GemmArgs()
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
    args.K = (kk + 63) // 64 * 64
    args.G1, args.G2 = G1, G2
    args.kpt, args.dil, args.pad = kpt, dil, pad
    args.a_c2_g1, args.a_c2_g2 = a_c2
    args.a_c3_g1, args.a_c3_g2 = a_c3
    args.b_c2_g1, args.b_c2_g2, args.b_c2_tile = b_c2
    args.bn = bn
    args.tile_group, args.m_tiles_used = _ptr(tile_group), _ptr(m_tiles_used)
    args.mode, args.alpha, args.swiglu = mode, alpha, int(swiglu)
    _check_f32(bias, "bias"); _check_f32(row_scale, "row_scale"); _check_f32(gate, "gate")
    _check_f32(residual, "residual"); _check_f32(out_f32, "out_f32")
    args.bias, args.bias_g1, args.bias_tile = _ptr(bias), bias_g1, bias_tile
    args.act, args.act_slope = act, act_slope
    args.row_scale = _ptr(row_scale)
    args.gate, args.gate_rows, args.gate_ld = _ptr(gate), gate_rows, gate_ld
    args.row_map = _ptr(row_map)
    args.row_mul, args.row_off, args.row_off_g1 = row_mul, row_off, row_off_g1
    args.rows_out = args.M if rows_out is None else rows_out
    args.residual, args.accumulate = _ptr(residual), int(accumulate)
    args.out_f32, args.out_ld = _ptr(out_f32), out_ld
    args.out_g1, args.out_g2 = out_g
    if out_bf16 is not None:
        assert out_bf16.dtype == torch.bfloat16
    args.out_bf16, args.outb_ld = _ptr(out_bf16), outb_ld
    args.outb_g1, args.outb_g2 = outb_g
    args.act_out, args.act_out_slope, args.out_scale = act_out, act_out_slope, out_scale
    _check_f32(rope_cs, "rope_cs")
    args.rope_cs, args.T, args.heads, args.head_dim, args.Tp = _ptr(rope_cs), T, heads, head_dim, Tp
    args.q_out, args.k_out, args.vt_out = _ptr(q_out), _ptr(k_out), _ptr(vt_out)
    _lib.check(_lib.lib().vb_gemm(C.byref(args), _stream()), "vb_gemm")
