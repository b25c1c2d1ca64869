"""ctypes binding of libversband_sm100.so — the C-ABI declared in include/versband.h.

The library is the product path: if it is missing or the device is not sm_100 we raise; there
is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libversband_sm100.so"

c_i32, c_i64, c_f32, c_ptr = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class GemmArgs(C.Structure):
    """Mirror of `struct vb_gemm_args` (include/versband.h)."""

    _fields_ = [
        ("a", c_ptr), ("a_dims", c_i64 * 4), ("a_strides", c_i64 * 4),
        ("b", c_ptr), ("b_dims", c_i64 * 3), ("b_strides", c_i64 * 3),
        ("M", c_i32), ("N", c_i32), ("K", c_i32), ("G1", c_i32), ("G2", c_i32),
        ("kpt", c_i32), ("dil", c_i32), ("pad", c_i32),
        ("a_c2_g1", c_i32), ("a_c2_g2", c_i32), ("a_c3_g1", c_i32), ("a_c3_g2", c_i32),
        ("b_c2_g1", c_i32), ("b_c2_g2", c_i32), ("b_c2_tile", c_i32),
        ("bn", c_i32), ("operand_f32", c_i32),
        ("tile_group", c_ptr), ("m_tiles_used", c_ptr), ("tile_row0", c_ptr),
        ("mode", c_i32), ("alpha", c_f32), ("swiglu", c_i32),
        ("bias", c_ptr), ("bias_g1", c_i64), ("bias_tile", c_i64),
        ("act", c_i32), ("act_slope", c_f32),
        ("row_scale", c_ptr), ("gate", c_ptr), ("gate_rows", c_i32), ("gate_ld", c_i64), ("gate_g1", c_i64),
        ("row_map", c_ptr),
        ("row_mul", c_i32), ("row_off", c_i32), ("row_off_g1", c_i32), ("rows_out", c_i32),
        ("residual", c_ptr), ("res16", c_ptr), ("res16_neg", c_f32), ("accumulate", c_i32),
        ("out_f32", c_ptr), ("out_ld", c_i64), ("out_g1", c_i64), ("out_g2", c_i64),
        ("out_fp16", c_ptr), ("outb_f32", c_i32), ("outb_ld", c_i64), ("outb_g1", c_i64), ("outb_g2", c_i64),
        ("act_out", c_i32), ("act_out_slope", c_f32), ("out_scale", c_f32),
        ("rope_cs", c_ptr), ("T", c_i32), ("heads", c_i32), ("head_dim", c_i32), ("Tp", c_i32), ("qkv_sec0", c_i32),
        ("q_out", c_ptr), ("k_out", c_ptr), ("vt_out", c_ptr),
        ("route_vg", c_ptr), ("route_vg_g1", c_i64),
    ]


class AttnArgs(C.Structure):
    """Mirror of `struct vb_attn_args`."""

    _fields_ = [
        ("q", c_ptr), ("q_st", c_i64), ("q_sh", c_i64), ("q_sb", c_i64),
        ("k_self", c_ptr), ("vt_self", c_ptr), ("k_text", c_ptr), ("vt_text", c_ptr),
        ("gate", c_ptr), ("out_fp16", c_ptr), ("out_f32", c_ptr),
        ("B", c_i32), ("T", c_i32), ("H", c_i32), ("hd", c_i32), ("Tp", c_i32), ("L", c_i32), ("Lp", c_i32),
        ("scale", c_f32),
    ]


VOC_MAX_STAGES = 8


class VocDesc(C.Structure):
    """Mirror of `struct vb_voc_desc`."""

    _fields_ = [
        ("n_mel", c_i32), ("c0", c_i32), ("pre_k", c_i32), ("n_stages", c_i32), ("n_kernels", c_i32), ("post_k", c_i32),
        ("lrelu", c_f32), ("post_b", c_f32),
        ("rates", c_i32 * VOC_MAX_STAGES), ("up_k", c_i32 * VOC_MAX_STAGES), ("up_cout", c_i32 * VOC_MAX_STAGES),
        ("rb_k", c_i32 * 3), ("n_dil", c_i32 * 3), ("rb_d", (c_i32 * 4) * 3),
        ("pre_w", c_ptr), ("pre_b", c_ptr),
        ("up_w", c_ptr * VOC_MAX_STAGES), ("up_b", c_ptr * VOC_MAX_STAGES),
        ("rb_w1", ((c_ptr * 4) * 3) * VOC_MAX_STAGES), ("rb_w2", ((c_ptr * 4) * 3) * VOC_MAX_STAGES),
        ("rb_b1", ((c_ptr * 4) * 3) * VOC_MAX_STAGES), ("rb_b2", ((c_ptr * 4) * 3) * VOC_MAX_STAGES),
        ("post_w", c_ptr),
    ]


VAE_MAX_OPS = 48
VAE_RES, VAE_ATTN, VAE_UP, VAE_DOWN = 0, 1, 2, 3


class VaeOp(C.Structure):
    """Mirror of `struct vb_vae_op`."""

    _fields_ = [("kind", c_i32), ("cin", c_i32), ("cout", c_i32), ("k1", c_i32), ("k2", c_i32), ("ks", c_i32),
                ("n1_w", c_ptr), ("n1_b", c_ptr), ("n2_w", c_ptr), ("n2_b", c_ptr),
                ("w1", c_ptr), ("w2", c_ptr), ("ws", c_ptr), ("b1", c_ptr), ("b2", c_ptr), ("bs", c_ptr)]


class VaeDesc(C.Structure):
    """Mirror of `struct vb_vae_desc`."""

    _fields_ = [("zc", c_i32), ("zq_c", c_i32), ("c0", c_i32), ("out_ch", c_i32), ("groups", c_i32), ("n_ops", c_i32),
                ("pq_k", c_i32), ("in_k", c_i32), ("out_k", c_i32), ("eps", c_f32),
                ("pq_w", c_ptr), ("pq_b", c_ptr), ("in_w", c_ptr), ("in_b", c_ptr), ("no_w", c_ptr), ("no_b", c_ptr),
                ("out_w", c_ptr), ("out_b", c_ptr), ("ops", VaeOp * VAE_MAX_OPS)]


DIT_MAX_DEPTH = 16


class DitBlock(C.Structure):
    """Mirror of `struct vb_dit_block`."""

    _fields_ = [(n, c_ptr) for n in ("wqkv", "wo", "gate", "attn_norm", "ffn_norm", "capg_b", "w13", "w2", "f13", "f2")]


class DitDesc(C.Structure):
    """Mirror of `struct vb_dit_desc`."""

    _fields_ = [("C", c_i32), ("D", c_i32), ("depth", c_i32), ("H", c_i32), ("E", c_i32), ("Fh", c_i32),
                ("max_len", c_i32), ("proj_in_k", c_i32), ("operand_f32", c_i32), ("eps", c_f32), ("temperature", c_f32),
                ("mod_ld", c_i64)] + [(n, c_ptr) for n in (
                    "proj_in_w", "proj_in_b", "t_w0", "t_w2", "ada_w", "t_b0", "t_b2", "ada_b", "t_freqs", "hl_w", "hl_b",
                    "final_w", "final_b", "rope")] + [("blocks", DitBlock * DIT_MAX_DEPTH)]


class DitCond(C.Structure):
    """Mirror of `struct vb_dit_cond`."""

    _fields_ = [("B", c_i32), ("T", c_i32), ("L", c_i32), ("Lp", c_i32), ("acoustic", c_ptr), ("ac_logits", c_ptr),
                ("cap_emb", c_ptr)] + [(n, c_ptr * DIT_MAX_DEPTH) for n in ("k_y", "vt_y", "wk_m", "bs_m", "vg_m")]


class DitPrepBlock(C.Structure):
    """Mirror of `struct vb_dit_prep_block`."""

    _fields_ = [(n, c_ptr) for n in ("y_norm", "wkvy", "mkv_w", "mkv_b", "mq_wT", "mq_bs", "capg_w")]


class DitPrepDesc(C.Structure):
    """Mirror of `struct vb_dit_prep_desc`."""

    _fields_ = [("ori_dim", c_i32), ("stem_k", c_i32), ("Vm", c_i32), ("Vb", c_i32)] + [(n, c_ptr) for n in (
        "midi_tab", "beats_tab", "midi_proj_b", "beats_proj_b", "final_proj_w3", "final_proj_b", "acg_w", "acg_b",
        "c_w0", "c_w2", "cap_w", "c_b0", "c_b2", "c_ln_w", "c_ln_b", "cap_ln_w", "cap_ln_b", "cap_b")] + [
            ("blocks", DitPrepBlock * DIT_MAX_DEPTH)]


class DitIO(C.Structure):
    """Mirror of `struct vb_dit_io`."""

    _fields_ = [("x", c_ptr), ("Bx", c_i32), ("x_cl_ready", c_i32), ("x_cl", c_ptr), ("t", c_ptr), ("mod", c_ptr),
                ("hl", c_ptr), ("e0", c_ptr * DIT_MAX_DEPTH), ("e1", c_ptr * DIT_MAX_DEPTH), ("e2", c_ptr * DIT_MAX_DEPTH),
                ("v", c_ptr), ("out_v", c_ptr), ("lb", c_ptr)]


class SampleArgs(C.Structure):
    """Mirror of `struct vb_sample_args`."""

    _fields_ = [("n_steps", c_i32), ("Bc", c_i32), ("guidance", c_f32), ("t_all", c_ptr), ("dt", C.POINTER(c_f32)),
                ("e0", c_ptr), ("e1", c_ptr), ("e2", c_ptr), ("x", c_ptr), ("lb", c_ptr)]


class VersbandError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load the shared library once; fail loudly if it was not built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise VersbandError(
                f"{LIB_PATH} not found — run `python -m versband_b200.build` (or "
                "__graft_entry__.build()); versband_b200 has no fallback path")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.vb_last_error.restype = C.c_char_p
        _lib.vb_launch_count.restype = C.c_longlong
        _lib.vb_gemm.argtypes = [C.POINTER(GemmArgs), c_ptr]
        _lib.vb_gemm_chain.argtypes = [C.POINTER(GemmArgs), C.POINTER(c_i32), c_i32, c_i32, c_ptr, c_i64, c_ptr]
        _declare(_lib)
        _lib.vb_attn_fused.argtypes = [C.POINTER(AttnArgs), c_ptr]
        if _lib.vb_sizeof(1) != C.sizeof(AttnArgs):
            raise VersbandError("vb_attn_args ABI mismatch between versband.h and _lib.py")
        _lib.vb_voc_forward.argtypes = [C.POINTER(VocDesc), c_ptr, c_i32, c_i32, c_ptr, c_ptr, c_i64, c_ptr]
        _lib.vb_voc_workspace_bytes.argtypes = [C.POINTER(VocDesc), c_i32, c_i32]
        _lib.vb_voc_workspace_bytes.restype = c_i64
        if _lib.vb_sizeof(2) != C.sizeof(VocDesc):
            raise VersbandError("vb_voc_desc ABI mismatch between versband.h and _lib.py")
        _lib.vb_vae_decode.argtypes = [C.POINTER(VaeDesc), c_ptr, c_f32, c_i32, c_i32, c_ptr, c_ptr, c_i64, c_ptr]
        _lib.vb_vae_workspace_bytes.argtypes = [C.POINTER(VaeDesc), c_i32, c_i32]
        _lib.vb_vae_workspace_bytes.restype = c_i64
        _lib.vb_vae_encode.argtypes = [C.POINTER(VaeDesc), c_ptr, c_i32, c_i32, c_ptr, c_ptr, c_i64, c_ptr]
        _lib.vb_vae_encode_workspace_bytes.argtypes = [C.POINTER(VaeDesc), c_i32, c_i32]
        _lib.vb_vae_encode_workspace_bytes.restype = c_i64
        if _lib.vb_sizeof(3) != C.sizeof(VaeDesc):
            raise VersbandError("vb_vae_desc ABI mismatch between versband.h and _lib.py")
        _lib.vb_dit_forward.argtypes = [C.POINTER(DitDesc), C.POINTER(DitCond), C.POINTER(DitIO), c_ptr, c_i64, c_ptr]
        _lib.vb_dit_workspace_bytes.argtypes = [C.POINTER(DitDesc), c_i32, c_i32]
        _lib.vb_dit_workspace_bytes.restype = c_i64
        _lib.vb_dit_prepare.argtypes = [C.POINTER(DitDesc), C.POINTER(DitPrepDesc), c_ptr, c_ptr, c_i32, c_i32, c_ptr, c_i32,
                                        c_i32, C.POINTER(DitCond), c_ptr, c_ptr, c_i64, c_ptr]
        _lib.vb_dit_prepare_workspace_bytes.argtypes = [C.POINTER(DitDesc), C.POINTER(DitPrepDesc)] + [c_i32] * 5
        _lib.vb_dit_prepare_workspace_bytes.restype = c_i64
        _lib.vb_sample_cfg.argtypes = [C.POINTER(DitDesc), C.POINTER(DitCond), C.POINTER(SampleArgs), c_ptr, c_i64, c_ptr]
        _lib.vb_sample_workspace_bytes.argtypes = [C.POINTER(DitDesc), c_i32, c_i32, c_i32]
        _lib.vb_sample_workspace_bytes.restype = c_i64
        for which, st in ((4, DitDesc), (5, DitCond), (6, DitIO), (7, DitPrepDesc), (8, SampleArgs)):
            if _lib.vb_sizeof(which) != C.sizeof(st):
                raise VersbandError(f"{st.__name__} ABI mismatch between versband.h and _lib.py")
        if _lib.vb_sizeof(0) != C.sizeof(GemmArgs):
            raise VersbandError("vb_gemm_args ABI mismatch between versband.h and _lib.py")
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().vb_last_error().decode()
        raise VersbandError(f"{what} failed (status {rc}): {msg}")


def launch_count() -> int:
    return int(lib().vb_launch_count())


def _declare(l: C.CDLL) -> None:
    """argtypes for the plain-argument entry points (include/versband.h)."""
    P, I, L, F = c_ptr, c_i32, c_i64, c_f32
    sigs = {
        "vb_norm_mod": [I, P, P, P, P, P, L, I, F, P, P, I, I, P],
        "vb_softmax_rows": [P, L, P, L, L, I, P],
        "vb_attn_smallkv": [P, L, L, L, P, P, L, I, I, I, I, I, F, P, P, P, P, P],
        "vb_mean_tokens": [P, I, I, I, P, P],
        "vb_timestep_embed": [P, P, P, P, I, I, P],
        "vb_linear_small": [P, L, P, P, P, I, I, I, P],
        "vb_router": [P, P, P, P, P, I, P, I, P, P, P, F, I, I, I, I, P, P, P, P, P, P, P, P, P, P, I, P],
        "vb_route_prepare": [P, I, I, I, I, P, P, P, P, P],
        "vb_route_dispatch": [P, I, P, P, I, P, I, P, P, P, F, I, I, I, I, P, P, I, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P],
        "vb_moe_finalize": [P, P, I, P, P, P, P, P, I, P, P],
        "vb_moe_dispatch": [P, P, P, P, P, P, P, P, P, P, I, I, I, P, P, P, P, P, P, I, P, P],
        "vb_moe_scatter": [P, P, P, P, P, P, P, I, I, I, P, P, P, P, P],
        "vb_moe_sort": [P, P, P, P, I, I, P, P, P, P, I, P, P],
        "vb_gather_rows": [P, P, P, I, I, P, I, P],
        "vb_ew": [P, P, F, F, I, F, P, P, L, I, I, P],
        "vb_timestamp": [P, P],
        "vb_router_noise": [C.c_uint64, C.c_uint64, I, I, I, I, P, P, P, P],
        "vb_nct_to_ntc": [P, F, P, I, I, I, I, I, P],
        "vb_ntc_to_nct": [P, P, L, F, F, P, I, I, I, P, I, I, P],
        "vb_embed_rows": [P, P, I, I, P, L, I, P],
        "vb_pool2_sum": [P, P, I, I, I, I, P, P, P],
        "vb_gn_stats": [P, I, I, I, I, F, P, P, P, P],
        "vb_gn_apply": [P, P, P, P, P, I, I, I, I, I, P, P],
        "vb_gn_fused": [P, P, P, I, I, I, I, F, I, P, P],
        "vb_conv_post": [P, P, F, I, I, I, I, P, P],
        "vb_resblock_pair": [P, P, P, P, P, P, I, I, I, I, I, F, P],
        "vb_adamw": [P, P, P, P, L, F, F, F, F, F, F, F, P],
        "vb_transpose_f32": [P, L, L, L, L, P, L, L, I, P],
        "vb_wave_prepare": [P, I, L, I, L, P, P],
        "vb_loudness_pcm16": [P, I, L, F, P, P, P],
        "vb_stem_gather": [P, P, P, P, P, P, I, I, I, I, I, I, I, F, P, I, P],
        "vb_mrf_combine": [P, P, P, F, F, F, P, L, P],
    }
    for name, at in sigs.items():
        fn = getattr(l, name)
        fn.argtypes = at
        fn.restype = C.c_int


PROF_KINDS = ["gemm", "norm", "softmax", "attn_small", "router", "moe_sort", "elementwise", "layout",
              "groupnorm", "conv_post"]


def prof_enable(on: bool) -> None:
    check(lib().vb_prof_enable(int(on)), "vb_prof_enable")


def prof_report() -> dict:
    n = len(PROF_KINDS)
    ms = (C.c_double * n)()
    cnt = (C.c_longlong * n)()
    check(lib().vb_prof_report(ms, cnt, n), "vb_prof_report")
    return {k: {"ms": ms[i], "launches": int(cnt[i])} for i, k in enumerate(PROF_KINDS)}
