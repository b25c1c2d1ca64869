// Device code of vb_gemm: tcgen05 / TMEM / TMA GEMM kernels for sm_100a with implicit-conv A
// addressing and fused epilogues, templated on the epilogue class (vb_gemm_params.h). Included by the
// vb_gemm_ec*.cu instantiation units only. See include/versband.h for the contract.
//
// CTA = 320 threads: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..9 = epilogue (TMEM lane quadrant = warp_idx % 4). Persistent: one CTA per SM walks the
// tile list; two TMEM accumulators let the epilogue of tile i overlap the mainloop of tile i+1.
#pragma once
#include <math.h>
#include "vb_gemm_params.h"
#include "vb_ptx.cuh"

namespace vb {


__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define VB_TRACE(slot) do { if (p.trace) p.trace[blockIdx.x * 8 + (slot)] = gtime_ns(); } while (0)

// Activations are applied by whole-chunk loops selected ONCE per chunk (one uniform branch, then a
// straight-line 32-wide loop): keeps the executed instruction stream short — the epilogue runs per
// tile on four warps and was instruction-fetch bound when the switch sat inside the unrolled loop.
template <int ACT>
__device__ __forceinline__ float act_one(float v, float slope) {
  if (ACT == VB_ACT_SILU) return v / (1.f + __expf(-v));
  if (ACT == VB_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (ACT == VB_ACT_GELU_ERF) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  if (ACT == VB_ACT_TANH) return tanhf(v);
  if (ACT == VB_ACT_LOG10) return log10f(fmaxf(v, slope));  // slope = clip value (dynamic range compression)
  return v;
}
template <int ACT, int N>
__device__ __forceinline__ void act_loop(float (&v)[N], float slope, float pre) {
#pragma unroll
  for (int j = 0; j < N; ++j) v[j] = act_one<ACT>(v[j] * pre, slope);
}
template <int N>
__device__ __forceinline__ void apply_act_all(float (&v)[N], int act, float slope, float pre = 1.f) {
  switch (act) {
    case VB_ACT_SILU: act_loop<VB_ACT_SILU>(v, slope, pre); break;
    case VB_ACT_LRELU: act_loop<VB_ACT_LRELU>(v, slope, pre); break;
    case VB_ACT_GELU_ERF: act_loop<VB_ACT_GELU_ERF>(v, slope, pre); break;
    case VB_ACT_TANH: act_loop<VB_ACT_TANH>(v, slope, pre); break;
    case VB_ACT_LOG10: act_loop<VB_ACT_LOG10>(v, slope, pre); break;
    default:
      if (pre != 1.f) act_loop<VB_ACT_NONE>(v, slope, pre);
      break;
  }
}

__device__ __forceinline__ uint32_t pack_fp16(float a, float b) {
  __half2 t = vb_f2h2(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// ------------------------------------------------------------------ coalesced epilogue
// tcgen05.ld hands each thread one accumulator ROW (32 consecutive columns). Storing straight from
// that layout makes every warp-level 16-byte access touch 32 different cache lines. Each epilogue
// warp therefore transposes its 32x32 chunk through a private, XOR-swizzled 4 KB smem block:
// afterwards 8 lanes cover 128 contiguous bytes of one row, one warp instruction touches 4 rows x
// 128 B, and every load of the fused operands (bias, gate, residual, rope table) and every store is
// coalesced.
//   lane -> (lr = lane >> 3, cq = lane & 7); item i in 0..7 = row 4i + lr of the warp's 32 rows,
//   columns 4cq .. 4cq+3 of the chunk; v[4i + k] is component k of item i.
// The epilogue is instruction-issue bound (a few warps running a long dependent chain), so the
// fast path keeps per-chunk work minimal: row offsets are computed once per tile as 32-bit element
// offsets, alignment / raggedness is decided once per chunk (edge chunks fall back to the row-wise
// code), and invalid rows are handled by predicated stores.
struct EpiRows {
  unsigned off[8];   // generic: orow * out_ld;  QKV: (b*H*T + t) * hd         (0 for rows not stored)
  unsigned offb[8];  // generic: orow * outb_ld; QKV: t * hd (row of the rope table)
  float rs[8];       // row_scale
  unsigned valid;    // bit i: item row i is stored
  unsigned hi;       // bit i: item row i belongs to gate batch b_lo + 1
  int b_lo;          // gate batch of the lane's first row
  bool multi;        // the lane's rows span more than two gate batches -> row-wise path
};

template <int EC>
__device__ __forceinline__ void epi_prepare_rows(const GemmKParams& p, EpiRows& rows, int m_base, int g1, int lane) {
  const int m0 = m_base + (lane >> 3);
  rows.valid = 0; rows.hi = 0; rows.multi = false; rows.b_lo = 0;
  if (EC == EC_QKV) {
    const int mc = m0 < p.M ? m0 : p.M - 1;
    int b = mc / p.T, t = mc - b * p.T;
    t += m0 - mc;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      while (t >= p.T) { t -= p.T; ++b; }
      const bool ok = m0 + 4 * i < p.M;
      rows.off[i] = ok ? static_cast<unsigned>((b * p.heads * p.T + t) * p.head_dim) : 0u;
      rows.offb[i] = ok ? static_cast<unsigned>(t * p.head_dim) : 0u;
      rows.rs[i] = 1.f;
      if (ok) rows.valid |= 1u << i;
      t += 4;
    }
    return;
  }
  if (!p.row_map && !p.row_scale) {
    // affine rows (every conv, most linears): a handful of instructions per row instead of the general path
    const bool gated = EC != EC_SWIGLU && p.gate != nullptr;
    int r0 = 0;
    bool simple = true;
    if (gated) {  // gate batch of the first row; all 8 rows (span 29) must fall in at most two batches
      const int mc = m0 < p.M ? m0 : p.M - 1;
      const int b0 = mc / p.gate_rows;
      r0 = mc - b0 * p.gate_rows + (m0 - mc);
      rows.b_lo = b0;
      simple = r0 + 28 < 2 * p.gate_rows;
    }
    if (simple) {
      const int o0 = m0 * p.row_mul + p.row_off + g1 * p.row_off_g1, ostep = 4 * p.row_mul;
      const unsigned ld = static_cast<unsigned>(p.out_ld), ldb = static_cast<unsigned>(p.outb_ld);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int o = o0 + i * ostep;
        const bool ok = (m0 + 4 * i < p.M) && (static_cast<unsigned>(o) < static_cast<unsigned>(p.rows_out));
        if (ok) rows.valid |= 1u << i;
        if (gated && ok && r0 + 4 * i >= p.gate_rows) rows.hi |= 1u << i;
        rows.off[i] = ok ? static_cast<unsigned>(o) * ld : 0u;
        rows.offb[i] = ok ? static_cast<unsigned>(o) * ldb : 0u;
        rows.rs[i] = 1.f;
      }
      return;
    }
    rows.b_lo = 0;
  }
  int b = 0, r = 0;
  if (EC != EC_SWIGLU && p.gate) {
    const int mc = m0 < p.M ? m0 : p.M - 1;
    b = mc / p.gate_rows;
    r = mc - b * p.gate_rows + (m0 - mc);
    rows.b_lo = b;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + 4 * i;
    int o = -1;
    float rs = 1.f;
    if (m < p.M) {
      o = p.row_map ? p.row_map[m] : m * p.row_mul + p.row_off + g1 * p.row_off_g1;
      if (o >= p.rows_out) o = -1;
      if (p.row_scale) rs = p.row_scale[m];
      if (EC != EC_SWIGLU && p.gate) {
        while (r >= p.gate_rows) { r -= p.gate_rows; ++b; }
        if (b - rows.b_lo == 1) rows.hi |= 1u << i;
        if (b - rows.b_lo > 1) rows.multi = true;
        r += 4;
      }
    }
    if (o >= 0) rows.valid |= 1u << i;
    rows.off[i] = o >= 0 ? static_cast<unsigned>(o) * static_cast<unsigned>(p.out_ld) : 0u;
    rows.offb[i] = o >= 0 ? static_cast<unsigned>(o) * static_cast<unsigned>(p.outb_ld) : 0u;
    rows.rs[i] = rs;
  }
}

// thread `lane` owns accumulator row `lane` of the warp: write it, read back the item layout
__device__ __forceinline__ void epi_transpose(const uint32_t (&raw)[32], uint32_t stg, int lane, float (&v)[32]) {
  __syncwarp();  // previous chunk's reads are done
#pragma unroll
  for (int j = 0; j < 8; ++j)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 128 + ((j ^ (lane & 7)) << 4)),
                 "r"(raw[4 * j]), "r"(raw[4 * j + 1]), "r"(raw[4 * j + 2]), "r"(raw[4 * j + 3])
                 : "memory");
  __syncwarp();
  const int lr = lane >> 3, cq = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = 4 * i + lr;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[4 * i]), "=f"(v[4 * i + 1]), "=f"(v[4 * i + 2]), "=f"(v[4 * i + 3])
                 : "r"(stg + row * 128 + ((cq ^ (row & 7)) << 4))
                 : "memory");
  }
}

__device__ __forceinline__ float silu_mul(float a, float b) { return __fdividef(a, 1.f + __expf(-a)) * b; }
// gated-GELU of T5 v1.1 / FLAN-T5 ("gelu_new": tanh approximation) x gate: DenseGatedActDense
__device__ __forceinline__ float gelu_new_mul(float a, float b) {
  const float u = 0.7978845608028654f * (a + 0.044715f * a * a * a);
  return 0.5f * a * (1.f + tanhf(u)) * b;
}
// kind 1 = SwiGLU (silu), 2 = GEGLU (gelu_new), 3 = magnitude of a (re, im) pair: sqrt(re^2 + im^2 + 1e-9) (STFT, NAT_mel.py:75)
__device__ __forceinline__ float glu_mul(float a, float b, int kind) {
  if (kind == 3) return sqrtf(fmaf(a, a, fmaf(b, b, 1e-9f)));
  return kind == 2 ? gelu_new_mul(a, b) : silu_mul(a, b);
}

// v[j] (op)= src[j] for the 32 columns of this chunk; vector path when the chunk is full & aligned
template <bool MUL>
__device__ __forceinline__ void combine32(float (&v)[32], const float* __restrict__ src, bool vec,
                                          int n_left) {
  if (vec) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(src) + j);
      if (MUL) { v[4 * j] *= r.x; v[4 * j + 1] *= r.y; v[4 * j + 2] *= r.z; v[4 * j + 3] *= r.w; }
      else { v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w; }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < n_left) { if (MUL) v[j] *= src[j]; else v[j] += src[j]; }
  }
}

// ------------------------------------------------------------------ row-wise (edge) epilogue
// One thread owns one output row (TMEM lane) and 32 consecutive columns of it: handles ragged N,
// unaligned pointers and every fusion; only used for chunks the coalesced path below rejects.
__device__ __forceinline__ void epilogue_rowwise(const GemmKParams& p, float (&v)[32], int m,
                                                 int n_base, int g1, int g2, int tg,
                                                 long long orow) {
  const int N = p.N;
  if (p.alpha != 1.f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
  }

  if (p.swiglu) {
    // columns (2i, 2i+1) = (w1 x, w3 x)  ->  silu(w1 x) * (w3 x), 16 outputs
    const int n_out0 = n_base >> 1;
    const int n_out_total = N >> 1;
    const float rs = p.row_scale ? p.row_scale[m] : 1.f;
    float o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      o[j] = glu_mul(v[2 * j], v[2 * j + 1], p.swiglu) * rs;
    }
    if (p.out_fp16 && p.outb_f32) {
      float* dstf = reinterpret_cast<float*>(p.out_fp16) + g1 * p.outb_g1 + g2 * p.outb_g2 + orow * p.outb_ld + n_out0;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n_out0 + j < n_out_total) dstf[j] = o[j];
    } else if (p.out_fp16) {
      __half* dst = p.out_fp16 + g1 * p.outb_g1 + g2 * p.outb_g2 + orow * p.outb_ld + n_out0;
      if (p.vec_fp16 && n_out0 + 16 <= n_out_total) {
        uint4 w0, w1;
        w0.x = pack_fp16(o[0], o[1]);   w0.y = pack_fp16(o[2], o[3]);
        w0.z = pack_fp16(o[4], o[5]);   w0.w = pack_fp16(o[6], o[7]);
        w1.x = pack_fp16(o[8], o[9]);   w1.y = pack_fp16(o[10], o[11]);
        w1.z = pack_fp16(o[12], o[13]); w1.w = pack_fp16(o[14], o[15]);
        reinterpret_cast<uint4*>(dst)[0] = w0;
        reinterpret_cast<uint4*>(dst)[1] = w1;
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n_out0 + j < n_out_total) dst[j] = vb_f2h(o[j]);
      }
    }
    if (p.out_f32) {
      float* dst = p.out_f32 + g1 * p.out_g1 + g2 * p.out_g2 + orow * p.out_ld + n_out0;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n_out0 + j < n_out_total) dst[j] = o[j];
    }
    return;
  }

  const bool full = (n_base + 32 <= N);
  const int n_left = N - n_base;
  if (p.bias) combine32<false>(v, p.bias + g1 * p.bias_g1 + tg * p.bias_tile + n_base, full && p.vec_aux, n_left);
  if (p.act != VB_ACT_NONE) apply_act_all(v, p.act, p.act_slope);
  if (p.row_scale) {
    const float rs = p.row_scale[m];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= rs;
  }
  if (p.gate)
    combine32<true>(v, p.gate + (long long)(m / p.gate_rows) * p.gate_ld + g1 * p.gate_g1 + n_base,
                    full && p.vec_aux, n_left);
  const long long off32 = g1 * p.out_g1 + g2 * p.out_g2 + orow * p.out_ld + n_base;
  if (p.residual) combine32<false>(v, p.residual + off32, full && p.vec_f32, n_left);
  if (p.res16) {
    const __half* r16 = p.res16 + g1 * p.outb_g1 + g2 * p.outb_g2 + orow * p.outb_ld + n_base;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < n_left) { const float y = __half2float(r16[j]); v[j] += y < 0.f ? y * p.res16_neg : y; }
  }
  if (p.out_f32) {
    float* dst = p.out_f32 + off32;
    if (p.accumulate) combine32<false>(v, dst, full && p.vec_f32, n_left);
    if (full && p.vec_f32) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(dst)[j] =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < n_left) dst[j] = v[j];
    }
  }
  if (p.out_fp16) {
    const long long offb = g1 * p.outb_g1 + g2 * p.outb_g2 + orow * p.outb_ld + n_base;
    apply_act_all(v, p.act_out, p.act_out_slope, p.out_scale);
    if (p.outb_f32) {
      float* dstf = reinterpret_cast<float*>(p.out_fp16) + offb;
      if (full && p.vec_fp16) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          reinterpret_cast<float4*>(dstf)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < n_left) dstf[j] = v[j];
      }
      return;
    }
    __half* dst = p.out_fp16 + offb;
    if (full && p.vec_fp16) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 w;
        w.x = pack_fp16(v[8 * j], v[8 * j + 1]);
        w.y = pack_fp16(v[8 * j + 2], v[8 * j + 3]);
        w.z = pack_fp16(v[8 * j + 4], v[8 * j + 5]);
        w.w = pack_fp16(v[8 * j + 6], v[8 * j + 7]);
        reinterpret_cast<uint4*>(dst)[j] = w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < n_left) dst[j] = vb_f2h(v[j]);
    }
  }
}

struct Raw32 { uint32_t r[32]; };
// Out of line on purpose: keeps the (rarely executed) edge code out of the hot instruction stream.
static __device__ __noinline__ void epilogue_rowwise_chunk(const GemmKParams& p, Raw32 raw, int m, int n_base, int g1,
                                                    int g2, int tg) {
  if (m >= p.M) return;
  const long long orow = p.row_map ? (long long)p.row_map[m]
                                   : (long long)m * p.row_mul + p.row_off + (long long)g1 * p.row_off_g1;
  if (orow < 0 || orow >= p.rows_out) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw.r[j]);
  epilogue_rowwise(p, v, m, n_base, g1, g2, tg, orow);
}

// Global operands of one chunk's epilogue, requested ahead of use (together with the chunk's TMEM
// load): all loads are unconditional (rows that are not stored read row 0) so that they are in
// flight together.
struct EpiOperands {
  float4 res[8];      // generic: residual (or the accumulate target) per item; QKV: rope (cos,sin) x2 per item
  float4 bias;        // per column quad
  float4 g_lo, g_hi;  // gate rows of the two batches this lane's rows fall in
};

template <int EC>
__device__ __forceinline__ void epi_load_operands(const GemmKParams& p, const EpiRows& rows, EpiOperands& op,
                                                  int n_base, int g1, int g2, int tg, int lane) {
  const int n = n_base + 4 * (lane & 7);
  if (EC == EC_QKV) {
    const int hd = p.head_dim;
    if (n_base / (p.heads * hd) + p.qkv_sec0 == 2) return;
    op.bias = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.rope_cs) {
      const float* cs = p.rope_cs + (n_base % hd) + 4 * (lane & 7);  // 2 floats per pair: column d <-> floats d, d+1
#pragma unroll
      for (int i = 0; i < 8; ++i) op.res[i] = __ldg(reinterpret_cast<const float4*>(cs + rows.offb[i]));
    }
    return;
  }
  if (EC == EC_SWIGLU || ((EC == EC_FULL || EC == EC_CHAIN) && p.swiglu)) return;
  if (p.bias) op.bias = __ldg(reinterpret_cast<const float4*>(p.bias + g1 * p.bias_g1 + tg * p.bias_tile + n));
  if (p.gate) {
    const float* gsrc = p.gate + g1 * p.gate_g1 + n + (long long)rows.b_lo * p.gate_ld;
    op.g_lo = __ldg(reinterpret_cast<const float4*>(gsrc));
    op.g_hi = __ldg(reinterpret_cast<const float4*>(gsrc + (rows.hi ? p.gate_ld : 0)));
  }
  if ((EC == EC_RES16 || EC == EC_FULL) && p.res16) {
    // fp16 residual in the activated ("y = leaky_relu(x)") representation: x = y >= 0 ? y : y * res16_neg
    const __half* r16 = p.res16 + g1 * p.outb_g1 + g2 * p.outb_g2 + n;
    const float ng = p.res16_neg;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(r16 + rows.offb[i]));
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      op.res[i] = make_float4(a.x < 0.f ? a.x * ng : a.x, a.y < 0.f ? a.y * ng : a.y,
                              b.x < 0.f ? b.x * ng : b.x, b.y < 0.f ? b.y * ng : b.y);
    }
    return;
  }
  const float* rsrc = p.residual ? p.residual : ((p.accumulate && p.out_f32) ? p.out_f32 : nullptr);
  if (rsrc) {
    rsrc += g1 * p.out_g1 + g2 * p.out_g2 + n;
#pragma unroll
    for (int i = 0; i < 8; ++i) op.res[i] = __ldg(reinterpret_cast<const float4*>(rsrc + rows.off[i]));
  }
}

// ------------------------------------------------------------------ generic epilogue, coalesced fast path
// (full 32-column chunk, every pointer / stride 16-byte friendly: GemmKParams::fast_ok)
template <int EC>
__device__ __forceinline__ void epilogue_generic(const GemmKParams& p, float (&v)[32], const EpiRows& rows,
                                                 const EpiOperands& op, int n_base, int g1, int g2, int tg, int lane) {
  const int n = n_base + 4 * (lane & 7);  // first of this lane's 4 columns
  if (EC == EC_FULL && p.alpha != 1.f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
  }

  if (EC == EC_SWIGLU || ((EC == EC_FULL || EC == EC_CHAIN) && p.swiglu)) {
    // columns (2i, 2i+1) = (w1 x, w3 x)  ->  silu(w1 x) * (w3 x): two outputs per quad
    const int no = n >> 1;
    float o[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // the DiT's hot classes (SWIGLU, CHAIN) carry the SiLU gate only; GEGLU / magnitude live in EC_FULL
      if (EC == EC_FULL) {
        o[2 * i] = glu_mul(v[4 * i], v[4 * i + 1], p.swiglu) * rows.rs[i];
        o[2 * i + 1] = glu_mul(v[4 * i + 2], v[4 * i + 3], p.swiglu) * rows.rs[i];
      } else {
        o[2 * i] = silu_mul(v[4 * i], v[4 * i + 1]) * rows.rs[i];
        o[2 * i + 1] = silu_mul(v[4 * i + 2], v[4 * i + 3]) * rows.rs[i];
      }
    }
    if (p.out_fp16 && p.outb_f32) {
      float* dst = reinterpret_cast<float*>(p.out_fp16) + g1 * p.outb_g1 + g2 * p.outb_g2 + no;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (rows.valid >> i & 1) *reinterpret_cast<float2*>(dst + rows.offb[i]) = make_float2(o[2 * i], o[2 * i + 1]);
    } else if (p.out_fp16) {
      __half* dst = p.out_fp16 + g1 * p.outb_g1 + g2 * p.outb_g2 + no;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (rows.valid >> i & 1) *reinterpret_cast<uint32_t*>(dst + rows.offb[i]) = pack_fp16(o[2 * i], o[2 * i + 1]);
    }
    if (EC == EC_FULL && p.out_f32) {
      float* dst = p.out_f32 + g1 * p.out_g1 + g2 * p.out_g2 + no;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (rows.valid >> i & 1) *reinterpret_cast<float2*>(dst + rows.off[i]) = make_float2(o[2 * i], o[2 * i + 1]);
    }
    return;
  }
  if (EC == EC_SWIGLU) return;

  if (p.bias) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[4 * i] += op.bias.x; v[4 * i + 1] += op.bias.y; v[4 * i + 2] += op.bias.z; v[4 * i + 3] += op.bias.w; }
  }
  if (EC == EC_FULL) {
    if (p.act != VB_ACT_NONE) apply_act_all(v, p.act, p.act_slope);
  } else if (p.act == VB_ACT_LRELU) {
    act_loop<VB_ACT_LRELU>(v, p.act_slope, 1.f);
  }
  if (p.row_scale) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[4 * i] *= rows.rs[i]; v[4 * i + 1] *= rows.rs[i]; v[4 * i + 2] *= rows.rs[i]; v[4 * i + 3] *= rows.rs[i]; }
  }
  if (p.gate) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 g4 = (rows.hi >> i & 1) ? op.g_hi : op.g_lo;
      v[4 * i] *= g4.x; v[4 * i + 1] *= g4.y; v[4 * i + 2] *= g4.z; v[4 * i + 3] *= g4.w;
    }
  }
  const long long gofs32 = g1 * p.out_g1 + g2 * p.out_g2 + n;
  if (p.residual || ((EC == EC_RES16 || EC == EC_FULL) && p.res16) || (p.accumulate && p.out_f32)) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[4 * i] += op.res[i].x; v[4 * i + 1] += op.res[i].y; v[4 * i + 2] += op.res[i].z; v[4 * i + 3] += op.res[i].w; }
    if ((p.residual || ((EC == EC_RES16 || EC == EC_FULL) && p.res16)) && p.accumulate && p.out_f32) {  // both: the accumulate target is read here
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 r4 = *reinterpret_cast<const float4*>(p.out_f32 + gofs32 + rows.off[i]);
        v[4 * i] += r4.x; v[4 * i + 1] += r4.y; v[4 * i + 2] += r4.z; v[4 * i + 3] += r4.w;
      }
    }
  }
  if (p.out_f32) {
    float* dst = p.out_f32 + gofs32;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (rows.valid >> i & 1)
        *reinterpret_cast<float4*>(dst + rows.off[i]) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  if (p.out_fp16) {
    const long long gofsb = g1 * p.outb_g1 + g2 * p.outb_g2 + n;
    if (EC == EC_FULL) {
      apply_act_all(v, p.act_out, p.act_out_slope, p.out_scale);
    } else if (p.act_out == VB_ACT_LRELU) {
      act_loop<VB_ACT_LRELU>(v, p.act_out_slope, p.out_scale);
    } else if (p.out_scale != 1.f) {
      act_loop<VB_ACT_NONE>(v, 0.f, p.out_scale);
    }
    if (p.outb_f32) {
      float* dst = reinterpret_cast<float*>(p.out_fp16) + gofsb;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (rows.valid >> i & 1)
          *reinterpret_cast<float4*>(dst + rows.offb[i]) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
      __half* dst = p.out_fp16 + gofsb;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (rows.valid >> i & 1)
          *reinterpret_cast<uint2*>(dst + rows.offb[i]) =
              make_uint2(pack_fp16(v[4 * i], v[4 * i + 1]), pack_fp16(v[4 * i + 2], v[4 * i + 3]));
    }
  }
}

// ------------------------------------------------------------------ QKV + RoPE epilogue
// Row m = b*T + t, column n in [0, 3*H*hd): section = n / (H*hd) (q,k,v), head, dim.
// Rotary: pairs (2i, 2i+1) rotated by angle(t, i) — flag_large_dit_moe.py:263-269.
// A 32-column chunk never straddles a head or section because head_dim % 32 == 0.
// q / k go through the transposed (coalesced) layout; V^T[b][h][d][t] is written straight from the
// TMEM layout (lanes = consecutive t -> coalesced 2-byte stores).
__device__ __forceinline__ void epilogue_qkv(const GemmKParams& p, const uint32_t (&raw)[32], uint32_t stg,
                                             const EpiRows& rows, const EpiOperands& op, int m_base, int n_base,
                                             int lane) {
  const int hd = p.head_dim, H = p.heads, T = p.T;
  const int D = H * hd;
  const int sec0 = n_base / D;
  const int rem = n_base - sec0 * D;
  const int sec = sec0 + p.qkv_sec0;
  const int h = rem / hd;
  const int d0 = rem - h * hd;
  if (sec == 2) {
    const int m = m_base + lane;
    if (m >= p.M) return;
    const int b = m / T;
    const int t = m - b * T;
    __half* dst = p.vt_out + (((long long)b * H + h) * hd + d0) * p.Tp + t;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float x = __uint_as_float(raw[j]) * p.alpha;
      if (p.bias) x += __ldg(p.bias + n_base + j);
      dst[(long long)j * p.Tp] = vb_f2h(x);
    }
    return;
  }
  float v[32];
  epi_transpose(raw, stg, lane, v);
  __half* out = (sec == 0 ? p.q_out : p.k_out) + (long long)h * T * hd + d0 + 4 * (lane & 7);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float x0 = v[4 * i] * p.alpha + op.bias.x, x1 = v[4 * i + 1] * p.alpha + op.bias.y;
    float x2 = v[4 * i + 2] * p.alpha + op.bias.z, x3 = v[4 * i + 3] * p.alpha + op.bias.w;
    if (p.rope_cs) {
      const float4 cs = op.res[i];
      const float y0 = x0 * cs.x - x1 * cs.y, y1 = x0 * cs.y + x1 * cs.x;
      const float y2 = x2 * cs.z - x3 * cs.w, y3 = x2 * cs.w + x3 * cs.z;
      x0 = y0; x1 = y1; x2 = y2; x3 = y3;
    }
    if (rows.valid >> i & 1)
      *reinterpret_cast<uint2*>(out + rows.off[i]) = make_uint2(pack_fp16(x0, x1), pack_fp16(x2, x3));
  }
}

// ------------------------------------------------------------------ ROUTE epilogue (BN = 96 = one head)
// Thread = accumulator row (token). Online softmax over the head's 96 score columns (3 TMEM chunks),
// contracted on the fly with route_vg[batch][head][j][0..3]; one float4 per (token, head).
__device__ __forceinline__ void epilogue_route_tile(const GemmKParams& p, uint32_t tmem_d, int m, int n_tile, int g1) {
  const float4* bias4 = reinterpret_cast<const float4*>(p.bias + g1 * p.bias_g1 + n_tile * 96);
  const float4* vg = reinterpret_cast<const float4*>(p.route_vg + g1 * p.route_vg_g1) + n_tile * 96;
  float mx = -1e30f, sum = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int c = 0; c < 96; c += 32) {
    uint32_t raw[32];
    tmem_ld32(tmem_d + c, raw);
    tmem_ld_wait();
    float v[32];
    float cmax = -1e30f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b4 = __ldg(bias4 + (c >> 2) + j);
      v[4 * j] = __uint_as_float(raw[4 * j]) + b4.x;
      v[4 * j + 1] = __uint_as_float(raw[4 * j + 1]) + b4.y;
      v[4 * j + 2] = __uint_as_float(raw[4 * j + 2]) + b4.z;
      v[4 * j + 3] = __uint_as_float(raw[4 * j + 3]) + b4.w;
      cmax = fmaxf(cmax, fmaxf(fmaxf(v[4 * j], v[4 * j + 1]), fmaxf(v[4 * j + 2], v[4 * j + 3])));
    }
    const float m_new = fmaxf(mx, cmax);
    const float resc = __expf(mx - m_new);
    sum *= resc; acc.x *= resc; acc.y *= resc; acc.z *= resc; acc.w *= resc;
    mx = m_new;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float pj = __expf(v[j] - mx);
      const float4 g = __ldg(vg + c + j);
      sum += pj;
      acc.x = fmaf(pj, g.x, acc.x); acc.y = fmaf(pj, g.y, acc.y);
      acc.z = fmaf(pj, g.z, acc.z); acc.w = fmaf(pj, g.w, acc.w);
    }
  }
  if (m < p.M) {
    const float inv = 1.f / sum;
    *reinterpret_cast<float4*>(p.out_f32 + ((long long)g1 * p.M + m) * p.out_ld + n_tile * 4) =
        make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
}

// The chunk loop shared by the three kernels. Eight epilogue warps: two per TMEM lane quadrant, the
// pair splitting the tile's 32-column chunks even/odd (`c_first` = 0 or 32) — a single warp per
// scheduler runs the chunk body latency-bound, so warps, not bytes, set the pace.
// A chunk's global operands are requested together with its TMEM load, before either is consumed.
#define VB_EPILOGUE_TILE(BN_, N0, G1_, G2_, TG_, WAIT_STMT)                                              \
  {                                                                                                      \
    EpiRows rows;                                                                                        \
    epi_prepare_rows<EC>(p, rows, m_base, (G1_), lane);                                                  \
    const bool tile_fast = EC == EC_QKV || (p.fast_ok && !__any_sync(0xffffffffu, rows.multi));          \
    WAIT_STMT;                                                                                           \
    tc_fence_after();                                                                                    \
    _Pragma("unroll 1") for (int c = c_first; c < (BN_); c += kEpiColStep) {                             \
      const int n_base = (N0) + c;                                                                       \
      if (n_base >= p.N) break;                                                                          \
      uint32_t raw[32];                                                                                  \
      tmem_ld32(tmem_d + c, raw);                                                                        \
      if (tile_fast && (EC == EC_QKV || n_base + 32 <= p.N)) {                                           \
        EpiOperands op;                                                                                  \
        epi_load_operands<EC>(p, rows, op, n_base, (G1_), (G2_), (TG_), lane);                           \
        tmem_ld_wait();                                                                                  \
        if (EC == EC_QKV) {                                                                              \
          epilogue_qkv(p, raw, stg, rows, op, m_base, n_base, lane);                                     \
        } else {                                                                                         \
          float v[32];                                                                                   \
          epi_transpose(raw, stg, lane, v);                                                              \
          epilogue_generic<EC>(p, v, rows, op, n_base, (G1_), (G2_), (TG_), lane);                       \
        }                                                                                                \
      } else {                                                                                           \
        tmem_ld_wait();                                                                                  \
        Raw32 rr;                                                                                        \
        _Pragma("unroll") for (int j = 0; j < 32; ++j) rr.r[j] = raw[j];                                 \
        epilogue_rowwise_chunk(p, rr, m_base + lane, n_base, (G1_), (G2_), (TG_));                       \
      }                                                                                                  \
    }                                                                                                    \
  }

// ------------------------------------------------------------------ kernel
// Persistent: one CTA per SM walks the tile list (m fastest). The smem ring is as deep as
// shared memory allows (6-10 stages) so a latency-bound small GEMM has all of its k-blocks in
// flight, and two TMEM accumulators let the epilogue of tile i overlap the mainloop of tile i+1.
template <int BN>
struct GemmCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStages = kPipeBytes / (kABytes + kBBytes) > 12 ? 12 : kPipeBytes / (kABytes + kBBytes);
  static constexpr int kAccStride = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kSmemBytes = kStages * (kABytes + kBBytes) + 1024 /*align*/ + 1024 /*barriers*/ + kEpiStageBytes;
  static_assert(BN % 32 == 0 && BN <= 256, "N tile");
};

struct TileCoord {
  int m_tile, n_tile, g1, g2, tg;
  int m0;  // first A / output row of the tile: m_tile * 128, or tile_row0[m_tile] (sparse grouped layouts)
  bool skip;
};

// `m_used` = number of M tiles actually present (device value in grouped mode, else p.m_tiles): the flat tile
// space is COMPACT over the used tiles, so a sparse tile table (most of p.m_tiles unused) still spreads evenly
// over the persistent CTAs.
__device__ __forceinline__ TileCoord decode_tile(const GemmKParams& p, int t, int m_used) {
  TileCoord c;
  const int per_g = m_used * p.n_tiles;
  const int g = t / per_g;
  const int r = t - g * per_g;
  c.n_tile = r / m_used;
  c.m_tile = r - c.n_tile * m_used;
  c.g1 = g % p.G1;
  c.g2 = g / p.G1;
  c.skip = false;
  c.tg = p.tile_group ? p.tile_group[c.m_tile] : 0;
  c.m0 = p.tile_row0 ? p.tile_row0[c.m_tile] : c.m_tile * BM;
  return c;
}

template <int BN, bool TF32, int EC>
__global__ void __launch_bounds__(kThreads, 1)
vb_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmKParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kStages * Cfg::kBBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + kStages * (Cfg::kABytes + Cfg::kBBytes) + 1024);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    VB_TRACE(0);
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], BN == 32 ? kEpiWarps * 16 : kEpiWarps * 32);
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // prologue above overlaps the previous kernel; global memory is touched only below
  if (threadIdx.x == 0) VB_TRACE(1);

  const int num_kb = p.num_kb;
  const int m_used = p.m_tiles_used ? min(*p.m_tiles_used, p.m_tiles) : p.m_tiles;
  const int num_tiles = m_used * p.n_tiles * (p.num_tiles / (p.m_tiles * p.n_tiles));

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t, m_used);
        if (tc.skip) continue;
        const int m0 = tc.m0, n0 = tc.n_tile * BN;
        const int a_c2 = tc.g1 * p.a_c2_g1 + tc.g2 * p.a_c2_g2;
        const int a_c3 = tc.g1 * p.a_c3_g1 + tc.g2 * p.a_c3_g2;
        const int b_c2 = tc.g1 * p.b_c2_g1 + tc.g2 * p.b_c2_g2 + tc.tg * p.b_c2_tile;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kABytes + Cfg::kBBytes);
          const int tap = kb / p.kpt;
          constexpr int kBkElems = TF32 ? 32 : 64;
          const int c0 = (kb - tap * p.kpt) * kBkElems;
          const int c1 = m0 + tap * p.dil - p.pad - (p.dbg_shift & 0xffff);
          tma_load_4d(smem_a + stage * Cfg::kABytes, &tmA, &full_bar[stage], c0, c1, a_c2, a_c3);
          tma_load_3d(smem_b + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * kBkElems, n0, b_c2);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = TF32 ? make_idesc_tf32(BM, BN) : make_idesc_fp16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t, m_used);
        if (tc.skip) continue;
        const int acc = local & 1;
        mbar_wait(&tmem_empty_bar[acc], ((local >> 1) & 1) ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * Cfg::kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kb == 0 && local == 0) VB_TRACE(2);
          const int dsh = p.dbg_shift & 0xffff;
          const uint64_t da = make_kmajor_sw128_desc(smem_u32(smem_a + stage * Cfg::kABytes) + dsh * 128) |
                              ((p.dbg_shift >> 16) ? 0ull : (static_cast<uint64_t>(dsh & 7) << 49));  // base_offset
          const uint64_t db = make_kmajor_sw128_desc(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // advance 32 bytes (16 fp16 or 8 tf32 = one MMA K-step) inside the 128-byte swizzle
            // row: +2 in 16-byte units
            if (TF32) umma_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else umma_fp16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete
        if (local == 0) VB_TRACE(3);
        ++local;
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int q = warp_idx & 3;  // TMEM lane quadrant this warp may access
    // BN >= 64: the two warps of a quadrant split a tile's chunks even/odd. BN == 32 (one chunk):
    // they alternate TILES instead — warp set s owns TMEM accumulator s.
    const int wset = (warp_idx - 2) >> 2;
    const int c_first = BN == 32 ? 0 : wset * 32;
    const uint32_t stg = smem_u32(epi_stage) + (warp_idx - 2) * 4096;
    int local = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t, m_used);
      if (tc.skip) continue;
      const int acc = local & 1;
      if (BN == 32 && acc != wset) { ++local; continue; }
      const int m_base = tc.m0 + q * 32, n0 = tc.n_tile * BN;
      const uint32_t tmem_d = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(q * 32) << 16);
      if (EC == EC_ROUTE) {
        mbar_wait(&tmem_full_bar[acc], (local >> 1) & 1);
        tc_fence_after();
        if (wset == 0) epilogue_route_tile(p, tmem_d, m_base + lane, tc.n_tile, tc.g1);
      } else
      VB_EPILOGUE_TILE(BN, n0, tc.g1, tc.g2, tc.tg,
                       mbar_wait(&tmem_full_bar[acc], (local >> 1) & 1); if (local == 0 && threadIdx.x == 64) VB_TRACE(4))
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);  // all epilogue threads release the accumulator to the MMA warp
      if (local == 0 && threadIdx.x == 64) VB_TRACE(5);
      ++local;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) VB_TRACE(6);
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------ 2-CTA kernel
// CTA pairs (cluster of 2 along M) run tcgen05.mma.cta_group::2: UMMA M = 256 (128 rows per CTA),
// N = BN. Each CTA loads its own A tile and only HALF of the B tile; the tensor core reads both
// halves across the pair. Per output element the operand bytes an SM has to pull from L2 drop
// from (1/128 + 1/BN) to (1/128 + 1/(2 BN)) ... with BN = 256: half of the 128x128 kernel.
// Protocol (leader = cluster rank 0): both CTAs' TMA credit the LEADER's full barrier, the leader
// issues the MMAs and its commits arrive (multicast) on both CTAs' empty / accumulator-full
// barriers, both epilogues release the accumulator on the leader's tmem_empty barrier.
template <int BN>
struct Gemm2Cfg {
  static constexpr int kABytes = BM * 128;
  static constexpr int kBBytes = (BN / 2) * 128;
  static constexpr int kStages = kPipeBytes / (kABytes + kBBytes) > 12 ? 12 : kPipeBytes / (kABytes + kBBytes);
  static constexpr int kAccStride = BN <= 128 ? 128 : 256;
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kSmemBytes = kStages * (kABytes + kBBytes) + 1024 + 1024 + kEpiStageBytes;
};

template <int BN, bool TF32, int EC>
__global__ void __launch_bounds__(kThreads, 1)
vb_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ GemmKParams p) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kBkElems = TF32 ? 32 : 64;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kStages * Cfg::kBBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + kStages * (Cfg::kABytes + Cfg::kBBytes) + 1024);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 2);   // leader's arrive.expect_tx + the peer producer's arrive
      mbar_init(&empty_bar[s], 1);  // one multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 2 * kEpiWarps * 32);  // the epilogue threads of both CTAs
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc2<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();

  const int num_kb = p.num_kb;
  const int m_pairs = (p.m_tiles + 1) >> 1;
  const int per_g = m_pairs * p.n_tiles;
  const int num_pairs = per_g * (p.num_tiles / (p.m_tiles * p.n_tiles));  // x groups
  const int n_clusters = gridDim.x >> 1;
  const int first = blockIdx.x >> 1;

  if (warp_idx == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = first; t < num_pairs; t += n_clusters) {
        const int g = t / per_g, r = t - g * per_g;
        const int n_tile = r / m_pairs, m_tile = 2 * (r - n_tile * m_pairs) + (int)rank;
        const int g1 = g % p.G1, g2 = g / p.G1;
        const int m0 = m_tile * BM, n0 = n_tile * BN + (int)rank * (BN / 2);
        const int a_c2 = g1 * p.a_c2_g1 + g2 * p.a_c2_g2;
        const int a_c3 = g1 * p.a_c3_g1 + g2 * p.a_c3_g2;
        const int b_c2 = g1 * p.b_c2_g1 + g2 * p.b_c2_g2;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t lead_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          const int tap = kb / p.kpt;
          const int c0 = (kb - tap * p.kpt) * kBkElems;
          const int c1 = m0 + tap * p.dil - p.pad;
          tma_load_4d_2sm(smem_a + stage * Cfg::kABytes, &tmA, lead_full, c0, c1, a_c2, a_c3);
          tma_load_3d_2sm(smem_b + stage * Cfg::kBBytes, &tmB, lead_full, kb * kBkElems, n0, b_c2);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * (Cfg::kABytes + Cfg::kBBytes));
          else mbar_arrive_cluster(lead_full);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (leader CTA, single thread) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = TF32 ? make_idesc_tf32(2 * BM, BN) : make_idesc_fp16(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int t = first; t < num_pairs; t += n_clusters) {
        const int acc = local & 1;
        mbar_wait(&tmem_empty_bar[acc], ((local >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * Cfg::kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_kmajor_sw128_desc(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t db = make_kmajor_sw128_desc(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (TF32) umma2_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else umma2_fp16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit2_mc(&empty_bar[stage], 3);  // frees the slot in both CTAs
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit2_mc(&tmem_full_bar[acc], 3);
        ++local;
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs) =====================
    const int q = warp_idx & 3;
    const int c_first = ((warp_idx - 2) >> 2) * 32;
    const uint32_t stg = smem_u32(epi_stage) + (warp_idx - 2) * 4096;
    int local = 0;
    for (int t = first; t < num_pairs; t += n_clusters) {
      const int g = t / per_g, rr = t - g * per_g;
      const int n_tile = rr / m_pairs, m_tile = 2 * (rr - n_tile * m_pairs) + (int)rank;
      const int g1 = g % p.G1, g2 = g / p.G1;
      const int acc = local & 1;
      const int m_base = m_tile * BM + q * 32, n0 = n_tile * BN;
      const uint32_t tmem_d = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(q * 32) << 16);
      VB_EPILOGUE_TILE(BN, n0, g1, g2, 0, mbar_wait(&tmem_full_bar[acc], (local >> 1) & 1))
      tc_fence_before();
      if (leader) mbar_arrive(&tmem_empty_bar[acc]);
      else mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty_bar[acc]), 0));
      ++local;
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still signal my barriers / read my smem until here
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc2<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------ haloed small-C conv kernel
// Conv1d with C_in <= 64 (one 128-byte k-block per tap) and N <= 64 output channels — the two last
// HiFi-GAN stages. The per-tap implicit-GEMM kernel above re-fetches the same activation rows once
// per tap (k = 3..11); here each tile's rows are fetched ONCE as a tall tile of 128 + (taps-1)*dil
// rows, tap j is the same shared-memory tile seen through a UMMA descriptor shifted by j*dil rows
// (start address + rows*128 B; base_offset stays 0 because the 128B swizzle is a pure function of the
// shared-memory address; tests/test_conv_halo_gpu.py), and all tap weights stay resident in shared
// memory for the lifetime of the persistent CTA. Operand traffic per output row drops ~taps-fold.

template <int BN, int KSTEPS, int EC>
__global__ void __launch_bounds__(kThreads, 1)
vb_conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ GemmKParams p, const HaloParams hp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  constexpr int kWTap = BN * 128;             // bytes of one tap's [BN x 64] weight block
  const int a_slot = hp.rows8 * 128;          // bytes of one haloed activation tile
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + ((hp.taps * kWTap + 1023) & ~1023);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_a + hp.stages * a_slot);
  uint64_t* a_empty = a_full + 8;
  uint64_t* w_full = a_empty + 8;
  uint64_t* tmem_full_bar = w_full + 1;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* epi_stage = reinterpret_cast<float*>(smem_a + hp.stages * a_slot + 1024);
  constexpr int kAcc = BN <= 32 ? 32 : 64;

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < hp.stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], BN == 32 ? kEpiWarps * 16 : kEpiWarps * 32); }
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc<2 * kAcc>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();

  if (warp_idx == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_full, hp.taps * kWTap);
      for (int j = 0; j < hp.taps; ++j) tma_load_3d(smem_w + j * kWTap, &tmB, w_full, j * 64, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int g = t / p.m_tiles, m_tile = t - g * p.m_tiles;
        mbar_wait(&a_empty[stage], phase ^ 1);
        mbar_expect_tx(&a_full[stage], hp.rows * 128);
        tma_load_4d(smem_a + stage * a_slot, &tmA, &a_full[stage], 0, m_tile * BM - p.pad, g * p.a_c2_g1, 0);
        if (++stage == hp.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_fp16(BM, BN);
      mbar_wait(w_full, 0);
      const uint64_t w_desc0 = make_kmajor_sw128_desc(smem_u32(smem_w));
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int acc = local & 1;
        mbar_wait(&tmem_empty_bar[acc], ((local >> 1) & 1) ^ 1);
        mbar_wait(&a_full[stage], phase);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kAcc;
        // The single issuing thread is the bottleneck for these skinny (128 x BN x 16) MMAs: keep the
        // per-MMA instruction count minimal — descriptors advance by plain adds (16-byte units).
        uint64_t da = make_kmajor_sw128_desc(smem_u32(smem_a + stage * a_slot));
        uint64_t db = w_desc0;
        const uint64_t da_step = static_cast<uint64_t>(hp.row_shift * 8);  // row_shift rows * 128 B / 16
        constexpr uint64_t db_step = kWTap / 16;
        umma_fp16(tmem_d, da, db, idesc, 0u);
#pragma unroll
        for (int k = 1; k < KSTEPS; ++k) umma_fp16(tmem_d, da + 2 * k, db + 2 * k, idesc, 1u);
#pragma unroll 1
        for (int j = 1; j < hp.taps; ++j) {
          da += da_step;
          db += db_step;
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k) umma_fp16(tmem_d, da + 2 * k, db + 2 * k, idesc, 1u);
        }
        umma_commit(&a_empty[stage]);
        umma_commit(&tmem_full_bar[acc]);
        if (++stage == hp.stages) { stage = 0; phase ^= 1; }
        ++local;
      }
    }
  } else {
    const int q = warp_idx & 3;
    const int wset = (warp_idx - 2) >> 2;  // BN == 32: warp set s owns accumulator s (see vb_gemm_kernel)
    const int c_first = BN == 32 ? 0 : wset * 32;
    const uint32_t stg = smem_u32(epi_stage) + (warp_idx - 2) * 4096;
    int local = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      const int g = t / p.m_tiles, m_tile = t - g * p.m_tiles;
      const int g1 = g % p.G1, g2 = g / p.G1;
      const int acc = local & 1;
      if (BN == 32 && acc != wset) { ++local; continue; }
      const int m_base = m_tile * BM + q * 32;
      const uint32_t tmem_d = tmem_base + acc * kAcc + (static_cast<uint32_t>(q * 32) << 16);
      VB_EPILOGUE_TILE(BN, 0, g1, g2, 0, mbar_wait(&tmem_full_bar[acc], (local >> 1) & 1))
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      ++local;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<2 * kAcc>(tmem_base);
  }
}


// ------------------------------------------------------------------ chained problems in one launch
// vb_chain_kernel: the persistent 1-CTA GEMM above walking a flat tile list that spans up to kChainMaxOps
// dependent problems (vb_gemm_params.h: ChainOp). One launch, one prologue, one TMEM allocation; the smem
// ring and the two TMEM accumulators run straight through problem boundaries, so the epilogue tail of
// problem o overlaps the mainloop of problem o+1 and no SM idles between them except where the data
// dependency itself forces it. All CTAs are co-resident (grid <= SM count, 1 CTA/SM) and every CTA walks its
// tiles in increasing flat order while dependencies only point to smaller indices, so waiting cannot deadlock.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct ChainTile {
  int o, m_tile, n_tile, g1, g2, g, tg, m0;
  bool skip;
};
__device__ __forceinline__ ChainTile chain_decode(const ChainParams& cp, int t, const int* m_used) {
  ChainTile c;
  int o = 0;
#pragma unroll
  for (int i = 1; i < kChainMaxOps; ++i)
    if (i < cp.n_ops && t >= cp.op[i].tile_start) o = i;
  const GemmKParams& p = cp.op[o].p;
  const int local = t - cp.op[o].tile_start;
  const int per_g = p.m_tiles * p.n_tiles;
  c.o = o;
  c.g = local / per_g;
  const int r = local - c.g * per_g;
  if (cp.dbg & 4) {
    c.n_tile = r / p.m_tiles;
    c.m_tile = r - c.n_tile * p.m_tiles;
  } else {
    c.m_tile = r / p.n_tiles;            // n fastest: a row block's N tiles are adjacent in the list
    c.n_tile = r - c.m_tile * p.n_tiles;
  }
  c.g1 = c.g % p.G1;
  c.g2 = c.g / p.G1;
  c.skip = c.m_tile >= m_used[o];
  c.tg = (p.tile_group && !c.skip) ? p.tile_group[c.m_tile] : 0;
  c.m0 = (p.tile_row0 && !c.skip) ? p.tile_row0[c.m_tile] : c.m_tile * BM;
  return c;
}

template <int BN, bool TF32>
__global__ void __launch_bounds__(kThreads, 1)
vb_chain_kernel(const __grid_constant__ ChainParams cp) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  constexpr int EC = EC_CHAIN;
  static_assert(BN >= 64, "chain tiles use all eight epilogue warps");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kStages * Cfg::kBBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + kStages * (Cfg::kABytes + Cfg::kBBytes) + 1024);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    for (int o = 0; o < cp.n_ops; ++o) {
      tma_prefetch_desc(&cp.op[o].tmA);
      tma_prefetch_desc(&cp.op[o].tmB);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], kEpiWarps * 32);
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();

  int m_used[kChainMaxOps];
#pragma unroll
  for (int o = 0; o < kChainMaxOps; ++o)
    m_used[o] = o < cp.n_ops ? (cp.op[o].p.m_tiles_used ? *cp.op[o].p.m_tiles_used : cp.op[o].p.m_tiles) : 0;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < cp.total_tiles; t += gridDim.x) {
        const ChainTile tc = chain_decode(cp, t, m_used);
        if (tc.skip) continue;
        const ChainOp& op = cp.op[tc.o];
        const GemmKParams& p = op.p;
        if (op.dep_mode != 0) {
          // wait for the producer problem's rows, then order the generic-proxy stores we just observed
          // before this thread's async-proxy (TMA) reads
          const ChainOp& prev = cp.op[tc.o - 1];
          const int* cnt = op.dep_mode == 1 ? prev.cnt_rows + (tc.g * p.m_tiles + tc.m_tile) : prev.cnt_all;
          const int expect = op.dep_mode == 1 ? prev.p.n_tiles : m_used[tc.o - 1] * prev.p.n_tiles * (prev.p.num_tiles / (prev.p.m_tiles * prev.p.n_tiles));
          const long long t0 = clock64();
          while (ld_acquire_gpu(cnt) < expect) {
            if (clock64() - t0 > 4000000000ll) { *cp.err_flag = 1; break; }  // ~2 s: a bug, not a wait
          }
          fence_proxy_async_all();
        }
        const int m0 = tc.m0, n0 = tc.n_tile * BN;
        const int a_c2 = tc.g1 * p.a_c2_g1 + tc.g2 * p.a_c2_g2;
        const int a_c3 = tc.g1 * p.a_c3_g1 + tc.g2 * p.a_c3_g2;
        const int b_c2 = tc.g1 * p.b_c2_g1 + tc.g2 * p.b_c2_g2 + tc.tg * p.b_c2_tile;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kABytes + Cfg::kBBytes);
          const int tap = kb / p.kpt;
          constexpr int kBkElems = TF32 ? 32 : 64;
          const int c0 = (kb - tap * p.kpt) * kBkElems;
          const int c1 = m0 + tap * p.dil - p.pad;
          tma_load_4d(smem_a + stage * Cfg::kABytes, &op.tmA, &full_bar[stage], c0, c1, a_c2, a_c3);
          tma_load_3d(smem_b + stage * Cfg::kBBytes, &op.tmB, &full_bar[stage], kb * kBkElems, n0, b_c2);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = TF32 ? make_idesc_tf32(BM, BN) : make_idesc_fp16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int t = blockIdx.x; t < cp.total_tiles; t += gridDim.x) {
        const ChainTile tc = chain_decode(cp, t, m_used);
        if (tc.skip) continue;
        const int num_kb = cp.op[tc.o].p.num_kb;
        const int acc = local & 1;
        mbar_wait(&tmem_empty_bar[acc], ((local >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * Cfg::kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_kmajor_sw128_desc(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t db = make_kmajor_sw128_desc(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (TF32) umma_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else umma_fp16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);
        ++local;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp_idx & 3;
    const int wset = (warp_idx - 2) >> 2;
    const int c_first = wset * 32;
    const uint32_t stg = smem_u32(epi_stage) + (warp_idx - 2) * 4096;
    int local = 0;
    for (int t = blockIdx.x; t < cp.total_tiles; t += gridDim.x) {
      const ChainTile tc = chain_decode(cp, t, m_used);
      if (tc.skip) continue;
      const ChainOp& op = cp.op[tc.o];
      const GemmKParams& p = op.p;
      const int acc = local & 1;
      const int m_base = tc.m0 + q * 32, n0 = tc.n_tile * BN;
      const uint32_t tmem_d = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(q * 32) << 16);
      VB_EPILOGUE_TILE(BN, n0, tc.g1, tc.g2, tc.tg, mbar_wait(&tmem_full_bar[acc], (local >> 1) & 1))
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      if (op.sig_mode != 0) {
        // publish: every epilogue thread's stores of this tile are fenced, then ONE arrival per tile
        if (!(cp.dbg & 2)) {
          fence_proxy_async_all();
          __threadfence();
        }
        named_bar_sync(1, kEpiWarps * 32);
        if (threadIdx.x == 64) {
          if (op.sig_mode == 1) atomicAdd(op.cnt_rows + (tc.g * p.m_tiles + tc.m_tile), 1);
          else atomicAdd(op.cnt_all, 1);
        }
      }
      ++local;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN, bool TF32>
static int launch_chain_t(const ChainParams& cp, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static PerDevice<bool> attr_flag;
  bool& attr_set = attr_flag.get();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(vb_chain_kernel<BN, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "cudaFuncSetAttribute(chain smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
    attr_set = true;
  }
  cudaError_t e = launch_k(vb_chain_kernel<BN, TF32>, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, stream, cp);
  if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "vb_gemm_chain launch failed: %s", cudaGetErrorString(e));
  return check_launch("vb_gemm_chain");
}

template <int BN, bool TF32, int EC>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p,
                  dim3 grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static PerDevice<bool> attr_flag;
  bool& attr_set = attr_flag.get();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(vb_gemm_kernel<BN, TF32, EC>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess)
      return set_error(VB_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes,
                       cudaGetErrorString(e));
    attr_set = true;
  }
  launch_k(vb_gemm_kernel<BN, TF32, EC>, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, stream, tmA, tmB, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return set_error(VB_ERR_CUDA, "vb_gemm launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return VB_OK;
}

template <int BN, bool TF32, int EC>
static int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p, int grid,
                   cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN>;
  static PerDevice<bool> attr_flag;
  bool& attr_set = attr_flag.get();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(vb_gemm2_kernel<BN, TF32, EC>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess)
      return set_error(VB_ERR_CUDA, "cudaFuncSetAttribute(gemm2 smem=%d): %s", Cfg::kSmemBytes,
                       cudaGetErrorString(e));
    attr_set = true;
  }
  cudaError_t e = launch_k_cluster(vb_gemm2_kernel<BN, TF32, EC>, 2, dim3(grid), dim3(kThreads), Cfg::kSmemBytes,
                                   stream, tmA, tmB, p);
  if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "vb_gemm (2-CTA) launch failed: %s", cudaGetErrorString(e));
  return check_launch("vb_gemm2");
}


template <int EC>
int launch_gemm_class(int bn, bool tf32, bool use2, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p,
                      int grid_, cudaStream_t stream) {
  if constexpr (EC == EC_ROUTE) {
    if (use2 || bn != 96) return set_error(VB_ERR_INVALID, "vb_gemm: the ROUTE epilogue runs on 128x96 tiles only");
    return tf32 ? launch<96, true, EC>(tmA, tmB, p, dim3(grid_, 1, 1), stream)
                : launch<96, false, EC>(tmA, tmB, p, dim3(grid_, 1, 1), stream);
  } else {
  if (use2) {
    if (bn == 256) return tf32 ? launch2<256, true, EC>(tmA, tmB, p, grid_, stream) : launch2<256, false, EC>(tmA, tmB, p, grid_, stream);
    if (bn == 192) return tf32 ? launch2<192, true, EC>(tmA, tmB, p, grid_, stream) : launch2<192, false, EC>(tmA, tmB, p, grid_, stream);
    return tf32 ? launch2<128, true, EC>(tmA, tmB, p, grid_, stream) : launch2<128, false, EC>(tmA, tmB, p, grid_, stream);
  }
  dim3 grid(grid_, 1, 1);
  switch (bn) {
    case 32: return tf32 ? launch<32, true, EC>(tmA, tmB, p, grid, stream) : launch<32, false, EC>(tmA, tmB, p, grid, stream);
    case 64: return tf32 ? launch<64, true, EC>(tmA, tmB, p, grid, stream) : launch<64, false, EC>(tmA, tmB, p, grid, stream);
    case 96: return tf32 ? launch<96, true, EC>(tmA, tmB, p, grid, stream) : launch<96, false, EC>(tmA, tmB, p, grid, stream);
    case 128: return tf32 ? launch<128, true, EC>(tmA, tmB, p, grid, stream) : launch<128, false, EC>(tmA, tmB, p, grid, stream);
    case 192: return tf32 ? launch<192, true, EC>(tmA, tmB, p, grid, stream) : launch<192, false, EC>(tmA, tmB, p, grid, stream);
    case 256: return tf32 ? launch<256, true, EC>(tmA, tmB, p, grid, stream) : launch<256, false, EC>(tmA, tmB, p, grid, stream);
    default: return set_error(VB_ERR_INVALID, "vb_gemm: unsupported N tile %d", bn);
  }
  }
}

template <int EC>
int launch_halo_class(int hbn, bool k2, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p,
                      const HaloParams& hp, int grid, int smem_bytes, cudaStream_t stream) {
  cudaError_t e;
  auto go = [&](auto kern) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (err == cudaSuccess) err = launch_k(kern, dim3(grid), dim3(kThreads), smem_bytes, stream, tmA, tmB, p, hp);
    return err;
  };
  if (hbn == 32) e = k2 ? go(vb_conv_halo_kernel<32, 2, EC>) : go(vb_conv_halo_kernel<32, 4, EC>);
  else e = k2 ? go(vb_conv_halo_kernel<64, 2, EC>) : go(vb_conv_halo_kernel<64, 4, EC>);
  if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "vb_gemm (conv halo) launch failed: %s", cudaGetErrorString(e));
  return check_launch("vb_conv_halo");
}

}  // namespace vb
