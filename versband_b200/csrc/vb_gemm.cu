// vb_gemm — tcgen05 / TMEM / TMA GEMM for sm_100a with implicit-conv A addressing and
// fused epilogues. See include/versband.h for the contract.
//
// CTA = 192 threads: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4). One 128 x BN output tile per
// CTA; several CTAs are co-resident per SM (smem permitting) so one CTA's epilogue overlaps
// another's mainloop.
#include <math.h>
#include "vb_internal.h"
#include "vb_ptx.cuh"

namespace vb {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 elements per k-block = one 128-byte swizzle row (32 for fp32/tf32 operands)
constexpr int kThreads = 192;

struct GemmKParams {
  int M, N, K;
  int G1;
  int num_kb, kpt, dil, pad;
  int m_tiles, n_tiles, num_tiles;
  int a_c2_g1, a_c2_g2, a_c3_g1, a_c3_g2;
  int b_c2_g1, b_c2_g2, b_c2_tile;
  const int* tile_group;
  const int* m_tiles_used;
  // epilogue
  int mode;
  float alpha;
  int swiglu;
  const float* bias;
  long long bias_g1, bias_tile;
  int act;
  float act_slope;
  const float* row_scale;
  const float* gate;
  int gate_rows;
  long long gate_ld, gate_g1;
  const int* row_map;
  int row_mul, row_off, row_off_g1, rows_out;
  const float* residual;
  int accumulate;
  float* out_f32;
  long long out_ld, out_g1, out_g2;
  __nv_bfloat16* out_bf16;
  int outb_f32;  // the operand-output slot stores fp32 (tf32 operand mode) instead of bf16
  long long outb_ld, outb_g1, outb_g2;
  int act_out;
  float act_out_slope, out_scale;
  int dbg_shift;  // experiment: A tile loaded `dbg_shift` rows early, UMMA descriptor shifted back
  int vec_f32, vec_bf16, vec_aux;  // alignment allows 16-byte accesses (outputs / bias+gate)
  // qkv-rope
  const float* rope_cs;
  int T, heads, head_dim, Tp, qkv_sec0;
  __nv_bfloat16* q_out;
  __nv_bfloat16* k_out;
  __nv_bfloat16* vt_out;
};

// Activations are applied by whole-chunk loops selected ONCE per chunk (one uniform branch, then a
// straight-line 32-wide loop): keeps the executed instruction stream short — the epilogue runs per
// tile on four warps and was instruction-fetch bound when the switch sat inside the unrolled loop.
template <int ACT>
__device__ __forceinline__ float act_one(float v, float slope) {
  if (ACT == VB_ACT_SILU) return v / (1.f + __expf(-v));
  if (ACT == VB_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (ACT == VB_ACT_GELU_ERF) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  if (ACT == VB_ACT_TANH) return tanhf(v);
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
    default:
      if (pre != 1.f) act_loop<VB_ACT_NONE>(v, slope, pre);
      break;
  }
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
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

// ------------------------------------------------------------------ generic epilogue
// One thread owns one output row (TMEM lane) and 32 consecutive columns of it.
__device__ __forceinline__ void epilogue_generic(const GemmKParams& p, float (&v)[32], int m,
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
      const float a = v[2 * j], b = v[2 * j + 1];
      o[j] = (a / (1.f + __expf(-a))) * b * rs;
    }
    if (p.out_bf16 && p.outb_f32) {
      float* dstf = reinterpret_cast<float*>(p.out_bf16) + g1 * p.outb_g1 + g2 * p.outb_g2 + orow * p.outb_ld + n_out0;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n_out0 + j < n_out_total) dstf[j] = o[j];
    } else if (p.out_bf16) {
      __nv_bfloat16* dst = p.out_bf16 + g1 * p.outb_g1 + g2 * p.outb_g2 + orow * p.outb_ld + n_out0;
      if (p.vec_bf16 && n_out0 + 16 <= n_out_total) {
        uint4 w0, w1;
        w0.x = pack_bf16(o[0], o[1]);   w0.y = pack_bf16(o[2], o[3]);
        w0.z = pack_bf16(o[4], o[5]);   w0.w = pack_bf16(o[6], o[7]);
        w1.x = pack_bf16(o[8], o[9]);   w1.y = pack_bf16(o[10], o[11]);
        w1.z = pack_bf16(o[12], o[13]); w1.w = pack_bf16(o[14], o[15]);
        reinterpret_cast<uint4*>(dst)[0] = w0;
        reinterpret_cast<uint4*>(dst)[1] = w1;
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n_out0 + j < n_out_total) dst[j] = __float2bfloat16(o[j]);
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
  if (p.out_bf16) {
    const long long offb = g1 * p.outb_g1 + g2 * p.outb_g2 + orow * p.outb_ld + n_base;
    apply_act_all(v, p.act_out, p.act_out_slope, p.out_scale);
    if (p.outb_f32) {
      float* dstf = reinterpret_cast<float*>(p.out_bf16) + offb;
      if (full && p.vec_bf16) {
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
    __nv_bfloat16* dst = p.out_bf16 + offb;
    if (full && p.vec_bf16) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 w;
        w.x = pack_bf16(v[8 * j], v[8 * j + 1]);
        w.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
        w.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
        w.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
        reinterpret_cast<uint4*>(dst)[j] = w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < n_left) dst[j] = __float2bfloat16(v[j]);
    }
  }
}

// ------------------------------------------------------------------ QKV + RoPE epilogue
// Row m = b*T + t, column n in [0, 3*H*hd): section = n / (H*hd) (q,k,v), head, dim.
// Rotary: pairs (2i, 2i+1) rotated by angle(t, i) — flag_large_dit_moe.py:263-269.
// A 32-column chunk never straddles a head or section because head_dim % 32 == 0.
__device__ __forceinline__ void epilogue_qkv(const GemmKParams& p, float (&v)[32], int m,
                                             int n_base) {
  const int hd = p.head_dim, H = p.heads, T = p.T;
  const int D = H * hd;
  const int sec0 = n_base / D;
  const int rem = n_base - sec0 * D;
  const int sec = sec0 + p.qkv_sec0;
  const int h = rem / hd;
  const int d0 = rem - h * hd;
  const int b = m / T;
  const int t = m - b * T;
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + n_base + j);
  }
  if (sec < 2) {
    float o[32];
    if (p.rope_cs) {
      const float* cs = p.rope_cs + ((long long)t * (hd >> 1) + (d0 >> 1)) * 2;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float c = __ldg(cs + 2 * j), s = __ldg(cs + 2 * j + 1);
        float x0 = v[2 * j], x1 = v[2 * j + 1];
        o[2 * j] = x0 * c - x1 * s;
        o[2 * j + 1] = x0 * s + x1 * c;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = v[j];
    }
    __nv_bfloat16* dst = (sec == 0 ? p.q_out : p.k_out) +
                         (((long long)b * H + h) * T + t) * hd + d0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 w;
      w.x = pack_bf16(o[8 * j], o[8 * j + 1]);
      w.y = pack_bf16(o[8 * j + 2], o[8 * j + 3]);
      w.z = pack_bf16(o[8 * j + 4], o[8 * j + 5]);
      w.w = pack_bf16(o[8 * j + 6], o[8 * j + 7]);
      reinterpret_cast<uint4*>(dst)[j] = w;
    }
  } else {
    // V^T[b][h][d][t]: lanes of a warp hold consecutive t -> coalesced 2-byte stores
    __nv_bfloat16* dst = p.vt_out + (((long long)b * H + h) * hd + d0) * p.Tp + t;
#pragma unroll
    for (int j = 0; j < 32; ++j) dst[(long long)j * p.Tp] = __float2bfloat16(v[j]);
  }
}

// ------------------------------------------------------------------ kernel
// Persistent: one CTA per SM walks the tile list (m fastest). The smem ring is as deep as
// shared memory allows (6-10 stages) so a latency-bound small GEMM has all of its k-blocks in
// flight, and two TMEM accumulators let the epilogue of tile i overlap the mainloop of tile i+1.
template <int BN>
struct GemmCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStages = (200 * 1024) / (kABytes + kBBytes) > 12 ? 12 : (200 * 1024) / (kABytes + kBBytes);
  static constexpr int kAccStride = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kSmemBytes = kStages * (kABytes + kBBytes) + 1024 /*align*/ + 512;
  static_assert(BN % 32 == 0 && BN <= 256, "N tile");
};

struct TileCoord {
  int m_tile, n_tile, g1, g2, tg;
  bool skip;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmKParams& p, int t, int m_used) {
  TileCoord c;
  const int per_g = p.m_tiles * p.n_tiles;
  const int g = t / per_g;
  const int r = t - g * per_g;
  c.n_tile = r / p.m_tiles;
  c.m_tile = r - c.n_tile * p.m_tiles;
  c.g1 = g % p.G1;
  c.g2 = g / p.G1;
  c.skip = c.m_tile >= m_used;
  c.tg = (p.tile_group && !c.skip) ? p.tile_group[c.m_tile] : 0;
  return c;
}

template <int BN, bool TF32>
__global__ void __launch_bounds__(kThreads, 1)
vb_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmKParams p) {
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

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 128);
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // prologue above overlaps the previous kernel; global memory is touched only below

  const int num_kb = p.num_kb;
  const int m_used = p.m_tiles_used ? *p.m_tiles_used : p.m_tiles;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t, m_used);
        if (tc.skip) continue;
        const int m0 = tc.m_tile * BM, n0 = tc.n_tile * BN;
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
      constexpr uint32_t idesc = TF32 ? make_idesc_tf32(BM, BN) : make_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t, m_used);
        if (tc.skip) continue;
        const int acc = local & 1;
        mbar_wait(&tmem_empty_bar[acc], ((local >> 1) & 1) ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * Cfg::kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const int dsh = p.dbg_shift & 0xffff;
          const uint64_t da = make_kmajor_sw128_desc(smem_u32(smem_a + stage * Cfg::kABytes) + dsh * 128) |
                              ((p.dbg_shift >> 16) ? 0ull : (static_cast<uint64_t>(dsh & 7) << 49));  // base_offset
          const uint64_t db = make_kmajor_sw128_desc(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // advance 32 bytes (16 bf16 or 8 tf32 = one MMA K-step) inside the 128-byte swizzle
            // row: +2 in 16-byte units
            if (TF32) umma_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete
        ++local;
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int q = warp_idx & 3;  // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;
    int local = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t, m_used);
      if (tc.skip) continue;
      const int acc = local & 1;
      const int m0 = tc.m_tile * BM, n0 = tc.n_tile * BN;
      const int m = m0 + r;
      const bool row_ok = m < p.M;
      long long orow = 0;
      bool store_ok = row_ok;
      if (row_ok && p.mode == VB_EPI_GENERIC) {
        if (p.row_map) orow = p.row_map[m];
        else orow = (long long)m * p.row_mul + p.row_off + (long long)tc.g1 * p.row_off_g1;
        store_ok = (orow >= 0 && orow < p.rows_out);
      }
      mbar_wait(&tmem_full_bar[acc], (local >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t raw[32];
        tmem_ld32(tmem_d + c, raw);
        tmem_ld_wait();
        const int n_base = n0 + c;
        if (store_ok && n_base < p.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          if (p.mode == VB_EPI_QKV_ROPE) epilogue_qkv(p, v, m, n_base);
          else epilogue_generic(p, v, m, n_base, tc.g1, tc.g2, tc.tg, orow);
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);  // 128 arrivals release the accumulator to the MMA warp
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
  static constexpr int kStages = (200 * 1024) / (kABytes + kBBytes) > 12 ? 12 : (200 * 1024) / (kABytes + kBBytes);
  static constexpr int kAccStride = BN <= 128 ? 128 : 256;
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kSmemBytes = kStages * (kABytes + kBBytes) + 1024 + 512;
};

template <int BN, bool TF32>
__global__ void __launch_bounds__(kThreads, 1)
vb_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const GemmKParams p) {
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
      mbar_init(&tmem_empty_bar[i], 256);  // 128 epilogue threads of each CTA
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
      constexpr uint32_t idesc = TF32 ? make_idesc_tf32(2 * BM, BN) : make_idesc_bf16(2 * BM, BN);
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
            else umma2_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
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
    const int r_in = q * 32 + lane;
    int local = 0;
    for (int t = first; t < num_pairs; t += n_clusters) {
      const int g = t / per_g, rr = t - g * per_g;
      const int n_tile = rr / m_pairs, m_tile = 2 * (rr - n_tile * m_pairs) + (int)rank;
      const int g1 = g % p.G1, g2 = g / p.G1;
      const int acc = local & 1;
      const int m0 = m_tile * BM, n0 = n_tile * BN;
      const int m = m0 + r_in;
      const bool row_ok = m < p.M;
      long long orow = 0;
      bool store_ok = row_ok;
      if (row_ok && p.mode == VB_EPI_GENERIC) {
        if (p.row_map) orow = p.row_map[m];
        else orow = (long long)m * p.row_mul + p.row_off + (long long)g1 * p.row_off_g1;
        store_ok = (orow >= 0 && orow < p.rows_out);
      }
      mbar_wait(&tmem_full_bar[acc], (local >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t raw[32];
        tmem_ld32(tmem_d + c, raw);
        tmem_ld_wait();
        const int n_base = n0 + c;
        if (store_ok && n_base < p.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          if (p.mode == VB_EPI_QKV_ROPE) epilogue_qkv(p, v, m, n_base);
          else epilogue_generic(p, v, m, n_base, g1, g2, 0, orow);
        }
      }
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
// shared-memory address — verified by tools/exp_rowshift.py), and all tap weights stay resident in shared
// memory for the lifetime of the persistent CTA. Operand traffic per output row drops ~taps-fold.
struct HaloParams {
  int taps, row_shift, rows, rows8, n_ksteps, stages, base_off;
};

template <int BN, int KSTEPS>
__global__ void __launch_bounds__(kThreads, 1)
vb_conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmKParams p, const HaloParams hp) {
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
  constexpr int kAcc = BN <= 32 ? 32 : 64;

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < hp.stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 128); }
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
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
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
        umma_bf16(tmem_d, da, db, idesc, 0u);
#pragma unroll
        for (int k = 1; k < KSTEPS; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, 1u);
#pragma unroll 1
        for (int j = 1; j < hp.taps; ++j) {
          da += da_step;
          db += db_step;
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, 1u);
        }
        umma_commit(&a_empty[stage]);
        umma_commit(&tmem_full_bar[acc]);
        if (++stage == hp.stages) { stage = 0; phase ^= 1; }
        ++local;
      }
    }
  } else {
    const int q = warp_idx & 3;
    const int r = q * 32 + lane;
    int local = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      const int g = t / p.m_tiles, m_tile = t - g * p.m_tiles;
      const int g1 = g % p.G1, g2 = g / p.G1;
      const int acc = local & 1;
      const int m = m_tile * BM + r;
      const bool row_ok = m < p.M;
      long long orow = 0;
      bool store_ok = row_ok;
      if (row_ok) {
        if (p.row_map) orow = p.row_map[m];
        else orow = (long long)m * p.row_mul + p.row_off + (long long)g1 * p.row_off_g1;
        store_ok = (orow >= 0 && orow < p.rows_out);
      }
      mbar_wait(&tmem_full_bar[acc], (local >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * kAcc + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t raw[32];
        tmem_ld32(tmem_d + c, raw);
        tmem_ld_wait();
        if (store_ok && c < p.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          epilogue_generic(p, v, m, c, g1, g2, 0, orow);
        }
      }
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

// ------------------------------------------------------------------ host side
static int encode_map(CUtensorMap* map, const void* base, int rank, const int64_t* dims,
                      const int64_t* strides, const uint32_t* box, bool tf32) {
  const int esz = tf32 ? 4 : 2;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = static_cast<cuuint64_t>(dims[i]);
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = static_cast<cuuint64_t>(strides[i]) * esz;  // bytes
      if (gstr[i - 1] % 16 != 0 || gstr[i - 1] == 0)
        return set_error(VB_ERR_INVALID, "vb_gemm: tensor stride %lld of dim %d is not a positive multiple of 16 bytes",
                         (long long)strides[i], i);
    }
    if (dims[i] <= 0) return set_error(VB_ERR_INVALID, "vb_gemm: non-positive tensor dim");
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return set_error(VB_ERR_INVALID, "vb_gemm: operand base pointer not 16-byte aligned");
  auto fn = get_tensor_map_encoder();
  if (!fn) return set_error(VB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  CUresult r = fn(map, tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr,
                  bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(VB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return VB_OK;
}

template <int BN, bool TF32>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p,
                  dim3 grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(vb_gemm_kernel<BN, TF32>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess)
      return set_error(VB_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes,
                       cudaGetErrorString(e));
    attr_set = true;
  }
  launch_k(vb_gemm_kernel<BN, TF32>, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, stream, tmA, tmB, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return set_error(VB_ERR_CUDA, "vb_gemm launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return VB_OK;
}

template <int BN, bool TF32>
static int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p, int grid,
                   cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(vb_gemm2_kernel<BN, TF32>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess)
      return set_error(VB_ERR_CUDA, "cudaFuncSetAttribute(gemm2 smem=%d): %s", Cfg::kSmemBytes,
                       cudaGetErrorString(e));
    attr_set = true;
  }
  cudaError_t e = launch_k_cluster(vb_gemm2_kernel<BN, TF32>, 2, dim3(grid), dim3(kThreads), Cfg::kSmemBytes,
                                   stream, tmA, tmB, p);
  if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "vb_gemm (2-CTA) launch failed: %s", cudaGetErrorString(e));
  return check_launch("vb_gemm2");
}

static int g_dbg_shift = 0;
static int g_conv_halo = 1;  // 0 off, 1 on
static int g_gemm2_mode = 1;  // 0 = off, 1 = cost model, 2 = whenever legal

}  // namespace vb

extern "C" int vb_set_conv_halo(int mode) {
  vb::g_conv_halo = mode;
  return VB_OK;
}

extern "C" int vb_set_debug_shift(int rows) {
  vb::g_dbg_shift = rows;
  return VB_OK;
}

extern "C" int vb_set_gemm2(int mode) {
  vb::g_gemm2_mode = mode;
  return VB_OK;
}

extern "C" int vb_gemm(const vb_gemm_args* a, void* stream_) {
  vb::ProfScope prof_scope(vb::PK_GEMM, stream_);
  using namespace vb;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a || !a->a || !a->b) return set_error(VB_ERR_INVALID, "vb_gemm: null operand");
  const bool tf32 = a->operand_f32 != 0;
  const int bk = tf32 ? 32 : 64;  // elements per 128-byte k-block
  if (a->M <= 0 || a->N <= 0 || a->K <= 0 || a->K % bk != 0)
    return set_error(VB_ERR_INVALID, "vb_gemm: bad problem size M=%d N=%d K=%d (K must be a multiple of %d)",
                     a->M, a->N, a->K, bk);
  if (a->G1 <= 0 || a->G2 <= 0) return set_error(VB_ERR_INVALID, "vb_gemm: bad group grid");
  const int num_kb = a->K / bk;
  const int kpt = a->kpt > 0 ? a->kpt : num_kb;
  if (num_kb % kpt != 0) return set_error(VB_ERR_INVALID, "vb_gemm: K/64 not a multiple of kpt");
  if (a->a_strides[0] != 1 || a->b_strides[0] != 1)
    return set_error(VB_ERR_INVALID, "vb_gemm: reduction dim must be contiguous");

  // N tile
  int bn = a->bn;
  const int m_tiles = (a->M + BM - 1) / BM;
  const int G = a->G1 * a->G2;
  // ---- tile selection -------------------------------------------------------------------
  // Cost model fitted to device timings on B200 (tools/bench_bn.py): operand ingest is bounded per
  // SM (~64 B/clk) and chip-wide (L2 throughput, ~half of 148 x 64 B/clk), every tile pays an
  // epilogue proportional to its width plus a fixed pipeline fill. Units: 128-byte rows.
  bool use2 = false;
  {
    const bool user_bn = bn != 0;
    const int cands[6] = {32, 64, 96, 128, 192, 256};
    double best = 1e30;
    int best_bn = 32;
    const int n_cap = (a->N + 31) / 32 * 32;
    for (int i = 0; i < 6; ++i) {
      const int c = cands[i];
      if (user_bn && c != bn) continue;
      if (c > n_cap && c != 32) continue;
      const double tiles = (double)m_tiles * ((a->N + c - 1) / c) * G;
      const double waves = ceil(tiles / 148.0);
      const double u = (double)num_kb * (128 + c);
      const double cost = fmax(waves * u, tiles * u / 74.0) + waves * (40.0 * c + 1500.0);
      if (cost < best) { best = cost; best_bn = c; }
    }
    // 2-CTA pairs (cta_group::2): each CTA loads only half of the B tile. Non-grouped problems only.
    if (g_gemm2_mode > 0 && !user_bn && !a->tile_group && m_tiles >= 2 && a->N >= 128) {
      const int c2s[3] = {128, 192, 256};
      for (int i = 0; i < 3; ++i) {
        const int c = c2s[i];
        if (c > n_cap) continue;
        const double ctas = 2.0 * ((m_tiles + 1) / 2) * ((a->N + c - 1) / c) * G;
        const double waves = ceil(ctas / 148.0);
        const double u = (double)num_kb * (128 + c / 2);
        const double cost = fmax(waves * u, ctas * u / 74.0) + waves * (40.0 * c + 2100.0);
        if (cost < best || g_gemm2_mode == 2) {
          if (g_gemm2_mode == 2 && use2 && cost >= best) continue;
          best = cost; best_bn = c; use2 = true;
        }
      }
    }
    bn = best_bn;
  }
  if (a->mode == VB_EPI_QKV_ROPE) {
    if (!a->k_out || !a->vt_out || a->head_dim % 32 != 0 ||
        a->N != (3 - a->qkv_sec0) * a->heads * a->head_dim || a->qkv_sec0 < 0 || a->qkv_sec0 > 1 || (a->qkv_sec0 == 0 && !a->q_out) || a->T <= 0 || a->M % a->T != 0)
      return set_error(VB_ERR_INVALID, "vb_gemm: bad QKV_ROPE arguments");
    if ((bn == 96 || bn == 192) && (a->head_dim % 32) != 0) { bn = 64; use2 = false; }
  }
  if (a->swiglu && (a->N % 2 != 0))
    return set_error(VB_ERR_INVALID, "vb_gemm: swiglu needs an even N");

  CUtensorMap tmA, tmB;
  {
    uint32_t box[4] = {static_cast<uint32_t>(bk), BM, 1, 1};
    int rc = encode_map(&tmA, a->a, 4, a->a_dims, a->a_strides, box, tf32);
    if (rc) return rc;
    uint32_t boxb[3] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(use2 ? bn / 2 : bn), 1};
    rc = encode_map(&tmB, a->b, 3, a->b_dims, a->b_strides, boxb, tf32);
    if (rc) return rc;
  }

  GemmKParams p{};
  p.M = a->M; p.N = a->N; p.K = a->K; p.G1 = a->G1;
  p.dbg_shift = g_dbg_shift;
  p.num_kb = num_kb; p.kpt = kpt; p.dil = a->dil; p.pad = a->pad;
  p.a_c2_g1 = a->a_c2_g1; p.a_c2_g2 = a->a_c2_g2; p.a_c3_g1 = a->a_c3_g1; p.a_c3_g2 = a->a_c3_g2;
  p.b_c2_g1 = a->b_c2_g1; p.b_c2_g2 = a->b_c2_g2; p.b_c2_tile = a->b_c2_tile;
  p.tile_group = a->tile_group; p.m_tiles_used = a->m_tiles_used;
  p.mode = a->mode; p.alpha = a->alpha; p.swiglu = a->swiglu;
  p.bias = a->bias; p.bias_g1 = a->bias_g1; p.bias_tile = a->bias_tile;
  p.act = a->act; p.act_slope = a->act_slope;
  p.row_scale = a->row_scale; p.gate = a->gate;
  p.gate_rows = a->gate_rows > 0 ? a->gate_rows : 1; p.gate_ld = a->gate_ld; p.gate_g1 = a->gate_g1;
  p.row_map = a->row_map;
  p.row_mul = a->row_mul; p.row_off = a->row_off; p.row_off_g1 = a->row_off_g1;
  p.rows_out = a->rows_out;
  p.residual = a->residual; p.accumulate = a->accumulate;
  p.out_f32 = a->out_f32; p.out_ld = a->out_ld; p.out_g1 = a->out_g1; p.out_g2 = a->out_g2;
  p.out_bf16 = static_cast<__nv_bfloat16*>(a->out_bf16);
  p.outb_f32 = a->outb_f32;
  p.outb_ld = a->outb_ld; p.outb_g1 = a->outb_g1; p.outb_g2 = a->outb_g2;
  p.act_out = a->act_out; p.act_out_slope = a->act_out_slope; p.out_scale = a->out_scale;
  p.rope_cs = a->rope_cs; p.T = a->T; p.heads = a->heads; p.head_dim = a->head_dim; p.Tp = a->Tp; p.qkv_sec0 = a->qkv_sec0;
  p.q_out = static_cast<__nv_bfloat16*>(a->q_out);
  p.k_out = static_cast<__nv_bfloat16*>(a->k_out);
  p.vt_out = static_cast<__nv_bfloat16*>(a->vt_out);
  if (a->mode == VB_EPI_GENERIC) {
    if (!a->out_f32 && !a->out_bf16)
      return set_error(VB_ERR_INVALID, "vb_gemm: no output pointer");
    if (a->rows_out <= 0) return set_error(VB_ERR_INVALID, "vb_gemm: rows_out must be set");
    auto al = [](const void* ptr, int bytes) {
      return (reinterpret_cast<uintptr_t>(ptr) % bytes) == 0;
    };
    p.vec_f32 = (a->out_ld % 4 == 0) && (a->out_g1 % 4 == 0) && (a->out_g2 % 4 == 0) &&
                (!a->out_f32 || al(a->out_f32, 16)) && (!a->residual || al(a->residual, 16));
    p.vec_bf16 = (a->outb_ld % (a->outb_f32 ? 4 : 8) == 0) && (a->outb_g1 % 8 == 0) && (a->outb_g2 % 8 == 0) &&
                 (!a->out_bf16 || al(a->out_bf16, 16));
    p.vec_aux = (!a->bias || (al(a->bias, 16) && a->bias_g1 % 4 == 0 && a->bias_tile % 4 == 0)) &&
                (!a->gate || (al(a->gate, 16) && a->gate_ld % 4 == 0 && a->gate_g1 % 4 == 0));
    if (a->accumulate && !a->out_f32)
      return set_error(VB_ERR_INVALID, "vb_gemm: accumulate needs out_f32");
    if (a->residual && a->out_ld <= 0)
      return set_error(VB_ERR_INVALID, "vb_gemm: residual needs out_ld (it shares the fp32 output addressing)");
  }

  p.m_tiles = m_tiles;
  p.n_tiles = (a->N + bn - 1) / bn;
  p.num_tiles = p.m_tiles * p.n_tiles * G;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  // haloed small-C conv path (see vb_conv_halo_kernel)
  if (g_conv_halo > 0 && !tf32 && !use2 && !a->tile_group && a->mode == VB_EPI_GENERIC && kpt == 1 && num_kb >= 2 &&
      num_kb <= 16 && a->N <= 64 && a->a_dims[0] <= 64 && a->dil >= 1 && !a->swiglu && a->G2 == 1 &&
      a->a_c2_g2 == 0 && a->a_c3_g1 == 0 && a->a_c3_g2 == 0 && a->b_c2_g1 == 0 && a->b_c2_g2 == 0 && m_tiles >= 64) {
    HaloParams hp{};
    hp.taps = num_kb;
    hp.row_shift = a->dil;
    hp.rows = BM + (num_kb - 1) * a->dil;
    hp.rows8 = (hp.rows + 7) / 8 * 8;
    hp.n_ksteps = (int)((a->a_dims[0] + 15) / 16);
    hp.base_off = g_conv_halo == 3;  // (experiment: base_offset must stay 0 — the 128B swizzle is a pure function of the smem address)
    const int hbn = a->N <= 32 ? 32 : 64;
    const int w_bytes = ((hp.taps * hbn * 128 + 1023) / 1024) * 1024;
    if (hp.rows <= 256) {
      int stages = (200 * 1024 - w_bytes) / (hp.rows8 * 128);
      if (stages > 6) stages = 6;
      if (stages >= 2) {
        hp.stages = stages;
        CUtensorMap tmA2, tmB2;
        uint32_t boxa[4] = {64, static_cast<uint32_t>(hp.rows), 1, 1};
        int rc = encode_map(&tmA2, a->a, 4, a->a_dims, a->a_strides, boxa, false);
        if (rc) return rc;
        uint32_t boxb2[3] = {64, static_cast<uint32_t>(hbn), 1};
        rc = encode_map(&tmB2, a->b, 3, a->b_dims, a->b_strides, boxb2, false);
        if (rc) return rc;
        GemmKParams ph = p;
        ph.n_tiles = 1;
        ph.num_tiles = m_tiles * G;
        const int smem_bytes = w_bytes + stages * hp.rows8 * 128 + 1024 + 512;
        const int grid_h = ph.num_tiles < num_sms ? ph.num_tiles : num_sms;
        cudaError_t e;
        auto go = [&](auto kern) {
          cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
          if (err == cudaSuccess) err = launch_k(kern, dim3(grid_h), dim3(kThreads), smem_bytes, stream, tmA2, tmB2, ph, hp);
          return err;
        };
        const bool k2 = hp.n_ksteps <= 2;
        if (hbn == 32) e = k2 ? go(vb_conv_halo_kernel<32, 2>) : go(vb_conv_halo_kernel<32, 4>);
        else e = k2 ? go(vb_conv_halo_kernel<64, 2>) : go(vb_conv_halo_kernel<64, 4>);
        if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "vb_gemm (conv halo) launch failed: %s", cudaGetErrorString(e));
        return check_launch("vb_conv_halo");
      }
    }
  }
  if (use2) {
    const long long pairs = (long long)((m_tiles + 1) / 2) * p.n_tiles * G;
    int clusters = (int)(pairs < num_sms / 2 ? pairs : num_sms / 2);
    const int g2 = clusters * 2;
    if (bn == 256) return tf32 ? launch2<256, true>(tmA, tmB, p, g2, stream) : launch2<256, false>(tmA, tmB, p, g2, stream);
    if (bn == 192) return tf32 ? launch2<192, true>(tmA, tmB, p, g2, stream) : launch2<192, false>(tmA, tmB, p, g2, stream);
    return tf32 ? launch2<128, true>(tmA, tmB, p, g2, stream) : launch2<128, false>(tmA, tmB, p, g2, stream);
  }
  dim3 grid(p.num_tiles < num_sms ? p.num_tiles : num_sms, 1, 1);
  switch (bn) {
    case 32: return tf32 ? launch<32, true>(tmA, tmB, p, grid, stream) : launch<32, false>(tmA, tmB, p, grid, stream);
    case 64: return tf32 ? launch<64, true>(tmA, tmB, p, grid, stream) : launch<64, false>(tmA, tmB, p, grid, stream);
    case 96: return tf32 ? launch<96, true>(tmA, tmB, p, grid, stream) : launch<96, false>(tmA, tmB, p, grid, stream);
    case 128: return tf32 ? launch<128, true>(tmA, tmB, p, grid, stream) : launch<128, false>(tmA, tmB, p, grid, stream);
    case 192: return tf32 ? launch<192, true>(tmA, tmB, p, grid, stream) : launch<192, false>(tmA, tmB, p, grid, stream);
    case 256: return tf32 ? launch<256, true>(tmA, tmB, p, grid, stream) : launch<256, false>(tmA, tmB, p, grid, stream);
    default: return set_error(VB_ERR_INVALID, "vb_gemm: unsupported N tile %d", bn);
  }
}
