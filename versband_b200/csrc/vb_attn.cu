// vb_attn_fused — flash-style attention on tcgen05/TMEM/TMA for sm_100a.
//
// One CTA = 128 query rows of one (batch, head). Key tiles of 128 keys stream through 2-stage TMA rings
// (K and V^T separately); per tile:
//   MMA1  S = Q K^T          (UMMA 128x128x16, fp32 S in TMEM, double-buffered)
//   softmax warps: TMEM -> registers, running row maximum, P = 2^(s*c - m) as fp16 -> swizzled smem
//   MMA2  O += P [V | 1]     (UMMA 128x(HD+16)x16 ACCUMULATING in TMEM: the B operand is the V^T tile with one
//                             extra all-ones row, so column HD of O is the softmax denominator — the row sum is
//                             taken by the tensor core from the very fp16 P values it multiplies with V)
// O never leaves TMEM between tiles: the running maximum is only raised when a tile's maximum exceeds it by
// more than 2^8 (P then stays below 256, well inside fp16), in which case the row's accumulator is rescaled in
// TMEM (tcgen05.ld / st) before the tile's P is published — rare after the first tile. The per-tile work of a
// softmax thread is therefore: one TMEM read of S, max, exp2, pack, store. Round 1 read O_j back and folded it
// into registers every tile and summed rows on the CUDA cores (2250 cycles per tile, issue bound).
// Two KV segments with independent softmax (flag_large_dit_moe.py:381-402), separate accumulators:
//   segment 0: self-attention over T keys            out  = O0 / l0
//   segment 1: one tile of L <= 128 "text" keys      out += tanh(gate[h]) * O1 / l1
// With segment 0 disabled the kernel is the MoE cross-attention of vocal2music_moe.py:119.
//
// Warp roles (576 threads): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2..17 softmax/epilogue:
// FOUR threads per query row (one per 32-column group of S and per HD/4 output dims) — the softmax is
// instruction-issue bound, so 16 warps (4 per scheduler) rather than 8 hide its latencies.
#include <math_constants.h>
#include "vb_internal.h"
#include "vb_ptx.cuh"

namespace vb {

struct AttnKParams {
  int T, L, H, n_self;
  int has_text;
  float scale;
  const float* gate;
  __half* out_fp16;
  float* out_f32;
  long long* trace;  // debug (vb_attn_set_trace): clock64 stamps of CTA (0,0,0), 8 per tile + MMA-thread stamps at [64 + 4*j]
};

static long long* g_attn_trace = nullptr;

template <int HD>
struct AttnCfg {
  static constexpr int VR = HD + 16;                   // V^T rows per k-block incl. the ones row (+ 15 zero rows)
  static constexpr int kQBytes = 2 * 16384;            // Q tile: 2 k-blocks of [128 x 64] fp16
  static constexpr int kKBytes = 2 * 16384;            // K tile per stage
  static constexpr int kVKb = VR * 128;                // one k-block of the V^T tile [VR x 64 keys]
  static constexpr int kVBytes = 2 * kVKb;             // per stage
  static constexpr int kSmem = kQBytes + 2 * kKBytes + 2 * kVBytes + 1024 + 256 + 4 * 1024;
  static constexpr int kThreads = 64 + 16 * 32;
};

template <int HD>
__global__ void __launch_bounds__(AttnCfg<HD>::kThreads, 1)
attn_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKs,
                  const __grid_constant__ CUtensorMap tmVs, const __grid_constant__ CUtensorMap tmKt,
                  const __grid_constant__ CUtensorMap tmVt, const AttnKParams p) {
  using Cfg = AttnCfg<HD>;
  constexpr int VR = Cfg::VR;
  constexpr int DPT = HD / 4;  // output dims per thread (four threads per query row)
  static_assert(DPT == 24 || DPT == 16, "head_dim 96 or 64");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::kQBytes;          // [2 stages][2 kb][16384]
  uint8_t* sV = sK + 2 * Cfg::kKBytes;      // [2 stages][2 kb][VR*128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * Cfg::kVBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]  K and V^T tiles ride separate rings: a K slot is free as soon as
  uint64_t* k_empty = bars + 3;   // [2]  its S = Q K^T retires, a whole tile before the V slot
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* p_full = bars + 11;   // [2]  512 softmax threads published P_j
  uint64_t* p_empty = bars + 13;  // [2]  MMA2_j retired: P buffer free, O holds tile j
  uint64_t* o_done = bars + 15;   // [1]
  uint64_t* s_free = bars + 17;   // [1]  the 16 softmax warps have read S_j (single S buffer)
  uint64_t* q_ready = bars + 18;  // [1]  Q copied into tensor memory (A operand of every S = Q K^T)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);  // (bars + 16: 8 bytes)
  float* xch = reinterpret_cast<float*>(bars + 32);  // [2 parities][4 column groups][128 rows] tile maxima

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_self = p.n_self;
  const int n_tiles = n_self + (p.has_text ? 1 : 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 512);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_done, 1);
    mbar_init(s_free, 16);
    mbar_init(q_ready, 4 * (HD / 32));  // the warps that copy Q (one 32-dim group each)
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  if (warp >= 2) {
    // rows HD .. HD+15 of every V^T k-block (TMA only ever writes rows 0..HD-1): one row of ones, fifteen of zeros
    const int i = threadIdx.x - 64;            // 0..511 = (stage, kb, row 0..15, chunk 0..7)
    const int chunk = i & 7, row = (i >> 3) & 15, kb = (i >> 7) & 1, st = i >> 8;
    const uint32_t v = row == 0 ? 0x3C003C00u : 0u;  // fp16 1.0 pairs
    *reinterpret_cast<uint4*>(sV + st * Cfg::kVBytes + kb * Cfg::kVKb + (HD + row) * 128 + chunk * 16) = make_uint4(v, v, v, v);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // tensor-memory map (512 columns): S [0,128) | P [128,192) | O_self [192,192+VR) | O_text [304,304+VR) | Q [416,416+HD/2)
  const uint32_t tmem_S = tmem_base, tmem_P = tmem_base + 128;
  const uint32_t tmem_Os = tmem_base + 192, tmem_Ot = tmem_base + 304, tmem_Q = tmem_base + 416;
  pdl_sync();  // prologue above overlaps the previous kernel

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, HD > 64 ? Cfg::kQBytes : 16384);
      tma_load_4d(sQ, &tmQ, q_full, 0, q0, h, b);
      if (HD > 64) tma_load_4d(sQ + 16384, &tmQ, q_full, 64, q0, h, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        const bool text = j >= n_self;
        const CUtensorMap* mk = text ? &tmKt : &tmKs;
        const CUtensorMap* mv = text ? &tmVt : &tmVs;
        const int key0 = text ? 0 : j * 128;
        uint8_t* k_dst = sK + st * Cfg::kKBytes;
        uint8_t* v_dst = sV + st * Cfg::kVBytes;
        mbar_wait(&k_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], HD > 64 ? Cfg::kKBytes : 16384);
        tma_load_4d(k_dst, mk, &k_full[st], 0, key0, h, b);
        if (HD > 64) tma_load_4d(k_dst + 16384, mk, &k_full[st], 64, key0, h, b);
        mbar_wait(&v_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&v_full[st], 2 * HD * 128);
        tma_load_4d(v_dst, mv, &v_full[st], key0, 0, h, b);
        tma_load_4d(v_dst + Cfg::kVKb, mv, &v_full[st], key0 + 64, 0, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_fp16(128, 128);
      constexpr uint32_t idesc_o = make_idesc_fp16(128, VR);
      mbar_wait(q_ready, 0);
      tc_fence_after();
      auto mma1 = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        if (j > 0) mbar_wait(s_free, (j - 1) & 1);  // the softmax warps hold S_{j-1} in registers
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const int kb = k >> 2, off = (k & 3) * 2;
          const uint64_t db = make_kmajor_sw128_desc(smem_u32(sK + st * Cfg::kKBytes + kb * 16384)) + off;
          umma_fp16_ts(tmem_S, tmem_Q + 8 * k, db, idesc_s, k > 0 ? 1u : 0u);  // Q from tensor memory: only K streams from smem
        }
        umma_commit(&s_full[0]);
        umma_commit(&k_empty[st]);
      };
      mma1(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        if (j + 1 < n_tiles) mma1(j + 1);  // overlaps the exponentials of tile j (waits until S_j has been read)
        const bool trm = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
        if (trm) p.trace[64 + 4 * j] = clock64();
        mbar_wait(&p_full[st], (j >> 1) & 1);
        if (trm) p.trace[64 + 4 * j + 1] = clock64();
        mbar_wait(&v_full[st], (j >> 1) & 1);
        tc_fence_after();
        const bool text = j >= n_self;
        const bool first = text || j == 0;  // first tile of its segment: overwrite the accumulator
        const int keys = text ? p.L : min(128, p.T - j * 128);
        const int ksteps = (keys + 15) >> 4;
        const uint32_t tO = text ? tmem_Ot : tmem_Os;
        const uint32_t tP = tmem_P;
        for (int k = 0; k < ksteps; ++k) {
          const int kb = k >> 2, off = (k & 3) * 2;
          const uint64_t db = make_kmajor_sw128_desc(smem_u32(sV + st * Cfg::kVBytes + kb * Cfg::kVKb)) + off;
          umma_fp16_ts(tO, tP + 8 * k, db, idesc_o, (!first || k > 0) ? 1u : 0u);
        }
        umma_commit(&p_empty[st]);
        umma_commit(&v_empty[st]);
        if (trm) p.trace[64 + 4 * j + 2] = clock64();
      }
      umma_commit(o_done);
    }
  } else {
    const int qd = warp & 3;             // TMEM lane quadrant = rows 32*qd .. 32*qd+31
    const int cg = (warp - 2) >> 2;      // 32-column group of S, HD/4 output dims
    const int row = qd * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kRaise = 8.f;        // the running maximum follows a tile's maximum only beyond 2^8
    const float c1 = p.scale * kLog2e;
    float m_used = -CUDART_INF_F;        // running maximum (log2 domain) the published P values refer to
    if (cg < HD / 32) {
      // Q (the A operand of every S = Q K^T) moves from its swizzled smem tile into tensor memory once: lane = row,
      // column c = dims (2c, 2c+1). The tensor core then fetches only the K tile from shared memory per MMA
      // (8 KB -> 4 KB: the S MMAs were bound by that port, tools/profile_attn.py).
      mbar_wait(q_full, 0);
      const int d0 = cg * 32;  // 32 dims = 64 bytes = 4 chunks of k-block d0 / 64
      const uint8_t* qrow = sQ + (d0 >> 6) * 16384 + row * 128;
      uint32_t w[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(qrow + (((((d0 & 63) >> 3) + c) ^ (row & 7)) << 4));
        w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
      }
      tmem_st16(tmem_Q + cg * 16 + lane_sel, w);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_ready);
    }
    for (int j = 0; j < n_tiles; ++j) {
      const bool text = j >= n_self;
      const bool first = text || j == 0;
      if (first) m_used = -CUDART_INF_F;
      const int valid = (text ? p.L : min(128, p.T - j * 128)) - cg * 32;  // valid columns of my group
      const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && row == 0 && cg == 0;
      if (tr) p.trace[j * 8 + 0] = clock64();
      mbar_wait(&s_full[0], j & 1);
      tc_fence_after();
      if (tr) p.trace[j * 8 + 1] = clock64();
      float sv[32];
      {
        uint32_t r0[32];
        tmem_ld32(tmem_S + cg * 32 + lane_sel, r0);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);  // S_j is in registers: S_{j+1} may be computed
#pragma unroll
        for (int i = 0; i < 32; ++i) sv[i] = __uint_as_float(r0[i]);
      }
      if (valid < 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) if (i >= valid) sv[i] = -CUDART_INF_F;
      }
      float tmax = -CUDART_INF_F;
#pragma unroll
      for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, sv[i]);
      float* xm = xch + (j & 1) * 512;
      xm[cg * 128 + row] = tmax;
      named_bar_sync(1 + qd, 128);  // the four warps sharing this lane quadrant
      if (tr) p.trace[j * 8 + 2] = clock64();
      tmax = fmaxf(fmaxf(xm[row], xm[128 + row]), fmaxf(xm[256 + row], xm[384 + row]));
      const float m_tile = tmax * c1;
      // (all four threads of a row take the same decision from the same numbers)
      float resc = 1.f;
      if (first) m_used = m_tile;
      else if (m_tile > m_used + kRaise) { resc = ex2_ftz(m_used - m_tile); m_used = m_tile; }
      if (!first && __any_sync(0xffffffffu, resc != 1.f)) {
        // raise the maximum of some rows of this warp: O (and its denominator column) *= 2^(m_old - m_new), in TMEM.
        // MMA2_{j-1} must have retired first; MMA2_j cannot start before this warp publishes P_j below.
        mbar_wait(&p_empty[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        const uint32_t tO = (text ? tmem_Ot : tmem_Os) + lane_sel;
        if (DPT == 24) {
          uint32_t a[16], c[8];
          tmem_ld16(tO + cg * DPT, a);
          tmem_ld8(tO + cg * DPT + 16, c);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) a[i] = __float_as_uint(__uint_as_float(a[i]) * resc);
#pragma unroll
          for (int i = 0; i < 8; ++i) c[i] = __float_as_uint(__uint_as_float(c[i]) * resc);
          tmem_st16(tO + cg * DPT, a);
          tmem_st8(tO + cg * DPT + 16, c);
        } else {
          uint32_t a[16];
          tmem_ld16(tO + cg * DPT, a);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) a[i] = __float_as_uint(__uint_as_float(a[i]) * resc);
          tmem_st16(tO + cg * DPT, a);
        }
        if (cg == 3) {  // the denominator column block HD .. HD+7
          uint32_t c[8];
          tmem_ld8(tO + HD, c);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) c[i] = __float_as_uint(__uint_as_float(c[i]) * resc);
          tmem_st8(tO + HD, c);
        }
        tmem_st_wait();
        tc_fence_before();
      }
      // p = 2^(s*c1 - m_used) evaluated TWO columns per MUFU operation (ex2.approx.f16x2 on the packed
      // exponents): the exponential unit (16 ops/clk/SM) was the bound of this loop — 1400 of a tile's 2400 cycles
      // (tools/profile_attn.py). The exponent is rounded to fp16 first: its absolute error is <= 2^-11 |x|, i.e. a
      // relative error in p of <= 7e-4 at x = -4 and below the fp16 rounding of p itself for the terms that carry
      // the weight (|x| < 1.5); masked columns are -inf -> exactly 0. The result already is the packed fp16 P.
      // P_j goes straight back into TENSOR MEMORY (64 columns next to S): lane = row,
      // column 16*cg + i = keys 32*cg + 2i, +1 — the A operand of MMA2 is read from there (tcgen05.mma with a TMEM
      // A operand), so P never touches shared memory: no swizzled stores, no proxy fence, and the tensor core's
      // shared-memory port only streams V^T. (Every warp of the quadrant has read its S columns: named barrier above.)
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float x0 = fmaf(sv[2 * i], c1, -m_used), x1 = fmaf(sv[2 * i + 1], c1, -m_used);
        uint32_t xx;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(xx) : "f"(x1), "f"(x0));  // (-inf stays -inf)
        asm("ex2.approx.f16x2 %0, %1;" : "=r"(w[i]) : "r"(xx));
      }
      if (j >= 1) {  // single P buffer: MMA2_{j-1} must have consumed P_{j-1}
        mbar_wait(&p_empty[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
      }
      tmem_st16(tmem_P + cg * 16 + lane_sel, w);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[j & 1]);
      if (tr) p.trace[j * 8 + 5] = clock64();
    }
    // ---- epilogue: out = O_self / l_self + tanh(gate) * O_text / l_text
    mbar_wait(o_done, 0);
    tc_fence_after();
    float acc[DPT];
#pragma unroll
    for (int i = 0; i < DPT; ++i) acc[i] = 0.f;
    auto add_segment = [&](uint32_t tO, float g) {
      uint32_t c[8];
      tmem_ld8(tO + lane_sel + HD, c);  // column HD = sum of the row's P values (the ones row of [V | 1])
      uint32_t a[16];
      tmem_ld16(tO + lane_sel + cg * DPT, a);
      tmem_ld_wait();
      const float f = g / __uint_as_float(c[0]);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(__uint_as_float(a[i]), f, acc[i]);
      if (DPT == 24) {
        uint32_t d[8];
        tmem_ld8(tO + lane_sel + cg * DPT + 16, d);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[16 + i] = fmaf(__uint_as_float(d[i]), f, acc[16 + i]);
      }
    };
    if (n_self > 0) add_segment(tmem_Os, 1.f);
    if (p.has_text) add_segment(tmem_Ot, (p.gate && n_self > 0) ? tanhf(p.gate[h]) : (p.gate ? tanhf(p.gate[h]) : 1.f));
    const int t = q0 + row;
    if (t < p.T) {
      const long long o = ((long long)b * p.T + t) * ((long long)p.H * HD) + (long long)h * HD + cg * DPT;
      if (p.out_fp16) {
        uint4* dst = reinterpret_cast<uint4*>(p.out_fp16 + o);
#pragma unroll
        for (int i = 0; i < DPT / 8; ++i) {
          const __half2 a0 = vb_f2h2(acc[8 * i], acc[8 * i + 1]);
          const __half2 a1 = vb_f2h2(acc[8 * i + 2], acc[8 * i + 3]);
          const __half2 a2 = vb_f2h2(acc[8 * i + 4], acc[8 * i + 5]);
          const __half2 a3 = vb_f2h2(acc[8 * i + 6], acc[8 * i + 7]);
          dst[i] = make_uint4(*reinterpret_cast<const uint32_t*>(&a0), *reinterpret_cast<const uint32_t*>(&a1),
                              *reinterpret_cast<const uint32_t*>(&a2), *reinterpret_cast<const uint32_t*>(&a3));
        }
      }
      if (p.out_f32) {
        float4* dst = reinterpret_cast<float4*>(p.out_f32 + o);
#pragma unroll
        for (int i = 0; i < DPT / 4; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static int encode4(CUtensorMap* map, const void* base, const int64_t* dims, const int64_t* strides_el,
                   uint32_t box0, uint32_t box1) {
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t box[4] = {box0, box1, 1, 1}, es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) gdim[i] = (cuuint64_t)dims[i];
  for (int i = 1; i < 4; ++i) {
    gstr[i - 1] = (cuuint64_t)strides_el[i] * 2;
    if (gstr[i - 1] % 16 != 0 || gstr[i - 1] == 0)
      return set_error(VB_ERR_INVALID, "vb_attn_fused: stride %lld of dim %d not a positive multiple of 8 elements",
                       (long long)strides_el[i], i);
  }
  if (reinterpret_cast<uintptr_t>(base) & 15) return set_error(VB_ERR_INVALID, "vb_attn_fused: unaligned operand");
  auto fn = get_tensor_map_encoder();
  if (!fn) return set_error(VB_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(VB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return VB_OK;
}

template <int HD>
static int launch_attn(const vb_attn_args* a, cudaStream_t stream) {
  using Cfg = AttnCfg<HD>;
  CUtensorMap tmQ, tmKs, tmVs, tmKt, tmVt;
  int rc;
  {
    int64_t d[4] = {HD, a->T, a->H, a->B};
    int64_t s[4] = {1, a->q_st, a->q_sh, a->q_sb};
    if ((rc = encode4(&tmQ, a->q, d, s, 64, 128))) return rc;
  }
  const bool self = a->k_self != nullptr;
  const bool text = a->k_text != nullptr;
  if (self) {
    int64_t d[4] = {HD, a->T, a->H, a->B};
    int64_t s[4] = {1, HD, (int64_t)a->T * HD, (int64_t)a->H * a->T * HD};
    if ((rc = encode4(&tmKs, a->k_self, d, s, 64, 128))) return rc;
    int64_t dv[4] = {a->T, HD, a->H, a->B};
    int64_t sv[4] = {1, a->Tp, (int64_t)HD * a->Tp, (int64_t)a->H * HD * a->Tp};
    if ((rc = encode4(&tmVs, a->vt_self, dv, sv, 64, HD))) return rc;
  }
  if (text) {
    int64_t d[4] = {HD, a->L, a->H, a->B};
    int64_t s[4] = {1, HD, (int64_t)a->L * HD, (int64_t)a->H * a->L * HD};
    if ((rc = encode4(&tmKt, a->k_text, d, s, 64, 128))) return rc;
    int64_t dv[4] = {a->L, HD, a->H, a->B};
    int64_t sv[4] = {1, a->Lp, (int64_t)HD * a->Lp, (int64_t)a->H * HD * a->Lp};
    if ((rc = encode4(&tmVt, a->vt_text, dv, sv, 64, HD))) return rc;
  }
  if (!self) { tmKs = tmKt; tmVs = tmVt; }
  if (!text) { tmKt = tmKs; tmVt = tmVs; }
  AttnKParams p{};
  p.T = a->T; p.L = a->L; p.H = a->H;
  p.n_self = self ? (a->T + 127) / 128 : 0;
  p.has_text = text ? 1 : 0;
  p.scale = a->scale; p.gate = a->gate;
  p.out_fp16 = static_cast<__half*>(a->out_fp16);
  p.out_f32 = a->out_f32;
  p.trace = g_attn_trace;
  static PerDevice<bool> attr_flag;
  bool& attr = attr_flag.get();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fused_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "attn_fused smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid((a->T + 127) / 128, a->H, a->B);
  launch_k(attn_fused_kernel<HD>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmem, stream, tmQ, tmKs, tmVs, tmKt, tmVt, p);
  return check_launch("vb_attn_fused");
}

}  // namespace vb

extern "C" int vb_attn_set_trace(void* buf) {
  vb::g_attn_trace = static_cast<long long*>(buf);
  return VB_OK;
}

extern "C" int vb_attn_fused(const vb_attn_args* a, void* stream) {
  using namespace vb;
  ProfScope prof_scope(PK_ATTN_SMALL, stream);
  if (!a || !a->q || (!a->k_self && !a->k_text) || (!a->out_fp16 && !a->out_f32))
    return set_error(VB_ERR_INVALID, "vb_attn_fused: null argument");
  if (a->k_self && (!a->vt_self || a->Tp < a->T || a->Tp % 8))
    return set_error(VB_ERR_INVALID, "vb_attn_fused: bad self K/V^T");
  if (a->k_text && (!a->vt_text || a->L <= 0 || a->L > 128 || a->Lp < a->L || a->Lp % 8))
    return set_error(VB_ERR_INVALID, "vb_attn_fused: text segment needs 0 < L <= 128 and Lp %% 8 == 0");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (a->hd) {
    case 96: return launch_attn<96>(a, s);
    case 64: return launch_attn<64>(a, s);
    default:  // (128 would need 228 KB of shared memory with the double-buffered P tile)
      return set_error(VB_ERR_INVALID, "vb_attn_fused: head_dim %d not instantiated (64/96)", a->hd);
  }
}
