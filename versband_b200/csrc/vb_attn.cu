// vb_attn_fused — flash-style attention on tcgen05/TMEM/TMA for sm_100a.
//
// One CTA = 128 query rows of one (batch, head). Key tiles of 128 keys stream through a
// 2-stage TMA ring; per tile:
//   MMA1  S = Q K^T        (UMMA 128x128x16, fp32 S in TMEM, double-buffered)
//   softmax warps: TMEM -> registers, online max/sum, P (fp16) -> swizzled smem
//   MMA2  O_j = P V        (UMMA 128xHDx16, B operand = V^T tile), TMEM -> registers,
//                           acc = acc * alpha + O_j   (running output kept in registers)
// P and O are double-buffered: the softmax warps start on tile j+1 right after publishing P_j and
// fold O_j into their accumulators one tile later, so they never wait for MMA2.
// Two KV segments with independent softmax (flag_large_dit_moe.py:381-402):
//   segment 0: self-attention over T keys            out  = O0 / l0
//   segment 1: one tile of L <= 128 "text" keys      out += tanh(gate[h]) * O1 / l1
// With segment 0 disabled the kernel is the MoE cross-attention of vocal2music_moe.py:119.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer,
// warps 2..9 softmax/epilogue (two threads per query row, each owning half of the columns).
#include <math_constants.h>
#include "vb_internal.h"
#include "vb_ptx.cuh"

namespace vb {

static long long* g_attn_trace = nullptr;

struct AttnKParams {
  int T, L, H, n_self;
  int has_text;
  float scale;
  const float* gate;
  __half* out_fp16;
  float* out_f32;
  long long* trace;  // debug: per-tile clock64 stamps of CTA (0,0,0)
};

template <int HD>
struct AttnCfg {
  static constexpr int kQBytes = 2 * 16384;            // Q tile: 2 k-blocks of [128 x 64] fp16
  static constexpr int kKBytes = 2 * 16384;            // K tile per stage
  static constexpr int kVBytes = 2 * HD * 128;         // V^T tile per stage: 2 k-blocks of [HD x 64]
  static constexpr int kPBytes = 2 * 16384;            // P tile (two buffers)
  static constexpr int kSmem = kQBytes + 2 * kKBytes + 2 * kVBytes + 2 * kPBytes + 1024 + 256 + 3 * 1024;  // + row exchange buffers
};

template <int HD>
__global__ void __launch_bounds__(320, 1)
attn_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKs,
                  const __grid_constant__ CUtensorMap tmVs, const __grid_constant__ CUtensorMap tmKt,
                  const __grid_constant__ CUtensorMap tmVt, const AttnKParams p) {
  using Cfg = AttnCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::kQBytes;          // [2 stages][2 kb][16384]
  uint8_t* sV = sK + 2 * Cfg::kKBytes;      // [2 stages][2 kb][HD*128]
  uint8_t* sP = sV + 2 * Cfg::kVBytes;      // [2 buffers][2 kb][16384]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * Cfg::kPBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]  K and V^T tiles ride separate rings: a K slot is free as soon as
  uint64_t* k_empty = bars + 3;   // [2]  its S = Q K^T retires (a whole tile before the V slot), so K_{j+1} is
  uint64_t* v_full = bars + 13;   // [2]  already in flight while tile j-1's P V is still running
  uint64_t* v_empty = bars + 15;  // [2]
  uint64_t* s_full = bars + 5;    // [2]
  uint64_t* p_full = bars + 7;    // [2]
  uint64_t* o_full = bars + 9;    // [2]
  uint64_t* o_empty = bars + 11;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

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
      mbar_init(&p_full[i], 256);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;  // two O buffers of 128 columns
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
        mbar_expect_tx(&v_full[st], Cfg::kVBytes);
        tma_load_4d(v_dst, mv, &v_full[st], key0, 0, h, b);
        tma_load_4d(v_dst + HD * 128, mv, &v_full[st], key0 + 64, 0, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_fp16(128, 128);
      constexpr uint32_t idesc_o = make_idesc_fp16(128, HD);
      mbar_wait(q_full, 0);
      auto mma1 = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const int kb = k >> 2, off = (k & 3) * 2;
          const uint64_t da = make_kmajor_sw128_desc(smem_u32(sQ + kb * 16384)) + off;
          const uint64_t db = make_kmajor_sw128_desc(smem_u32(sK + st * Cfg::kKBytes + kb * 16384)) + off;
          umma_fp16(tmem_base + st * 128, da, db, idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[st]);
        umma_commit(&k_empty[st]);
      };
      mma1(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        if (j + 1 < n_tiles) mma1(j + 1);  // overlaps the softmax of tile j
        mbar_wait(&p_full[st], (j >> 1) & 1);
        if (j >= 2) mbar_wait(&o_empty[st], ((j - 2) >> 1) & 1);  // O_{j-2} has been folded into the registers
        mbar_wait(&v_full[st], (j >> 1) & 1);
        tc_fence_after();
        const int keys = (j >= n_self) ? p.L : min(128, p.T - j * 128);
        const int ksteps = (keys + 15) >> 4;
        for (int k = 0; k < ksteps; ++k) {
          const int kb = k >> 2, off = (k & 3) * 2;
          const uint64_t da = make_kmajor_sw128_desc(smem_u32(sP + st * Cfg::kPBytes + kb * 16384)) + off;
          const uint64_t db = make_kmajor_sw128_desc(smem_u32(sV + st * Cfg::kVBytes + kb * HD * 128)) + off;
          umma_fp16(tmem_O + st * 128, da, db, idesc_o, k > 0 ? 1u : 0u);
        }
        umma_commit(&o_full[st]);
        umma_commit(&v_empty[st]);
      }
    }
  } else {
    // Two threads per query row: half 0 (warps 2..5) owns score columns 0..63 and output dims
    // 0..HD/2-1, half 1 (warps 6..9) the rest. S is read from TMEM once (it is the TMEM read port,
    // 64 B/clk/SM, that bounds this loop), the row maximum is exchanged through shared memory.
    constexpr int HH = HD / 2;
    const int half = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
    float acc[HH];
#pragma unroll
    for (int i = 0; i < HH; ++i) acc[i] = 0.f;
    constexpr float kLog2e = 1.4426950408889634f;
    float m2 = -CUDART_INF_F, l = 0.f;  // running max (log2 domain) and this half's partial sum
    const float c1 = p.scale * kLog2e;
    const float g = p.gate ? tanhf(p.gate[h]) : 1.f;
    const int rsw = row & 7;
    uint8_t* prow = sP + half * 16384 + (row >> 3) * 1024 + rsw * 128;
    // row exchange between the two threads of a query row: [2 halves][128 rows] floats per buffer; the
    // per-tile maxima alternate between buffers 0/1 (a thread may be a tile ahead of its partner's read),
    // buffer 2 carries the segment sums (those exchanges are bracketed by their own barriers)
    float* xch_base = reinterpret_cast<float*>(tmem_slot + 4);
    float* xch = xch_base + 2 * 256;
    // acc += O_k of a tile whose P was published one iteration earlier (f_old/f_new are that tile's factors)
    int pend = -1;
    float pend_old = 0.f, pend_new = 0.f;
    auto fold = [&]() {
      if (pend < 0) return;
      const int pb = pend & 1;
      mbar_wait(&o_full[pb], (pend >> 1) & 1);
      tc_fence_after();
      const uint32_t tO = tmem_O + pb * 128 + half * HH + lane_sel;
      if (HH == 48) {
        uint32_t r0[32], r1[16];
        tmem_ld32(tO, r0);
        tmem_ld16(tO + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = acc[i] * pend_old + __uint_as_float(r0[i]) * pend_new;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[32 + i] = acc[32 + i] * pend_old + __uint_as_float(r1[i]) * pend_new;
      } else {
#pragma unroll
        for (int c = 0; c < HH / 32; ++c) {
          uint32_t r0[32];
          tmem_ld32(tO + c * 32, r0);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[c * 32 + i] = acc[c * 32 + i] * pend_old + __uint_as_float(r0[i]) * pend_new;
        }
      }
      tc_fence_before();
      mbar_arrive(&o_empty[pb]);
      pend = -1;
    };
    for (int j = 0; j < n_tiles; ++j) {
      const bool text = j >= n_self;
      if (text) {
        if (n_self > 0) {  // close the self segment: acc /= (l_half0 + l_half1)
          fold();          // (the last self tile's O first)
          xch[half * 128 + row] = l;
          named_bar_sync(1 + qd, 64);
          const float inv = 1.f / (l + xch[(half ^ 1) * 128 + row]);
          named_bar_sync(1 + qd, 64);
#pragma unroll
          for (int i = 0; i < HH; ++i) acc[i] *= inv;
        }
        m2 = -CUDART_INF_F;
        l = 0.f;
      }
      const int valid = (text ? p.L : min(128, p.T - j * 128)) - half * 64;  // valid columns of my half
      const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && row == 0 && half == 0;
      if (tr) p.trace[j * 8 + 0] = clock64();
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (tr) p.trace[j * 8 + 1] = clock64();
      float sv[64];
      {
        const uint32_t tS = tmem_base + (j & 1) * 128 + half * 64 + lane_sel;
        uint32_t r0[32], r1[32];
        tmem_ld32(tS, r0);
        tmem_ld32(tS + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) { sv[i] = __uint_as_float(r0[i]); sv[32 + i] = __uint_as_float(r1[i]); }
      }
      if (valid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i) if (i >= valid) sv[i] = -CUDART_INF_F;
      }
      float tmax = -CUDART_INF_F;
#pragma unroll
      for (int i = 0; i < 64; ++i) tmax = fmaxf(tmax, sv[i]);
      float* xm = xch_base + (j & 1) * 256;
      xm[half * 128 + row] = tmax;
      named_bar_sync(1 + qd, 64);
      tmax = fmaxf(tmax, xm[(half ^ 1) * 128 + row]);
      if (tr) p.trace[j * 8 + 2] = clock64();
      const float m_new = fmaxf(m2, tmax * c1);
      const float alpha = (m2 == -CUDART_INF_F) ? 0.f : ex2_ftz(m2 - m_new);
      // p = 2^(s*c1 - m_new)  (ftz: masked columns are -inf -> exactly 0)
#pragma unroll
      for (int i = 0; i < 64; ++i) sv[i] = ex2_ftz(fmaf(sv[i], c1, -m_new));
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
#pragma unroll
      for (int i = 0; i < 64; i += 4) { rs0 += sv[i]; rs1 += sv[i + 1]; rs2 += sv[i + 2]; rs3 += sv[i + 3]; }
      l = l * alpha + ((rs0 + rs1) + (rs2 + rs3));
      m2 = m_new;
      uint8_t* prow_j = prow + (j & 1) * Cfg::kPBytes;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint4 w;
        __half2 t0 = vb_f2h2(sv[8 * ch], sv[8 * ch + 1]);
        __half2 t1 = vb_f2h2(sv[8 * ch + 2], sv[8 * ch + 3]);
        __half2 t2 = vb_f2h2(sv[8 * ch + 4], sv[8 * ch + 5]);
        __half2 t3 = vb_f2h2(sv[8 * ch + 6], sv[8 * ch + 7]);
        w.x = *reinterpret_cast<uint32_t*>(&t0); w.y = *reinterpret_cast<uint32_t*>(&t1);
        w.z = *reinterpret_cast<uint32_t*>(&t2); w.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(prow_j + ((ch ^ rsw) << 4)) = w;
      }
      tc_fence_before();
      fence_proxy_async();  // make the generic-proxy smem writes visible to the tensor core
      mbar_arrive(&p_full[j & 1]);
      if (tr) p.trace[j * 8 + 3] = clock64();
      float f_old = alpha, f_new = 1.f;
      if (text) {  // single text tile: l is final after the exchange
        named_bar_sync(1 + qd, 64);  // (both halves are done reading the max exchange)
        xch[half * 128 + row] = l;
        named_bar_sync(1 + qd, 64);
        f_old = 1.f;
        f_new = g / (l + xch[(half ^ 1) * 128 + row]);
      }
      // O_{j-1} is complete by now (MMA2 of tile j-1 ran under this tile's softmax): fold it in, then
      // queue this tile
      fold();
      if (tr) p.trace[j * 8 + 4] = clock64();
      pend = j;
      pend_old = f_old;
      pend_new = f_new;
    }
    fold();
    if (!p.has_text && n_self > 0) {
      xch[half * 128 + row] = l;
      named_bar_sync(1 + qd, 64);
      const float inv = 1.f / (l + xch[(half ^ 1) * 128 + row]);
#pragma unroll
      for (int i = 0; i < HH; ++i) acc[i] *= inv;
    }
    const int t = q0 + row;
    if (t < p.T) {
      const long long o = ((long long)b * p.T + t) * ((long long)p.H * HD) + (long long)h * HD + half * HH;
      if (p.out_fp16) {
        uint4* dst = reinterpret_cast<uint4*>(p.out_fp16 + o);
#pragma unroll
        for (int i = 0; i < HH / 8; ++i) {
          __half2 a0 = vb_f2h2(acc[8 * i], acc[8 * i + 1]);
          __half2 a1 = vb_f2h2(acc[8 * i + 2], acc[8 * i + 3]);
          __half2 a2 = vb_f2h2(acc[8 * i + 4], acc[8 * i + 5]);
          __half2 a3 = vb_f2h2(acc[8 * i + 6], acc[8 * i + 7]);
          dst[i] = make_uint4(*reinterpret_cast<uint32_t*>(&a0), *reinterpret_cast<uint32_t*>(&a1),
                              *reinterpret_cast<uint32_t*>(&a2), *reinterpret_cast<uint32_t*>(&a3));
        }
      }
      if (p.out_f32) {
        float4* dst = reinterpret_cast<float4*>(p.out_f32 + o);
#pragma unroll
        for (int i = 0; i < HH / 4; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
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
  launch_k(attn_fused_kernel<HD>, dim3(grid), dim3(320), Cfg::kSmem, stream, tmQ, tmKs, tmVs, tmKt, tmVt, p);
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
