// vb_resblock_pair — one HiFi-GAN ResBlock1 conv PAIR in a single persistent tcgen05 kernel
// (reference vocoder/hifigan/modules/hifigan.py:51-58):
//
//     x' = x + conv2_{k,1}( lrelu( conv1_{k,d}( lrelu(x) ) + b1 ) ) + b2
//
// with every activation in HBM held ONLY as y = lrelu(x) in fp16 (see vocoder.py): the input tile is
// directly conv1's A operand, the residual x is recovered from the same shared-memory tile
// (x = y >= 0 ? y : y / slope) and the result is written back as y' = lrelu(x'). Per element the pair
// moves 2 B in + 2 B out of HBM (the two separate conv launches moved 10, with an fp32 residual stream 16),
// and the intermediate t1 never leaves the SM.
//
// One tile = R = 128 - (k-1) output rows (time steps) x C channels, C in {32, 64}:
//   TMA     : y rows [t0-h2-h1, t0-h2+128+h1)  (h1 = (k-1)/2*d, h2 = (k-1)/2; zero-filled outside [0,T))
//   MMA #1  : conv1 as k taps of UMMA 128 x C x C — tap j reads the SAME tile through a descriptor shifted
//             by j*d rows (the 128-byte swizzle is a pure function of the smem address) -> acc1 (TMEM)
//   epi #1  : acc1 + b1 -> lrelu -> fp16 -> swizzled smem tile t1 (rows outside [0,T) forced to 0: conv2
//             zero-pads ITS input, hifigan.py:23-24)
//   MMA #2  : conv2 as k taps over t1 (shift j rows) -> acc2 (TMEM); rows >= R of the tile are discarded
//   epi #2  : acc2 + b2 + x -> lrelu -> fp16 -> global
// Both convs' weights (2 x k taps x [C x 64] fp16) stay resident in shared memory for the lifetime of the
// persistent CTA. Warp roles (576 threads): 0 TMA producer, 1 MMA issuer (single thread), 2-9 epilogue #1,
// 10-17 epilogue #2 (two warps per TMEM lane quadrant, each thread = one accumulator row x half the channels;
// no transposes: the A-operand layout of t1 and the channels-last output are both row-contiguous).
#include <cstdlib>
#include "vb_internal.h"
#include "vb_ptx.cuh"

namespace vb {

struct ResPairParams {
  int T, B, C, k, dil, h1, h2, R;
  int y_rows, y_rows8, t1_rows8, n_t, num_tiles, NS, NT;
  float slope, inv_slope;
  const float* b1;
  const float* b2;
  const __half* in;
  __half* out;
  long long* trace;  // debug (vb_resblock_set_trace): clock64 stamps of CTA 0, 16 per tile
};
static long long* g_pair_trace = nullptr;

template <int C>
__global__ void __launch_bounds__(576, 1)
resblock_pair_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const ResPairParams p) {
  constexpr int KSTEPS = C / 16;
  constexpr int kWTap = C * 128;  // bytes of one tap's [C x 64] weight block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int w_bytes = 2 * p.k * kWTap;  // multiple of 4096
  const int y_slot = p.y_rows8 * 128, t_slot = p.t1_rows8 * 128;
  uint8_t* smem_w = smem;
  uint8_t* smem_y = smem_w + w_bytes;
  uint8_t* smem_t = smem_y + p.NS * y_slot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_t + p.NT * t_slot);
  uint64_t* w_full = bars;           // [1]
  uint64_t* y_full = bars + 1;       // [4]  input ring (up to 4 stages: the loop is TMA-latency bound otherwise)
  uint64_t* y_empty = bars + 5;      // [4]
  uint64_t* a1_full = bars + 9;      // [2]
  uint64_t* a1_empty = bars + 11;    // [2]
  uint64_t* t1_full = bars + 13;     // [2]
  uint64_t* t1_empty = bars + 15;    // [2]
  uint64_t* a2_full = bars + 17;     // [2]
  uint64_t* a2_empty = bars + 19;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  float* s_bias = reinterpret_cast<float*>(bars + 22);  // [2][C]

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    mbar_init(w_full, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], 1);        // conv1's commit (the only reader of the input stage)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a1_full[i], 1);
      mbar_init(&a1_empty[i], 256);
      mbar_init(&t1_full[i], 256);
      mbar_init(&t1_empty[i], 1);
      mbar_init(&a2_full[i], 1);
      mbar_init(&a2_empty[i], 256);
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc<4 * C>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 2 * C) {
    const int i = threadIdx.x - 64;
    s_bias[i] = i < C ? p.b1[i] : p.b2[i - C];
  }
  __syncthreads();

  const int first = blockIdx.x, step = gridDim.x;
  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(w_full, w_bytes);
      for (int j = 0; j < p.k; ++j) {
        tma_load_3d(smem_w + j * kWTap, &tmW1, w_full, 0, 0, j);
        tma_load_3d(smem_w + (p.k + j) * kWTap, &tmW2, w_full, 0, 0, j);
      }
      int i = 0;
      for (int t = first; t < p.num_tiles; t += step, ++i) {
        const int s = i % p.NS, use = i / p.NS;
        const int b = t / p.n_t, tt = t - b * p.n_t;
        mbar_wait(&y_empty[s], (use & 1) ^ 1);
        mbar_expect_tx(&y_full[s], p.y_rows * 128);
        tma_load_3d(smem_y + s * y_slot, &tmY, &y_full[s], 0, tt * p.R - p.h2 - p.h1, b);
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_fp16(128, C);
      mbar_wait(w_full, 0);
      const uint64_t w1_desc = make_kmajor_sw128_desc(smem_u32(smem_w));
      const uint64_t w2_desc = make_kmajor_sw128_desc(smem_u32(smem_w + p.k * kWTap));
      constexpr uint64_t w_step = kWTap / 16;
      int n_local = 0;
      for (int t = first; t < p.num_tiles; t += step) ++n_local;
      const bool trm = p.trace && blockIdx.x == 0;
      auto mma1 = [&](int i) {
        const int s = i % p.NS, acc = i & 1;
        if (trm && i < 12) p.trace[i * 16 + 0] = clock64();
        mbar_wait(&y_full[s], (i / p.NS) & 1);
        if (trm && i < 12) p.trace[i * 16 + 1] = clock64();
        mbar_wait(&a1_empty[acc], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * C;
        uint64_t da = make_kmajor_sw128_desc(smem_u32(smem_y + s * y_slot));
        uint64_t db = w1_desc;
        const uint64_t da_step = static_cast<uint64_t>(p.dil * 8);  // dil rows * 128 B / 16
        for (int j = 0; j < p.k; ++j) {
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks) umma_fp16(d, da + 2 * ks, db + 2 * ks, idesc, (j > 0 || ks > 0) ? 1u : 0u);
          da += da_step;
          db += w_step;
        }
        umma_commit(&a1_full[acc]);
        umma_commit(&y_empty[s]);
        if (trm && i < 12) p.trace[i * 16 + 2] = clock64();
      };
      auto mma2 = [&](int i) {
        const int tb = i % p.NT, acc = i & 1;
        if (trm && i < 12) p.trace[i * 16 + 3] = clock64();
        mbar_wait(&t1_full[tb], (i / p.NT) & 1);
        if (trm && i < 12) p.trace[i * 16 + 4] = clock64();
        mbar_wait(&a2_empty[acc], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + 2 * C + acc * C;
        uint64_t da = make_kmajor_sw128_desc(smem_u32(smem_t + tb * t_slot));
        uint64_t db = w2_desc;
        for (int j = 0; j < p.k; ++j) {
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks) umma_fp16(d, da + 2 * ks, db + 2 * ks, idesc, (j > 0 || ks > 0) ? 1u : 0u);
          da += 8;  // one row
          db += w_step;
        }
        umma_commit(&a2_full[acc]);
        umma_commit(&t1_empty[tb]);
        if (trm && i < 12) p.trace[i * 16 + 5] = clock64();
      };
      // Issue order: conv1 runs TWO tiles ahead of conv2. The chain MMA1(i) -> epilogue #1 -> MMA2(i) carries
      // ~3 k cycles of fixed latency (TMEM read, smem write, proxy fence, barrier wake-ups); with conv1 only one
      // tile ahead the tensor pipe idled for most of it (measured: 3.75 k cycles per tile + 65 per MMA).
      if (n_local > 0) mma1(0);
      if (n_local > 1) mma1(1);
      for (int i = 0; i < n_local; ++i) {
        mma2(i);
        if (i + 2 < n_local) mma1(i + 2);
      }
    }
  } else if (warp_idx < 10) {
    // ===================== epilogue #1: acc1 -> t1 (shared memory), 8 warps =====================
    // Two warps per TMEM lane quadrant, each owning half of the channels of its 32 rows: the epilogues are
    // instruction-issue bound (long dependent chains, few warps per scheduler), so thread-level parallelism —
    // not bytes — sets the pace. Activation math runs on packed half2.
    constexpr int CH = C / 2;            // channels per warp
    const int q = warp_idx & 3, half = (warp_idx - 2) >> 2;
    const int r = q * 32 + lane;         // accumulator row = t1 row
    const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
    const __half2 slope2 = __float2half2_rn(p.slope);
    int i = 0;
    for (int t = first; t < p.num_tiles; t += step, ++i) {
      const int tb = i % p.NT, acc = i & 1;
      const int b = t / p.n_t, tt = t - b * p.n_t;
      const int time = tt * p.R - p.h2 + r;
      const bool inside = time >= 0 && time < p.T;
      const bool tr1 = p.trace && blockIdx.x == 0 && threadIdx.x == 64 && i < 12;
      if (tr1) p.trace[i * 16 + 6] = clock64();
      mbar_wait(&a1_full[acc], (i >> 1) & 1);
      if (tr1) p.trace[i * 16 + 7] = clock64();
      mbar_wait(&t1_empty[tb], ((i / p.NT) & 1) ^ 1);
      tc_fence_after();
      const uint32_t row = smem_u32(smem_t + tb * t_slot) + r * 128;
      uint32_t raw[CH];
      if (CH == 32) tmem_ld32(tmem_base + acc * C + half * CH + lane_sel, *reinterpret_cast<uint32_t(*)[32]>(raw));
      else tmem_ld16(tmem_base + acc * C + half * CH + lane_sel, *reinterpret_cast<uint32_t(*)[16]>(raw));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&a1_empty[acc]);  // accumulator is in registers
#pragma unroll
      for (int ch = 0; ch < CH / 8; ++ch) {  // 8 channels = one 16-byte chunk of the A-operand row
        uint32_t w[4];
        const float4 ba = *reinterpret_cast<const float4*>(&s_bias[half * CH + ch * 8]);
        const float4 bb = *reinterpret_cast<const float4*>(&s_bias[half * CH + ch * 8 + 4]);
        const float bs[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const __half2 v = vb_f2h2(__uint_as_float(raw[ch * 8 + 2 * h]) + bs[2 * h],
                                    __uint_as_float(raw[ch * 8 + 2 * h + 1]) + bs[2 * h + 1]);
          const __half2 a = __hmax2(v, __hmul2(v, slope2));  // leaky relu (slope < 1)
          w[h] = inside ? *reinterpret_cast<const uint32_t*>(&a) : 0u;  // conv2 zero-pads its input
        }
        const int chunk = (half * CH >> 3) + ch;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((chunk ^ (r & 7)) << 4)), "r"(w[0]),
                     "r"(w[1]), "r"(w[2]), "r"(w[3])
                     : "memory");
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core's async proxy
      mbar_arrive(&t1_full[tb]);
      if (tr1) p.trace[i * 16 + 8] = clock64();
    }
  } else {
    // ===================== epilogue #2: acc2 + residual -> global, 8 warps =====================
    constexpr int CH = C / 2;
    const int q = warp_idx & 3, half = (warp_idx - 10) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
    const __half2 slope2 = __float2half2_rn(p.slope), inv2 = __float2half2_rn(p.inv_slope);
    int i = 0;
    for (int t = first; t < p.num_tiles; t += step, ++i) {
      const int acc = i & 1;
      const int b = t / p.n_t, tt = t - b * p.n_t;
      const int time = tt * p.R + r;
      const bool valid = r < p.R && time < p.T;
      // residual row = the centre of the receptive field, read straight from global memory (L2-resident: the
      // TMA has just fetched the same lines), issued before the accumulator wait so its latency is hidden.
      // (Reading it from the shared-memory input stage raced with the producer's refill of that stage.)
      uint4 res[CH / 8];
      const long long goff = ((long long)b * p.T + (valid ? time : 0)) * C + half * CH;
      {
        const uint4* src = reinterpret_cast<const uint4*>(p.in + goff);
#pragma unroll
        for (int ch = 0; ch < CH / 8; ++ch) res[ch] = __ldg(src + ch);
      }
      const bool tr2 = p.trace && blockIdx.x == 0 && threadIdx.x == 320 && i < 12;
      if (tr2) p.trace[i * 16 + 9] = clock64();
      mbar_wait(&a2_full[acc], (i >> 1) & 1);
      if (tr2) p.trace[i * 16 + 10] = clock64();
      tc_fence_after();
      uint32_t raw[CH];
      if (CH == 32) tmem_ld32(tmem_base + 2 * C + acc * C + half * CH + lane_sel, *reinterpret_cast<uint32_t(*)[32]>(raw));
      else tmem_ld16(tmem_base + 2 * C + acc * C + half * CH + lane_sel, *reinterpret_cast<uint32_t(*)[16]>(raw));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&a2_empty[acc]);
      uint4* dst = reinterpret_cast<uint4*>(p.out + goff);
#pragma unroll
      for (int ch = 0; ch < CH / 8; ++ch) {
        const uint32_t rw[4] = {res[ch].x, res[ch].y, res[ch].z, res[ch].w};
        const float4 ba = *reinterpret_cast<const float4*>(&s_bias[C + half * CH + ch * 8]);
        const float4 bb = *reinterpret_cast<const float4*>(&s_bias[C + half * CH + ch * 8 + 4]);
        const float bs[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const __half2 y = *reinterpret_cast<const __half2*>(&rw[h]);
          const float2 x = __half22float2(__hmin2(y, __hmul2(y, inv2)));  // undo the leaky relu: x = y < 0 ? y/slope : y
          const __half2 v = vb_f2h2(__uint_as_float(raw[ch * 8 + 2 * h]) + bs[2 * h] + x.x,
                                    __uint_as_float(raw[ch * 8 + 2 * h + 1]) + bs[2 * h + 1] + x.y);
          const __half2 a = __hmax2(v, __hmul2(v, slope2));
          w[h] = *reinterpret_cast<const uint32_t*>(&a);
        }
        if (valid) dst[ch] = make_uint4(w[0], w[1], w[2], w[3]);
      }
      if (tr2) p.trace[i * 16 + 11] = clock64();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<4 * C>(tmem_base);
  }
}

// out = lrelu_{slope_out}( (x0 + x1 + x2) * scale ), inputs in the activated representation (x = y >= 0 ? y : y * neg)
__global__ void mrf_combine_kernel(const __half* __restrict__ r0, const __half* __restrict__ r1,
                                   const __half* __restrict__ r2, float neg, float scale, float slope_out,
                                   __half* __restrict__ out, long long n8) {
  pdl_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(r0) + i);
    const uint4 b = r1 ? __ldg(reinterpret_cast<const uint4*>(r1) + i) : make_uint4(0, 0, 0, 0);
    const uint4 c = r2 ? __ldg(reinterpret_cast<const uint4*>(r2) + i) : make_uint4(0, 0, 0, 0);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, cw[4] = {c.x, c.y, c.z, c.w};
    uint32_t o[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[h]));
      const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&bw[h]));
      const float2 fc = __half22float2(*reinterpret_cast<const __half2*>(&cw[h]));
      auto inv = [neg](float y) { return y < 0.f ? y * neg : y; };
      float v0 = (inv(fa.x) + inv(fb.x) + inv(fc.x)) * scale;
      float v1 = (inv(fa.y) + inv(fb.y) + inv(fc.y)) * scale;
      v0 = v0 > 0.f ? v0 : v0 * slope_out;
      v1 = v1 > 0.f ? v1 : v1 * slope_out;
      const __half2 hh = vb_f2h2(v0, v1);
      o[h] = *reinterpret_cast<const uint32_t*>(&hh);
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

static int encode3(CUtensorMap* map, const void* base, const uint64_t* dims, const uint64_t* strides_el,
                   const uint32_t* box, const char* what) {
  cuuint64_t gdim[3] = {dims[0], dims[1], dims[2]};
  cuuint64_t gstr[2] = {strides_el[1] * 2, strides_el[2] * 2};
  cuuint32_t bx[3] = {box[0], box[1], box[2]}, es[3] = {1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) || gstr[0] % 16 || gstr[1] % 16)
    return set_error(VB_ERR_INVALID, "vb_resblock_pair: %s not 16-byte aligned", what);
  auto fn = get_tensor_map_encoder();
  if (!fn) return set_error(VB_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(VB_ERR_CUDA, "vb_resblock_pair: cuTensorMapEncodeTiled(%s) failed (%d)", what, (int)r);
  return VB_OK;
}

// shared-memory plan: weights + NS input stages + NT t1 buffers + barriers/bias; 0 stages = does not fit
static void plan_smem(int C, int k, int dil, ResPairParams* p, int* smem_bytes) {
  p->h2 = (k - 1) / 2;
  p->h1 = p->h2 * dil;
  p->R = 128 - 2 * p->h2;
  p->y_rows = 128 + 2 * p->h1;
  p->y_rows8 = (p->y_rows + 7) / 8 * 8;
  p->t1_rows8 = (128 + 2 * p->h2 + 7) / 8 * 8;
  const int w = 2 * k * C * 128, fixed = 1024 /*align*/ + 1024 /*barriers + bias*/;
  const int lim = 227 * 1024;
  p->NS = p->NT = 0;
  for (int ns = 4; ns >= 1 && !p->NS; --ns)
    for (int nt = 2; nt >= 1; --nt)
      if (w + ns * p->y_rows8 * 128 + nt * p->t1_rows8 * 128 + fixed <= lim) { p->NS = ns; p->NT = nt; break; }
  if (const char* f = getenv("VB_RESPAIR_BUFFERS")) {  // debug: "<NS><NT>" caps, e.g. 21 = 2 input stages, 1 t1 buffer
    if (p->NS > f[0] - '0' && f[0] >= '1') p->NS = f[0] - '0';
    if (p->NT > f[1] - '0' && f[1] >= '1') p->NT = f[1] - '0';
  }
  *smem_bytes = w + p->NS * p->y_rows8 * 128 + p->NT * p->t1_rows8 * 128 + fixed;
}

}  // namespace vb

using namespace vb;

extern "C" int vb_resblock_set_trace(void* buf) {
  vb::g_pair_trace = static_cast<long long*>(buf);
  return VB_OK;
}

extern "C" int vb_resblock_pair_fits(int C, int k, int dil) {
  if ((C != 32 && C != 64) || k < 1 || !(k & 1) || dil < 1) return 0;
  ResPairParams p{};
  int smem = 0;
  plan_smem(C, k, dil, &p, &smem);
  return (p.NS > 0 && 128 + 2 * p.h1 <= 256 && p.R >= 64) ? 1 : 0;
}

extern "C" int vb_resblock_pair(const void* y_in, void* y_out, const void* w1, const void* w2, const float* b1,
                                const float* b2, int B, int T, int C, int k, int dil, float slope, void* stream) {
  ProfScope prof_scope(PK_GEMM, stream);
  if (!y_in || !y_out || !w1 || !w2 || !b1 || !b2 || B <= 0 || T <= 0 || slope <= 0.f)
    return set_error(VB_ERR_INVALID, "vb_resblock_pair: null/invalid argument");
  if (y_in == y_out) return set_error(VB_ERR_INVALID, "vb_resblock_pair: in-place operation is not supported (halo rows)");
  if (!vb_resblock_pair_fits(C, k, dil))
    return set_error(VB_ERR_INVALID, "vb_resblock_pair: (C=%d, k=%d, dil=%d) not supported (C in {32,64}, odd k, weights + tiles must fit 227 KB)", C, k, dil);
  ResPairParams p{};
  int smem = 0;
  plan_smem(C, k, dil, &p, &smem);
  p.T = T; p.B = B; p.C = C; p.k = k; p.dil = dil;
  p.slope = slope; p.inv_slope = 1.f / slope;
  p.b1 = b1; p.b2 = b2; p.in = static_cast<const __half*>(y_in); p.out = static_cast<__half*>(y_out);
  p.trace = g_pair_trace;
  p.n_t = (T + p.R - 1) / p.R;
  p.num_tiles = p.n_t * B;
  CUtensorMap tmY, tmW1, tmW2;
  int rc;
  {
    const uint64_t d[3] = {(uint64_t)C, (uint64_t)T, (uint64_t)B}, s[3] = {1, (uint64_t)C, (uint64_t)T * C};
    const uint32_t box[3] = {64, (uint32_t)p.y_rows, 1};
    if ((rc = encode3(&tmY, y_in, d, s, box, "y_in"))) return rc;
  }
  {
    // packed conv weight [C_out, k * 64] (tap-major, C_in zero-padded to 64) viewed as [tap][C_out][64]
    const uint64_t d[3] = {64, (uint64_t)C, (uint64_t)k}, s[3] = {1, (uint64_t)k * 64, 64};
    const uint32_t box[3] = {64, (uint32_t)C, 1};
    if ((rc = encode3(&tmW1, w1, d, s, box, "w1"))) return rc;
    if ((rc = encode3(&tmW2, w2, d, s, box, "w2"))) return rc;
  }
  const int sms = device_sm_count();
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  cudaError_t e;
  if (C == 64) {
    e = cudaFuncSetAttribute(resblock_pair_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = launch_k(resblock_pair_kernel<64>, dim3(grid), dim3(576), smem, static_cast<cudaStream_t>(stream), tmY, tmW1, tmW2, p);
  } else {
    e = cudaFuncSetAttribute(resblock_pair_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = launch_k(resblock_pair_kernel<32>, dim3(grid), dim3(576), smem, static_cast<cudaStream_t>(stream), tmY, tmW1, tmW2, p);
  }
  if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "vb_resblock_pair launch failed: %s", cudaGetErrorString(e));
  return check_launch("vb_resblock_pair");
}

extern "C" int vb_mrf_combine(const void* r0, const void* r1, const void* r2, float neg, float scale, float slope_out,
                              void* out, int64_t n, void* stream) {
  ProfScope prof_scope(PK_EW, stream);
  if (!r0 || !out || n <= 0 || n % 8) return set_error(VB_ERR_INVALID, "vb_mrf_combine: bad arguments (n %% 8 == 0)");
  const long long n8 = n / 8;
  int grid = (int)((n8 + 255) / 256);
  const int cap = device_sm_count() * 8;
  if (grid > cap) grid = cap;
  launch_k(mrf_combine_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const __half*>(r0),
           static_cast<const __half*>(r1), static_cast<const __half*>(r2), neg, scale, slope_out,
           static_cast<__half*>(out), n8);
  return check_launch("vb_mrf_combine");
}
