// vb_gemm — tcgen05 / TMEM / TMA GEMM for sm_100a with implicit-conv A addressing and
// fused epilogues. See include/versband.h for the contract.
//
// CTA = 192 threads: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4). One 128 x BN output tile per
// CTA; several CTAs are co-resident per SM (smem permitting) so one CTA's epilogue overlaps
// another's mainloop.
#include "vb_internal.h"
#include "vb_ptx.cuh"

namespace vb {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kThreads = 192;

struct GemmKParams {
  int M, N, K;
  int G1;
  int num_kb, kpt, dil, pad;
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
  long long outb_ld, outb_g1, outb_g2;
  int act_out;
  float act_out_slope, out_scale;
  int vec_f32, vec_bf16;  // alignment allows 16-byte accesses
  // qkv-rope
  const float* rope_cs;
  int T, heads, head_dim, Tp, qkv_sec0;
  __nv_bfloat16* q_out;
  __nv_bfloat16* k_out;
  __nv_bfloat16* vt_out;
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case VB_ACT_SILU:
      return v / (1.f + __expf(-v));
    case VB_ACT_LRELU:
      return v > 0.f ? v : v * slope;
    case VB_ACT_GELU_ERF:
      return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case VB_ACT_TANH:
      return tanhf(v);
    default:
      return v;
  }
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// ------------------------------------------------------------------ generic epilogue
// One thread owns one output row (TMEM lane) and 32 consecutive columns of it.
__device__ __forceinline__ void epilogue_generic(const GemmKParams& p, float (&v)[32], int m,
                                                 int n_base, int g1, int g2, int tg,
                                                 long long orow) {
  const int N = p.N;
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] *= p.alpha;

  if (p.swiglu) {
    // columns (2i, 2i+1) = (w1 x, w3 x)  ->  silu(w1 x) * (w3 x), 16 outputs
    const int n_out0 = n_base >> 1;
    const int n_out_total = N >> 1;
    float o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float a = v[2 * j], b = v[2 * j + 1];
      o[j] = (a / (1.f + __expf(-a))) * b;
    }
    if (p.row_scale) {
      float rs = p.row_scale[m];
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] *= rs;
    }
    if (p.out_bf16) {
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
  if (p.bias) {
    const float* bp = p.bias + g1 * p.bias_g1 + tg * p.bias_tile + n_base;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += __ldg(bp + j);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n_base + j < N) v[j] += __ldg(bp + j);
    }
  }
  if (p.act != VB_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act, p.act_slope);
  }
  if (p.row_scale) {
    float rs = p.row_scale[m];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= rs;
  }
  if (p.gate) {
    const float* gp = p.gate + (long long)(m / p.gate_rows) * p.gate_ld + g1 * p.gate_g1 + n_base;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= __ldg(gp + j);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n_base + j < N) v[j] *= __ldg(gp + j);
    }
  }
  const long long off32 = g1 * p.out_g1 + g2 * p.out_g2 + orow * p.out_ld + n_base;
  if (p.residual) {
    const float* rp = p.residual + off32;
    if (full && p.vec_f32) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 r = reinterpret_cast<const float4*>(rp)[j];
        v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n_base + j < N) v[j] += rp[j];
    }
  }
  if (p.out_f32) {
    float* dst = p.out_f32 + off32;
    if (full && p.vec_f32) {
      if (p.accumulate) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 r = reinterpret_cast<const float4*>(dst)[j];
          v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(dst)[j] =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n_base + j < N) {
          if (p.accumulate) v[j] += dst[j];
          dst[j] = v[j];
        }
    }
  }
  if (p.out_bf16) {
    __nv_bfloat16* dst =
        p.out_bf16 + g1 * p.outb_g1 + g2 * p.outb_g2 + orow * p.outb_ld + n_base;
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = apply_act(v[j] * p.out_scale, p.act_out, p.act_out_slope);
    if (full && p.vec_bf16) {
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
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n_base + j < N) dst[j] = __float2bfloat16(o[j]);
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
template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN <= 32) ? 4 : 4;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int kSmemBytes = kStages * (kABytes + kBBytes) + 1024 /*align*/ + 256;
};

template <int BN>
__global__ void __launch_bounds__(kThreads)
vb_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmKParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;

  const int m_tile = blockIdx.y;
  if (p.m_tiles_used != nullptr && m_tile >= *p.m_tiles_used) return;  // uniform per CTA

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kStages * Cfg::kBBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = m_tile * BM;
  const int g = blockIdx.z;
  const int g1 = g % p.G1;
  const int g2 = g / p.G1;
  const int tg = p.tile_group ? p.tile_group[m_tile] : 0;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = p.num_kb;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int a_c2 = g1 * p.a_c2_g1 + g2 * p.a_c2_g2;
      const int a_c3 = g1 * p.a_c3_g1 + g2 * p.a_c3_g2;
      const int b_c2 = g1 * p.b_c2_g1 + g2 * p.b_c2_g2 + tg * p.b_c2_tile;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], Cfg::kABytes + Cfg::kBBytes);
        const int tap = kb / p.kpt;
        const int c0 = (kb - tap * p.kpt) * BK;
        const int c1 = m0 + tap * p.dil - p.pad;
        tma_load_4d(smem_a + stage * Cfg::kABytes, &tmA, &full_bar[stage], c0, c1, a_c2, a_c3);
        tma_load_3d(smem_b + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * BK, n0, b_c2);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t da = make_kmajor_sw128_desc(smem_u32(smem_a + stage * Cfg::kABytes));
        const uint64_t db = make_kmajor_sw128_desc(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in 16-byte units
          umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs retire
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int q = warp_idx & 3;  // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;
    const int m = m0 + r;
    const bool row_ok = m < p.M;
    long long orow = 0;
    bool store_ok = row_ok;
    if (row_ok && p.mode == VB_EPI_GENERIC) {
      if (p.row_map) orow = p.row_map[m];
      else orow = (long long)m * p.row_mul + p.row_off + (long long)g1 * p.row_off_g1;
      store_ok = (orow >= 0 && orow < p.rows_out);
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t raw[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, raw);
      tmem_ld_wait();
      const int n_base = n0 + c;
      if (store_ok && n_base < p.N) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        if (p.mode == VB_EPI_QKV_ROPE) epilogue_qkv(p, v, m, n_base);
        else epilogue_generic(p, v, m, n_base, g1, g2, tg, orow);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side
static int encode_map(CUtensorMap* map, const void* base, int rank, const int64_t* dims,
                      const int64_t* strides, const uint32_t* box) {
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = static_cast<cuuint64_t>(dims[i]);
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = static_cast<cuuint64_t>(strides[i]) * 2;  // bytes
      if (gstr[i - 1] % 16 != 0 || gstr[i - 1] == 0)
        return set_error(VB_ERR_INVALID, "vb_gemm: tensor stride %lld of dim %d is not a positive multiple of 8 elements",
                         (long long)strides[i], i);
    }
    if (dims[i] <= 0) return set_error(VB_ERR_INVALID, "vb_gemm: non-positive tensor dim");
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return set_error(VB_ERR_INVALID, "vb_gemm: operand base pointer not 16-byte aligned");
  auto fn = get_tensor_map_encoder();
  if (!fn) return set_error(VB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr,
                  bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(VB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return VB_OK;
}

template <int BN>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p,
                  dim3 grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(vb_gemm_kernel<BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess)
      return set_error(VB_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes,
                       cudaGetErrorString(e));
    attr_set = true;
  }
  vb_gemm_kernel<BN><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return set_error(VB_ERR_CUDA, "vb_gemm launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return VB_OK;
}

}  // namespace vb

extern "C" int vb_gemm(const vb_gemm_args* a, void* stream_) {
  vb::ProfScope prof_scope(vb::PK_GEMM, stream_);
  using namespace vb;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a || !a->a || !a->b) return set_error(VB_ERR_INVALID, "vb_gemm: null operand");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0 || a->K % BK != 0)
    return set_error(VB_ERR_INVALID, "vb_gemm: bad problem size M=%d N=%d K=%d (K must be a multiple of 64)",
                     a->M, a->N, a->K);
  if (a->G1 <= 0 || a->G2 <= 0) return set_error(VB_ERR_INVALID, "vb_gemm: bad group grid");
  const int num_kb = a->K / BK;
  const int kpt = a->kpt > 0 ? a->kpt : num_kb;
  if (num_kb % kpt != 0) return set_error(VB_ERR_INVALID, "vb_gemm: K/64 not a multiple of kpt");
  if (a->a_strides[0] != 1 || a->b_strides[0] != 1)
    return set_error(VB_ERR_INVALID, "vb_gemm: reduction dim must be contiguous");

  // N tile
  int bn = a->bn;
  const int m_tiles = (a->M + BM - 1) / BM;
  const int G = a->G1 * a->G2;
  if (bn == 0) {
    auto tiles = [&](int t) { return (long long)m_tiles * ((a->N + t - 1) / t) * G; };
    if (a->N <= 32) bn = 32;
    else if (a->N % 96 == 0 && a->N % 128 != 0 && a->N <= 192) bn = 96;
    else if (tiles(128) >= 2 * 148 && a->N >= 128) bn = 128;
    else if (tiles(64) >= 148 && a->N >= 64) bn = 64;
    else bn = (a->N >= 64 && tiles(32) > 4 * 148) ? 64 : 32;
  }
  if (a->mode == VB_EPI_QKV_ROPE) {
    if (!a->k_out || !a->vt_out || a->head_dim % 32 != 0 ||
        a->N != (3 - a->qkv_sec0) * a->heads * a->head_dim || a->qkv_sec0 < 0 || a->qkv_sec0 > 1 || (a->qkv_sec0 == 0 && !a->q_out) || a->T <= 0 || a->M % a->T != 0)
      return set_error(VB_ERR_INVALID, "vb_gemm: bad QKV_ROPE arguments");
    if (bn == 96 && (a->head_dim % 96) != 0) bn = 64;
  }
  if (a->swiglu && (a->N % 2 != 0))
    return set_error(VB_ERR_INVALID, "vb_gemm: swiglu needs an even N");

  CUtensorMap tmA, tmB;
  {
    uint32_t box[4] = {BK, BM, 1, 1};
    int rc = encode_map(&tmA, a->a, 4, a->a_dims, a->a_strides, box);
    if (rc) return rc;
    uint32_t boxb[3] = {BK, static_cast<uint32_t>(bn), 1};
    rc = encode_map(&tmB, a->b, 3, a->b_dims, a->b_strides, boxb);
    if (rc) return rc;
  }

  GemmKParams p{};
  p.M = a->M; p.N = a->N; p.K = a->K; p.G1 = a->G1;
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
    p.vec_bf16 = (a->outb_ld % 8 == 0) && (a->outb_g1 % 8 == 0) && (a->outb_g2 % 8 == 0) &&
                 (!a->out_bf16 || al(a->out_bf16, 16));
    if ((a->residual || a->accumulate) && !a->out_f32)
      return set_error(VB_ERR_INVALID, "vb_gemm: residual/accumulate need out_f32 addressing");
  }

  dim3 grid((a->N + bn - 1) / bn, m_tiles, G);
  switch (bn) {
    case 32: return launch<32>(tmA, tmB, p, grid, stream);
    case 64: return launch<64>(tmA, tmB, p, grid, stream);
    case 96: return launch<96>(tmA, tmB, p, grid, stream);
    case 128: return launch<128>(tmA, tmB, p, grid, stream);
    default: return set_error(VB_ERR_INVALID, "vb_gemm: unsupported N tile %d", bn);
  }
}
