// Row-wise SIMT kernels of the DiT step: RMSNorm/LayerNorm (+adaLN modulate), row softmax,
// small-KV attention (text cross-attention / MoE cross-attention), timestep embedding.
// All are HBM/L2-bound: one warp per row, 16-byte vector loads, fp32 math, fp16 operand output.
#include <math_constants.h>
#include "vb_internal.h"

namespace vb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------
// norm + modulate:  y = norm(x) [* w] [+ b];  y = y * (1 + scale[batch]) + shift[batch]
// kind 0 = RMSNorm (flag_large_dit_moe.py:52-77), kind 1 = LayerNorm (biased variance).
// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void norm_mod_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                const float* __restrict__ b, const float* __restrict__ shift,
                                const float* __restrict__ scale, long long mod_ld,
                                int rows_per_batch, float eps, __half* __restrict__ out_fp16,
                                float* __restrict__ out_f32, int rows, int D) {
  pdl_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (long long)row * D;
  float s = 0.f, ss = 0.f;
  for (int d = lane * 4; d < D; d += 128) {
    float4 v = *reinterpret_cast<const float4*>(xr + d);
    s += v.x + v.y + v.z + v.w;
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  float mean = 0.f, rstd;
  if (KIND == 0) {
    rstd = rsqrtf(ss / D + eps);
  } else {
    mean = s / D;
    // two-pass variance for accuracy
    float vs = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
      float4 v = *reinterpret_cast<const float4*>(xr + d);
      float a = v.x - mean, bb = v.y - mean, c = v.z - mean, e = v.w - mean;
      vs += a * a + bb * bb + c * c + e * e;
    }
    vs = warp_sum(vs);
    rstd = rsqrtf(vs / D + eps);
  }
  const int bidx = row / rows_per_batch;
  const float* sh = shift ? shift + (long long)bidx * mod_ld : nullptr;
  const float* sc = scale ? scale + (long long)bidx * mod_ld : nullptr;
  for (int d = lane * 4; d < D; d += 128) {
    float4 v = *reinterpret_cast<const float4*>(xr + d);
    float y[4] = {(v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (w) y[j] *= w[d + j];
      if (b) y[j] += b[d + j];
      if (sc) y[j] = y[j] * (1.f + sc[d + j]);
      if (sh) y[j] += sh[d + j];
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + (long long)row * D + d) = make_float4(y[0], y[1], y[2], y[3]);
    if (out_fp16) {
      __half2 p0 = vb_f2h2(y[0], y[1]);
      __half2 p1 = vb_f2h2(y[2], y[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&p0);
      u.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(out_fp16 + (long long)row * D + d) = u;
    }
  }
}


// Register-resident variant (D <= 128 * NV, every pointer 16-byte aligned): the row is read once,
// the modulation vectors are requested together with it, and nothing in the statistics -> output
// chain goes back to memory. With ~10 warps per SM the kernel is bound by its dependent memory round
// trips, so this halves its duration against the re-reading kernel above.
template <int KIND, int NV>
__global__ void norm_mod_reg_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                    const float* __restrict__ b, const float* __restrict__ shift,
                                    const float* __restrict__ scale, long long mod_ld,
                                    int rows_per_batch, float eps, __half* __restrict__ out_fp16,
                                    float* __restrict__ out_f32, int rows, int D) {
  pdl_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (long long)row * D;
  const int bidx = row / rows_per_batch;
  const float* sh = shift ? shift + (long long)bidx * mod_ld : nullptr;
  const float* sc = scale ? scale + (long long)bidx * mod_ld : nullptr;
  float4 v[NV], wv[NV], bv[NV], scv[NV], shv[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int d = (k * 32 + lane) * 4;
    const bool ok = d < D;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f), one = make_float4(1.f, 1.f, 1.f, 1.f);
    v[k] = ok ? *reinterpret_cast<const float4*>(xr + d) : z;
    wv[k] = (ok && w) ? __ldg(reinterpret_cast<const float4*>(w + d)) : one;
    bv[k] = (ok && b) ? __ldg(reinterpret_cast<const float4*>(b + d)) : z;
    scv[k] = (ok && sc) ? *reinterpret_cast<const float4*>(sc + d) : z;
    shv[k] = (ok && sh) ? *reinterpret_cast<const float4*>(sh + d) : z;
  }
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    s += v[k].x + v[k].y + v[k].z + v[k].w;
    ss += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  float mean = 0.f, rstd;
  if (KIND == 0) {
    rstd = rsqrtf(ss / D + eps);
  } else {
    mean = s / D;
    float vs = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if ((k * 32 + lane) * 4 < D) {
        const float a = v[k].x - mean, bb = v[k].y - mean, c = v[k].z - mean, e = v[k].w - mean;
        vs += a * a + bb * bb + c * c + e * e;
      }
    }
    vs = warp_sum(vs);
    rstd = rsqrtf(vs / D + eps);
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int d = (k * 32 + lane) * 4;
    if (d >= D) continue;
    float y[4] = {(v[k].x - mean) * rstd, (v[k].y - mean) * rstd, (v[k].z - mean) * rstd, (v[k].w - mean) * rstd};
    const float wj[4] = {wv[k].x, wv[k].y, wv[k].z, wv[k].w}, bj[4] = {bv[k].x, bv[k].y, bv[k].z, bv[k].w};
    const float scj[4] = {scv[k].x, scv[k].y, scv[k].z, scv[k].w}, shj[4] = {shv[k].x, shv[k].y, shv[k].z, shv[k].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // same operation order as norm_mod_kernel
      if (w) y[j] *= wj[j];
      if (b) y[j] += bj[j];
      if (sc) y[j] = y[j] * (1.f + scj[j]);
      if (sh) y[j] += shj[j];
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + (long long)row * D + d) = make_float4(y[0], y[1], y[2], y[3]);
    if (out_fp16) {
      __half2 p0 = vb_f2h2(y[0], y[1]);
      __half2 p1 = vb_f2h2(y[2], y[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&p0);
      u.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(out_fp16 + (long long)row * D + d) = u;
    }
  }
}

// ---------------------------------------------------------------------------------------
// row softmax: P[r, :n] = softmax(S[r, :n]) (fp32 in, fp16 out), P[r, n:ldp] = 0
// ---------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(const float* __restrict__ S, long long lds,
                                    __half* __restrict__ P, long long ldp, long long rows,
                                    int n) {
  pdl_sync();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* s = S + row * lds;
  float m = -CUDART_INF_F;
  for (int c = lane; c < n; c += 32) m = fmaxf(m, s[c]);
  m = warp_max(m);
  float sum = 0.f;
  for (int c = lane; c < n; c += 32) sum += __expf(s[c] - m);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  __half* p = P + row * ldp;
  for (int c = lane; c < ldp; c += 32)
    p[c] = vb_f2h(c < n ? __expf(s[c] - m) * inv : 0.f);
}

// ---------------------------------------------------------------------------------------
// small-KV attention: ctx[t, h, :] = softmax(q[t,h,:] . K[:,h,:]^T * scale) V[:,h,:]
// out[row, h*hd + d] = (base ? base[...] : 0) + (gate ? tanh(gate[h]) : 1) * ctx
// K, V fp32 [B, L, H*hd] staged in shared memory (rows padded by 1 float: no bank conflicts).
// flag_large_dit_moe.py:388-402 (text cross-attention, 80 keys) and vocal2music_moe.py:119.
// ---------------------------------------------------------------------------------------
template <int HD>
__global__ void attn_smallkv_kernel(const __half* __restrict__ q, long long q_sb,
                                    long long q_st, long long q_sh, const float* __restrict__ K,
                                    const float* __restrict__ V, long long kv_ld, int L, int H, int T, float scale,
                                    const float* __restrict__ gate, const float* __restrict__ base,
                                    __half* __restrict__ out_fp16,
                                    float* __restrict__ out_f32, int tokens_per_block) {
  pdl_sync();
  extern __shared__ float sm[];
  constexpr int HP = HD + 1;
  float* sK = sm;
  float* sV = sm + L * HP;
  float* sQ = sV + L * HP;  // [warps][HD]
  const int b = blockIdx.z, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const long long kv_off = ((long long)b * L) * kv_ld + h * HD;
  for (int i = threadIdx.x; i < L * HD; i += blockDim.x) {
    int l = i / HD, d = i - l * HD;
    sK[l * HP + d] = K[kv_off + (long long)l * kv_ld + d];
    sV[l * HP + d] = V[kv_off + (long long)l * kv_ld + d];
  }
  __syncthreads();
  const float g = gate ? tanhf(gate[h]) : 1.f;
  const int t0 = blockIdx.x * tokens_per_block;
  float* myq = sQ + warp * HD;
  constexpr int MAXL = 4;  // up to 128 keys
  for (int tt = warp; tt < tokens_per_block; tt += nw) {
    const int t = t0 + tt;
    if (t >= T) break;
    const __half* qp = q + b * q_sb + (long long)t * q_st + h * q_sh;
    for (int d = lane; d < HD; d += 32) myq[d] = __half2float(qp[d]);
    __syncwarp();
    float sc[MAXL];
    float m = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < MAXL; ++j) {
      const int l = lane + 32 * j;
      float acc = -CUDART_INF_F;
      if (l < L) {
        acc = 0.f;
        const float* kr = sK + l * HP;
#pragma unroll 8
        for (int d = 0; d < HD; ++d) acc = fmaf(myq[d], kr[d], acc);
        acc *= scale;
      }
      sc[j] = acc;
      m = fmaxf(m, acc);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < MAXL; ++j) {
      sc[j] = (lane + 32 * j < L) ? __expf(sc[j] - m) : 0.f;
      sum += sc[j];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float o[(HD + 31) / 32];
#pragma unroll
    for (int i = 0; i < (HD + 31) / 32; ++i) o[i] = 0.f;
#pragma unroll
    for (int j = 0; j < MAXL; ++j) {
      for (int src = 0; src < 32; ++src) {
        const int l = src + 32 * j;
        if (l >= L) break;
        const float p = __shfl_sync(0xffffffffu, sc[j], src) * inv;
        const float* vr = sV + l * HP;
#pragma unroll
        for (int i = 0; i < (HD + 31) / 32; ++i) {
          const int d = lane + 32 * i;
          if (d < HD) o[i] = fmaf(p, vr[d], o[i]);
        }
      }
    }
    const long long orow = ((long long)b * T + t) * (H * HD) + h * HD;
#pragma unroll
    for (int i = 0; i < (HD + 31) / 32; ++i) {
      const int d = lane + 32 * i;
      if (d < HD) {
        float val = g * o[i];
        if (base) val += base[orow + d];
        if (out_fp16) out_fp16[orow + d] = vb_f2h(val);
        if (out_f32) out_f32[orow + d] = val;
      }
    }
    __syncwarp();
  }
}

// timestep embedding: out[b, :] = cat[cos(t f_i), sin(t f_i)], f_i = exp(-ln(1e4) i / half)
// flag_large_dit_moe.py:109-127
// `freqs` is the host-computed table (identical to the reference's torch.exp result).
__global__ void timestep_embed_kernel(const long long* __restrict__ t, const float* __restrict__ freqs,
                                      float* __restrict__ out_f32,
                                      __half* __restrict__ out_fp16, int B, int dim) {
  pdl_sync();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i - b * half;
  const float a = (float)t[b] * freqs[k];
  const float c = cosf(a), s = sinf(a);
  if (out_f32) { out_f32[(long long)b * dim + k] = c; out_f32[(long long)b * dim + half + k] = s; }
  if (out_fp16) {
    out_fp16[(long long)b * dim + k] = vb_f2h(c);
    out_fp16[(long long)b * dim + half + k] = vb_f2h(s);
  }
}

}  // namespace vb

using namespace vb;

extern "C" int vb_norm_mod(int kind, const float* x, const float* w, const float* b,
                           const float* shift, const float* scale, int64_t mod_ld,
                           int rows_per_batch, float eps, void* out_fp16, float* out_f32, int rows,
                           int D, void* stream) {
  vb::ProfScope prof_scope(vb::PK_NORM, stream);
  if (!x || rows <= 0 || D <= 0 || D % 4 != 0 || (!out_fp16 && !out_f32))
    return set_error(VB_ERR_INVALID, "vb_norm_mod: bad arguments (D must be a multiple of 4)");
  if (rows_per_batch <= 0) rows_per_batch = rows;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
  if (D <= 1024 && al16(x) && al16(w) && al16(b) && al16(shift) && al16(scale) && mod_ld % 4 == 0 &&
      al16(out_f32) && al16(out_fp16)) {
    const int wpb4 = 4;  // more, smaller CTAs: the rows of one DiT step spread over all SMs
    dim3 grid4((rows + wpb4 - 1) / wpb4), block4(wpb4 * 32);
    auto go = [&](auto kern) {
      launch_k(kern, dim3(grid4), dim3(block4), 0, s, x, w, b, shift, scale, mod_ld, rows_per_batch, eps,
               static_cast<__half*>(out_fp16), out_f32, rows, D);
      return check_launch("vb_norm_mod");
    };
    if (D <= 512) return kind == 0 ? go(norm_mod_reg_kernel<0, 4>) : go(norm_mod_reg_kernel<1, 4>);
    if (D <= 768) return kind == 0 ? go(norm_mod_reg_kernel<0, 6>) : go(norm_mod_reg_kernel<1, 6>);
    return kind == 0 ? go(norm_mod_reg_kernel<0, 8>) : go(norm_mod_reg_kernel<1, 8>);
  }
  const int wpb = 8;
  dim3 grid((rows + wpb - 1) / wpb), block(wpb * 32);
  if (kind == 0)
    launch_k(norm_mod_kernel<0>, dim3(grid), dim3(block), 0, s, x, w, b, shift, scale, mod_ld, rows_per_batch, eps,
                                              static_cast<__half*>(out_fp16), out_f32, rows, D);
  else
    launch_k(norm_mod_kernel<1>, dim3(grid), dim3(block), 0, s, x, w, b, shift, scale, mod_ld, rows_per_batch, eps,
                                              static_cast<__half*>(out_fp16), out_f32, rows, D);
  return check_launch("vb_norm_mod");
}

extern "C" int vb_softmax_rows(const float* S, int64_t lds, void* P, int64_t ldp, int64_t rows,
                               int n, void* stream) {
  vb::ProfScope prof_scope(vb::PK_SOFTMAX, stream);
  if (!S || !P || rows <= 0 || n <= 0 || ldp < n)
    return set_error(VB_ERR_INVALID, "vb_softmax_rows: bad arguments");
  const int wpb = 8;
  dim3 grid((unsigned)((rows + wpb - 1) / wpb)), block(wpb * 32);
  launch_k(softmax_rows_kernel, dim3(grid), dim3(block), 0, static_cast<cudaStream_t>(stream), 
      S, lds, static_cast<__half*>(P), ldp, rows, n);
  return check_launch("vb_softmax_rows");
}

extern "C" int vb_attn_smallkv(const void* q, int64_t q_sb, int64_t q_st, int64_t q_sh,
                               const float* K, const float* V, int64_t kv_ld, int B, int T, int H,
                               int hd, int L, float scale, const float* gate, const float* base, void* out_fp16,
                               float* out_f32, void* stream) {
  vb::ProfScope prof_scope(vb::PK_ATTN_SMALL, stream);
  if (!q || !K || !V || L <= 0 || L > 128 || (!out_fp16 && !out_f32))
    return set_error(VB_ERR_INVALID, "vb_attn_smallkv: bad arguments (L must be <= 128)");
  if (hd != 96 && hd != 64 && hd != 128)
    return set_error(VB_ERR_INVALID, "vb_attn_smallkv: head_dim %d not instantiated (64/96/128)", hd);
  const int threads = 256, tpb = 32;
  const size_t smem = (size_t)(2 * L * (hd + 1) + (threads / 32) * hd) * sizeof(float);
  dim3 grid((T + tpb - 1) / tpb, H, B);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto go = [&](auto kern) -> int {
    static PerDevice<size_t> attr_dev;  // one per instantiation of this generic lambda, per device
    size_t& attr_smem = attr_dev.get();
    if (smem > attr_smem) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "attn_smallkv smem attr: %s", cudaGetErrorString(e));
      attr_smem = smem;
    }
    launch_k(kern, dim3(grid), dim3(threads), smem, s, static_cast<const __half*>(q), q_sb, q_st, q_sh, K, V, kv_ld, L,
                                     H, T, scale, gate, base, static_cast<__half*>(out_fp16),
                                     out_f32, tpb);
    return check_launch("vb_attn_smallkv");
  };
  if (hd == 96) return go(attn_smallkv_kernel<96>);
  if (hd == 64) return go(attn_smallkv_kernel<64>);
  return go(attn_smallkv_kernel<128>);
}

extern "C" int vb_timestep_embed(const int64_t* t, const float* freqs, float* out_f32,
                                 void* out_fp16, int B, int dim, void* stream) {
  vb::ProfScope prof_scope(vb::PK_EW, stream);
  if (!t || !freqs || B <= 0 || dim <= 0 || dim % 2) return set_error(VB_ERR_INVALID, "vb_timestep_embed: bad arguments");
  const int n = B * dim / 2;
  launch_k(timestep_embed_kernel, dim3((n + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const long long*>(t), freqs, out_f32, static_cast<__half*>(out_fp16), B, dim);
  return check_launch("vb_timestep_embed");
}

// mean over the token axis: out[b, d] = (1/L) sum_l x[b, l, d]   (vocal2music_moe.py:410-412)
namespace vb {
__global__ void mean_tokens_kernel(const float* __restrict__ x, int L, int D, float* __restrict__ out) {
  pdl_sync();
  const int b = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float s = 0.f;
  for (int l = 0; l < L; ++l) s += x[((long long)b * L + l) * D + d];
  out[(long long)b * D + d] = s / (float)L;
}
}  // namespace vb

extern "C" int vb_mean_tokens(const float* x, int B, int L, int D, float* out, void* stream) {
  vb::ProfScope prof_scope(vb::PK_EW, stream);
  if (!x || !out || B <= 0 || L <= 0) return set_error(VB_ERR_INVALID, "vb_mean_tokens: bad arguments");
  launch_k(mean_tokens_kernel, dim3(dim3((D + 127) / 128, B)), dim3(128), 0, static_cast<cudaStream_t>(stream), x, L, D, out);
  return check_launch("vb_mean_tokens");
}
