// Elementwise / layout / GroupNorm kernels (HBM-bound; coalesced, vectorised where aligned).
#include "vb_internal.h"
#include "vb_ptx.cuh"

namespace vb {

__device__ __forceinline__ float act_f(float v, int act, float slope) {
  switch (act) {
    case VB_ACT_SILU: return v / (1.f + __expf(-v));
    case VB_ACT_LRELU: return v > 0.f ? v : v * slope;
    case VB_ACT_GELU_ERF: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case VB_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// out = act(alpha*a + beta*b); each input row is written `rep` times (nearest-neighbour x2
// upsample of autoencoder1d.py:288 when rep == 2). Row length D % 4 == 0.
__global__ void ew_kernel(const float* __restrict__ a, const float* __restrict__ b, float alpha,
                          float beta, int act, float slope, float* __restrict__ out_f32,
                          __half* __restrict__ out_fp16, long long rows, int D, int rep) {
  pdl_sync();
  const long long nvec = rows * (D / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (D / 4);
    const int c = (int)(i - r * (D / 4)) * 4;
    float4 v = *reinterpret_cast<const float4*>(a + r * D + c);
    float y[4] = {v.x * alpha, v.y * alpha, v.z * alpha, v.w * alpha};
    if (b) {
      float4 w = *reinterpret_cast<const float4*>(b + r * D + c);
      y[0] += beta * w.x; y[1] += beta * w.y; y[2] += beta * w.z; y[3] += beta * w.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = act_f(y[j], act, slope);
    for (int k = 0; k < rep; ++k) {
      const long long o = (r * rep + k) * D + c;
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(y[0], y[1], y[2], y[3]);
      if (out_fp16) {
        __half2 p0 = vb_f2h2(y[0], y[1]);
        __half2 p1 = vb_f2h2(y[2], y[3]);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&p0);
        u.y = *reinterpret_cast<uint32_t*>(&p1);
        *reinterpret_cast<uint2*>(out_fp16 + o) = u;
      }
    }
  }
}

// [B, C, T] fp32 -> [B, T, Cp] fp16 (channels-last, zero padded to Cp), value * scale.
// 32x32 shared-memory transpose: coalesced on both sides.
__global__ void nct_to_ntc_kernel(const float* __restrict__ in, float scale,
                                  __half* __restrict__ out, float* __restrict__ out_f32, int C, int T, int Cp) {
  pdl_sync();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, t = t0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && t < T) ? in[((long long)b * C + c) * T + t] * scale : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int t = t0 + j, c = c0 + threadIdx.x;
    if (t < T && c < Cp) {
      if (out_f32) out_f32[((long long)b * T + t) * Cp + c] = tile[threadIdx.x][j];
      else out[((long long)b * T + t) * Cp + c] = vb_f2h(tile[threadIdx.x][j]);
    }
  }
}

// [B, T, ld] fp32 (first C columns) -> [B, C, T] fp32:  out = alpha*out_prev? no: plain transpose,
// or the fused CFG + Euler update (cfm1_audio.py:160, torchdyn Euler):
//   x[b,c,t] += dt * (e_u + s * (e_c - e_u))       when e_u != nullptr
//   out[b,c,t] = e_c[b,t,c]                          otherwise
__global__ void ntc_to_nct_kernel(const float* __restrict__ e_c, const float* __restrict__ e_u,
                                  long long ld, float s, float dt, float* __restrict__ x, int C,
                                  int T, void* __restrict__ x_cl, int x_cl_f32, int rep, long long rep_stride) {
  pdl_sync();
  __shared__ float tile[32][33];
  __shared__ float tile2[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int t = t0 + j, c = c0 + threadIdx.x;
    float v = 0.f;
    if (t < T && c < C) {
      const long long o = ((long long)b * T + t) * ld + c;
      v = e_c[o];
      if (e_u) { const float u = e_u[o]; v = u + s * (v - u); }
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, t = t0 + threadIdx.x;
    if (c < C && t < T) {
      const long long o = ((long long)b * C + c) * T + t;
      float nv = tile[threadIdx.x][j];
      if (e_u) nv = x[o] + dt * nv;
      x[o] = nv;
      tile2[threadIdx.x][j] = nv;
    }
  }
  if (!x_cl) return;
  // also refresh the channels-last operand copies of the latent ([rep][B][T][32], the DiT stem's A
  // operand for every CFG branch) so the next evaluation needs no layout launch
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int t = t0 + j, c = c0 + threadIdx.x;
    if (t < T && c < C) {
      const long long o = ((long long)b * T + t) * 32 + c;
      for (int r = 0; r < rep; ++r) {
        if (x_cl_f32) static_cast<float*>(x_cl)[r * rep_stride + o] = tile2[j][threadIdx.x];
        else static_cast<__half*>(x_cl)[r * rep_stride + o] = vb_f2h(tile2[j][threadIdx.x]);
      }
    }
  }
}

// out[r, :] = table[idx[r], :] as fp16 (nn.Embedding, vocal2music_moe.py:336-337,388-389)
__global__ void embed_rows_kernel(const long long* __restrict__ idx, const float* __restrict__ table,
                                  int V, int D, __half* __restrict__ out, float* __restrict__ out_f32,
                                  long long rows) {
  pdl_sync();
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  long long id = idx[r];
  id = id < 0 ? 0 : (id >= V ? V - 1 : id);
  const float* src = table + id * D;
  for (int d = lane * 2; d < D; d += 64) {
    if (out_f32) {
      *reinterpret_cast<float2*>(out_f32 + r * D + d) = make_float2(src[d], src[d + 1]);
    } else {
      __half2 p = vb_f2h2(src[d], src[d + 1]);
      *reinterpret_cast<__half2*>(out + r * D + d) = p;
    }
  }
}

// out[b, t, :] = 0.5*(a[b,2t]+a[b,2t+1]) + 0.5*(b[b,2t]+b[b,2t+1])  (AvgPool1d(2) of both stem
// branches + their sum, vocal2music_moe.py:343,349,392); inputs already LeakyReLU'd.
__global__ void pool2_sum_kernel(const float* __restrict__ a, const float* __restrict__ b2, int Tin,
                                 int Tout, int D, __half* __restrict__ out_fp16,
                                 float* __restrict__ out_f32, int B) {
  pdl_sync();
  const long long n = (long long)B * Tout * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const long long bt = i / D;
    const int t = (int)(bt % Tout);
    const int b = (int)(bt / Tout);
    const long long base = ((long long)b * Tin + 2 * t) * D + d;
    float v = (a[base] + a[base + D]) * 0.5f;
    if (b2) v += (b2[base] + b2[base + D]) * 0.5f;
    if (out_fp16) out_fp16[i] = vb_f2h(v);
    if (out_f32) out_f32[i] = v;
  }
}

// Acoustic stem without a GEMM: Embedding -> Conv1d(k, pad k/2) -> LeakyReLU -> AvgPool1d(2), for the midi and
// the beats stream, summed (vocal2music_moe.py:336-350,388-393). The embedding tables have only Vm = 130 / Vb = 3
// rows, so conv(emb[idx])[t] = bias + sum_tap M[tap][idx[t + tap - pad]] with M[tap][v] = W[:, :, tap] . emb[v]
// precomputed at pack time in fp64: the 2 x 9.3 GMAC convolutions become 2 x k row gathers per time step, exact
// to fp32 rounding. The pooled sum is written as the split operand [hi | hi | lo] (hi = operand-precision value,
// lo = residue) of the 1x1 final_proj GEMM against [W_hi | W_lo | W_hi], which recovers fp32-class accuracy from
// 11-bit tensor-core operands; the acoustic gate logits derived from it then match the reference to ~1e-6.
__global__ void stem_gather_kernel(const long long* __restrict__ midi, const long long* __restrict__ beats,
                                   const float* __restrict__ tab_m, const float* __restrict__ tab_b,
                                   const float* __restrict__ bias_m, const float* __restrict__ bias_b, int Vm, int Vb,
                                   int k, int Tm, int Tpool, int D, float slope, __half* __restrict__ out16,
                                   float* __restrict__ out32, long long total) {
  pdl_sync();
  const int pad = k / 2, D4 = D / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d4 = (int)(i % D4);
    const long long bt = i / D4;
    const int tp = (int)(bt % Tpool), b = (int)(bt / Tpool);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < 2; ++s) {  // stream: midi, beats
      const long long* idx = (s ? beats : midi) + (long long)b * Tm;
      const float* tab = s ? tab_b : tab_m;
      const int V = s ? Vb : Vm;
      const float4 bs = __ldg(reinterpret_cast<const float4*>(s ? bias_b : bias_m) + d4);
#pragma unroll
      for (int u = 0; u < 2; ++u) {  // the two time steps of the pooling window
        const int t = 2 * tp + u;
        float4 v = bs;
        for (int j = 0; j < k; ++j) {
          const int tt = t + j - pad;
          if (tt < 0 || tt >= Tm) continue;
          long long id = idx[tt];
          id = id < 0 ? 0 : (id >= V ? V - 1 : id);
          const float4 m = __ldg(reinterpret_cast<const float4*>(tab + ((long long)j * V + id) * D) + d4);
          v.x += m.x; v.y += m.y; v.z += m.z; v.w += m.w;
        }
        acc.x += 0.5f * (v.x > 0.f ? v.x : v.x * slope);
        acc.y += 0.5f * (v.y > 0.f ? v.y : v.y * slope);
        acc.z += 0.5f * (v.z > 0.f ? v.z : v.z * slope);
        acc.w += 0.5f * (v.w > 0.f ? v.w : v.w * slope);
      }
    }
    const float a[4] = {acc.x, acc.y, acc.z, acc.w};
    const long long row = bt * 3 * D + 4 * d4;
    if (out16) {
      __half hi[4], lo[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) { hi[c] = vb_f2h(a[c]); lo[c] = vb_f2h(a[c] - __half2float(hi[c])); }
      *reinterpret_cast<uint2*>(out16 + row) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(out16 + row + D) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(out16 + row + 2 * D) = *reinterpret_cast<const uint2*>(lo);
    } else {
      float hi[4], lo[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {  // TF32 grid: 10-bit mantissa, round to nearest
        hi[c] = __uint_as_float((__float_as_uint(a[c]) + 0x1000u) & ~0x1FFFu);
        lo[c] = a[c] - hi[c];
      }
      *reinterpret_cast<float4*>(out32 + row) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(out32 + row + D) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(out32 + row + 2 * D) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// waveform -> [hi | lo | hi] fp16 planes with reflect padding (see include/versband.h: vb_wave_prepare)
__global__ void wave_prepare_kernel(const float* __restrict__ wav, long long n, int pad, long long plane,
                                    __half* __restrict__ out) {
  pdl_sync();
  const int b = blockIdx.y;
  const float* w = wav + (long long)b * n;
  __half* o = out + (long long)b * 3 * plane;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < plane; i += (long long)gridDim.x * blockDim.x) {
    long long j = i - pad;  // reflect (no edge repeat): -1 -> 1, n -> n-2
    if (j < 0) j = -j;
    if (j >= n) j = 2 * (n - 1) - j;
    float v = 0.f;
    if (i < n + 2 * pad && j >= 0 && j < n) v = fminf(fmaxf(w[j], -1.f), 1.f);
    const __half hi = vb_f2h(v);
    const __half lo = vb_f2h(v - __half2float(hi));
    o[i] = hi;
    o[plane + i] = lo;
    o[2 * plane + i] = hi;
  }
}

// loudness normalisation + PCM16: partial sums of squares (64 per clip), then gain and conversion
__global__ void loudness_partial_kernel(const float* __restrict__ wav, long long n, float* __restrict__ part) {
  pdl_sync();
  const int b = blockIdx.y;
  const float* w = wav + (long long)b * n;
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) acc += (double)w[i] * w[i];
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[b * 64 + blockIdx.x] = (float)sh[0];
}
__global__ void loudness_apply_kernel(const float* __restrict__ wav, long long n, const float* __restrict__ part,
                                      float target_db, short* __restrict__ out) {
  pdl_sync();
  const int b = blockIdx.y;
  __shared__ float s_gain;
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < 64; ++k) tot += part[b * 64 + k];
    const float rms = sqrtf((float)(tot / (double)n));
    const float loud = 20.f * log10f(rms);
    s_gain = powf(10.f, (target_db - loud) / 20.f);
  }
  __syncthreads();
  const float g = s_gain * 32767.f;
  const float* w = wav + (long long)b * n;
  short* o = out + (long long)b * n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = fminf(fmaxf(rintf(w[i] * g), -32768.f), 32767.f);
    o[i] = (short)v;
  }
}

// AdamW (decoupled weight decay), one launch per parameter tensor: the optimiser update of the training step
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                             float c1, float c2) {
  pdl_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float pi = p[i] * (1.f - lr * wd);
    p[i] = pi - lr * (mi / c1) / (sqrtf(vi / c2) + eps);
  }
}

// GroupNorm statistics over channels-last x[B, T, C]: per (b, group) mean and rstd
// (autoencoder1d.py:165-166: 32 groups, eps 1e-6). Two launches, deterministic: gn_partial splits T
// into `nsplit` slabs (grid = G x B x nsplit, coalesced rows, fp64 accumulation), gn_finish combines.
__global__ void gn_partial_kernel(const float* __restrict__ x, int T, int C, int G, int nsplit,
                                  double* __restrict__ part) {
  pdl_sync();
  const int g = blockIdx.x, b = blockIdx.y, sp = blockIdx.z;
  const int cpg = C / G;
  const int t0 = (int)((long long)T * sp / nsplit), t1 = (int)((long long)T * (sp + 1) / nsplit);
  const float* base = x + (long long)b * T * C + g * cpg;
  double s = 0.0, ss = 0.0;
  const int n = (t1 - t0) * cpg;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int t = i / cpg, c = i - t * cpg;
    const float v = __ldg(base + (long long)(t0 + t) * C + c);
    s += v;
    ss += (double)v * v;
  }
  __shared__ double sh[2][32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = ss; }
  __syncthreads();
  if (warp == 0) {
    s = lane < (blockDim.x >> 5) ? sh[0][lane] : 0.0;
    ss = lane < (blockDim.x >> 5) ? sh[1][lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (lane == 0) {
      double* p = part + (((long long)b * G + g) * nsplit + sp) * 2;
      p[0] = s;
      p[1] = ss;
    }
  }
}

__global__ void gn_finish_kernel(const double* __restrict__ part, int nsplit, long long n_per_group, float eps,
                                 float* __restrict__ mean, float* __restrict__ rstd, int total) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  double s = 0.0, ss = 0.0;
  for (int k = 0; k < nsplit; ++k) { s += part[((long long)i * nsplit + k) * 2]; ss += part[((long long)i * nsplit + k) * 2 + 1]; }
  const double m = s / (double)n_per_group;
  const double var = ss / (double)n_per_group - m * m;
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt((var > 0 ? var : 0) + (double)eps));
}

// y = act((x - mean) * rstd * w + b) -> fp16 (A operand of the following conv GEMM)
__global__ void gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const float* __restrict__ w,
                                const float* __restrict__ bias, int T, int C, int G, int act,
                                __half* __restrict__ out, long long total_vec) {
  pdl_sync();
  const int cpg = C / G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 4;
    const int c = (int)(e % C);
    const int b = (int)(e / ((long long)T * C));
    float4 v = *reinterpret_cast<const float4*>(x + e);
    float y[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = (c + j) / cpg;
      const float m = mean[b * G + g], r = rstd[b * G + g];
      y[j] = act_f((y[j] - m) * r * w[c + j] + bias[c + j], act, 0.f);
    }
    __half2 p0 = vb_f2h2(y[0], y[1]);
    __half2 p1 = vb_f2h2(y[2], y[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0);
    u.y = *reinterpret_cast<uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(out + e) = u;
  }
}

// GroupNorm (+ activation) -> fp16 in ONE launch: a thread-block cluster of GN_CL CTAs owns one (batch, group); each
// CTA stages its slab of rows in shared memory while accumulating fp64 partial sums, the partials are exchanged
// through distributed shared memory (every CTA sums them in the same order: deterministic), and the slab is
// normalised from shared memory — x is read from HBM/L2 once, no statistics round trip, one launch instead of three.
constexpr int GN_CL = 8;
__global__ void __launch_bounds__(256) gn_fused_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, int T, int C, int G, float eps,
                                                       int act, __half* __restrict__ out) {
  pdl_sync();
  extern __shared__ float gn_tile[];  // [rows][cpg]
  __shared__ double part[2];
  __shared__ double sh[2][8];
  __shared__ float stat[2];
  const int sp = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
  const int cpg = C / G, cq = cpg >> 2;
  const int t0 = (int)((long long)T * sp / GN_CL), t1 = (int)((long long)T * (sp + 1) / GN_CL);
  const long long base = ((long long)b * T + t0) * C + g * cpg;
  const int nvec = (t1 - t0) * cq;
  double s = 0.0, ss = 0.0;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const int r = i / cq, c4 = i - r * cq;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + base + (long long)r * C + c4 * 4));
    reinterpret_cast<float4*>(gn_tile)[i] = v;
    s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
    ss += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = 0.0; ss = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { s += sh[0][k]; ss += sh[1][k]; }
    part[0] = s; part[1] = ss;
  }
  cluster_sync_all();  // partials of all GN_CL CTAs visible cluster-wide
  if (threadIdx.x == 0) {
    s = 0.0; ss = 0.0;
    const uint32_t local = smem_u32(part);
    for (int k = 0; k < GN_CL; ++k) {
      const uint32_t remote = mapa_shared(local, (uint32_t)k);
      double a, q;
      asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(a) : "r"(remote));
      asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(q) : "r"(remote + 8));
      s += a; ss += q;
    }
    const double n = (double)T * cpg;
    const double m = s / n;
    const double var = ss / n - m * m;
    stat[0] = (float)m;
    stat[1] = (float)(1.0 / sqrt((var > 0 ? var : 0) + (double)eps));
  }
  cluster_sync_all();  // (also: no CTA exits while a peer may still read its partials)
  const float m = stat[0], r_ = stat[1];
  const float* wg = w + g * cpg;
  const float* bg = bias + g * cpg;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const int r = i / cq, c4 = i - r * cq;
    const float4 v = reinterpret_cast<const float4*>(gn_tile)[i];
    const float4 ww = __ldg(reinterpret_cast<const float4*>(wg + c4 * 4));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bg + c4 * 4));
    const __half2 p0 = vb_f2h2(act_f((v.x - m) * r_ * ww.x + bb.x, act, 0.f), act_f((v.y - m) * r_ * ww.y + bb.y, act, 0.f));
    const __half2 p1 = vb_f2h2(act_f((v.z - m) * r_ * ww.z + bb.z, act, 0.f), act_f((v.w - m) * r_ * ww.w + bb.w, act, 0.f));
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&p0);
    u.y = *reinterpret_cast<const uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(out + base + (long long)r * C + c4 * 4) = u;
  }
}

// HiFi-GAN conv_post (modules/hifigan.py:139-141): x is the already leaky-relu'd fp16
// activation [B, T, C]; wav[b, t] = tanh(bias + sum_{j,c} w[j, c] * x[b, t + j - pad, c]).
// One warp per 32 output samples: each lane owns one sample; weights in shared memory.
__global__ void conv_post_kernel(const __half* __restrict__ x, const float* __restrict__ w,
                                 float bias, int T, int C, int k, float* __restrict__ wav) {
  pdl_sync();
  extern __shared__ float sw[];
  for (int i = threadIdx.x; i < k * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int pad = k / 2;
  float acc = bias;
  for (int j = 0; j < k; ++j) {
    const int tt = t + j - pad;
    if (tt < 0 || tt >= T) continue;
    const uint4* row = reinterpret_cast<const uint4*>(x + ((long long)b * T + tt) * C);
    for (int c8 = 0; c8 < C / 8; ++c8) {
      const uint4 u = row[c8];
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __half22float2(h[q]);
        acc = fmaf(f.x, sw[j * C + c8 * 8 + 2 * q], acc);
        acc = fmaf(f.y, sw[j * C + c8 * 8 + 2 * q + 1], acc);
      }
    }
  }
  wav[(long long)b * T + t] = tanhf(acc);
}

}  // namespace vb

using namespace vb;

static inline int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" int vb_ew(const float* a, const float* b, float alpha, float beta, int act, float slope,
                     float* out_f32, void* out_fp16, int64_t rows, int D, int rep, void* stream) {
  vb::ProfScope prof_scope(vb::PK_EW, stream);
  if (!a || rows <= 0 || D <= 0 || D % 4 != 0 || rep < 1 || (!out_f32 && !out_fp16))
    return set_error(VB_ERR_INVALID, "vb_ew: bad arguments (D %% 4 == 0)");
  const long long nvec = rows * (D / 4);
  launch_k(ew_kernel, dim3(grid_for(nvec, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      a, b, alpha, beta, act, slope, out_f32, static_cast<__half*>(out_fp16), rows, D, rep);
  return check_launch("vb_ew");
}

extern "C" int vb_nct_to_ntc(const float* in, float scale, void* out_fp16, int B, int C, int T, int Cp,
                             int out_is_f32, void* stream) {
  vb::ProfScope prof_scope(vb::PK_LAYOUT, stream);
  if (!in || !out_fp16 || Cp < C) return set_error(VB_ERR_INVALID, "vb_nct_to_ntc: bad arguments");
  dim3 grid((T + 31) / 32, (Cp + 31) / 32, B), block(32, 8);
  launch_k(nct_to_ntc_kernel, dim3(grid), dim3(block), 0, static_cast<cudaStream_t>(stream), 
      in, scale, out_is_f32 ? nullptr : static_cast<__half*>(out_fp16),
      out_is_f32 ? static_cast<float*>(out_fp16) : nullptr, C, T, Cp);
  return check_launch("vb_nct_to_ntc");
}

extern "C" int vb_ntc_to_nct(const float* e_c, const float* e_u, int64_t ld, float s, float dt,
                             float* x, int B, int C, int T, void* x_cl, int x_cl_is_f32, int x_cl_rep, void* stream) {
  vb::ProfScope prof_scope(vb::PK_LAYOUT, stream);
  if (!e_c || !x || (x_cl && C > 32)) return set_error(VB_ERR_INVALID, "vb_ntc_to_nct: bad arguments");
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  launch_k(ntc_to_nct_kernel, dim3(grid), dim3(block), 0, static_cast<cudaStream_t>(stream), e_c, e_u, ld, s, dt, x, C, T,
           x_cl, x_cl_is_f32, x_cl_rep, (long long)B * T * 32);
  return check_launch("vb_ntc_to_nct");
}

extern "C" int vb_embed_rows(const int64_t* idx, const float* table, int V, int D, void* out_fp16,
                             int64_t rows, int out_is_f32, void* stream) {
  vb::ProfScope prof_scope(vb::PK_LAYOUT, stream);
  if (!idx || !table || !out_fp16 || D % 2) return set_error(VB_ERR_INVALID, "vb_embed_rows: bad arguments");
  launch_k(embed_rows_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const long long*>(idx), table, V, D, out_is_f32 ? nullptr : static_cast<__half*>(out_fp16),
      out_is_f32 ? static_cast<float*>(out_fp16) : nullptr, rows);
  return check_launch("vb_embed_rows");
}

extern "C" int vb_pool2_sum(const float* a, const float* b, int B, int Tin, int Tout, int D,
                            void* out_fp16, float* out_f32, void* stream) {
  vb::ProfScope prof_scope(vb::PK_EW, stream);
  if (!a || Tout * 2 > Tin) return set_error(VB_ERR_INVALID, "vb_pool2_sum: bad arguments");
  const long long n = (long long)B * Tout * D;
  launch_k(pool2_sum_kernel, dim3(grid_for(n, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      a, b, Tin, Tout, D, static_cast<__half*>(out_fp16), out_f32, B);
  return check_launch("vb_pool2_sum");
}

extern "C" int vb_gn_stats(const float* x, int B, int T, int C, int G, float eps, float* mean,
                           float* rstd, void* scratch, void* stream) {
  vb::ProfScope prof_scope(vb::PK_GN, stream);
  if (!x || !mean || !rstd || !scratch || C % G) return set_error(VB_ERR_INVALID, "vb_gn_stats: bad arguments");
  const int nsplit = VB_GN_SPLIT;
  launch_k(gn_partial_kernel, dim3(G, B, nsplit), dim3(256), 0, static_cast<cudaStream_t>(stream), x, T, C, G, nsplit,
           static_cast<double*>(scratch));
  int rc = check_launch("vb_gn_stats");
  if (rc) return rc;
  launch_k(gn_finish_kernel, dim3((B * G + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream),
           static_cast<const double*>(scratch), nsplit, (long long)T * (C / G), eps, mean, rstd, B * G);
  return check_launch("vb_gn_stats");
}

extern "C" int vb_gn_apply(const float* x, const float* mean, const float* rstd, const float* w,
                           const float* bias, int B, int T, int C, int G, int act, void* out_fp16,
                           void* stream) {
  vb::ProfScope prof_scope(vb::PK_GN, stream);
  if (!x || !mean || !rstd || !w || !bias || !out_fp16 || C % 4 || C % G)
    return set_error(VB_ERR_INVALID, "vb_gn_apply: bad arguments");
  const long long nvec = (long long)B * T * C / 4;
  launch_k(gn_apply_kernel, dim3(grid_for(nvec, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      x, mean, rstd, w, bias, T, C, G, act, static_cast<__half*>(out_fp16), nvec);
  return check_launch("vb_gn_apply");
}

// GroupNorm + activation -> fp16 in one cluster launch (falls back to stats + apply for shapes whose slab does not
// fit in shared memory: scratch / mean / rstd as for vb_gn_stats, may be NULL when the fused kernel applies)
extern "C" int vb_gn_fused_fits(int T, int C, int G) {
  if (C % G || (C / G) % 4 || T < vb::GN_CL) return 0;
  const long long rows = ((long long)T + vb::GN_CL - 1) / vb::GN_CL + 1;
  return rows * (C / G) * 4 <= 200 * 1024;
}

extern "C" int vb_gn_fused(const float* x, const float* w, const float* bias, int B, int T, int C, int G, float eps, int act,
                           void* out_fp16, void* stream) {
  vb::ProfScope prof_scope(vb::PK_GN, stream);
  if (!x || !w || !bias || !out_fp16 || !vb_gn_fused_fits(T, C, G)) return set_error(VB_ERR_INVALID, "vb_gn_fused: bad arguments / shape does not fit");
  const long long rows = ((long long)T + vb::GN_CL - 1) / vb::GN_CL + 1;
  const size_t smem = (size_t)rows * (C / G) * 4;
  static vb::PerDevice<size_t> attr_smem;
  if (smem > 48 * 1024 && attr_smem.get() < smem) {
    cudaError_t e = cudaFuncSetAttribute(vb::gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "vb_gn_fused: %s", cudaGetErrorString(e));
    attr_smem.get() = 200 * 1024;
  }
  launch_k_cluster(vb::gn_fused_kernel, vb::GN_CL, dim3(vb::GN_CL, G, B), dim3(256), smem, static_cast<cudaStream_t>(stream), x, w, bias,
                   T, C, G, eps, act, static_cast<__half*>(out_fp16));
  return check_launch("vb_gn_fused");
}

extern "C" int vb_conv_post(const void* x_fp16, const float* w, float bias, int B, int T, int C, int k,
                            float* wav, void* stream) {
  vb::ProfScope prof_scope(vb::PK_CONV_POST, stream);
  if (!x_fp16 || !w || !wav || C % 8) return set_error(VB_ERR_INVALID, "vb_conv_post: bad arguments");
  dim3 grid((T + 127) / 128, B);
  launch_k(conv_post_kernel, dim3(grid), dim3(128), k * C * sizeof(float), static_cast<cudaStream_t>(stream), 
      static_cast<const __half*>(x_fp16), w, bias, T, C, k, wav);
  return check_launch("vb_conv_post");
}

// Debug aid: write the GPU global timer to *slot (in-graph timeline of an op sequence).
namespace vb {
__global__ void timestamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}
}  // namespace vb

extern "C" int vb_timestamp(uint64_t* slot, void* stream) {
  vb::timestamp_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned long long*>(slot));
  return vb::check_launch("vb_timestamp");
}

extern "C" int vb_stem_gather(const int64_t* midi, const int64_t* beats, const float* tab_m, const float* tab_b,
                              const float* bias_m, const float* bias_b, int Vm, int Vb, int k, int B, int Tm, int Tpool,
                              int D, float slope, void* out, int out_is_f32, void* stream) {
  vb::ProfScope prof_scope(vb::PK_EW, stream);
  if (!midi || !beats || !tab_m || !tab_b || !bias_m || !bias_b || !out || D % 4 || k < 1 || Tpool * 2 > Tm)
    return set_error(VB_ERR_INVALID, "vb_stem_gather: bad arguments");
  const long long total = (long long)B * Tpool * (D / 4);
  launch_k(stem_gather_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
           reinterpret_cast<const long long*>(midi), reinterpret_cast<const long long*>(beats), tab_m, tab_b, bias_m, bias_b,
           Vm, Vb, k, Tm, Tpool, D, slope, out_is_f32 ? nullptr : static_cast<__half*>(out),
           out_is_f32 ? static_cast<float*>(out) : nullptr, total);
  return check_launch("vb_stem_gather");
}

extern "C" int vb_wave_prepare(const float* wav, int B, int64_t n, int pad, int64_t plane, void* out_fp16, void* stream) {
  vb::ProfScope prof_scope(vb::PK_EW, stream);
  if (!wav || !out_fp16 || B <= 0 || n <= pad || pad < 0 || plane < n + 2 * pad || plane % 8)
    return set_error(VB_ERR_INVALID, "vb_wave_prepare: bad arguments (plane >= n + 2*pad, multiple of 8; n > pad)");
  launch_k(wave_prepare_kernel, dim3(grid_for(plane, 256) > 1024 ? 1024 : grid_for(plane, 256), B), dim3(256), 0,
           static_cast<cudaStream_t>(stream), wav, (long long)n, pad, (long long)plane, static_cast<__half*>(out_fp16));
  return check_launch("vb_wave_prepare");
}

extern "C" int vb_loudness_pcm16(const float* wav, int B, int64_t n, float target_db, float* scratch, int16_t* out,
                                 void* stream) {
  vb::ProfScope prof_scope(vb::PK_EW, stream);
  if (!wav || !scratch || !out || B <= 0 || n <= 0) return set_error(VB_ERR_INVALID, "vb_loudness_pcm16: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  launch_k(loudness_partial_kernel, dim3(64, B), dim3(256), 0, s, wav, (long long)n, scratch);
  int rc = check_launch("vb_loudness_pcm16");
  if (rc) return rc;
  launch_k(loudness_apply_kernel, dim3(256, B), dim3(256), 0, s, wav, (long long)n, static_cast<const float*>(scratch),
           target_db, reinterpret_cast<short*>(out));
  return check_launch("vb_loudness_pcm16");
}

extern "C" int vb_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                        float wd, float c1, float c2, void* stream) {
  vb::ProfScope prof_scope(vb::PK_EW, stream);
  if (!p || !g || !m || !v || n <= 0) return set_error(VB_ERR_INVALID, "vb_adamw: bad arguments");
  int grid = grid_for(n, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  launch_k(adamw_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), p, g, m, v, (long long)n, lr, b1, b2,
           eps, wd, c1, c2);
  return check_launch("vb_adamw");
}

// out[g][c][r] = in[g][r][c] for fp32 matrices (rows x cols, leading dimensions ld_in / ld_out >= rows; the
// ld_out - rows padding columns are zero-filled): the operand transposes of the training backward pass
// (dW = dY^T X needs both factors K-major over the tokens). 32 x 32 tiles through padded shared memory, both sides coalesced.
namespace vb {
__global__ void __launch_bounds__(256) transpose_f32_kernel(const float* __restrict__ in, long long rows, long long cols,
                                                            long long ld_in, long long in_g, float* __restrict__ out,
                                                            long long ld_out, long long out_g) {
  pdl_sync();
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
  const float* src = in + (long long)blockIdx.z * in_g;
  float* dst = out + (long long)blockIdx.z * out_g;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const long long r = r0 + ty + j, c = c0 + tx;
    tile[ty + j][tx] = (r < rows && c < cols) ? __ldg(src + r * ld_in + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const long long c = c0 + ty + j, r = r0 + tx;   // out row = input column
    if (c < cols && r < ld_out) dst[c * ld_out + r] = tile[tx][ty + j];
  }
}
}  // namespace vb

extern "C" int vb_transpose_f32(const float* in, int64_t rows, int64_t cols, int64_t ld_in, int64_t in_batch_stride, float* out,
                                int64_t ld_out, int64_t out_batch_stride, int batch, void* stream) {
  vb::ProfScope prof_scope(vb::PK_LAYOUT, stream);
  if (!in || !out || rows <= 0 || cols <= 0 || ld_in < cols || ld_out < rows || batch <= 0 || batch > 65535)
    return set_error(VB_ERR_INVALID, "vb_transpose_f32: bad arguments");
  const long long gy = (ld_out + 31) / 32, gx = (cols + 31) / 32;
  if (gy > 65535) return set_error(VB_ERR_INVALID, "vb_transpose_f32: too many rows");
  launch_k(vb::transpose_f32_kernel, dim3((unsigned)gx, (unsigned)gy, (unsigned)batch), dim3(256), 0, static_cast<cudaStream_t>(stream), in,
           (long long)rows, (long long)cols, (long long)ld_in, (long long)in_batch_stride, out, (long long)ld_out,
           (long long)out_batch_stride);
  return check_launch("vb_transpose_f32");
}
