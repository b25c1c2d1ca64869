// Band-MoE routing kernels (vocal2music_moe.py:117-185):
//   vb_linear_small : fp32 row x [E<=8] gating logits (exact fp32 FMA dot products)
//   vb_router       : Gumbel-softmax group gate (soft) + hard top-1 per routed group
//   vb_moe_sort     : counting sort of (token, group) pairs by expert into 128-row tiles
//   vb_gather_rows  : x_sorted[slot] = x[row_map[slot] % n_tokens]
#include <math_constants.h>
#include "vb_internal.h"

namespace vb {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[r, e] = bias[e] + sum_d x[r, d] * W[e, d]      (one warp per (row, output) pair)
__global__ void linear_small_kernel(const float* __restrict__ x, long long ldx,
                                    const float* __restrict__ W, const float* __restrict__ bias,
                                    float* __restrict__ out, int rows, int D, int E) {
  pdl_sync();
  const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pair >= rows * E) return;
  const int row = pair / E, e = pair - row * E;
  const float* xr = x + (long long)row * ldx;
  const float* wr = W + (long long)e * D;
  float acc = 0.f;
#pragma unroll 4
  for (int d = lane; d < D; d += 32) acc = fmaf(__ldg(xr + d), __ldg(wr + d), acc);
  acc = wsum(acc);
  if (lane == 0) out[(long long)row * E + e] = acc + (bias ? bias[e] : 0.f);
}

// One warp per token n (= b*T + t).
//  group gate : p = softmax(hl_logits[b] - log(E0[n]))           (tau = 1, always soft)  :132-138
//  expert gate: y = softmax((logits - log(E)) / tau); idx = argmax (first max); hard weight
//               w = (1 - y[idx]) + y[idx]  (straight-through value)                       :81-93,150-151
//  scale_c[n] = w_c * p[0], scale_a[n] = w_a * p[1]                                       :157-165
template <int E>
__global__ void router_kernel(const float* __restrict__ cap_feat, const __half* __restrict__ cap_feat_h,
                              const float* __restrict__ Wc,
                              const float* __restrict__ bc, const float* __restrict__ ac_logits,
                              const float* __restrict__ hl_logits, int hl_ld, int ac_ld, const float* __restrict__ e0,
                              const float* __restrict__ e1, const float* __restrict__ e2, float tau,
                              int N, int T, int D, int* __restrict__ idx_c, int* __restrict__ idx_a,
                              float* __restrict__ scale_c, float* __restrict__ scale_a,
                              float* __restrict__ cap_logits_out, int* __restrict__ bucket_count,
                              float* __restrict__ usage, int* __restrict__ slot_c,
                              int* __restrict__ slot_a, const float* __restrict__ cap_partial, int n_heads) {
  pdl_sync();
  __shared__ int s_cnt[2 * E], s_base[2 * E];
  __shared__ float s_use[2 * E];
  if (threadIdx.x < 2 * E) { s_cnt[threadIdx.x] = 0; s_use[threadIdx.x] = 0.f; }
  __syncthreads();
  const int n_raw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool tok_ok = n_raw < N;
  const int n = tok_ok ? n_raw : N - 1;  // surplus warps recompute the last token and store nothing
  const int lane = threadIdx.x & 31;
  float acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = 0.f;
  if (cap_partial) {
    // per-head contributions from the VB_EPI_ROUTE GEMM, summed in head order (deterministic)
    static_assert(E == 4, "cap_partial layout is [N, heads, 4]");
    const float4* pp = reinterpret_cast<const float4*>(cap_partial) + (long long)n * n_heads;
    for (int h = 0; h < n_heads; ++h) {
      const float4 v = __ldg(pp + h);
      acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
    }
  } else if ((D & 127) == 0) {  // 16-byte loads: lane owns 4 consecutive features per 128-feature slab
    for (int d = lane * 4; d < D; d += 128) {
      float4 xv;
      if (cap_feat_h) {
        const uint2 u = *reinterpret_cast<const uint2*>(cap_feat_h + (long long)n * D + d);
        const __half2 p0 = *reinterpret_cast<const __half2*>(&u.x);
        const __half2 p1 = *reinterpret_cast<const __half2*>(&u.y);
        xv = make_float4(__low2float(p0), __high2float(p0), __low2float(p1), __high2float(p1));
      } else {
        xv = *reinterpret_cast<const float4*>(cap_feat + (long long)n * D + d);
      }
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(Wc + (long long)e * D + d));
        acc[e] = fmaf(xv.x, wv.x, acc[e]);
        acc[e] = fmaf(xv.y, wv.y, acc[e]);
        acc[e] = fmaf(xv.z, wv.z, acc[e]);
        acc[e] = fmaf(xv.w, wv.w, acc[e]);
      }
    }
  } else if (cap_feat_h) {
    const __half* xr = cap_feat_h + (long long)n * D;
#pragma unroll 4
    for (int d = lane; d < D; d += 32) {
      const float xv = __half2float(xr[d]);
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] = fmaf(xv, __ldg(Wc + (long long)e * D + d), acc[e]);
    }
  } else {
    const float* xr = cap_feat + (long long)n * D;
#pragma unroll 4
    for (int d = lane; d < D; d += 32) {
      const float xv = xr[d];
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] = fmaf(xv, __ldg(Wc + (long long)e * D + d), acc[e]);
    }
  }
  if (!cap_partial) {
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = wsum(acc[e]);
  }
  // The three Gumbel-softmaxes of this token run in parallel on lanes 0..E-1 (caption gate),
  // E..2E-1 (acoustic gate) and 2E, 2E+1 (group gate); every lane of a gate gathers the gate's
  // values with shuffles and reduces them in index order, i.e. the same arithmetic as a serial loop.
  const int b = n / T;
  const int gate = lane < E ? 0 : (lane < 2 * E ? 1 : 2);
  const int base = gate == 0 ? 0 : (gate == 1 ? E : 2 * E);
  const int width = gate == 2 ? 2 : E;
  const int j = lane - base;
  const bool active = lane < 2 * E + 2;
  float logit = 0.f, ex = 1.f, tau_g = tau;
  if (gate == 0) {
    float lcv = acc[0];
#pragma unroll
    for (int e = 1; e < E; ++e) lcv = (j == e) ? acc[e] : lcv;
    logit = lcv + bc[j < E ? j : 0];
    if (cap_logits_out && lane < E && tok_ok) cap_logits_out[(long long)n * E + j] = logit;
    ex = e1[(long long)n * E + (j < E ? j : 0)];
  } else if (gate == 1) {
    logit = ac_logits[(long long)n * ac_ld + j];
    ex = e2[(long long)n * E + j];
  } else if (active) {
    logit = hl_logits[b * hl_ld + j];
    ex = e0[(long long)n * 2 + j];
    tau_g = 1.0f;
  }
  const float z = (logit + (-logf(ex))) / tau_g;
  float zz[E];
#pragma unroll
  for (int e = 0; e < E; ++e) zz[e] = __shfl_sync(0xffffffffu, z, base + (e < width ? e : 0));
  float m = zz[0];
#pragma unroll
  for (int e = 1; e < E; ++e) if (e < width) m = fmaxf(m, zz[e]);
  float ssum = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) if (e < width) { zz[e] = expf(zz[e] - m); ssum += zz[e]; }
  // (lanes of one gate now hold identical zz[], m, ssum)
  int best = 0;
  float pb = zz[0] / ssum;
#pragma unroll
  for (int e = 1; e < E; ++e) {
    if (e < width) {
      const float pe = zz[e] / ssum;
      if (pe > pb) { pb = pe; best = e; }
    }
  }
  const float w_hard = (1.f - pb) + pb;
  // group masks: soft probabilities of the group gate (lane 2E holds p[0] in zz[0]/ssum, etc.)
  const float mask_c = __shfl_sync(0xffffffffu, zz[0] / ssum, 2 * E);
  const float mask_a = __shfl_sync(0xffffffffu, zz[1] / ssum, 2 * E);
  const int ia = __shfl_sync(0xffffffffu, best, E);
  const float wa = __shfl_sync(0xffffffffu, w_hard, E);
  const int ic = best;
  const float wc = w_hard;
  const bool writer = lane == 0 && tok_ok;
  if (writer) {
    idx_c[n] = ic;
    idx_a[n] = ia;
    scale_c[n] = wc * mask_c;
    scale_a[n] = wa * mask_a;
  }
  if (bucket_count) {
    // unordered slot claim inside the (group, expert) bucket: ranks inside the CTA from shared-memory
    // atomics, ONE global atomic per bucket and CTA (16 addresses chip-wide would serialise otherwise)
    int rc = 0, ra = 0;
    if (writer) {
      rc = atomicAdd(&s_cnt[ic], 1);
      ra = atomicAdd(&s_cnt[E + ia], 1);
      atomicAdd(&s_use[ic], wc * mask_c);
      atomicAdd(&s_use[E + ia], wa * mask_a);
    }
    __syncthreads();
    if (threadIdx.x < 2 * E && s_cnt[threadIdx.x] > 0) {
      s_base[threadIdx.x] = atomicAdd(&bucket_count[threadIdx.x], s_cnt[threadIdx.x]);
      atomicAdd(&usage[threadIdx.x], s_use[threadIdx.x]);
    }
    __syncthreads();
    if (writer) {
      slot_c[n] = s_base[ic] + rc;
      slot_a[n] = s_base[E + ia] + ra;
    }
  }
}

// One small CTA: padded bucket starts, tile table, tile count, load-balancing loss; marks padding
// rows invalid; re-zeroes the counters for the next evaluation.
__global__ void moe_finalize_kernel(int* __restrict__ bucket_count, float* __restrict__ usage, int nb, int E,
                                    int* __restrict__ bucket_start, int* __restrict__ row_map,
                                    float* __restrict__ row_scale, int* __restrict__ tile_group,
                                    int* __restrict__ m_tiles_used, int max_rows, float* __restrict__ lb_out) {
  pdl_sync();
  __shared__ int s_start[17], s_cnt[16];
  if (threadIdx.x == 0) {
    int acc = 0;
    float tot = 0.f;
    for (int k = 0; k < nb; ++k) {
      const int c = bucket_count[k];
      s_cnt[k] = c;
      s_start[k] = acc;
      bucket_start[k] = acc;
      const int padded = (c + 127) / 128 * 128;
      for (int t = acc / 128; t < (acc + padded) / 128; ++t) tile_group[t] = k;
      acc += padded;
      tot += usage[k];
    }
    s_start[nb] = acc;
    *m_tiles_used = acc / 128;
    if (lb_out) {
      float loss = 0.f;
      const float total = tot * (float)E;
      for (int k = 0; k < nb; ++k) {
        const float u = usage[k] / (total + 1e-10f);
        loss += u * logf(u + 1e-10f);
      }
      *lb_out = loss / (float)nb;
    }
  }
  __syncthreads();
  for (int k = 0; k < nb; ++k)
    for (int s = s_start[k] + s_cnt[k] + threadIdx.x; s < s_start[k + 1]; s += blockDim.x) { row_map[s] = -1; row_scale[s] = 0.f; }
  for (int s = s_start[nb] + threadIdx.x; s < max_rows; s += blockDim.x) { row_map[s] = -1; row_scale[s] = 0.f; }
  __syncthreads();
  if (threadIdx.x < nb) { bucket_count[threadIdx.x] = 0; usage[threadIdx.x] = 0.f; }
}

// One warp per (token, group) entry: place it in its bucket slot and copy its activation row.
__global__ void moe_scatter_kernel(const int* __restrict__ idx_c, const int* __restrict__ idx_a,
                                   const int* __restrict__ slot_c, const int* __restrict__ slot_a,
                                   const float* __restrict__ scale_c, const float* __restrict__ scale_a,
                                   const int* __restrict__ bucket_start, int N, int E, int row_bytes,
                                   const uint8_t* __restrict__ x, uint8_t* __restrict__ xs,
                                   int* __restrict__ row_map, float* __restrict__ row_scale) {
  pdl_sync();
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= 2 * N) return;
  const int g = i / N, n = i - g * N;
  const int bucket = g * E + (g ? idx_a[n] : idx_c[n]);
  const int slot = bucket_start[bucket] + (g ? slot_a[n] : slot_c[n]);
  if (lane == 0) { row_map[slot] = i; row_scale[slot] = g ? scale_a[n] : scale_c[n]; }
  const uint4* src = reinterpret_cast<const uint4*>(x + (long long)n * row_bytes);
  uint4* dst = reinterpret_cast<uint4*>(xs + (long long)slot * row_bytes);
  for (int k = lane; k < row_bytes / 16; k += 32) dst[k] = src[k];
}

// Single CTA. Entries i in [0, 2N): group g = i / N (0 caption, 1 acoustic), token = i % N,
// bucket = g*E + idx. Buckets are laid out back to back, each padded to a multiple of 128 rows.
// Warp w (< 2E) owns bucket w: ordered ballot scan -> deterministic slots.
// Also reduces the reference's load-balancing loss (vocal2music_moe.py:95-115) for this block.
__global__ void moe_sort_kernel(const int* __restrict__ idx_c, const int* __restrict__ idx_a,
                                const float* __restrict__ scale_c, const float* __restrict__ scale_a,
                                int N, int E, int* __restrict__ row_map, float* __restrict__ row_scale,
                                int* __restrict__ tile_group, int* __restrict__ m_tiles_used,
                                int max_rows, float* __restrict__ lb_out) {
  pdl_sync();
  __shared__ int s_count[16];
  __shared__ int s_start[17];
  __shared__ float s_usage[16];
  __shared__ float s_masksum;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = 2 * E;
  if (threadIdx.x < 16) { s_count[threadIdx.x] = 0; s_usage[threadIdx.x] = 0.f; }
  if (threadIdx.x == 0) s_masksum = 0.f;
  __syncthreads();
  // pass 1: counts + usage sums
  if (warp < nb) {
    const int g = warp / E, e = warp - g * E;
    const int* idx = g ? idx_a : idx_c;
    const float* sc = g ? scale_a : scale_c;
    int cnt = 0;
    float us = 0.f;
    for (int n = lane; n < N; n += 32)
      if (idx[n] == e) { cnt++; us += sc[n]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      us += __shfl_xor_sync(0xffffffffu, us, o);
    }
    if (lane == 0) { s_count[warp] = cnt; s_usage[warp] = us; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int k = 0; k < nb; ++k) {
      s_start[k] = acc;
      const int padded = (s_count[k] + 127) / 128 * 128;
      for (int t = acc / 128; t < (acc + padded) / 128; ++t) tile_group[t] = k;
      acc += padded;
    }
    s_start[nb] = acc;
    *m_tiles_used = acc / 128;
  }
  __syncthreads();
  // pass 2: ordered slot assignment
  if (warp < nb) {
    const int g = warp / E, e = warp - g * E;
    const int* idx = g ? idx_a : idx_c;
    const float* sc = g ? scale_a : scale_c;
    int base = s_start[warp];
    for (int n0 = 0; n0 < N; n0 += 32) {
      const int n = n0 + lane;
      const bool mine = (n < N) && (idx[n] == e);
      const unsigned bal = __ballot_sync(0xffffffffu, mine);
      if (mine) {
        const int slot = base + __popc(bal & ((1u << lane) - 1));
        row_map[slot] = g * N + n;
        row_scale[slot] = sc[n];
      }
      base += __popc(bal);
    }
    for (int s = base + lane; s < s_start[warp + 1]; s += 32) { row_map[s] = -1; row_scale[s] = 0.f; }
  }
  // rows past the used tiles: mark invalid
  for (int s = s_start[nb] + threadIdx.x; s < max_rows; s += blockDim.x) { row_map[s] = -1; row_scale[s] = 0.f; }
  // load-balancing loss: usage_k / (E * (sum mask_c + sum mask_a)); in eval mode the masked hard
  // probabilities sum, per token and group, to scale_g[n]; sum of masks = sum over tokens of
  // (mask_c + mask_a) = N (softmax over two groups) up to rounding -> recomputed from usage.
  if (lb_out && threadIdx.x == 0) {
    float total = 0.f;  // sum of all expanded-mask entries = E * sum_n (mask_c + mask_a)
    // hard weights are 1 (+-1ulp), so sum_n mask_g = sum_k usage_{g,k}
    for (int k = 0; k < nb; ++k) total += s_usage[k];
    total *= (float)E;
    float loss = 0.f;
    for (int k = 0; k < nb; ++k) {
      const float u = s_usage[k] / (total + 1e-10f);
      loss += u * logf(u + 1e-10f);
    }
    *lb_out = loss / (float)nb;
  }
}

// x_sorted[slot, :] = x[row_map[slot] % n_tokens, :] (fp16 rows, 16-byte vectors); slot invalid -> zeros
__global__ void gather_rows_kernel(const __half* __restrict__ x, const int* __restrict__ row_map,
                                   const int* __restrict__ m_tiles_used, int n_tokens, int D,
                                   __half* __restrict__ out, int max_rows) {
  pdl_sync();
  const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (slot >= max_rows || slot >= *m_tiles_used * 128) return;
  const int r = row_map[slot];
  uint4* dst = reinterpret_cast<uint4*>(out + (long long)slot * D);
  if (r < 0) {
    for (int i = lane; i < D / 8; i += 32) dst[i] = make_uint4(0, 0, 0, 0);
    return;
  }
  const uint4* src = reinterpret_cast<const uint4*>(x + (long long)(r % n_tokens) * D);
  for (int i = lane; i < D / 8; i += 32) dst[i] = src[i];
}

// Dispatch in one launch. Every CTA rebuilds the (<= 16) padded bucket starts from the router's
// counters; one warp per (token, group) entry places the entry in its bucket slot and copies its
// activation row; CTA 0 additionally writes the tile table, tile count, load-balancing loss,
// marks padding rows invalid and zeroes the NEXT block's counters (this block's counters were
// zeroed by the previous block's dispatch, so no separate memset/finalize launch is needed).
__global__ void moe_dispatch_kernel(const int* __restrict__ idx_c, const int* __restrict__ idx_a,
                                    const int* __restrict__ slot_c, const int* __restrict__ slot_a,
                                    const float* __restrict__ scale_c, const float* __restrict__ scale_a,
                                    const int* __restrict__ bucket_count, const float* __restrict__ usage,
                                    int* __restrict__ next_count, float* __restrict__ next_usage, int N, int E,
                                    int row_bytes, const uint8_t* __restrict__ x, uint8_t* __restrict__ xs,
                                    int* __restrict__ row_map, float* __restrict__ row_scale,
                                    int* __restrict__ tile_group, int* __restrict__ m_tiles_used, int max_rows,
                                    float* __restrict__ lb_out) {
  pdl_sync();
  __shared__ int s_start[17], s_cnt[16];
  const int nb = 2 * E;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int k = 0; k < nb; ++k) {
      const int c = bucket_count[k];
      s_cnt[k] = c;
      s_start[k] = acc;
      acc += (c + 127) / 128 * 128;
    }
    s_start[nb] = acc;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int k = 0; k < nb; ++k) {
        for (int t = s_start[k] / 128; t < s_start[k + 1] / 128; ++t) tile_group[t] = k;
        tot += usage[k];
      }
      *m_tiles_used = s_start[nb] / 128;
      if (lb_out) {
        float loss = 0.f;
        const float total = tot * (float)E;
        for (int k = 0; k < nb; ++k) {
          const float u = usage[k] / (total + 1e-10f);
          loss += u * logf(u + 1e-10f);
        }
        *lb_out = loss / (float)nb;
      }
    }
    for (int k = 0; k < nb; ++k)
      for (int sl = s_start[k] + s_cnt[k] + threadIdx.x; sl < s_start[k + 1]; sl += blockDim.x) { row_map[sl] = -1; row_scale[sl] = 0.f; }
    for (int sl = s_start[nb] + threadIdx.x; sl < max_rows; sl += blockDim.x) { row_map[sl] = -1; row_scale[sl] = 0.f; }
    if (threadIdx.x < nb && next_count) { next_count[threadIdx.x] = 0; next_usage[threadIdx.x] = 0.f; }
  }
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= 2 * N) return;
  const int g = i / N, n = i - g * N;
  const int bucket = g * E + (g ? idx_a[n] : idx_c[n]);
  const int slot = s_start[bucket] + (g ? slot_a[n] : slot_c[n]);
  if (lane == 0) { row_map[slot] = i; row_scale[slot] = g ? scale_a[n] : scale_c[n]; }
  const uint4* src = reinterpret_cast<const uint4*>(x + (long long)n * row_bytes);
  uint4* dst = reinterpret_cast<uint4*>(xs + (long long)slot * row_bytes);
  for (int k = lane; k < row_bytes / 16; k += 32) dst[k] = src[k];
}


// ------------------------------------------------------------------ router + dispatch in ONE launch
// vb_route_dispatch: the Gumbel-softmax router (above) and the expert dispatch fused. Buckets have a FIXED
// capacity of `cap_rows` rows each (cap_rows = N rounded up to 128: one group's tokens could all pick the same
// expert), bucket k occupies rows [k*cap_rows, (k+1)*cap_rows) of xs / row_map / row_scale, so a token's row is
// known as soon as its slot inside the bucket is claimed — no prefix over the bucket counts, hence no second
// launch: one warp per token routes it, claims its two slots (CTA-aggregated atomics), and copies its operand
// row into both. The LAST CTA to finish (ticket counter) builds the compact tile table of the grouped GEMMs
// from the final counts — (group, first row) of every 128-row tile that holds at least one token — marks the
// padding rows of partial tiles invalid, reduces the load-balancing loss and zeroes the next block's counters.
template <int E>
__global__ void route_dispatch_kernel(const float* __restrict__ cap_partial, int n_heads, const float* __restrict__ bc,
                                      const float* __restrict__ ac_logits, int ac_ld,
                                      const float* __restrict__ hl_logits, int hl_ld, const float* __restrict__ e0,
                                      const float* __restrict__ e1, const float* __restrict__ e2, float tau, int N,
                                      int T, int row_bytes, const uint8_t* __restrict__ x, uint8_t* __restrict__ xs,
                                      int cap_rows, int* __restrict__ idx_c, int* __restrict__ idx_a,
                                      float* __restrict__ cap_logits_out, int* __restrict__ bucket_count,
                                      float* __restrict__ usage, int* __restrict__ done_ticket,
                                      int* __restrict__ next_count, float* __restrict__ next_usage,
                                      int* __restrict__ row_map, float* __restrict__ row_scale,
                                      int* __restrict__ tile_group, int* __restrict__ tile_row0,
                                      int* __restrict__ m_tiles_used, float* __restrict__ lb_out) {
  pdl_sync();
  static_assert(E == 4, "cap_partial layout is [N, heads, 4]");
  __shared__ int s_cnt[2 * E], s_base[2 * E];
  __shared__ float s_use[2 * E];
  __shared__ int s_last;
  if (threadIdx.x < 2 * E) { s_cnt[threadIdx.x] = 0; s_use[threadIdx.x] = 0.f; }
  __syncthreads();
  const int n_raw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool tok_ok = n_raw < N;
  const int n = tok_ok ? n_raw : N - 1;  // surplus warps recompute the last token and store nothing
  const int lane = threadIdx.x & 31;
  float acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = 0.f;
  {  // per-head contributions from the VB_EPI_ROUTE GEMM, summed in head order (deterministic)
    const float4* pp = reinterpret_cast<const float4*>(cap_partial) + (long long)n * n_heads;
    for (int h = 0; h < n_heads; ++h) {
      const float4 v = __ldg(pp + h);
      acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
    }
  }
  // the three Gumbel-softmaxes: identical arithmetic to router_kernel (lanes 0..E-1 caption gate,
  // E..2E-1 acoustic gate, 2E..2E+1 group gate)
  const int b = n / T;
  const int gate = lane < E ? 0 : (lane < 2 * E ? 1 : 2);
  const int base = gate == 0 ? 0 : (gate == 1 ? E : 2 * E);
  const int width = gate == 2 ? 2 : E;
  const int j = lane - base;
  const bool active = lane < 2 * E + 2;
  float logit = 0.f, ex = 1.f, tau_g = tau;
  if (gate == 0) {
    float lcv = acc[0];
#pragma unroll
    for (int e = 1; e < E; ++e) lcv = (j == e) ? acc[e] : lcv;
    logit = lcv + bc[j < E ? j : 0];
    if (cap_logits_out && lane < E && tok_ok) cap_logits_out[(long long)n * E + j] = logit;
    ex = e1[(long long)n * E + (j < E ? j : 0)];
  } else if (gate == 1) {
    logit = ac_logits[(long long)n * ac_ld + j];
    ex = e2[(long long)n * E + j];
  } else if (active) {
    logit = hl_logits[b * hl_ld + j];
    ex = e0[(long long)n * 2 + j];
    tau_g = 1.0f;
  }
  const float z = (logit + (-logf(ex))) / tau_g;
  float zz[E];
#pragma unroll
  for (int e = 0; e < E; ++e) zz[e] = __shfl_sync(0xffffffffu, z, base + (e < width ? e : 0));
  float m = zz[0];
#pragma unroll
  for (int e = 1; e < E; ++e) if (e < width) m = fmaxf(m, zz[e]);
  float ssum = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) if (e < width) { zz[e] = expf(zz[e] - m); ssum += zz[e]; }
  int best = 0;
  float pb = zz[0] / ssum;
#pragma unroll
  for (int e = 1; e < E; ++e) {
    if (e < width) {
      const float pe = zz[e] / ssum;
      if (pe > pb) { pb = pe; best = e; }
    }
  }
  const float w_hard = (1.f - pb) + pb;
  const float mask_c = __shfl_sync(0xffffffffu, zz[0] / ssum, 2 * E);
  const float mask_a = __shfl_sync(0xffffffffu, zz[1] / ssum, 2 * E);
  const int ic = __shfl_sync(0xffffffffu, best, 0);
  const float wc = __shfl_sync(0xffffffffu, w_hard, 0);
  const int ia = __shfl_sync(0xffffffffu, best, E);
  const float wa = __shfl_sync(0xffffffffu, w_hard, E);
  const float sc_c = wc * mask_c, sc_a = wa * mask_a;
  const bool writer = lane == 0 && tok_ok;
  int rc = 0, ra = 0;
  if (writer) {
    idx_c[n] = ic;
    idx_a[n] = ia;
    rc = atomicAdd(&s_cnt[ic], 1);
    ra = atomicAdd(&s_cnt[E + ia], 1);
    atomicAdd(&s_use[ic], sc_c);
    atomicAdd(&s_use[E + ia], sc_a);
  }
  __syncthreads();
  if (threadIdx.x < 2 * E && s_cnt[threadIdx.x] > 0) {
    s_base[threadIdx.x] = atomicAdd(&bucket_count[threadIdx.x], s_cnt[threadIdx.x]);
    atomicAdd(&usage[threadIdx.x], s_use[threadIdx.x]);
  }
  __syncthreads();
  if (tok_ok) {
    // rows of this token in the two groups' buckets (all lanes need them for the copy)
    const int row_c = ic * cap_rows + s_base[ic] + __shfl_sync(0xffffffffu, rc, 0);
    const int row_a = (E + ia) * cap_rows + s_base[E + ia] + __shfl_sync(0xffffffffu, ra, 0);
    if (lane == 0) {
      row_map[row_c] = n;        row_scale[row_c] = sc_c;
      row_map[row_a] = N + n;    row_scale[row_a] = sc_a;
    }
    const uint4* src = reinterpret_cast<const uint4*>(x + (long long)n * row_bytes);
    uint4* dc = reinterpret_cast<uint4*>(xs + (long long)row_c * row_bytes);
    uint4* da = reinterpret_cast<uint4*>(xs + (long long)row_a * row_bytes);
    for (int k = lane; k < row_bytes / 16; k += 32) {
      const uint4 v = __ldg(src + k);
      dc[k] = v;
      da[k] = v;
    }
  }
  // ---- completion ticket: the last CTA finalises
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(done_ticket, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ int f_cnt[2 * E], f_tile0[2 * E + 1];
  const int nb = 2 * E;
  if (threadIdx.x == 0) {
    int tiles = 0;
    float tot = 0.f;
    float us[2 * E];
    for (int k = 0; k < nb; ++k) {
      const int c = *reinterpret_cast<volatile int*>(&bucket_count[k]);
      f_cnt[k] = c;
      f_tile0[k] = tiles;
      for (int t = 0; t * 128 < c; ++t) {
        tile_group[tiles] = k;
        tile_row0[tiles] = k * cap_rows + t * 128;
        ++tiles;
      }
      us[k] = *reinterpret_cast<volatile float*>(&usage[k]);
      tot += us[k];
    }
    f_tile0[nb] = tiles;
    *m_tiles_used = tiles;
    if (lb_out) {
      float loss = 0.f;
      const float total = tot * (float)E;
      for (int k = 0; k < nb; ++k) {
        const float u = us[k] / (total + 1e-10f);
        loss += u * logf(u + 1e-10f);
      }
      *lb_out = loss / (float)nb;
    }
    *done_ticket = 0;  // ready for the next launch (graph replay)
  }
  __syncthreads();
  // padding rows of each bucket's last (partial) tile: invalid, so the grouped GEMMs drop them
  for (int k = 0; k < nb; ++k) {
    const int c = f_cnt[k], end = (c + 127) / 128 * 128;
    for (int r = c + threadIdx.x; r < end; r += blockDim.x) { row_map[k * cap_rows + r] = -1; row_scale[k * cap_rows + r] = 0.f; }
  }
  __syncthreads();
  if (threadIdx.x < nb) {
    // this block's counters are consumed: zero them for the next evaluation; and the next block's too
    bucket_count[threadIdx.x] = 0; usage[threadIdx.x] = 0.f;
    if (next_count) { next_count[threadIdx.x] = 0; next_usage[threadIdx.x] = 0.f; }
  }
}

// Step-invariant operands of the VB_EPI_ROUTE GEMM (include/versband.h): one CTA per (batch, head),
// thread j = key. kv = [B*L, 2*H*hd] fp32 rows (k | v) of the MoE cross-attention's in-projection.
//   bias_s[b][h][j] = <k_h[j], bq_h>            (bq already carries 1/sqrt(hd)); -1e30 for j >= L
//   vg[b][h][j][e]  = <v_h[j], capg_w[e][h*hd : (h+1)*hd]>;                        0 for j >= L
__global__ void route_prepare_kernel(const float* __restrict__ kv, int L, int H, int hd,
                                     const float* __restrict__ bq, const float* __restrict__ capg_w,
                                     float* __restrict__ bias_s, float* __restrict__ vg) {
  pdl_sync();
  const int h = blockIdx.x, b = blockIdx.y, j = threadIdx.x;  // blockDim.x == 96
  const int D = H * hd;
  float s = -1e30f;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j < L) {
    const float* kr = kv + ((long long)b * L + j) * 2 * D + h * hd;
    const float* vr = kr + D;
    s = 0.f;
    for (int d = 0; d < hd; ++d) {
      s = fmaf(kr[d], bq[h * hd + d], s);
      const float v = vr[d];
      g.x = fmaf(v, capg_w[0 * D + h * hd + d], g.x);
      g.y = fmaf(v, capg_w[1 * D + h * hd + d], g.y);
      g.z = fmaf(v, capg_w[2 * D + h * hd + d], g.z);
      g.w = fmaf(v, capg_w[3 * D + h * hd + d], g.w);
    }
  }
  const long long o = ((long long)b * H + h) * 96 + j;
  bias_s[o] = s;
  reinterpret_cast<float4*>(vg)[o] = g;
}

}  // namespace vb

using namespace vb;

extern "C" int vb_linear_small(const float* x, int64_t ldx, const float* W, const float* bias,
                               float* out, int rows, int D, int E, void* stream) {
  vb::ProfScope prof_scope(vb::PK_ROUTER, stream);
  if (!x || !W || !out || rows <= 0 || E <= 0 || E > 16)
    return set_error(VB_ERR_INVALID, "vb_linear_small: bad arguments (E <= 16)");
  launch_k(linear_small_kernel, dim3((rows * E + 7) / 8), dim3(256), 0, static_cast<cudaStream_t>(stream), x, ldx, W, bias,
                                                                                     out, rows, D, E);
  return check_launch("vb_linear_small");
}

extern "C" int vb_router(const float* cap_feat, const void* cap_feat_fp16, const float* Wc, const float* bc,
                         const float* ac_logits, int ac_ld, const float* hl_logits, int hl_ld,
                         const float* e0, const float* e1, const float* e2, float tau, int N, int T, int D, int E,
                         int32_t* idx_c, int32_t* idx_a, float* scale_c, float* scale_a,
                         float* cap_logits_out, int32_t* bucket_count, float* usage, int32_t* slot_c,
                         int32_t* slot_a, const float* cap_partial, int n_heads, void* stream) {
  vb::ProfScope prof_scope(vb::PK_ROUTER, stream);
  if ((!cap_feat && !cap_feat_fp16 && !cap_partial) || (!Wc && !cap_partial) || !bc || !ac_logits || !hl_logits || !e0 || !e1 || !e2 || !idx_c || !idx_a ||
      !scale_c || !scale_a || N <= 0 || T <= 0)
    return set_error(VB_ERR_INVALID, "vb_router: null/invalid argument");
  if (E != 4) return set_error(VB_ERR_INVALID, "vb_router: num_experts %d not instantiated (4)", E);
  launch_k(router_kernel<4>, dim3((N + 15) / 16), dim3(512), 0, static_cast<cudaStream_t>(stream), 
      cap_feat, static_cast<const __half*>(cap_feat_fp16), Wc, bc, ac_logits, hl_logits, hl_ld, ac_ld, e0, e1, e2, tau, N, T, D, idx_c, idx_a, scale_c, scale_a,
      cap_logits_out, bucket_count, usage, slot_c, slot_a, cap_partial, n_heads);
  return check_launch("vb_router");
}

extern "C" int vb_route_prepare(const float* kv, int B, int L, int H, int hd, const float* bq, const float* capg_w,
                                float* bias_s, float* vg, void* stream) {
  vb::ProfScope prof_scope(vb::PK_ROUTER, stream);
  if (!kv || !bq || !capg_w || !bias_s || !vg || B <= 0 || L <= 0 || L > 96 || H <= 0 || hd <= 0)
    return set_error(VB_ERR_INVALID, "vb_route_prepare: bad arguments (1 <= L <= 96)");
  launch_k(route_prepare_kernel, dim3(H, B), dim3(96), 0, static_cast<cudaStream_t>(stream), kv, L, H, hd, bq, capg_w,
           bias_s, vg);
  return check_launch("vb_route_prepare");
}

extern "C" int vb_moe_sort(const int32_t* idx_c, const int32_t* idx_a, const float* scale_c,
                           const float* scale_a, int N, int E, int32_t* row_map, float* row_scale,
                           int32_t* tile_group, int32_t* m_tiles_used, int max_rows, float* lb_out,
                           void* stream) {
  vb::ProfScope prof_scope(vb::PK_MOE_SORT, stream);
  if (!idx_c || !idx_a || !row_map || !row_scale || !tile_group || !m_tiles_used || E <= 0 || E > 8)
    return set_error(VB_ERR_INVALID, "vb_moe_sort: bad arguments");
  if (max_rows < ((2 * N + 127) / 128 + 2 * E) * 128)
    return set_error(VB_ERR_INVALID, "vb_moe_sort: max_rows %d too small for N=%d E=%d", max_rows, N, E);
  launch_k(moe_sort_kernel, dim3(1), dim3(512), 0, static_cast<cudaStream_t>(stream), 
      idx_c, idx_a, scale_c, scale_a, N, E, row_map, row_scale, tile_group, m_tiles_used, max_rows, lb_out);
  return check_launch("vb_moe_sort");
}

extern "C" int vb_gather_rows(const void* x, const int32_t* row_map, const int32_t* m_tiles_used,
                              int n_tokens, int D, void* out, int max_rows, void* stream) {
  vb::ProfScope prof_scope(vb::PK_MOE_SORT, stream);
  if (!x || !row_map || !out || D % 8 != 0) return set_error(VB_ERR_INVALID, "vb_gather_rows: bad arguments");
  launch_k(gather_rows_kernel, dim3((max_rows + 7) / 8), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __half*>(x), row_map, m_tiles_used, n_tokens, D,
      static_cast<__half*>(out), max_rows);
  return check_launch("vb_gather_rows");
}

extern "C" int vb_moe_finalize(int32_t* bucket_count, float* usage, int E, int32_t* bucket_start,
                               int32_t* row_map, float* row_scale, int32_t* tile_group,
                               int32_t* m_tiles_used, int max_rows, float* lb_out, void* stream) {
  vb::ProfScope prof_scope(vb::PK_MOE_SORT, stream);
  if (!bucket_count || !usage || !bucket_start || !row_map || !row_scale || !tile_group || !m_tiles_used ||
      E <= 0 || E > 8)
    return set_error(VB_ERR_INVALID, "vb_moe_finalize: bad arguments");
  launch_k(moe_finalize_kernel, dim3(1), dim3(256), 0, static_cast<cudaStream_t>(stream), bucket_count, usage,
           2 * E, E, bucket_start, row_map, row_scale, tile_group, m_tiles_used, max_rows, lb_out);
  return check_launch("vb_moe_finalize");
}

extern "C" int vb_moe_scatter(const int32_t* idx_c, const int32_t* idx_a, const int32_t* slot_c,
                              const int32_t* slot_a, const float* scale_c, const float* scale_a,
                              const int32_t* bucket_start, int N, int E, int row_bytes, const void* x,
                              void* xs, int32_t* row_map, float* row_scale, void* stream) {
  vb::ProfScope prof_scope(vb::PK_MOE_SORT, stream);
  if (!idx_c || !idx_a || !slot_c || !slot_a || !bucket_start || !x || !xs || row_bytes % 16)
    return set_error(VB_ERR_INVALID, "vb_moe_scatter: bad arguments");
  launch_k(moe_scatter_kernel, dim3((2 * N + 7) / 8), dim3(256), 0, static_cast<cudaStream_t>(stream), idx_c,
           idx_a, slot_c, slot_a, scale_c, scale_a, bucket_start, N, E, row_bytes,
           static_cast<const uint8_t*>(x), static_cast<uint8_t*>(xs), row_map, row_scale);
  return check_launch("vb_moe_scatter");
}

extern "C" int vb_moe_dispatch(const int32_t* idx_c, const int32_t* idx_a, const int32_t* slot_c,
                               const int32_t* slot_a, const float* scale_c, const float* scale_a,
                               const int32_t* bucket_count, const float* usage, int32_t* next_count,
                               float* next_usage, int N, int E, int row_bytes, const void* x, void* xs,
                               int32_t* row_map, float* row_scale, int32_t* tile_group,
                               int32_t* m_tiles_used, int max_rows, float* lb_out, void* stream) {
  vb::ProfScope prof_scope(vb::PK_MOE_SORT, stream);
  if (!idx_c || !idx_a || !slot_c || !slot_a || !bucket_count || !usage || !x || !xs || !row_map || !row_scale ||
      !tile_group || !m_tiles_used || row_bytes % 16 || E <= 0 || E > 8)
    return set_error(VB_ERR_INVALID, "vb_moe_dispatch: bad arguments");
  if (max_rows < ((2 * N + 127) / 128 + 2 * E) * 128)
    return set_error(VB_ERR_INVALID, "vb_moe_dispatch: max_rows %d too small for N=%d E=%d", max_rows, N, E);
  launch_k(moe_dispatch_kernel, dim3((2 * N + 7) / 8), dim3(256), 0, static_cast<cudaStream_t>(stream), idx_c, idx_a,
           slot_c, slot_a, scale_c, scale_a, bucket_count, usage, next_count, next_usage, N, E, row_bytes,
           static_cast<const uint8_t*>(x), static_cast<uint8_t*>(xs), row_map, row_scale, tile_group, m_tiles_used,
           max_rows, lb_out);
  return check_launch("vb_moe_dispatch");
}

extern "C" int vb_route_dispatch(const float* cap_partial, int n_heads, const float* bc, const float* ac_logits,
                                 int ac_ld, const float* hl_logits, int hl_ld, const float* e0, const float* e1,
                                 const float* e2, float tau, int N, int T, int E, int row_bytes, const void* x,
                                 void* xs, int cap_rows, int32_t* idx_c, int32_t* idx_a, float* cap_logits_out,
                                 int32_t* bucket_count, float* usage, int32_t* done_ticket, int32_t* next_count,
                                 float* next_usage, int32_t* row_map, float* row_scale, int32_t* tile_group,
                                 int32_t* tile_row0, int32_t* m_tiles_used, float* lb_out, void* stream) {
  vb::ProfScope prof_scope(vb::PK_ROUTER, stream);
  if (!cap_partial || !bc || !ac_logits || !hl_logits || !e0 || !e1 || !e2 || !x || !xs || !idx_c || !idx_a ||
      !bucket_count || !usage || !done_ticket || !row_map || !row_scale || !tile_group || !tile_row0 || !m_tiles_used ||
      N <= 0 || T <= 0 || n_heads <= 0 || row_bytes % 16)
    return set_error(VB_ERR_INVALID, "vb_route_dispatch: null/invalid argument");
  if (E != 4) return set_error(VB_ERR_INVALID, "vb_route_dispatch: num_experts %d not instantiated (4)", E);
  if (cap_rows < N || cap_rows % 128)
    return set_error(VB_ERR_INVALID, "vb_route_dispatch: cap_rows %d must be a multiple of 128 and >= N=%d", cap_rows, N);
  launch_k(route_dispatch_kernel<4>, dim3((N + 7) / 8), dim3(256), 0, static_cast<cudaStream_t>(stream), cap_partial,
           n_heads, bc, ac_logits, ac_ld, hl_logits, hl_ld, e0, e1, e2, tau, N, T, row_bytes,
           static_cast<const uint8_t*>(x), static_cast<uint8_t*>(xs), cap_rows, idx_c, idx_a, cap_logits_out, bucket_count,
           usage, done_ticket, next_count, next_usage, row_map, row_scale, tile_group, tile_row0, m_tiles_used, lb_out);
  return check_launch("vb_route_dispatch");
}
