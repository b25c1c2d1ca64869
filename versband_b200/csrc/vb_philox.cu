// vb_router_noise — all Exp(1) draws of one sampling run in ONE launch, bit-identical to the
// sequence of `torch.empty(N, k, device='cuda').exponential_()` calls the reference issues
// (vocal2music_moe.py:81-83,134,150,151), i.e. PyTorch's Philox4x32-10 stream:
//   per call c (numel n_c): grid_c = min(ceil(n_c/256), 8*numSMs) blocks of 256 threads, thread idx
//   takes ONE Philox block with counter (offset_c/4, 0, idx, 0) and key = seed, its 4 outputs go to
//   elements idx + 256*grid_c*{0,1,2,3}; u = (x + 0.5) * 2^-32; E = -log(u) (u ~ 1 guarded);
//   every call advances the generator offset by 4 (n_c <= 1024*grid_c).
// The host side advances torch's generator offset accordingly; tests/test_philox_gpu.py checks
// bit-equality against torch on the GPU and the sampler falls back to torch's own draws otherwise.
#include "vb_internal.h"

namespace vb {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__global__ void router_noise_kernel(unsigned long long seed, unsigned long long offset0, int N, int E,
                                    int depth, float* __restrict__ e0, float* __restrict__ e1,
                                    float* __restrict__ e2, int max_grid) {
  pdl_sync();
  // call index c = ((step*2 + branch)*depth + block)*3 + which
  const int c = blockIdx.y;
  const int which = c % 3;
  const int rest = c / 3;
  const int k = rest % depth;
  const int sb = rest / depth;
  const int br = sb & 1, s = sb >> 1;
  const int w = which == 0 ? 2 : E;
  const int n = N * w;
  int grid_c = (n + 255) / 256;
  if (grid_c > max_grid) grid_c = max_grid;
  if ((int)blockIdx.x >= grid_c) return;
  float* dst = (which == 0 ? e0 : (which == 1 ? e1 : e2)) + (((long long)s * depth + k) * 2 * N + (long long)br * N) * w;
  const unsigned long long off = offset0 + 4ull * (unsigned long long)c;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int stride = 256 * grid_c;
  const int rounded = ((n - 1) / (stride * 4) + 1) * stride * 4;
  unsigned long long ctr_lo = off / 4;  // (offsets are multiples of 4: STATE == 0)
  for (int li0 = idx; li0 < rounded; li0 += stride * 4) {
    uint32_t ctr[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)idx, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(ctr, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int li = li0 + stride * ii;
      if (li < n) {
        const float u = (float)ctr[ii] * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
        const float lg = (u >= 1.0f - 5.9604644775390625e-08f) ? -5.9604644775390625e-08f : __logf(u);  // torch (at::log in device code) resolves to the fast intrinsic
        dst[li] = -1.0f * lg;
      }
    }
    ctr_lo += 1;
  }
}

}  // namespace vb

extern "C" int vb_router_noise(uint64_t seed, uint64_t offset, int N, int E, int depth, int steps,
                               float* e0, float* e1, float* e2, void* stream) {
  using namespace vb;
  ProfScope prof_scope(PK_EW, stream);
  if (!e0 || !e1 || !e2 || N <= 0 || E <= 0 || depth <= 0 || steps <= 0 || (offset & 3))
    return set_error(VB_ERR_INVALID, "vb_router_noise: bad arguments (offset must be a multiple of 4)");
  static PerDevice<int> max_grid_dev;
  int& max_grid = max_grid_dev.get();
  if (max_grid == 0) {
    int dev = 0, sms = 0, tpm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&tpm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
    max_grid = sms * (tpm / 256);
  }
  const int n_max = N * (E > 2 ? E : 2);
  if ((long long)n_max > 1024LL * ((n_max + 255) / 256 < max_grid ? (n_max + 255) / 256 : max_grid))
    return set_error(VB_ERR_INVALID, "vb_router_noise: numel too large for the single-iteration offset rule");
  int gx = (n_max + 255) / 256;
  if (gx > max_grid) gx = max_grid;
  dim3 grid(gx, steps * 2 * depth * 3);
  launch_k(router_noise_kernel, grid, dim3(256), 0, static_cast<cudaStream_t>(stream),
           (unsigned long long)seed, (unsigned long long)offset, N, E, depth, e0, e1, e2, max_grid);
  return check_launch("vb_router_noise");
}
