// vb_gemm_chain kernel instances (several dependent GEMM problems per persistent launch), see vb_gemm_kernels.cuh
#include "vb_gemm_kernels.cuh"

namespace vb {
int launch_chain(int bn, bool tf32, const ChainParams& cp, int grid, cudaStream_t stream) {
  if (bn == 64) return tf32 ? launch_chain_t<64, true>(cp, grid, stream) : launch_chain_t<64, false>(cp, grid, stream);
  if (bn == 128) return tf32 ? launch_chain_t<128, true>(cp, grid, stream) : launch_chain_t<128, false>(cp, grid, stream);
  return set_error(VB_ERR_INVALID, "vb_gemm_chain: N tile %d not instantiated (64/128)", bn);
}
}  // namespace vb
