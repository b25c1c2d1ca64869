// vb_gemm kernels specialised for epilogue class EC_RES16 (LINEAR + fp16 activated-representation residual)
#include "vb_gemm_kernels.cuh"

template int vb::launch_gemm_class<vb::EC_RES16>(int, bool, bool, const CUtensorMap&, const CUtensorMap&, const vb::GemmKParams&,
                                              int, cudaStream_t);
template int vb::launch_halo_class<vb::EC_RES16>(int, bool, const CUtensorMap&, const CUtensorMap&, const vb::GemmKParams&,
                                              const vb::HaloParams&, int, int, cudaStream_t);
