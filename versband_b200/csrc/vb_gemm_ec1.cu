// vb_gemm kernels specialised for epilogue class EC_LINEAR (see vb_gemm_params.h)
#include "vb_gemm_kernels.cuh"

template int vb::launch_gemm_class<vb::EC_LINEAR>(int, bool, bool, const CUtensorMap&, const CUtensorMap&, const vb::GemmKParams&,
                                               int, cudaStream_t);
template int vb::launch_halo_class<vb::EC_LINEAR>(int, bool, const CUtensorMap&, const CUtensorMap&, const vb::GemmKParams&,
                                               const vb::HaloParams&, int, int, cudaStream_t);
