// vb_gemm kernels specialised for epilogue class EC_ROUTE (see vb_gemm_params.h)
#include "vb_gemm_kernels.cuh"

template int vb::launch_gemm_class<vb::EC_ROUTE>(int, bool, bool, const CUtensorMap&, const CUtensorMap&, const vb::GemmKParams&,
                                                 int, cudaStream_t);
