// Internal helpers shared by the translation units of libversband_sm100.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/versband.h"

namespace vb {

// printf-style: records the thread-local error string and returns `code`.
int set_error(int code, const char* fmt, ...);
// Bumps the process-wide kernel-launch counter (vb_launch_count()).
void count_launch();

typedef CUresult (*tensor_map_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                         const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                         const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion,
                                         CUtensorMapFloatOOBfill);
// Resolved once through cudaGetDriverEntryPoint (no link-time libcuda dependency).
tensor_map_encode_fn get_tensor_map_encoder();

// Optional per-launch CUDA-event timing by kernel family (vb_prof_enable / vb_prof_report).
enum ProfKind { PK_GEMM = 0, PK_NORM, PK_SOFTMAX, PK_ATTN_SMALL, PK_ROUTER, PK_MOE_SORT, PK_EW, PK_LAYOUT,
                PK_GN, PK_CONV_POST, PK_COUNT };
bool prof_enabled();
void prof_start(int kind, cudaStream_t s);
void prof_stop(int kind, cudaStream_t s);
struct ProfScope {
  int kind; cudaStream_t s; bool on;
  ProfScope(int k, void* stream) : kind(k), s(static_cast<cudaStream_t>(stream)), on(prof_enabled()) {
    if (on) prof_start(kind, s);
  }
  ~ProfScope() { if (on) prof_stop(kind, s); }
};

// Programmatic dependent launch (PDL): every kernel of this library starts with pdl_sync() and is
// launched with programmaticStreamSerializationAllowed, so the launch latency and prologue of
// kernel N+1 overlap the execution of kernel N. griddepcontrol.wait blocks until the preceding
// grid has fully completed (memory visible), which keeps stream-order semantics intact.
bool pdl_enabled();
#ifdef __CUDACC__
// fp32 -> fp16 operand conversion, round-to-nearest-even, saturating to +-65504 (never inf)
__device__ __forceinline__ __half2 vb_f2h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ __half vb_f2h(float x) {
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return *reinterpret_cast<__half*>(&r);
}
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// same, as thread-block clusters of `cluster_x` CTAs
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_cluster(void (*kern)(KArgs...), int cluster_x, dim3 grid, dim3 block, size_t smem,
                                    cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// Per-device one-time state (function attributes are per device; SM counts differ between SKUs): a process
// that drives several GPUs initialises each of them, not just whichever was current first.
template <typename T>
struct PerDevice {
  T v[64] = {};
  T& get() {
    int d = 0;
    cudaGetDevice(&d);
    return v[d & 63];
  }
};
inline int device_sm_count() {
  static PerDevice<int> sms;
  int& n = sms.get();
  if (n == 0) {
    int d = 0;
    cudaGetDevice(&d);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
    if (n <= 0) n = 148;
  }
  return n;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  count_launch();
  return VB_OK;
}

}  // namespace vb
