// Library-level entry points: version, error string, device check, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <vector>
#include "vb_internal.h"

namespace vb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

tensor_map_encode_fn get_tensor_map_encoder() {
  static tensor_map_encode_fn fn = []() -> tensor_map_encode_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<tensor_map_encode_fn>(p);
  }();
  return fn;
}

static bool g_pdl = true;
bool pdl_enabled() { return g_pdl; }

// ---- per-family event profiler ----
struct ProfEv { cudaEvent_t a, b; int kind; };
static bool g_prof_on = false;
static std::vector<ProfEv> g_prof;
static std::mutex g_prof_mu;
bool prof_enabled() { return g_prof_on; }
void prof_start(int kind, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfEv e; e.kind = kind;
  cudaEventCreate(&e.a); cudaEventCreate(&e.b);
  cudaEventRecord(e.a, s);
  g_prof.push_back(e);
}
void prof_stop(int kind, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (size_t i = g_prof.size(); i-- > 0;)
    if (g_prof[i].kind == kind) { cudaEventRecord(g_prof[i].b, s); break; }
}

}  // namespace vb

extern "C" int vb_set_pdl(int on) {
  vb::g_pdl = on != 0;
  return VB_OK;
}

extern "C" int vb_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(vb::g_prof_mu);
  for (auto& e : vb::g_prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  vb::g_prof.clear();
  vb::g_prof_on = on != 0;
  return VB_OK;
}

// Synchronises the device and sums event-pair durations per kernel family.
extern "C" int vb_prof_report(double* ms_by_kind, long long* count_by_kind, int nkinds) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return vb::set_error(VB_ERR_CUDA, "vb_prof_report: %s", cudaGetErrorString(e));
  for (int i = 0; i < nkinds; ++i) { ms_by_kind[i] = 0; count_by_kind[i] = 0; }
  std::lock_guard<std::mutex> lk(vb::g_prof_mu);
  for (auto& ev : vb::g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess && ev.kind < nkinds) {
      ms_by_kind[ev.kind] += ms; count_by_kind[ev.kind] += 1;
    }
  }
  return VB_OK;
}

extern "C" int vb_version(void) { return VB_ABI_VERSION; }
extern "C" const char* vb_last_error(void) { return vb::g_err; }
extern "C" long long vb_launch_count(void) { return vb::g_launches.load(); }

extern "C" int vb_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return vb::set_error(VB_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10)
    return vb::set_error(VB_ERR_ARCH, "versband_b200 needs an sm_100 (B200) device, found sm_%d%d", major, minor);
  return VB_OK;
}

// ABI self-check used by the host binding: sizeof of the public argument structs.
extern "C" int vb_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(vb_gemm_args);
    case 1: return (int)sizeof(vb_attn_args);
    case 2: return (int)sizeof(vb_voc_desc);
    case 3: return (int)sizeof(vb_vae_desc);
    case 4: return (int)sizeof(vb_dit_desc);
    case 5: return (int)sizeof(vb_dit_cond);
    case 6: return (int)sizeof(vb_dit_io);
    case 7: return (int)sizeof(vb_dit_prep_desc);
    case 8: return (int)sizeof(vb_sample_args);
    default: return -1;
  }
}
