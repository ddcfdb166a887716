// Library-level entry points: version, status strings, last CUDA error.
#include "hia_common.cuh"
#include <stdlib.h>

namespace hia {
static thread_local char g_last_err[256] = "";
void set_last_cuda_error(cudaError_t e) {
  const char* s = cudaGetErrorString(e);
  strncpy(g_last_err, s ? s : "unknown", sizeof(g_last_err) - 1);
  g_last_err[sizeof(g_last_err) - 1] = 0;
  cudaGetLastError();  // clear the sticky-free error state
}
static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
// read on every call (a getenv, next to kernel launches): tools time both settings in one process
int sel_aggregate() {
  const char* e = getenv("HIA_SEL_AGG");
  return !(e && e[0] == '0');
}
int num_sms() {
  static int cached = 0;
  int v = __atomic_load_n(&cached, __ATOMIC_RELAXED);
  if (v == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) {
      (void)cudaGetLastError();
      return 148;                       // B200; not cached, so a later call can still query
    }
    __atomic_store_n(&cached, v, __ATOMIC_RELAXED);
  }
  return v;
}
}  // namespace hia

HIA_API uint64_t hia_launch_count(void) { return __atomic_load_n(&hia::g_launches, __ATOMIC_RELAXED); }

HIA_API int hia_version(void) { return 100; }

HIA_API const char* hia_last_cuda_error(void) { return hia::g_last_err; }

HIA_API const char* hia_status_string(int s) {
  switch (s) {
    case HIA_OK: return "ok";
    case HIA_ERR_BAD_ARG: return "bad argument";
    case HIA_ERR_ALIGN: return "alignment requirement not met";
    case HIA_ERR_CUDA: return "CUDA error";
    case HIA_ERR_UNSUPPORTED: return "unsupported size";
    case HIA_ERR_WORKSPACE: return "workspace too small";
    case HIA_ERR_ARCH: return "device is not sm_100";
    default: return "unknown status";
  }
}
