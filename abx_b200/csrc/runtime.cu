// Library-level entry points: last-error string, launch counter, device check.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace abx {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local bool g_pdl_scope = false;
bool pdl_scope_active() { return g_pdl_scope; }
void pdl_scope_set(bool on) { g_pdl_scope = on; }

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return ABX_ERR_CUDA;
  }
  return ABX_OK;
}

}  // namespace abx

extern "C" const char* abx_last_error(void) { return abx::g_err; }
extern "C" int abx_version(void) { return 1; }
extern "C" uint64_t abx_launch_count(void) { return abx::g_launches.load(std::memory_order_relaxed); }
extern "C" void abx_reset_launch_count(void) { abx::g_launches.store(0, std::memory_order_relaxed); }

extern "C" int abx_device_check(int device, int* sm_count) {
  cudaDeviceProp p;
  ABX_CUDA(cudaGetDeviceProperties(&p, device));
  ABX_REQUIRE(p.major == 10, "abx_b200 is built for sm_100a only; device %d is sm_%d%d", device, p.major, p.minor);
  if (sm_count) *sm_count = p.multiProcessorCount;
  return ABX_OK;
}
