// Library-wide state behind the C ABI: thread-local error text, launch counter, device facts.
#include <atomic>
#include <cuda_runtime.h>

#include "common.cuh"
#include "vllmini_b200.h"

namespace vb200 {

static thread_local char t_error[512] = {0};
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_pdl{1};

bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }
void disable_pdl() { g_pdl.store(0, std::memory_order_relaxed); }

char* error_buffer() { return t_error; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    cudaGetLastError();  // clear the sticky-free error so later launches are not blamed for it
    return 148;
  }
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = 148;
    }
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace vb200

extern "C" const char* vb200_last_error(void) { return vb200::t_error; }
extern "C" int32_t vb200_abi_version(void) { return VB200_ABI_VERSION; }
extern "C" int32_t vb200_set_pdl(int32_t enable) {
  return vb200::g_pdl.exchange(enable ? 1 : 0, std::memory_order_relaxed);
}
extern "C" uint64_t vb200_launch_count(void) { return vb200::g_launches.load(std::memory_order_relaxed); }

