// Shared device helpers (sm_100a PTX wrappers) and host-side error plumbing for libvllmini_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "vllmini_b200.h"

namespace vb200 {

// ------------------------------------------------------------------------------------------
// host: error state, device guard, launch counter
// ------------------------------------------------------------------------------------------
char* error_buffer();              // thread-local, 512 bytes (abi.cu)
void count_launch();               // atomic ++ of the library-wide launch counter (abi.cu)
int device_sm_count();             // SMs of the current device (cached; 148 if it cannot be queried)

inline int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define VB200_CUDA_CHECK(expr)                                                                   \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::vb200::set_error(VB200_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

// Equivalent of the reference's at::cuda::OptionalCUDAGuard (attention_kernels.cu:736,
// cache_kernels.cu:276): make `device` current for the launch, restore on scope exit.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t status = cudaSuccess;
  explicit DeviceGuard(int device, bool enabled = true) {
    if (!enabled) return;  // shape-only queries never touch the CUDA runtime
    status = cudaGetDevice(&prev);
    if (status == cudaSuccess && device >= 0 && device != prev) {
      status = cudaSetDevice(device);
      switched = (status == cudaSuccess);
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

// Programmatic dependent launch (PDL): when enabled, kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization so that their CTAs can become resident while the previous
// kernel on the stream drains; every kernel calls griddep_wait() before it touches global memory, which blocks
// until the previous kernel has completed and flushed (so stream-order semantics are unchanged).
bool pdl_enabled();                // abi.cu (vb200_set_pdl)
void disable_pdl();

// cluster_z > 1 launches thread-block clusters of (1, 1, cluster_z) CTAs (grid.z must be a multiple of it).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t stream, unsigned cluster_z, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_z > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = 1;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = cluster_z;
    ++n;
  }
  const int n_fixed = n;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  if (e == cudaErrorNotSupported && n > n_fixed) {  // driver / capture mode without programmatic launches
    cudaGetLastError();
    disable_pdl();
    cfg.numAttrs = n_fixed;
    e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  }
  return e;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args... args) {
  return launch_kernel_cluster(kern, grid, block, smem, stream, 1u, args...);
}

// ------------------------------------------------------------------------------------------
// device: PTX wrappers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// PDL: let the next kernel on the stream start scheduling its CTAs / block until the previous kernel is done.
// Both are no-ops when the kernel was launched without the programmatic attribute.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- thread-block cluster: barrier over all CTAs of the cluster, distributed shared memory ----
// (release/acquire: shared-memory writes before the arrive are visible to every CTA of the cluster after the wait)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// one lane of the (converged) warp: true in exactly one lane (elect.sync)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier (shared::cta) ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  // make the initialised barriers visible to the async (TMA) proxy
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// named barrier over N threads (ids 1..15; 0 is __syncthreads)
template <int ID, int N>
__device__ __forceinline__ void named_bar_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory");
}
// non-blocking probe: true if the phase with the given parity has completed
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "VB200_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra VB200_DONE;\n"
      "bra VB200_WAIT;\n"
      "VB200_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// order this thread's prior generic-proxy global writes before later async-proxy (bulk copy) reads
__device__ __forceinline__ void fence_proxy_async_global() {
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// ---- 1-D bulk async copy global -> shared, completion on an mbarrier (SASS: UBLKCP) ----
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---- vector loads / stores ----
// volatile + "memory": every use sits after griddep_wait() (or a barrier) and reads what the PREVIOUS kernel on the
// stream may have written (q, the new K/V rows) -- the compiler must not hoist it above that wait.  Independent
// loads still issue back to back; only their order relative to other memory operations is pinned.
__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"(addr)
               : "memory");
  return r;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
// L2-only load (never allocates in L1): for SPECULATIVE reads ahead of griddepcontrol.wait -- a later ordinary load of
// the same address must not find a stale line in this SM's L1
__device__ __forceinline__ int32_t ld_cg_s32(const int32_t* p) {
  int32_t r;
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}
// hint: bring the 128-byte line holding p into L2; no destination, no fault on an unmapped address
// (tools/prefetch_wild_repro.cu)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p) : "memory"); }
__device__ __forceinline__ float ld_cg_f32(const float* p) {
  float r;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
  return r;
}

// ---- mixed-precision FMA: d = lo/hi(a) * lo/hi(b) + c, fp16 x fp16 exact, fp32 accumulate.
//      PTX fma.rn.f32.f16 (sm_100+); SASS FHFMA with free .H0/.H1 operand selects. ----
__device__ __forceinline__ float fhfma_lo(uint32_t a, uint32_t b, float c) {
  float d;
  asm("{\n"
      ".reg .b16 al, ah, bl, bh;\n"
      "mov.b32 {al, ah}, %1;\n"
      "mov.b32 {bl, bh}, %2;\n"
      "fma.rn.f32.f16 %0, al, bl, %3;\n"
      "}\n"
      : "=f"(d)
      : "r"(a), "r"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fhfma_hi(uint32_t a, uint32_t b, float c) {
  float d;
  asm("{\n"
      ".reg .b16 al, ah, bl, bh;\n"
      "mov.b32 {al, ah}, %1;\n"
      "mov.b32 {bl, bh}, %2;\n"
      "fma.rn.f32.f16 %0, ah, bh, %3;\n"
      "}\n"
      : "=f"(d)
      : "r"(a), "r"(b), "f"(c));
  return d;
}

// ---- warp-wide fp32 max in one instruction (redux.sync.max.f32, sm_100a; SASS CREDUX.MAX.F32) ----
__device__ __forceinline__ float warp_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

__device__ __forceinline__ float ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

#endif  // __CUDACC__

}  // namespace vb200
