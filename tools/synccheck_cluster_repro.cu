// Minimal stand-alone kernel for `compute-sanitizer --tool synccheck`: does the tool track an mbarrier of a CTA whose
// rank in a thread-block cluster is >= 8 (non-portable cluster sizes 9..16)?
//
// Grid (12, 2, n) in clusters of (1,1,n).  Mode 0: every CTA does the textbook sequence on ONE mbarrier in its own
// shared memory: init(count 1) -> fence -> __syncthreads -> [thread 0] arrive.expect_tx + cp.async.bulk (1 KB) -> [all]
// try_wait parity 0 -> barrier.cluster -> exit.  Mode 1: the product's pattern (per-warp two-slot rings, several
// phases per barrier, see below).  Every arrive is matched by a wait of all the threads that use the data, so a
// correct tool reports nothing for any n.  Usage: synccheck_cluster_repro <mode> <n> [<n> ...]; prints the checksum per n (all CTAs must see the data).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/bin/synccheck_cluster_repro tools/synccheck_cluster_repro.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

// mode 0: one barrier, one phase.  mode 2 = mode 1 on a grid of 4800 clusters (several waves).  mode 1: the product's pattern -- every warp owns a ring of two 1 KB slots with one
// mbarrier each and streams `rounds` stages through it (lane 0: arrive.expect_tx + bulk copy; all lanes: try_wait on
// the slot's parity; refill after use), the ranks of a cluster running different numbers of rounds.
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D;\n"
      "bra W;\n"
      "D:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void issue_copy(uint32_t dst, const void* src, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 1024;" ::"r"(bar) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 1024, [%2];" ::"r"(dst),
               "l"(src), "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(128) repro_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ out,
                                                    int mode) {
  __shared__ __align__(128) uint32_t buf[4 * 2 * 256];   // [warp][slot][256 words]
  __shared__ __align__(8) uint64_t bars[4 * 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  const uint32_t buf0 = (uint32_t)__cvta_generic_to_shared(buf);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + i * 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  uint32_t s = 0;
  if (mode == 0) {
    if (threadIdx.x == 0) issue_copy(buf0, src + 256 * blockIdx.z, bar0);
    mbar_wait_parity(bar0, 0);
    s = buf[threadIdx.x] + buf[threadIdx.x + 128];
  } else {   // modes 1 and 2
    const int rounds = 3 + (int)(blockIdx.z % 5);          // stages of this warp; ranks differ
    const uint32_t* mine = src + 256 * blockIdx.z;          // every stage re-reads the rank's 1 KB
    const uint32_t wbar = bar0 + warp * 16, wbuf = buf0 + warp * 2048;
    for (int k = 0; k < 2 && k < rounds; ++k) {
      if (lane == 0) issue_copy(wbuf + k * 1024, mine, wbar + k * 8);
      __syncwarp();
    }
    for (int k = 0; k < rounds; ++k) {
      const int slot = k & 1;
      mbar_wait_parity(wbar + slot * 8, (k >> 1) & 1);
      const uint32_t* b = buf + warp * 512 + slot * 256;
      if (k == rounds - 1 && warp == 0)
        for (int i = lane; i < 256; i += 32) s += b[i];
      __syncwarp();
      if (k + 2 < rounds) {
        if (lane == 0) issue_copy(wbuf + slot * 1024, mine, wbar + slot * 8);
        __syncwarp();
      }
    }
  }
  atomicAdd(out, s);
  if (threadIdx.x == 0) atomicAdd(out + 1, bar0 >> 24);   // the rank bits of the shared-window address, summed
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main(int argc, char** argv) {
  uint32_t *src, *out;
  const int max_n = 16;
  cudaMalloc(&src, max_n * 1024);
  cudaMalloc(&out, 8);
  uint32_t* h = (uint32_t*)malloc(max_n * 1024);
  for (int i = 0; i < max_n * 256; ++i) h[i] = i;
  cudaMemcpy(src, h, max_n * 1024, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(repro_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  for (int a = 2; a < argc; ++a) {
    const int n = atoi(argv[a]);
    if (n < 1 || n > max_n) continue;
    cudaMemset(out, 0, 8);
    cudaLaunchConfig_t cfg{};
    // several clusters, like the product's (heads, seqs, splits); mode 2: more clusters than the GPU holds at once
    cfg.gridDim = dim3(12, mode == 2 ? 400 : 2, n);
    cfg.blockDim = dim3(128);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = n;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, repro_kernel, (const uint32_t*)src, out, mode);
    cudaError_t e2 = cudaDeviceSynchronize();
    uint32_t r[2] = {0, 0};
    cudaMemcpy(r, out, 8, cudaMemcpyDeviceToHost);
    uint32_t want = 0;
    for (int i = 0; i < n * 256; ++i) want += (mode == 2 ? 4800u : 24u) * h[i];
    printf("mode %d cluster (1,1,%d): launch=%s sync=%s checksum %s (%u), sum of address rank bits %u (expect %d)\n", mode, n,
           cudaGetErrorString(e), cudaGetErrorString(e2), r[0] == want ? "ok" : "WRONG", r[0], r[1], (mode == 2 ? 4800 : 24) * n * (n - 1) / 2);
  }
  return 0;
}
