// Minimal stand-alone kernel for `compute-sanitizer --tool synccheck`: does the tool track an mbarrier of a CTA whose
// rank in a thread-block cluster is >= 8 (non-portable cluster sizes 9..16)?
//
// Every CTA of a (1,1,n) cluster does the textbook sequence on ONE mbarrier in its own shared memory:
//   init(count 1) -> fence -> __syncthreads -> [thread 0] arrive.expect_tx + cp.async.bulk (1 KB) -> [all] try_wait
//   parity 0 -> barrier.cluster -> exit.
// There is exactly one arrive and every thread waits for the phase it completes, so a correct tool reports nothing
// for any n.  Usage: synccheck_cluster_repro <n> [<n> ...]; prints the checksum per n (all CTAs must see the data).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/bin/synccheck_cluster_repro tools/synccheck_cluster_repro.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void __launch_bounds__(128) repro_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ out) {
  __shared__ __align__(128) uint32_t buf[256];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t buf_a = (uint32_t)__cvta_generic_to_shared(buf);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 1024;" ::"r"(bar_a) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 1024, [%2];" ::"r"(buf_a),
                 "l"(src + 256 * blockIdx.z), "r"(bar_a)
                 : "memory");
  }
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@p bra D;\n"
      "bra W;\n"
      "D:\n"
      "}\n" ::"r"(bar_a)
      : "memory");
  uint32_t s = buf[threadIdx.x] + buf[threadIdx.x + 128];
  atomicAdd(out, s);
  if (threadIdx.x == 0) atomicAdd(out + 1, bar_a >> 24);   // the rank bits of the shared-window address, summed
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
  for (int a = 1; a < argc; ++a) {
    const int n = atoi(argv[a]);
    if (n < 1 || n > max_n) continue;
    cudaMemset(out, 0, 8);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1, 1, n);
    cfg.blockDim = dim3(128);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = n;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, repro_kernel, (const uint32_t*)src, out);
    cudaError_t e2 = cudaDeviceSynchronize();
    uint32_t r[2] = {0, 0};
    cudaMemcpy(r, out, 8, cudaMemcpyDeviceToHost);
    uint32_t want = 0;
    for (int i = 0; i < n * 256; ++i) want += h[i];
    printf("cluster (1,1,%d): launch=%s sync=%s checksum %s (%u), sum of address rank bits %u (expect %d)\n", n,
           cudaGetErrorString(e), cudaGetErrorString(e2), r[0] == want ? "ok" : "WRONG", r[0], r[1], n * (n - 1) / 2);
  }
  return 0;
}
