// Is `prefetch.global.L2` a pure hint on B200 -- i.e. silently dropped when the address is unmapped, misaligned, null,
// or points into another allocation?  (The attention kernel prefetches its first ring of K/V tiles from SPECULATIVELY
// read block-table entries before `griddepcontrol.wait`; a stale or uninitialised entry turns into an arbitrary
// address.)  Every thread prefetches a few thousand addresses derived from a hash: inside a small allocation, just
// past it, far beyond any mapping, negative offsets, the null page.  Expected: no error from the launch, the sync, or
// compute-sanitizer memcheck; the checksum kernel afterwards still reads the allocation correctly.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/bin/prefetch_wild_repro tools/prefetch_wild_repro.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void wild_prefetch(const char* base, long long span, int rounds) {
  unsigned long long x = (blockIdx.x * 1315423911ull + threadIdx.x) * 0x9E3779B97F4A7C15ull + 1;
  for (int r = 0; r < rounds; ++r) {
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;                      // xorshift
    const unsigned long long h = x * 0x2545F4914F6CDD1Dull;
    long long off;
    switch (r & 7) {
      case 0: off = (long long)(h % (unsigned long long)span); break;          // inside the allocation
      case 1: off = span + (long long)(h % 4096); break;                       // just past the end
      case 2: off = (long long)(h >> 20); break;                               // up to 2^44 bytes beyond
      case 3: off = -(long long)(h >> 22); break;                              // far below
      case 4: off = (long long)(int)(h >> 32) * 49152ll; break;                // garbage int32 block id x block stride
      case 5: off = -(long long)(uintptr_t)base; break;                        // the null page
      case 6: off = -(long long)(uintptr_t)base + (long long)(h >> 8); break;  // anywhere in the 56-bit space
      default: off = (long long)(h % (unsigned long long)span) | 1; break;     // misaligned, inside
    }
    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off) : "memory");
  }
}

__global__ void checksum(const unsigned* p, size_t n, unsigned long long* out) {
  unsigned long long s = 0;
  for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += p[i];
  atomicAdd(out, s);
}

int main() {
  const size_t n = 1 << 20;   // 4 MB
  unsigned* d;
  unsigned long long* out;
  cudaMalloc(&d, n * 4);
  cudaMalloc(&out, 8);
  cudaMemset(out, 0, 8);
  unsigned* h = (unsigned*)malloc(n * 4);
  unsigned long long want = 0;
  for (size_t i = 0; i < n; ++i) { h[i] = (unsigned)(i * 2654435761u); want += h[i]; }
  cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
  wild_prefetch<<<148 * 4, 128>>>((const char*)d, (long long)(n * 4), 4096);
  cudaError_t e1 = cudaGetLastError();
  cudaError_t e2 = cudaDeviceSynchronize();
  checksum<<<148, 256>>>(d, n, out);
  cudaError_t e3 = cudaDeviceSynchronize();
  unsigned long long got = 0;
  cudaMemcpy(&got, out, 8, cudaMemcpyDeviceToHost);
  printf("wild prefetch: launch=%s sync=%s; checksum after: %s (%s)\n", cudaGetErrorString(e1), cudaGetErrorString(e2),
         got == want ? "ok" : "WRONG", cudaGetErrorString(e3));
  return (e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess && got == want) ? 0 : 1;
}
