// The rest of the reference's cache_ops surface: copy_blocks, swap_blocks, reshape_and_cache_flash.
//
// None of these is called by vllmini: they are exported by
// the reference extension (paged_attention_cuda.cpp:57-60) and provided here so that the drop-in package
// covers the same surface.  All three are pure copies -> bit-exact.
//   copy_blocks              cache_kernels.cu:67-94 (kernel), :96-148 (entry)
//   swap_blocks              cache_kernels.cu:24-63
//   reshape_and_cache_flash  cache_kernels.cu:209-240 (kernel), :283-317 (entry)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "vllmini_b200.h"

namespace vb200 {

// grid (pairs, layers, 2): z = 0 key cache, 1 value cache.  16-byte vectors when everything is aligned.
template <typename V>
__global__ void __launch_bounds__(256) copy_blocks_kernel(const int64_t* __restrict__ key_cache_ptrs,
                                                          const int64_t* __restrict__ value_cache_ptrs,
                                                          const int64_t* __restrict__ block_mapping,
                                                          int64_t vecs_per_block) {
  const int pair = blockIdx.x, layer = blockIdx.y;
  V* cache = reinterpret_cast<V*>(blockIdx.z == 0 ? key_cache_ptrs[layer] : value_cache_ptrs[layer]);
  const int64_t src = block_mapping[2 * pair] * vecs_per_block;
  const int64_t dst = block_mapping[2 * pair + 1] * vecs_per_block;
  for (int64_t i = threadIdx.x; i < vecs_per_block; i += blockDim.x) cache[dst + i] = cache[src + i];
}

// one thread per 16-byte chunk of a token row; k_cache/v_cache [NB, B, H, D]
__global__ void __launch_bounds__(256) reshape_and_cache_flash_vec8_kernel(
    const __half* __restrict__ key, const __half* __restrict__ value, __half* __restrict__ k_cache,
    __half* __restrict__ v_cache, const int64_t* __restrict__ slot_mapping, int64_t total_chunks,
    int chunks_per_token, int block_size, int64_t block_stride, int64_t key_stride, int64_t value_stride) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total_chunks) return;
  const int64_t t = idx / chunks_per_token;
  const int c = (int)(idx - t * chunks_per_token);
  const int64_t slot = slot_mapping[t];
  if (slot < 0) return;  // padded token (cache_kernels.cu:223-226)
  const int64_t blk = slot / block_size, off = slot % block_size;
  const int64_t dst = blk * block_stride + off * (int64_t)chunks_per_token * 8 + (int64_t)c * 8;
  *reinterpret_cast<uint4*>(k_cache + dst) = ldg128(key + t * key_stride + (int64_t)c * 8);
  *reinterpret_cast<uint4*>(v_cache + dst) = ldg128(value + t * value_stride + (int64_t)c * 8);
}

__global__ void __launch_bounds__(256) reshape_and_cache_flash_generic_kernel(
    const uint16_t* __restrict__ key, const uint16_t* __restrict__ value, uint16_t* __restrict__ k_cache,
    uint16_t* __restrict__ v_cache, const int64_t* __restrict__ slot_mapping, int64_t total, int n, int block_size,
    int64_t block_stride, int64_t key_stride, int64_t value_stride) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t t = idx / n;
  const int i = (int)(idx - t * n);
  const int64_t slot = slot_mapping[t];
  if (slot < 0) return;
  const int64_t dst = (slot / block_size) * block_stride + (slot % block_size) * (int64_t)n + i;
  k_cache[dst] = key[t * key_stride + i];
  v_cache[dst] = value[t * value_stride + i];
}

}  // namespace vb200

using namespace vb200;

extern "C" int vb200_copy_blocks(const int64_t* key_cache_ptrs, const int64_t* value_cache_ptrs,
                                 const int64_t* block_mapping, int32_t num_layers, int32_t num_pairs,
                                 int64_t block_bytes, int32_t all_16b_aligned, int32_t device, void* stream) {
  if (num_layers < 0 || num_pairs < 0 || block_bytes <= 0)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "negative size argument");
  if (num_layers == 0 || num_pairs == 0) return VB200_OK;
  if (!key_cache_ptrs || !value_cache_ptrs || !block_mapping)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "null pointer argument");
  if (num_layers > 65535) return set_error(VB200_ERR_INVALID_ARGUMENT, "too many layers");
  DeviceGuard guard(device);
  if (guard.status != cudaSuccess)
    return set_error(VB200_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.status));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid(num_pairs, num_layers, 2);
  if (all_16b_aligned && block_bytes % 16 == 0) {
    copy_blocks_kernel<uint4><<<grid, 256, 0, s>>>(key_cache_ptrs, value_cache_ptrs, block_mapping, block_bytes / 16);
  } else if (block_bytes % 2 == 0) {
    copy_blocks_kernel<uint16_t><<<grid, 256, 0, s>>>(key_cache_ptrs, value_cache_ptrs, block_mapping, block_bytes / 2);
  } else {
    copy_blocks_kernel<uint8_t><<<grid, 256, 0, s>>>(key_cache_ptrs, value_cache_ptrs, block_mapping, block_bytes);
  }
  count_launch();
  VB200_CUDA_CHECK(cudaGetLastError());
  return VB200_OK;
}

extern "C" int vb200_swap_blocks(const void* src, void* dst, const int64_t* block_mapping_host, int64_t num_pairs,
                                 int64_t block_bytes, int32_t kind, int32_t device, void* stream) {
  if (num_pairs < 0 || block_bytes <= 0) return set_error(VB200_ERR_INVALID_ARGUMENT, "negative size argument");
  if (num_pairs == 0) return VB200_OK;
  if (!src || !dst || !block_mapping_host) return set_error(VB200_ERR_INVALID_ARGUMENT, "null pointer argument");
  cudaMemcpyKind k;
  switch (kind) {
    case 0: k = cudaMemcpyDeviceToDevice; break;
    case 1: k = cudaMemcpyDeviceToHost; break;
    case 2: k = cudaMemcpyHostToDevice; break;
    default: return set_error(VB200_ERR_INVALID_ARGUMENT, "Invalid device combination");
  }
  DeviceGuard guard(device);
  if (guard.status != cudaSuccess)
    return set_error(VB200_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.status));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const char* sp = static_cast<const char*>(src);
  char* dp = static_cast<char*>(dst);
  for (int64_t i = 0; i < num_pairs; ++i) {  // one async copy per block, as the reference (cache_kernels.cu:54-62)
    const int64_t so = block_mapping_host[2 * i] * block_bytes;
    const int64_t d_o = block_mapping_host[2 * i + 1] * block_bytes;
    VB200_CUDA_CHECK(cudaMemcpyAsync(dp + d_o, sp + so, (size_t)block_bytes, k, s));
  }
  return VB200_OK;
}

extern "C" int vb200_reshape_and_cache_flash_f16(const void* key, const void* value, void* k_cache, void* v_cache,
                                                 const int64_t* slot_mapping, int32_t num_tokens, int32_t num_heads,
                                                 int32_t head_size, int32_t block_size, int64_t block_stride,
                                                 int64_t key_stride, int64_t value_stride, int32_t device,
                                                 void* stream) {
  if (num_tokens < 0 || num_heads <= 0 || head_size <= 0 || block_size <= 0)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "negative or zero size argument");
  if (num_tokens == 0) return VB200_OK;
  if (!key || !value || !k_cache || !v_cache || !slot_mapping)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "null pointer argument");
  DeviceGuard guard(device);
  if (guard.status != cudaSuccess)
    return set_error(VB200_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.status));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = num_heads * head_size;
  const bool vec_ok = n % 8 == 0 && key_stride % 8 == 0 && value_stride % 8 == 0 && block_stride % 8 == 0 &&
                      ((((uintptr_t)key | (uintptr_t)value | (uintptr_t)k_cache | (uintptr_t)v_cache) & 15) == 0);
  if (vec_ok) {
    const int64_t total = (int64_t)num_tokens * (n / 8);
    reshape_and_cache_flash_vec8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
        static_cast<const __half*>(key), static_cast<const __half*>(value), static_cast<__half*>(k_cache),
        static_cast<__half*>(v_cache), slot_mapping, total, n / 8, block_size, block_stride, key_stride, value_stride);
  } else {
    const int64_t total = (int64_t)num_tokens * n;
    reshape_and_cache_flash_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
        static_cast<const uint16_t*>(key), static_cast<const uint16_t*>(value), static_cast<uint16_t*>(k_cache),
        static_cast<uint16_t*>(v_cache), slot_mapping, total, n, block_size, block_stride, key_stride, value_stride);
  }
  count_launch();
  VB200_CUDA_CHECK(cudaGetLastError());
  return VB200_OK;
}

#ifdef VB200_EXPERIMENTS
// Test helper of the experiments build (tests/test_parity_gpu.py::test_speculative_prefetch_reads_stale_indices): a
// kernel that behaves like a PDL-aware producer of the index tensors -- it lets its dependents launch at once, keeps
// them waiting for `delay_ns`, and only then writes the block table and seq_lens.  The attention launch queued behind
// it is resident (and speculating on whatever the two buffers held before) during the whole delay.
namespace vb200 {
__global__ void __launch_bounds__(256) debug_delayed_copy2_kernel(int32_t* dst1, const int32_t* src1, int64_t n1,
                                                                  int32_t* dst2, const int32_t* src2, int64_t n2,
                                                                  uint32_t delay_ns) {
  griddep_launch_dependents();
  griddep_wait();
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    __nanosleep(1000);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < delay_ns);
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = i0; i < n1; i += step) dst1[i] = src1[i];
  for (int64_t i = i0; i < n2; i += step) dst2[i] = src2[i];
}
}  // namespace vb200

extern "C" int vb200_debug_delayed_copy2(int32_t* dst1, const int32_t* src1, int64_t n1, int32_t* dst2,
                                         const int32_t* src2, int64_t n2, uint32_t delay_ns, void* stream) {
  VB200_CUDA_CHECK(vb200::launch_kernel(vb200::debug_delayed_copy2_kernel, dim3(8), dim3(256), 0,
                                        static_cast<cudaStream_t>(stream), dst1, src1, n1, dst2, src2, n2, delay_ns));
  return VB200_OK;
}
#endif
