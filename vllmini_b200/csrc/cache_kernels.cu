// reshape_and_cache for sm_100a: scatter new K/V rows into the paged caches by slot.
//
// Behavioural spec: /root/reference/paged_attention_ext/paged_attention_cuda/cache_kernels.cu
//   :152-207 (kernel), :256-281 (host entry).  Pure copy ("auto" kv-cache dtype) -> bit-exact.
//
//   reference: one CTA per token, one thread per scalar: 2-byte loads, 2-byte stores.
//   here:      one thread per 16-byte chunk (8 dims of one head): one 128-bit load of K and of V,
//              one 128-bit store into the K cache (its layout keeps 8-dim groups contiguous:
//              key_cache[blk][h][d/8][off][0..7]) and eight 2-byte stores into the V cache, whose
//              layout value_cache[blk][h][d][off] makes a token's elements B*2 bytes apart by
//              construction.  Consecutive lanes take consecutive dim groups, so the K/V reads are
//              fully coalesced; the grid is flat over (token, chunk).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "vllmini_b200.h"

namespace vb200 {

// x == 8, head_size % 8 == 0, 16-byte aligned rows
__global__ void __launch_bounds__(256) reshape_and_cache_vec8_kernel(
    const __half* __restrict__ key, const __half* __restrict__ value, __half* __restrict__ key_cache,
    __half* __restrict__ value_cache, const int64_t* __restrict__ slot_mapping, int64_t total_chunks,
    int chunks_per_token, int num_heads, int head_size, int block_size, int64_t key_stride, int64_t value_stride) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total_chunks) return;
  const int64_t t = idx / chunks_per_token;
  const int c = (int)(idx - t * chunks_per_token);
  const int64_t slot = slot_mapping[t];
  if (slot < 0) return;  // padding token (cache_kernels.cu:166-169)
  const int64_t blk = slot / block_size;
  const int off = (int)(slot % block_size);
  const int groups = head_size >> 3;
  const int h = c / groups;
  const int g = c - h * groups;

  const uint4 kv = ldg128(key + t * key_stride + (int64_t)c * 8);
  const uint4 vv = ldg128(value + t * value_stride + (int64_t)c * 8);

  // key_cache[blk][h][g][off][0..7]
  __half* kdst = key_cache + (((blk * num_heads + h) * groups + g) * block_size + off) * 8;
  *reinterpret_cast<uint4*>(kdst) = kv;

  // value_cache[blk][h][g*8 + i][off]
  __half* vdst = value_cache + ((blk * num_heads + h) * (int64_t)head_size + g * 8) * block_size + off;
  const uint32_t w[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    reinterpret_cast<uint16_t*>(vdst)[(2 * i) * block_size] = (uint16_t)(w[i] & 0xffffu);
    reinterpret_cast<uint16_t*>(vdst)[(2 * i + 1) * block_size] = (uint16_t)(w[i] >> 16);
  }
}

// any x / alignment: one thread per element, literal index formulas of cache_kernels.cu:178-199
__global__ void __launch_bounds__(256) reshape_and_cache_generic_kernel(
    const uint16_t* __restrict__ key, const uint16_t* __restrict__ value, uint16_t* __restrict__ key_cache,
    uint16_t* __restrict__ value_cache, const int64_t* __restrict__ slot_mapping, int64_t total, int n,
    int num_heads, int head_size, int block_size, int x, int64_t key_stride, int64_t value_stride) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t t = idx / n;
  const int i = (int)(idx - t * n);
  const int64_t slot = slot_mapping[t];
  if (slot < 0) return;
  const int64_t blk = slot / block_size;
  const int64_t off = slot % block_size;
  const int h = i / head_size;
  const int d = i - h * head_size;
  const int64_t tk = ((blk * num_heads + h) * (head_size / x) + d / x) * block_size * x + off * x + d % x;
  const int64_t tv = ((blk * num_heads + h) * (int64_t)head_size + d) * block_size + off;
  key_cache[tk] = key[t * key_stride + i];
  value_cache[tv] = value[t * value_stride + i];
}

}  // namespace vb200

using namespace vb200;

extern "C" int vb200_reshape_and_cache_f16(const void* key, const void* value, void* key_cache, void* value_cache,
                                           const int64_t* slot_mapping, int32_t num_tokens, int32_t num_heads,
                                           int32_t head_size, int32_t block_size, int32_t x, int64_t key_stride,
                                           int64_t value_stride, int32_t device, void* stream) {
  if (num_tokens < 0 || num_heads <= 0 || head_size <= 0 || block_size <= 0 || x <= 0)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "negative or zero size argument");
  if (head_size % x != 0)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "head_size %d not a multiple of x %d", head_size, x);
  if (num_tokens == 0) return VB200_OK;
  if (!key || !value || !key_cache || !value_cache || !slot_mapping)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "null pointer argument");
  DeviceGuard guard(device);
  if (guard.status != cudaSuccess)
    return set_error(VB200_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.status));
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  const bool vec_ok = x == 8 && head_size % 8 == 0 && key_stride % 8 == 0 && value_stride % 8 == 0 &&
                      ((((uintptr_t)key | (uintptr_t)value | (uintptr_t)key_cache) & 15) == 0);
  if (vec_ok) {
    const int chunks = num_heads * head_size / 8;
    const int64_t total = (int64_t)num_tokens * chunks;
    const int threads = 256;
    const int64_t blocks = (total + threads - 1) / threads;
    reshape_and_cache_vec8_kernel<<<(unsigned)blocks, threads, 0, s>>>(
        static_cast<const __half*>(key), static_cast<const __half*>(value), static_cast<__half*>(key_cache),
        static_cast<__half*>(value_cache), slot_mapping, total, chunks, num_heads, head_size, block_size,
        key_stride, value_stride);
  } else {
    const int n = num_heads * head_size;
    const int64_t total = (int64_t)num_tokens * n;
    const int threads = 256;
    const int64_t blocks = (total + threads - 1) / threads;
    reshape_and_cache_generic_kernel<<<(unsigned)blocks, threads, 0, s>>>(
        static_cast<const uint16_t*>(key), static_cast<const uint16_t*>(value), static_cast<uint16_t*>(key_cache),
        static_cast<uint16_t*>(value_cache), slot_mapping, total, n, num_heads, head_size, block_size, x,
        key_stride, value_stride);
  }
  count_launch();
  VB200_CUDA_CHECK(cudaGetLastError());
  return VB200_OK;
}
