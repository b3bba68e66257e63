// reshape_and_cache for sm_100a: scatter new K/V rows into the paged caches by slot.
//
// Behavioural spec: /root/reference/paged_attention_ext/paged_attention_cuda/cache_kernels.cu
//   :152-207 (kernel), :256-281 (host entry).  Pure copy ("auto" kv-cache dtype) -> bit-exact.
//
//   reference: one CTA per token, one thread per scalar: 2-byte loads, 2-byte stores; consecutive threads
//              write consecutive head dims, i.e. addresses B*2 bytes apart -> one sector per element.
//   here:      one CTA per (32-token tile, head): coalesced 128-bit row loads staged in shared memory,
//              then lane <-> token stores, so that tokens in consecutive slots (prefill) are written as
//              whole sectors (see reshape_and_cache_tile_kernel).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "vllmini_b200.h"

namespace vb200 {

// x == 8, head_size % 8 == 0, 16-byte aligned rows.
// One CTA per (16/32-token tile, head): 16 tokens on 8 warps for 512 < T <= 2048, 32 tokens on 4 warps above.
// The tile's slots are read first and classified (warp ballot):
//  * consecutive slots (prefill).  Phase 1: the tile's K and V rows (D halves each) are read with coalesced
//    128-bit loads and parked in shared memory (row pitch D*2+16 bytes: conflict-free for both access patterns).
//    Phase 2: lane <-> token.  K: lane stores the 16-byte group g of its token to key_cache[blk][h][g][off][0..7];
//    V: lane stores element d of its token to value_cache[blk][h][d][off]: the lanes of a store instruction hit
//    consecutive addresses and the LSU coalesces them into whole sectors.
//  * scattered slots (a large decode batch): every store is its own sector whatever the mapping -- that is what
//    the layout dictates -- so the rows go straight from registers to the caches, warp <-> token, without the
//    staging pass and its CTA barrier (the staged fallback of round 1 lost to the reference kernel here, 14.2 vs
//    13.2 us at 2048 tokens).

template <int kTileTokens, int kTileThreads, bool kEager = false>   // tokens: 32, or 16 for small batches
__global__ void __launch_bounds__(kTileThreads) reshape_and_cache_tile_kernel(
    const __half* __restrict__ key, const __half* __restrict__ value, __half* __restrict__ key_cache,
    __half* __restrict__ value_cache, const int64_t* __restrict__ slot_mapping, int num_tokens, int num_heads,
    int head_size, int block_size, int64_t key_stride, int64_t value_stride) {
  extern __shared__ __align__(16) uint8_t tile_smem[];
  // head is the fast grid dimension: the CTAs that fill the heads of the same cache blocks run together,
  // so the partial-sector V writes of a block reach DRAM close in time
  const int h = blockIdx.x;
  const int t0 = blockIdx.y * kTileTokens;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  griddep_launch_dependents();  // PDL: see common.cuh
  griddep_wait();
  const int groups = head_size >> 3;        // 16-byte chunks per row
  const int pitch = head_size * 2 + 16;     // K rows: 16-byte aligned, conflict-free for 128-bit column reads
  const int vpitch = head_size / 2 + 1;     // V rows in 32-bit words, odd pitch: conflict-free column reads
  uint8_t* ks = tile_smem;
  uint32_t* vs = reinterpret_cast<uint32_t*>(tile_smem + kTileTokens * pitch);

  // slot of this lane's token (every warp holds the same 32 slots)
  const int my_tok = t0 + lane;
  const int64_t slot = (lane < kTileTokens && my_tok < num_tokens) ? slot_mapping[my_tok] : -1;

  // rows of this warp's tokens, warp <-> token, lanes <-> 32-bit words (the scattered path's mapping)
  constexpr int kTok = kTileTokens / (kTileThreads / 32);   // tokens per warp
  constexpr int kMaxWords = 4;                              // head_size <= 256
  uint32_t kw[kTok][kMaxWords], vw[kTok][kMaxWords];
  auto load_rows = [&]() {
#pragma unroll
    for (int n = 0; n < kTok; ++n) {
      const int tok = warp + n * (kTileThreads / 32);
      const bool live = t0 + tok < num_tokens;
      const uint32_t* krow = reinterpret_cast<const uint32_t*>(key + (int64_t)(t0 + tok) * key_stride + (int64_t)h * head_size);
      const uint32_t* vrow = reinterpret_cast<const uint32_t*>(value + (int64_t)(t0 + tok) * value_stride + (int64_t)h * head_size);
#pragma unroll
      for (int u = 0; u < kMaxWords; ++u) {
        const int i = lane + 32 * u;
        if (live && i < head_size / 2) {
          kw[n][u] = __ldg(krow + i);
          vw[n][u] = __ldg(vrow + i);
        }
      }
    }
  };
  if constexpr (kEager) load_rows();   // in flight together with the slot load: one round trip for either path

  // Are the tile's tokens (mostly) in consecutive slots?  Every warp holds the same slots -> same answer.
  const int64_t prev = __shfl_up_sync(0xffffffffu, slot, 1);
  const unsigned adj = __ballot_sync(0xffffffffu, lane > 0 && slot >= 0 && slot == prev + 1);
  const bool run_like = __popc(adj) >= kTileTokens / 2;

  if (!run_like) {
    // ---- decode-like (scattered slots): no staging, no CTA barrier.  warp <-> token, lanes <-> 32-bit words of the
    //      row, so the stores of one instruction stay inside one (block, head) tile; the rows of all of this warp's
    //      tokens are loaded before the first store. ----
    if constexpr (!kEager) load_rows();
#pragma unroll
    for (int n = 0; n < kTok; ++n) {
      const int tok = warp + n * (kTileThreads / 32);
      const int64_t s_tok = __shfl_sync(0xffffffffu, slot, tok);
      if (s_tok < 0) continue;   // padding token (cache_kernels.cu:166-169); also tokens past num_tokens
      const int64_t blk = s_tok / block_size;
      const int off = (int)(s_tok % block_size);
      uint32_t* kdst32 = reinterpret_cast<uint32_t*>(key_cache + (blk * num_heads + h) * (int64_t)head_size * block_size);
      uint16_t* vdst =
          reinterpret_cast<uint16_t*>(value_cache) + (blk * num_heads + h) * (int64_t)head_size * block_size + off;
#pragma unroll
      for (int u = 0; u < kMaxWords; ++u) {
        const int i = lane + 32 * u;   // 32-bit word i of the row: dims 2i, 2i+1
        if (i < head_size / 2) {
          kdst32[((int64_t)(i >> 2) * block_size + off) * 4 + (i & 3)] = kw[n][u];
          vdst[(int64_t)(2 * i) * block_size] = (uint16_t)(vw[n][u] & 0xffffu);
          vdst[(int64_t)(2 * i + 1) * block_size] = (uint16_t)(vw[n][u] >> 16);
        }
      }
    }
    return;
  }

  if constexpr (kEager) {
    // the rows are already in registers: park them (32-bit words) instead of re-reading them with 128-bit loads
#pragma unroll
    for (int n = 0; n < kTok; ++n) {
      const int tok = warp + n * (kTileThreads / 32);
#pragma unroll
      for (int u = 0; u < kMaxWords; ++u) {
        const int i = lane + 32 * u;
        if (t0 + tok < num_tokens && i < head_size / 2) {
          reinterpret_cast<uint32_t*>(ks + tok * pitch)[i] = kw[n][u];
          vs[tok * vpitch + i] = vw[n][u];
        }
      }
    }
  } else {
  // ---- prefill-like.  Phase 1: coalesced 128-bit row loads -> smem.  All loads of a batch are issued before the
  //      first store so a thread pays one memory round trip per batch. ----
  const int nchunks = kTileTokens * groups;
  constexpr int kBatch = 4;
  for (int c0 = threadIdx.x; c0 < nchunks; c0 += kTileThreads * kBatch) {
    uint4 kv[kBatch], vv[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int c = c0 + u * kTileThreads;
      const int tok = c / groups, g = c - tok * groups;
      if (c < nchunks && t0 + tok < num_tokens) {
        kv[u] = ldg128(key + (int64_t)(t0 + tok) * key_stride + (int64_t)h * head_size + g * 8);
        vv[u] = ldg128(value + (int64_t)(t0 + tok) * value_stride + (int64_t)h * head_size + g * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int c = c0 + u * kTileThreads;
      const int tok = c / groups, g = c - tok * groups;
      if (c < nchunks && t0 + tok < num_tokens) {
        *reinterpret_cast<uint4*>(ks + tok * pitch + g * 16) = kv[u];
        uint32_t* vr = vs + tok * vpitch + g * 4;
        vr[0] = vv[u].x;
        vr[1] = vv[u].y;
        vr[2] = vv[u].z;
        vr[3] = vv[u].w;
      }
    }
  }
  }
  __syncthreads();

  // ---- phase 2: lane <-> token, so one store instruction covers consecutive slots ----
  if (slot < 0) return;  // padding token (cache_kernels.cu:166-169); also lanes past num_tokens
  const int64_t blk = slot / block_size;
  const int off = (int)(slot % block_size);
  __half* kdst = key_cache + ((blk * num_heads + h) * (int64_t)groups * block_size + off) * 8;
  for (int g = warp; g < groups; g += kTileThreads / 32)
    *reinterpret_cast<uint4*>(kdst + (int64_t)g * block_size * 8) =
        *reinterpret_cast<const uint4*>(ks + lane * pitch + g * 16);
  uint16_t* vdst =
      reinterpret_cast<uint16_t*>(value_cache) + (blk * num_heads + h) * (int64_t)head_size * block_size + off;
  const uint32_t* vrow = vs + lane * vpitch;
  for (int d2 = warp; d2 < head_size / 2; d2 += kTileThreads / 32) {
    const uint32_t w = vrow[d2];
    vdst[(int64_t)(2 * d2) * block_size] = (uint16_t)(w & 0xffffu);
    vdst[(int64_t)(2 * d2 + 1) * block_size] = (uint16_t)(w >> 16);
  }
}

// Small batches (decode): one warp per (token, head), no shared memory, no CTA barrier.  The slot and the
// row words are loaded independently (one memory round trip), lanes <-> 32-bit words of the 2*D-byte row.
__global__ void __launch_bounds__(128) reshape_and_cache_warp_kernel(
    const __half* __restrict__ key, const __half* __restrict__ value, __half* __restrict__ key_cache,
    __half* __restrict__ value_cache, const int64_t* __restrict__ slot_mapping, int num_tokens, int num_heads,
    int head_size, int block_size, int64_t key_stride, int64_t value_stride) {
  const int lane = threadIdx.x & 31;
  griddep_launch_dependents();  // PDL: see common.cuh
  griddep_wait();
  const int64_t w = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int64_t t = w / num_heads;
  const int h = (int)(w - t * num_heads);
  if (t >= num_tokens) return;
  const int64_t slot = slot_mapping[t];
  const uint32_t* krow = reinterpret_cast<const uint32_t*>(key + t * key_stride + (int64_t)h * head_size);
  const uint32_t* vrow = reinterpret_cast<const uint32_t*>(value + t * value_stride + (int64_t)h * head_size);
  constexpr int kMaxWords = 4;  // head_size <= 256
  uint32_t kw[kMaxWords], vw[kMaxWords];
#pragma unroll
  for (int u = 0; u < kMaxWords; ++u) {
    const int i = lane + 32 * u;
    if (i < head_size / 2) {
      kw[u] = __ldg(krow + i);
      vw[u] = __ldg(vrow + i);
    }
  }
  if (slot < 0) return;  // padding token (cache_kernels.cu:166-169)
  const int64_t blk = slot / block_size;
  const int off = (int)(slot % block_size);
  uint32_t* kdst32 = reinterpret_cast<uint32_t*>(key_cache + (blk * num_heads + h) * (int64_t)head_size * block_size);
  uint16_t* vdst =
      reinterpret_cast<uint16_t*>(value_cache) + (blk * num_heads + h) * (int64_t)head_size * block_size + off;
#pragma unroll
  for (int u = 0; u < kMaxWords; ++u) {
    const int i = lane + 32 * u;  // 32-bit word i of the row: dims 2i, 2i+1
    if (i < head_size / 2) {
      kdst32[((int64_t)(i >> 2) * block_size + off) * 4 + (i & 3)] = kw[u];
      vdst[(int64_t)(2 * i) * block_size] = (uint16_t)(vw[u] & 0xffffu);
      vdst[(int64_t)(2 * i + 1) * block_size] = (uint16_t)(vw[u] >> 16);
    }
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

#ifdef VB200_EXPERIMENTS
// sweep knob of the experiments build (tools/sweep_cache.py): 0 = shipped dispatch, 1 = warp kernel for any size,
// 2 = 16-token tiles on 4 warps, 3 = 32-token tiles, 6 = 16-token tiles on 8 warps (shipped for 512 < T <= 2048),
// 7 = 32-token tiles on 8 warps, 8 / 9 = eager row loads
static int g_cache_mode = 0;
extern "C" int32_t vb200_debug_set_cache_kernel(int32_t mode) {
  const int prev = g_cache_mode;
  g_cache_mode = mode;
  return prev;
}
static int cache_mode() { return g_cache_mode; }
#else
static constexpr int cache_mode() { return 0; }
#endif

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
  if (vec_ok && head_size <= 256 && (num_tokens + 15) / 16 <= 65535) {
    const int mode = cache_mode();
    if (mode == 1 || (mode == 0 && num_tokens <= 512)) {  // decode-sized: latency-bound, one warp per (token, head)
      const int64_t warps = (int64_t)num_tokens * num_heads;
      VB200_CUDA_CHECK(launch_kernel(reshape_and_cache_warp_kernel, dim3((unsigned)((warps + 3) / 4)), dim3(128), 0, s,
                                     static_cast<const __half*>(key), static_cast<const __half*>(value),
                                     static_cast<__half*>(key_cache), static_cast<__half*>(value_cache), slot_mapping,
                                     num_tokens, num_heads, head_size, block_size, key_stride, value_stride));
      count_launch();
      return VB200_OK;
    }
    // 512 < T <= 2048: 16-token tiles on 8 warps (2 tokens per warp) -- scattered batches of that size then run at the
    // reference kernel's speed (L2 sector-write bound, 0.99x) instead of 0.76x with 4 warps, consecutive slots still
    // 2x the reference; above that 32-token tiles on 4 warps (profiles/r02g_reshape_and_cache.md)
    const int tt = (mode == 2 || mode == 6 || mode == 8) ? 16
                   : (mode == 3 || mode == 7 || mode == 9) ? 32 : (num_tokens <= 2048 ? 16 : 32);
    const int smem = tt * (head_size * 2 + 16) + tt * (head_size / 2 + 1) * 4;
    dim3 grid(num_heads, (num_tokens + tt - 1) / tt);
    auto kern = tt == 16 ? reshape_and_cache_tile_kernel<16, 256> : reshape_and_cache_tile_kernel<32, 128>;
    int threads = tt == 16 ? 256 : 128;
#ifdef VB200_EXPERIMENTS
    if (mode == 2) {                // round 2's first dispatch: 16-token tiles on 4 warps
      kern = reshape_and_cache_tile_kernel<16, 128>;
      threads = 128;
    }
    if (mode == 7) {                // tried: twice the warps per 32-token tile
      kern = reshape_and_cache_tile_kernel<32, 256>;
      threads = 256;
    }
    if (mode == 8 || mode == 9) {   // tried: row loads issued before the slots are known, registers -> smem
      kern = tt == 16 ? reshape_and_cache_tile_kernel<16, 128, true> : reshape_and_cache_tile_kernel<32, 128, true>;
      threads = 128;
    }
#endif
    VB200_CUDA_CHECK(launch_kernel(kern, grid, dim3(threads), (size_t)smem, s, static_cast<const __half*>(key),
                                   static_cast<const __half*>(value), static_cast<__half*>(key_cache),
                                   static_cast<__half*>(value_cache), slot_mapping, num_tokens, num_heads, head_size,
                                   block_size, key_stride, value_stride));
    count_launch();
    return VB200_OK;
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
