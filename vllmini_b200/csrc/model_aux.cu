// Residual add + LayerNorm in one pass, for the batched decode driver (vllmini_b200/decode.py).
//
// Between two hot-path launches a GPT-2 block does  x = x + proj(...) ; h = LayerNorm(x)  (reference:
// vllmini/model/gpt2.py:150-160, nn.LayerNorm + residual adds, each its own elementwise kernel in eager torch).
// At decode batch sizes these are launch-latency bound (393 KB per tensor at 256 x 768), so the pair is fused:
// one warp per row, the row lives in registers, two-pass mean / variance, one read of each input and one
// write of each output.  With row index arrays the same kernel does the embedding step
// x = wte[ids] + wpe[pos] ; h = ln_1(x)  (gpt2.py:258-262).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "vllmini_b200.h"

namespace vb200 {

constexpr int kLnWarps = 2;   // rows per CTA: small CTAs spread the rows of a decode batch over all SMs

template <int VPL>  // 16-byte chunks per lane: hidden <= VPL * 256
__global__ void __launch_bounds__(kLnWarps * 32)
    add_layernorm_kernel(const __half* a, const int64_t* __restrict__ a_rows, int64_t a_stride, const __half* x,
                         const int64_t* __restrict__ x_rows, int64_t x_stride, __half* x_out, __half* ln_out,
                         const __half* __restrict__ gamma, const __half* __restrict__ beta, int num_rows, int hidden,
                         float eps, int64_t a_num_rows, int64_t x_num_rows) {
  griddep_launch_dependents();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= num_rows) return;
  const int nchunks = hidden >> 3;
  const int64_t ar = a_rows ? a_rows[row] : row;
  const int64_t xr = x_rows ? x_rows[row] : row;
  // an index outside its source table (token id >= vocab, position >= n_positions) fails loudly, like the device
  // assert of the eager embedding lookup it replaces -- never a silent out-of-bounds read
  if (ar < 0 || ar >= a_num_rows || xr < 0 || xr >= x_num_rows) __trap();
  const uint4* ap = reinterpret_cast<const uint4*>(a + ar * a_stride);
  const uint4* xp = reinterpret_cast<const uint4*>(x + xr * x_stride);  // may alias x_out: plain loads

  // all loads of the row are independent: issue them together (one memory round trip per launch)
  uint4 av_[VPL], xv_[VPL], gv_[VPL], bv_[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = lane + 32 * v;
    if (c < nchunks) {
      av_[v] = ap[c];
      xv_[v] = xp[c];
      gv_[v] = ldg128(gamma + c * 8);
      bv_[v] = ldg128(beta + c * 8);
    }
  }
  float s[VPL][8];
  float sum = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = lane + 32 * v;
    if (c < nchunks) {
      const uint4 av = av_[v];
      const uint4 xv = xv_[v];
      const __half2* ah = reinterpret_cast<const __half2*>(&av);
      const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 af = __half22float2(ah[i]);
        const float2 xf = __half22float2(xh[i]);
        // the residual stream is fp16: round the sum first, normalise what was stored
        s[v][2 * i] = __half2float(__float2half_rn(af.x + xf.x));
        s[v][2 * i + 1] = __half2float(__float2half_rn(af.y + xf.y));
        sum += s[v][2 * i] + s[v][2 * i + 1];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) s[v][i] = 0.f;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)hidden;
  float sq = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    if (lane + 32 * v < nchunks) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = s[v][i] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)hidden + eps);

  uint4* xo = reinterpret_cast<uint4*>(x_out + (int64_t)row * hidden);
  uint4* lo = reinterpret_cast<uint4*>(ln_out + (int64_t)row * hidden);
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = lane + 32 * v;
    if (c < nchunks) {
      const uint4 gv = gv_[v];
      const uint4 bv = bv_[v];
      const __half2* gh = reinterpret_cast<const __half2*>(&gv);
      const __half2* bh = reinterpret_cast<const __half2*>(&bv);
      uint4 xs, ys;
      __half2* xsh = reinterpret_cast<__half2*>(&xs);
      __half2* ysh = reinterpret_cast<__half2*>(&ys);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 g = __half22float2(gh[i]);
        const float2 b = __half22float2(bh[i]);
        xsh[i] = __floats2half2_rn(s[v][2 * i], s[v][2 * i + 1]);
        ysh[i] = __floats2half2_rn((s[v][2 * i] - mean) * rstd * g.x + b.x,
                                   (s[v][2 * i + 1] - mean) * rstd * g.y + b.y);
      }
      xo[c] = xs;
      lo[c] = ys;
    }
  }
}

// Greedy sampling: row-wise argmax of the logits (first index among equal maxima, NaN counts as the maximum --
// torch.argmax's behaviour, which vllmini's sampler reduces to at temperature 0).  One CTA per row.
constexpr int kArgmaxThreads = 256;

// (value, index) order: NaN above everything, then larger value, then smaller index
__device__ __forceinline__ bool argmax_better(float v, int i, float bv, int bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn != bn) return vn;
  if (!vn && v != bv) return v > bv;
  return i < bi;
}

__global__ void __launch_bounds__(kArgmaxThreads)
    argmax_rows_kernel(const __half* __restrict__ x, int64_t stride, int n, int64_t* __restrict__ out) {
  griddep_launch_dependents();
  griddep_wait();
  const __half* row = x + (int64_t)blockIdx.x * stride;
  float best = -INFINITY;   // largest non-NaN value seen by this thread (indices only grow: strict > keeps the first)
  int besti = 0x7fffffff;
  int nani = 0x7fffffff;    // first NaN seen by this thread
  auto scan8 = [&](const uint4& v, int base) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
    // chunk maximum first (NaN-propagating): almost every chunk is rejected with one compare
    const __half2 m2 = __hmax2_nan(__hmax2_nan(h[0], h[1]), __hmax2_nan(h[2], h[3]));
    const float m = __half2float(__hmax_nan(__low2half(m2), __high2half(m2)));
    if (!(m <= best)) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float f = __half2float(reinterpret_cast<const __half*>(&v)[i]);
        if (f != f) {
          nani = min(nani, base + i);
        } else if (f > best) {
          best = f;
          besti = base + i;
        }
      }
    }
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  const int nvec = vec ? (n >> 3) : 0;
  int c = threadIdx.x;
  for (; c + 3 * kArgmaxThreads < nvec; c += 4 * kArgmaxThreads) {  // four independent 16-byte loads in flight
    const uint4 v0 = ldg128(row + (c + 0 * kArgmaxThreads) * 8);
    const uint4 v1 = ldg128(row + (c + 1 * kArgmaxThreads) * 8);
    const uint4 v2 = ldg128(row + (c + 2 * kArgmaxThreads) * 8);
    const uint4 v3 = ldg128(row + (c + 3 * kArgmaxThreads) * 8);
    scan8(v0, (c + 0 * kArgmaxThreads) * 8);
    scan8(v1, (c + 1 * kArgmaxThreads) * 8);
    scan8(v2, (c + 2 * kArgmaxThreads) * 8);
    scan8(v3, (c + 3 * kArgmaxThreads) * 8);
  }
  for (; c < nvec; c += kArgmaxThreads) scan8(ldg128(row + c * 8), c * 8);
  for (int i = nvec * 8 + threadIdx.x; i < n; i += kArgmaxThreads) {
    const float f = __half2float(row[i]);
    if (f != f) {
      nani = min(nani, i);
    } else if (f > best) {
      best = f;
      besti = i;
    }
  }
  if (nani != 0x7fffffff) {
    best = __int_as_float(0x7fc00000);
    besti = nani;
  } else if (besti == 0x7fffffff) {   // nothing above -inf in this thread's share: any -inf element qualifies
    besti = threadIdx.x < n ? ((vec && threadIdx.x < nvec) ? threadIdx.x * 8 : threadIdx.x) : 0x7fffffff;
    if (besti >= n) besti = 0x7fffffff;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (argmax_better(ov, oi, best, besti)) best = ov, besti = oi;
  }
  __shared__ float sv[kArgmaxThreads / 32];
  __shared__ int si[kArgmaxThreads / 32];
  if ((threadIdx.x & 31) == 0) sv[threadIdx.x >> 5] = best, si[threadIdx.x >> 5] = besti;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < kArgmaxThreads / 32; ++w)
      if (argmax_better(sv[w], si[w], best, besti)) best = sv[w], besti = si[w];
    out[blockIdx.x] = besti == 0x7fffffff ? 0 : besti;
  }
}

}  // namespace vb200

using namespace vb200;

extern "C" int vb200_argmax_rows_f16(int64_t* out, const void* x, int64_t row_stride, int32_t num_rows, int32_t n,
                                     int32_t device, void* stream) {
  if (num_rows < 0 || n <= 0) return set_error(VB200_ERR_INVALID_ARGUMENT, "negative or zero size argument");
  if (num_rows == 0) return VB200_OK;
  if (!out || !x) return set_error(VB200_ERR_INVALID_ARGUMENT, "null pointer argument");
  DeviceGuard guard(device);
  if (guard.status != cudaSuccess)
    return set_error(VB200_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.status));
  VB200_CUDA_CHECK(launch_kernel(argmax_rows_kernel, dim3(num_rows), dim3(kArgmaxThreads), 0,
                                 static_cast<cudaStream_t>(stream), static_cast<const __half*>(x), row_stride, (int)n,
                                 out));
  count_launch();
  return VB200_OK;
}

extern "C" int vb200_add_layernorm_f16(void* x_out, void* ln_out, const void* a, const int64_t* a_rows,
                                       int64_t a_num_rows, int64_t a_stride, const void* x, const int64_t* x_rows,
                                       int64_t x_num_rows, int64_t x_stride, const void* gamma, const void* beta,
                                       int32_t num_rows, int32_t hidden, float eps, int32_t device, void* stream) {
  if (num_rows < 0 || hidden <= 0) return set_error(VB200_ERR_INVALID_ARGUMENT, "negative or zero size argument");
  if ((!a_rows && a_num_rows < num_rows) || (!x_rows && x_num_rows < num_rows))
    return set_error(VB200_ERR_INVALID_ARGUMENT, "fewer source rows than num_rows and no row index array");
  if (num_rows == 0) return VB200_OK;
  if (!x_out || !ln_out || !a || !x || !gamma || !beta)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "null pointer argument");
  if (hidden % 8 != 0 || hidden > 8 * 256)
    return set_error(VB200_ERR_UNSUPPORTED, "hidden size %d: must be a multiple of 8, at most 2048", hidden);
  if ((((uintptr_t)x_out | (uintptr_t)ln_out | (uintptr_t)a | (uintptr_t)x | (uintptr_t)gamma | (uintptr_t)beta) & 15) ||
      (a_stride % 8) || (x_stride % 8))
    return set_error(VB200_ERR_INVALID_ARGUMENT, "pointers and row strides must be 16-byte aligned");
  DeviceGuard guard(device);
  if (guard.status != cudaSuccess)
    return set_error(VB200_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.status));
  const dim3 grid((num_rows + kLnWarps - 1) / kLnWarps), block(kLnWarps * 32);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __half* ah = static_cast<const __half*>(a);
  const __half* xh = static_cast<const __half*>(x);
  __half* xo = static_cast<__half*>(x_out);
  __half* lo = static_cast<__half*>(ln_out);
  const __half* g = static_cast<const __half*>(gamma);
  const __half* b = static_cast<const __half*>(beta);
  const int vpl = (hidden / 8 + 31) / 32;
#define VB200_LN_CASE(V)                                                                                            \
  case V:                                                                                                           \
    VB200_CUDA_CHECK(launch_kernel(add_layernorm_kernel<V>, grid, block, 0, s, ah, a_rows, a_stride, xh, x_rows,    \
                                   x_stride, xo, lo, g, b, (int)num_rows, (int)hidden, eps, a_num_rows,           \
                                   x_num_rows));                                                                    \
    break;
  switch (vpl) {
    VB200_LN_CASE(1)
    VB200_LN_CASE(2)
    VB200_LN_CASE(3)
    VB200_LN_CASE(4)
    VB200_LN_CASE(5)
    VB200_LN_CASE(6)
    VB200_LN_CASE(7)
    default:
      VB200_CUDA_CHECK(launch_kernel(add_layernorm_kernel<8>, grid, block, 0, s, ah, a_rows, a_stride, xh, x_rows,
                                     x_stride, xo, lo, g, b, (int)num_rows, (int)hidden, eps, a_num_rows, x_num_rows));
  }
#undef VB200_LN_CASE
  count_launch();
  return VB200_OK;
}
