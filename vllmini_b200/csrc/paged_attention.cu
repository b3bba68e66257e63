// paged_attention_v1 for sm_100a -- per-warp bulk-copy (TMA 1-D) pipelines, online softmax,
// FHFMA (fp16 x fp16 + fp32) inner products, fused split-KV merge.
//
// Behavioural spec: /root/reference/paged_attention_ext/paged_attention_cuda/attention_kernels.cu
//   :86-496 (kernel body), :693-767 (launcher), :805-826 (host entry).  Written from that spec with
// a different decomposition:
//
//   reference                                   this kernel
//   ---------                                   -----------
//   grid (H, S), 4 warps, warp <- block          grid (H, S, splits), W warps, warp <- 32-token stage
//   K: 8-byte LDG.CONSTANT per lane, dependent   K/V tiles of a stage fetched by cp.async.bulk
//      on a per-block table load                    (UBLKCP) into a per-warp smem ring, NST stages
//   full-sequence fp32 logits in smem,           deep; table entries prefetched one refill ahead;
//      3 CTA barriers, K pass then V pass        streaming (online) softmax: O(1) smem in ctx,
//   fp16 p, fp16 PV partial sums                 no CTA barrier in the main loop
//   no split-KV in v1                            fp16 p, exact fp16xfp16 products, fp32 accumulation
//                                                (FHFMA = fma.rn.f32.f16, sm_100+)
//                                                split-KV with an in-kernel last-arriver merge
//
// Layout facts used (vllmini/kv_cache.py:13-14): for one (physical block,
// kv head) the K tile [D/8][B][8] and the V tile [D][B] are each D*B*2 contiguous bytes, so one
// bulk copy moves a whole tile.  A lane reads 16-byte chunks c = lane + 32*j of a tile:
//   K chunk c -> token c % B, dims 8*(c / B) .. +7        (lane owns ONE token, D*B/256 dim groups)
//   V chunk c -> row   c / (B/8), tokens 8*(c % (B/8)) .. +7  (lane owns D*B/256 rows, 8 tokens)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "vllmini_b200.h"

namespace vb200 {

struct PaParams {
  __half* __restrict__ out;
  const __half* __restrict__ q;
  const __half* __restrict__ kc;
  const __half* __restrict__ vc;
  const int32_t* __restrict__ block_tables;
  const int32_t* __restrict__ seq_lens;
  int num_heads;
  int num_kv_heads;
  int max_blocks_per_seq;
  int max_seq_len;
  int64_t q_stride;
  int64_t kv_block_stride;
  int64_t kv_head_stride;
  float scale_log2e;
  int num_splits;
  int split_tokens;  // multiple of 32; ignored when num_splits == 1
  int32_t* ws_cnt;   // [num_seqs * num_heads] zero between launches
  float* ws_ml;      // [num_seqs * num_heads * num_splits * 2]
  float* ws_o;       // [num_seqs * num_heads * num_splits * D]
};

template <int D, int B, int W, int NST>
struct PaCfg {
  static constexpr int BPS = 32 / B;             // KV blocks per 32-token stage
  static constexpr int TILE = D * B * 2;         // bytes of one (block, head) K tile (== V tile)
  static constexpr int NJ = D * B / 256;         // 16-byte chunks per lane per tile
  static constexpr int TPR = B / 8;              // 16-byte chunks per V row
  static constexpr int STAGE = 2 * BPS * TILE;   // bytes per stage: K tiles then V tiles (= 128*D)
  static constexpr int RING = W * NST * STAGE;
  static constexpr int PBUF_OFF = RING;                     // W * 32 halves
  static constexpr int BAR_OFF = PBUF_OFF + W * 64;         // W * NST mbarriers
  static constexpr int RED_OFF = BAR_OFF + W * NST * 8;     // W * (D + 2) floats
  static constexpr int SMEM = RED_OFF + W * (D + 2) * 4;
  static_assert(B == 8 || B == 16 || B == 32, "block size");
  static_assert(D % 8 == 0 && NJ >= 1 && (D * B) % 256 == 0, "head size");
};

// 8 fp16 x fp16 products accumulated in fp32 on two chains (FHFMA).
__device__ __forceinline__ void dot8(const uint4& a, const uint4& b, float& c0, float& c1) {
  c0 = fhfma_lo(a.x, b.x, c0);
  c1 = fhfma_hi(a.x, b.x, c1);
  c0 = fhfma_lo(a.y, b.y, c0);
  c1 = fhfma_hi(a.y, b.y, c1);
  c0 = fhfma_lo(a.z, b.z, c0);
  c1 = fhfma_hi(a.z, b.z, c1);
  c0 = fhfma_lo(a.w, b.w, c0);
  c1 = fhfma_hi(a.w, b.w, c1);
}

// keep the first nv (0..8) halves of a 16-byte chunk, zero the rest (NaN-safe tail, the
// reference zeroes out-of-context V elements the same way: attention_kernels.cu:420-430)
__device__ __forceinline__ uint4 keep_first(uint4 v, int nv) {
  auto word = [&](uint32_t w, int i) -> uint32_t {
    const uint32_t lo = (2 * i < nv) ? 0x0000ffffu : 0u;
    const uint32_t hi = (2 * i + 1 < nv) ? 0xffff0000u : 0u;
    return w & (lo | hi);
  };
  return make_uint4(word(v.x, 0), word(v.y, 1), word(v.z, 2), word(v.w, 3));
}

template <int D, int B, int W, int NST>
__global__ void __launch_bounds__(W * 32) pa_v1_kernel(const PaParams p) {
  using C = PaCfg<D, B, W, NST>;
  extern __shared__ __align__(128) uint8_t smem[];

  const int head = blockIdx.x;
  const int seq = blockIdx.y;
  const int split = blockIdx.z;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;

  int len = p.seq_lens[seq];
  len = len < p.max_blocks_per_seq * B ? len : p.max_blocks_per_seq * B;  // never read past the table row
  if (p.num_splits > 1 && len > p.num_splits * p.split_tokens) len = p.num_splits * p.split_tokens;
  const int tok_begin = p.num_splits > 1 ? split * p.split_tokens : 0;
  if (split > 0 && tok_begin >= len) return;  // this split holds no token (split 0 always runs)
  int tok_end = len;
  if (p.num_splits > 1 && tok_begin + p.split_tokens < len) tok_end = tok_begin + p.split_tokens;
  const int ntok = tok_end > tok_begin ? tok_end - tok_begin : 0;
  const int nstages = (ntok + 31) >> 5;
  const int nblk = (ntok + B - 1) / B;
  const int nsplit_seq = p.num_splits > 1 && len > 0 ? (len + p.split_tokens - 1) / p.split_tokens : 1;

  const int kv_head = head / (p.num_heads / p.num_kv_heads);
  const int32_t* table = p.block_tables + (int64_t)seq * p.max_blocks_per_seq + tok_begin / B;
  const __half* kbase = p.kc + (int64_t)kv_head * p.kv_head_stride;
  const __half* vbase = p.vc + (int64_t)kv_head * p.kv_head_stride;

  const uint32_t ring = smem_u32(smem) + warp * (NST * C::STAGE);
  const uint32_t pbuf = smem_u32(smem + C::PBUF_OFF) + warp * 64;
  const uint32_t bars = smem_u32(smem + C::BAR_OFF) + warp * (NST * 8);

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < W * NST; ++i) mbar_init(smem_u32(smem + C::BAR_OFF) + i * 8, 1);
    fence_mbar_init();
  }
  __syncthreads();

  // ---- producer side: lane i < 2*BPS issues copy i of a stage (K tiles 0..BPS-1, then V tiles) ----
  // phys: lane b < BPS holds the physical block id of block b of the stage being issued.
  auto issue_stage = [&](int st, int slot, int phys) {
    const int blk0 = st * C::BPS;
    const int nb = nblk - blk0 < C::BPS ? nblk - blk0 : C::BPS;  // blocks that exist in this stage
    const uint32_t bar = bars + slot * 8;
    if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(2 * nb * C::TILE));
    __syncwarp();
    const int b = lane % C::BPS;
    const int is_v = lane / C::BPS;
    const int phys_b = __shfl_sync(0xffffffffu, phys, b);
    if (lane < 2 * C::BPS && b < nb) {
      const __half* src = (is_v ? vbase : kbase) + (int64_t)phys_b * p.kv_block_stride;
      const uint32_t dst = ring + slot * C::STAGE + (is_v * C::BPS + b) * C::TILE;
      bulk_g2s(dst, src, C::TILE, bar);
    }
  };
  auto load_phys = [&](int st) -> int {  // table entries of stage st, one per lane < BPS
    const int blk = st * C::BPS + lane;
    return (lane < C::BPS && blk < nblk) ? __ldg(table + blk) : 0;
  };

  // prologue: fill the ring
  {
    int ph[NST];
#pragma unroll
    for (int k = 0; k < NST; ++k) ph[k] = (warp + k * W < nstages) ? load_phys(warp + k * W) : 0;
#pragma unroll
    for (int k = 0; k < NST; ++k)
      if (warp + k * W < nstages) issue_stage(warp + k * W, k, ph[k]);
  }
  int next_phys = (warp + NST * W < nstages) ? load_phys(warp + NST * W) : 0;

  // ---- q: dims 8*(lane/B + j*BPS) .. +7 of this head's row, packed fp16 ----
  uint4 qv[C::NJ];
  {
    const __half* qrow = p.q + (int64_t)seq * p.q_stride + (int64_t)head * D;
#pragma unroll
    for (int j = 0; j < C::NJ; ++j) qv[j] = ldg128(qrow + 8 * (lane / B + j * C::BPS));
  }

  float m = -INFINITY;  // running max (log2 domain), warp-uniform
  float lsum = 0.f;     // per-lane partial sum of p
  float acc[C::NJ];     // rows lane/TPR + j*(32/TPR), partial over this lane's 8-token groups
#pragma unroll
  for (int j = 0; j < C::NJ; ++j) acc[j] = 0.f;

  int k = 0;
  for (int st = warp; st < nstages; st += W, ++k) {
    const int slot = k % NST;
    const uint32_t sbase = ring + slot * C::STAGE;
    mbar_wait(bars + slot * 8, (k / NST) & 1);

    // ---- QK^T: lane owns token lane % B of every block of the stage ----
    float part[C::BPS];
#pragma unroll
    for (int b = 0; b < C::BPS; ++b) {
      float c0 = 0.f, c1 = 0.f;
#pragma unroll
      for (int j = 0; j < C::NJ; ++j) {
        const uint4 kv = lds128(sbase + b * C::TILE + (lane + 32 * j) * 16);
        dot8(qv[j], kv, c0, c1);
      }
      part[b] = c0 + c1;
    }
    // transpose-reduce: sum over the BPS lanes sharing lane % B; lane keeps block lane / B
    float s;
    if constexpr (C::BPS == 1) {
      s = part[0];
    } else if constexpr (C::BPS == 2) {
      const bool up = lane & 16;
      const float give = up ? part[0] : part[1];
      s = (up ? part[1] : part[0]) + __shfl_xor_sync(0xffffffffu, give, 16);
    } else {  // BPS == 4
      const bool up = lane & 16;
      const float g0 = up ? part[0] : part[2];
      const float g1 = up ? part[1] : part[3];
      const float k0 = (up ? part[2] : part[0]) + __shfl_xor_sync(0xffffffffu, g0, 16);
      const float k1 = (up ? part[3] : part[1]) + __shfl_xor_sync(0xffffffffu, g1, 16);
      const bool up2 = lane & 8;
      const float give = up2 ? k0 : k1;
      s = (up2 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, give, 8);
    }
    // lane now holds the logit of stage token `lane`
    const int stage_tok0 = tok_begin + st * 32;
    const bool tail = stage_tok0 + 32 > tok_end;  // warp-uniform
    s *= p.scale_log2e;
    if (stage_tok0 + lane >= tok_end) s = -INFINITY;  // select, so NaN garbage cannot leak

    // ---- online softmax (warp-uniform running max) ----
    const float smax = warp_max_f32(s);
    if (smax > m) {
      const float alpha = ex2(m - smax);  // m = -inf on first use -> 0
      m = smax;
      lsum *= alpha;
#pragma unroll
      for (int j = 0; j < C::NJ; ++j) acc[j] *= alpha;
    }
    const __half p16 = __float2half_rn(ex2(s - m));
    lsum += __half2float(p16);  // normalise by the weights actually applied
    sts16(pbuf + lane * 2, __half_as_ushort(p16));
    __syncwarp();

    // ---- PV: lane owns rows lane/TPR + j*(32/TPR), tokens 8*(lane%TPR)..+7 of each block ----
#pragma unroll
    for (int b = 0; b < C::BPS; ++b) {
      const uint4 pv = lds128(pbuf + (b * B + (lane % C::TPR) * 8) * 2);
      float c1[C::NJ];
#pragma unroll
      for (int j = 0; j < C::NJ; ++j) c1[j] = 0.f;
      if (!tail) {
#pragma unroll
        for (int j = 0; j < C::NJ; ++j) {
          const uint4 vv = lds128(sbase + (C::BPS + b) * C::TILE + (lane + 32 * j) * 16);
          dot8(pv, vv, acc[j], c1[j]);
        }
      } else {
        const int nv = tok_end - (stage_tok0 + b * B + (lane % C::TPR) * 8);
#pragma unroll
        for (int j = 0; j < C::NJ; ++j) {
          const uint4 vv = keep_first(lds128(sbase + (C::BPS + b) * C::TILE + (lane + 32 * j) * 16), nv);
          dot8(pv, vv, acc[j], c1[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < C::NJ; ++j) acc[j] += c1[j];
    }
    __syncwarp();  // every lane is done with this slot and with pbuf

    // ---- refill this slot with the stage NST rounds ahead ----
    const int st_next = st + NST * W;
    if (st_next < nstages) {
      issue_stage(st_next, slot, next_phys);
      next_phys = (st_next + W < nstages) ? load_phys(st_next + W) : 0;
    }
  }

  // ---- per-warp totals -> smem ----
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
#pragma unroll
  for (int j = 0; j < C::NJ; ++j) {
#pragma unroll
    for (int o = C::TPR / 2; o >= 1; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  float* red = reinterpret_cast<float*>(smem + C::RED_OFF);  // [W][D + 2]
  if (lane % C::TPR == 0) {
#pragma unroll
    for (int j = 0; j < C::NJ; ++j) red[warp * (D + 2) + lane / C::TPR + j * (32 / C::TPR)] = acc[j];
  }
  if (lane == 0) {
    red[warp * (D + 2) + D] = m;
    red[warp * (D + 2) + D + 1] = lsum;
  }
  __syncthreads();

  // ---- merge the W warps (log-sum-exp) ----
  float M = -INFINITY;
#pragma unroll
  for (int w = 0; w < W; ++w) M = fmaxf(M, red[w * (D + 2) + D]);
  float L = 0.f;
  float wgt[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const float mw = red[w * (D + 2) + D];
    wgt[w] = (mw == -INFINITY) ? 0.f : ex2(mw - M);
    L += red[w * (D + 2) + D + 1] * wgt[w];
  }
  const int64_t row = (int64_t)seq * p.num_heads + head;

  if (nsplit_seq == 1) {
    // reference normaliser: 1 / (sum + 1e-6)  (attention_kernels.cu:342)
    const float inv = __fdividef(1.f, L + 1e-6f);
    for (int d = threadIdx.x; d < D; d += W * 32) {
      float o = 0.f;
#pragma unroll
      for (int w = 0; w < W; ++w) o += red[w * (D + 2) + d] * wgt[w];
      p.out[row * D + d] = __float2half_rn(o * inv);
    }
    return;
  }

  // ---- split-KV: publish the partial, last arriver of this (seq, head) merges ----
  {
    const int64_t part_idx = row * p.num_splits + split;
    for (int d = threadIdx.x; d < D; d += W * 32) {
      float o = 0.f;
#pragma unroll
      for (int w = 0; w < W; ++w) o += red[w * (D + 2) + d] * wgt[w];
      p.ws_o[part_idx * D + d] = o;
    }
    if (threadIdx.x == 0) {
      p.ws_ml[part_idx * 2] = M;
      p.ws_ml[part_idx * 2 + 1] = L;
    }
  }
  __threadfence();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(p.ws_cnt + row, 1);
    s_last = (prev == nsplit_seq - 1);
    if (s_last) p.ws_cnt[row] = 0;  // leave the workspace reusable
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    const float* ml = p.ws_ml + row * p.num_splits * 2;
    float Mg = -INFINITY;
    for (int i = 0; i < nsplit_seq; ++i) Mg = fmaxf(Mg, ld_cg_f32(ml + 2 * i));
    float Lg = 0.f;
    for (int i = 0; i < nsplit_seq; ++i) {
      const float mi = ld_cg_f32(ml + 2 * i);
      Lg += ld_cg_f32(ml + 2 * i + 1) * ((mi == -INFINITY) ? 0.f : ex2(mi - Mg));
    }
    const float inv = __fdividef(1.f, Lg + 1e-6f);
    for (int d = threadIdx.x; d < D; d += W * 32) {
      float o = 0.f;
      for (int i = 0; i < nsplit_seq; ++i) {
        const float mi = ld_cg_f32(ml + 2 * i);
        const float wi = (mi == -INFINITY) ? 0.f : ex2(mi - Mg);
        o += ld_cg_f32(p.ws_o + (row * p.num_splits + i) * D + d) * wi;
      }
      p.out[row * D + d] = __float2half_rn(o * inv);
    }
  }
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
constexpr int kCounterCap = 16384;  // int32 counters at the start of the workspace

struct Tuning {
  int force_splits = 0;
  int variant = 0;
};
static Tuning g_tuning;

struct Variant {
  int W, NST;
};
// variant 0: 4 warps x 2 stages (64 KB/CTA at D=64 -> 3 CTAs/SM); 1: 4 x 3; 2: 8 x 1
static constexpr Variant kVariants[] = {{4, 2}, {4, 3}, {8, 1}};
constexpr int kNumVariants = 3;

template <int D, int B, int W, int NST>
static int launch_one(const PaParams& p, int num_seqs, cudaStream_t stream) {
  using C = PaCfg<D, B, W, NST>;
  auto kern = pa_v1_kernel<D, B, W, NST>;
  static bool attr_set[64] = {};
  int dev = 0;
  VB200_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    VB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set[dev] = true;
  }
  dim3 grid(p.num_heads, num_seqs, p.num_splits);
  kern<<<grid, W * 32, C::SMEM, stream>>>(p);
  count_launch();
  VB200_CUDA_CHECK(cudaGetLastError());
  return VB200_OK;
}

template <int D, int B>
static int launch_variant(const PaParams& p, int num_seqs, int variant, cudaStream_t stream) {
  switch (variant) {
    case 1: return launch_one<D, B, 4, 3>(p, num_seqs, stream);
    case 2: return launch_one<D, B, 8, 1>(p, num_seqs, stream);
    default: return launch_one<D, B, 4, 2>(p, num_seqs, stream);
  }
}

template <int D>
static int launch_block(const PaParams& p, int num_seqs, int block_size, int variant, cudaStream_t stream) {
  switch (block_size) {
    case 8: return launch_variant<D, 8>(p, num_seqs, variant, stream);
    case 16: return launch_variant<D, 16>(p, num_seqs, variant, stream);
    case 32: return launch_variant<D, 32>(p, num_seqs, variant, stream);
    default: return set_error(VB200_ERR_BLOCK_SIZE, "Unsupported block size: %d", block_size);
  }
}

static int smem_bytes_for(int D, int variant) {
  const Variant v = kVariants[variant];
  return v.W * v.NST * 128 * D + v.W * 64 + v.W * v.NST * 8 + v.W * (D + 2) * 4;
}

// split-KV factor: split only when (seq, head) pairs alone cannot fill ~2 waves of resident CTAs
static void choose_split(int num_seqs, int num_heads, int head_size, int max_seq_len, int variant, int num_sms,
                         int* num_splits, int* split_tokens) {
  const Variant v = kVariants[variant];
  const int smem = smem_bytes_for(head_size, variant) + 1024;
  int ctas_per_sm = (227 * 1024) / smem;
  ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 16 ? 16 : ctas_per_sm);
  const int64_t slots = (int64_t)num_sms * ctas_per_sm;
  const int64_t base = (int64_t)num_seqs * num_heads;
  const int quantum = 32 * v.W;  // tokens consumed by one round of the CTA's warps
  int ns = 1;
  if (g_tuning.force_splits > 0) {
    ns = g_tuning.force_splits;
  } else if (base > 0 && base < 2 * slots && base <= kCounterCap) {
    ns = (int)((2 * slots + base - 1) / base);
    const int max_by_len = max_seq_len / (2 * quantum);  // >= 2 rounds per split
    if (ns > max_by_len) ns = max_by_len;
    if (ns > 64) ns = 64;
  }
  if (ns < 1) ns = 1;
  int st = max_seq_len > 0 ? (max_seq_len + ns - 1) / ns : quantum;
  st = (st + quantum - 1) / quantum * quantum;
  ns = max_seq_len > 0 ? (max_seq_len + st - 1) / st : 1;
  if (base > kCounterCap) ns = 1;
  *num_splits = ns < 1 ? 1 : ns;
  *split_tokens = st;
}

static size_t workspace_bytes_for(int num_seqs, int num_heads, int head_size, int num_splits) {
  if (num_splits <= 1) return 0;
  const size_t rows = (size_t)num_seqs * num_heads * num_splits;
  return (size_t)kCounterCap * 4 + rows * 2 * 4 + rows * head_size * 4;
}

}  // namespace vb200

using namespace vb200;

extern "C" size_t vb200_paged_attention_v1_workspace_bytes(int32_t num_seqs, int32_t num_heads, int32_t head_size,
                                                           int32_t block_size, int32_t max_seq_len) {
  (void)block_size;
  if (num_seqs <= 0 || num_heads <= 0 || (head_size != 64 && head_size != 128)) return 0;
  int variant = g_tuning.variant;
  if (variant < 0 || variant >= kNumVariants) variant = 0;
  int ns = 1, st = 0;
  choose_split(num_seqs, num_heads, head_size, max_seq_len, variant, device_sm_count(), &ns, &st);
  return workspace_bytes_for(num_seqs, num_heads, head_size, ns);
}

extern "C" int vb200_paged_attention_v1_f16(void* out, const void* query, const void* key_cache,
                                            const void* value_cache, const int32_t* block_tables,
                                            const int32_t* seq_lens, int32_t num_seqs, int32_t num_heads,
                                            int32_t num_kv_heads, int32_t head_size, int32_t block_size,
                                            int32_t max_num_blocks_per_seq, int32_t max_seq_len, int64_t q_stride,
                                            int64_t kv_block_stride, int64_t kv_head_stride, float scale,
                                            void* workspace, size_t workspace_bytes, int32_t device, void* stream) {
  if (num_seqs < 0 || num_heads <= 0 || num_kv_heads <= 0 || max_num_blocks_per_seq < 0)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "negative or zero size argument");
  if (block_size != 8 && block_size != 16 && block_size != 32)
    return set_error(VB200_ERR_BLOCK_SIZE, "Unsupported block size: %d", block_size);
  if (head_size != 64 && head_size != 128)
    return set_error(VB200_ERR_HEAD_SIZE, "Unsupported head size: %d", head_size);
  if (num_heads % num_kv_heads != 0)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "num_heads %d not a multiple of num_kv_heads %d", num_heads,
                     num_kv_heads);
  if (num_seqs == 0) return VB200_OK;
  if (!out || !query || !key_cache || !value_cache || !block_tables || !seq_lens)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "null pointer argument");
  if (num_seqs > 65535) return set_error(VB200_ERR_INVALID_ARGUMENT, "num_seqs %d exceeds grid.y limit", num_seqs);
  // 16-byte vector / bulk-copy alignment (the reference's vector loads assume the same)
  if (((uintptr_t)query | (uintptr_t)key_cache | (uintptr_t)value_cache) & 15)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "query/key_cache/value_cache must be 16-byte aligned");
  if ((q_stride % 8) || (kv_block_stride % 8) || (kv_head_stride % 8))
    return set_error(VB200_ERR_INVALID_ARGUMENT, "strides must be multiples of 8 elements (16 bytes)");
  if (kv_head_stride < (int64_t)head_size * block_size)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "kv_head_stride smaller than one (block, head) tile");

  DeviceGuard guard(device);
  if (guard.status != cudaSuccess) return set_error(VB200_ERR_CUDA, "cudaSetDevice(%d): %s", device,
                                                     cudaGetErrorString(guard.status));
  int variant = g_tuning.variant;
  if (variant < 0 || variant >= kNumVariants) variant = 0;

  PaParams p;
  p.out = static_cast<__half*>(out);
  p.q = static_cast<const __half*>(query);
  p.kc = static_cast<const __half*>(key_cache);
  p.vc = static_cast<const __half*>(value_cache);
  p.block_tables = block_tables;
  p.seq_lens = seq_lens;
  p.num_heads = num_heads;
  p.num_kv_heads = num_kv_heads;
  p.max_blocks_per_seq = max_num_blocks_per_seq;
  p.max_seq_len = max_seq_len;
  p.q_stride = q_stride;
  p.kv_block_stride = kv_block_stride;
  p.kv_head_stride = kv_head_stride;
  p.scale_log2e = (float)((double)scale * 1.4426950408889634074);
  int ns = 1, st = 0;
  if (workspace != nullptr && max_seq_len > 0)
    choose_split(num_seqs, num_heads, head_size, max_seq_len, variant, device_sm_count(), &ns, &st);
  if (ns > 1 && workspace_bytes_for(num_seqs, num_heads, head_size, ns) > workspace_bytes) ns = 1;
  if (ns > 1 && ((uintptr_t)workspace & 15)) ns = 1;
  p.num_splits = ns;
  p.split_tokens = st;
  p.ws_cnt = nullptr;
  p.ws_ml = nullptr;
  p.ws_o = nullptr;
  if (ns > 1) {
    const size_t rows = (size_t)num_seqs * num_heads * ns;
    p.ws_cnt = static_cast<int32_t*>(workspace);
    p.ws_ml = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + (size_t)kCounterCap * 4);
    p.ws_o = p.ws_ml + rows * 2;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (head_size == 64) return launch_block<64>(p, num_seqs, block_size, variant, s);
  return launch_block<128>(p, num_seqs, block_size, variant, s);
}

extern "C" int32_t vb200_set_tuning(int32_t force_num_splits, int32_t variant) {
  const int32_t prev = (g_tuning.force_splits << 8) | (g_tuning.variant & 0xff);
  g_tuning.force_splits = force_num_splits < 0 ? 0 : force_num_splits;
  g_tuning.variant = (variant < 0 || variant >= kNumVariants) ? 0 : variant;
  return prev;
}
