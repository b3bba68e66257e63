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
//                                                split-KV: the splits of a (seq, head) form a thread-block
//                                                cluster and merge through distributed shared memory
//
// Layout facts used (vllmini/kv_cache.py:13-14): for one (physical block,
// kv head) the K tile [D/8][B][8] and the V tile [D][B] are each D*B*2 contiguous bytes, so one
// bulk copy moves a whole tile.  A lane reads 16-byte chunks c = lane + 32*j of a tile:
//   K chunk c -> token c % B, dims 8*(c / B) .. +7        (lane owns ONE token, D*B/256 dim groups)
//   V chunk c -> row   c / (B/8), tokens 8*(c % (B/8)) .. +7  (lane owns D*B/256 rows, 8 tokens)
#include <algorithm>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>

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
  int spec_ctas;     // CTAs with a linear block index below this (= the launch's first wave: SMs x CTAs per SM) prefetch
                     // their first ring into L2 ahead of griddepcontrol.wait; 0 = nobody does
  int cluster_merge; // != 0: the num_splits CTAs of a (seq, head) form one thread-block cluster (1, 1, num_splits) and
                     // merge their partials through distributed shared memory -- no workspace, no counters
  // fused append (vb200_paged_attention_v1_append_f16): this step's K/V rows [num_seqs, num_kv_heads, D] are
  // scattered into the caches by the CTA that attends over the new token.  new_k == nullptr: plain v1.
  const __half* __restrict__ new_k;
  const __half* __restrict__ new_v;
  const int64_t* __restrict__ slot_mapping;
  int64_t new_k_stride;
  int64_t new_v_stride;
  __half* kc_w;
  __half* vc_w;
  // paged_attention_v2 (vb200_paged_attention_v2_f16): grid.z = the caller's 512-token partitions; every CTA writes
  // its partition's softmax statistics and normalised output into the CALLER's scratch tensors and a second kernel
  // reduces them (attention_kernels.cu:349-357, 564-669).  v2_exp_sums == nullptr: v1.
  float* v2_exp_sums;    // [num_seqs, num_heads, num_splits]
  float* v2_max_logits;  // [num_seqs, num_heads, num_splits]
  __half* v2_tmp_out;    // [num_seqs, num_heads, num_splits, D]
#ifdef VB200_TRACE
  unsigned long long* trace;  // [num_ctas][8]: smid, t_start, t_first_data, t_end, t_range_known, t_loop_end,
                              //                t_merge_start, t_after_first_cluster_barrier (globaltimer ns)
#endif
};

template <int D, int B, int W, int NST>
struct PaCfg {
  static constexpr int BPS = 32 / B;             // KV blocks per 32-token stage
  static constexpr int TILE = D * B * 2;         // bytes of one (block, head) K tile (== V tile)
  static constexpr int CH = D * B / 8;           // 16-byte chunks per tile
  static constexpr int NJ = (CH + 31) / 32;      // chunks per lane per tile (the last one partial unless FULL)
  static constexpr bool FULL = CH % 32 == 0;     // false only for head sizes 80 / 112 at block size 8
  static constexpr bool QSMEM = NJ > 16;         // q stays in shared memory (head sizes 192 / 256 at block size 32)
  static constexpr int NQ = QSMEM ? 1 : NJ;      // q chunks held in registers
  static constexpr int TPR = B / 8;              // 16-byte chunks per V row
  static constexpr int STAGE = 2 * BPS * TILE;   // bytes per stage: K tiles then V tiles (= 128*D)
  static constexpr int RING = W * NST * STAGE;
  static constexpr int PBUF_OFF = RING;                     // W * 32 halves
  static constexpr int BAR_OFF = PBUF_OFF + W * 64;         // W * NST mbarriers
  static constexpr int RED_OFF = BAR_OFF + W * NST * 8;     // W * (D + 2) floats
  static constexpr int Q_OFF = (RED_OFF + W * (D + 2) * 4 + 15) / 16 * 16;  // D halves (used when QSMEM)
  static constexpr int SMEM = Q_OFF + D * 2;
  static constexpr int kD = D, kB = B;
  static_assert(B == 8 || B == 16 || B == 32, "block size");
  static_assert(D % 16 == 0 && D >= 16 && D <= 256, "head size");
  static constexpr bool FITS = SMEM <= 227 * 1024;
};

// does chunk j of this lane exist in the tile?  (compile-time true except in the last chunk of a partial tile)
template <class C>
__device__ __forceinline__ bool chunk_ok(int lane, int j) {
  return C::FULL || j < C::NJ - 1 || lane + 32 * j < C::CH;
}

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

// ---------------------------------------------------------------------------------------------
// Shared device pieces
// ---------------------------------------------------------------------------------------------

// One 32-token stage resident in shared memory at `sbase` (K tiles then V tiles): QK^T, online
// softmax update of (m, lsum), PV accumulation into acc[].  Warp-synchronous; `pbuf` is the warp's
// 64-byte fp16 probability buffer.  Ends with a __syncwarp(): afterwards every lane is done with the
// stage's shared memory and with pbuf.
template <class C>
__device__ __forceinline__ void process_stage(uint32_t sbase, uint32_t pbuf, int lane, const uint4 (&qv)[C::NQ],
                                              uint32_t qs, float& m, float& lsum, float (&acc)[C::NJ], int stage_tok0,
                                              int tok_end, float scale_log2e) {
  constexpr int B = C::kB;
  auto qchunk = [&](int j) -> uint4 {  // q dims 8*((lane + 32j)/B) .. +7
    if constexpr (C::QSMEM) {
      return lds128(qs + ((lane + 32 * j) / B) * 16);
    } else {
      return qv[j];
    }
  };
  // ---- QK^T: lane owns token lane % B of every block of the stage ----
  float part[C::BPS];
#pragma unroll
  for (int b = 0; b < C::BPS; ++b) {
    float c0 = 0.f, c1 = 0.f;
#pragma unroll
    for (int j = 0; j < C::NJ; ++j) {
      if (chunk_ok<C>(lane, j)) {
        const uint4 kv = lds128(sbase + b * C::TILE + (lane + 32 * j) * 16);
        dot8(qchunk(j), kv, c0, c1);
      }
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
  const bool tail = stage_tok0 + 32 > tok_end;  // warp-uniform
  s *= scale_log2e;
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
        if (chunk_ok<C>(lane, j)) {
          const uint4 vv = lds128(sbase + (C::BPS + b) * C::TILE + (lane + 32 * j) * 16);
          dot8(pv, vv, acc[j], c1[j]);
        }
      }
    } else {
      const int nv = tok_end - (stage_tok0 + b * B + (lane % C::TPR) * 8);
#pragma unroll
      for (int j = 0; j < C::NJ; ++j) {
        if (chunk_ok<C>(lane, j)) {
          const uint4 vv = keep_first(lds128(sbase + (C::BPS + b) * C::TILE + (lane + 32 * j) * 16), nv);
          dot8(pv, vv, acc[j], c1[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < C::NJ; ++j) acc[j] += c1[j];
  }
  __syncwarp();
}

// Issue the bulk copies of one stage (K and V tile of each block that exists).
// `phys`: lane b < BPS holds the physical block id of block b of the stage; nb = blocks that exist.
template <class C>
__device__ __forceinline__ void issue_stage_copies(uint32_t dst_stage, uint32_t bar, int lane, int phys, int nb,
                                                   const __half* kbase, const __half* vbase,
                                                   int64_t kv_block_stride) {
  // All operands of the copies are warp-uniform (constant-lane shuffles, a warp index broadcast from lane 0), and one
  // elected lane issues them in straight-line code: the compiler can keep addresses in uniform registers instead of
  // looping over the lanes that each hold one copy.
  int pb[C::BPS];
#pragma unroll
  for (int b = 0; b < C::BPS; ++b) pb[b] = __shfl_sync(0xffffffffu, phys, b);
  if (elect_one()) {
    mbar_arrive_expect_tx(bar, (uint32_t)(2 * nb * C::TILE));
#pragma unroll
    for (int b = 0; b < C::BPS; ++b) {
      if (b < nb) {
        const int64_t off = (int64_t)pb[b] * kv_block_stride;
        bulk_g2s(dst_stage + b * C::TILE, kbase + off, C::TILE, bar);
        bulk_g2s(dst_stage + (C::BPS + b) * C::TILE, vbase + off, C::TILE, bar);
      }
    }
  }
}

// Warp totals -> red[warp][D + 2] = {acc rows..., m, lsum}
template <class C>
__device__ __forceinline__ void warp_publish(float* red, int warp, int lane, float m, float lsum,
                                             float (&acc)[C::NJ]) {
  constexpr int D = C::kD;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
#pragma unroll
  for (int j = 0; j < C::NJ; ++j) {
#pragma unroll
    for (int o = C::TPR / 2; o >= 1; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  if (lane % C::TPR == 0) {
#pragma unroll
    for (int j = 0; j < C::NJ; ++j)
      if (chunk_ok<C>(lane, j)) red[warp * (D + 2) + lane / C::TPR + j * (32 / C::TPR)] = acc[j];
  }
  if (lane == 0) {
    red[warp * (D + 2) + D] = m;
    red[warp * (D + 2) + D + 1] = lsum;
  }
}

// CTA-level log-sum-exp merge of the W warps' partials (after a barrier), then the final normalised fp16 row, the
// paged_attention_v2 partition record, or -- split-KV -- the merge across the CTAs of the cluster.
// `tid` in [0, W*32); `part`: (D + 2) floats of this CTA's shared memory, at the same offset in every CTA.
template <class C, int W>
__device__ __forceinline__ void cta_merge_and_store(const PaParams& p, const float* red, int tid, int seq, int head,
                                                    int split, int nsplit_seq, float* part) {
  constexpr int D = C::kD;
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

  if (p.v2_exp_sums != nullptr) {
    // paged_attention_v2 partition record, as the reference's partition kernel leaves it (attention_kernels.cu:
    // 349-357, 484-495): max logit in natural-log units, exp_sum = sum exp(logit - max), tmp_out = the partition's
    // own softmax-weighted V (normalised by 1 / (exp_sum + 1e-6)), fp16.
    const int64_t part_idx = row * p.num_splits + split;
    const float inv = __fdividef(1.f, L + 1e-6f);
    for (int d = tid; d < D; d += W * 32) {
      float o = 0.f;
#pragma unroll
      for (int w = 0; w < W; ++w) o += red[w * (D + 2) + d] * wgt[w];
      p.v2_tmp_out[part_idx * D + d] = __float2half_rn(o * inv);
    }
    if (tid == 0) {
      p.v2_max_logits[part_idx] = M * 0.69314718055994530942f;  // log2 domain -> natural log
      p.v2_exp_sums[part_idx] = L;
    }
    return;
  }

  if (p.cluster_merge) {
    // ---- split-KV inside a thread-block cluster: every CTA of the cluster (also one whose split holds no token)
    //      parks its partial {o[D], M, L} in its own shared memory (`part`: the start of the ring, idle by now),
    //      a cluster barrier publishes them, the CTA of split 0 merges the live splits through distributed shared
    //      memory and writes the row, and a second barrier keeps every CTA's shared memory alive until then. ----
    for (int d = tid; d < D; d += W * 32) {
      float o = 0.f;
#pragma unroll
      for (int w = 0; w < W; ++w) o += red[w * (D + 2) + d] * wgt[w];
      part[d] = o;
    }
    if (tid == 0) {
      part[D] = M;
      part[D + 1] = L;
    }
    cluster_sync_all();
#ifdef VB200_TRACE
    if (tid == 0 && p.trace != nullptr) {
      unsigned long long t_cb;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_cb));
      p.trace[(((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + 7] = t_cb;
    }
#endif
    if (split == 0) {
      // A remote shared-memory read costs ~0.1 us and the cluster holds up to 16 partials, so the reads are spread
      // and batched instead of looped over: lane i of warp 0 fetches {M_i, L_i} of split i (one round trip for all
      // splits), the 16 lanes reduce max and sum with a fixed shuffle tree and leave the weights 2^(M_i - M) in this
      // CTA's shared memory; then thread d fetches o_i[d] of 8 splits at a time, back to back, and accumulates in
      // split-index order.
      float* wgt_s = part + D + 2;             // 16 weights + 1 / (L + 1e-6), behind this CTA's own partial
      if (tid < 32) {
        const uint32_t base = smem_u32(part);
        float mi = -INFINITY, li = 0.f;
        if (tid < nsplit_seq) {
          mi = ld_dsmem_f32(dsmem_addr(base + D * 4, tid));
          li = ld_dsmem_f32(dsmem_addr(base + (D + 1) * 4, tid));
        }
        float Mg = mi;
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) Mg = fmaxf(Mg, __shfl_xor_sync(0xffffffffu, Mg, o));
        const float wi = (mi == -INFINITY) ? 0.f : ex2(mi - Mg);
        float Lg = li * wi;
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) Lg += __shfl_xor_sync(0xffffffffu, Lg, o);
        if (tid < 16) wgt_s[tid] = wi;
        if (tid == 0) wgt_s[16] = __fdividef(1.f, Lg + 1e-6f);   // reference normaliser (attention_kernels.cu:342)
      }
      __syncthreads();
      const float inv = wgt_s[16];
      for (int d = tid; d < D; d += W * 32) {
        const uint32_t mine = smem_u32(part + d);
        float o = 0.f;
#pragma unroll 1
        for (int i0 = 0; i0 < nsplit_seq; i0 += 8) {
          float oi[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) oi[u] = (i0 + u < nsplit_seq) ? ld_dsmem_f32(dsmem_addr(mine, i0 + u)) : 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) o += oi[u] * wgt_s[i0 + u];
        }
        p.out[row * D + d] = __float2half_rn(o * inv);
      }
    }
    cluster_sync_all();
    return;
  }

  if (nsplit_seq == 1) {
    // reference normaliser: 1 / (sum + 1e-6)  (attention_kernels.cu:342)
    const float inv = __fdividef(1.f, L + 1e-6f);
    for (int d = tid; d < D; d += W * 32) {
      float o = 0.f;
#pragma unroll
      for (int w = 0; w < W; ++w) o += red[w * (D + 2) + d] * wgt[w];
      p.out[row * D + d] = __float2half_rn(o * inv);
    }
    return;
  }

  // (a split launch always runs as a cluster: nsplit_seq > 1 without cluster_merge does not occur)
}

// Token range of (seq, split) and derived counts.
struct UnitRange {
  int tok_begin, tok_end, nstages, nblk, nsplit_seq;
  bool empty_split;  // split > 0 and no token in it: nothing to do, nobody waits for it
};
template <int B>
__device__ __forceinline__ UnitRange unit_range_of_len(const PaParams& p, int len, int split) {
  len = len < p.max_blocks_per_seq * B ? len : p.max_blocks_per_seq * B;  // never read past the table row
  const bool v2 = p.v2_exp_sums != nullptr;
  const bool parts = p.num_splits > 1 || v2;
  // v2: the caller's scratch has room for num_splits partitions and no more.  v1: max_seq_len is only a hint for the
  // split factor -- the LAST split runs to seq_len, so the result never depends on whether the launch was split.
  if (v2 && len > p.num_splits * p.split_tokens) len = p.num_splits * p.split_tokens;
  UnitRange r;
  r.tok_begin = parts ? split * p.split_tokens : 0;
  // nothing to do, nobody waits for it (v2: also partition 0 of an empty sequence, attention_kernels.cu:119-122)
  r.empty_split = (split > 0 || v2) && r.tok_begin >= len;
  r.tok_end = len;
  if (parts && split < p.num_splits - 1 && r.tok_begin + p.split_tokens < len) r.tok_end = r.tok_begin + p.split_tokens;
  const int ntok = r.tok_end > r.tok_begin ? r.tok_end - r.tok_begin : 0;
  r.nstages = (ntok + 31) >> 5;
  r.nblk = (ntok + B - 1) / B;
  r.nsplit_seq = 1;
  if (p.num_splits > 1 && len > 0) {
    const int n = (len + p.split_tokens - 1) / p.split_tokens;
    r.nsplit_seq = n < p.num_splits ? n : p.num_splits;
  }
  return r;
}

template <int B>
__device__ __forceinline__ UnitRange unit_range(const PaParams& p, int seq, int split) {
  return unit_range_of_len<B>(p, p.seq_lens[seq], split);
}

// ---------------------------------------------------------------------------------------------
// Classic kernel: one CTA per (head, seq, split)
// ---------------------------------------------------------------------------------------------
template <int D, int B, int W, int NST, bool APPEND>
__global__ void __launch_bounds__(W * 32) pa_v1_kernel(const PaParams p) {
  using C = PaCfg<D, B, W, NST>;
  extern __shared__ __align__(128) uint8_t smem[];

  const int head = blockIdx.x;
  const int seq = blockIdx.y;
  const int split = blockIdx.z;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // broadcast: warp-uniform for the compiler

  // PDL: this CTA may have become resident while the previous kernel on the stream was still draining.  Let
  // our own successor do the same, set up the barriers, then wait for the predecessor before the first global
  // access (seq_lens, block tables, q and the caches may all be its outputs).
#ifdef VB200_TRACE
  unsigned long long t_start = 0, t_first = 0, t_range = 0, t_loop = 0, t_merge = 0;
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
#endif
  griddep_launch_dependents();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < W * NST; ++i) mbar_init(smem_u32(smem + PaCfg<D, B, W, NST>::BAR_OFF) + i * 8, 1);
    fence_mbar_init();
  }
  // ---- speculative L2 prefetch of this warp's first ring, AHEAD of griddepcontrol.wait.  With programmatic
  //      dependent launch the CTAs of the first wave are resident while the previous kernel on the stream still runs
  //      (its tail leaves HBM underused); seq_lens and the block table may be that kernel's output, so whatever is
  //      read here can be stale or garbage -- it is used for HINTS only (prefetch.global.L2: no destination, dropped on
  //      an unmapped address, tools/prefetch_wild_repro.cu), read with L2-only loads (no stale line is left in L1),
  //      and everything is read again after the wait.  A prefetched K/V line that the predecessor then writes (the new
  //      token's slot) is updated in L2, the point of coherence, like any other.  CTAs of later waves start when the
  //      launch is under way and skip this (their hints would only duplicate the copies that follow). ----
  if ((int)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) < p.spec_ctas) {
    int len_s = ld_cg_s32(p.seq_lens + seq);
    len_s = len_s < 0 ? 0 : len_s;
    const UnitRange rs = unit_range_of_len<B>(p, len_s, split);
    const int kv_head_s = head / (p.num_heads / p.num_kv_heads);
    const char* kb = reinterpret_cast<const char*>(p.kc + (int64_t)kv_head_s * p.kv_head_stride);
    const char* vb = reinterpret_cast<const char*>(p.vc + (int64_t)kv_head_s * p.kv_head_stride);
    const int32_t* table_s = p.block_tables + (int64_t)seq * p.max_blocks_per_seq + rs.tok_begin / B;
    constexpr int LINES = C::TILE / 128;   // 128-byte lines per K (or V) tile
    // (one ring deep: two rings gain less, eight lose on large batches -- profiles/r02j_spec_prefetch.md)
#pragma unroll
    for (int k = 0; k < NST; ++k) {
      const int st = warp + k * W;
      if (st >= rs.nstages) break;   // (warp-uniform)
#pragma unroll
      for (int b = 0; b < C::BPS; ++b) {
        const int blk = st * C::BPS + b;
        if (blk < rs.nblk) {
          const int64_t off = (int64_t)ld_cg_s32(table_s + blk) * p.kv_block_stride * 2;   // bytes
          for (int l = lane; l < 2 * LINES; l += 32)
            prefetch_l2(l < LINES ? kb + off + l * 128 : vb + off + (l - LINES) * 128);
        }
      }
    }
    __syncwarp();   // 2 * LINES need not be a multiple of 32: converge again before the CTA barrier below
  }
  griddep_wait();

  const UnitRange r = unit_range<B>(p, seq, split);
  // a split without tokens has nothing to do -- unless its cluster waits for it at the merge barriers
  if (r.empty_split && !p.cluster_merge) return;
  const int nstages = r.nstages, nblk = r.nblk;
#ifdef VB200_TRACE
  if (threadIdx.x == 0 && nstages >= 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_range));
#endif

  const int kv_head = head / (p.num_heads / p.num_kv_heads);
  const int32_t* table = p.block_tables + (int64_t)seq * p.max_blocks_per_seq + r.tok_begin / B;
  const __half* kbase = p.kc + (int64_t)kv_head * p.kv_head_stride;
  const __half* vbase = p.vc + (int64_t)kv_head * p.kv_head_stride;

  const uint32_t ring = smem_u32(smem) + warp * (NST * C::STAGE);
  const uint32_t pbuf = smem_u32(smem + C::PBUF_OFF) + warp * 64;
  const uint32_t bars = smem_u32(smem + C::BAR_OFF) + warp * (NST * 8);

  // ---- fused append (APPEND instantiation only): the CTA of the LAST live split of (seq, head) also stores this
  //      step's K/V row of its kv head into the caches -- the same element moves as reshape_and_cache
  //      (cache_kernels.cu:178-199).  Nothing of it sits on the critical path: the warp that owns the last
  //      stage prefetches the slot and the row into registers now, patches the row into its shared-memory
  //      tile when that stage has landed (if the row is the newest attended token, position seq_len - 1; the
  //      bulk copy may have raced with nothing, the row is not in global memory yet), and writes the row to
  //      global memory after its last stage.  Contract: the new token is at position >= seq_len - 1, so no
  //      other split reads that slot. ----
  const int row_warp = nstages > 0 ? (nstages - 1) % W : 0;
  const bool row_owner = APPEND && p.new_k != nullptr && split == r.nsplit_seq - 1 && warp == row_warp;
  int64_t new_slot = -1;
  int64_t phys_last = -1;
  uint4 patch_k = make_uint4(0, 0, 0, 0);   // lane g < D/8: the row's 16-byte group g
  uint32_t patch_v[(D + 63) / 64];          // lane: words lane, lane + 32, ... of the V row (2 dims per word)
#pragma unroll
  for (int u = 0; u < (D + 63) / 64; ++u) patch_v[u] = 0;
  if (row_owner) {
    new_slot = p.slot_mapping[seq];
    const __half* ksrc = p.new_k + (int64_t)seq * p.new_k_stride + (int64_t)kv_head * D;
    const __half* vsrc = p.new_v + (int64_t)seq * p.new_v_stride + (int64_t)kv_head * D;
    if (lane < D / 8) patch_k = ldg128(ksrc + lane * 8);
#pragma unroll
    for (int u = 0; u < (D + 63) / 64; ++u)
      if (lane + 32 * u < D / 2) patch_v[u] = __ldg(reinterpret_cast<const uint32_t*>(vsrc) + lane + 32 * u);
    if (nstages > 0) phys_last = __ldg(table + (r.tok_end - 1 - r.tok_begin) / B);
  }
  __syncthreads();

  auto issue_stage = [&](int st, int slot, int phys) {
    const int blk0 = st * C::BPS;
    const int nb = nblk - blk0 < C::BPS ? nblk - blk0 : C::BPS;
    issue_stage_copies<C>(ring + slot * C::STAGE, bars + slot * 8, lane, phys, nb, kbase, vbase, p.kv_block_stride);
  };
  auto load_phys = [&](int st) -> int {  // table entries of stage st, one per lane < BPS
    const int blk = st * C::BPS + lane;
    return (lane < C::BPS && blk < nblk) ? __ldg(table + blk) : 0;
  };

  // prologue: fill the ring.  (Fetching these table entries together with seq_len -- one dependent round trip less --
  // was measured three times, the last time with the first ring already prefetched into L2: profiles/
  // r01i_cta_timeline.md, r02d_spec_table_prefetch.md, r02j_spec_prefetch.md -- and gains nothing.)
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
  uint4 qv[C::NQ];
  const uint32_t qs = smem_u32(smem + C::Q_OFF);
  {
    const __half* qrow = p.q + (int64_t)seq * p.q_stride + (int64_t)head * D;
    if constexpr (C::QSMEM) {  // large head x block: the row lives in shared memory, read per chunk (broadcast)
      qv[0] = make_uint4(0, 0, 0, 0);
      if (threadIdx.x < D / 8) *reinterpret_cast<uint4*>(smem + C::Q_OFF + threadIdx.x * 16) = ldg128(qrow + 8 * threadIdx.x);
      __syncthreads();
    } else {
#pragma unroll
      for (int j = 0; j < C::NJ; ++j)
        qv[j] = chunk_ok<C>(lane, j) ? ldg128(qrow + 8 * (lane / B + j * C::BPS)) : make_uint4(0, 0, 0, 0);
    }
  }

  float m = -INFINITY;  // running max (log2 domain), warp-uniform
  float lsum = 0.f;     // per-lane partial sum of p
  float acc[C::NJ];     // rows lane/TPR + j*(32/TPR), partial over this lane's 8-token groups
#pragma unroll
  for (int j = 0; j < C::NJ; ++j) acc[j] = 0.f;

  int k = 0;
  for (int st = warp; st < nstages; st += W, ++k) {
    const int slot = k % NST;
    mbar_wait(bars + slot * 8, (k / NST) & 1);
#ifdef VB200_TRACE
    if (threadIdx.x == 0 && k == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_first));
#endif
    if (APPEND && row_owner && st == nstages - 1 && new_slot >= 0 &&
        new_slot == phys_last * B + (r.tok_end - 1) % B) {
      // the new row is the newest attended token: put it into the landed tiles of this (last) stage
      const int in_stage = (r.tok_end - 1 - r.tok_begin) - st * 32;   // 0..31
      const int b = in_stage / B;
      const int off = (int)(new_slot % B);
      uint8_t* kt = smem + (warp * NST + slot) * C::STAGE + b * C::TILE;
      uint8_t* vt = smem + (warp * NST + slot) * C::STAGE + (C::BPS + b) * C::TILE;
      if (lane < D / 8) *reinterpret_cast<uint4*>(kt + (lane * B + off) * 16) = patch_k;
#pragma unroll
      for (int u = 0; u < (D + 63) / 64; ++u) {
        const int i = lane + 32 * u;   // dims 2i, 2i+1
        if (i < D / 2) {
          reinterpret_cast<uint16_t*>(vt)[(2 * i) * B + off] = (uint16_t)(patch_v[u] & 0xffffu);
          reinterpret_cast<uint16_t*>(vt)[(2 * i + 1) * B + off] = (uint16_t)(patch_v[u] >> 16);
        }
      }
      __syncwarp();
    }
    process_stage<C>(ring + slot * C::STAGE, pbuf, lane, qv, qs, m, lsum, acc, r.tok_begin + st * 32, r.tok_end,
                     p.scale_log2e);
    // ---- refill this slot with the stage NST rounds ahead ----
    const int st_next = st + NST * W;
    if (st_next < nstages) {
      issue_stage(st_next, slot, next_phys);
      next_phys = (st_next + W < nstages) ? load_phys(st_next + W) : 0;
    }
  }

  if (APPEND && row_owner && new_slot >= 0) {  // fused append: the row goes to the caches, off the critical path
    const int64_t blk = new_slot / B;
    const int off = (int)(new_slot % B);
    __half* kdst = p.kc_w + blk * p.kv_block_stride + (int64_t)kv_head * p.kv_head_stride;
    uint16_t* vdst =
        reinterpret_cast<uint16_t*>(p.vc_w + blk * p.kv_block_stride + (int64_t)kv_head * p.kv_head_stride) + off;
    if (lane < D / 8) *reinterpret_cast<uint4*>(kdst + ((int64_t)lane * B + off) * 8) = patch_k;
#pragma unroll
    for (int u = 0; u < (D + 63) / 64; ++u) {
      const int i = lane + 32 * u;
      if (i < D / 2) {
        vdst[(int64_t)(2 * i) * B] = (uint16_t)(patch_v[u] & 0xffffu);
        vdst[(int64_t)(2 * i + 1) * B] = (uint16_t)(patch_v[u] >> 16);
      }
    }
  }

#ifdef VB200_TRACE
  if (threadIdx.x == 0 && k >= 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_loop));
#endif
  float* red = reinterpret_cast<float*>(smem + C::RED_OFF);  // [W][D + 2]
  warp_publish<C>(red, warp, lane, m, lsum, acc);
  __syncthreads();
#ifdef VB200_TRACE
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_merge));
#endif
  cta_merge_and_store<C, W>(p, red, threadIdx.x, seq, head, split, r.nsplit_seq, reinterpret_cast<float*>(smem));
#ifdef VB200_TRACE
  if (threadIdx.x == 0 && p.trace != nullptr) {
    unsigned long long t_end;
    unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    p.trace[cta * 8 + 0] = smid;
    p.trace[cta * 8 + 1] = t_start;
    p.trace[cta * 8 + 2] = t_first;
    p.trace[cta * 8 + 3] = t_end;
    p.trace[cta * 8 + 4] = t_range;
    p.trace[cta * 8 + 5] = t_loop;
    p.trace[cta * 8 + 6] = t_merge;
  }
#endif
}

template <int D>
constexpr int default_warps_c() { return (D == 64 || D == 96 || D == 192) ? 4 : 3; }   // == default_warps(D), see there

// ---------------------------------------------------------------------------------------------
// Grouped-query variant: one CTA per (G query heads of one kv head, seq, split)
// ---------------------------------------------------------------------------------------------
// With num_kv_heads < num_heads the classic grid gives every QUERY head its own CTA, so a kv head's tiles are
// fetched once per query head (the repeats hit L2, but L2 -> SM bandwidth then bounds the kernel:
// profiles/r01k_sweep_gqa.json).  Here a landed K/V tile is applied to G query heads of its kv head before the
// ring slot is refilled: HBM and L2 traffic per query head drop by G.  The query rows live in shared memory
// (read per chunk, broadcast); softmax state and accumulators are per query head.  Not reached by vllmini (MHA).
template <int D, int B, int W, int NST, int G>
struct PaCfgG : PaCfg<D, B, W, NST> {
  using Base = PaCfg<D, B, W, NST>;
  static constexpr int PBUF_OFF = Base::RING;                                     // W * G * 32 halves
  static constexpr int BAR_OFF = PBUF_OFF + W * G * 64;                           // W * NST mbarriers
  static constexpr int RED_OFF = BAR_OFF + W * NST * 8;                           // G * W * (D + 2) floats
  static constexpr int Q_OFF = (RED_OFF + G * W * (D + 2) * 4 + 15) / 16 * 16;    // G * D halves
  static constexpr int SMEM = Q_OFF + G * D * 2;
  static constexpr bool FITS = SMEM <= 227 * 1024;
  static constexpr int CTAS_PER_SM = FITS ? (227 * 1024) / (SMEM + 1024) : 1;   // what shared memory allows (launch bound)
  static constexpr int kG = G;
};

// CTA shape of the grouped-query kernel.  Head size 128 with 4 heads per CTA needs ~250 registers at 3 warps x 2
// stages (2 CTAs/SM, 6 warps); 4 warps x 1 stage under a 3-CTAs/SM register bound (128 registers, 12 warps) is
// 8-16 % faster (profiles/r01k_gqa.md).  Everything else keeps the default shape of its head size.
template <int D, int G> struct GqaShape { static constexpr int W = default_warps_c<D>(), NST = 2; };
template <> struct GqaShape<128, 4> { static constexpr int W = 4, NST = 1; };

template <class C>
__device__ __forceinline__ void process_stage_gqa(uint32_t sbase, uint32_t pbuf, int lane, uint32_t qs,
                                                  float (&m)[C::kG], float (&lsum)[C::kG],
                                                  float (&acc)[C::kG][C::NJ], int stage_tok0, int tok_end,
                                                  float scale_log2e) {
  constexpr int B = C::kB, G = C::kG;
  // ---- QK^T: every K chunk is read once and multiplied with the G query rows ----
  float s[G];
  {
    float part[G][C::BPS];
#pragma unroll
    for (int b = 0; b < C::BPS; ++b) {
      float c0[G], c1[G];
#pragma unroll
      for (int g = 0; g < G; ++g) c0[g] = c1[g] = 0.f;
#pragma unroll
      for (int j = 0; j < C::NJ; ++j) {
        if (chunk_ok<C>(lane, j)) {
          const uint4 kv = lds128(sbase + b * C::TILE + (lane + 32 * j) * 16);
#pragma unroll
          for (int g = 0; g < G; ++g)
            dot8(lds128(qs + (g * (C::kD / 8) + (lane + 32 * j) / B) * 16), kv, c0[g], c1[g]);
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) part[g][b] = c0[g] + c1[g];
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {  // transpose-reduce as in process_stage
      if constexpr (C::BPS == 1) {
        s[g] = part[g][0];
      } else if constexpr (C::BPS == 2) {
        const bool up = lane & 16;
        const float give = up ? part[g][0] : part[g][1];
        s[g] = (up ? part[g][1] : part[g][0]) + __shfl_xor_sync(0xffffffffu, give, 16);
      } else {
        const bool up = lane & 16;
        const float g0 = up ? part[g][0] : part[g][2];
        const float g1 = up ? part[g][1] : part[g][3];
        const float k0 = (up ? part[g][2] : part[g][0]) + __shfl_xor_sync(0xffffffffu, g0, 16);
        const float k1 = (up ? part[g][3] : part[g][1]) + __shfl_xor_sync(0xffffffffu, g1, 16);
        const bool up2 = lane & 8;
        const float give = up2 ? k0 : k1;
        s[g] = (up2 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, give, 8);
      }
    }
  }
  const bool tail = stage_tok0 + 32 > tok_end;  // warp-uniform
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float sg = s[g] * scale_log2e;
    if (stage_tok0 + lane >= tok_end) sg = -INFINITY;
    const float smax = warp_max_f32(sg);
    if (smax > m[g]) {
      const float alpha = ex2(m[g] - smax);
      m[g] = smax;
      lsum[g] *= alpha;
#pragma unroll
      for (int j = 0; j < C::NJ; ++j) acc[g][j] *= alpha;
    }
    const __half p16 = __float2half_rn(ex2(sg - m[g]));
    lsum[g] += __half2float(p16);
    sts16(pbuf + g * 64 + lane * 2, __half_as_ushort(p16));
  }
  __syncwarp();

  // ---- PV: every V chunk is read once and multiplied with the G probability rows ----
#pragma unroll
  for (int b = 0; b < C::BPS; ++b) {
    uint4 pv[G];
#pragma unroll
    for (int g = 0; g < G; ++g) pv[g] = lds128(pbuf + g * 64 + (b * B + (lane % C::TPR) * 8) * 2);
    const int nv = tok_end - (stage_tok0 + b * B + (lane % C::TPR) * 8);
#pragma unroll
    for (int j = 0; j < C::NJ; ++j) {
      if (chunk_ok<C>(lane, j)) {
        uint4 vv = lds128(sbase + (C::BPS + b) * C::TILE + (lane + 32 * j) * 16);
        if (tail) vv = keep_first(vv, nv);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          float c1 = 0.f;
          dot8(pv[g], vv, acc[g][j], c1);
          acc[g][j] += c1;
        }
      }
    }
  }
  __syncwarp();
}

template <int D, int B, int W, int NST, bool APPEND, int G>
__global__ void __launch_bounds__(W * 32, PaCfgG<D, B, W, NST, G>::CTAS_PER_SM) pa_v1_gqa_kernel(const PaParams p) {
  using C = PaCfgG<D, B, W, NST, G>;
  extern __shared__ __align__(128) uint8_t smem[];

  const int head0 = blockIdx.x * G;  // first of this CTA's G query heads (all of one kv head)
  const int seq = blockIdx.y;
  const int split = blockIdx.z;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;

  griddep_launch_dependents();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < W * NST; ++i) mbar_init(smem_u32(smem + C::BAR_OFF) + i * 8, 1);
    fence_mbar_init();
  }
  griddep_wait();

  const UnitRange r = unit_range<B>(p, seq, split);
  if (r.empty_split) return;
  const int nstages = r.nstages, nblk = r.nblk;

  const int kv_head = head0 / (p.num_heads / p.num_kv_heads);
  const int32_t* table = p.block_tables + (int64_t)seq * p.max_blocks_per_seq + r.tok_begin / B;
  const __half* kbase = p.kc + (int64_t)kv_head * p.kv_head_stride;
  const __half* vbase = p.vc + (int64_t)kv_head * p.kv_head_stride;

  const uint32_t ring = smem_u32(smem) + warp * (NST * C::STAGE);
  const uint32_t pbuf = smem_u32(smem + C::PBUF_OFF) + warp * (G * 64);
  const uint32_t bars = smem_u32(smem + C::BAR_OFF) + warp * (NST * 8);
  const uint32_t qs = smem_u32(smem + C::Q_OFF);

  // the G query rows -> shared memory (consecutive heads: one contiguous G*D run of the query row)
  {
    const __half* qrow = p.q + (int64_t)seq * p.q_stride + (int64_t)head0 * D;
    for (int c = threadIdx.x; c < G * D / 8; c += W * 32)
      *reinterpret_cast<uint4*>(smem + C::Q_OFF + c * 16) = ldg128(qrow + 8 * c);
  }

  // fused append: as in pa_v1_kernel, but a kv head's row is now written by exactly one CTA when G covers the group
  const int row_warp = nstages > 0 ? (nstages - 1) % W : 0;
  const bool row_owner = APPEND && p.new_k != nullptr && split == r.nsplit_seq - 1 && warp == row_warp;
  int64_t new_slot = -1;
  int64_t phys_last = -1;
  uint4 patch_k = make_uint4(0, 0, 0, 0);
  uint32_t patch_v[(D + 63) / 64];
#pragma unroll
  for (int u = 0; u < (D + 63) / 64; ++u) patch_v[u] = 0;
  if (row_owner) {
    new_slot = p.slot_mapping[seq];
    const __half* ksrc = p.new_k + (int64_t)seq * p.new_k_stride + (int64_t)kv_head * D;
    const __half* vsrc = p.new_v + (int64_t)seq * p.new_v_stride + (int64_t)kv_head * D;
    if (lane < D / 8) patch_k = ldg128(ksrc + lane * 8);
#pragma unroll
    for (int u = 0; u < (D + 63) / 64; ++u)
      if (lane + 32 * u < D / 2) patch_v[u] = __ldg(reinterpret_cast<const uint32_t*>(vsrc) + lane + 32 * u);
    if (nstages > 0) phys_last = __ldg(table + (r.tok_end - 1 - r.tok_begin) / B);
  }
  __syncthreads();

  auto issue_stage = [&](int st, int slot, int phys) {
    const int blk0 = st * C::BPS;
    const int nb = nblk - blk0 < C::BPS ? nblk - blk0 : C::BPS;
    issue_stage_copies<C>(ring + slot * C::STAGE, bars + slot * 8, lane, phys, nb, kbase, vbase, p.kv_block_stride);
  };
  auto load_phys = [&](int st) -> int {
    const int blk = st * C::BPS + lane;
    return (lane < C::BPS && blk < nblk) ? __ldg(table + blk) : 0;
  };
  {
    int ph[NST];
#pragma unroll
    for (int k = 0; k < NST; ++k) ph[k] = (warp + k * W < nstages) ? load_phys(warp + k * W) : 0;
#pragma unroll
    for (int k = 0; k < NST; ++k)
      if (warp + k * W < nstages) issue_stage(warp + k * W, k, ph[k]);
  }
  int next_phys = (warp + NST * W < nstages) ? load_phys(warp + NST * W) : 0;

  float m[G], lsum[G], acc[G][C::NJ];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    m[g] = -INFINITY;
    lsum[g] = 0.f;
#pragma unroll
    for (int j = 0; j < C::NJ; ++j) acc[g][j] = 0.f;
  }

  int k = 0;
  for (int st = warp; st < nstages; st += W, ++k) {
    const int slot = k % NST;
    mbar_wait(bars + slot * 8, (k / NST) & 1);
    if (APPEND && row_owner && st == nstages - 1 && new_slot >= 0 &&
        new_slot == phys_last * B + (r.tok_end - 1) % B) {
      const int in_stage = (r.tok_end - 1 - r.tok_begin) - st * 32;
      const int b = in_stage / B;
      const int off = (int)(new_slot % B);
      uint8_t* kt = smem + (warp * NST + slot) * C::STAGE + b * C::TILE;
      uint8_t* vt = smem + (warp * NST + slot) * C::STAGE + (C::BPS + b) * C::TILE;
      if (lane < D / 8) *reinterpret_cast<uint4*>(kt + (lane * B + off) * 16) = patch_k;
#pragma unroll
      for (int u = 0; u < (D + 63) / 64; ++u) {
        const int i = lane + 32 * u;
        if (i < D / 2) {
          reinterpret_cast<uint16_t*>(vt)[(2 * i) * B + off] = (uint16_t)(patch_v[u] & 0xffffu);
          reinterpret_cast<uint16_t*>(vt)[(2 * i + 1) * B + off] = (uint16_t)(patch_v[u] >> 16);
        }
      }
      __syncwarp();
    }
    process_stage_gqa<C>(ring + slot * C::STAGE, pbuf, lane, qs, m, lsum, acc, r.tok_begin + st * 32, r.tok_end,
                         p.scale_log2e);
    const int st_next = st + NST * W;
    if (st_next < nstages) {
      issue_stage(st_next, slot, next_phys);
      next_phys = (st_next + W < nstages) ? load_phys(st_next + W) : 0;
    }
  }

  if (APPEND && row_owner && new_slot >= 0) {
    const int64_t blk = new_slot / B;
    const int off = (int)(new_slot % B);
    __half* kdst = p.kc_w + blk * p.kv_block_stride + (int64_t)kv_head * p.kv_head_stride;
    uint16_t* vdst =
        reinterpret_cast<uint16_t*>(p.vc_w + blk * p.kv_block_stride + (int64_t)kv_head * p.kv_head_stride) + off;
    if (lane < D / 8) *reinterpret_cast<uint4*>(kdst + ((int64_t)lane * B + off) * 8) = patch_k;
#pragma unroll
    for (int u = 0; u < (D + 63) / 64; ++u) {
      const int i = lane + 32 * u;
      if (i < D / 2) {
        vdst[(int64_t)(2 * i) * B] = (uint16_t)(patch_v[u] & 0xffffu);
        vdst[(int64_t)(2 * i + 1) * B] = (uint16_t)(patch_v[u] >> 16);
      }
    }
  }

  float* red = reinterpret_cast<float*>(smem + C::RED_OFF);  // [G][W][D + 2]
#pragma unroll
  for (int g = 0; g < G; ++g) warp_publish<C>(red + g * (W * (D + 2)), warp, lane, m[g], lsum[g], acc[g]);
  __syncthreads();
  for (int g = 0; g < G; ++g)
    cta_merge_and_store<C, W>(p, red + g * (W * (D + 2)), threadIdx.x, seq, head0 + g, split, r.nsplit_seq,
                              reinterpret_cast<float*>(smem));
}

// ---------------------------------------------------------------------------------------------
// paged_attention_v2: reduce over the caller's partition records (attention_kernels.cu:564-669)
// ---------------------------------------------------------------------------------------------
// Grid (H, S), D threads' worth of work per CTA: out = sum_j tmp_out[j] * exp_sums[j] * exp(max_logits[j] - M) /
// (sum_j exp_sums[j] * exp(max_logits[j] - M) + 1e-6); a single live partition is copied (reference :580-592); an
// empty sequence yields zeros (the reference's loop over zero partitions does the same).
template <int D>
__global__ void __launch_bounds__(128) pa_v2_reduce_kernel(__half* __restrict__ out, const float* __restrict__ exp_sums,
                                                           const float* __restrict__ max_logits,
                                                           const __half* __restrict__ tmp_out,
                                                           const int32_t* __restrict__ seq_lens, int num_heads,
                                                           int max_num_partitions, int partition_tokens,
                                                           int max_tokens) {
  const int head = blockIdx.x, seq = blockIdx.y;
  griddep_launch_dependents();
  griddep_wait();  // the partition kernel's records
  int len = seq_lens[seq];
  len = len < max_tokens ? len : max_tokens;
  int np = (len + partition_tokens - 1) / partition_tokens;
  np = np < max_num_partitions ? np : max_num_partitions;
  const int64_t row = (int64_t)seq * num_heads + head;
  const float* ml = max_logits + row * max_num_partitions;
  const float* es = exp_sums + row * max_num_partitions;
  const __half* to = tmp_out + row * max_num_partitions * D;
  __half* o = out + row * D;
  if (np == 1) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) o[d] = to[d];
    return;
  }
  float M = -INFINITY;
  for (int j = 0; j < np; ++j) M = fmaxf(M, ld_cg_f32(ml + j));
  float G = 0.f;
  for (int j = 0; j < np; ++j) G += ld_cg_f32(es + j) * __expf(ld_cg_f32(ml + j) - M);
  const float inv = __fdividef(1.f, G + 1e-6f);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < np; ++j)
      acc += __half2float(to[(int64_t)j * D + d]) * (ld_cg_f32(es + j) * __expf(ld_cg_f32(ml + j) - M)) * inv;
    o[d] = __float2half_rn(acc);
  }
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
constexpr int kV2PartitionTokens = 512;  // PARTITION_SIZE of the reference's v2 launcher (attention_kernels.cu:847)

#ifdef VB200_TRACE
static unsigned long long* g_trace = nullptr;
extern "C" void vb200_debug_set_trace(void* buf) { g_trace = static_cast<unsigned long long*>(buf); }
#endif

// CTA shape: W warps x NST ring stages.  The shipped library has ONE shape per head size (default_warps(D) x 2);
// a -DVB200_EXPERIMENTS build adds the other shapes and a forced split factor for sweeps (tools/sweep.py), set
// through vb200_debug_set_tuning -- not declared in the public header, not present in the default build.
struct Shape {
  int W, NST;
};
#ifdef VB200_EXPERIMENTS
struct Tuning {
  int force_splits = 0;
  int variant = 0;
  int max_cluster = 16;  // cap on the cluster size / split factor
  int spec_prefetch = 1; // speculative first-ring L2 prefetch ahead of griddepcontrol.wait (A/B knob)
};
static Tuning make_tuning() {
  Tuning t;
  // VB200_MAX_CLUSTER=8: lets a whole test run stay within the cluster sizes compute-sanitizer's synccheck can follow
  // (profiles/r02f_sanitizer.md); experiments build only
  if (const char* e = getenv("VB200_MAX_CLUSTER")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 16) t.max_cluster = v;
  }
  if (const char* e = getenv("VB200_SPEC_PREFETCH")) t.spec_prefetch = atoi(e) ? 1 : 0;
  return t;
}
static Tuning g_tuning = make_tuning();
// 0: default; 1: 4x3; 2: 8x1; 3: 2x2; 4: default kernel with the grouped-query path off
static constexpr Shape kVariants[] = {{0, 2}, {4, 3}, {8, 1}, {2, 2}, {0, 2}};
constexpr int kNumVariants = 5;
static int tuning_variant() { return g_tuning.variant; }
static int tuning_force_splits() { return g_tuning.force_splits; }
static int tuning_max_cluster() { return g_tuning.max_cluster; }
static int tuning_spec_prefetch() { return g_tuning.spec_prefetch; }
#else
static constexpr int tuning_variant() { return 0; }
static constexpr int tuning_force_splits() { return 0; }
static constexpr int tuning_max_cluster() { return 16; }
static constexpr int tuning_spec_prefetch() { return 1; }
#endif

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel instantiation, device): a lock-free
// test-and-set per device (the attribute call itself is idempotent, so two racing first callers are harmless).
struct DeviceOnce {
  std::atomic<uint64_t> done{0};
  template <class F>
  cudaError_t run(int dev, F&& f) {
    if (dev < 0 || dev >= 64) return f();
    const uint64_t bit = 1ull << dev;
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    const cudaError_t e = f();
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
  }
};

// what a launch did, for vb200_paged_attention_plan (a pure function of the problem shape)
struct Plan {
  const char* kernel = "";
  int D = 0, B = 0, W = 0, NST = 0, G = 1, append = 0;
  unsigned gx = 0, gy = 0, gz = 0;
  int smem = 0;
};

template <int D, int B, int W, int NST, bool APPEND>
static int launch_classic_impl(const PaParams& p, int num_seqs, cudaStream_t stream, Plan* plan) {
  using C = PaCfg<D, B, W, NST>;
  static_assert(C::FITS, "shared memory");
  auto kern = pa_v1_kernel<D, B, W, NST, APPEND>;
  dim3 grid(p.num_heads, num_seqs, p.num_splits);
  if (plan) {
    *plan = Plan{"pa_v1_kernel", D, B, W, NST, 1, APPEND, grid.x, grid.y, grid.z, C::SMEM};
    return VB200_OK;
  }
  static DeviceOnce once;
  int dev = 0;
  VB200_CUDA_CHECK(cudaGetDevice(&dev));
  VB200_CUDA_CHECK(once.run(dev, [&] {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);   // clusters of up to 16
  }));
  PaParams pl = p;
  if (p.spec_ctas) pl.spec_ctas = device_sm_count() * ((227 * 1024) / (C::SMEM + 1024));   // the launch's first wave
  VB200_CUDA_CHECK(launch_kernel_cluster(kern, grid, dim3(W * 32), (size_t)C::SMEM, stream,
                                         pl.cluster_merge ? (unsigned)pl.num_splits : 1u, pl));
  count_launch();
  return VB200_OK;
}

// Largest thread-block cluster (1, 1, z) the device can co-schedule for the default kernel of head size D (its
// shared-memory footprint does not depend on the block size): 16 (non-portable, opt-in), else 8, 4, 2; 1 = none.
// Queried once per device.
template <int D>
static int cluster_cap() {
  constexpr int W = (D == 64 || D == 96 || D == 192) ? 4 : 3;
  using C = PaCfg<D, 16, W, 2>;
  auto kern = pa_v1_kernel<D, 16, W, 2, false>;
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    cudaGetLastError();
    return 1;
  }
  int cap = cached[dev].load(std::memory_order_acquire);
  if (cap) return cap;
  cap = 1;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) == cudaSuccess &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    for (unsigned z = 16; z >= 2 && cap == 1; z >>= 1) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(1, 1, z);
      cfg.blockDim = dim3(W * 32);
      cfg.dynamicSmemBytes = C::SMEM;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 1;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = z;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) cap = (int)z;
    }
  }
  cudaGetLastError();
  cached[dev].store(cap, std::memory_order_release);
  return cap;
}
// How many (1, 1, n) clusters of the default kernel of head size D the device holds at once (co-scheduling inside a
// GPC makes this less than slots / n: 48 clusters of 3 do not fit 148 one-CTA SMs).  Queried once per (device, n).
template <int D>
static int max_active_clusters(int n) {
  constexpr int W = (D == 64 || D == 96 || D == 192) ? 4 : 3;
  using C = PaCfg<D, 16, W, 2>;
  auto kern = pa_v1_kernel<D, 16, W, 2, false>;
  static std::atomic<int> cached[64][17];
  int dev = 0;
  if (n < 2 || n > 16 || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    cudaGetLastError();
    return 1 << 30;   // unknown: do not constrain
  }
  int v = cached[dev][n].load(std::memory_order_acquire);
  if (v) return v - 1;
  int num = 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1, 1, (unsigned)n);
  cfg.blockDim = dim3(W * 32);
  cfg.dynamicSmemBytes = C::SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = (unsigned)n;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
      cudaOccupancyMaxActiveClusters(&num, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    num = (1 << 30) - 1;
  }
  cached[dev][n].store(num + 1, std::memory_order_release);
  return num;
}
static int max_active_clusters_for(int head_size, int n) {
  switch (head_size) {
    case 64: return max_active_clusters<64>(n);
    case 80: return max_active_clusters<80>(n);
    case 96: return max_active_clusters<96>(n);
    case 112: return max_active_clusters<112>(n);
    case 128: return max_active_clusters<128>(n);
    case 192: return max_active_clusters<192>(n);
    default: return max_active_clusters<256>(n);
  }
}

static int cluster_cap_for(int head_size) {
  switch (head_size) {
    case 64: return cluster_cap<64>();
    case 80: return cluster_cap<80>();
    case 96: return cluster_cap<96>();
    case 112: return cluster_cap<112>();
    case 128: return cluster_cap<128>();
    case 192: return cluster_cap<192>();
    default: return cluster_cap<256>();
  }
}

// Default CTA shape per head size: W warps x 2 stages of 128*D bytes, W chosen so that the resident CTAs of an SM
// keep the most bytes in flight within 227 KB (measured, profiles/r01j_head_sizes.md): 4 warps at 64 (3 CTAs/SM),
// 96 (2) and 192 (1); 3 warps at 80 (3), 112 (2), 128 (2) and 256 (1).
constexpr int default_warps(int head_size) {
  return (head_size == 64 || head_size == 96 || head_size == 192) ? 4 : 3;
}

template <int D, int B, int W, int NST>
static int launch_classic(const PaParams& p, int num_seqs, cudaStream_t stream, Plan* plan) {
  if (p.new_k != nullptr) {  // the fused append lives in the default instantiation of each head size
    if constexpr (W == default_warps(D) && NST == 2)
      return launch_classic_impl<D, B, W, 2, true>(p, num_seqs, stream, plan);
  }
  return launch_classic_impl<D, B, W, NST, false>(p, num_seqs, stream, plan);
}

template <int D, int B>
static int launch_variant(const PaParams& p, int num_seqs, int variant, cudaStream_t stream, Plan* plan) {
#ifdef VB200_EXPERIMENTS
  if constexpr (D == 64 || D == 128) {
    switch (variant) {
      case 1: return launch_classic<D, B, 4, 3>(p, num_seqs, stream, plan);
      case 2: return launch_classic<D, B, 8, 1>(p, num_seqs, stream, plan);
      case 3: return launch_classic<D, B, 2, 2>(p, num_seqs, stream, plan);
      default: break;
    }
  }
#endif
  (void)variant;
  return launch_classic<D, B, default_warps(D), 2>(p, num_seqs, stream, plan);
}

// Grouped-query launch: G query heads of one kv head per CTA.  Returns 1 if this shape has no such instantiation.
template <int D, int B, int G, bool APPEND>
static int launch_gqa_impl(const PaParams& p, int num_seqs, cudaStream_t stream, Plan* plan) {
  constexpr int W = GqaShape<D, G>::W, NST = GqaShape<D, G>::NST;
  static_assert(default_warps_c<D>() == default_warps(D), "default shape");
  using C = PaCfgG<D, B, W, NST, G>;
  static_assert(C::FITS, "shared memory");
  auto kern = pa_v1_gqa_kernel<D, B, W, NST, APPEND, G>;
  dim3 grid(p.num_heads / G, num_seqs, p.num_splits);
  if (plan) {
    *plan = Plan{"pa_v1_gqa_kernel", D, B, W, NST, G, APPEND, grid.x, grid.y, grid.z, C::SMEM};
    return VB200_OK;
  }
  static DeviceOnce once;
  int dev = 0;
  VB200_CUDA_CHECK(cudaGetDevice(&dev));
  VB200_CUDA_CHECK(once.run(dev, [&] {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  }));
  VB200_CUDA_CHECK(launch_kernel(kern, grid, dim3(W * 32), (size_t)C::SMEM, stream, p));
  count_launch();
  return VB200_OK;
}
template <int D, int B>
static int launch_gqa(const PaParams& p, int num_seqs, cudaStream_t stream, Plan* plan) {
  const int ratio = p.num_heads / p.num_kv_heads;
  constexpr int NJ = (D * B / 8 + 31) / 32;
  if constexpr (4 * NJ <= 32) {  // accumulators of 4 heads fit in registers
    if (ratio % 4 == 0)
      return p.new_k ? launch_gqa_impl<D, B, 4, true>(p, num_seqs, stream, plan)
                     : launch_gqa_impl<D, B, 4, false>(p, num_seqs, stream, plan);
  }
  if constexpr (2 * NJ <= 32) {
    if (ratio % 2 == 0)
      return p.new_k ? launch_gqa_impl<D, B, 2, true>(p, num_seqs, stream, plan)
                     : launch_gqa_impl<D, B, 2, false>(p, num_seqs, stream, plan);
  }
  return 1;
}
template <int D>
static int launch_gqa_block(const PaParams& p, int num_seqs, int block_size, cudaStream_t stream, Plan* plan) {
  switch (block_size) {
    case 8: return launch_gqa<D, 8>(p, num_seqs, stream, plan);
    case 16: return launch_gqa<D, 16>(p, num_seqs, stream, plan);
    case 32: return launch_gqa<D, 32>(p, num_seqs, stream, plan);
    default: return 1;
  }
}

template <int D>
static int launch_block(const PaParams& p, int num_seqs, int block_size, int variant, cudaStream_t stream,
                        Plan* plan) {
  switch (block_size) {
    case 8: return launch_variant<D, 8>(p, num_seqs, variant, stream, plan);
    case 16: return launch_variant<D, 16>(p, num_seqs, variant, stream, plan);
    case 32: return launch_variant<D, 32>(p, num_seqs, variant, stream, plan);
    default: return set_error(VB200_ERR_BLOCK_SIZE, "Unsupported block size: %d", block_size);
  }
}

static bool head_size_supported(int d) {  // the reference launcher's switch, attention_kernels.cu:738-766
  return d == 64 || d == 80 || d == 96 || d == 112 || d == 128 || d == 192 || d == 256;
}

// Query heads per CTA the launch will use: 4 / 2 through the grouped-query kernel (mirrors launch_gqa), else 1.
// The grouped kernel trades parallelism for traffic: it wins once the per-head grid is several waves deep
// (bandwidth-bound), and loses on small batches where the launch is latency-bound and the extra CTAs of the
// per-head grid are free (profiles/r01k_gqa.md: crossover near 3-4 units per SM).
static int heads_per_cta(int num_seqs, int num_heads, int num_kv_heads, int head_size, int block_size, int variant,
                         int num_sms) {
  if (num_heads == num_kv_heads || variant != 0 || (head_size != 64 && head_size != 128)) return 1;
  if ((int64_t)num_seqs * num_heads < 4 * (int64_t)num_sms) return 1;
  const int ratio = num_heads / num_kv_heads;
  const int nj = (head_size * block_size / 8 + 31) / 32;
  if (ratio % 4 == 0 && 4 * nj <= 32) return 4;
  if (ratio % 2 == 0 && 2 * nj <= 32) return 2;
  return 1;
}

// split_tokens / num_splits for a wanted factor (the tail of choose_split): splits are whole rounds of the CTA's warps
static void choose_split_exact(int max_seq_len, int warps, int want, int* num_splits, int* split_tokens) {
  const int quantum = 32 * warps;
  int ns = want < 1 ? 1 : want;
  int st = max_seq_len > 0 ? (max_seq_len + ns - 1) / ns : quantum;
  st = (st + quantum - 1) / quantum * quantum;
  ns = max_seq_len > 0 ? (max_seq_len + st - 1) / st : 1;
  *num_splits = ns < 1 ? 1 : ns;
  *split_tokens = st;
}

// Resident CTAs per SM of the default pa_v1_kernel of a head size (what shared memory allows; 3 / 3 / 2 / 2 / 2 / 1 / 1).
static int classic_ctas_per_sm(int head_size) {
  auto per = [](int smem) { return (227 * 1024) / (smem + 1024); };
  switch (head_size) {
    case 64: return per(PaCfg<64, 16, default_warps(64), 2>::SMEM);
    case 80: return per(PaCfg<80, 16, default_warps(80), 2>::SMEM);
    case 96: return per(PaCfg<96, 16, default_warps(96), 2>::SMEM);
    case 112: return per(PaCfg<112, 16, default_warps(112), 2>::SMEM);
    case 128: return per(PaCfg<128, 16, default_warps(128), 2>::SMEM);
    case 192: return per(PaCfg<192, 16, default_warps(192), 2>::SMEM);
    default: return per(PaCfg<256, 16, default_warps(256), 2>::SMEM);
  }
}

// split-KV factor.  Measured on B200 (profiles/r01d_small_shapes.md): splitting costs ~2.6 us of prologue /
// merge per extra CTA and only pays while (seq, head) pairs alone leave SMs idle, so split only when there
// is less than one unit per SM, aim for ~2 CTAs per SM, and keep >= 256 tokens (2 stages per warp) per split.
// `group` = query heads per CTA (heads_per_cta): the launch has num_seqs * num_heads / group units.
static void choose_split(int num_seqs, int num_heads, int head_size, int max_seq_len, int variant, int num_sms,
                         int group, int max_splits, int* num_splits, int* split_tokens) {
  const int64_t base = (int64_t)num_seqs * num_heads / group;
  int warps = default_warps(head_size);
#ifdef VB200_EXPERIMENTS
  if (kVariants[variant].W) warps = kVariants[variant].W;
#endif
  (void)variant;
  if (group == 4 && head_size == 128) warps = 4;   // GqaShape<128, 4>
  const int quantum = 32 * warps;  // tokens consumed by one round of the CTA's warps
  int ns = 1;
  if (tuning_force_splits() > 0) {
    ns = tuning_force_splits();
  } else if (base > 0 && base <= num_sms) {
    ns = (int)((2 * (int64_t)num_sms + base - 1) / base);
    // ... but never more CTAs than are resident at once: the clusters of a second wave would start when the first
    // has drained (head sizes >= 96 hold 2 or 1 CTAs per SM; 120 units x 3 splits at head size 96 were 360 CTAs on
    // 296 slots)
    const int by_slots = (int)((int64_t)num_sms * (variant == 0 ? classic_ctas_per_sm(head_size) : 2) / base);
    if (ns > by_slots) ns = by_slots;
    const int max_by_len = max_seq_len / 256;
    if (ns > max_by_len) ns = max_by_len;
    if (ns > 64) ns = 64;
  }
  if (ns > max_splits) ns = max_splits;
  if (ns < 1) ns = 1;
  int st = max_seq_len > 0 ? (max_seq_len + ns - 1) / ns : quantum;
  st = (st + quantum - 1) / quantum * quantum;
  ns = max_seq_len > 0 ? (max_seq_len + st - 1) / st : 1;
  *num_splits = ns < 1 ? 1 : ns;
  *split_tokens = st;
}

}  // namespace vb200

using namespace vb200;

namespace {
struct ExtraArgs {
  // fused append
  const void* new_k = nullptr;
  const void* new_v = nullptr;
  const int64_t* slot_mapping = nullptr;
  int64_t new_k_stride = 0, new_v_stride = 0;
  // paged_attention_v2 scratch
  float* v2_exp_sums = nullptr;
  float* v2_max_logits = nullptr;
  void* v2_tmp_out = nullptr;
  // plan query instead of a launch
  Plan* plan = nullptr;
};
}  // namespace

static int paged_attention_impl(void* out, const void* query, const void* key_cache, const void* value_cache,
                                const int32_t* block_tables, const int32_t* seq_lens, int32_t num_seqs,
                                int32_t num_heads, int32_t num_kv_heads, int32_t head_size, int32_t block_size,
                                int32_t max_num_blocks_per_seq, int32_t max_seq_len, int64_t q_stride,
                                int64_t kv_block_stride, int64_t kv_head_stride, float scale, int32_t device,
                                void* stream, const ExtraArgs& a) {
  const bool planning = a.plan != nullptr;
  const bool v2 = a.v2_exp_sums != nullptr;
  if (num_seqs < 0 || num_heads <= 0 || num_kv_heads <= 0 || max_num_blocks_per_seq < 0)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "negative or zero size argument");
  if (block_size != 8 && block_size != 16 && block_size != 32)
    return set_error(VB200_ERR_BLOCK_SIZE, "Unsupported block size: %d", block_size);
  if (!head_size_supported(head_size)) return set_error(VB200_ERR_HEAD_SIZE, "Unsupported head size: %d", head_size);
  if (num_heads % num_kv_heads != 0)
    return set_error(VB200_ERR_INVALID_ARGUMENT, "num_heads %d not a multiple of num_kv_heads %d", num_heads,
                     num_kv_heads);
  if (num_seqs == 0) return VB200_OK;
  if (num_seqs > 65535) return set_error(VB200_ERR_INVALID_ARGUMENT, "num_seqs %d exceeds grid.y limit", num_seqs);
  if (!planning) {
    if (!out || !query || !key_cache || !value_cache || !block_tables || !seq_lens)
      return set_error(VB200_ERR_INVALID_ARGUMENT, "null pointer argument");
    // 16-byte vector / bulk-copy alignment (the reference's vector loads assume the same)
    if (((uintptr_t)query | (uintptr_t)key_cache | (uintptr_t)value_cache) & 15)
      return set_error(VB200_ERR_INVALID_ARGUMENT, "query/key_cache/value_cache must be 16-byte aligned");
    if ((q_stride % 8) || (kv_block_stride % 8) || (kv_head_stride % 8))
      return set_error(VB200_ERR_INVALID_ARGUMENT, "strides must be multiples of 8 elements (16 bytes)");
    if (kv_head_stride < (int64_t)head_size * block_size)
      return set_error(VB200_ERR_INVALID_ARGUMENT, "kv_head_stride smaller than one (block, head) tile");
  }

  DeviceGuard guard(device, /*enabled=*/!planning);
  if (guard.status != cudaSuccess) return set_error(VB200_ERR_CUDA, "cudaSetDevice(%d): %s", device,
                                                     cudaGetErrorString(guard.status));
  int variant = tuning_variant();
  if (a.new_k != nullptr || v2) variant = 0;  // the fused append and v2 live in the default instantiations
  const int num_sms = planning ? 148 : device_sm_count();
  const int group = v2 ? 1 : heads_per_cta(num_seqs, num_heads, num_kv_heads, head_size, block_size, variant, num_sms);

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
  p.cluster_merge = 0;
  // speculative first-ring prefetch (pa_v1_kernel): pays when CTAs are resident before the previous kernel has
  // finished (PDL) and the K/V they will read is not already in L2; a launch over a few MB gains nothing and pays
  // ~0.08 us for the two extra index loads (profiles/r02j_spec_prefetch.md)
  const int64_t ctx_bound = std::min<int64_t>(max_seq_len, (int64_t)max_num_blocks_per_seq * block_size);
  p.spec_ctas = (pdl_enabled() && tuning_spec_prefetch() &&
                 (int64_t)num_seqs * ctx_bound * num_kv_heads * head_size * 4 >= (int64_t)(4 << 20)) ? 1 : 0;
  int ns = 1, st = 0;
  if (v2) {
    // the reference's fixed 512-token partitions; the partition count is the scratch tensors' stride
    // (max_num_partitions = ceil(max_seq_len / 512), attention_kernels.cu:885)
    st = kV2PartitionTokens;
    ns = (max_seq_len + st - 1) / st;
    if (ns < 1) ns = 1;
    if (ns > 65535) return set_error(VB200_ERR_INVALID_ARGUMENT, "max_seq_len %d exceeds grid.z limit", max_seq_len);
  } else if (group == 1 && max_seq_len > 0) {
    // split-KV, merged inside a thread-block cluster of the splits of a (sequence, head): no scratch, no state
    int cap = planning ? 16 : cluster_cap_for(head_size);
    if (cap > tuning_max_cluster()) cap = tuning_max_cluster();
    choose_split(num_seqs, num_heads, head_size, max_seq_len, variant, num_sms, 1, cap, &ns, &st);
    // every cluster of the launch must be resident at once (a second wave of clusters starts when the first has
    // drained: 48 units x 3 splits at head size 192 took 24.9 us, x 2 splits 16.5): ask the occupancy calculator and
    // fall back to fewer, longer splits.  (A plan query has no device and keeps the static estimate.)
    if (!planning && variant == 0 && tuning_force_splits() == 0) {
      const int64_t units = (int64_t)num_seqs * num_heads;
      while (ns > 1 && max_active_clusters_for(head_size, ns) < units) {
        const int fewer = ns - 1;
        choose_split_exact(max_seq_len, default_warps(head_size), fewer, &ns, &st);
        if (ns > fewer) break;   // (cannot happen: fewer splits never round up to more)
      }
    }
    p.cluster_merge = ns > 1;
  }
  p.num_splits = ns;
  p.split_tokens = st;
  p.new_k = static_cast<const __half*>(a.new_k);
  p.new_v = static_cast<const __half*>(a.new_v);
  p.slot_mapping = a.slot_mapping;
  p.new_k_stride = a.new_k_stride;
  p.new_v_stride = a.new_v_stride;
  p.kc_w = static_cast<__half*>(const_cast<void*>(key_cache));
  p.vc_w = static_cast<__half*>(const_cast<void*>(value_cache));
  p.v2_exp_sums = a.v2_exp_sums;
  p.v2_max_logits = a.v2_max_logits;
  p.v2_tmp_out = static_cast<__half*>(a.v2_tmp_out);
#ifdef VB200_TRACE
  p.trace = g_trace;
#endif
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (group > 1) {
    // grouped-query shapes, batch large enough to be bandwidth-bound: share each landed tile among the group
    int rc = 1;
    if (head_size == 64) rc = launch_gqa_block<64>(p, num_seqs, block_size, s, a.plan);
    if (head_size == 128) rc = launch_gqa_block<128>(p, num_seqs, block_size, s, a.plan);
    if (rc != 1) return rc;
  }
  int rc;
  switch (head_size) {
    case 64: rc = launch_block<64>(p, num_seqs, block_size, variant, s, a.plan); break;
    case 128: rc = launch_block<128>(p, num_seqs, block_size, variant, s, a.plan); break;
    case 80: rc = launch_block<80>(p, num_seqs, block_size, variant, s, a.plan); break;
    case 96: rc = launch_block<96>(p, num_seqs, block_size, variant, s, a.plan); break;
    case 112: rc = launch_block<112>(p, num_seqs, block_size, variant, s, a.plan); break;
    case 192: rc = launch_block<192>(p, num_seqs, block_size, variant, s, a.plan); break;
    default: rc = launch_block<256>(p, num_seqs, block_size, variant, s, a.plan); break;
  }
  if (rc != VB200_OK || !v2 || planning) return rc;

  // ---- v2: reduce the partition records into `out` (second launch; PDL lets it become resident early) ----
  const int max_tokens = max_num_blocks_per_seq * block_size;
  auto reduce = [&](auto kern) -> cudaError_t {
    return launch_kernel(kern, dim3(num_heads, num_seqs), dim3(128), 0, s, static_cast<__half*>(out),
                         static_cast<const float*>(a.v2_exp_sums), static_cast<const float*>(a.v2_max_logits),
                         static_cast<const __half*>(a.v2_tmp_out), seq_lens, (int)num_heads, ns, st, max_tokens);
  };
  cudaError_t e;
  switch (head_size) {
    case 64: e = reduce(pa_v2_reduce_kernel<64>); break;
    case 80: e = reduce(pa_v2_reduce_kernel<80>); break;
    case 96: e = reduce(pa_v2_reduce_kernel<96>); break;
    case 112: e = reduce(pa_v2_reduce_kernel<112>); break;
    case 128: e = reduce(pa_v2_reduce_kernel<128>); break;
    case 192: e = reduce(pa_v2_reduce_kernel<192>); break;
    default: e = reduce(pa_v2_reduce_kernel<256>); break;
  }
  VB200_CUDA_CHECK(e);
  count_launch();
  return VB200_OK;
}

extern "C" int vb200_paged_attention_v1_f16(void* out, const void* query, const void* key_cache,
                                            const void* value_cache, const int32_t* block_tables,
                                            const int32_t* seq_lens, int32_t num_seqs, int32_t num_heads,
                                            int32_t num_kv_heads, int32_t head_size, int32_t block_size,
                                            int32_t max_num_blocks_per_seq, int32_t max_seq_len, int64_t q_stride,
                                            int64_t kv_block_stride, int64_t kv_head_stride, float scale,
                                            int32_t device, void* stream) {
  return paged_attention_impl(out, query, key_cache, value_cache, block_tables, seq_lens, num_seqs, num_heads,
                              num_kv_heads, head_size, block_size, max_num_blocks_per_seq, max_seq_len, q_stride,
                              kv_block_stride, kv_head_stride, scale, device, stream, ExtraArgs{});
}

extern "C" int vb200_paged_attention_v1_append_f16(void* out, const void* query, const void* key, const void* value,
                                                   void* key_cache, void* value_cache, const int64_t* slot_mapping,
                                                   const int32_t* block_tables, const int32_t* seq_lens,
                                                   int32_t num_seqs, int32_t num_heads, int32_t num_kv_heads,
                                                   int32_t head_size, int32_t block_size,
                                                   int32_t max_num_blocks_per_seq, int32_t max_seq_len,
                                                   int64_t q_stride, int64_t key_stride, int64_t value_stride,
                                                   int64_t kv_block_stride, int64_t kv_head_stride, float scale,
                                                   int32_t device, void* stream) {
  if (num_seqs > 0 && (!key || !value || !slot_mapping))
    return set_error(VB200_ERR_INVALID_ARGUMENT, "null pointer argument");
  if ((((uintptr_t)key | (uintptr_t)value) & 15) || (key_stride % 8) || (value_stride % 8))
    return set_error(VB200_ERR_INVALID_ARGUMENT, "key/value rows must be 16-byte aligned");
  ExtraArgs a;
  a.new_k = key;
  a.new_v = value;
  a.slot_mapping = slot_mapping;
  a.new_k_stride = key_stride;
  a.new_v_stride = value_stride;
  return paged_attention_impl(out, query, key_cache, value_cache, block_tables, seq_lens, num_seqs, num_heads,
                              num_kv_heads, head_size, block_size, max_num_blocks_per_seq, max_seq_len, q_stride,
                              kv_block_stride, kv_head_stride, scale, device, stream, a);
}

extern "C" int vb200_paged_attention_v2_f16(void* out, float* exp_sums, float* max_logits, void* tmp_out,
                                            const void* query, const void* key_cache, const void* value_cache,
                                            const int32_t* block_tables, const int32_t* seq_lens, int32_t num_seqs,
                                            int32_t num_heads, int32_t num_kv_heads, int32_t head_size,
                                            int32_t block_size, int32_t max_num_blocks_per_seq, int32_t max_seq_len,
                                            int64_t q_stride, int64_t kv_block_stride, int64_t kv_head_stride,
                                            float scale, int32_t device, void* stream) {
  if (num_seqs > 0 && (!exp_sums || !max_logits || !tmp_out))
    return set_error(VB200_ERR_INVALID_ARGUMENT, "null pointer argument");
  if (max_seq_len <= 0) return set_error(VB200_ERR_INVALID_ARGUMENT, "max_seq_len must be positive");
  ExtraArgs a;
  a.v2_exp_sums = exp_sums;
  a.v2_max_logits = max_logits;
  a.v2_tmp_out = tmp_out;
  return paged_attention_impl(out, query, key_cache, value_cache, block_tables, seq_lens, num_seqs, num_heads,
                              num_kv_heads, head_size, block_size, max_num_blocks_per_seq, max_seq_len, q_stride,
                              kv_block_stride, kv_head_stride, scale, device, stream, a);
}

extern "C" int vb200_paged_attention_plan(char* buf, size_t buf_bytes, int32_t num_seqs, int32_t num_heads,
                                          int32_t num_kv_heads, int32_t head_size, int32_t block_size,
                                          int32_t max_seq_len, int32_t fused_append) {
  if (!buf || buf_bytes == 0) return set_error(VB200_ERR_INVALID_ARGUMENT, "null pointer argument");
  buf[0] = 0;
  Plan plan;
  ExtraArgs a;
  a.plan = &plan;
  static const int dummy = 0;
  if (fused_append) a.new_k = &dummy;   // selects the APPEND instantiation; never dereferenced while planning
  const int rc = paged_attention_impl(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, num_seqs, num_heads,
                                      num_kv_heads, head_size, block_size, 0, max_seq_len, 0, 0, 0, 1.f, -1, nullptr, a);
  if (rc != VB200_OK) return rc;
  if (plan.G > 1)
    snprintf(buf, buf_bytes, "vb200::%s<%d,%d,%d,%d,%d,%d> grid=(%u,%u,%u) block=%d smem=%d", plan.kernel, plan.D,
             plan.B, plan.W, plan.NST, plan.append, plan.G, plan.gx, plan.gy, plan.gz, plan.W * 32, plan.smem);
  else
    snprintf(buf, buf_bytes, "vb200::%s<%d,%d,%d,%d,%d> grid=(%u,%u,%u) block=%d smem=%d", plan.kernel, plan.D,
             plan.B, plan.W, plan.NST, plan.append, plan.gx, plan.gy, plan.gz, plan.W * 32, plan.smem);
  return VB200_OK;
}

#ifdef VB200_EXPERIMENTS
// Sweep knobs of the experiments build (tools/sweep.py); process-global on purpose, never in the shipped library.
extern "C" int32_t vb200_debug_set_tuning(int32_t force_num_splits, int32_t variant) {
  const int32_t prev = (g_tuning.force_splits << 8) | (g_tuning.variant & 0xff);
  g_tuning.force_splits = force_num_splits < 0 ? 0 : force_num_splits;
  g_tuning.variant = (variant < 0 || variant >= kNumVariants) ? 0 : variant;
  return prev;
}
// cap on the cluster size = split factor (2..16); merge_mode is accepted for old sweep scripts and ignored (round 1's
// workspace merge was measured against the cluster merge in profiles/r02c_split_merge.md and removed)
extern "C" void vb200_debug_set_spec_prefetch(int32_t on) { g_tuning.spec_prefetch = on ? 1 : 0; }
extern "C" void vb200_debug_set_merge(int32_t merge_mode, int32_t max_cluster) {
  (void)merge_mode;
  g_tuning.max_cluster = max_cluster < 1 ? 16 : max_cluster;
}
#endif
