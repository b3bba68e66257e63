/*
 * vllmini_b200 -- C ABI of the B200-native paged-attention decode path.
 *
 * This is the drop-in boundary for the ONE hot path of MDK8888/vllmini that this
 * project replaces: the two native entry points its Python host calls,
 *
 *     paged_attention_cuda.paged_attention_v1            (vllmini/model/gpt2.py:94)
 *     paged_attention_cuda.cache_ops.reshape_and_cache   (vllmini/model/gpt2.py:81)
 *
 * which the reference implements as a torch/pybind11 extension
 * (paged_attention_ext/paged_attention_cuda/paged_attention_cuda.cpp:7-25, :56-59;
 *  cache_kernels.h:11-14).  Here the same two operations are exported from a plain
 * shared library (libvllmini_b200.so): extern "C", raw device pointers, sizes and
 * strides in ELEMENTS, a CUDA stream handle passed as void*.  No torch types, no
 * exceptions, no allocation, no synchronisation: every call only enqueues work on
 * `stream` and returns.  The Python shim that reproduces the reference's positional
 * signatures on top of this ABI is vllmini_b200/dropin/paged_attention_cuda/.
 *
 * Return value: 0 on success, a negative VB200_ERR_* code otherwise;
 * vb200_last_error() returns a thread-local human-readable message for the last
 * failing call on the calling thread.
 *
 * Data layout (owned by the caller, vllmini/kv_cache.py:13-14):
 *   key_cache   [num_blocks, num_kv_heads, head_size/8, block_size, 8]  fp16
 *   value_cache [num_blocks, num_kv_heads, head_size,   block_size]     fp16
 *   block_tables [num_seqs, max_num_blocks_per_seq] int32, seq_lens [num_seqs] int32
 *   slot_mapping [num_tokens] int64 (negative = padding, skipped)
 */
#ifndef VLLMINI_B200_H_
#define VLLMINI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB200_ABI_VERSION 2

enum {
  VB200_OK = 0,
  VB200_ERR_CUDA = -1,              /* a CUDA runtime call failed (message has the cudaError string) */
  VB200_ERR_BLOCK_SIZE = -2,        /* reference: TORCH_CHECK "Unsupported block size" (attention_kernels.cu:800-802) */
  VB200_ERR_HEAD_SIZE = -3,         /* reference: TORCH_CHECK "Unsupported head size" (attention_kernels.cu:763-765) */
  VB200_ERR_INVALID_ARGUMENT = -4,  /* null pointer, negative size, misaligned pointer/stride */
  VB200_ERR_UNSUPPORTED = -5        /* a feature of the reference signature this build does not implement */
};

/*
 * paged_attention_v1, fp16 query/cache ("auto" kv-cache dtype), no ALiBi, no block-sparsity:
 * the one arm of the reference dispatch vllmini reaches.
 * Replaces: paged_attention_v1(...) host entry, attention_kernels.cu:805-826, its launcher
 * :693-767 and kernel :86-527.
 *
 *   head_size  64, 80, 96, 112, 128, 192 or 256 (the launcher's switch, :738-766); block_size 8, 16 or 32
 *                (:789-803); anything else returns VB200_ERR_HEAD_SIZE / VB200_ERR_BLOCK_SIZE
 *   out        [num_seqs, num_heads, head_size] fp16, contiguous (written)
 *   query      [num_seqs, num_heads, head_size] fp16, row stride q_stride elements
 *   kv_block_stride = key_cache.stride(0), kv_head_stride = key_cache.stride(1); the same two
 *                strides address value_cache (reference: attention_kernels.cu:402-403)
 *   scale      softmax scale (head_size ** -0.5 in vllmini, gpt2.py:13)
 *   max_seq_len upper bound on seq_lens[] (vllmini passes capacity, scheduler.py:97); used only
 *                to pick the split-KV factor -- a seq_lens[] entry above it is still attended in
 *                full (the last split runs to seq_len), so results never depend on the split
 *   device     CUDA device ordinal the pointers live on (reference: OptionalCUDAGuard, :736)
 *   stream     cudaStream_t (reference: at::cuda::getCurrentCUDAStream(), :737)
 *
 * seq_lens[s] == 0 yields zeros for that row, as in the reference.  Entries of block_tables
 * at or beyond ceil(seq_len/block_size) are never read (attention_kernels.cu:227).
 *
 * Like the reference's v1 (SURVEY.md 8b "ownership"), the call needs no workspace and keeps no
 * state: when the batch alone leaves SMs idle (num_seqs * num_heads <= #SMs, long max_seq_len)
 * the context is split over up to 16 CTAs that form one thread-block cluster per (sequence,
 * head) and merge their partial results through distributed shared memory inside the launch.
 * Any number of calls may run concurrently on different streams.
 */
int vb200_paged_attention_v1_f16(void* out, const void* query, const void* key_cache, const void* value_cache,
                                 const int32_t* block_tables, const int32_t* seq_lens, int32_t num_seqs,
                                 int32_t num_heads, int32_t num_kv_heads, int32_t head_size, int32_t block_size,
                                 int32_t max_num_blocks_per_seq, int32_t max_seq_len, int64_t q_stride,
                                 int64_t kv_block_stride, int64_t kv_head_stride, float scale, int32_t device,
                                 void* stream);

/*
 * Fused decode step of one layer (no reference counterpart -- the reference
 * always issues the two calls separately, gpt2.py:44,62): exactly reshape_and_cache(key, value, ...,
 * slot_mapping) followed by paged_attention_v1(...), in ONE launch.  key / value [num_seqs, num_kv_heads,
 * head_size]: one new token per sequence (decode), row strides key_stride / value_stride.  The token of
 * sequence s must sit at position >= seq_lens[s] - 1 of that sequence (seq_lens[s] - 1 if it attends to
 * itself; seq_lens[s] with the reference scheduler's convention).  slot_mapping[s] < 0 skips the write.
 */
int vb200_paged_attention_v1_append_f16(void* out, const void* query, const void* key, const void* value,
                                        void* key_cache, void* value_cache, const int64_t* slot_mapping,
                                        const int32_t* block_tables, const int32_t* seq_lens, int32_t num_seqs,
                                        int32_t num_heads, int32_t num_kv_heads, int32_t head_size,
                                        int32_t block_size, int32_t max_num_blocks_per_seq, int32_t max_seq_len,
                                        int64_t q_stride, int64_t key_stride, int64_t value_stride,
                                        int64_t kv_block_stride, int64_t kv_head_stride, float scale,
                                        int32_t device, void* stream);

/*
 * paged_attention_v2, fp16: the reference's partitioned entry (paged_attention_cuda.cpp:27-48; host entry
 * attention_kernels.cu:966-990, launcher :844-931, partition kernel :529-562, reduce kernel :564-669).
 * Same result as v1, through the CALLER's scratch: the context is cut into the reference's fixed 512-token
 * partitions, P = ceil(max_seq_len / 512) of them per (sequence, head);
 *
 *   exp_sums, max_logits [num_seqs, num_heads, P] fp32, tmp_out [num_seqs, num_heads, P, head_size] fp16
 *
 * receive, for every partition that holds at least one token of its sequence, exactly what the reference's
 * partition kernel writes (:349-357, :484-495): the partition's max logit (scale * q.k, natural-log units), the
 * sum of exp(logit - max) and the partition's own normalised softmax(QK^T)V; partitions beyond a sequence's
 * length are left untouched, as in the reference (:119-122).  A second launch reduces the records into `out`
 * (:637-668; one live partition is copied, :580-592).  No other scratch, no state.
 */
int vb200_paged_attention_v2_f16(void* out, float* exp_sums, float* max_logits, void* tmp_out, const void* query,
                                 const void* key_cache, const void* value_cache, const int32_t* block_tables,
                                 const int32_t* seq_lens, int32_t num_seqs, int32_t num_heads, int32_t num_kv_heads,
                                 int32_t head_size, int32_t block_size, int32_t max_num_blocks_per_seq,
                                 int32_t max_seq_len, int64_t q_stride, int64_t kv_block_stride,
                                 int64_t kv_head_stride, float scale, int32_t device, void* stream);

/*
 * Which kernel instantiation and grid a v1 / fused-append call of this shape launches on a 148-SM B200, as text:
 * "vb200::pa_v1_kernel<64,16,4,2,0> grid=(12,256,1) block=128 smem=67088" (template arguments: head size, block
 * size, warps, ring stages, fused append[, query heads per CTA]).  A pure function of its arguments -- touches no
 * device; bench.py labels its roofline record with it.  (No reference counterpart.)
 */
int vb200_paged_attention_plan(char* buf, size_t buf_bytes, int32_t num_seqs, int32_t num_heads,
                               int32_t num_kv_heads, int32_t head_size, int32_t block_size, int32_t max_seq_len,
                               int32_t fused_append);

/*
 * reshape_and_cache, fp16 ("auto"): scatter new K/V rows into the paged caches by slot.
 * Replaces: reshape_and_cache(...) host entry cache_kernels.cu:256-281 and kernel :152-207.
 *
 *   key, value [num_tokens, num_heads, head_size] fp16, row strides key_stride / value_stride
 *   x          key_cache.size(4) (8 for fp16)
 * Bit-exact copy.  slot_mapping[t] < 0 skips token t (cache_kernels.cu:166-169).
 */
int vb200_reshape_and_cache_f16(const void* key, const void* value, void* key_cache, void* value_cache,
                                const int64_t* slot_mapping, int32_t num_tokens, int32_t num_heads,
                                int32_t head_size, int32_t block_size, int32_t x, int64_t key_stride,
                                int64_t value_stride, int32_t device, void* stream);

/*
 * The rest of the reference's cache_ops surface (exported by paged_attention_cuda.cpp:57-60, not called by
 * vllmini; pure copies, bit-exact):
 *
 * copy_blocks: for every layer and every (src, dst) pair of block_mapping [num_pairs, 2] (device, int64)
 *   copy block src over block dst in that layer's key and value cache.  key_cache_ptrs / value_cache_ptrs
 *   are DEVICE arrays of num_layers cache base addresses.  Replaces cache_kernels.cu:67-148.
 * swap_blocks: one async copy per (src, dst) pair of block_mapping_host (HOST, int64) between two cache
 *   tensors; kind 0 = device->device, 1 = device->host, 2 = host->device.  Replaces cache_kernels.cu:24-63.
 * reshape_and_cache_flash: scatter rows into k_cache / v_cache laid out [num_blocks, block_size, num_heads,
 *   head_size] (block_stride = k_cache.stride(0)).  Replaces cache_kernels.cu:209-240, 283-317.
 */
int vb200_copy_blocks(const int64_t* key_cache_ptrs, const int64_t* value_cache_ptrs, const int64_t* block_mapping,
                      int32_t num_layers, int32_t num_pairs, int64_t block_bytes, int32_t all_16b_aligned,
                      int32_t device, void* stream);
int vb200_swap_blocks(const void* src, void* dst, const int64_t* block_mapping_host, int64_t num_pairs,
                      int64_t block_bytes, int32_t kind, int32_t device, void* stream);
int vb200_reshape_and_cache_flash_f16(const void* key, const void* value, void* k_cache, void* v_cache,
                                      const int64_t* slot_mapping, int32_t num_tokens, int32_t num_heads,
                                      int32_t head_size, int32_t block_size, int64_t block_stride, int64_t key_stride,
                                      int64_t value_stride, int32_t device, void* stream);

/*
 * Residual add + LayerNorm for the batched decode driver (vllmini_b200/decode.py; not part of the reference's
 * extension -- in vllmini these are eager torch ops between the two hot-path calls, gpt2.py:150-160, 258-262):
 *
 *   s          = fp16(a[ar] + x[xr])             ar = a_rows ? a_rows[r] : r,  xr = x_rows ? x_rows[r] : r
 *   x_out[r]   = s                                [num_rows, hidden] contiguous; may alias x when x_rows is NULL
 *   ln_out[r]  = fp16((s - mean(s)) * rsqrt(var(s) + eps) * gamma + beta)     statistics in fp32
 *
 * a, x: fp16 tables of a_num_rows / x_num_rows rows of `hidden` elements, row strides a_stride / x_stride
 * (elements); hidden % 8 == 0, <= 2048.  With a = wte, a_rows = token ids, x = wpe, x_rows = positions it is the
 * embedding step; a row index outside [0, a_num_rows) / [0, x_num_rows) traps the kernel (the eager embedding
 * lookup it replaces raises a device assert) instead of reading out of bounds.
 */
int vb200_add_layernorm_f16(void* x_out, void* ln_out, const void* a, const int64_t* a_rows, int64_t a_num_rows,
                            int64_t a_stride, const void* x, const int64_t* x_rows, int64_t x_num_rows,
                            int64_t x_stride, const void* gamma, const void* beta, int32_t num_rows, int32_t hidden,
                            float eps, int32_t device, void* stream);

/* Greedy sampling for the decode driver: out[r] = argmax over x[r, 0..n) (fp16 rows, row_stride elements apart);
 * first index among equal maxima, NaN counts as the maximum (torch.argmax semantics). */
int vb200_argmax_rows_f16(int64_t* out, const void* x, int64_t row_stride, int32_t num_rows, int32_t n,
                          int32_t device, void* stream);

/* Thread-local message of the last failing call on this thread ("" if none). */
const char* vb200_last_error(void);

/* VB200_ABI_VERSION the library was built with. */
int32_t vb200_abi_version(void);

/* Number of kernel launches this library has enqueued since load (all threads); used by
 * bench.py to report gpu_launches and by tests to prove the CUDA path ran. */
uint64_t vb200_launch_count(void);

/* Programmatic dependent launch (default on): kernels are launched so that their CTAs may become resident
 * while the previous kernel on the stream drains, and block (griddepcontrol.wait) until it has completed
 * before reading or writing global memory -- stream-order semantics are unchanged.  Returns the previous
 * setting.  Turn off if a tool or driver in the stack mishandles programmatic launch attributes.
 *
 * With it on, vb200_paged_attention_v1* / v2 launches over more than 4 MB of K/V also issue L2 PREFETCH HINTS ahead of
 * that wait: the first wave of CTAs reads seq_lens / block_tables entries as they are at that moment (L2-only loads) and
 * hints the K/V tiles they name.  Hints have no architectural effect -- stale or uninitialised indices make them
 * useless, not wrong; every value is read again after the wait.  Both index tensors only have to be valid allocations
 * when the call is enqueued.  Off together with programmatic launch. */
int32_t vb200_set_pdl(int32_t enable);

#ifdef __cplusplus
}
#endif
#endif /* VLLMINI_B200_H_ */
