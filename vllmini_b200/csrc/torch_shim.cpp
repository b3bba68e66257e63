// Thin torch-facing shim over the C ABI (include/vllmini_b200.h).
//
// The reference's binding layer is a pybind11/torch extension (paged_attention_ext/paged_attention_cuda/
// paged_attention_cuda.cpp:51-61) whose functions take torch::Tensor& positionally.  This file is the same
// layer for the B200 path: it only validates arguments (mirroring the reference's TORCH_CHECKs and
// data_ptr<T>() dtype checks), extracts pointers / strides / the current CUDA stream, and calls the
// extern "C" entry points of libvllmini_b200.so.  No kernel, no arithmetic, no allocation and no state lives
// here (nor below: the library has no workspace -- split-KV merges inside a thread-block cluster -- so, like the
// reference's v1, any number of calls may be in flight on different streams).
// vllmini_b200/ops.py falls back to an equivalent ctypes path when this module is not built; both end in the
// same C ABI and the same kernels.
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include <string>

#include "vllmini_b200.h"

namespace {

void check_rc(int rc) {
  // reference: TORCH_CHECK -> c10::Error -> Python RuntimeError
  TORCH_CHECK(rc == 0, "vllmini_b200: ", vb200_last_error(), " (code ", rc, ")");
}

void paged_attention_v1(at::Tensor& out, at::Tensor& query, at::Tensor& key_cache, at::Tensor& value_cache,
                        int64_t num_kv_heads, double scale, at::Tensor& block_tables, at::Tensor& seq_lens,
                        int64_t block_size, int64_t max_seq_len, const c10::optional<at::Tensor>& alibi_slopes,
                        const std::string& kv_cache_dtype, double kv_scale, const int64_t tp_rank,
                        const int64_t blocksparse_local_blocks, const int64_t blocksparse_vert_stride,
                        const int64_t blocksparse_block_size, const int64_t blocksparse_head_sliding_step) {
  (void)kv_scale; (void)tp_rank; (void)blocksparse_local_blocks; (void)blocksparse_block_size;
  (void)blocksparse_head_sliding_step;
  TORCH_CHECK(query.is_cuda(), "vllmini_b200: query must be a CUDA tensor, got device ", query.device());
  TORCH_CHECK(query.scalar_type() == at::kHalf, "vllmini_b200: unsupported input type of kv cache: ",
              query.scalar_type());
  TORCH_CHECK(kv_cache_dtype == "auto", "vllmini_b200: Unsupported data type of kv cache: ", kv_cache_dtype);
  TORCH_CHECK(!alibi_slopes.has_value(), "vllmini_b200: alibi_slopes is not supported (vllmini passes None)");
  TORCH_CHECK(blocksparse_vert_stride <= 1, "vllmini_b200: block-sparse attention is not supported");
  TORCH_CHECK(out.is_cuda() && key_cache.is_cuda() && value_cache.is_cuda() && block_tables.is_cuda() &&
                  seq_lens.is_cuda(),
              "vllmini_b200: out, key_cache, value_cache, block_tables and seq_lens must be CUDA tensors");
  TORCH_CHECK(block_tables.scalar_type() == at::kInt, "vllmini_b200: expected scalar type Int but found ",
              block_tables.scalar_type(), " (block_tables)");
  TORCH_CHECK(seq_lens.scalar_type() == at::kInt, "vllmini_b200: expected scalar type Int but found ",
              seq_lens.scalar_type(), " (seq_lens)");
  TORCH_CHECK(out.scalar_type() == at::kHalf && key_cache.scalar_type() == at::kHalf &&
                  value_cache.scalar_type() == at::kHalf,
              "vllmini_b200: out, key_cache and value_cache must be float16");
  TORCH_CHECK(query.dim() == 3 && query.stride(2) == 1 && query.stride(1) == query.size(2),
              "vllmini_b200: query must be [num_seqs, num_heads, head_size] with contiguous heads");
  TORCH_CHECK(out.is_contiguous() && block_tables.is_contiguous() && seq_lens.is_contiguous(),
              "vllmini_b200: out, block_tables and seq_lens must be contiguous");
  TORCH_CHECK(out.numel() == query.numel(), "vllmini_b200: out must have num_seqs*num_heads*head_size elements");
  const int num_seqs = (int)query.size(0), num_heads = (int)query.size(1), head_size = (int)query.size(2);
  const int max_blocks = block_tables.dim() == 2 ? (int)block_tables.size(1) : 0;
  const int dev = query.get_device();
  const c10::cuda::CUDAGuard device_guard((c10::DeviceIndex)dev);   // reference: OptionalCUDAGuard(device_of(query))
  cudaStream_t stream = c10::cuda::getCurrentCUDAStream(dev).stream();
  check_rc(vb200_paged_attention_v1_f16(out.data_ptr(), query.data_ptr(), key_cache.data_ptr(),
                                        value_cache.data_ptr(), block_tables.data_ptr<int>(),
                                        seq_lens.data_ptr<int>(), num_seqs, num_heads, (int)num_kv_heads, head_size,
                                        (int)block_size, max_blocks, (int)max_seq_len, query.stride(0),
                                        key_cache.stride(0), key_cache.stride(1), (float)scale, dev, (void*)stream));
}

void reshape_and_cache(at::Tensor& key, at::Tensor& value, at::Tensor& key_cache, at::Tensor& value_cache,
                       at::Tensor& slot_mapping, const std::string& kv_cache_dtype, const double kv_scale) {
  (void)kv_scale;
  TORCH_CHECK(key.is_cuda(), "vllmini_b200: key must be a CUDA tensor, got device ", key.device());
  TORCH_CHECK(kv_cache_dtype == "auto", "vllmini_b200: Unsupported data type of kv cache: ", kv_cache_dtype);
  TORCH_CHECK(key.scalar_type() == at::kHalf, "vllmini_b200: unsupported input type of kv cache: ",
              key.scalar_type());
  TORCH_CHECK(value.is_cuda() && key_cache.is_cuda() && value_cache.is_cuda() && slot_mapping.is_cuda(),
              "vllmini_b200: value, key_cache, value_cache and slot_mapping must be CUDA tensors");
  TORCH_CHECK(slot_mapping.scalar_type() == at::kLong, "vllmini_b200: expected scalar type Long but found ",
              slot_mapping.scalar_type(), " (slot_mapping)");
  TORCH_CHECK(value.scalar_type() == at::kHalf && key_cache.scalar_type() == at::kHalf &&
                  value_cache.scalar_type() == at::kHalf,
              "vllmini_b200: value, key_cache and value_cache must be float16");
  TORCH_CHECK(key.dim() == 3 && key.stride(2) == 1 && key.stride(1) == key.size(2) && value.stride(2) == 1 &&
                  value.stride(1) == value.size(2),
              "vllmini_b200: key/value must be [num_tokens, num_heads, head_size] with contiguous heads");
  TORCH_CHECK(key_cache.dim() == 5 && key_cache.is_contiguous() && value_cache.is_contiguous(),
              "vllmini_b200: key_cache must be contiguous [NB, H, D/x, B, x]; value_cache [NB, H, D, B]");
  TORCH_CHECK(slot_mapping.is_contiguous(), "vllmini_b200: slot_mapping must be contiguous");
  const int dev = key.get_device();
  cudaStream_t stream = c10::cuda::getCurrentCUDAStream(dev).stream();
  check_rc(vb200_reshape_and_cache_f16(key.data_ptr(), value.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(),
                                       slot_mapping.data_ptr<int64_t>(), (int)key.size(0), (int)key.size(1),
                                       (int)key.size(2), (int)key_cache.size(3), (int)key_cache.size(4),
                                       key.stride(0), value.stride(0), dev, (void*)stream));
}

}  // namespace

PYBIND11_MODULE(_torch_shim, m) {
  m.doc() = "torch-facing shim over libvllmini_b200 (C ABI)";
  m.def("paged_attention_v1", &paged_attention_v1, "Paged Attention V1 (sm_100a)");
  m.def("reshape_and_cache", &reshape_and_cache, "Reshape and cache key/value tensors (sm_100a)");
}
