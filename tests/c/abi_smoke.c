/* Plain-C consumer of include/vllmini_b200.h: proves the header is valid C (no C++, no torch types) and that the
 * argument validation of the ABI works without a GPU.  Built and run by tests/test_abi_cpu.py with gcc. */
#include <stdio.h>
#include <string.h>

#include "vllmini_b200.h"

static int fails = 0;
#define CHECK(cond)                                                  \
  do {                                                               \
    if (!(cond)) {                                                   \
      printf("FAIL line %d: %s\n", __LINE__, #cond);                 \
      ++fails;                                                       \
    }                                                                \
  } while (0)

int main(void) {
  char one[64] = {0};
  CHECK(vb200_abi_version() == VB200_ABI_VERSION);
  /* unsupported block size / head size are rejected before any CUDA call (reference: TORCH_CHECK) */
  int rc = vb200_paged_attention_v1_f16(one, one, one, one, (const int32_t*)one, (const int32_t*)one, 1, 12, 12, 64, 7,
                                        4, 64, 768, 12288, 1024, 0.125f, 0, NULL);
  CHECK(rc == VB200_ERR_BLOCK_SIZE);
  CHECK(strstr(vb200_last_error(), "Unsupported block size: 7") != NULL);
  rc = vb200_paged_attention_v1_f16(one, one, one, one, (const int32_t*)one, (const int32_t*)one, 1, 12, 12, 72, 16, 4,
                                    64, 864, 13824, 1152, 0.125f, 0, NULL);
  CHECK(rc == VB200_ERR_HEAD_SIZE);
  rc = vb200_paged_attention_v1_f16(NULL, one, one, one, (const int32_t*)one, (const int32_t*)one, 1, 12, 12, 64, 16,
                                    4, 64, 768, 12288, 1024, 0.125f, 0, NULL);
  CHECK(rc == VB200_ERR_INVALID_ARGUMENT);
  /* zero sequences / tokens: nothing to do, no device needed */
  CHECK(vb200_paged_attention_v1_f16(one, one, one, one, (const int32_t*)one, (const int32_t*)one, 0, 12, 12, 64, 16, 4,
                                     64, 768, 12288, 1024, 0.125f, 0, NULL) == VB200_OK);
  CHECK(vb200_reshape_and_cache_f16(one, one, one, one, (const int64_t*)one, 0, 12, 64, 16, 8, 768, 768, 0, NULL) ==
        VB200_OK);
  CHECK(vb200_reshape_and_cache_f16(one, one, one, one, (const int64_t*)one, 1, 12, 60, 16, 8, 768, 768, 0, NULL) ==
        VB200_ERR_INVALID_ARGUMENT); /* head_size not a multiple of x */
  CHECK(vb200_add_layernorm_f16(one, one, one, NULL, 1, 768, one, NULL, 1, 768, one, one, 1, 770, 1e-5f, 0, NULL) ==
        VB200_ERR_UNSUPPORTED);
  CHECK(vb200_add_layernorm_f16(one, one, one, NULL, 1, 768, one, NULL, 1, 768, one, one, 2, 768, 1e-5f, 0, NULL) ==
        VB200_ERR_INVALID_ARGUMENT); /* two rows asked of one-row tables without index arrays */
  CHECK(vb200_argmax_rows_f16(NULL, one, 8, 1, 8, 0, NULL) == VB200_ERR_INVALID_ARGUMENT);
  /* v2: same validation as v1, plus its scratch pointers */
  CHECK(vb200_paged_attention_v2_f16(one, (float*)one, (float*)one, one, one, one, one, (const int32_t*)one,
                                     (const int32_t*)one, 1, 12, 12, 64, 7, 4, 64, 768, 12288, 1024, 0.125f, 0,
                                     NULL) == VB200_ERR_BLOCK_SIZE);
  CHECK(vb200_paged_attention_v2_f16(one, NULL, (float*)one, one, one, one, one, (const int32_t*)one,
                                     (const int32_t*)one, 1, 12, 12, 64, 16, 4, 64, 768, 12288, 1024, 0.125f, 0,
                                     NULL) == VB200_ERR_INVALID_ARGUMENT);
  {
    char plan[128];
    CHECK(vb200_paged_attention_plan(plan, sizeof plan, 256, 12, 12, 64, 16, 1024, 0) == VB200_OK);
    CHECK(strstr(plan, "pa_v1_kernel<64,16,4,2,0> grid=(12,256,1)") != NULL);
  }
  CHECK(vb200_launch_count() == 0); /* nothing above may have launched a kernel */
  printf(fails ? "abi_smoke: %d failure(s)\n" : "abi_smoke: ok\n", fails);
  return fails ? 1 : 0;
}
