/*
 * CPU oracle for vllmini's paged-attention decode path -- plain C restatement.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker*: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load the library built from it (oracle/_build/libpaged_oracle.so).  The
 * product (vllmini_b200/) never links, loads or calls it.
 *
 * Parity pinning: see the header of oracle/paged_oracle.py (golden vectors from
 * the reference's own Python attention and from the reference's own CUDA
 * kernels run on a B200 live in tests/golden/).
 *
 * Each function cites the reference lines it follows; paths are relative to
 * /root/reference/paged_attention_ext/paged_attention_cuda/.
 *
 * Build: gcc -O3 -fopenmp -fPIC -shared -o libpaged_oracle.so paged_oracle.c -lm
 *        (no -march=native: the .so is built in the authoring container and
 *         runs on the GPU box's host CPU; no -ffast-math: rounding matters)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- IEEE binary16 <-> binary32, round-to-nearest-even (what cvt.rn.f16.f32 /
 *      cvt.f32.f16 do: dtype_float16.cuh:86-147) ------------------------------- */
static inline float h2f(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu;
  uint32_t man = h & 0x3ffu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else { /* subnormal: normalise */
      int e = -1;
      do { man <<= 1; e++; } while (!(man & 0x400u));
      man &= 0x3ffu;
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7f800000u | (man << 13);
  } else {
    bits = sign | ((exp + 112u) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

static inline uint16_t f2h(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t ax = x & 0x7fffffffu;
  if (ax >= 0x7f800000u) { /* inf / nan */
    return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x200u : 0u));
  }
  if (ax >= 0x477ff000u) { /* rounds to >= 65520 -> inf */
    return (uint16_t)(sign | 0x7c00u);
  }
  if (ax < 0x33000001u) { /* < 2^-25 (or == 2^-25: ties to even -> 0) */
    return (uint16_t)sign;
  }
  int32_t e = (int32_t)(ax >> 23) - 127;
  uint32_t m = (ax & 0x7fffffu) | 0x800000u; /* 24-bit significand */
  uint32_t shift;
  uint32_t base;
  if (e < -14) { /* subnormal result */
    shift = (uint32_t)(13 + (-14 - e));
    base = 0;
  } else {
    shift = 13;
    base = ((uint32_t)(e + 15) << 10) - 0x400u; /* implicit bit adds 0x400 below */
  }
  uint32_t q = m >> shift;
  uint32_t rem = m & ((1u << shift) - 1u);
  uint32_t half = 1u << (shift - 1);
  if (rem > half || (rem == half && (q & 1u))) q++;
  return (uint16_t)(sign | (base + q));
}

/* fp16 op helpers: one rounding to fp16 after an exact (fp32/double) op, i.e.
 * mul.f16x2 / add.f16x2 per lane (dtype_float16.cuh:150-178, 217-224). */
static inline uint16_t hmul(uint16_t a, uint16_t b) { return f2h(h2f(a) * h2f(b)); }
static inline uint16_t d2h(double s) { /* double -> fp16 with ONE rounding (via round-to-odd fp32) */
  float f = (float)s;
  if ((double)f != s) {
    uint32_t b;
    memcpy(&b, &f, 4);
    if (fabs((double)f) > fabs(s)) b -= 1; /* truncate toward zero */
    b |= 1u;                               /* sticky bit */
    memcpy(&f, &b, 4);
  }
  return f2h(f);
}
static inline uint16_t hadd(uint16_t a, uint16_t b) { return d2h((double)h2f(a) + (double)h2f(b)); }
/* HFMA2 lane: a*b + c with a single rounding to fp16 */
static inline uint16_t hfma(uint16_t a, uint16_t b, uint16_t c) {
  return d2h((double)h2f(a) * (double)h2f(b) + (double)h2f(c));
}

/* ---------------------------------------------------------------------------
 * reshape_and_cache  (cache_kernels.cu:152-207; host entry :256-281)
 * key/value: [T, H, D] with row strides key_stride/value_stride (elements);
 * key_cache [NB, H, D/x, B, x]; value_cache [NB, H, D, B]; slot_mapping int64.
 * ------------------------------------------------------------------------- */
void oracle_reshape_and_cache_f16(const uint16_t* key, const uint16_t* value, uint16_t* key_cache,
                                  uint16_t* value_cache, const int64_t* slot_mapping, int num_tokens,
                                  int num_heads, int head_size, int block_size, int x, int64_t key_stride,
                                  int64_t value_stride) {
  const int n = num_heads * head_size;
  for (int64_t t = 0; t < num_tokens; ++t) {
    const int64_t slot = slot_mapping[t];
    if (slot < 0) continue;                             /* :166-169 */
    const int64_t block_idx = slot / block_size;        /* :172 */
    const int64_t block_offset = slot % block_size;     /* :173 */
    for (int i = 0; i < n; ++i) {                       /* :178-199 */
      const int h = i / head_size, d = i % head_size;
      const int x_idx = d / x, x_off = d % x;
      const int64_t tk = block_idx * num_heads * (head_size / x) * block_size * x +
                         (int64_t)h * (head_size / x) * block_size * x + (int64_t)x_idx * block_size * x +
                         block_offset * x + x_off;
      const int64_t tv = block_idx * num_heads * head_size * block_size + (int64_t)h * head_size * block_size +
                         (int64_t)d * block_size + block_offset;
      key_cache[tk] = key[t * key_stride + i];
      value_cache[tv] = value[t * value_stride + i];
    }
  }
}

/* ---------------------------------------------------------------------------
 * paged_attention_v1  (attention_kernels.cu:90-496 body, :693-767 launcher)
 *
 * mode 0: "exact"  -- double accumulation everywhere (ground truth)
 * mode 1: "fp32"   -- fp32 logits/softmax/PV, the plain restatement used as the
 *                     timed CPU port (fast path, OpenMP over (seq, head))
 * mode 2: "ref-rounding" -- p -> fp16, fp16 products and fp16 partial sums per
 *                     8-token group, fp32 across groups in the reference's
 *                     warp order (see paged_oracle.py for the line citations)
 * out: [S, H, D] contiguous fp16.  K address (phys,h,tok,d) =
 *   phys*kv_block_stride + h*kv_head_stride + (d/8)*B*8 + tok*8 + d%8  (:272-290)
 * V address = phys*kv_block_stride + h*kv_head_stride + d*B + tok       (:402-412)
 * ------------------------------------------------------------------------- */
int oracle_paged_attention_v1_f16(uint16_t* out, const uint16_t* query, const uint16_t* key_cache,
                                  const uint16_t* value_cache, const int32_t* block_tables,
                                  const int32_t* seq_lens, int num_seqs, int num_heads, int num_kv_heads,
                                  int head_size, int block_size, int max_num_blocks_per_seq, int64_t q_stride,
                                  int64_t kv_block_stride, int64_t kv_head_stride, float scale, int mode) {
  if (block_size != 8 && block_size != 16 && block_size != 32) return -2; /* :800-802 */
  if (head_size % 8 != 0 || head_size > 256) return -3;                   /* :763-765 */
  const int D = head_size, B = block_size;
  const int qpk = num_heads / num_kv_heads;                               /* :152 */
  int rc = 0;
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
  for (int s = 0; s < num_seqs; ++s) {
    for (int h = 0; h < num_heads; ++h) {
      const int L = seq_lens[s];
      uint16_t* o = out + ((int64_t)s * num_heads + h) * D;
      if (L <= 0) {
        for (int d = 0; d < D; ++d) o[d] = 0;
        continue;
      }
      const int kvh = h / qpk;                                            /* :153 */
      const int nblk = (L + B - 1) / B;                                   /* :121 */
      const int32_t* table = block_tables + (int64_t)s * max_num_blocks_per_seq;
      const uint16_t* q = query + (int64_t)s * q_stride + (int64_t)h * D; /* :179 */
      float qf[256];
      for (int d = 0; d < D; ++d) qf[d] = h2f(q[d]);
      float* logits = (float*)malloc(sizeof(float) * (size_t)nblk * B);
      if (!logits) { rc = -1; continue; }
      /* ---- logits (:227-308) ---- */
      double mx = -INFINITY;
      for (int b = 0; b < nblk && mode != 0; ++b) {
        const uint16_t* kb = key_cache + (int64_t)table[b] * kv_block_stride + (int64_t)kvh * kv_head_stride;
        for (int t = 0; t < B; ++t) {
          const int tok = b * B + t;
          float val = 0.f;
          if (tok < L) {
            /* Qk_dot, fp32 FMA in the reference's (lane, vector) order:
             * attention_utils.cuh:30-47 with TG = max(32/B,1), VEC = 16/(TG*2) */
            const int TG = 32 / B > 1 ? 32 / B : 1, VEC = 8 / TG, NVEC = D / TG / VEC;
            float lane_sum[4];
            for (int g = 0; g < TG; ++g) {
              float acc[8] = {0};
              for (int j = 0; j < NVEC; ++j)
                for (int c = 0; c < VEC; ++c) {
                  const int d = (g + j * TG) * VEC + c;
                  acc[c] = qf[d] * h2f(kb[(d / 8) * B * 8 + t * 8 + d % 8]) + acc[c]; /* exact product */
                }
              float sacc = acc[0];
              for (int c = 1; c < VEC; ++c) sacc += acc[c];
              lane_sum[g] = sacc;
            }
            for (int w = TG; w > 1; w /= 2)
              for (int g = 0; g < w / 2; ++g) lane_sum[g] = lane_sum[g] + lane_sum[g + w / 2];
            val = scale * lane_sum[0];                                    /* :294 */
            if (val > mx) mx = val;                                       /* :305 */
          }
          logits[tok] = val;                                              /* masked -> 0.f (:303) */
        }
      }
      /* ---- softmax (:332-346) ---- */
      double sum = 0.0;
      float sumf = 0.f;
      for (int t = 0; t < L && mode != 0; ++t) {
        const float e = expf(logits[t] - (float)mx);                      /* __expf (:335) */
        logits[t] = e;
        sumf += e;
      }
      double acc_out[256];
      for (int d = 0; d < D; ++d) acc_out[d] = 0.0;
      if (mode == 0) {
        /* exact: double logits, double softmax, double PV */
        double* pl = (double*)malloc(sizeof(double) * (size_t)L);
        if (!pl) { rc = -1; free(logits); continue; }
        for (int b = 0; b < nblk; ++b) {
          const uint16_t* kb = key_cache + (int64_t)table[b] * kv_block_stride + (int64_t)kvh * kv_head_stride;
          for (int t = 0; t < B && b * B + t < L; ++t) {
            double acc = 0.0;
            for (int d = 0; d < D; ++d) acc += (double)qf[d] * (double)h2f(kb[(d / 8) * B * 8 + t * 8 + d % 8]);
            pl[b * B + t] = acc * (double)scale;
          }
        }
        for (int t = 0; t < L; ++t) if (pl[t] > mx) mx = pl[t];
        for (int t = 0; t < L; ++t) { pl[t] = exp(pl[t] - mx); sum += pl[t]; }
        const double inv = 1.0 / (sum + 1e-6);                            /* :342 */
        for (int b = 0; b < nblk; ++b) {
          const uint16_t* vb = value_cache + (int64_t)table[b] * kv_block_stride + (int64_t)kvh * kv_head_stride;
          for (int t = 0; t < B && b * B + t < L; ++t)
            for (int d = 0; d < D; ++d) acc_out[d] += pl[b * B + t] * inv * (double)h2f(vb[d * B + t]);
        }
        free(pl);
        for (int d = 0; d < D; ++d) o[d] = f2h((float)acc_out[d]);
      } else if (mode == 1) {
        const float inv = 1.f / (sumf + 1e-6f);
        float accf[256];
        for (int d = 0; d < D; ++d) accf[d] = 0.f;
        for (int b = 0; b < nblk; ++b) {
          const uint16_t* vb = value_cache + (int64_t)table[b] * kv_block_stride + (int64_t)kvh * kv_head_stride;
          const int tmax = (L - b * B) < B ? (L - b * B) : B;
          for (int d = 0; d < D; ++d) {
            float a = 0.f;
            for (int t = 0; t < tmax; ++t) a += logits[b * B + t] * h2f(vb[d * B + t]);
            accf[d] += a;
          }
        }
        for (int d = 0; d < D; ++d) o[d] = f2h(accf[d] * inv);
      } else { /* mode 2: reference rounding (:398-481) */
        const float inv = 1.f / (sumf + 1e-6f);
        const int VPR = B / 8; /* NUM_V_VECS_PER_ROW with V_VEC_SIZE = 8 */
        float accw[4][4][256];  /* [warp][vec-in-row][d] */
        memset(accw, 0, sizeof(accw));
        for (int b = 0; b < nblk; ++b) {
          const int w = b % 4;                                            /* :380-381 */
          const uint16_t* vb = value_cache + (int64_t)table[b] * kv_block_stride + (int64_t)kvh * kv_head_stride;
          for (int g = 0; g < VPR; ++g) {
            uint16_t p16[8];
            for (int j = 0; j < 8; ++j) {
              const int tok = b * B + g * 8 + j;
              p16[j] = f2h(tok < L ? logits[tok] * inv : 0.f);            /* :344, :398-400 */
            }
            for (int d = 0; d < D; ++d) {
              /* Source: mul.f16x2 products + three add.f16x2 (sum(uint4), dtype_float16.cuh:
               * 451-457).  As compiled (nvcc 12.9 / sm_100) ptxas contracts them into
               *   c = rn(p1*v1); c = fma(p0,v0,c); c = fma(p2,v2,c); c = fma(p3,v3,c)
               * per half2 lane (words x,y,z,w = 0..3); this reproduces the golden outputs of the
               * reference kernel bit for bit (tests/golden/refkernel_outputs.npz). */
              uint16_t vv[8];
              for (int j = 0; j < 8; ++j) {
                const int tok = b * B + g * 8 + j;
                vv[j] = tok < L ? vb[d * B + g * 8 + j] : (uint16_t)0;    /* :420-430 */
              }
              uint16_t lo = hmul(p16[2], vv[2]);
              lo = hfma(p16[0], vv[0], lo); lo = hfma(p16[4], vv[4], lo); lo = hfma(p16[6], vv[6], lo);
              uint16_t hi = hmul(p16[3], vv[3]);
              hi = hfma(p16[1], vv[1], hi); hi = hfma(p16[5], vv[5], hi); hi = hfma(p16[7], vv[7], hi);
              accw[w][g][d] += h2f(lo) + h2f(hi);
            }
          }
        }
        for (int d = 0; d < D; ++d) {
          float a[4];
          for (int w = 0; w < 4; ++w) {
            float l[4];
            for (int g = 0; g < VPR; ++g) l[g] = accw[w][g][d];
            for (int wd = VPR; wd > 1; wd /= 2)
              for (int g = 0; g < wd / 2; ++g) l[g] = l[g] + l[g + wd / 2];
            a[w] = l[0];
          }
          a[0] += a[2]; a[1] += a[3]; a[0] += a[1];                       /* :452-481 */
          o[d] = f2h(a[0]);
        }
      }
      free(logits);
    }
  }
  return rc;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* exposed for the conversion self-tests */
uint16_t oracle_f2h(float f) { return f2h(f); }
float oracle_h2f(uint16_t h) { return h2f(h); }
