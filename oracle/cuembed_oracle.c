/*
 * cuembed_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the cuEmbed CPU reference for the embedding hot path
 * (forward pool, index transforms, backward).  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load this
 * library; the product (cuembed_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * against (a) the known-answer vectors of the reference's own gtest suites
 * (tests/golden/kat.py, transcribed with file:line) and (b) the reference's CPU
 * templates compiled unchanged from /root/reference into oracle/_ref/ (see
 * oracle/ref_shim.cu, oracle/Makefile) on the randomised shape matrix of
 * tests/test_embedding_against_cpu.cu:236-293.
 *
 * Each function cites the reference lines it restates (paths relative to the
 * reference root).  Low-precision arithmetic is modelled the way cuda_fp16.h /
 * cuda_bf16.h do it on the host: operate in float, round to nearest-even to the
 * storage type after every operation (innocuous double rounding: 24 >= 2p+2).
 *
 * dtype codes: 0 = float32, 1 = float16, 2 = bfloat16.
 * itype codes: 0 = int32,   1 = int64.
 * mode codes : 0 = kSum, 1 = kMean, 2 = kConcat
 *              (cuembed/include/embedding_lookup_types.cuh:29).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- scalars */

static inline uint32_t f2u(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float u2f(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}

/* float -> IEEE binary16 bits, round-to-nearest-even (== __float2half). */
static uint16_t f32_to_f16_bits(float f) {
  uint32_t x = f2u(f);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t abs = x & 0x7fffffffu;
  if (abs >= 0x7f800000u) { /* inf / nan */
    return (uint16_t)(sign | (abs > 0x7f800000u ? 0x7fffu : 0x7c00u));
  }
  if (abs >= 0x477ff000u) { /* >= 65520 rounds to inf */
    return (uint16_t)(sign | 0x7c00u);
  }
  if (abs < 0x33000001u) { /* <= 2^-25 rounds to zero */
    return (uint16_t)sign;
  }
  int32_t exp = (int32_t)(abs >> 23) - 127;
  uint32_t man = (abs & 0x7fffffu) | 0x800000u;
  uint32_t shift;
  uint32_t half_exp;
  if (exp < -14) { /* subnormal half */
    shift = (uint32_t)(13 + (-14 - exp));
    half_exp = 0;
  } else {
    shift = 13;
    half_exp = (uint32_t)(exp + 15);
  }
  uint32_t q = man >> shift;
  uint32_t rem = man & ((1u << shift) - 1u);
  uint32_t halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (q & 1u))) q++;
  uint32_t out;
  if (half_exp == 0) {
    out = q; /* may carry into exponent 1: still correct */
  } else {
    out = ((half_exp - 1) << 10) + q; /* q has the implicit bit at 1<<10 */
  }
  return (uint16_t)(sign | out);
}

static float f16_bits_to_f32(uint16_t h) {
  uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu;
  uint32_t man = h & 0x3ffu;
  if (exp == 0) {
    if (man == 0) return u2f(sign);
    float v = (float)man * 5.9604644775390625e-08f; /* 2^-24 */
    return sign ? -v : v;
  }
  if (exp == 31) return u2f(sign | 0x7f800000u | (man << 13));
  return u2f(sign | ((exp + 112u) << 23) | (man << 13));
}

/* float -> bfloat16 bits, round-to-nearest-even (== __float2bfloat16). */
static uint16_t f32_to_bf16_bits(float f) {
  uint32_t x = f2u(f);
  if ((x & 0x7fffffffu) > 0x7f800000u) return 0x7fffu;
  uint32_t lsb = (x >> 16) & 1u;
  x += 0x7fffu + lsb;
  return (uint16_t)(x >> 16);
}
static float bf16_bits_to_f32(uint16_t b) { return u2f((uint32_t)b << 16); }

static inline float round_to(int dt, float v) {
  if (dt == 1) return f16_bits_to_f32(f32_to_f16_bits(v));
  if (dt == 2) return bf16_bits_to_f32(f32_to_bf16_bits(v));
  return v;
}
static inline float load_elem(const void* p, int dt, int64_t i) {
  if (dt == 1) return f16_bits_to_f32(((const uint16_t*)p)[i]);
  if (dt == 2) return bf16_bits_to_f32(((const uint16_t*)p)[i]);
  return ((const float*)p)[i];
}
static inline void store_elem(void* p, int dt, int64_t i, float v) {
  if (dt == 1)
    ((uint16_t*)p)[i] = f32_to_f16_bits(v);
  else if (dt == 2)
    ((uint16_t*)p)[i] = f32_to_bf16_bits(v);
  else
    ((float*)p)[i] = v;
}
static inline int64_t load_idx(const void* p, int it, int64_t i) {
  return it ? ((const int64_t*)p)[i] : (int64_t)((const int32_t*)p)[i];
}
static inline void store_idx(void* p, int it, int64_t i, int64_t v) {
  if (it)
    ((int64_t*)p)[i] = v;
  else
    ((int32_t*)p)[i] = (int32_t)v;
}
static inline size_t esize(int dt) { return dt == 0 ? 4 : 2; }

/* exported for tests: scalar conversions */
uint16_t oracle_f32_to_f16(float f) { return f32_to_f16_bits(f); }
float oracle_f16_to_f32(uint16_t h) { return f16_bits_to_f32(h); }
uint16_t oracle_f32_to_bf16(float f) { return f32_to_bf16_bits(f); }
float oracle_bf16_to_f32(uint16_t h) { return bf16_bits_to_f32(h); }

/* ---------------------------------------------------------------- forward */

/*
 * Restates EmbeddingForwardCpu, utils/include/embedding_lookup_cpu.hpp:35-94.
 *   loop nest sample -> element -> hot (:57-65); SumT = float unless
 *   fp16_math, then the element type (:59); weight multiply then add, each
 *   rounded in SumT (:73-75); sum: cast once (:80-81); mean: multiply by
 *   (SumT)(1.0f/hotness), empty bag -> sum*0 (:82-90); concat: raw row copy to
 *   [index_start + j] (:69-71).
 * Extension (not in the CPU reference, which CHECK-fails on it at :51):
 *   weighted mean follows the GPU combiner,
 *   cuembed/include/embedding_lookup_ops.cuh:255-289 -- float accumulated
 *   weight, output = sum * (1.0f / accw) in the reduce type, zero vector when
 *   accw == 0.  Pinned only by tests/test_embedding_ops.cu:281-286.
 * Returns 0, or a negative code for the argument checks at :50-56.
 * [sample_begin, sample_end) lets the caller slice the batch across threads
 * (pointer-offset slicing) without changing results.
 */
int oracle_forward(const void* params, int in_dt, int embed_width,
                   int batch_size, int num_hots, const void* indices, int it,
                   const void* offsets, int ot, const void* weights, void* ret,
                   int out_dt, int mode, int fp16_math, int sample_begin,
                   int sample_end) {
  if (weights != NULL && mode == 2) return -1;
  if (!((offsets != NULL && num_hots == 0) ||
        (offsets == NULL && num_hots > 0)))
    return -2;
  if (offsets != NULL && mode == 2) return -3;
  const int sum_dt = fp16_math ? in_dt : 0;
  if (sample_end > batch_size) sample_end = batch_size;
  for (int i = sample_begin; i < sample_end; ++i) {
    int64_t index_start =
        offsets ? load_idx(offsets, ot, i) : (int64_t)i * num_hots;
    int hotness = offsets ? (int)(load_idx(offsets, ot, i + 1) - index_start)
                          : num_hots;
    if (mode == 2) {
      size_t es = esize(in_dt);
      for (int j = 0; j < hotness; ++j) {
        int64_t row = load_idx(indices, it, index_start + j);
        memcpy((char*)ret + (size_t)(index_start + j) * embed_width * es,
               (const char*)params + (size_t)row * embed_width * es,
               (size_t)embed_width * es);
      }
      continue;
    }
    float accw = 0.f;
    if (weights != NULL && mode == 1) {
      for (int j = 0; j < hotness; ++j)
        accw += load_elem(weights, in_dt, index_start + j);
    }
    for (int k = 0; k < embed_width; ++k) {
      float sum = 0.f;
      for (int j = 0; j < hotness; ++j) {
        int64_t row = load_idx(indices, it, index_start + j);
        float w =
            weights ? load_elem(weights, in_dt, index_start + j) : 1.0f;
        float x = load_elem(params, in_dt, row * embed_width + k);
        float prod = round_to(sum_dt, x * w);
        sum = round_to(sum_dt, sum + prod);
      }
      float out;
      if (mode == 0) {
        out = sum;
      } else if (weights != NULL) { /* GPU/TF weighted mean */
        if (accw == 0.f) {
          out = 0.f;
        } else {
          float scale = round_to(sum_dt, 1.0f / accw);
          out = round_to(sum_dt, sum * scale);
        }
      } else if (hotness == 0) {
        out = round_to(sum_dt, sum * 0.0f);
      } else {
        float scale = round_to(sum_dt, 1.0f / (float)hotness);
        out = round_to(sum_dt, sum * scale);
      }
      store_elem(ret, out_dt, (int64_t)i * embed_width + k, out);
    }
  }
  return 0;
}

/* ------------------------------------------------------- index transforms */

/* ExtractRowIdsFromFixedCpu, utils/include/index_transforms_cpu.hpp:35-44 */
void oracle_extract_row_ids_fixed(int batch_size, int num_hots, void* row_ids,
                                  int it) {
  for (int b = 0; b < batch_size; ++b)
    for (int h = 0; h < num_hots; ++h)
      store_idx(row_ids, it, (int64_t)b * num_hots + h, b);
}

/* ExtractRowIdsFromCSRCpu, utils/include/index_transforms_cpu.hpp:46-57 */
void oracle_extract_row_ids_csr(const void* offsets, int ot, int batch_size,
                                void* row_ids, int it) {
  int64_t cnt = 0;
  for (int b = 0; b < batch_size; ++b) {
    int64_t lo = load_idx(offsets, ot, b), hi = load_idx(offsets, ot, b + 1);
    for (int64_t o = lo; o < hi; ++o) store_idx(row_ids, it, cnt++, b);
  }
}

/* ExtractRowIdsForConcatCpu, utils/include/index_transforms_cpu.hpp:59-64 */
void oracle_extract_row_ids_concat(int nnz, void* row_ids, int it) {
  for (int i = 0; i < nnz; ++i) store_idx(row_ids, it, i, i);
}

/* ComputeCompressedGradIndicesCpu, index_transforms_cpu.hpp:66-77 */
void oracle_compressed_grad_indices(const void* indices, int it, int nnz,
                                    void* remapped) {
  int64_t unique_cnt = 0;
  for (int64_t c = 0; c < nnz; ++c) {
    if (c > 0 && load_idx(indices, it, c) != load_idx(indices, it, c - 1))
      unique_cnt++;
    store_idx(remapped, it, c, unique_cnt);
  }
}

typedef struct {
  int64_t idx;
  int64_t sid;
  float wt;
  int64_t pos;
} tuple_t;

static int tuple_cmp(const void* pa, const void* pb) {
  const tuple_t* a = (const tuple_t*)pa;
  const tuple_t* b = (const tuple_t*)pb;
  if (a->idx < b->idx) return -1;
  if (a->idx > b->idx) return 1;
  if (a->sid < b->sid) return -1;
  if (a->sid > b->sid) return 1;
  if (a->wt < b->wt) return -1;
  if (a->wt > b->wt) return 1;
  /* full ties are indistinguishable in the reference's output; keep input
   * order so the result is deterministic. */
  return (a->pos > b->pos) - (a->pos < b->pos);
}

/*
 * TransposeCpu, utils/include/index_transforms_cpu.hpp:86-125: sort the COO
 * triples by (table index, sample id, weight) (:105-115); weights optional.
 * wdt is the weight dtype code.
 */
int oracle_transpose(const void* rows, const void* cols, const void* weights,
                     int wdt, int nnz, int it, void* t_rows, void* t_cols,
                     void* t_weights) {
  if (nnz <= 0) return 0;
  tuple_t* t = (tuple_t*)malloc((size_t)nnz * sizeof(tuple_t));
  if (!t) return -1;
  for (int c = 0; c < nnz; ++c) {
    t[c].idx = load_idx(cols, it, c);
    t[c].sid = load_idx(rows, it, c);
    t[c].wt = weights ? load_elem(weights, wdt, c) : 0.f;
    t[c].pos = c;
  }
  qsort(t, (size_t)nnz, sizeof(tuple_t), tuple_cmp);
  for (int c = 0; c < nnz; ++c) {
    store_idx(t_rows, it, c, t[c].idx);
    store_idx(t_cols, it, c, t[c].sid);
    if (t_weights && weights) {
      /* bit-exact move of the weight payload */
      size_t es = esize(wdt);
      memcpy((char*)t_weights + (size_t)c * es,
             (const char*)weights + (size_t)t[c].pos * es, es);
    }
  }
  free(t);
  return 0;
}

/* --------------------------------------------------------------- backward */

/*
 * Restates EmbeddingBackwardCpu, utils/include/embedding_lookup_cpu.hpp:96-144:
 *   inverse_mapping from run starts of the remapped array (:110-123); optional
 *   memset (:124-129); then for each nz in order,
 *   grad[index] += grad_y[sample] * weight with BOTH operations rounded in
 *   GradT (:131-143).
 * acc_f32 = 1 is NOT the reference: it accumulates in float in nz order and
 * rounds once at the end of each run of equal destination rows (used to state
 * the fp16/bf16 tolerance of the deterministic fp32-accumulating CUDA path,
 * SURVEY.md section 8(c)).  With acc_f32 the destination row is overwritten
 * (or accumulated onto the existing content when skip_grad_init) per run.
 * [nz_begin, nz_end) lets a caller slice the nz range at run boundaries.
 */
int oracle_backward(const void* grad_y, int dt, int embed_width,
                    int num_grad_embedding_rows, int nnz, int it,
                    const void* t_indices, const void* t_sample_ids,
                    const void* t_remapped, const void* t_weights,
                    int skip_grad_init, void* grad_embedding,
                    void* inverse_mapping, int acc_f32, int nz_begin,
                    int nz_end) {
  if (nnz == 0) return 0;
  if (nz_end > nnz) nz_end = nnz;
  if (t_remapped != NULL && nz_begin == 0) {
    if (!grad_embedding || !inverse_mapping) return -1;
    store_idx(inverse_mapping, it, 0, load_idx(t_indices, it, 0));
    int64_t cnt = 1;
    for (int i = 1; i < nnz; ++i) {
      if (load_idx(t_remapped, it, i - 1) != load_idx(t_remapped, it, i)) {
        store_idx(inverse_mapping, it, cnt, load_idx(t_indices, it, i));
        cnt++;
      }
    }
  }
  if (!skip_grad_init && nz_begin == 0) {
    memset(grad_embedding, 0,
           (size_t)num_grad_embedding_rows * (size_t)embed_width * esize(dt));
  }
  const void* idxp = t_remapped ? t_remapped : t_indices;
  if (!acc_f32) {
    for (int nz = nz_begin; nz < nz_end; ++nz) {
      int64_t index = load_idx(idxp, it, nz);
      int64_t sid = load_idx(t_sample_ids, it, nz);
      float w = t_weights ? load_elem(t_weights, dt, nz) : 1.0f;
      for (int e = 0; e < embed_width; ++e) {
        float g = load_elem(grad_y, dt, sid * embed_width + e);
        float prod = round_to(dt, g * w);
        float cur = load_elem(grad_embedding, dt, index * embed_width + e);
        store_elem(grad_embedding, dt, index * embed_width + e,
                   round_to(dt, cur + prod));
      }
    }
    return 0;
  }
  float* acc = (float*)malloc((size_t)embed_width * sizeof(float));
  if (!acc) return -2;
  int nz = nz_begin;
  while (nz < nz_end) {
    int64_t index = load_idx(idxp, it, nz);
    for (int e = 0; e < embed_width; ++e) acc[e] = 0.f;
    int end = nz;
    while (end < nz_end && load_idx(idxp, it, end) == index) {
      int64_t sid = load_idx(t_sample_ids, it, end);
      float w = t_weights ? load_elem(t_weights, dt, end) : 1.0f;
      for (int e = 0; e < embed_width; ++e)
        acc[e] += load_elem(grad_y, dt, sid * embed_width + e) * w;
      ++end;
    }
    for (int e = 0; e < embed_width; ++e)
      store_elem(grad_embedding, dt, index * embed_width + e, acc[e]);
    nz = end;
  }
  free(acc);
  return 0;
}
