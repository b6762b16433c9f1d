/*
 * cuembed_b200.h -- C ABI of the B200-native embedding-lookup library.
 *
 * This is the drop-in boundary for the cuEmbed hot path (forward pool ->
 * index transpose -> backward).  The reference exposes that path as
 * header-only C++ templates in namespace cuembed; each entry point below is
 * the type-erased launcher behind one of those templates (file:line refer to
 * the reference tree, NVIDIA/cuEmbed @ 90dd8436):
 *
 *   cuembed_forward                     EmbeddingForward
 *                                       cuembed/include/embedding_lookup.cuh:245-308
 *   cuembed_extract_row_ids_fixed       ExtractRowIdsFromFixed
 *                                       cuembed/include/index_transforms.cuh:45-55
 *   cuembed_extract_row_ids_csr         ExtractRowIdsFromCSR     :66-74
 *   cuembed_extract_row_ids_concat      ExtractRowIdsForConcat   :85-93
 *   cuembed_transpose                   Transpose                :224-250
 *   cuembed_compressed_grad_indices     ComputeCompressedGradIndices :278-323
 *   cuembed_backward                    EmbeddingBackward
 *                                       cuembed/include/embedding_lookup.cuh:423-483
 *
 * The C++ templates with the reference's exact signatures live in
 * include/cuembed/include/{embedding_lookup,index_transforms}.cuh and forward
 * to these functions; INTEGRATION.md shows both bindings.
 *
 * Conventions (same as the reference, cuembed/README.md):
 *   - every pointer is caller-owned DEVICE (or managed) memory unless named
 *     lwork; nothing here synchronises; all work is enqueued on `stream`;
 *   - scratch memory is caller-provided through the work/lwork two-call query
 *     (call with work == NULL to get the size in *lwork);
 *   - functions return CUEMBED_OK or a negative error code instead of
 *     aborting; the C++ templates turn a non-zero code into the reference's
 *     "Check failed ... abort()" behaviour.
 * There is no CPU fallback: every call launches sm_100a kernels.
 */
#ifndef CUEMBED_B200_H_
#define CUEMBED_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cudaStream_t without dragging cuda_runtime.h into C callers. */
typedef struct CUstream_st* cuembed_stream_t;

/* element dtype codes */
#define CUEMBED_F32 0
#define CUEMBED_F16 1
#define CUEMBED_BF16 2
/* index / offset type codes */
#define CUEMBED_I32 0
#define CUEMBED_I64 1
/* combine modes, same order as cuembed::CombineMode
 * (cuembed/include/embedding_lookup_types.cuh:29) */
#define CUEMBED_SUM 0
#define CUEMBED_MEAN 1
#define CUEMBED_CONCAT 2

/* error codes */
#define CUEMBED_OK 0
#define CUEMBED_ERR_WEIGHTED_CONCAT (-1) /* embedding_lookup.cuh:260-261 */
#define CUEMBED_ERR_CSR_XOR_FIXED (-2)   /* embedding_lookup.cuh:263-265 */
#define CUEMBED_ERR_CSR_CONCAT (-3)      /* embedding_lookup.cuh:266-267 */
#define CUEMBED_ERR_ROW_BYTES (-4)       /* row bytes % 4, :163 */
#define CUEMBED_ERR_DTYPE (-5)           /* unsupported dtype combination */
#define CUEMBED_ERR_WORKSPACE (-6)       /* *lwork too small, index_transforms.cuh:126 */
#define CUEMBED_ERR_ARGUMENT (-7)        /* null / negative argument */
#define CUEMBED_ERR_CUDA (-8)            /* a CUDA runtime call failed */
#define CUEMBED_ERR_NNZ_LIMIT (-9)       /* nnz >= 2^30 in transpose */
#define CUEMBED_ERR_INDEX_RANGE (-10)    /* debug check: an index outside [0, rows) */
#define CUEMBED_ERR_OFFSETS (-11)        /* debug check: offsets not ascending / out of range */

/* Library / ABI version and the SM architecture the kernels were built for. */
int cuembed_version(void);
const char* cuembed_build_arch(void);
const char* cuembed_error_string(int code);

/*
 * Pooled embedding lookup.  params [rows, embed_width] row-major of in_dtype;
 * indices [nnz] of idx_type; offsets [batch_size + 1] of off_type (CSR) or
 * NULL with num_hots > 0 (fixed hotness); weights [nnz] of in_dtype or NULL;
 * ret [batch_size, embed_width] of out_dtype (sum / mean) or
 * [nnz, embed_width] of in_dtype (concat).  fp16_math != 0 accumulates in the
 * input type (meaningful for F16/BF16 only).  Accumulation is sequential in
 * bag order per output element, so fp32 results are bit-identical to the
 * reference's CPU implementation.
 */
int cuembed_forward(const void* params, int in_dtype, int embed_width,
                    const void* indices, int idx_type, const void* offsets,
                    int off_type, const void* weights, int batch_size,
                    int num_hots, int mode, int fp16_math, void* ret,
                    int out_dtype, cuembed_stream_t stream);

/*
 * Pooled lookup through an addresser indirection (new): the hook the reference
 * reserves for an embedding cache ("sample_id directly corresponds to the
 * physical row ... this may change with an index mapping or an embedding
 * cache", cuembed/include/embedding_lookup_ops.cuh:62-63; "templatize the
 * addresser with the cache", embedding_lookup_kernels.cuh:114-115).
 * row_map [rows of params] of idx_type:
 *   row_map[i] >= 0 : lookup i reads row row_map[i] of cache_params
 *                     (of params itself when cache_params is NULL: a pure
 *                     index remapping);
 *   row_map[i] <  0 : lookup i reads row i of params (not cached); params may
 *                     be any device-accessible memory, e.g. pinned host memory.
 * cache_params has the dtype and row width of params.  Sum / mean only, fp32
 * accumulation; everything else as cuembed_forward, and the result is
 * bit-identical to cuembed_forward on the table the mapping describes.
 */
int cuembed_forward_mapped(const void* params, int in_dtype, int embed_width,
                           const void* indices, int idx_type,
                           const void* offsets, int off_type,
                           const void* weights, int batch_size, int num_hots,
                           int mode, void* ret, int out_dtype,
                           const void* row_map, const void* cache_params,
                           cuembed_stream_t stream);

/*
 * Multi-table batched lookup (new; the reference is "single table",
 * README.md:110): num_tables pooled lookups in one launch per 32 tables.  All
 * tables share in_dtype, embed_width, idx_type, off_type, out_dtype and
 * weighted-ness (weights == NULL, or one non-NULL pointer per table); batch
 * size, fixed hotness / CSR offsets and the mode (CUEMBED_SUM / CUEMBED_MEAN;
 * modes == NULL means sum) are per table.  params / indices / offsets /
 * weights / rets / batch_sizes / num_hots / modes are HOST arrays of num_tables
 * entries (device pointers inside).  out_row_stride is the pitch of every
 * output in elements of out_dtype (0 = embed_width), so all tables can write
 * their slice of one [batch, num_tables * embed_width] matrix.  Every table's
 * result is bit-identical to its own cuembed_forward call.
 */
int cuembed_forward_multi(int num_tables, const void* const* params,
                          int in_dtype, int embed_width,
                          const void* const* indices, int idx_type,
                          const void* const* offsets, int off_type,
                          const void* const* weights, const int* batch_sizes,
                          const int* num_hots, const int* modes,
                          void* const* rets, int out_dtype,
                          long long out_row_stride, cuembed_stream_t stream);

/*
 * Pooled lookup with a shared-memory cache of hot table rows (new; the
 * power-law policy of BASELINE.json's north_star, measured in
 * profiles/r02_notes.md).  hot_rows [capacity] int32 row ids and *hot_count
 * (device int32, clamped to capacity) name the rows to keep in shared memory;
 * any list is correct (rows not in the batch, duplicates, negative entries are
 * ignored), a good one holds the most frequent rows of the batch --
 * cuembed_hot_rows_from_sorted derives it from the sorted indices of a
 * transposed batch (rows hit >= min_count times).  Results are bit-identical
 * to cuembed_forward.  Scope: row bytes 128 / 256 / 512, int32 indices, sum or
 * mean; cuembed_forward_hot_capacity returns the number of rows that fit (0:
 * shape not supported, use cuembed_forward).
 */
int cuembed_forward_hot_capacity(int in_dtype, int embed_width);
int cuembed_forward_hot(const void* params, int in_dtype, int embed_width,
                        const void* indices, int idx_type, const void* offsets,
                        int off_type, const void* weights, int batch_size,
                        int num_hots, int mode, void* ret, int out_dtype,
                        const int* hot_rows, const int* hot_count, int capacity,
                        cuembed_stream_t stream);
int cuembed_hot_rows_from_sorted(const void* sorted_keys, int idx_type, int nnz,
                                 int min_count, int* hot_rows, int capacity,
                                 int* hot_count, cuembed_stream_t stream);

/* row_ids[i] = i / num_hots, nnz = batch_size * num_hots. */
int cuembed_extract_row_ids_fixed(int batch_size, int num_hots, void* row_ids,
                                  int idx_type, cuembed_stream_t stream);
/* row_ids[offsets[b] .. offsets[b+1]) = b. */
int cuembed_extract_row_ids_csr(const void* offsets, int off_type,
                                int batch_size, void* row_ids, int idx_type,
                                cuembed_stream_t stream);
/* row_ids[i] = i. */
int cuembed_extract_row_ids_concat(int nnz, void* row_ids, int idx_type,
                                   cuembed_stream_t stream);

/*
 * Stable sort of the COO triples by `cols` (table index).  rows = sample ids,
 * weights optional (weight_dtype is the element dtype of weights).
 * transpose_rows receives the sorted table indices, transpose_cols the sample
 * ids, transpose_weights the weights.  Two-call workspace protocol.
 */
int cuembed_transpose(const void* rows, const void* cols, const void* weights,
                      int weight_dtype, int nnz, int idx_type,
                      void* transpose_rows, void* transpose_cols,
                      void* transpose_weights, char* work, size_t* lwork,
                      cuembed_stream_t stream);

/*
 * Fixed-hotness COO transposed in ONE call (new): equivalent to
 * cuembed_extract_row_ids_fixed + cuembed_transpose, but the sample id of
 * position i (= i / num_hots) is synthesised inside the first sort pass, so the
 * row-id array is neither written nor read.  Same outputs, same workspace size
 * as cuembed_transpose with nnz = batch_size * num_hots.
 */
int cuembed_transpose_fixed(const void* cols, int batch_size, int num_hots,
                            const void* weights, int weight_dtype, int idx_type,
                            void* transpose_rows, void* transpose_cols,
                            void* transpose_weights, char* work, size_t* lwork,
                            cuembed_stream_t stream);

/* Dense rank of each element of a grouped index array:
 * [4,4,7,8,8,8,18] -> [0,0,1,2,2,2,3].  Two-call workspace protocol. */
int cuembed_compressed_grad_indices(const void* indices, int idx_type, int nnz,
                                    void* remapped_indices, char* work,
                                    size_t* lwork, cuembed_stream_t stream);

/*
 * Gradient w.r.t. the table: grad_embedding[r] = sum over the run of equal
 * transpose_indices of weight * grad_y[sample].  Deterministic: fp32
 * accumulation in a fixed order, one rounding to `dtype` per output element,
 * no atomics.  transpose_remapped_indices != NULL selects compressed
 * gradients and fills inverse_mapping.  skip_grad_init == 0 zero-fills
 * grad_embedding first; with skip_grad_init != 0 rows that receive a gradient
 * are overwritten and all other rows are left untouched.
 *
 * cuembed_backward takes its scratch (a few MB of partial sums for runs that
 * span thread blocks) from a library-owned stream-ordered memory pool, because
 * the reference signature has no workspace argument.
 * cuembed_backward_ws is the same operation with caller-provided scratch
 * (two-call protocol) for callers that must not allocate.
 */
int cuembed_backward(const void* grad_y, int dtype, int embed_width,
                     int num_grad_embedding_rows, int nnz, int idx_type,
                     const void* transpose_indices,
                     const void* transpose_sample_ids,
                     const void* transpose_remapped_indices,
                     const void* transpose_weights, int skip_grad_init,
                     void* grad_embedding, void* inverse_mapping,
                     cuembed_stream_t stream);

int cuembed_backward_ws(const void* grad_y, int dtype, int embed_width,
                        int num_grad_embedding_rows, int nnz, int idx_type,
                        const void* transpose_indices,
                        const void* transpose_sample_ids,
                        const void* transpose_remapped_indices,
                        const void* transpose_weights, int skip_grad_init,
                        void* grad_embedding, void* inverse_mapping,
                        char* work, size_t* lwork, cuembed_stream_t stream);

/*
 * Fused backward + sparse optimizer step (new; the reference lists "optimizer"
 * as a future kernel type, README.md:119).  Same inputs as cuembed_backward
 * with full-table indexing (transpose_indices are table rows; no compressed
 * indices needed), but instead of writing a gradient the finished sum g of
 * every touched row is applied to `params` [rows, embed_width] (dtype) in
 * place:
 *   CUEMBED_OPT_SGD      p <- p (+) round(-(lr * g)): the update is rounded to
 *                        the table's type and added by the L2 reduction unit
 *                        (one vector red per 16 bytes, nothing is read back
 *                        into the SM); fp32 tables: p - lr * g with subnormal
 *                        results flushed to zero.  One reduction per row, so
 *                        the result is deterministic.
 *   CUEMBED_OPT_ADAGRAD  s <- s + g * g;  p <- p - (lr * g) / (sqrt(s) + eps),
 *                        every operation rounded separately in fp32, with
 *                        `state` = s [rows, embed_width] fp32.
 * Rows that receive no gradient are not touched.  Two-call workspace protocol.
 */
#define CUEMBED_OPT_NONE 0
#define CUEMBED_OPT_SGD 1
#define CUEMBED_OPT_ADAGRAD 2
int cuembed_backward_update(const void* grad_y, int dtype, int embed_width,
                            int nnz, int idx_type,
                            const void* transpose_indices,
                            const void* transpose_sample_ids,
                            const void* transpose_weights, int optimizer,
                            float lr, float eps, void* params, float* state,
                            char* work, size_t* lwork, cuembed_stream_t stream);

/*
 * Row-sharded multi-GPU mode (new: the reference is single-GPU, README.md:110).
 * A GPU that owns table rows [row_lo, row_hi) selects, from the replicated
 * lookup indices of the global batch, the lookups that fall into its range:
 * local_offsets [batch_size + 1] (int32 CSR offsets), local_indices (rebased
 * to the shard: index - row_lo, same integer type as `indices`, capacity nnz)
 * and local_weights keep bag order, so cuembed_forward / cuembed_transpose /
 * cuembed_backward run unchanged on the shard.  Two-call workspace protocol.
 */
int cuembed_shard_select(const void* indices, int idx_type, const void* offsets,
                         int off_type, const void* weights, int weight_dtype,
                         int batch_size, int num_hots, long long row_lo,
                         long long row_hi, int* local_offsets,
                         void* local_indices, void* local_weights, char* work,
                         size_t* lwork, cuembed_stream_t stream);

/*
 * The same selection producing the COO form the transpose consumes directly:
 * local_sample_ids (optional, same integer type as indices) receives the bag of
 * every selected lookup, so no row-id extraction pass is needed; counts
 * (optional, [batch_size] int32, e.g. from cuembed_shard_pool_push) are the
 * per-bag numbers of selected lookups if the caller already has them, which
 * saves the counting pass over the indices.
 */
int cuembed_shard_select_coo(const void* indices, int idx_type,
                             const void* offsets, int off_type,
                             const void* weights, int weight_dtype,
                             int batch_size, int num_hots, long long row_lo,
                             long long row_hi, const int* counts,
                             int* local_offsets, void* local_indices,
                             void* local_sample_ids, void* local_weights,
                             char* work, size_t* lwork,
                             cuembed_stream_t stream);

/*
 * Epilogue after the reduce-scatter of the fp32 partial sums: for the samples
 * [sample0, sample0 + n_samples) of the global batch, out = partial (sum) or
 * partial / global bag length (mean; / sum of weights if weights != NULL, zero
 * vector if that sum is 0), cast to out_dtype.  offsets / num_hots / weights
 * describe the GLOBAL batch.
 */
int cuembed_shard_finalize(const void* partial_f32, int n_samples,
                           int embed_width, int mode, const void* offsets,
                           int off_type, int num_hots, int sample0,
                           const void* weights, int weight_dtype, void* out,
                           int out_dtype, cuembed_stream_t stream);

/*
 * Row-sharded mode, exchange fused with the kernels over NVLink / NVSwitch peer
 * memory (one process per GPU; no collective library call on the data path).
 *
 * Peer memory: cuembed_peer_alloc returns zero-filled device memory that can
 * be exported (cuembed_peer_export -> an opaque CUEMBED_PEER_HANDLE_BYTES
 * handle to ship to the other processes by any means) and mapped by them
 * (cuembed_peer_open; cuembed_peer_close unmaps).  The functions below take
 * HOST arrays of `world` device pointers, entry o being rank o's buffer as
 * mapped in the calling process (the caller's own buffer at entry `rank`).
 *
 * Flag buffer: every rank owns CUEMBED_PEER_FLAG_BYTES of peer memory:
 * uint32 flags[CUEMBED_PEER_CHANNELS][CUEMBED_MAX_WORLD] followed by one
 * uint32 status word (non-zero after a wait timed out).  cuembed_shard_signal
 * release-stores `epoch` into flags[channel][rank] of every rank;
 * waits succeed once all `world` flags of the channel reached `epoch`.
 */
#define CUEMBED_MAX_WORLD 16
#define CUEMBED_PEER_HANDLE_BYTES 64
#define CUEMBED_PEER_CHANNELS 4
#define CUEMBED_PEER_FLAG_BYTES 512

int cuembed_peer_alloc(size_t bytes, void** ptr);
int cuembed_peer_free(void* ptr);
int cuembed_peer_export(void* ptr, unsigned char* handle);
int cuembed_peer_open(const unsigned char* handle, void** ptr);
int cuembed_peer_close(void* ptr);

/*
 * Forward, step 1: pool the lookups of the GLOBAL batch (replicated indices /
 * offsets / weights, batch_size % world == 0) that fall into this rank's rows
 * [row_lo, row_hi) -- local_params holds exactly those rows -- in bag order
 * with fp32 accumulation, and store each partial row directly into the
 * exchange buffer of the rank that owns the bag:
 *   slot_ptrs[o] + (rank * per + bag - o * per) * embed_width   (partial_dtype
 * elements, per = batch_size / world, o = bag / per).
 * counts (optional, [batch_size] int32) receives the number of lookups of each
 * bag that this rank owns.
 */
int cuembed_shard_pool_push(const void* local_params, int in_dtype,
                            int embed_width, const void* indices, int idx_type,
                            const void* offsets, int off_type,
                            const void* weights, int batch_size, int num_hots,
                            long long row_lo, long long row_hi,
                            void* const* slot_ptrs, int world, int rank,
                            int partial_dtype, int* counts,
                            cuembed_stream_t stream);

/* Sharded concat (fixed hotness): the owner of each looked-up row copies it to
 * out_ptrs[o] + ((b - o * per) * num_hots + j) * embed_width, o = b / per. */
int cuembed_shard_concat_push(const void* local_params, int dtype,
                              int embed_width, const void* indices,
                              int idx_type, int batch_size, int num_hots,
                              long long row_lo, long long row_hi,
                              void* const* out_ptrs, int world, int rank,
                              cuembed_stream_t stream);

/* Tell every rank that this rank's pushes of `epoch` are complete (enqueue
 * after the pushing kernel / copies on the same stream). */
int cuembed_shard_signal(void* const* flag_ptrs, int world, int rank,
                         int channel, unsigned epoch, cuembed_stream_t stream);
/* Hold the stream until all ranks signalled `epoch` on `channel`.  A wait
 * gives up after the peer timeout (default 60 s, CUEMBED_PEER_TIMEOUT_MS or
 * cuembed_shard_set_timeout_ms): the status word (4 bytes at
 * flags + CUEMBED_PEER_CHANNELS * CUEMBED_MAX_WORLD words) then holds 1 + the
 * missing rank, and -- so that stale data never passes for a result -- slice
 * [rank] of `poison` ([world] slices of poison_bytes; may be NULL) is filled
 * with 0xff bytes (NaN in every element type).  cuembed_shard_reduce_finalize
 * poisons its whole output the same way. */
int cuembed_shard_wait(const void* flags, int world, int channel,
                       unsigned epoch, void* poison, size_t poison_bytes,
                       cuembed_stream_t stream);
int cuembed_shard_set_timeout_ms(long long timeout_ms);


/*
 * Forward, step 2 (on the bag owner): wait for all ranks, then
 * out[s, :] = cast(scale * (slot 0 + slot 1 + ... in rank order)) for the
 * samples [sample0, sample0 + n_samples) of the global batch; slots is this
 * rank's exchange buffer [world][n_samples][embed_width] of partial_dtype;
 * scale as in cuembed_shard_finalize.
 */
int cuembed_shard_reduce_finalize(const void* slots, int partial_dtype,
                                  int world, const void* flags, int channel,
                                  unsigned epoch, int n_samples,
                                  int embed_width, int mode,
                                  const void* offsets, int off_type,
                                  int num_hots, int sample0,
                                  const void* weights, int weight_dtype,
                                  void* out, int out_dtype,
                                  cuembed_stream_t stream);

/* Backward: copy this rank's `bytes` of grad_y into every rank's gather buffer
 * at offset rank * bytes (copy engines; follow with cuembed_shard_signal). */
int cuembed_shard_allgather_push(const void* src, size_t bytes,
                                 void* const* gather_ptrs, int world, int rank,
                                 cuembed_stream_t stream);

/*
 * Diagnostics: measured gather ceiling for the roofline (csrc/microbench.cu).
 * Gathers n rows of row_bytes (128, 256 or 512) at rows[i] * row_bytes from
 * buf with the access shape of the forward / backward kernels (lane group per
 * row, 16-byte loads, 8 rows in flight, persistent grid) and no arithmetic
 * beyond one XOR per word.  On an L2-resident buffer this is the L2 -> SM
 * gather ceiling, on a multi-GB buffer the DRAM gather ceiling.
 * no_l1_allocate: 1 = ld.global.nc.L1::no_allocate, 2 = 16 rows in flight per
 * warp, 3 = L2 cache hints (evict_last on rows, evict_first on the index list);
 * 2 and 3 for 512-byte rows only.  sink: 4 bytes.
 */
int cuembed_microbench_gather(const void* buf, int row_bytes, const int* rows,
                              long long n, int no_l1_allocate, unsigned* sink,
                              cuembed_stream_t stream);
/* The same gather with the rows landing in shared memory through per-lane bulk
 * copies (cp.async.bulk + mbarrier) and read back with LDS.128; row_bytes must
 * be 512.  variant selects rows per stage x stages x warps per SM:
 * 0 = 32x2x6, 1 = 16x3x8, 2 = 32x1x12, 3 = 16x2x12, 4 = 8x4x12. */
int cuembed_microbench_gather_bulk(const void* buf, int row_bytes,
                                   const int* rows, long long n, int variant,
                                   unsigned* sink, cuembed_stream_t stream);
/* The same gather with per-lane 16-byte asynchronous copies (cp.async.cg, SASS
 * LDGSTS) into the lane's own shared-memory slot, waited for with
 * cp.async.wait_group (no barrier); row_bytes must be 512.  variant selects
 * batches of 8 rows in flight per warp x warps per CTA x CTAs per SM:
 * 0 = 2x4x6, 1 = 3x4x4, 2 = 2x4x4, 3 = 4x4x3, 4 = 2x8x3. */
int cuembed_microbench_gather_async(const void* buf, int row_bytes,
                                    const int* rows, long long n, int variant,
                                    unsigned* sink, cuembed_stream_t stream);

/*
 * Debug aid (new).  The reference removes every bounds check from its kernels
 * (cuembed/include/embedding_lookup_ops.cuh:59) and so does this library; this
 * call validates a lookup's inputs on the device instead: every index must lie
 * in [0, num_rows), and (offsets != NULL) the offsets must be non-negative,
 * non-decreasing and end at or below nnz.  It SYNCHRONISES the stream.  Returns
 * CUEMBED_OK, CUEMBED_ERR_OFFSETS or CUEMBED_ERR_INDEX_RANGE, with the first
 * offending position (bag for offsets, lookup for indices) in
 * *first_bad_position (-1 if none; may be NULL).
 * With CUEMBED_NVTX=1 in the environment every entry point of this library
 * also opens an NVTX range named after itself around its launches.
 */
int cuembed_debug_check_lookup(const void* indices, int idx_type, long long nnz,
                               long long num_rows, const void* offsets,
                               int off_type, int batch_size,
                               long long* first_bad_position,
                               cuembed_stream_t stream);

/* Number of kernels this library has launched in this process (all threads);
 * used by bench.py to report `gpu_launches`. */
unsigned long long cuembed_launch_count(void);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* CUEMBED_B200_H_ */
