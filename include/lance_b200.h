/*
 * lance_b200.h -- C ABI of the B200-native IVF-PQ / IVF-FLAT hot path.
 *
 * This is the drop-in boundary for lancedb/lance: every entry point replaces one function (or one
 * trait method) of the reference's `lance-index::vector::{kmeans,ivf,pq,flat}` / `lance-linalg`
 * crates; the reference file:line it replaces is cited above each declaration (paths relative to
 * /root/reference/rust).  INTEGRATION.md shows the Rust `extern "C"` block + shim a maintainer adds.
 *
 * Conventions
 *   - Plain pointers and sizes only.  Every data pointer may be a HOST pointer (pageable or pinned)
 *     or a DEVICE pointer of the current device; the library detects which (cudaPointerGetAttributes)
 *     and stages host buffers through the device itself.  Outputs are written where they point.
 *   - Caller allocates inputs AND outputs; the library never frees or keeps a caller pointer after
 *     the call returns, except inside an `lb2_index` handle, which owns private device copies.
 *   - Vectors are Arrow FixedSizeList values buffers: contiguous row-major n x d.
 *   - Every function returns lb2_status and never unwinds/aborts across the boundary; the message
 *     of the last failure on the calling thread is available from lb2_last_error().
 *   - There is NO CPU fallback: without a usable CUDA device every compute entry point returns
 *     LB2_NO_DEVICE.
 *   - Calls are blocking (results are complete on return).  All entry points are thread-safe;
 *     each calling thread uses the CUDA device selected by lb2_set_device() on that thread and its
 *     own CUDA stream, so concurrent callers (the reference searches up to ncpu-2 partitions at a
 *     time, rust/lance/src/io/exec/knn.rs:881) overlap on the device.
 *   - Stream variants: lb2_set_stream() orders every later call of the thread on a caller-owned
 *     cudaStream_t; lb2_index_search_async() enqueues a whole search and returns without waiting.
 */
#ifndef LANCE_B200_H_
#define LANCE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  LB2_OK = 0,
  LB2_INVALID_ARG = 1,
  LB2_UNSUPPORTED = 2, /* valid request that this build does not implement (never silently emulated) */
  LB2_CUDA_ERROR = 3,
  LB2_NCCL_ERROR = 4,
  LB2_OOM = 5,
  LB2_NO_DEVICE = 6
} lb2_status;

/* element type of a vector buffer (arrow DataType of the FixedSizeList child) */
typedef enum { LB2_F32 = 0, LB2_F16 = 1, LB2_BF16 = 2, LB2_U8 = 3 } lb2_dtype;

/* lance_linalg::distance::DistanceType (lance-linalg/src/distance.rs) */
typedef enum { LB2_L2 = 0, LB2_COSINE = 1, LB2_DOT = 2 } lb2_metric;

typedef struct lb2_index lb2_index; /* device-resident IVF_PQ / IVF_FLAT index */

/* ---- runtime -------------------------------------------------------------------------------- */
const char* lb2_version(void);
size_t lb2_last_error(char* buf, size_t len); /* copies the calling thread's last message */
int lb2_device_count(void);                   /* 0 when no CUDA device is usable */
lb2_status lb2_set_device(int device);
lb2_status lb2_synchronize(void);
/* Give cached device memory back: the calling thread's bulk-copy staging buffer (host-sourced builds keep one
 * of up to LB2_STAGING_CACHE_MB, default 1024 MB, between calls so that the copy of the next build starts at
 * once) and every free block of the stream-ordered pool.  The reference has no counterpart (its buffers are
 * Arrow arrays dropped with the batch); call it when a burst of index builds is over. */
lb2_status lb2_trim_memory(void);
/* Bind the calling thread's library context to a caller-owned CUDA stream (cudaStream_t passed as
 * void*): all work of later calls from this thread -- kernels, copies, stream-ordered allocations --
 * is enqueued on it, after whatever the caller enqueued before.  Blocking entry points still wait for
 * their own results (that is a cudaStreamSynchronize of this stream).  NULL returns to the thread's
 * private stream.  The stream must belong to the thread's device and outlive the binding. */
lb2_status lb2_set_stream(void* cuda_stream);
/* device / pinned-host buffers for callers that keep data resident (bench, the Rust shim's ring) */
lb2_status lb2_malloc(void** ptr, size_t bytes);
lb2_status lb2_free(void* ptr);
lb2_status lb2_malloc_host(void** ptr, size_t bytes); /* pinned */
lb2_status lb2_free_host(void* ptr);
lb2_status lb2_memcpy(void* dst, const void* src, size_t bytes); /* direction auto-detected */
/* instrumentation: number of kernels this library launched on the calling thread's device since
 * the last reset, and per-kernel CUDA-event timing (name = kernel family, e.g. "pq_scan"). */
lb2_status lb2_launch_count(uint64_t* count, int reset);
lb2_status lb2_profile_enable(int on);
lb2_status lb2_profile_get(const char* name, uint64_t* launches, double* total_ms);
lb2_status lb2_profile_reset(void);
/* all entries as "name\tlaunches\ttotal_ms\n" lines; returns the full length (like snprintf) */
size_t lb2_profile_dump(char* buf, size_t len);
lb2_status lb2_timer_start(void);          /* CUDA event on the library's stream */
lb2_status lb2_timer_stop(float* ms_out);  /* records, synchronises, returns elapsed ms */

/* ---- lance-linalg distance API -------------------------------------------------------------- */
/* l2_distance_batch / dot_distance_batch / cosine_distance_batch
 * (lance-linalg/src/distance/l2.rs:194-203, dot.rs:164-172, cosine.rs:266-290):
 * out[i] = dist(from, to[i]) for i < n.  f32 L2/Dot are bit-exact to the reference's 16-lane order. */
lb2_status lb2_distance_batch(const void* from, const void* to, uint64_t n, uint32_t d,
                              lb2_dtype dtype, lb2_metric metric, float* out);
/* normalize_fsl (lance-linalg/src/kernels.rs:141-146,201-211): out[i] = x[i] / ||x[i]|| */
lb2_status lb2_normalize(const void* vectors, uint64_t n, uint32_t d, lb2_dtype dtype, void* out);

/* ---- k-means (lance-index/src/vector/kmeans.rs) --------------------------------------------- */
typedef struct {
  uint32_t max_iters;      /* KMeansParams::max_iters, default 50 (kmeans.rs:92-103) */
  double tolerance;        /* 1e-4 */
  uint32_t redos;          /* 1.  Every redo restarts from the same rng state (kmeans.rs:645-653), so without a
                              balance bias redos > 1 equals one run; with a bias redos > 1 -> LB2_UNSUPPORTED */
  float balance_factor;    /* BEFORE the division by n that train_kmeans applies (kmeans.rs:1344);
                              IVF training passes 1.0 (rust/lance/src/index/vector/ivf.rs:1858) */
  uint32_t hierarchical_k; /* 16: for k > 256 (and no init_centroids) the reference's hierarchical
                              scheme is used (kmeans.rs:746-1003, 1027); 0/1 = flat Lloyd for every k */
  uint64_t sample_rate;    /* 256: only the first sample_rate*k rows are used (kmeans.rs:1328-1340) */
  uint64_t seed;           /* the reference is unseeded (kmeans.rs:645); we are reproducible */
  const void* init_centroids; /* KMeanInit::Incremental (k x d, same dtype) or NULL = random rows */
  lb2_metric metric;       /* L2 or DOT (cosine callers normalise first, as the reference does) */
} lb2_kmeans_params;
void lb2_kmeans_params_default(lb2_kmeans_params* p);

/* train_kmeans<T>(array, params, dimension, k, sample_rate) -> KMeans  (kmeans.rs:1309-1347) */
lb2_status lb2_kmeans_train(const void* data, uint64_t n, uint32_t d, lb2_dtype dtype, uint32_t k,
                            const lb2_kmeans_params* params, void* centroids_out, double* loss_out,
                            uint32_t* iters_out);

/* compute_partitions_arrow_array / compute_partitions_with_dists (kmeans.rs:1187-1294):
 * part_out[i] = argmin_k dist(vectors[i], centroids[k]) (first minimum), dist_out[i] that distance,
 * valid_out[i] = 0 where the reference returns None (all NaN/Inf).  dist_out/valid_out nullable. */
lb2_status lb2_compute_partitions(const void* centroids, uint32_t k, uint32_t d, lb2_dtype dtype,
                                  lb2_metric metric, const void* vectors, uint64_t n,
                                  uint32_t* part_out, float* dist_out, uint8_t* valid_out);

/* kmeans_find_partitions_arrow_array (kmeans.rs:1076-1158), batched over nq queries:
 * ids/dists are [nq][nprobes], ascending by (distance, id). */
lb2_status lb2_find_partitions(const void* centroids, uint32_t k, uint32_t d, lb2_dtype dtype,
                               lb2_metric metric, const void* queries, uint64_t nq,
                               uint32_t nprobes, uint32_t* ids_out, float* dists_out);

/* compute_residual (lance-index/src/vector/residual.rs:111-154): out[i] = v[i] - centroids[part[i]] */
lb2_status lb2_compute_residual(const void* centroids, uint32_t k, uint32_t d, lb2_dtype dtype,
                                const void* vectors, uint64_t n, const uint32_t* part_ids,
                                void* out);

/* ---- product quantisation (lance-index/src/vector/pq*.rs) ------------------------------------ */
typedef struct {
  uint32_t num_sub_vectors; /* PQBuildParams (pq/builder.rs:27-59): 16 */
  uint32_t num_bits;        /* 8 or 4 */
  uint32_t max_iters;       /* 50 */
  uint32_t kmeans_redos;    /* 1 (PQ k-means has no balance bias: any value equals one run, see redos above) */
  uint64_t sample_rate;     /* 256 */
  const void* codebook;     /* user codebook to continue from, or NULL */
  uint64_t seed;
} lb2_pq_params;
void lb2_pq_params_default(lb2_pq_params* p);

/* PQBuildParams::build(data, distance_type) -> ProductQuantizer (pq/builder.rs:162-194):
 * codebook_out is the flat [M][2^nbits][d/M] layout of pq/utils.rs:59-76; iters_out[M] nullable. */
lb2_status lb2_pq_train(const void* data, uint64_t n, uint32_t d, lb2_dtype dtype,
                        lb2_metric metric, const lb2_pq_params* params, void* codebook_out,
                        uint32_t* iters_out);

/* ProductQuantizer::quantize / transform_impl (pq.rs:116-191,430).  When `centroids`
 * ([num_centroids][d]) and `part_ids` are given the residual (residual.rs:161-205) is fused: codes of
 * v - centroids[part]; a part id >= num_centroids is LB2_INVALID_ARG.
 * codes_out is row-major [n][M] (8-bit) or [n][M/2] (4-bit: byte i = code[2i+1] << 4 | code[2i],
 * pq.rs:168-173; 16 codewords per sub-space, M even). */
lb2_status lb2_pq_encode(const void* codebook, uint32_t num_sub_vectors, uint32_t num_bits,
                         uint32_t d, lb2_dtype dtype, lb2_metric metric, const void* centroids,
                         uint32_t num_centroids, const uint32_t* part_ids, const void* vectors, uint64_t n,
                         uint8_t* codes_out);

/* build_distance_table_l2 / _dot (pq/distance.rs:24-92): lut_out[M * 2^nbits] f32 */
lb2_status lb2_pq_build_lut(const void* codebook, uint32_t num_sub_vectors, uint32_t num_bits,
                            uint32_t d, lb2_metric metric, const float* query, float* lut_out);

/* compute_pq_distance (pq/distance.rs:109-144) on TRANSPOSED codes [M][n], as the reference's
 * storage holds them; PQDistCalculator::distance_all's Dot correction (pq/storage.rs:957-958)
 * is applied when metric == LB2_DOT. */
lb2_status lb2_pq_scan(const float* lut, uint32_t num_sub_vectors, uint32_t num_bits,
                       lb2_metric metric, const uint8_t* codes_transposed, uint64_t n,
                       float* dists_out);

/* 4-bit PQ: compute_pq_distance_4bit (pq/distance.rs:147-242) with PQDistCalculator::distance_all's
 * Dot correction.  lut = M x 16 f32 (lb2_pq_build_lut with num_bits = 4), codes_transposed = packed
 * [M/2][n].  The first min(max(200, k_hint), n) rows and the last n % 16 rows are exact f32 sums; the
 * rest is the reference's u8-quantised table with saturating u8 accumulation, dequantised.  k_hint =
 * the k of the search (DistCalculator::distance_all(k_hint), flat/index.rs:99). */
lb2_status lb2_pq_scan_4bit(const float* lut, uint32_t num_sub_vectors, lb2_metric metric,
                            const uint8_t* codes_transposed, uint64_t n, uint64_t k_hint,
                            float* dists_out);

/* FlatIndex::search fast path over a distance array (flat/index.rs:97-127): the k smallest
 * (distance, position) pairs; out sorted ascending by (distance, row id).  *count_out <= k. */
lb2_status lb2_flat_topk(const float* dists, const uint64_t* row_ids, uint64_t n, uint32_t k,
                         uint64_t* ids_out, float* dists_out, uint32_t* count_out);
/* The same with FlatIndex::search's range branch (flat/index.rs:100-115): only rows with
 * lower <= dist < upper (f32::total_cmp order; an absent bound is f32::MIN / f32::MAX, as the reference
 * unwraps it) are offered to the heap.  Both calls return exactly the SET the reference's BinaryHeap ends
 * with -- also when more rows tie at the k-th distance than fit (the heap's sift order is restated). */
lb2_status lb2_flat_topk_range(const float* dists, const uint64_t* row_ids, uint64_t n, uint32_t k,
                               int has_lower, float lower, int has_upper, float upper,
                               uint64_t* ids_out, float* dists_out, uint32_t* count_out);

/* IvfTransformer::transform for IVF_PQ (lance-index/src/vector/ivf.rs:188-236,357): for a batch,
 * [normalise if cosine] -> partition id -> residual -> PQ code, in one pass over the vectors.
 * `metric` is the index metric: it selects the partition assignment and whether residuals are taken
 * (not for dot, PQBuildParams::use_residual); the PQ codes are L2 codes in every case, because the
 * index builder trains its quantizer with DistanceType::L2 (rust/lance/src/index/vector/builder.rs:460).
 * valid_out[i] = 0 marks rows KeepFiniteVectors would drop (transform.rs:112-159). */
lb2_status lb2_ivfpq_transform(const void* centroids, uint32_t k, const void* codebook,
                               uint32_t num_sub_vectors, uint32_t num_bits, uint32_t d,
                               lb2_dtype dtype, lb2_metric metric, const void* vectors, uint64_t n,
                               uint32_t* part_out, uint8_t* codes_out, uint8_t* valid_out);

/* ---- device-resident index: IVFIndex<FlatIndex, ProductQuantizer> ----------------------------
 * (rust/lance/src/index/vector/ivf/v2.rs:104; storage lance-index/src/vector/pq/storage.rs:151) */
lb2_status lb2_index_create(const void* centroids, uint32_t k, uint32_t d, lb2_dtype dtype,
                            lb2_metric metric, const void* codebook, uint32_t num_sub_vectors,
                            uint32_t num_bits, lb2_index** out);
/* num_bits = 8 or 4 (4: 16 codewords per sub-space, codes are [n][M/2] packed bytes, M even; searches
 * use compute_pq_distance_4bit's flat-rows + u8-quantised-table rule with k_hint = k, and exact
 * row-by-row distances when a prefilter is given, as the reference does).
 * load the (row_id, __ivf_part_id, __pq_code) shuffle output (builder.rs:685-937): rows are grouped
 * by partition on the device (stable, i.e. input order inside a partition). Replaces prior content. */
lb2_status lb2_index_load(lb2_index* index, const uint32_t* part_ids, const uint8_t* codes,
                          const uint64_t* row_ids /* NULL = 0..n */, uint64_t n);
/* IVFIndex::find_partitions + search_in_partition for a BATCH of queries, then the global merge
 * SortExec(_distance, _rowid).fetch(k) (v2.rs:455-500, rust/lance/src/dataset/scanner.rs:3450-3466).
 * Outputs [nq][k]; unused slots: row id = UINT64_MAX, distance = +inf; counts_out[nq] nullable. */
lb2_status lb2_index_search(lb2_index* index, const void* queries, uint64_t nq, uint32_t k,
                            uint32_t nprobes, uint64_t* row_ids_out, float* dists_out,
                            uint32_t* counts_out);
/* The same, followed by the refine step of the reference's plan (Take vectors + flat KNN re-rank,
 * rust/lance/src/dataset/scanner.rs:2884-2905; lance-index/src/vector/flat.rs:95-148): the index
 * returns k * refine_factor candidates, their EXACT distances to the query are recomputed from the
 * raw column `vectors` ([>= max row id + 1][d], element type = the index's dtype, row id = row
 * number; device-resident for speed) with the index's true metric, and the k best by
 * (distance, row id) are returned.  k * refine_factor <= 1024. */
lb2_status lb2_index_search_refine(lb2_index* index, const void* vectors, uint64_t num_vectors,
                                   const void* queries, uint64_t nq, uint32_t k, uint32_t nprobes,
                                   uint32_t refine_factor, uint64_t* row_ids_out, float* dists_out,
                                   uint32_t* counts_out);
/* Extended search: prefilter and refine in one call.
 * Prefilter = the reference's PreFilter / RowIdMask (lance-index/src/prefilter.rs:27-51): when the
 * mask is not empty FlatIndex::search visits the partition row by row and skips rows for which
 * RowIdMask::selected(row_id) is false BEFORE they can enter the heap (flat/index.rs:129-165).
 * Here the mask is a bitmap over STORAGE positions (the order of lb2_index_export's row_ids, i.e.
 * partition-local offset + part_offsets[p]); bit i (word i/64, bit i%64) set = row may be returned.
 * lb2_index_row_mask builds it on the device from the RowIdMask's allow / block lists.
 * allow_bitmap NULL = the fast unfiltered path. Host or device pointer, (num_rows+63)/64 words. */
typedef struct {
  uint32_t k;
  uint32_t nprobes;
  uint32_t refine_factor;       /* 0 = no refine step */
  const void* refine_vectors;   /* raw column, see lb2_index_search_refine; required when refine_factor > 0 */
  uint64_t num_vectors;
  const uint64_t* allow_bitmap; /* nullable */
  /* range query (Query::lower_bound / upper_bound, lance-index/src/vector.rs:83-86): inside a partition
   * only rows with lower <= _distance < upper enter the top-k (flat/index.rs:100-115, index distances);
   * with refine the plan filters the exact distances the same way afterwards (scanner.rs:3342-3377) */
  uint32_t has_lower_bound, has_upper_bound;
  float lower_bound, upper_bound;
} lb2_search_params;
lb2_status lb2_index_search_ex(lb2_index* index, const void* queries, uint64_t nq,
                               const lb2_search_params* params, uint64_t* row_ids_out, float* dists_out,
                               uint32_t* counts_out);
/* Incremental update: the device half of optimize_indices / split / join.
 * The reference expresses an optimize step as per-partition AssignOp::Add / AssignOp::Remove lists against a new
 * centroid set (rust/lance/src/index/vector/builder.rs:1219-1333 split_partition_impl, :1476-1530
 * join_partition_impl, :1534-1650 build_assign_batch) and merges them with the stored partitions.  The decisions
 * (should_split :1152, should_join :1343, assign_vectors :1690 -- built from lb2_kmeans_train(k = 2),
 * lb2_distance_batch and lb2_ivfpq_transform on the moved rows) stay with the host, which owns the dataset;
 * this call is the merge on the device and returns a NEW index:
 *   - new_centroids [new_k][d] in the model's element type (NULL = unchanged, then new_k must equal the old k);
 *   - part_map[old_k] (nullable = identity): new partition id of every old partition, UINT32_MAX = the partition's
 *     rows are dropped (split: the split partition's rows come back through the add list; join: ids after the
 *     deleted partition shift down by one);
 *   - remove_row_ids (sorted ascending): old rows to drop (AssignOp::Remove; also deletions);
 *   - add_*: n_add rows already transformed (partition id, PQ code, row id); inside a partition the surviving old
 *     rows keep their order and the added rows follow in list order.
 * Appending new data to an index (optimize without retraining) is the add list alone. */
lb2_status lb2_index_update(const lb2_index* old_index, const void* new_centroids, uint32_t new_k,
                            const uint32_t* part_map, const uint32_t* add_part_ids, const uint8_t* add_codes,
                            const uint64_t* add_row_ids, uint64_t n_add, const uint64_t* remove_row_ids,
                            uint64_t n_remove, lb2_index** out);
/* Asynchronous search (SURVEY 8b "Threading": `_async` variants taking a stream/event).  Same
 * arguments and results as lb2_index_search_ex, but the call only ENQUEUES the work on `cuda_stream`
 * (cudaStream_t; NULL = the calling thread's current library stream) and returns: probe selection, LUT
 * build, scan, tie replay, merge and the optional refine run in stream order, with no host round trip.
 * Buffers must be device memory or pinned host memory and stay valid until the stream reaches the end of
 * the search; if `done_event` (cudaEvent_t) is not NULL it is recorded there.  Errors detected while
 * enqueueing are returned; the results are defined once the stream (or the event) has completed. */
lb2_status lb2_index_search_async(lb2_index* index, const void* queries, uint64_t nq,
                                  const lb2_search_params* params, uint64_t* row_ids_out, float* dists_out,
                                  uint32_t* counts_out, void* cuda_stream, void* done_event);
/* bitmap_out[(num_rows+63)/64]: bit i = RowIdMask::selected(row id stored at position i)
 * (lance-core/src/utils/mask.rs:84-93: in the allow list if there is one, and not in the block
 * list if there is one). Lists are sorted ascending (RoaringTreemap order); has_* = list present. */
lb2_status lb2_index_row_mask(const lb2_index* index, const uint64_t* allow_ids, uint64_t n_allow,
                              int has_allow, const uint64_t* block_ids, uint64_t n_block, int has_block,
                              uint64_t* bitmap_out);
lb2_status lb2_index_info(const lb2_index* index, uint32_t* k, uint32_t* d, uint32_t* num_sub_vectors,
                          uint32_t* num_bits, uint64_t* num_rows);
/* export for the host to write index files: any pointer may be NULL.
 * part_offsets[k+1]; codes [num_rows][M] and row_ids [num_rows] in partition order. */
lb2_status lb2_index_export(const lb2_index* index, void* centroids_out, void* codebook_out,
                            uint64_t* part_offsets_out, uint8_t* codes_out, uint64_t* row_ids_out);
/* One partition in the layout the reference's storage holds and merge_partitions writes (`__pq_code` with
 * "transposed": true -- lance-index/src/vector/pq/storage.rs:52-67,430-450; rust/lance/src/index/vector/builder.rs:
 * 938-1079): codes column-major [code bytes per row][n_p], row ids [n_p].  Call with NULL outputs first to get
 * *num_rows_out.  tests/: the bytes equal the reference's own fixture test_data/v0.27.1/pq_in_schema. */
lb2_status lb2_index_export_partition(const lb2_index* index, uint32_t partition, uint8_t* codes_transposed_out,
                                      uint64_t* row_ids_out, uint64_t* num_rows_out);
lb2_status lb2_index_destroy(lb2_index* index);

/* IvfIndexBuilder::build (rust/lance/src/index/vector/builder.rs:236): sample -> train IVF ->
 * residuals -> train PQ -> assign + encode every row -> group by partition, all on the device. */
typedef struct {
  uint32_t num_partitions;
  lb2_kmeans_params ivf;  /* balance_factor 1.0, sample_rate 256 (ivf/builder.rs:62-78) */
  lb2_pq_params pq;
  uint64_t seed;          /* training-sample selection */
} lb2_ivfpq_build_params;
void lb2_ivfpq_build_params_default(lb2_ivfpq_build_params* p);
typedef struct {
  float ms_ivf_train, ms_pq_train, ms_transform, ms_group, ms_total; /* CUDA-event times */
  uint32_t ivf_iters, pq_iters_max;
  double ivf_loss;
} lb2_build_stats;
lb2_status lb2_ivfpq_build(const void* data, uint64_t n, uint32_t d, lb2_dtype dtype,
                           lb2_metric metric, const lb2_ivfpq_build_params* params,
                           const uint64_t* row_ids /* NULL = 0..n */, lb2_index** out,
                           lb2_build_stats* stats /* nullable */);

/* ---- IVF_FLAT: IVFIndex<FlatIndex, FlatQuantizer> (lance-index/src/vector/flat/{index,storage}.rs) --
 * The partitions hold the raw f32 vectors (normalised first when the metric is cosine, as
 * IvfTransformer::new_flat does, lance-index/src/vector/ivf.rs:149-185); search scores every row of
 * the probed partitions exactly (FlatDistanceCal::distance_all, flat/storage.rs:397-403) and keeps
 * the k smallest (FlatIndex::search, flat/index.rs:82-177).  lb2_index_search / _info / _destroy
 * work on both index kinds. */
typedef struct {
  uint32_t num_partitions;
  lb2_kmeans_params ivf;
  uint64_t seed;
} lb2_ivfflat_build_params;
void lb2_ivfflat_build_params_default(lb2_ivfflat_build_params* p);
lb2_status lb2_ivfflat_build(const void* data, uint64_t n, uint32_t d, lb2_dtype dtype,
                             lb2_metric metric, const lb2_ivfflat_build_params* params,
                             const uint64_t* row_ids, lb2_index** out, lb2_build_stats* stats);
lb2_status lb2_index_create_flat(const void* centroids, uint32_t k, uint32_t d, lb2_dtype dtype,
                                 lb2_metric metric, lb2_index** out);
/* vectors [n][d] (already normalised for cosine), grouped by partition on the device (stable) */
lb2_status lb2_index_load_flat(lb2_index* index, const uint32_t* part_ids, const void* vectors,
                               const uint64_t* row_ids, uint64_t n);
lb2_status lb2_index_export_flat(const lb2_index* index, void* centroids_out,
                                 uint64_t* part_offsets_out, void* vectors_out,
                                 uint64_t* row_ids_out);

/* ---- multi-GPU (one process per GPU) -----------------------------------------------------------
 * With a communicator every training / build call takes THIS RANK'S ROW SHARD: the k-means loops
 * (flat and hierarchical) exchange their packed per-cluster partial results once per Lloyd iteration
 * (one collective, reduced in rank order on every rank -> bit-identical models on all ranks), the
 * transform and the index are local to the shard. */
/* unique_id is the 128-byte ncclUniqueId produced by rank 0 (lb2_comm_unique_id) and broadcast by
 * the host runtime (torch.distributed / MPI / the Rust side). */
lb2_status lb2_comm_unique_id(void* unique_id_128);
lb2_status lb2_comm_init(const void* unique_id_128, int rank, int nranks);
lb2_status lb2_comm_destroy(void);
lb2_status lb2_comm_info(int* rank, int* nranks); /* (0, 1) without a communicator */
/* Search of a ROW-SHARDED index (every rank built / loaded its own rows, row ids global): the local
 * lb2_index_search_ex result of every rank is exchanged in one collective and merged by (_distance,
 * _rowid) -- the reference's final SortExec.fetch(k) (rust/lance/src/dataset/scanner.rs:3450-3466) --
 * so every rank returns the global top-k.  All ranks must call it with the same queries and params. */
lb2_status lb2_index_search_sharded(lb2_index* index, const void* queries, uint64_t nq,
                                    const lb2_search_params* params, uint64_t* row_ids_out, float* dists_out,
                                    uint32_t* counts_out);
/* Partition ownership by device all-to-all (SURVEY 8e "partition build" / 8f-4).  The reference groups the
 * transformed rows by partition with a host/disk shuffler (rust/lance-index/src/vector/v3/shuffler.rs:105).
 * For a build sharded by rows over G ranks this call is that shuffle on the device: every rank passes its
 * row-shard index (same model on all ranks, global row ids); afterwards rank g holds ALL rows of the partitions
 * p with p % G == g (its other partitions are empty) in a NEW index.  Inside a partition rows are ordered by
 * source rank, then by the source's storage order -- with contiguous row shards that is the order a single-GPU
 * index has, so a partition is scanned exactly as on one GPU (heap tie order included).  One all-gather of the K
 * partition sizes, one grouped ncclSend/ncclRecv of (codes | vectors, row ids) over NVLink.  Without a
 * communicator it returns a copy.  lb2_index_search_sharded works on the result unchanged: partitions a rank
 * does not own are empty, so every probed partition is scanned by exactly one rank. */
lb2_status lb2_index_repartition(const lb2_index* shard, lb2_index** owned_out);

#ifdef __cplusplus
}
#endif
#endif /* LANCE_B200_H_ */
