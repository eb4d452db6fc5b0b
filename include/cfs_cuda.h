/*
 * cfs_cuda.h -- the C ABI of the B200 symmetric-SpMV path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * The C++ classes of include/ (SparseMatrix, CSRMatrix, SpDMV -- same names and
 * signatures as the reference's include/cfs.hpp) are thin callers of these
 * entry points; a maintainer of the reference would bind exactly these from
 * csr_matrix.tpp (see INTEGRATION.md for the stub).
 *
 * Each entry point names the reference interface it replaces
 * (file:line relative to the reference tree).
 *
 * Conventions
 *   - every function returns CFS_OK (0) or a CFS_ERR_* code; the message of the
 *     last failure on the calling thread is cfs_cuda_last_error().
 *     The C++ layer maps non-zero to the reference's convention: message on
 *     stdout + exit(1) (src/allocator.cpp:31-37, include/io/mmf.hpp:188-189).
 *   - there is NO CPU fallback: without a usable sm_100 device every compute
 *     entry point fails with CFS_ERR_NO_DEVICE.
 *   - indices are 32-bit like the reference's IndexT=int instantiations
 *     (src/cfs.cpp:11-12); values are float or double (is_double).
 *   - one process drives one GPU (cfs_cuda_init); multi-GPU runs are one
 *     process per GPU, each holding a row shard (cfs_cuda_matrix_create_shard).
 */
#ifndef CFS_CUDA_H
#define CFS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFS_OK 0
#define CFS_ERR_CUDA 1      /* a CUDA runtime call failed                    */
#define CFS_ERR_INVALID 2   /* bad argument                                  */
#define CFS_ERR_NO_DEVICE 3 /* no CUDA device / wrong architecture           */
#define CFS_ERR_STATE 4     /* call order (e.g. spmv before tune)            */
#define CFS_ERR_TOO_LARGE 5 /* reference-compatible metadata infeasible      */
#define CFS_ERR_NEEDS_HOST 6 /* input only the host loader may judge (errors
                                the reference reports itself, > 4 GiB text)   */

#define CFS_TUNING_NONE 0       /* util::Tuning::None       platform.hpp:22 */
#define CFS_TUNING_AGGRESSIVE 1 /* util::Tuning::Aggressive platform.hpp:22 */

typedef struct cfs_matrix_s *cfs_mat_t;

/* ---- runtime: replaces include/utils/runtime.hpp:22 / src/runtime.cpp:10-21
 * (get_num_threads picks the partition count; here the process also picks
 * its GPU). */
int cfs_cuda_device_count(int *count);
int cfs_cuda_init(int device);
const char *cfs_cuda_last_error(void);
const char *cfs_cuda_version(void);
/* run-time tunables; unknown key or bad value: CFS_ERR_INVALID. The
 * environment variable CFS_GPU_OPTIONS="key=value,key=value" applies the same
 * settings at cfs_cuda_init (for the reference's own bench / test binaries).
 *   "spmv_variant"   1 = one warp per slice, direct loads; 2 = persistent
 *                    TMA-staged kernel; 4 = shared-memory x/y windows; 5 =
 *                    compressed index stream + shuffle-merged REDs (default; on
 *                    irregular matrices with bounded column windows it hands
 *                    over to 6); 6 = transposed term transposed through shared
 *                    memory; 7 = 5 with every slice's value block staged in
 *                    shared memory by one TMA bulk copy per warp
 *   "l2_prefetch"    0/1  variant 5 with streamed values: one L2 prefetch per
 *                    128-byte line of the slice's value block up front
 *   "reg_blocks"     16/12  resident 128-thread CTAs per SM variant 5 asks for
 *   "deterministic"  0/1  y bitwise reproducible from run to run (integer
 *                    reductions of contributions rounded once; see det.cu)
 *   "keep_layouts"   0/1  (tune time) keep the layouts of the kernel variants
 *                    that are not selected and the P = 1 lower CSR; default:
 *                    only for matrices below 4 M entries. Needed to switch
 *                    spmv_variant after tune or to export CFS_META_LOWER_* /
 *                    CFS_META_SELL_COL on a large matrix
 *   "managed_prefetch", "managed_advise"   unified-memory vectors, see
 *                    cfs_cuda_host_alloc
 *   "tile6"          0/1  allow variant 6 (read at tune and at launch time)
 *   "value_index"    0/1  dictionary-coded values where the lower triangle has
 *                    <= 256 distinct values (lossless; tune and launch time)
 *   "sort_rows", "rechunk", "rechunk_pct", "hubs"   layout of ragged matrices
 *                    (tune time): length-sorting, virtual rows of rechunk_pct %
 *                    of the mean row length, column-wise hub columns
 *   "csr_layout"     0/1  Format::csr streams the sliced layout (default) or
 *                    runs the warp-per-row comparator kernel
 *   "pipeline", "pipeline_chunks", "pipeline_adaptive", "pipeline_split",
 *   "pipeline_taper", "pipeline_graph"
 *                    host-vector path of cfs_cuda_spmv: staged H2D / kernel /
 *                    D2H pipeline, its chunk count, head + rest stages, replay
 *                    as one CUDA graph
 *   "cg_batch"       iterations cfs_cuda_cg_solve enqueues between two looks
 *                    at the stop flag
 *   "ctas_per_sm"    persistent kernel (variant 2)
 *   "diag_mode", "pipeline_skip", "pipeline_smem", "pipeline_trace"
 *                    measurement aids (non-zero diag_mode / pipeline_skip
 *                    compute WRONG results) */
int cfs_cuda_set_option(const char *key, long long value);

/* ---- allocator: backs internal_alloc / internal_free
 * (include/utils/allocator.hpp:11-12, src/allocator.cpp:8-43).
 * 64-byte aligned like the reference. The reference's bench / test take x and y
 * from internal_alloc, fill x on the host, call y = A x in a loop that never
 * looks at the vectors (bench_spmv_mmf.cpp:154-167) and read y on the host
 * afterwards (test_spmv_mmf.cpp:94-104). So that such a caller runs at kernel
 * speed the default kind is UNIFIED (managed) memory: cfs_cuda_spmv launches
 * straight on it, the vectors live in HBM while the GPU works on them and come
 * back page by page when the host touches them. Kinds:
 *   CFS_ALLOC_MANAGED  cudaMallocManaged, preferred location = the process's GPU
 *   CFS_ALLOC_PINNED   page-locked host memory: cfs_cuda_spmv copies x in and y
 *                      out on every call (pipelined)
 *   CFS_ALLOC_PLAIN    posix_memalign (what the reference returns); the choice
 *                      for arrays that are uploaded once (the CSR of the host
 *                      loader)
 *   CFS_ALLOC_DEFAULT  the environment's CFS_GPU_ALLOC=managed|pinned|plain, else
 *                      managed (buffers below 64 KB: plain); plain when no GPU
 *                      is visible
 * cfs_cuda_host_alloc(bytes) = cfs_cuda_host_alloc_kind(bytes, CFS_ALLOC_DEFAULT).
 * Returns NULL on failure (the C++ wrapper prints and exit(1)s like the
 * reference). cfs_cuda_host_free takes any of them (and, like the reference's
 * free(), pointers it has never seen). */
#define CFS_ALLOC_DEFAULT 0
#define CFS_ALLOC_PLAIN 1
#define CFS_ALLOC_PINNED 2
#define CFS_ALLOC_MANAGED 3
void *cfs_cuda_host_alloc(size_t bytes);
void *cfs_cuda_host_alloc_kind(size_t bytes, int kind);
void cfs_cuda_host_free(void *ptr);
/* Where a vector should be before its next use, for callers that know (optional;
 * correctness never depends on it): to_device != 0 moves a managed range to the
 * GPU, 0 brings it to the host, in one bulk transfer instead of page faults.
 * No-op for the other kinds. */
int cfs_cuda_vector_prefetch(const void *ptr, size_t bytes, int to_device);

/* ---- matrix construction: replaces CSRMatrix(rowptr, colind, values, nrows,
 * ncols, symmetric, ...) (include/matrix/csr_matrix.tpp:114-144) and is what
 * the file constructor (csr_matrix.tpp:9-111) calls after the MMF parse.
 * FULL (expanded) CSR, 0-based. rowptr/colind/values may be host or device
 * pointers; host arrays are copied to the GPU, device arrays are borrowed until
 * cfs_cuda_matrix_tune returns. */
int cfs_cuda_matrix_create(cfs_mat_t *out, int32_t nrows, int32_t ncols,
                           const int32_t *rowptr, const int32_t *colind,
                           const void *values, int is_double, int symmetric);

/* ---- Matrix Market ingest on the GPU: replaces, for CSRMatrix(filename, ...),
 * the loader MMF<I,V> (include/io/mmf.hpp:179-343, src/mmf.cpp:6-44) and the
 * CSR fill (include/matrix/csr_matrix.tpp:74-107). The caller has read the
 * banner / comment / size lines (host code, cfs_host.h); the entry lines are
 * tokenised, converted (atoi / atof semantics, values correctly rounded like
 * strtod), mirrored when the file is symmetric, ordered by (row, col) and
 * turned into the full 0-based CSR on the device. The matrix owns that CSR;
 * cfs_cuda_matrix_tune consumes it like one made by cfs_cuda_matrix_create.
 * `symmetric` is the caller's wish (Format::sss); a general file quietly gives
 * plain CSR (csr_matrix.tpp:13-15). Returns CFS_ERR_NEEDS_HOST for anything
 * the reference reports as a fatal input error: the host loader then prints
 * the reference's message. */
typedef struct cfs_mmf_text {
  const char *text;       /* file image in host memory                        */
  size_t bytes;
  size_t entries_offset;  /* first byte of the first entry line               */
  int64_t declared;       /* entry lines announced by the size line           */
  int32_t nrows, ncols;
  int32_t file_symmetric; /* banner: symmetric -> mirror off-diagonal entries */
  int32_t zero_based;     /* banner token base-0                              */
} cfs_mmf_text;
typedef struct cfs_mmf_report {
  int64_t nnz;        /* expanded entry count = CSRMatrix::nnz()              */
  int64_t host_lines; /* lines handed to the host's strtol / strtod           */
  float ms_upload, ms_parse, ms_sort, ms_build; /* device time per phase      */
} cfs_mmf_report;
int cfs_cuda_matrix_create_from_mmf(cfs_mat_t *out, const cfs_mmf_text *in,
                                    int is_double, int symmetric,
                                    cfs_mmf_report *report);
/* The full CSR of a matrix back to the host (CSRMatrix::rowptr() / colind() /
 * values(), include/matrix/csr_matrix.hpp:73-75). NULL destinations are
 * skipped. CFS_ERR_STATE once tune() has released the full CSR of a symmetric
 * matrix (csr_matrix.tpp:1700-1706). */
int cfs_cuda_matrix_download_csr(cfs_mat_t m, int32_t *rowptr, int32_t *colind,
                                 void *values);

/* Multi-GPU row shard: this GPU owns global rows [row_begin, row_end) of an
 * (global_nrows x global_nrows) symmetric matrix (the reference's thread
 * partition, csr_matrix.tpp:404-435, lifted to GPUs). rowptr is shard-local
 * (row_end-row_begin+1 entries, starting at 0), colind holds GLOBAL ids. */
int cfs_cuda_matrix_create_shard(cfs_mat_t *out, int32_t global_nrows,
                                 int32_t row_begin, int32_t row_end,
                                 const int32_t *rowptr, const int32_t *colind,
                                 const void *values, int is_double);

/* ---- preprocessing: replaces CSRMatrix::tune (csr_matrix.tpp:231-310), i.e.
 * partition_by_nrows (404-435) + compress_symmetry (1684-1716) +
 * conflict_free_aposteriori (1206-1639) + color_greedy (2010-2213), all on the
 * GPU. nparts is the reference's P (= CFS_NUM_THREADS): it fixes the
 * reference-compatible metadata (row_split, per-partition lower CSR, conflict
 * graph, colours, ranges) that cfs_cuda_matrix_export returns bit-exactly.
 * The GPU execution layout (sliced, nnz-balanced lower triangle) is built in
 * the same call. Returns CFS_ERR_TOO_LARGE if the conflict graph of the
 * reference is infeasible for this input (the SpMV path is still usable). */
int cfs_cuda_matrix_tune(cfs_mat_t m, int nparts, int tuning);

/* ---- Format::hyb: replaces split_by_bandwidth (csr_matrix.tpp:314-401) and the
 * HYB kernels (:3031-3162), which the reference cannot reach (its tune() aborts
 * for P > 1 and switches HYB off for P = 1). Call between create
 * and tune on a symmetric matrix: entries with |col - row| >= threshold
 * (HybBwThreshold = 10000, csr_matrix.hpp:92) are kept in BOTH triangles and
 * only gathered (the Format::csr kernel adds them onto y); the band inside the
 * threshold runs through the whole symmetric pipeline. nnz(), size() follow the
 * reference's formulas; metadata export describes the near part. */
int cfs_cuda_matrix_set_hybrid(cfs_mat_t m, int32_t bandwidth_threshold);

void cfs_cuda_matrix_destroy(cfs_mat_t m); /* ~CSRMatrix, csr_matrix.tpp:147 */

typedef struct cfs_matrix_info {
  int32_t nrows, ncols;     /* of this shard's owned rows / global columns   */
  int32_t row_begin;        /* first owned global row (0 unless sharded)     */
  int32_t halo_begin;       /* smallest global column referenced (<= row_begin) */
  int64_t nnz_full;         /* nnz() of the reference: expanded count        */
  int64_t nnz_low, nnz_diag;
  int32_t nparts, ncolors, nranges;
  int32_t symmetric, is_double, tuned, refmeta; /* refmeta: metadata present  */
  int64_t size_bytes;       /* size() of the reference, csr_matrix.tpp:191    */
  int64_t device_bytes;     /* bytes this matrix really holds in HBM          */
  int64_t algorithmic_bytes;/* bytes one SpMV must move       */
  int64_t nvrows, nslices, padded_entries; /* execution layout statistics    */
  int64_t nconflict_edges;
  int64_t ntiles;      /* tiles of the persistent kernels                     */
  int64_t far_entries; /* entries outside the shared-memory windows (variant 3) */
  int64_t regular_slices; /* slices whose column stream is compressed to bases */
  int64_t index_rows;     /* 128-byte rows of the compressed column stream    */
  int64_t hub_columns;    /* columns whose transposed term runs column-wise   */
  int64_t hub_entries;    /* lower entries in those columns                   */
  int64_t sort_window;    /* 0 = natural row order; else rows were sorted by
                             length inside windows of this many rows          */
  int64_t transposed_tiles; /* tiles whose transposed term goes through shared
                               memory (variant 6); 0 = variant not applicable */
  int64_t tile_smem_bytes;  /* shared memory of the largest such tile         */
  int64_t value_dictionary; /* distinct values of the lower triangle when the
                               value stream is dictionary-coded (<= 256), else 0 */
  int64_t hyb_far_entries;  /* Format::hyb: entries (both triangles) kept in the
                               non-symmetric far part; 0 = none / not hybrid     */
} cfs_matrix_info;

int cfs_cuda_matrix_info(cfs_mat_t m, cfs_matrix_info *info);

/* ---- SpMV: replaces dense_vector_multiply / spmv_fn
 * (include/matrix/csr_matrix.hpp:67-70), i.e. cpu_mv_sym_conflict_free_v2
 * (csr_matrix.tpp:2966-3028), cpu_mv_sym_serial (2707-2729) and, for
 * symmetric == 0, cpu_mv (2684-2704).
 * cfs_cuda_spmv is synchronous and accepts host, pinned or device pointers
 * (y fully overwritten on return, like the reference).
 * cfs_cuda_spmv_async takes DEVICE vectors and a cudaStream_t; for a shard the
 * vectors are the extended local vectors covering global rows
 * [halo_begin, row_end). */
int cfs_cuda_spmv(cfs_mat_t m, void *y, const void *x);
int cfs_cuda_spmv_async(cfs_mat_t m, void *y_dev, const void *x_dev,
                        void *stream);

/* Multi-GPU SpMV fused with its exchange step. The transposed contributions to
 * columns below row_begin -- the reference's "direct conflicts" with the
 * partitions below (csr_matrix.tpp:1441-1451) -- are reduced by the kernel
 * straight into the y vector of the GPU that owns those rows, over NVLink:
 * y_lower_base is that (peer-mapped, e.g. symmetric-memory) vector's address
 * minus its halo_begin elements, so that y_lower_base[col] is y[col] there.
 * Requires that this shard's halo lies inside the row block of ONE GPU below.
 * The caller zeroes y on every GPU and synchronises the GPUs before and after
 * (y_is_zero != 0 skips the internal memset). */
int cfs_cuda_spmv_halo_async(cfs_mat_t m, void *y_dev, const void *x_dev,
                             void *y_lower_base, int y_is_zero, void *stream);

/* The same with the two things that take the step's skeleton off the critical
 * path (the kernel is 150 us; a y initialisation, an x-halo copy and a second
 * barrier around it cost 45-60 us more in round 1):
 *   x_lower_base: the x vector of the GPU below as a peer-mapped virtual base
 *     (like y_lower_base): halo entries of x are read from there by the kernel,
 *     no copy into the local halo beforehand; NULL: the local halo part of x_dev
 *     holds them;
 *   y_clear: a second vector like y_dev whose OWNED rows the kernel clears (one
 *     store per row by the lane that owns the row): with two result vectors
 *     used alternately, SpMV k clears the one SpMV k+1 reduces into, and no
 *     separate initialisation is left -- one synchronisation of the GPUs per
 *     SpMV (after it: reductions landed, other vector clear, x final) is then
 *     enough. NULL: no clearing.
 * Also usable on an unsharded matrix (y_lower_base = x_lower_base = NULL) to
 * drop the y initialisation from a loop of SpMVs. */
int cfs_cuda_spmv_shard_async(cfs_mat_t m, void *y_dev, const void *x_dev,
                              void *y_lower_base, const void *x_lower_base,
                              void *y_clear, int y_is_zero, void *stream);

/* The two parts of that step as separate launches, for callers that overlap
 * the synchronisation of the GPUs with the bulk of the work: part 1 = the
 * leading slices that reference columns below row_begin (they read the x of
 * and reduce into the y of the GPU below: y_lower_base / x_lower_base as
 * above; nothing to do on the lowest GPU), part 2 = all other slices (no
 * remote access; pass the same y_lower_base so that the split point is the
 * same, it is not dereferenced). y must be clear on entry; both parts clear
 * the rows they own in y_clear. cfs_spmv_b200/dist.py (P2PHalo.step) runs part
 * 2 on the main stream while a side stream runs barrier -> part 1. */
int cfs_cuda_spmv_shard_part_async(cfs_mat_t m, void *y_dev, const void *x_dev,
                                   void *y_lower_base, const void *x_lower_base,
                                   void *y_clear, int part, void *stream);

/* ---- one process, several GPUs: replaces the role of get_num_threads
 * (src/runtime.cpp:10-21) + partition_by_nnz (csr_matrix.tpp:438-541) for the
 * GPUs of one box. The C++ layer takes this path when CFS_NUM_GPUS > 1
 * (include/utils/runtime.hpp: get_num_gpus): the rows of a symmetric matrix are
 * cut into ngpus contiguous, 16-row aligned blocks with about the same number
 * of stored lower-triangle entries, GPU first_device + g holds the shard of
 * block g, and cfs_cuda_multi_spmv computes y = A x with HOST, managed or
 * device vectors of full length: every GPU fetches the piece of x it needs,
 * runs its kernel and writes its rows of y, concurrently. Contributions that
 * cross a block boundary are reduced straight into the y of the GPU below over
 * NVLink peer access when every halo lies inside that one block (banded /
 * stencil matrices, fused_halo = 1), else the owners add the strips the GPUs
 * above produced for them (a reduce-scatter restricted to the touched ranges;
 * R-MAT). With fused halos and x, y in unified memory (cfs_cuda_host_alloc)
 * the GPUs work on the caller's vectors in place: x advised read-mostly (local
 * duplicates per GPU), the rows of y preferring their owner and mapped into the
 * GPU above; nothing is copied (option multi_zero_copy, default 1).
 * rowptr / colind / values: FULL CSR in host memory, borrowed during the call
 * only. */
typedef struct cfs_multi_s *cfs_multi_t;
#define CFS_MULTI_MAX_GPUS 16
typedef struct cfs_multi_info {
  int32_t ngpus, fused_halo, nrows, reserved;
  int64_t nnz_full, nnz_low;
  int32_t device[CFS_MULTI_MAX_GPUS];
  int32_t row_begin[CFS_MULTI_MAX_GPUS], row_end[CFS_MULTI_MAX_GPUS];
  int32_t halo_begin[CFS_MULTI_MAX_GPUS];
  int64_t shard_nnz_low[CFS_MULTI_MAX_GPUS];
} cfs_multi_info;
int cfs_cuda_multi_create(cfs_multi_t *out, int ngpus, int first_device,
                          int32_t nrows, const int32_t *rowptr,
                          const int32_t *colind, const void *values,
                          int is_double);
int cfs_cuda_multi_tune(cfs_multi_t mm);
int cfs_cuda_multi_spmv(cfs_multi_t mm, void *y, const void *x);
int cfs_cuda_multi_info(cfs_multi_t mm, cfs_multi_info *info);
void cfs_cuda_multi_destroy(cfs_multi_t mm);

/* ---- conjugate gradients on the device: solves A x = b
 * for a tuned symmetric positive definite matrix with the SpMV above as its
 * only matrix operation. The reference has no solver; this is the loop its
 * bench emulates (bench_spmv_mmf.cpp:162-167: y = A x over and over), kept in
 * HBM: q = A p and p'q come out of ONE kernel, the vector updates are two more,
 * an iteration is replayed as a CUDA graph and the step lengths never visit
 * the host. x holds the initial guess on entry and the solution on return; x
 * and b may be host or device vectors of the matrix precision. Stops when
 * ||r_k|| <= rel_tol * ||r_0|| (recurrence residual) or after max_iters.
 * history (optional): ||r_k|| for k = 0..iterations, at most
 * history_capacity values. */
typedef struct cfs_cg_result {
  int32_t iterations; /* iterations until the stop criterion held (or max)  */
  int32_t executed;   /* iterations enqueued (whole batches, >= iterations) */
  int32_t converged;  /* ||r|| <= rel_tol * ||r0||                           */
  int32_t breakdown;  /* p'Ap <= 0 met: A is not positive definite           */
  double initial_residual_norm, residual_norm;
  float ms_total;     /* device time of the whole solve                      */
} cfs_cg_result;
int cfs_cuda_cg_solve(cfs_mat_t m, void *x, const void *b, int max_iters,
                      double rel_tol, cfs_cg_result *result, double *history,
                      int history_capacity);

/* Building blocks of the same loop over row SHARDS (one process per GPU): the
 * caller sums the two scalars of an iteration across its ranks (a 1-double
 * all-reduce each: the exchange step conjugate gradients really have) and the
 * vector work stays in these kernels. cfs_spmv_b200/dist.py (DistributedCG) is
 * the loop.
 *   cfs_cuda_spmv_halo_dot_async: cfs_cuda_spmv_halo_async that also ADDS this
 *     shard's part of x'(A x) to *dot_dev (every stored entry lives in exactly
 *     one shard, so the parts add up to the global value);
 *   cfs_cuda_cg_update_xr: alpha = scal[0] / scal[1] (r'r, p'Ap: global sums);
 *     x += alpha p, r -= alpha q, *rr_next_dev += this shard's r'r;
 *   cfs_cuda_cg_update_p: beta = scal[2] / scal[0]; p = r + beta p.
 * All pointers are device pointers over the shard's OWNED rows (n of them);
 * scal is a device array of 3 doubles {r'r, p'Ap, next r'r}. */
int cfs_cuda_spmv_halo_dot_async(cfs_mat_t m, void *y_dev, const void *x_dev,
                                 void *y_lower_base, int y_is_zero,
                                 double *dot_dev, void *stream);
int cfs_cuda_cg_update_xr(int64_t n, int is_double, const double *scal,
                          const void *p, const void *q, void *x, void *r,
                          double *rr_next_dev, void *stream);
int cfs_cuda_cg_update_p(int64_t n, int is_double, const double *scal,
                         const void *r, void *p, void *stream);

/* Measurement aid for bench.py (bench_spmv_mmf.cpp:162-167 times the same
 * loop with omp_get_wtime): runs `iters` SpMVs back to back on `stream` between
 * two CUDA events (total_ms: y initialisation + kernel, summed over the loop),
 * then the y initialisation alone the same way; kernel_ms = total_ms minus
 * that, i.e. the summed device time of the SpMV KERNEL, which by construction
 * cannot exceed the step time. Synchronises the stream; y = A x on return. */
int cfs_cuda_spmv_timed(cfs_mat_t m, void *y_dev, const void *x_dev,
                        void *stream, int iters, float *total_ms,
                        float *kernel_ms);

/* ---- metadata export (device -> host), for bit-exact parity with the
 * reference's private members (csr_matrix.hpp:96-124, 221-277).
 * dst == NULL: only *count is written. Element types: int32 except
 * LOWER_VALUES / DIAGONAL / SELL_VAL (matrix precision). */
enum {
  CFS_META_ROW_SPLIT = 1,  /* row_split_[P+1]                                */
  CFS_META_PART_NNZ_LOW,   /* SymThreadData::nnz_low_ per partition          */
  CFS_META_LOWER_ROWPTR,   /* per-partition local rowptr_, concatenated N+P   */
  CFS_META_LOWER_COLIND,   /* colind_ (global ids), concatenated             */
  CFS_META_LOWER_VALUES,   /* values_, concatenated                          */
  CFS_META_DIAGONAL,       /* diagonal_, concatenated (= N)                  */
  CFS_META_WEIGHT,         /* WeightedVertex::weight per 16-row block        */
  CFS_META_ADJ_PTR,        /* conflict graph, CSR with ascending neighbours  */
  CFS_META_ADJ,
  CFS_META_COLOR_FIRST,    /* colours after first-fit                        */
  CFS_META_COLOR,          /* colours after balancing (color_map)            */
  CFS_META_RANGE_PTR,      /* range_ptr_ per partition, P*(ncolors+1)        */
  CFS_META_PART_NRANGES,   /* SymThreadData::nranges_                        */
  CFS_META_RANGE_START,    /* range_start_ (local rows), concatenated        */
  CFS_META_RANGE_END,      /* range_end_ (inclusive), concatenated           */
  CFS_META_SELL_SLICE_PTR = 100, /* execution layout, for property tests     */
  CFS_META_SELL_VROW,
  CFS_META_SELL_COL,
  CFS_META_SELL_VAL
};
int cfs_cuda_matrix_export(cfs_mat_t m, int what, void *dst_host,
                           size_t capacity_elems, size_t *count);

/* ---- synthetic inputs (bench/test plumbing; cfs_gen.h holds the definition
 * shared with the oracle). Device-side so that a GPU can build its own shard
 * of the BASELINE.json configs in place. */
struct cfs_gen_spec;
int cfs_gen_host_count(const struct cfs_gen_spec *spec, int64_t row_begin,
                       int64_t row_end, int32_t *rowptr);
int cfs_gen_host_fill(const struct cfs_gen_spec *spec, int64_t row_begin,
                      int64_t row_end, const int32_t *rowptr, int32_t *colind,
                      void *values, int is_double);
int cfs_gen_host_x(uint64_t seed, int64_t begin, int64_t end, void *x,
                   int is_double);
int cfs_cuda_gen_count(const struct cfs_gen_spec *spec, int64_t row_begin,
                       int64_t row_end, int32_t *rowptr_dev, int64_t *nnz);
int cfs_cuda_gen_fill(const struct cfs_gen_spec *spec, int64_t row_begin,
                      int64_t row_end, const int32_t *rowptr_dev,
                      int32_t *colind_dev, void *values_dev, int is_double);
int cfs_cuda_gen_x(uint64_t seed, int64_t begin, int64_t end, void *x_dev,
                   int is_double);

#ifdef __cplusplus
}
#endif
#endif /* CFS_CUDA_H */
