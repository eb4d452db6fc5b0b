// common.cuh -- shared declarations of the CUDA side of the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>
#include <vector>

#include "cfs_cuda.h"

namespace cfsb {

// ---- error plumbing ------------------------------------------------------
void set_error(const char *fmt, ...);
// binds the process to its device on first use (cfs_cuda_init(0) by default);
// CFS_ERR_NO_DEVICE without an sm_100 GPU -- there is no CPU fallback
int require_device();
int current_device();
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define CFS_CUDA_TRY(call)                                                     \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess)                                                    \
      return ::cfsb::cuda_fail(e__, #call, __FILE__, __LINE__);                \
  } while (0)

#define CFS_TRY(call)                                                          \
  do {                                                                         \
    int s__ = (call);                                                          \
    if (s__ != CFS_OK)                                                         \
      return s__;                                                              \
  } while (0)

// Reference constants (include/matrix/csr_matrix.hpp:89-90)
constexpr int kBlkBits = 4;
constexpr int kBlkFactor = 1 << kBlkBits;

// Execution layout constants
constexpr int kSliceRows = 32;       // one warp lane per (virtual) row
constexpr int kMaxChunk = 32;        // longest virtual row; longer rows are split
constexpr int kVrowCont = 1 << 30;   // flag: virtual row continues an earlier chunk
constexpr int kVrowRowMask = kVrowCont - 1;
constexpr int kSliceRegular = 1 << 30; // flag in slice_cptr: regular slice
// hub columns (hubs.cu)
constexpr int kHubMinCount = 128; // lower entries that make a column a hub
constexpr int kHubChunk = 1024;   // entries of a hub column per warp
constexpr int kHubFlag = 1 << 30; // in the row stream: transposed term is done
                                  // column-wise, skip the RED
// Persistent TMA-staged kernel: a tile is up to kTileSlices consecutive slices
// whose entries (<= kTileSteps slice-steps of 32) are one contiguous piece of
// the value / index streams, fetched by one bulk copy each.
constexpr int kTileSlices = 4;
constexpr int kTileSteps = 56;

// variant 6 (spmv_tile.cuh, tiles6.cu): tiles of kT6Slices slices whose
// transposed term is transposed through shared memory
constexpr int kT6Slices = 32;          // one warp per slice, 1024 threads
constexpr int kT6MaxCols = 8192;       // columns a tile's window may span
constexpr int kT6MaxSmemBytes = 96 * 1024; // products of a tile (2 CTAs per SM)
constexpr int kMaxDict = 256; // distinct values a one-byte code can name
// partial sums of x'(A x) written by the DOT kernels (cg.cu)
constexpr int kDotSlots = 128;
constexpr int kDotStride = 4; // doubles

constexpr int kMaxWindows = 8;   // x / y windows per tile (variant 3)
constexpr int kWindowBlocks = 28; // 32-column blocks of window space per tile

// Per-tile record, 128 bytes, fetched into shared memory with the tile.
struct __align__(16) TileRec {
  int slice_begin, nslices, step_begin, nsteps;
  int slice_step[8 + 1]; // first step of slice w inside the tile (<= 8 slices)
  int nwin;              // windows in use (variant 3)
  int row_lo;            // first row if the tile's rows are consecutive, else -1
  int pad0;
  int win_lo[kMaxWindows];                // first global column, multiple of 32
  unsigned short win_nblk[kMaxWindows];   // 32-column blocks in the window
  int pad1[4];
};
static_assert(sizeof(TileRec) == 128, "TileRec must stay 128 bytes");

struct Options {
  int spmv_variant = 5;
  int ctas_per_sm = 2;
  int hubs = 1;      // column-wise handling of hub columns of ragged matrices
  int pipeline = 1;  // overlap H2D / kernel / D2H in cfs_cuda_spmv(host, host)
  int pipeline_trace = 0;   // print the per-chunk timeline (development aid)
  int pipeline_graph = 1;   // replay the pipelined step as one CUDA graph
  int pipeline_chunks = 6;  // row chunks of the pipeline (read at tune time)
  int pipeline_skip = 0;    // measurement aid: 1 no kernels, 2 no D2H, 4 no H2D
  int pipeline_smem = 0;    // dynamic smem of pipelined launches (occupancy cap)
  int pipeline_adaptive = 1; // fewer chunks for small vectors (>= 4 MB of x each)
  int pipeline_taper = 1;   // smaller chunks at both ends of the pipeline
  int pipeline_split = 1;   // chunks run as head + rest (see build_pipeline_plan)
  int sort_rows = 1; // allow length-sorting of ragged matrices at tune time
  int csr_layout = 1; // Format::csr streams the sliced layout (0: warp per row)
  int value_index = 1; // dictionary-coded values where <= 256 distinct (regular matrices)
  int rechunk_pct = 130; // ... as a percentage of the mean row length
  int rechunk = 1;     // ragged matrices: virtual rows of ~1.3 x the mean row length
  int tile6 = 1;      // variant 6 where it applies (bounded column windows)
  int multi_graph = 1; // ... and replays the step of all devices as one graph
  int multi_zero_copy = 1; // cfs_cuda_multi_spmv works on managed vectors in place
  int slot_banks = 1; // variant 6: bank-aware slot assignment (tune time)
  int deterministic = 0; // y bitwise reproducible: integer reductions (det.cu)
  int keep_layouts = 0; // keep the layouts of non-selected kernel variants (tune time)
  int reg_blocks = 16; // variant 5: resident 128-thread CTAs per SM asked for (16 or 12)
  int l2_prefetch = 1; // variant 5, streamed values: prefetch.global.L2 per slice
  int managed_prefetch = 1; // managed vectors: 0 never, 1 first use, 2 every call
  int managed_advise = 1;   // managed vectors: preferred location = the GPU
  int cg_batch = 16; // CG iterations enqueued between two looks at the stop flag
  int diag_mode = 0; // measurement aid, see spmv.cu (non-zero: wrong results)
};
extern Options g_options;

template <typename T> struct DevArray {
  T *p = nullptr;
  size_t n = 0;
  DevArray() {}
  DevArray(const DevArray &) = delete;
  DevArray &operator=(const DevArray &) = delete;
  ~DevArray() { release(); }
  int alloc(size_t count) {
    release();
    n = count;
    CFS_CUDA_TRY(cudaMalloc((void **)&p, (count ? count : 1) * sizeof(T)));
    return CFS_OK;
  }
  void release() {
    if (p)
      cudaFree(p);
    p = nullptr;
    n = 0;
  }
  size_t bytes() const { return p ? (n ? n : 1) * sizeof(T) : 0; }
};

} // namespace cfsb

// The opaque handle of the C ABI.
struct cfs_matrix_s {
  int device = 0;
  bool is_double = true, symmetric = true, tuned = false, sharded = false;
  int32_t nrows = 0, ncols = 0;          // owned rows, global columns
  int32_t row_begin = 0, global_nrows = 0;
  int32_t halo_begin = 0;
  int64_t nnz_full = 0, nnz_low = 0, nnz_diag = 0;
  size_t vsize() const { return is_double ? 8 : 4; }

  // full CSR of the owned rows (device). Borrowed or owned.
  const int32_t *csr_rowptr = nullptr;
  const int32_t *csr_colind = nullptr;
  const void *csr_values = nullptr;
  cfsb::DevArray<int32_t> own_rowptr, own_colind;
  cfsb::DevArray<char> own_values;

  // lower triangle in CSR order (reference: SymThreadData, csr_matrix.hpp:221)
  cfsb::DevArray<int32_t> low_rowptr; // nrows+1, shard-global offsets
  cfsb::DevArray<int32_t> low_colind; // global column ids
  cfsb::DevArray<char> low_values;
  cfsb::DevArray<char> diagonal;      // nrows

  // execution layout: sliced ELL over virtual rows
  int64_t nvrows = 0, nslices = 0, padded_entries = 0;
  // Format::hyb (hyb.cu): entries with |col - row| >= hyb_threshold live in
  // `far`, a non-symmetric matrix whose kernel adds onto y after this one's
  int hyb_threshold = 0;
  int64_t hyb_far_entries = 0;
  cfs_matrix_s *far = nullptr;
  int max_slice_steps = 0; // widest slice (steps of 32 entries)
  long long halo_slice_end = 0; // shard: slices beyond never touch the halo
  int max_row_nnz_full = 0; // longest row of the FULL matrix (terms one y gets)
  // deterministic mode (det.cu): 64-bit fixed-point y, {|A|max bits, |x|max
  // bits} and {scale, 1/scale}; all built on first use
  mutable cfsb::DevArray<long long> det_acc;
  mutable cfsb::DevArray<unsigned long long> det_max;
  mutable cfsb::DevArray<double> det_scale;
  mutable bool det_amax_done = false;
  int64_t sort_window = 0; // 0: natural order; else rows sorted by length in windows
  cfsb::DevArray<int32_t> slice_ptr;  // nslices+1, units of 32 entries
  cfsb::DevArray<int32_t> vrow_row;   // nslices*32
  cfsb::DevArray<int32_t> sell_col;   // padded_entries
  cfsb::DevArray<char> sell_val;      // padded_entries
  // compressed column stream (compress.cu): rows of 32 ints; a regular slice
  // keeps one row of step bases
  cfsb::DevArray<int32_t> slice_cptr; // nslices+1; bit 30 = regular
  cfsb::DevArray<int32_t> ccol;
  int64_t nregular = 0, ccol_rows = 0;
  // hub columns of ragged matrices (hubs.cu)
  int64_t nhubs = 0, hub_entries = 0, nhub_chunks = 0;
  cfsb::DevArray<int32_t> hub_ptr, hub_row, hub_colstream;
  cfsb::DevArray<char> hub_val;
  cfsb::DevArray<int4> hub_chunks;    // {column, first entry, end entry, -}
  // value indexing (valindex.cu): ndict == 0: values are streamed
  int ndict = 0;
  cfsb::DevArray<char> vdict;           // kMaxDict values of the matrix precision
  cfsb::DevArray<unsigned char> vcode;  // padded_entries (absent when ndict == 1)
  // variant 6: column-transposed tiles (tiles6.cu); nt6 == 0: not applicable
  int64_t nt6 = 0;
  int t6_smem_entries = 0;                  // products of the largest tile
  int t6_max_cols = 0;                      // widest column window of a tile
  int64_t t6_cptr_entries = 0;
  cfsb::DevArray<unsigned> t6_pack;         // padded_entries: lcol | slot << 16
  cfsb::DevArray<int> t6_lo, t6_ncols;      // per tile: first column, columns
  cfsb::DevArray<long long> t6_cptr_off;    // per tile: offset into t6_cptr
  cfsb::DevArray<unsigned short> t6_cptr;   // per tile: ncols + 1 slot offsets
  int64_t ntiles = 0;
  cfsb::DevArray<cfsb::TileRec> tile_rec;
  // variant 3: index stream rewritten to window slots / far codes
  cfsb::DevArray<int32_t> sell_slot;  // padded_entries
  int64_t far_entries = 0;            // entries left on the global-RED path

  // reference-compatible metadata for P partitions
  int32_t nparts = 1, ncolors = 0, nranges = 0, nblk = 0;
  bool refmeta = false;
  bool part_by_nnz = false; // non-symmetric: row_split came from partition_by_nnz
  int64_t nedges = 0;
  std::vector<int32_t> row_split;     // P+1 (host)
  cfsb::DevArray<int32_t> weight, adj_ptr, adj, color_first, color;
  cfsb::DevArray<int32_t> range_ptr, part_nranges, range_start, range_end;

  // partial sums of x'(A x) (cfs_cuda_spmv_halo_dot_async)
  cfsb::DevArray<double> dot_slots;
  // staging for the synchronous host-pointer entry point
  cfsb::DevArray<char> stage_x, stage_y;
  cudaStream_t stream = nullptr;
  // unified-memory vectors (cfs_cuda_spmv): calls that still prefetch after a
  // call that ran into page faults, and the size of that window
  int managed_hot = 0, managed_window = 0;
  // Host-vector pipeline (cfs_cuda_spmv with host x and y): the lower triangle
  // only looks DOWN, so a chunk of rows can run as soon as x up to its last row
  // has arrived, and its y rows are final once every chunk that reaches down
  // into them is done: H2D, kernel and D2H overlap chunk by chunk.
  // A stage = a range of slices (a whole chunk, or its head / its rest, see
  // build_pipeline_plan): the x rows it needs beyond the earlier stages and
  // the y row ranges that are final once it has run.
  struct Stage {
    long long slice0, slice1;
    int x_row0, x_row1;
    std::vector<std::pair<int, int>> y_ready;
  };
  std::vector<Stage> stages;
  // 0: untuned, 1: per-slice reach on the device, plan pending, 2: plan final
  int plan_state = 0;
  cfsb::DevArray<int> reach_min, reach_rlo, reach_rhi;
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  std::vector<cudaEvent_t> ev_x, ev_k, ev_d;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaGraphExec_t pipe_graph = nullptr; // the step captured for (pipe_x, pipe_y)
  const void *pipe_x = nullptr;
  void *pipe_y = nullptr;
  unsigned long long pipe_generation = 0; // options generation of the capture
};

namespace cfsb {

// preprocessing (preproc.cu)
int build_lower(cfs_matrix_s *m, cudaStream_t s);
int build_layout(cfs_matrix_s *m, cudaStream_t s);
int build_layout_from(cfs_matrix_s *m, const int32_t *src_rowptr,
                      const int32_t *src_colind, const void *src_values,
                      int64_t src_nnz, cudaStream_t s, int chunk = kMaxChunk,
                      bool rechunk = false);
// the non-symmetric path, Format::csr (csr_path.cu)
int tune_csr(cfs_matrix_s *m, int nparts, int tuning, cudaStream_t s);
// accumulate: add onto y instead of overwriting it (the far part of Format::hyb)
int launch_csr_sell(const cfs_matrix_s *m, void *y, const void *x,
                    cudaStream_t s, bool accumulate = false);
// Format::hyb (hyb.cu)
int split_hybrid(cfs_matrix_s *m, cudaStream_t s);
// x / y windows of the tiles (windows.cu)
int build_windows(cfs_matrix_s *m, cudaStream_t s);
// index-stream compression of regular slices (compress.cu)
int build_compressed_cols(cfs_matrix_s *m, cudaStream_t s);
// value dictionary + one-byte codes (valindex.cu)
int build_value_index(cfs_matrix_s *m, cudaStream_t s);
// column-transposed tiles of variant 6 (tiles6.cu)
int build_tiles6(cfs_matrix_s *m, cudaStream_t s);
// hub columns (hubs.cu)
int build_hubs(cfs_matrix_s *m, cudaStream_t s);
int launch_hub_spmv(const cfs_matrix_s *m, void *y, const void *x,
                    cudaStream_t s);
// reference metadata (refmeta.cu)
int build_refmeta(cfs_matrix_s *m, cudaStream_t s);
// kernels (spmv.cu)
// ev0/ev1 (optional) are recorded directly before/after the kernel launch
// y_lower_base: virtual base of the y vector of the GPU below (fused halo
// reduction over NVLink) or nullptr; y_is_zero: the caller cleared y already;
// x_lower_base: virtual base of the x vector of the GPU below (halo entries of
// x are then read from there instead of the local halo part); y_clear: a
// vector like y_ext whose OWNED rows the kernel clears (ping-pong results)
// xdoty: kDotSlots partial sums (stride kDotStride doubles) that the kernel
// adds x'(A x) into (the caller zeroes them), or nullptr
int launch_sym_spmv(const cfs_matrix_s *m, void *y_ext, const void *x_ext,
                    cudaStream_t s, cudaEvent_t ev0 = nullptr,
                    cudaEvent_t ev1 = nullptr, void *y_lower_base = nullptr,
                    bool y_is_zero = false, long long slice0 = 0,
                    long long slice1 = -1, double *xdoty = nullptr,
                    const void *x_lower_base = nullptr,
                    void *y_clear = nullptr);
// deterministic mode (det.cu): before / after the kernel
int det_prepare(const cfs_matrix_s *m, const void *x_ext, cudaStream_t s);
int det_finish(const cfs_matrix_s *m, void *y_ext, cudaStream_t s);
int build_halo_extent(cfs_matrix_s *m, cudaStream_t s);
// host-vector pipeline plan (preproc.cu): tune-time half and first-use half
int build_pipeline_reach(cfs_matrix_s *m, cudaStream_t s);
int build_pipeline_plan(cfs_matrix_s *m, cudaStream_t s);
int launch_csr_spmv(const cfs_matrix_s *m, void *y, const void *x,
                    cudaStream_t s);

} // namespace cfsb
