// common.cuh -- shared declarations of the CUDA side of the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "cfs_cuda.h"

namespace cfsb {

// ---- error plumbing ------------------------------------------------------
void set_error(const char *fmt, ...);
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
constexpr int kMaxChunk = 64;        // longest virtual row; longer rows are split
constexpr int kVrowCont = 1 << 30;   // flag: virtual row continues an earlier chunk
constexpr int kVrowRowMask = kVrowCont - 1;
// Persistent TMA-staged kernel: a tile is up to kTileSlices consecutive slices
// whose entries (<= kTileSteps slice-steps of 32) are one contiguous piece of
// the value / index streams, fetched by one bulk copy each.
constexpr int kTileSlices = 8;
constexpr int kTileSteps = 128;

struct Options {
  int spmv_variant = 1;
  int ctas_per_sm = 2;
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
  cfsb::DevArray<int32_t> slice_ptr;  // nslices+1, units of 32 entries
  cfsb::DevArray<int32_t> vrow_row;   // nslices*32
  cfsb::DevArray<int32_t> sell_col;   // padded_entries
  cfsb::DevArray<char> sell_val;      // padded_entries
  int64_t ntiles = 0;
  cfsb::DevArray<int4> tile_info;     // {first slice, #slices, first step, #steps}

  // reference-compatible metadata for P partitions
  int32_t nparts = 1, ncolors = 0, nranges = 0, nblk = 0;
  bool refmeta = false;
  int64_t nedges = 0;
  std::vector<int32_t> row_split;     // P+1 (host)
  cfsb::DevArray<int32_t> weight, adj_ptr, adj, color_first, color;
  cfsb::DevArray<int32_t> range_ptr, part_nranges, range_start, range_end;

  // staging for the synchronous host-pointer entry point
  cfsb::DevArray<char> stage_x, stage_y;
  cudaStream_t stream = nullptr;
};

namespace cfsb {

// preprocessing (preproc.cu)
int build_lower(cfs_matrix_s *m, cudaStream_t s);
int build_layout(cfs_matrix_s *m, cudaStream_t s);
// reference metadata (refmeta.cu)
int build_refmeta(cfs_matrix_s *m, cudaStream_t s);
// kernels (spmv.cu)
// ev0/ev1 (optional) are recorded directly before/after the kernel launch
int launch_sym_spmv(const cfs_matrix_s *m, void *y_ext, const void *x_ext,
                    cudaStream_t s, cudaEvent_t ev0 = nullptr,
                    cudaEvent_t ev1 = nullptr);
int launch_csr_spmv(const cfs_matrix_s *m, void *y, const void *x,
                    cudaStream_t s);

} // namespace cfsb
