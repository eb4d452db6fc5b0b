// compress.cu -- index-stream compression for regular slices (variant 5).
//
// A slice (32 virtual rows x w steps) is REGULAR when its 32 lanes hold 32
// consecutive, unsplit rows of equal length w and every step is unit-stride
// across the lanes: col[k][lane] = col[k][0] + lane. That is what the interior
// of any stencil / structured-grid matrix looks like under a natural ordering.
// For such a slice the column stream degenerates to w base columns: it is
// stored as ONE 32-int row (lane k = base of step k) instead of w rows, which
// removes ~4 of the 12 bytes/entry the kernel would otherwise stream from HBM,
// and it tells the kernel which steps hit shifted copies of the same y range,
// so their transposed-term updates can be merged with warp shuffles before one
// RED (spmv_reg.cuh). Every other slice keeps its full column rows.
//
// Output: ccol (compressed column stream, rows of 32 ints) and slice_cptr
// (first row of each slice in ccol; bit 30 set = regular).
#include <cub/cub.cuh>

#include "common.cuh"

namespace cfsb {

namespace {

constexpr int kThreads = 256;
inline unsigned blocks_for(size_t n) {
  return (unsigned)((n + kThreads - 1) / kThreads);
}

__global__ void classify_slices_kernel(long long nslices,
                                       const int *__restrict__ slice_ptr,
                                       const int *__restrict__ vrow_row,
                                       const int *__restrict__ sell_col,
                                       int *__restrict__ rows_needed,
                                       unsigned long long *__restrict__ nregular) {
  const long long s = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= nslices)
    return;
  const int p0 = slice_ptr[s], w = slice_ptr[s + 1] - p0;
  const int tag = vrow_row[s * kSliceRows + lane];
  const int tag0 = __shfl_sync(0xffffffffu, tag, 0);
  // consecutive, live, unsplit rows
  bool ok = tag0 >= 0 && !(tag0 & kVrowCont) && tag == tag0 + lane;
  // every step unit-stride, no padding; w <= 32 so the bases fit in one row
  ok = __all_sync(0xffffffffu, ok) && w >= 1 && w <= kSliceRows;
  const int *cp = sell_col + (size_t)p0 * kSliceRows + lane;
  for (int k = 0; k < w && ok; ++k) {
    const int c = cp[(size_t)k * kSliceRows];
    const int c0 = __shfl_sync(0xffffffffu, c, 0);
    ok = c0 >= 0 && c == c0 + lane;
    ok = __all_sync(0xffffffffu, ok);
  }
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) {
    rows_needed[s] = ok ? (1 | kSliceRegular) : w;
    if (ok)
      atomicAdd(nregular, 1ull);
  }
}

__global__ void strip_flag_kernel(long long n, const int *__restrict__ in,
                                  int *__restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = in[i] & ~kSliceRegular;
}

__global__ void fill_ccol_kernel(long long nslices,
                                 const int *__restrict__ slice_ptr,
                                 const int *__restrict__ rows_needed,
                                 const int *__restrict__ sell_col,
                                 int *__restrict__ slice_cptr,
                                 int *__restrict__ ccol) {
  const long long s = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= nslices)
    return;
  const int p0 = slice_ptr[s], w = slice_ptr[s + 1] - p0;
  const int at = slice_cptr[s]; // exclusive scan of the row counts
  const bool regular = (rows_needed[s] & kSliceRegular) != 0;
  const int *cp = sell_col + (size_t)p0 * kSliceRows;
  int *out = ccol + (size_t)at * kSliceRows;
  if (regular) {
    out[lane] = lane < w ? cp[(size_t)lane * kSliceRows] : -1; // bases
  } else {
    for (int k = 0; k < w; ++k)
      out[(size_t)k * kSliceRows + lane] = cp[(size_t)k * kSliceRows + lane];
  }
  __syncwarp();
  if (lane == 0 && regular)
    slice_cptr[s] = at | kSliceRegular;
}

} // namespace

int build_compressed_cols(cfs_matrix_s *m, cudaStream_t s) {
  m->nregular = 0;
  const long long ns = m->nslices;
  if (ns == 0)
    return CFS_OK;
  DevArray<int> need, need_plain;
  DevArray<unsigned long long> nreg;
  CFS_TRY(need.alloc((size_t)ns + 1));
  CFS_TRY(need_plain.alloc((size_t)ns + 1));
  CFS_TRY(nreg.alloc(1));
  CFS_CUDA_TRY(cudaMemsetAsync(need.p, 0, ((size_t)ns + 1) * 4, s));
  CFS_CUDA_TRY(cudaMemsetAsync(nreg.p, 0, 8, s));
  classify_slices_kernel<<<blocks_for((size_t)ns * 32), kThreads, 0, s>>>(
      ns, m->slice_ptr.p, m->vrow_row.p, m->sell_col.p, need.p, nreg.p);
  strip_flag_kernel<<<blocks_for((size_t)ns + 1), kThreads, 0, s>>>(
      ns + 1, need.p, need_plain.p);
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_TRY(m->slice_cptr.alloc((size_t)ns + 1));
  size_t tb = 0;
  CFS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, need_plain.p,
                                             m->slice_cptr.p, ns + 1, s));
  DevArray<char> tmp;
  CFS_TRY(tmp.alloc(tb));
  CFS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tb, need_plain.p,
                                             m->slice_cptr.p, ns + 1, s));
  int total_rows = 0;
  unsigned long long hreg = 0;
  CFS_CUDA_TRY(cudaMemcpyAsync(&total_rows, m->slice_cptr.p + ns, 4,
                               cudaMemcpyDeviceToHost, s));
  CFS_CUDA_TRY(cudaMemcpyAsync(&hreg, nreg.p, 8, cudaMemcpyDeviceToHost, s));
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  m->nregular = (int64_t)hreg;
  m->ccol_rows = total_rows;
  CFS_TRY(m->ccol.alloc((size_t)total_rows * kSliceRows));
  fill_ccol_kernel<<<blocks_for((size_t)ns * 32), kThreads, 0, s>>>(
      ns, m->slice_ptr.p, need.p, m->sell_col.p, m->slice_cptr.p, m->ccol.p);
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  return CFS_OK;
}

} // namespace cfsb
