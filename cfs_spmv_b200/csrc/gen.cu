// gen.cu -- synthetic inputs (bench / test plumbing). The matrix definitions
// live in cfs_gen.h and are shared with the oracle and the compiled reference
// harness; here they are evaluated on the host (OpenMP-free, plain loops) or by
// one GPU thread per row, so that each GPU can build its own row shard of the
// BASELINE.json configs directly in HBM.
#include <cub/cub.cuh>

#include "cfs_gen.h"
#include "common.cuh"

namespace {

__global__ void gen_count_kernel(cfs_gen_spec spec, long long row_begin,
                                 long long nrows, int *counts) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < nrows)
    counts[i] = cfs_gen_row(&spec, row_begin + i, nullptr, nullptr);
}

template <typename T>
__global__ void gen_fill_kernel(cfs_gen_spec spec, long long row_begin,
                                long long nrows, const int *rowptr, int *colind,
                                T *values) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= nrows)
    return;
  const int begin = rowptr[i];
  const long long row = row_begin + i;
  if (spec.kind != CFS_GEN_BANDED) {
    // Laplacian rows hold at most 27 entries: generate, then narrow to T
    int32_t cols[27];
    double vals[27];
    const int n = cfs_gen_row(&spec, row, cols, vals);
    for (int k = 0; k < n; ++k) {
      colind[begin + k] = cols[k];
      values[begin + k] = (T)vals[k];
    }
    return;
  }
  // banded rows have no useful length bound: write entries in place, in the
  // same order cfs_gen_row emits them (lower ascending, diagonal, upper)
  int32_t *c = colind + begin;
  T *v = values + begin;
  int k = 0;
  double absum = 0.0;
  int32_t dmax = (int32_t)(row < spec.bw ? row : spec.bw);
  for (int32_t d = dmax; d >= 1; --d)
    if (cfs_banded_has(&spec, row, d)) {
      const double a = cfs_banded_val(&spec, row, d);
      c[k] = (int32_t)(row - d);
      v[k] = (T)a;
      absum += -a;
      ++k;
    }
  const int diag_at = k++;
  dmax = (int32_t)((spec.nrows - 1 - row) < spec.bw ? (spec.nrows - 1 - row)
                                                    : spec.bw);
  for (int32_t d = 1; d <= dmax; ++d)
    if (cfs_banded_has(&spec, row + d, d)) {
      const double a = cfs_banded_val(&spec, row + d, d);
      c[k] = (int32_t)(row + d);
      v[k] = (T)a;
      absum += -a;
      ++k;
    }
  c[diag_at] = (int32_t)row;
  v[diag_at] = (T)(1.0 + absum);
}

template <typename T>
__global__ void gen_x_kernel(unsigned long long seed, long long begin,
                             long long n, T *x) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n)
    x[i] = (T)cfs_gen_x(seed, begin + i);
}

inline unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }

} // namespace

extern "C" {

int cfs_gen_host_count(const cfs_gen_spec *spec, int64_t row_begin,
                       int64_t row_end, int32_t *rowptr) {
  if (!spec || !rowptr || row_end < row_begin)
    return CFS_ERR_INVALID;
  rowptr[0] = 0;
  for (int64_t i = row_begin; i < row_end; ++i)
    rowptr[i - row_begin + 1] =
        rowptr[i - row_begin] + cfs_gen_row(spec, i, nullptr, nullptr);
  return CFS_OK;
}

int cfs_gen_host_fill(const cfs_gen_spec *spec, int64_t row_begin,
                      int64_t row_end, const int32_t *rowptr, int32_t *colind,
                      void *values, int is_double) {
  if (!spec || !rowptr || !colind || !values)
    return CFS_ERR_INVALID;
  std::vector<double> tmp;
  for (int64_t i = row_begin; i < row_end; ++i) {
    const int begin = rowptr[i - row_begin];
    const int n = rowptr[i - row_begin + 1] - begin;
    if (is_double) {
      cfs_gen_row(spec, i, colind + begin, (double *)values + begin);
    } else {
      tmp.resize(n);
      cfs_gen_row(spec, i, colind + begin, tmp.data());
      for (int k = 0; k < n; ++k)
        ((float *)values)[begin + k] = (float)tmp[k];
    }
  }
  return CFS_OK;
}

int cfs_gen_host_x(uint64_t seed, int64_t begin, int64_t end, void *x,
                   int is_double) {
  if (!x)
    return CFS_ERR_INVALID;
  for (int64_t i = begin; i < end; ++i) {
    if (is_double)
      ((double *)x)[i - begin] = cfs_gen_x(seed, i);
    else
      ((float *)x)[i - begin] = (float)cfs_gen_x(seed, i);
  }
  return CFS_OK;
}

int cfs_cuda_gen_count(const cfs_gen_spec *spec, int64_t row_begin,
                       int64_t row_end, int32_t *rowptr_dev, int64_t *nnz) {
  if (!spec || !rowptr_dev || row_end < row_begin)
    return CFS_ERR_INVALID;
  const long long n = row_end - row_begin;
  cfsb::DevArray<int> counts;
  CFS_TRY(counts.alloc((size_t)n + 1));
  CFS_CUDA_TRY(cudaMemset(counts.p, 0, ((size_t)n + 1) * 4));
  if (n)
    gen_count_kernel<<<blocks_for(n), 256>>>(*spec, row_begin, n, counts.p);
  CFS_CUDA_TRY(cudaGetLastError());
  size_t tb = 0;
  CFS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.p, rowptr_dev,
                                             n + 1));
  cfsb::DevArray<char> tmp;
  CFS_TRY(tmp.alloc(tb));
  CFS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tb, counts.p, rowptr_dev,
                                             n + 1));
  int total = 0;
  CFS_CUDA_TRY(cudaMemcpy(&total, rowptr_dev + n, 4, cudaMemcpyDeviceToHost));
  if (nnz)
    *nnz = total;
  return CFS_OK;
}

int cfs_cuda_gen_fill(const cfs_gen_spec *spec, int64_t row_begin,
                      int64_t row_end, const int32_t *rowptr_dev,
                      int32_t *colind_dev, void *values_dev, int is_double) {
  if (!spec || !rowptr_dev || !colind_dev || !values_dev)
    return CFS_ERR_INVALID;
  const long long n = row_end - row_begin;
  if (n) {
    if (is_double)
      gen_fill_kernel<double><<<blocks_for(n), 256>>>(
          *spec, row_begin, n, rowptr_dev, colind_dev, (double *)values_dev);
    else
      gen_fill_kernel<float><<<blocks_for(n), 256>>>(
          *spec, row_begin, n, rowptr_dev, colind_dev, (float *)values_dev);
  }
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_CUDA_TRY(cudaDeviceSynchronize());
  return CFS_OK;
}

int cfs_cuda_gen_x(uint64_t seed, int64_t begin, int64_t end, void *x_dev,
                   int is_double) {
  if (!x_dev || end < begin)
    return CFS_ERR_INVALID;
  const long long n = end - begin;
  if (n) {
    if (is_double)
      gen_x_kernel<double><<<blocks_for(n), 256>>>(seed, begin, n,
                                                   (double *)x_dev);
    else
      gen_x_kernel<float><<<blocks_for(n), 256>>>(seed, begin, n,
                                                  (float *)x_dev);
  }
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_CUDA_TRY(cudaDeviceSynchronize());
  return CFS_OK;
}

} // extern "C"
