// spmv.cu -- the sm_100a SpMV kernels.
//
// sym_spmv_sell_kernel replaces the reference's hot loop,
// cpu_mv_sym_conflict_free_v2 (include/matrix/csr_matrix.tpp:2966-3028) and
// cpu_mv_sym_serial (:2707-2729):  y = A*x with only the lower triangle stored;
// every stored (i, j, a) contributes  y[i] += a*x[j]  and  y[j] += a*x[i].
//
// Mapping to the GPU:
//   * one warp per slice of 32 virtual rows, one lane per virtual row; entry k
//     of all 32 rows is contiguous in memory, so each warp-wide load of values
//     (256 B for f64) and column ids (128 B) is a single coalesced request that
//     streams from HBM exactly once (L1 no-allocate).
//   * the direct term y[i] += a*x[j] accumulates in a register of the lane that
//     owns the row: no reduction tree at all for unsplit rows.
//   * the transposed term y[j] += a*x[i] is a no-return reduction (RED.ADD) to
//     L2. For stencil / banded matrices consecutive lanes hit consecutive
//     columns, so a warp-wide RED touches 2-3 cache lines.
//   * x gathers go through L1/L2 (x and y stay L2-resident: 126 MB L2).
//   * y is zeroed by a memset on the same stream before the kernel; the
//     diagonal term is folded into the direct-term accumulator.
#include "common.cuh"

namespace cfsb {

namespace {

// streaming loads: read once, do not pollute L1
__device__ __forceinline__ double ld_stream(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];"
               : "=d"(v)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];"
               : "=f"(v)
               : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];"
               : "=r"(v)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void red_add(double *p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void red_add(float *p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

constexpr int kSpmvThreads = 256;
constexpr int kUnroll = 4;

// x and y are "virtual base" pointers: indexed by GLOBAL row/column id (for a
// shard they point halo_begin elements before the extended local vector).
template <typename T>
__global__ void __launch_bounds__(kSpmvThreads)
    sym_spmv_sell_kernel(long long nslices, int row_begin,
                         const int *__restrict__ slice_ptr,
                         const int *__restrict__ vrow_row,
                         const int *__restrict__ sell_col,
                         const T *__restrict__ sell_val,
                         const T *__restrict__ diagonal,
                         const T *__restrict__ x, T *__restrict__ y) {
  const int lane = threadIdx.x & 31;
  const long long s =
      (blockIdx.x * (long long)kSpmvThreads + threadIdx.x) >> 5;
  if (s >= nslices)
    return;
  const int tag = vrow_row[s * kSliceRows + lane];
  const bool active = tag >= 0;
  const int row = tag & kVrowRowMask;
  T xr = 0, acc = 0;
  if (active) {
    xr = x[row];
    if (!(tag & kVrowCont))
      acc = diagonal[row - row_begin] * xr;
  }
  const int p0 = slice_ptr[s], p1 = slice_ptr[s + 1];
  const int *cp = sell_col + (size_t)p0 * kSliceRows + lane;
  const T *vp = sell_val + (size_t)p0 * kSliceRows + lane;
  int w = p1 - p0;
  for (; w >= kUnroll; w -= kUnroll) {
    int c[kUnroll];
    T a[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      c[u] = ld_stream(cp + u * kSliceRows);
      a[u] = ld_stream(vp + u * kSliceRows);
    }
    T xc[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      xc[u] = c[u] >= 0 ? x[c[u]] : T(0);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (c[u] >= 0) {
        acc += a[u] * xc[u];
        red_add(y + c[u], a[u] * xr);
      }
    }
    cp += kUnroll * kSliceRows;
    vp += kUnroll * kSliceRows;
  }
  for (; w > 0; --w) {
    const int c = ld_stream(cp);
    const T a = ld_stream(vp);
    if (c >= 0) {
      acc += a * x[c];
      red_add(y + c, a * xr);
    }
    cp += kSliceRows;
    vp += kSliceRows;
  }
  if (active)
    red_add(y + row, acc);
}

// Plain CSR y = A*x, one warp per row: the comparator the reference's test
// uses (cpu_mv / cpu_mv_serial, csr_matrix.tpp:2665-2704; test_spmv_mmf.cpp:
// 85-89). Not an optimisation target.
template <typename T>
__global__ void __launch_bounds__(kSpmvThreads)
    csr_spmv_kernel(int nrows, const int *__restrict__ rowptr,
                    const int *__restrict__ colind,
                    const T *__restrict__ values, const T *__restrict__ x,
                    T *__restrict__ y) {
  const int lane = threadIdx.x & 31;
  const long long row =
      (blockIdx.x * (long long)kSpmvThreads + threadIdx.x) >> 5;
  if (row >= nrows)
    return;
  T acc = 0;
  for (int j = rowptr[row] + lane; j < rowptr[row + 1]; j += 32)
    acc += values[j] * x[colind[j]];
  for (int o = 16; o; o >>= 1)
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0)
    y[row] = acc;
}

} // namespace

int launch_sym_spmv(const cfs_matrix_s *m, void *y_ext, const void *x_ext,
                    cudaStream_t s, cudaEvent_t ev0, cudaEvent_t ev1) {
  const size_t vs = m->vsize();
  const size_t ext_len = (size_t)(m->row_begin + m->nrows - m->halo_begin);
  CFS_CUDA_TRY(cudaMemsetAsync(y_ext, 0, ext_len * vs, s));
  if (m->nslices == 0)
    return CFS_OK;
  const unsigned grid = (unsigned)((m->nslices * 32 + kSpmvThreads - 1) /
                                   kSpmvThreads);
  if (ev0)
    CFS_CUDA_TRY(cudaEventRecord(ev0, s));
  if (m->is_double) {
    const double *xb = (const double *)x_ext - m->halo_begin;
    double *yb = (double *)y_ext - m->halo_begin;
    sym_spmv_sell_kernel<double><<<grid, kSpmvThreads, 0, s>>>(
        m->nslices, m->row_begin, m->slice_ptr.p, m->vrow_row.p, m->sell_col.p,
        (const double *)m->sell_val.p, (const double *)m->diagonal.p, xb, yb);
  } else {
    const float *xb = (const float *)x_ext - m->halo_begin;
    float *yb = (float *)y_ext - m->halo_begin;
    sym_spmv_sell_kernel<float><<<grid, kSpmvThreads, 0, s>>>(
        m->nslices, m->row_begin, m->slice_ptr.p, m->vrow_row.p, m->sell_col.p,
        (const float *)m->sell_val.p, (const float *)m->diagonal.p, xb, yb);
  }
  CFS_CUDA_TRY(cudaGetLastError());
  if (ev1)
    CFS_CUDA_TRY(cudaEventRecord(ev1, s));
  return CFS_OK;
}

int launch_csr_spmv(const cfs_matrix_s *m, void *y, const void *x,
                    cudaStream_t s) {
  if (m->nrows == 0)
    return CFS_OK;
  const unsigned grid =
      (unsigned)(((size_t)m->nrows * 32 + kSpmvThreads - 1) / kSpmvThreads);
  if (m->is_double)
    csr_spmv_kernel<double><<<grid, kSpmvThreads, 0, s>>>(
        m->nrows, m->csr_rowptr, m->csr_colind, (const double *)m->csr_values,
        (const double *)x, (double *)y);
  else
    csr_spmv_kernel<float><<<grid, kSpmvThreads, 0, s>>>(
        m->nrows, m->csr_rowptr, m->csr_colind, (const float *)m->csr_values,
        (const float *)x, (float *)y);
  CFS_CUDA_TRY(cudaGetLastError());
  return CFS_OK;
}

} // namespace cfsb
