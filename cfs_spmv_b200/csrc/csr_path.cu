// csr_path.cu -- the NON-symmetric path, Format::csr.
//
// Reference: tune() for symmetric_ == false (include/matrix/csr_matrix.tpp:
// 249-255): Tuning::Aggressive -> partition_by_nnz (:438-541), Tuning::None ->
// partition_by_nrows (:404-435); kernel cpu_mv / cpu_mv_serial (:2665-2704):
//     y[i] = sum_j values[j] * x[colind[j]]      one accumulator, CSR order.
//
// Here: row_split_ is computed on the GPU (bit-exact with the reference) and the
// full CSR is cut into the same sliced layout the symmetric kernels stream
// (build_layout_from): one lane per row, entries slice-column-major, so values
// and indices arrive in full cache lines instead of the 7-of-32 lanes a
// warp-per-row kernel keeps busy on a stencil matrix. The lane adds its
// products in CSR order with separate multiply and add -- the arithmetic the
// reference's `g++ -O2` build performs -- so y is BIT-IDENTICAL to cpu_mv for
// every row of at most kMaxChunk (32) entries; longer rows are cut into chunks
// whose partial sums meet in y through RED (normwise tolerance).
#include "common.cuh"
#include "spmv_tma.cuh"

namespace cfsb {
namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ double ld_stream(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// y_tmp += a * x without contraction into an FMA
__device__ __forceinline__ double mul_add(double acc, double a, double x) {
  return __dadd_rn(acc, __dmul_rn(a, x));
}
__device__ __forceinline__ float mul_add(float acc, float a, float x) {
  return __fadd_rn(acc, __fmul_rn(a, x));
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    csr_sell_kernel(long long nslices, const int *__restrict__ slice_ptr,
                    const int *__restrict__ vrow_row,
                    const int *__restrict__ sell_col,
                    const T *__restrict__ sell_val, const T *__restrict__ x,
                    T *__restrict__ y) {
  const int lane = threadIdx.x & 31;
  const long long s = (blockIdx.x * (long long)kThreads + threadIdx.x) >> 5;
  if (s >= nslices)
    return;
  const int tag = vrow_row[s * kSliceRows + lane];
  const int p0 = slice_ptr[s], p1 = slice_ptr[s + 1];
  const int *cp = sell_col + (size_t)p0 * kSliceRows + lane;
  const T *vp = sell_val + (size_t)p0 * kSliceRows + lane;
  T acc = 0;
  int w = p1 - p0;
  for (; w >= 4; w -= 4) {
    int c[4];
    T a[4], xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      c[u] = ld_stream(cp + u * kSliceRows);
      a[u] = ld_stream(vp + u * kSliceRows);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      xv[u] = c[u] >= 0 ? x[c[u]] : T(0);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (c[u] >= 0)
        acc = mul_add(acc, a[u], xv[u]);
    cp += 4 * kSliceRows;
    vp += 4 * kSliceRows;
  }
  for (; w > 0; --w) {
    const int c = ld_stream(cp);
    const T a = ld_stream(vp);
    if (c >= 0)
      acc = mul_add(acc, a, x[c]);
    cp += kSliceRows;
    vp += kSliceRows;
  }
  // y was cleared: 0 + acc = acc exactly for an unsplit row; the chunks of a
  // long row meet here
  if (tag >= 0)
    tma::red_add(y + (tag & kVrowRowMask), acc);
}

// partition_by_nnz, csr_matrix.tpp:438-541 (non-symmetric branch :504-515 and
// the tail :517-530), without the row-by-row scan: the running count only
// grows, so the first row where it reaches nnz_per_split is a binary search in
// rowptr and the split falls on the next row index i with (i+1) % 16 == 0.
__global__ void partition_by_nnz_kernel(int nrows, int P,
                                        const int *__restrict__ rowptr,
                                        int *__restrict__ row_split) {
  if (threadIdx.x != 0 || blockIdx.x != 0)
    return;
  const long long nnz_per_split = rowptr[nrows] / P;
  int split_cnt = 0, start = 0;
  long long curr = 0;
  row_split[0] = 0;
  while (start < nrows) {
    const long long target = (long long)rowptr[start] + nnz_per_split;
    int lo = start, hi = nrows; // first i in [start, nrows) with rowptr[i+1] >= target
    while (lo < hi) {
      const int mid = lo + (hi - lo) / 2;
      if (rowptr[mid + 1] >= target)
        hi = mid;
      else
        lo = mid + 1;
    }
    const long long split_row = (long long)(lo | (kBlkFactor - 1)) + 1;
    if (lo >= nrows || split_row > nrows) {
      curr = (long long)rowptr[nrows] - rowptr[start];
      break;
    }
    ++split_cnt;
    if (split_cnt <= P)
      row_split[split_cnt] = (int)split_row;
    start = (int)split_row;
    curr = 0;
  }
  if (curr < nnz_per_split && split_cnt <= P) {
    ++split_cnt; // the reference writes row_split_[P+1] here when split_cnt was P
    if (split_cnt <= P)
      row_split[split_cnt] = nrows;
  }
  if (split_cnt > P)
    row_split[P] = nrows;
  for (int i = split_cnt + 1; i <= P; ++i)
    row_split[i] = nrows;
}

} // namespace

int tune_csr(cfs_matrix_s *m, int nparts, int tuning, cudaStream_t s) {
  // ---- row_split_
  m->row_split.assign((size_t)nparts + 1, 0);
  m->row_split[nparts] = m->nrows;
  if (nparts > 1) {
    if (tuning == CFS_TUNING_AGGRESSIVE) {
      m->part_by_nnz = true; // counted by size(), csr_matrix.tpp:223-224
      DevArray<int> split;
      CFS_TRY(split.alloc((size_t)nparts + 1));
      partition_by_nnz_kernel<<<1, 32, 0, s>>>(m->nrows, nparts, m->csr_rowptr,
                                               split.p);
      CFS_CUDA_TRY(cudaGetLastError());
      CFS_CUDA_TRY(cudaMemcpyAsync(m->row_split.data(), split.p,
                                   ((size_t)nparts + 1) * 4,
                                   cudaMemcpyDeviceToHost, s));
      CFS_CUDA_TRY(cudaStreamSynchronize(s));
    } else {
      const long long S = ((m->nrows / nparts - 1) | (kBlkFactor - 1)) + 1;
      if (m->nrows / nparts < 1 || (long long)(nparts - 1) * S > m->nrows) {
        set_error("nparts=%d is not a valid partition count for %d rows",
                  nparts, m->nrows);
        return CFS_ERR_INVALID;
      }
      for (int t = 0; t < nparts; ++t)
        m->row_split[t] = (int32_t)(t * S);
    }
  }
  // ---- execution layout of the full CSR
  if (g_options.csr_layout && m->nnz_full > 0)
    CFS_TRY(build_layout_from(m, m->csr_rowptr, m->csr_colind, m->csr_values,
                              m->nnz_full, s));
  return CFS_OK;
}

// declared in common.cuh next to the warp-per-row comparator kernel (spmv.cu)
int launch_csr_sell(const cfs_matrix_s *m, void *y, const void *x,
                    cudaStream_t s, bool accumulate) {
  if (!accumulate)
    CFS_CUDA_TRY(cudaMemsetAsync(y, 0, (size_t)m->nrows * m->vsize(), s));
  if (m->nslices == 0)
    return CFS_OK;
  const unsigned grid =
      (unsigned)((m->nslices * 32 + kThreads - 1) / kThreads);
  if (m->is_double)
    csr_sell_kernel<double><<<grid, kThreads, 0, s>>>(
        m->nslices, m->slice_ptr.p, m->vrow_row.p, m->sell_col.p,
        (const double *)m->sell_val.p, (const double *)x, (double *)y);
  else
    csr_sell_kernel<float><<<grid, kThreads, 0, s>>>(
        m->nslices, m->slice_ptr.p, m->vrow_row.p, m->sell_col.p,
        (const float *)m->sell_val.p, (const float *)x, (float *)y);
  CFS_CUDA_TRY(cudaGetLastError());
  return CFS_OK;
}

} // namespace cfsb
