// hyb.cu -- Format::hyb: the reference's split of a symmetric matrix into a
// low-bandwidth part that is stored and multiplied symmetrically and a
// high-bandwidth remainder that is not (split_by_bandwidth,
// include/matrix/csr_matrix.tpp:314-401; kernels :3031-3162).
//
// In the reference the split exists but cannot run: with the default partitioner
// tune() reaches assert(false) for P > 1 (:283-296) and switches HYB off for
// P = 1 (:110, :143). The idea is sound, though: an entry far from the diagonal
// is the one whose transposed write y[col] += a*x[row] lands in somebody else's
// rows (a conflict on the CPU, a scattered L2 reduction here), so such entries
// are kept in BOTH triangles and only ever gathered. Here, done to the end:
//
//   near part  |col - row| < threshold (HybBwThreshold = 10000,
//              csr_matrix.hpp:92): goes through the whole symmetric pipeline --
//              lower extraction, layout, colouring metadata, CFS kernels;
//   far part   everything else, both triangles: its own sliced layout, the
//              Format::csr kernel (csr_path.cu) adds it onto y after the
//              symmetric kernel.
//
// The split itself is three passes over the full CSR on the GPU (count, scan,
// fill), order preserving, so both parts keep ascending columns.
#include <cub/cub.cuh>

#include "common.cuh"

namespace cfsb {
namespace {

constexpr int kThreads = 256;
inline unsigned blocks_for(size_t n) {
  return (unsigned)((n + kThreads - 1) / kThreads);
}

__global__ void hyb_count_kernel(int nrows, int row_begin, int threshold,
                                 const int *__restrict__ rowptr,
                                 const int *__restrict__ colind,
                                 int *__restrict__ near_cnt,
                                 int *__restrict__ far_cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows)
    return;
  const int row = row_begin + i;
  int nn = 0, nf = 0;
  for (int j = rowptr[i]; j < rowptr[i + 1]; ++j) {
    const int d = colind[j] - row;
    if ((d < 0 ? -d : d) < threshold)
      ++nn;
    else
      ++nf;
  }
  near_cnt[i] = nn;
  far_cnt[i] = nf;
}

template <typename T>
__global__ void hyb_fill_kernel(int nrows, int row_begin, int threshold,
                                const int *__restrict__ rowptr,
                                const int *__restrict__ colind,
                                const T *__restrict__ values,
                                const int *__restrict__ near_ptr,
                                const int *__restrict__ far_ptr,
                                int *__restrict__ near_col,
                                T *__restrict__ near_val,
                                int *__restrict__ far_col,
                                T *__restrict__ far_val) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows)
    return;
  const int row = row_begin + i;
  int on = near_ptr[i], of = far_ptr[i];
  for (int j = rowptr[i]; j < rowptr[i + 1]; ++j) {
    const int c = colind[j];
    const int d = c - row;
    if ((d < 0 ? -d : d) < threshold) {
      near_col[on] = c;
      near_val[on] = values[j];
      ++on;
    } else {
      far_col[of] = c;
      far_val[of] = values[j];
      ++of;
    }
  }
}

int scan(const int *in, int *out, size_t n, cudaStream_t s) {
  size_t tb = 0;
  CFS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out,
                                             (long long)n, s));
  DevArray<char> tmp;
  CFS_TRY(tmp.alloc(tb));
  CFS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, out, (long long)n,
                                             s));
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  return CFS_OK;
}

} // namespace

// Splits m's full CSR (m->csr_*) in place: afterwards m->csr_* is the near part
// (owned by m) and m->far a tuned non-symmetric matrix with the far part, or
// nullptr when nothing is far.
int split_hybrid(cfs_matrix_s *m, cudaStream_t s) {
  const int n = m->nrows;
  const size_t vs = m->vsize();
  DevArray<int> near_cnt, far_cnt;
  DevArray<int32_t> near_ptr, far_ptr;
  CFS_TRY(near_cnt.alloc((size_t)n + 1));
  CFS_TRY(far_cnt.alloc((size_t)n + 1));
  CFS_TRY(near_ptr.alloc((size_t)n + 1));
  CFS_TRY(far_ptr.alloc((size_t)n + 1));
  CFS_CUDA_TRY(cudaMemsetAsync(near_cnt.p, 0, ((size_t)n + 1) * 4, s));
  CFS_CUDA_TRY(cudaMemsetAsync(far_cnt.p, 0, ((size_t)n + 1) * 4, s));
  if (n > 0)
    hyb_count_kernel<<<blocks_for(n), kThreads, 0, s>>>(
        n, m->row_begin, m->hyb_threshold, m->csr_rowptr, m->csr_colind,
        near_cnt.p, far_cnt.p);
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_TRY(scan(near_cnt.p, near_ptr.p, (size_t)n + 1, s));
  CFS_TRY(scan(far_cnt.p, far_ptr.p, (size_t)n + 1, s));
  int nnear = 0, nfar = 0;
  CFS_CUDA_TRY(cudaMemcpy(&nnear, near_ptr.p + n, 4, cudaMemcpyDeviceToHost));
  CFS_CUDA_TRY(cudaMemcpy(&nfar, far_ptr.p + n, 4, cudaMemcpyDeviceToHost));
  m->hyb_far_entries = nfar;
  if (nfar == 0)
    return CFS_OK; // everything is near: plain SSS
  DevArray<int32_t> near_col, far_col;
  DevArray<char> near_val, far_val;
  CFS_TRY(near_col.alloc((size_t)nnear));
  CFS_TRY(near_val.alloc((size_t)nnear * vs));
  CFS_TRY(far_col.alloc((size_t)nfar));
  CFS_TRY(far_val.alloc((size_t)nfar * vs));
  if (m->is_double)
    hyb_fill_kernel<double><<<blocks_for(n), kThreads, 0, s>>>(
        n, m->row_begin, m->hyb_threshold, m->csr_rowptr, m->csr_colind,
        (const double *)m->csr_values, near_ptr.p, far_ptr.p, near_col.p,
        (double *)near_val.p, far_col.p, (double *)far_val.p);
  else
    hyb_fill_kernel<float><<<blocks_for(n), kThreads, 0, s>>>(
        n, m->row_begin, m->hyb_threshold, m->csr_rowptr, m->csr_colind,
        (const float *)m->csr_values, near_ptr.p, far_ptr.p, near_col.p,
        (float *)near_val.p, far_col.p, (float *)far_val.p);
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  // the near part replaces the full CSR of m (which m may only have borrowed)
  m->own_rowptr.release();
  m->own_colind.release();
  m->own_values.release();
  std::swap(m->own_rowptr.p, near_ptr.p);
  std::swap(m->own_rowptr.n, near_ptr.n);
  std::swap(m->own_colind.p, near_col.p);
  std::swap(m->own_colind.n, near_col.n);
  std::swap(m->own_values.p, near_val.p);
  std::swap(m->own_values.n, near_val.n);
  m->csr_rowptr = m->own_rowptr.p;
  m->csr_colind = m->own_colind.p;
  m->csr_values = m->own_values.p;
  // the far part: a non-symmetric matrix of its own
  cfs_matrix_s *f = new cfs_matrix_s;
  f->device = m->device;
  f->is_double = m->is_double;
  f->symmetric = false;
  f->nrows = m->nrows;
  f->ncols = m->ncols;
  f->global_nrows = m->global_nrows;
  f->nnz_full = nfar;
  std::swap(f->own_rowptr.p, far_ptr.p);
  std::swap(f->own_rowptr.n, far_ptr.n);
  std::swap(f->own_colind.p, far_col.p);
  std::swap(f->own_colind.n, far_col.n);
  std::swap(f->own_values.p, far_val.p);
  std::swap(f->own_values.n, far_val.n);
  f->csr_rowptr = f->own_rowptr.p;
  f->csr_colind = f->own_colind.p;
  f->csr_values = f->own_values.p;
  const int status = tune_csr(f, 1, CFS_TUNING_NONE, s);
  if (status != CFS_OK) {
    delete f;
    return status;
  }
  f->tuned = true;
  // the sliced layout is what runs; the far CSR itself has done its work
  f->own_colind.release();
  f->own_values.release();
  f->csr_colind = nullptr;
  f->csr_values = nullptr;
  m->far = f;
  return CFS_OK;
}

} // namespace cfsb
