// hubs.cu -- hub columns of power-law matrices.
//
// In a skewed matrix a few columns collect a large share of the lower-triangle
// entries (R-MAT scale 24: 136 k of 16.8 M columns hold 71 % of them, column 0
// alone 238 k). For such a column the transposed term y[c] += a*x[row] is a long
// dot product; done as one RED per entry it hammers a single L2 address. The
// reference meets the same columns as "conflicts with everybody" (its colouring
// degenerates to one partition per colour on R-MAT) and proposes
// a split format (HYB, csr_matrix.tpp:314-401) for them. Here:
//   * entries of hub columns (>= kHubMinCount entries) are ALSO stored column by
//     column (hub_ptr / hub_row / hub_val);
//   * in the row stream they carry a flag: the row kernel still uses them for the
//     direct term y[row] += a*x[c] (x of a hub column is always cache-hot) but
//     skips the transposed RED;
//   * hub_spmv_kernel then does the transposed term column-wise: one warp per
//     chunk of a hub column, coalesced reads, a shuffle reduction and ONE RED.
#include <cub/cub.cuh>

#include "common.cuh"

namespace cfsb {

namespace {

constexpr int kThreads = 256;
inline unsigned blocks_for(size_t n) {
  return (unsigned)((n + kThreads - 1) / kThreads);
}

__global__ void col_count_kernel(long long nnz, const int *__restrict__ colind,
                                 int *__restrict__ count) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < nnz)
    atomicAdd(&count[colind[i]], 1);
}

__global__ void hub_flag_kernel(int ncols, const int *__restrict__ count,
                                int *__restrict__ flag) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncols)
    flag[c] = count[c] >= kHubMinCount ? 1 : 0;
}

// hub_id[c] = index among the hubs or -1; per hub: column, entry count, chunks
__global__ void hub_table_kernel(int ncols, const int *__restrict__ count,
                                 const int *__restrict__ flag,
                                 const int *__restrict__ flag_scan,
                                 int *__restrict__ hub_id,
                                 int *__restrict__ hub_col,
                                 int *__restrict__ hub_cnt,
                                 int *__restrict__ hub_chunks) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols)
    return;
  if (!flag[c]) {
    hub_id[c] = -1;
    return;
  }
  const int k = flag_scan[c];
  hub_id[c] = k;
  hub_col[k] = c;
  hub_cnt[k] = count[c];
  hub_chunks[k] = (count[c] + kHubChunk - 1) / kHubChunk;
}

// one thread per row: copy its hub entries into the column lists
template <typename T>
__global__ void hub_fill_kernel(int nrows, int row_begin,
                                const int *__restrict__ low_rowptr,
                                const int *__restrict__ low_colind,
                                const T *__restrict__ low_values,
                                const int *__restrict__ hub_id,
                                const int *__restrict__ hub_ptr,
                                int *__restrict__ cursor,
                                int *__restrict__ hub_row,
                                T *__restrict__ hub_val) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows)
    return;
  for (int j = low_rowptr[i]; j < low_rowptr[i + 1]; ++j) {
    const int k = hub_id[low_colind[j]];
    if (k >= 0) {
      const int at = hub_ptr[k] + atomicAdd(&cursor[k], 1);
      hub_row[at] = row_begin + i;
      hub_val[at] = low_values[j];
    }
  }
}

__global__ void hub_chunk_table_kernel(int nhubs,
                                       const int *__restrict__ hub_col,
                                       const int *__restrict__ hub_ptr,
                                       const int *__restrict__ chunk_ptr,
                                       int4 *__restrict__ chunks) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nhubs)
    return;
  const int begin = hub_ptr[k], end = hub_ptr[k + 1];
  int at = chunk_ptr[k];
  for (int b = begin; b < end; b += kHubChunk, ++at)
    chunks[at] = make_int4(hub_col[k], b, min(b + kHubChunk, end), 0);
}

// flag hub entries in a copy of the column stream
__global__ void hub_mark_kernel(long long n, const int *__restrict__ sell_col,
                                const int *__restrict__ hub_id,
                                int *__restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const int c = sell_col[i];
  out[i] = (c >= 0 && hub_id[c] >= 0) ? (c | kHubFlag) : c;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    hub_spmv_kernel(int nchunks, const int4 *__restrict__ chunks,
                    const int *__restrict__ hub_row,
                    const T *__restrict__ hub_val, const T *__restrict__ x,
                    T *__restrict__ y) {
  const int w = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nchunks)
    return;
  const int4 ch = chunks[w];
  T acc = 0;
  for (int p = ch.y + lane; p < ch.z; p += 32)
    acc += hub_val[p] * x[hub_row[p]];
  for (int o = 16; o; o >>= 1)
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0)
    atomicAdd(y + ch.x, acc);
}

template <typename T>
int scan_i32(const int *in, int *out, size_t n, cudaStream_t s) {
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

int build_hubs(cfs_matrix_s *m, cudaStream_t s) {
  m->nhubs = 0;
  m->hub_entries = 0;
  m->nhub_chunks = 0;
  // only for ragged, unsharded matrices: structured ones have no hubs and the
  // fused multi-GPU path does not know about them
  if (m->sharded || m->nnz_low == 0 || !g_options.hubs ||
      m->nregular * 8 >= m->nslices)
    return CFS_OK;
  const int ncols = m->ncols;
  DevArray<int> count, flag, flag_scan, hub_id;
  CFS_TRY(count.alloc((size_t)ncols + 1));
  CFS_TRY(flag.alloc((size_t)ncols + 1));
  CFS_TRY(flag_scan.alloc((size_t)ncols + 1));
  CFS_TRY(hub_id.alloc((size_t)ncols));
  CFS_CUDA_TRY(cudaMemsetAsync(count.p, 0, ((size_t)ncols + 1) * 4, s));
  CFS_CUDA_TRY(cudaMemsetAsync(flag.p, 0, ((size_t)ncols + 1) * 4, s));
  col_count_kernel<<<blocks_for((size_t)m->nnz_low), kThreads, 0, s>>>(
      m->nnz_low, m->low_colind.p, count.p);
  hub_flag_kernel<<<blocks_for(ncols), kThreads, 0, s>>>(ncols, count.p,
                                                         flag.p);
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_TRY(scan_i32<int>(flag.p, flag_scan.p, (size_t)ncols + 1, s));
  int nhubs = 0;
  CFS_CUDA_TRY(cudaMemcpy(&nhubs, flag_scan.p + ncols, 4,
                          cudaMemcpyDeviceToHost));
  if (nhubs == 0)
    return CFS_OK;
  DevArray<int> hub_col, hub_cnt, hub_chunks, chunk_ptr, cursor;
  CFS_TRY(hub_col.alloc(nhubs));
  CFS_TRY(hub_cnt.alloc((size_t)nhubs + 1));
  CFS_TRY(hub_chunks.alloc((size_t)nhubs + 1));
  CFS_TRY(chunk_ptr.alloc((size_t)nhubs + 1));
  CFS_TRY(cursor.alloc(nhubs));
  CFS_TRY(m->hub_ptr.alloc((size_t)nhubs + 1));
  CFS_CUDA_TRY(cudaMemsetAsync(hub_cnt.p, 0, ((size_t)nhubs + 1) * 4, s));
  CFS_CUDA_TRY(cudaMemsetAsync(hub_chunks.p, 0, ((size_t)nhubs + 1) * 4, s));
  CFS_CUDA_TRY(cudaMemsetAsync(cursor.p, 0, (size_t)nhubs * 4, s));
  hub_table_kernel<<<blocks_for(ncols), kThreads, 0, s>>>(
      ncols, count.p, flag.p, flag_scan.p, hub_id.p, hub_col.p, hub_cnt.p,
      hub_chunks.p);
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_TRY(scan_i32<int>(hub_cnt.p, m->hub_ptr.p, (size_t)nhubs + 1, s));
  CFS_TRY(scan_i32<int>(hub_chunks.p, chunk_ptr.p, (size_t)nhubs + 1, s));
  int nent = 0, nchunks = 0;
  CFS_CUDA_TRY(cudaMemcpy(&nent, m->hub_ptr.p + nhubs, 4,
                          cudaMemcpyDeviceToHost));
  CFS_CUDA_TRY(cudaMemcpy(&nchunks, chunk_ptr.p + nhubs, 4,
                          cudaMemcpyDeviceToHost));
  CFS_TRY(m->hub_row.alloc((size_t)nent));
  CFS_TRY(m->hub_val.alloc((size_t)nent * m->vsize()));
  CFS_TRY(m->hub_chunks.alloc((size_t)nchunks));
  CFS_TRY(m->hub_colstream.alloc((size_t)m->padded_entries));
  if (m->is_double)
    hub_fill_kernel<double><<<blocks_for(m->nrows), kThreads, 0, s>>>(
        m->nrows, m->row_begin, m->low_rowptr.p, m->low_colind.p,
        (const double *)m->low_values.p, hub_id.p, m->hub_ptr.p, cursor.p,
        m->hub_row.p, (double *)m->hub_val.p);
  else
    hub_fill_kernel<float><<<blocks_for(m->nrows), kThreads, 0, s>>>(
        m->nrows, m->row_begin, m->low_rowptr.p, m->low_colind.p,
        (const float *)m->low_values.p, hub_id.p, m->hub_ptr.p, cursor.p,
        m->hub_row.p, (float *)m->hub_val.p);
  hub_chunk_table_kernel<<<blocks_for(nhubs), kThreads, 0, s>>>(
      nhubs, hub_col.p, m->hub_ptr.p, chunk_ptr.p, m->hub_chunks.p);
  hub_mark_kernel<<<blocks_for((size_t)m->padded_entries), kThreads, 0, s>>>(
      m->padded_entries, m->sell_col.p, hub_id.p, m->hub_colstream.p);
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  m->nhubs = nhubs;
  m->hub_entries = nent;
  m->nhub_chunks = nchunks;
  return CFS_OK;
}

int launch_hub_spmv(const cfs_matrix_s *m, void *y, const void *x,
                    cudaStream_t s) {
  if (m->nhub_chunks == 0)
    return CFS_OK;
  const unsigned grid = blocks_for((size_t)m->nhub_chunks * 32);
  if (m->is_double)
    hub_spmv_kernel<double><<<grid, kThreads, 0, s>>>(
        (int)m->nhub_chunks, m->hub_chunks.p, m->hub_row.p,
        (const double *)m->hub_val.p, (const double *)x, (double *)y);
  else
    hub_spmv_kernel<float><<<grid, kThreads, 0, s>>>(
        (int)m->nhub_chunks, m->hub_chunks.p, m->hub_row.p,
        (const float *)m->hub_val.p, (const float *)x, (float *)y);
  CFS_CUDA_TRY(cudaGetLastError());
  return CFS_OK;
}

} // namespace cfsb
