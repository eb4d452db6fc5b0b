// preproc.cu -- GPU format conversion.
//
//   build_lower : full CSR -> lower-triangle CSR + dense diagonal. This is the
//                 reference's symmetry compression, serial()
//                 (include/matrix/csr_matrix.tpp:642-706) and the extraction
//                 loop of conflict_free_aposteriori() (:1257-1292): entries in
//                 CSR order, col < row kept with GLOBAL column id, col == row
//                 goes to diagonal_, col > row dropped. Bit-exact by
//                 construction (pure copies).
//   build_layout: lower CSR -> the execution layout of the sm_100a kernel:
//                 rows cut into virtual rows of <= kMaxChunk entries, 32
//                 virtual rows per slice (one per warp lane), entries stored
//                 slice-column-major so that every warp load of values /
//                 indices is one fully coalesced 256 B / 128 B request.
#include <cub/cub.cuh>

#include "common.cuh"

namespace cfsb {

namespace {

constexpr int kThreads = 256;
inline unsigned blocks_for(size_t n, int per = kThreads) {
  return (unsigned)((n + per - 1) / per);
}

// ---- lower extraction ----------------------------------------------------
__global__ void count_lower_kernel(int nrows, int row_begin,
                                   const int *__restrict__ rowptr,
                                   const int *__restrict__ colind,
                                   int *__restrict__ low_count,
                                   int *__restrict__ min_col,
                                   unsigned long long *__restrict__ ndiag,
                                   int *__restrict__ max_len) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int my_min = INT_MAX, my_diag = 0, my_len = 0;
  if (i < nrows) {
    my_len = rowptr[i + 1] - rowptr[i];
    const int grow = row_begin + i;
    int cnt = 0;
    for (int j = rowptr[i]; j < rowptr[i + 1]; ++j) {
      const int c = colind[j];
      cnt += (c < grow);
      my_diag += (c == grow);
      my_min = min(my_min, c);
    }
    low_count[i] = cnt;
  }
  // block reduce of min column / diagonal count
  typedef cub::BlockReduce<int, kThreads> Reduce;
  __shared__ typename Reduce::TempStorage tmp;
  int bmin = Reduce(tmp).Reduce(my_min, cub::Min());
  __syncthreads();
  int bdiag = Reduce(tmp).Sum(my_diag);
  __syncthreads();
  int blen = Reduce(tmp).Reduce(my_len, cub::Max());
  if (threadIdx.x == 0) {
    atomicMin(min_col, bmin);
    atomicAdd(ndiag, (unsigned long long)bdiag);
    atomicMax(max_len, blen);
  }
}

template <typename T>
__global__ void fill_lower_kernel(int nrows, int row_begin,
                                  const int *__restrict__ rowptr,
                                  const int *__restrict__ colind,
                                  const T *__restrict__ values,
                                  const int *__restrict__ low_rowptr,
                                  int *__restrict__ low_colind,
                                  T *__restrict__ low_values,
                                  T *__restrict__ diagonal) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows)
    return;
  const int grow = row_begin + i;
  int out = low_rowptr[i];
  T d = 0; // the reference zero-fills diagonal_
  for (int j = rowptr[i]; j < rowptr[i + 1]; ++j) {
    const int c = colind[j];
    if (c < grow) {
      low_colind[out] = c;
      low_values[out] = values[j];
      ++out;
    } else if (c == grow) {
      d = values[j]; // last one wins, like the reference's overwrite
    }
  }
  diagonal[i] = d;
}

// ---- virtual rows ----------------------------------------------------------
__global__ void count_vrows_kernel(int nrows, int chunk,
                                   const int *__restrict__ low_rowptr,
                                   int *__restrict__ nv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows)
    return;
  const int len = low_rowptr[i + 1] - low_rowptr[i];
  nv[i] = len == 0 ? 1 : (len + chunk - 1) / chunk;
}

__global__ void fill_vrows_kernel(int nrows, int row_begin, int max_chunk,
                                  const int *__restrict__ low_rowptr,
                                  const int *__restrict__ voff,
                                  int *__restrict__ vrow_row,
                                  int *__restrict__ vrow_start,
                                  int *__restrict__ vrow_len) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows)
    return;
  const int begin = low_rowptr[i];
  const int len = low_rowptr[i + 1] - begin;
  int v = voff[i];
  int done = 0, chunk = 0;
  do {
    const int take = min(max_chunk, len - done);
    vrow_row[v] = (row_begin + i) | (chunk ? kVrowCont : 0);
    vrow_start[v] = begin + done;
    vrow_len[v] = take;
    done += take;
    ++v;
    ++chunk;
  } while (done < len);
}

// largest distance between a row and its first stored column (the CSR the
// layout is cut from keeps the columns of a row ascending)
__global__ void bandwidth_kernel(int nrows, int row_begin,
                                 const int *__restrict__ rowptr,
                                 const int *__restrict__ colind,
                                 int *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int d = 0;
  if (i < nrows && rowptr[i + 1] > rowptr[i])
    d = max(0, row_begin + i - colind[rowptr[i]]);
  typedef cub::BlockReduce<int, kThreads> Reduce;
  __shared__ typename Reduce::TempStorage tmp;
  d = Reduce(tmp).Reduce(d, cub::Max());
  if (threadIdx.x == 0)
    atomicMax(out, d);
}

__global__ void slice_width_kernel(long long nvrows, long long nslices,
                                   const int *__restrict__ vrow_len,
                                   int *__restrict__ width) {
  long long s = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= nslices)
    return;
  long long v = s * kSliceRows + lane;
  int len = v < nvrows ? vrow_len[v] : 0;
  for (int o = 16; o; o >>= 1)
    len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if (lane == 0)
    width[s] = len;
}

template <typename T>
__global__ void fill_sell_kernel(long long nvrows, long long nslices,
                                 const int *__restrict__ slice_ptr,
                                 const int *__restrict__ vrow_start,
                                 const int *__restrict__ vrow_len,
                                 const int *__restrict__ low_colind,
                                 const T *__restrict__ low_values,
                                 int *__restrict__ vrow_row,
                                 int *__restrict__ sell_col,
                                 T *__restrict__ sell_val) {
  long long s = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= nslices)
    return;
  const long long v = s * kSliceRows + lane;
  int start = 0, len = 0;
  if (v < nvrows) {
    start = vrow_start[v];
    len = vrow_len[v];
  } else {
    vrow_row[v] = -1; // inactive lane of the last slice
  }
  const int w = slice_ptr[s + 1] - slice_ptr[s];
  size_t base = (size_t)slice_ptr[s] * kSliceRows + lane;
  for (int k = 0; k < w; ++k) {
    int c = -1;
    T a = 0;
    if (k < len) {
      c = low_colind[start + k];
      a = low_values[start + k];
    }
    sell_col[base + (size_t)k * kSliceRows] = c;
    sell_val[base + (size_t)k * kSliceRows] = a;
  }
}

// ---- length sorting of virtual rows ------------------------------------------
// Ragged matrices waste bandwidth on padding (a slice is as wide as its longest
// row). Sorting the virtual rows by decreasing length inside windows of
// `window` rows (SELL-C-sigma) packs rows of similar length together; the window
// keeps x / y accesses local for banded matrices.
__global__ void sort_key_kernel(long long nvrows, int window,
                                const int *__restrict__ vlen,
                                unsigned *__restrict__ key,
                                int *__restrict__ idx) {
  long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= nvrows)
    return;
  key[v] = ((unsigned)(v / window) << 6) | (unsigned)(kMaxChunk - vlen[v]);
  idx[v] = (int)v;
}

__global__ void permute_vrows_kernel(long long nvrows,
                                     const int *__restrict__ perm,
                                     const int *__restrict__ row_in,
                                     const int *__restrict__ start_in,
                                     const int *__restrict__ len_in,
                                     int *__restrict__ row_out,
                                     int *__restrict__ start_out,
                                     int *__restrict__ len_out) {
  long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= nvrows)
    return;
  const int p = perm[v];
  row_out[v] = row_in[p];
  start_out[v] = start_in[p];
  len_out[v] = len_in[p];
}

// ---- tiles of the persistent kernel ---------------------------------------
// Slices are taken in groups of kTileSlices; a group whose slices hold more
// than kTileSteps slice-steps is cut greedily (a slice never exceeds
// kMaxChunk <= kTileSteps steps). emit == false counts the tiles of a group.
template <bool emit>
__global__ void build_tiles_kernel(long long nslices, long long ngroups,
                                   const int *__restrict__ slice_ptr,
                                   int *__restrict__ count,
                                   const int *__restrict__ offset,
                                   TileRec *__restrict__ tile_rec) {
  long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= ngroups)
    return;
  const long long s0 = g * kTileSlices;
  const long long s1 = min(s0 + kTileSlices, nslices);
  int n = 0;
  int at = emit ? offset[g] : 0;
  long long first = s0;
  int steps = 0;
  auto close_tile = [&](long long end) {
    if (emit) {
      TileRec r;
      memset(&r, 0, sizeof(r));
      r.slice_begin = (int)first;
      r.nslices = (int)(end - first);
      r.step_begin = slice_ptr[first];
      r.nsteps = steps;
      for (int w = 0; w <= 8; ++w) {
        const long long sl = first + w < end ? first + w : end;
        r.slice_step[w] = slice_ptr[sl] - r.step_begin;
      }
      tile_rec[at + n] = r;
    }
    ++n;
  };
  for (long long s = s0; s < s1; ++s) {
    const int w = slice_ptr[s + 1] - slice_ptr[s];
    if (steps + w > kTileSteps && s > first) {
      close_tile(s);
      first = s;
      steps = 0;
    }
    steps += w;
  }
  close_tile(s1);
  if (!emit)
    count[g] = n;
}

int exclusive_scan_i32(const int *in, int *out, size_t n, cudaStream_t s) {
  size_t tmp_bytes = 0;
  CFS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out,
                                             (long long)n, s));
  DevArray<char> tmp;
  CFS_TRY(tmp.alloc(tmp_bytes));
  CFS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out,
                                             (long long)n, s));
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  return CFS_OK;
}

} // namespace

int build_lower(cfs_matrix_s *m, cudaStream_t s) {
  const int n = m->nrows;
  DevArray<int> counts; // n+1 so that the scan yields rowptr[n]
  CFS_TRY(counts.alloc((size_t)n + 1));
  CFS_CUDA_TRY(cudaMemsetAsync(counts.p, 0, ((size_t)n + 1) * 4, s));
  DevArray<int> min_col, max_len;
  DevArray<unsigned long long> ndiag;
  CFS_TRY(min_col.alloc(1));
  CFS_TRY(max_len.alloc(1));
  CFS_TRY(ndiag.alloc(1));
  CFS_CUDA_TRY(cudaMemsetAsync(max_len.p, 0, 4, s));
  const int init_min = m->row_begin;
  CFS_CUDA_TRY(cudaMemcpyAsync(min_col.p, &init_min, 4, cudaMemcpyHostToDevice,
                               s));
  CFS_CUDA_TRY(cudaMemsetAsync(ndiag.p, 0, 8, s));
  if (n > 0)
    count_lower_kernel<<<blocks_for(n), kThreads, 0, s>>>(
        n, m->row_begin, m->csr_rowptr, m->csr_colind, counts.p, min_col.p,
        ndiag.p, max_len.p);
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_TRY(m->low_rowptr.alloc((size_t)n + 1));
  CFS_TRY(exclusive_scan_i32(counts.p, m->low_rowptr.p, (size_t)n + 1, s));
  int nlow = 0, hmin = 0;
  unsigned long long hdiag = 0;
  CFS_CUDA_TRY(cudaMemcpy(&nlow, m->low_rowptr.p + n, 4,
                          cudaMemcpyDeviceToHost));
  CFS_CUDA_TRY(cudaMemcpy(&hmin, min_col.p, 4, cudaMemcpyDeviceToHost));
  CFS_CUDA_TRY(cudaMemcpy(&hdiag, ndiag.p, 8, cudaMemcpyDeviceToHost));
  CFS_CUDA_TRY(cudaMemcpy(&m->max_row_nnz_full, max_len.p, 4,
                          cudaMemcpyDeviceToHost));
  m->nnz_low = nlow;
  m->nnz_diag = (int64_t)hdiag;
  // multiple of 32 columns: keeps every 32-column window block 16-byte
  // aligned inside the extended vectors (variant 3 bulk copies)
  m->halo_begin = m->sharded ? ((hmin < 0 ? 0 : hmin) & ~31) : 0;
  CFS_TRY(m->low_colind.alloc((size_t)nlow));
  CFS_TRY(m->low_values.alloc((size_t)nlow * m->vsize()));
  CFS_TRY(m->diagonal.alloc((size_t)n * m->vsize()));
  if (n > 0) {
    if (m->is_double)
      fill_lower_kernel<double><<<blocks_for(n), kThreads, 0, s>>>(
          n, m->row_begin, m->csr_rowptr, m->csr_colind,
          (const double *)m->csr_values, m->low_rowptr.p, m->low_colind.p,
          (double *)m->low_values.p, (double *)m->diagonal.p);
    else
      fill_lower_kernel<float><<<blocks_for(n), kThreads, 0, s>>>(
          n, m->row_begin, m->csr_rowptr, m->csr_colind,
          (const float *)m->csr_values, m->low_rowptr.p, m->low_colind.p,
          (float *)m->low_values.p, (float *)m->diagonal.p);
  }
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  return CFS_OK;
}

int build_layout(cfs_matrix_s *m, cudaStream_t s) {
  return build_layout_from(m, m->low_rowptr.p, m->low_colind.p,
                           m->low_values.p, m->nnz_low, s, kMaxChunk, true);
}

// rowptr / colind / values: the CSR the layout is cut from -- the lower
// triangle for a symmetric matrix, the full CSR for Format::csr (csr_path.cu)
// chunk: longest virtual row. rechunk: a ragged matrix (its natural order pads
// too much) may restart with shorter chunks, about 1.3 x the mean row length:
// after length-sorting the slices of a row block are then nearly equally long,
// which is what the tile kernel (one warp per slice, a barrier per tile) needs.
int build_layout_from(cfs_matrix_s *m, const int32_t *src_rowptr,
                      const int32_t *src_colind, const void *src_values,
                      int64_t src_nnz, cudaStream_t s, int chunk,
                      bool rechunk) {
  const int n = m->nrows;
  // virtual rows
  DevArray<int> nv, voff;
  CFS_TRY(nv.alloc((size_t)n + 1));
  CFS_TRY(voff.alloc((size_t)n + 1));
  CFS_CUDA_TRY(cudaMemsetAsync(nv.p, 0, ((size_t)n + 1) * 4, s));
  if (n > 0)
    count_vrows_kernel<<<blocks_for(n), kThreads, 0, s>>>(n, chunk,
                                                          src_rowptr, nv.p);
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_TRY(exclusive_scan_i32(nv.p, voff.p, (size_t)n + 1, s));
  int nvrows = 0;
  CFS_CUDA_TRY(cudaMemcpy(&nvrows, voff.p + n, 4, cudaMemcpyDeviceToHost));
  m->nvrows = nvrows;
  m->nslices = (nvrows + kSliceRows - 1) / kSliceRows;
  const size_t nlanes = (size_t)m->nslices * kSliceRows;
  DevArray<int> vstart, vlen;
  CFS_TRY(m->vrow_row.alloc(nlanes));
  CFS_TRY(vstart.alloc(nlanes));
  CFS_TRY(vlen.alloc(nlanes));
  if (n > 0)
    fill_vrows_kernel<<<blocks_for(n), kThreads, 0, s>>>(
        n, m->row_begin, chunk, src_rowptr, voff.p, m->vrow_row.p, vstart.p,
        vlen.p);
  CFS_CUDA_TRY(cudaGetLastError());
  // slice widths -> slice_ptr; if the natural order pads too much, sort the
  // virtual rows by length (windowed first, globally for heavy skew) and redo
  DevArray<int> width;
  CFS_TRY(width.alloc((size_t)m->nslices + 1));
  CFS_TRY(m->slice_ptr.alloc((size_t)m->nslices + 1));
  int total_width = 0;
  m->sort_window = 0;
  for (int attempt = 0; attempt < 3; ++attempt) {
    CFS_CUDA_TRY(
        cudaMemsetAsync(width.p, 0, ((size_t)m->nslices + 1) * 4, s));
    if (m->nslices > 0)
      slice_width_kernel<<<blocks_for(nlanes), kThreads, 0, s>>>(
          m->nvrows, m->nslices, vlen.p, width.p);
    CFS_CUDA_TRY(cudaGetLastError());
    CFS_TRY(exclusive_scan_i32(width.p, m->slice_ptr.p,
                               (size_t)m->nslices + 1, s));
    CFS_CUDA_TRY(cudaMemcpy(&total_width, m->slice_ptr.p + m->nslices, 4,
                            cudaMemcpyDeviceToHost));
    const double padding = src_nnz > 0 ? (double)total_width * kSliceRows /
                                                (double)src_nnz
                                          : 1.0;
    const double limit = attempt == 0 ? 1.15 : 1.5;
    if (padding <= limit || attempt == 2 || g_options.sort_rows == 0)
      break;
    if (attempt == 0 && rechunk && chunk == kMaxChunk && n > 0 &&
        g_options.rechunk && g_options.tile6) {
      // only where the tile kernel will run: the column window of a row block
      // (its rows + the bandwidth) has to fit (tiles6.cu); a power-law matrix
      // keeps the long chunks (R-MAT scale 24: 1.01 ms vs 1.33 ms re-chunked)
      DevArray<int> bw;
      CFS_TRY(bw.alloc(1));
      CFS_CUDA_TRY(cudaMemsetAsync(bw.p, 0, 4, s));
      bandwidth_kernel<<<blocks_for(n), kThreads, 0, s>>>(
          n, m->row_begin, src_rowptr, src_colind, bw.p);
      CFS_CUDA_TRY(cudaGetLastError());
      int bandwidth = 0;
      CFS_CUDA_TRY(cudaMemcpy(&bandwidth, bw.p, 4, cudaMemcpyDeviceToHost));
      const double mean = (double)src_nnz / (double)n;
      int shorter = (int)(g_options.rechunk_pct * 0.01 * mean + 0.999);
      shorter = shorter < 8 ? 8 : shorter;
      if (shorter < kMaxChunk &&
          bandwidth + 2 * kT6Slices * kSliceRows < kT6MaxCols)
        return build_layout_from(m, src_rowptr, src_colind, src_values,
                                 src_nnz, s, shorter, false);
    }
    // first try: sort inside the row blocks that become the tiles of variant 6
    const long long window =
        attempt == 0 ? kT6Slices * kSliceRows : (long long)1 << 30;
    if (nvrows / window >= (1 << 26))
      break;
    DevArray<unsigned> key, key_out;
    DevArray<int> idx, perm, row2, start2, len2;
    CFS_TRY(key.alloc(nvrows));
    CFS_TRY(key_out.alloc(nvrows));
    CFS_TRY(idx.alloc(nvrows));
    CFS_TRY(perm.alloc(nvrows));
    CFS_TRY(row2.alloc(nlanes));
    CFS_TRY(start2.alloc(nlanes));
    CFS_TRY(len2.alloc(nlanes));
    sort_key_kernel<<<blocks_for(nvrows), kThreads, 0, s>>>(
        nvrows, (int)window, vlen.p, key.p, idx.p);
    size_t tb = 0;
    CFS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(
        nullptr, tb, key.p, key_out.p, idx.p, perm.p, nvrows, 0, 32, s));
    DevArray<char> tmp;
    CFS_TRY(tmp.alloc(tb));
    CFS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(
        tmp.p, tb, key.p, key_out.p, idx.p, perm.p, nvrows, 0, 32, s));
    permute_vrows_kernel<<<blocks_for(nvrows), kThreads, 0, s>>>(
        nvrows, perm.p, m->vrow_row.p, vstart.p, vlen.p, row2.p, start2.p,
        len2.p);
    CFS_CUDA_TRY(cudaGetLastError());
    CFS_CUDA_TRY(cudaMemcpyAsync(m->vrow_row.p, row2.p, (size_t)nvrows * 4,
                                 cudaMemcpyDeviceToDevice, s));
    CFS_CUDA_TRY(cudaMemcpyAsync(vstart.p, start2.p, (size_t)nvrows * 4,
                                 cudaMemcpyDeviceToDevice, s));
    CFS_CUDA_TRY(cudaMemcpyAsync(vlen.p, len2.p, (size_t)nvrows * 4,
                                 cudaMemcpyDeviceToDevice, s));
    CFS_CUDA_TRY(cudaStreamSynchronize(s));
    m->sort_window = window;
  }
  m->padded_entries = (int64_t)total_width * kSliceRows;
  // widest slice: the stride of a warp's shared-memory piece in variant 7
  m->max_slice_steps = 0;
  if (m->nslices > 0) {
    DevArray<int> wmax;
    CFS_TRY(wmax.alloc(1));
    size_t tb = 0;
    CFS_CUDA_TRY(cub::DeviceReduce::Max(nullptr, tb, width.p, wmax.p,
                                        (int)m->nslices, s));
    DevArray<char> tmp;
    CFS_TRY(tmp.alloc(tb));
    CFS_CUDA_TRY(cub::DeviceReduce::Max(tmp.p, tb, width.p, wmax.p,
                                        (int)m->nslices, s));
    CFS_CUDA_TRY(cudaMemcpyAsync(&m->max_slice_steps, wmax.p, 4,
                                 cudaMemcpyDeviceToHost, s));
    CFS_CUDA_TRY(cudaStreamSynchronize(s));
  }
  CFS_TRY(m->sell_col.alloc((size_t)m->padded_entries));
  CFS_TRY(m->sell_val.alloc((size_t)m->padded_entries * m->vsize()));
  if (m->nslices > 0) {
    if (m->is_double)
      fill_sell_kernel<double><<<blocks_for(nlanes), kThreads, 0, s>>>(
          m->nvrows, m->nslices, m->slice_ptr.p, vstart.p, vlen.p,
          src_colind, (const double *)src_values, m->vrow_row.p,
          m->sell_col.p, (double *)m->sell_val.p);
    else
      fill_sell_kernel<float><<<blocks_for(nlanes), kThreads, 0, s>>>(
          m->nvrows, m->nslices, m->slice_ptr.p, vstart.p, vlen.p,
          src_colind, (const float *)src_values, m->vrow_row.p,
          m->sell_col.p, (float *)m->sell_val.p);
  }
  CFS_CUDA_TRY(cudaGetLastError());
  // tiles of the persistent kernel
  const long long ngroups = (m->nslices + kTileSlices - 1) / kTileSlices;
  m->ntiles = 0;
  if (ngroups > 0) {
    DevArray<int> tcount, toff;
    CFS_TRY(tcount.alloc((size_t)ngroups + 1));
    CFS_TRY(toff.alloc((size_t)ngroups + 1));
    CFS_CUDA_TRY(cudaMemsetAsync(tcount.p, 0, ((size_t)ngroups + 1) * 4, s));
    build_tiles_kernel<false><<<blocks_for(ngroups), kThreads, 0, s>>>(
        m->nslices, ngroups, m->slice_ptr.p, tcount.p, nullptr, nullptr);
    CFS_CUDA_TRY(cudaGetLastError());
    CFS_TRY(exclusive_scan_i32(tcount.p, toff.p, (size_t)ngroups + 1, s));
    int ntiles = 0;
    CFS_CUDA_TRY(cudaMemcpy(&ntiles, toff.p + ngroups, 4,
                            cudaMemcpyDeviceToHost));
    m->ntiles = ntiles;
    CFS_TRY(m->tile_rec.alloc((size_t)ntiles));
    build_tiles_kernel<true><<<blocks_for(ngroups), kThreads, 0, s>>>(
        m->nslices, ngroups, m->slice_ptr.p, nullptr, toff.p, m->tile_rec.p);
    CFS_CUDA_TRY(cudaGetLastError());
  }
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  return CFS_OK;
}

namespace {
// per slice: smallest column it touches (INT_MAX: none), smallest and largest
// row it holds (INT_MAX / -1: no live lane)
__global__ void slice_reach_kernel(long long nslices,
                                   const int *__restrict__ slice_ptr,
                                   const int *__restrict__ vrow_row,
                                   const int *__restrict__ sell_col,
                                   int *__restrict__ min_col,
                                   int *__restrict__ min_row,
                                   int *__restrict__ max_row) {
  const long long s = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= nslices)
    return;
  const int p0 = slice_ptr[s], w = slice_ptr[s + 1] - p0;
  const int tag = vrow_row[s * kSliceRows + lane];
  int lo = INT_MAX;
  int rlo = tag >= 0 ? (tag & kVrowRowMask) : INT_MAX;
  int rhi = tag >= 0 ? (tag & kVrowRowMask) : -1;
  const int *cp = sell_col + (size_t)p0 * kSliceRows + lane;
  for (int k = 0; k < w; ++k) {
    const int c = cp[(size_t)k * kSliceRows];
    if (c >= 0)
      lo = min(lo, c);
  }
  for (int o = 16; o; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    rlo = min(rlo, __shfl_xor_sync(0xffffffffu, rlo, o));
    rhi = max(rhi, __shfl_xor_sync(0xffffffffu, rhi, o));
  }
  if (lane == 0) {
    min_col[s] = lo;
    min_row[s] = rlo;
    max_row[s] = rhi;
  }
}
} // namespace

namespace {
__global__ void halo_extent_kernel(long long nslices, int row_begin,
                                   const int *__restrict__ min_col,
                                   int *__restrict__ out) {
  const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s < nslices && min_col[s] < row_begin)
    atomicMax(out, (int)(s + 1));
}
} // namespace

// A shard: one past the last slice that references a column below row_begin.
// Only slices [0, halo_slice_end) need the kernel instantiations that test
// every column against row_begin (x from / reductions into the GPU below); in
// a banded or stencil shard that is the first 0.1 - 2 % of the slices.
int build_halo_extent(cfs_matrix_s *m, cudaStream_t s) {
  m->halo_slice_end = m->nslices;
  if (!m->sharded || m->nslices == 0 || !m->sell_col.p)
    return CFS_OK;
  const long long ns = m->nslices;
  DevArray<int> d_min, d_rlo, d_rhi, d_out;
  CFS_TRY(d_min.alloc((size_t)ns));
  CFS_TRY(d_rlo.alloc((size_t)ns));
  CFS_TRY(d_rhi.alloc((size_t)ns));
  CFS_TRY(d_out.alloc(1));
  CFS_CUDA_TRY(cudaMemsetAsync(d_out.p, 0, 4, s));
  slice_reach_kernel<<<blocks_for((size_t)ns * 32), kThreads, 0, s>>>(
      ns, m->slice_ptr.p, m->vrow_row.p, m->sell_col.p, d_min.p, d_rlo.p,
      d_rhi.p);
  halo_extent_kernel<<<blocks_for((size_t)ns), kThreads, 0, s>>>(
      ns, m->row_begin, d_min.p, d_out.p);
  CFS_CUDA_TRY(cudaGetLastError());
  int end = 0;
  CFS_CUDA_TRY(cudaMemcpyAsync(&end, d_out.p, 4, cudaMemcpyDeviceToHost, s));
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  // whole tiles of variant 6 / whole CTAs of the others
  m->halo_slice_end = std::min<long long>(
      ns, ((long long)end + kT6Slices - 1) / kT6Slices * kT6Slices);
  return CFS_OK;
}

// Stages of the host-vector pipeline of cfs_cuda_spmv (see cfs_matrix_s::Stage),
// for an unsharded matrix whose slices follow the row order at least block-wise:
// natural order, or rows length-sorted inside windows (stage boundaries then
// fall on window boundaries).
//
// The slices are cut into pipeline_chunks equal chunks. The y rows of chunk c-1
// are final once every slice that reaches below the first row of chunk c has
// run -- in a banded / stencil matrix those are the FIRST few slices of chunk c
// (its "head": 40 k of 1.3 M rows for the 27-point Laplacian). So chunk c runs
// as two stages, head and rest: after the head (which needs only the head's
// rows of x) the D2H of chunk c-1 starts, while the H2D of the rest of chunk c
// is still under way. Without the split y lags x by a whole chunk in each
// direction (tools/e2e_probe.py).
// Tune-time half: what the plan needs from the column stream (which may be
// released after tune) -- per slice its smallest column and its row span, 12
// bytes per slice, kept on the device. The plan itself is made by
// build_pipeline_plan on the first call with host vectors: callers whose vectors
// live in unified or device memory never pay for it (7 ms of tune on config 2).
int build_pipeline_reach(cfs_matrix_s *m, cudaStream_t s) {
  m->stages.clear();
  m->plan_state = 2; // nothing to plan
  m->reach_min.release();
  m->reach_rlo.release();
  m->reach_rhi.release();
  if (m->sharded || m->far || m->nslices < 4096)
    return CFS_OK;
  // stage boundaries must not cut a sort window
  const long long unit =
      m->sort_window == 0 ? 1 : m->sort_window / kSliceRows;
  if (m->sort_window % kSliceRows != 0 || unit * 64 > m->nslices)
    return CFS_OK; // globally sorted (power-law matrices): no row order left
  const long long ns = m->nslices;
  CFS_TRY(m->reach_min.alloc((size_t)ns));
  CFS_TRY(m->reach_rlo.alloc((size_t)ns));
  CFS_TRY(m->reach_rhi.alloc((size_t)ns));
  slice_reach_kernel<<<blocks_for((size_t)ns * 32), kThreads, 0, s>>>(
      ns, m->slice_ptr.p, m->vrow_row.p, m->sell_col.p, m->reach_min.p,
      m->reach_rlo.p, m->reach_rhi.p);
  CFS_CUDA_TRY(cudaGetLastError());
  m->plan_state = 1; // reach known, plan pending
  return CFS_OK;
}

int build_pipeline_plan(cfs_matrix_s *m, cudaStream_t s) {
  m->stages.clear();
  if (m->plan_state != 1)
    return CFS_OK;
  m->plan_state = 2;
  const long long unit =
      m->sort_window == 0 ? 1 : m->sort_window / kSliceRows;
  const long long ns = m->nslices;
  std::vector<int> min_col((size_t)ns), min_row((size_t)ns),
      max_row((size_t)ns);
  CFS_CUDA_TRY(cudaMemcpyAsync(min_col.data(), m->reach_min.p, (size_t)ns * 4,
                               cudaMemcpyDeviceToHost, s));
  CFS_CUDA_TRY(cudaMemcpyAsync(min_row.data(), m->reach_rlo.p, (size_t)ns * 4,
                               cudaMemcpyDeviceToHost, s));
  CFS_CUDA_TRY(cudaMemcpyAsync(max_row.data(), m->reach_rhi.p, (size_t)ns * 4,
                               cudaMemcpyDeviceToHost, s));
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  m->reach_min.release();
  m->reach_rlo.release();
  m->reach_rhi.release();
  int K = g_options.pipeline_chunks;
  if (g_options.pipeline_adaptive) {
    // every stage costs 10-25 us of hand-overs: keep >= 4 MB of x per chunk
    const long long by_size =
        (long long)m->ncols * (long long)m->vsize() / (4ll << 20);
    K = (int)(by_size < 2 ? 2 : by_size < K ? by_size : K);
  }
  auto snap = [&](long long v) { // to a multiple of `unit`
    return (v + unit / 2) / unit * unit;
  };
  // pipeline_taper: the first H2D and the last D2H have nothing to overlap
  // with, so the chunks at both ends are smaller (weights 1, 2, 3, 3, ..., 2, 1)
  std::vector<long long> weight((size_t)K, 1);
  if (g_options.pipeline_taper)
    for (int c = 0; c < K; ++c) {
      const int edge = c < K - 1 - c ? c : K - 1 - c;
      weight[c] = edge + 1 < 3 ? edge + 1 : 3;
    }
  long long total_weight = 0;
  for (long long w : weight)
    total_weight += w;
  std::vector<long long> cut((size_t)K + 1, 0);
  long long acc_weight = 0;
  for (int c = 0; c < K; ++c) {
    acc_weight += weight[c];
    cut[c + 1] = c + 1 == K ? ns : snap(ns * acc_weight / total_weight);
  }
  for (int c = 0; c < K; ++c)
    if (cut[c + 1] <= cut[c])
      return CFS_OK;
  auto first_row_of = [&](long long s0, long long s1) {
    int r = INT_MAX;
    for (long long i = s0; i < s1; ++i)
      r = min_row[i] < r ? min_row[i] : r;
    return r;
  };
  // stage boundaries: chunk starts, and inside chunk c (c >= 1) the end of its
  // head = one past the last slice that reaches below the chunk's first row
  std::vector<long long> bound;
  std::vector<int> chunk_first_stage((size_t)K + 1, 0);
  std::vector<int> chunk_row0((size_t)K + 1, m->nrows);
  for (int c = 0; c < K; ++c) {
    chunk_first_stage[c] = (int)bound.size();
    bound.push_back(cut[c]);
    chunk_row0[c] = first_row_of(cut[c], cut[c + 1]);
    if (chunk_row0[c] == INT_MAX)
      return CFS_OK;
    if (c == 0 || !g_options.pipeline_split)
      continue;
    long long head_end = cut[c];
    for (long long i = cut[c]; i < cut[c + 1]; ++i)
      if (min_col[i] < chunk_row0[c])
        head_end = i + 1;
    head_end = (head_end + unit - 1) / unit * unit;
    // worth a stage of its own only if it is a small part of the chunk
    if (head_end > cut[c] && (head_end - cut[c]) * 4 <= cut[c + 1] - cut[c])
      bound.push_back(head_end);
  }
  chunk_first_stage[K] = (int)bound.size();
  bound.push_back(ns);
  const int S = (int)bound.size() - 1;
  std::vector<int> stage_min((size_t)S, INT_MAX);
  m->stages.resize((size_t)S);
  int prev_hi = -1;
  for (int j = 0; j < S; ++j) {
    cfs_matrix_s::Stage &st = m->stages[j];
    st.slice0 = bound[j];
    st.slice1 = bound[j + 1];
    int lo = INT_MAX, hi = -1;
    for (long long i = bound[j]; i < bound[j + 1]; ++i) {
      lo = min_row[i] < lo ? min_row[i] : lo;
      hi = max_row[i] > hi ? max_row[i] : hi;
      stage_min[j] = min_col[i] < stage_min[j] ? min_col[i] : stage_min[j];
    }
    // the stages have to follow the row order (a row split over a boundary
    // may appear on both sides)
    if (hi < 0 || lo < prev_hi) {
      m->stages.clear();
      return CFS_OK;
    }
    prev_hi = hi;
    // x rows this stage adds: from where the previous stage stopped up to its
    // own last row
    st.x_row0 = j == 0 ? 0 : m->stages[j - 1].x_row1;
    st.x_row1 = j + 1 == S ? m->ncols : hi + 1;
    if (st.x_row1 < st.x_row0)
      st.x_row1 = st.x_row0;
  }
  // y rows of chunk d: [first row of chunk d, first row of chunk d+1); final
  // after the chunk's own last stage, after a row shared with the next chunk
  // has received all its parts, and after the last stage reaching into it
  for (int d = 0; d < K; ++d) {
    const int row0 = d == 0 ? 0 : chunk_row0[d];
    const int row1 = d + 1 < K ? chunk_row0[d + 1] : m->nrows;
    int after = chunk_first_stage[d + 1] - 1;
    for (int j = after + 1; j < S; ++j)
      if (stage_min[j] < row1)
        after = j;
    if (row1 > row0)
      m->stages[after].y_ready.push_back(std::make_pair(row0, row1));
  }
  return CFS_OK;
}

} // namespace cfsb
