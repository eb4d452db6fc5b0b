// refmeta.cu -- the reference's conflict-free preprocessing, on the GPU,
// producing metadata BIT-EXACT with the reference for the same partition
// count P (= CFS_NUM_THREADS).
//
// Restated from (all in include/matrix/csr_matrix.tpp of the reference):
//   block-row conflict graph       :1364-1366, 1427-1477
//   color_greedy (first fit)       :2027-2077
//   2-colour load balancing        :2080-2213
//   per-colour row ranges          :1544-1627
//
// GPU formulation (not a translation of the OpenMP/TBB code):
//   * edges are emitted as 64-bit (u,v) keys, radix-sorted and uniqued, which
//     gives the set semantics of the reference's concurrent_unordered_set and
//     a CSR adjacency with ascending neighbours;
//   * sequential first-fit in natural vertex order == "colour = mex of the
//     colours of the lower-indexed neighbours". Every edge joins two different
//     partitions, so all lower-indexed neighbours of a vertex of partition t
//     live in partitions < t: P rounds, round t colours partition t in parallel;
//   * the balancing FIFO of the reference moves, in vertex order, every
//     vertex of the overloaded colour that has no neighbour of the target
//     colour while load - moved > mean: an ordered prefix sum. One CTA per
//     partition, block-wide scans;
//   * ranges are runs of equal colour over 16-row blocks: flag run heads,
//     compact, stable-sort by (partition, colour).
#include <cub/cub.cuh>

#include "common.cuh"

namespace cfsb {

namespace {

constexpr int kThreads = 256;
inline unsigned blocks_for(size_t n, int per = kThreads) {
  return (unsigned)((n + per - 1) / per);
}

struct PartInfo {
  int P;
  int per_split; // S of partition_by_nrows (csr_matrix.tpp:418)
  int nrows;
};

__device__ __forceinline__ int part_of(const PartInfo &pi, int row) {
  int t = row / pi.per_split;
  return t < pi.P ? t : pi.P - 1;
}
__device__ __forceinline__ int split_of(const PartInfo &pi, int t) {
  return t >= pi.P ? pi.nrows : t * pi.per_split;
}

__device__ __forceinline__ unsigned long long edge_key(int a, int b) {
  return ((unsigned long long)(unsigned)a << 32) | (unsigned)b;
}

// Pass 1 (emit == false): count the keys row i produces.
// Pass 2 (emit == true) : write them at offset[i].
template <bool emit>
__global__ void conflict_edges_kernel(PartInfo pi,
                                      const int *__restrict__ rowptr,
                                      const int *__restrict__ colind,
                                      const int *__restrict__ low_rowptr,
                                      const int *__restrict__ low_colind,
                                      unsigned long long *__restrict__ count,
                                      const unsigned long long *__restrict__ offset,
                                      unsigned long long *__restrict__ keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pi.nrows)
    return;
  const int t = part_of(pi, i);
  const int row_offset = split_of(pi, t);
  const int blk_row = i >> kBlkBits;
  unsigned long long n = 0;
  unsigned long long at = emit ? offset[i] : 0;
  // direct conflicts (:1441-1451)
  for (int j = low_rowptr[i]; j < low_rowptr[i + 1]; ++j) {
    const int col = low_colind[j];
    if (col < row_offset) {
      if (emit) {
        keys[at + n] = edge_key(blk_row, col >> kBlkBits);
        keys[at + n + 1] = edge_key(col >> kBlkBits, blk_row);
      }
      n += 2;
    }
  }
  // indirect conflicts (:1453-1475): the entries of row i from its first
  // upper-triangle entry on are the rows that write y[i] transposed
  const int end = rowptr[i + 1];
  int first_upper = end;
  for (int j = rowptr[i]; j < end; ++j)
    if (colind[j] > i) {
      first_upper = j;
      break;
    }
  for (int j = first_upper + 1; j < end; ++j) {
    const int cur = colind[j];
    const int pcur = part_of(pi, cur);
    for (int k = first_upper; k < j; ++k) {
      const int prev = colind[k];
      if (part_of(pi, prev) != pcur) {
        if (emit) {
          keys[at + n] = edge_key(prev >> kBlkBits, cur >> kBlkBits);
          keys[at + n + 1] = edge_key(cur >> kBlkBits, prev >> kBlkBits);
        }
        n += 2;
      }
    }
  }
  if (!emit)
    count[i] = n;
}

// vertex weights (:1434-1436): lower nnz of the block's 16 rows
__global__ void block_weight_kernel(int nblk, int nrows,
                                    const int *__restrict__ low_rowptr,
                                    int *__restrict__ weight) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblk)
    return;
  const int r0 = b << kBlkBits;
  const int r1 = min(r0 + kBlkFactor, nrows);
  weight[b] = low_rowptr[r1] - low_rowptr[r0];
}

__global__ void adjacency_kernel(int nblk, unsigned long long nkeys,
                                 const unsigned long long *__restrict__ keys,
                                 int *__restrict__ adj_ptr,
                                 int *__restrict__ adj) {
  unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (idx < nkeys)
    adj[idx] = (int)(unsigned)(keys[idx] & 0xffffffffull);
  if (idx <= (unsigned long long)nblk) {
    // adj_ptr[v] = first key whose high word is >= v
    const unsigned long long probe = idx << 32;
    unsigned long long lo = 0, hi = nkeys;
    while (lo < hi) {
      unsigned long long mid = (lo + hi) >> 1;
      if (keys[mid] < probe)
        lo = mid + 1;
      else
        hi = mid;
    }
    adj_ptr[idx] = (int)lo;
  }
}

// first-fit round for the vertices [b0, b1) of one partition
__global__ void first_fit_round_kernel(int b0, int b1,
                                       const int *__restrict__ adj_ptr,
                                       const int *__restrict__ adj,
                                       int *__restrict__ color,
                                       int *__restrict__ max_color) {
  int v = b0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= b1)
    return;
  const int begin = adj_ptr[v], end = adj_ptr[v + 1];
  int c = 0;
  // mex of the colours of lower-indexed (already coloured) neighbours
  for (int base = 0;; base += 64) {
    unsigned long long used = 0;
    for (int k = begin; k < end; ++k) {
      const int u = adj[k];
      if (u < v) {
        const int cu = color[u] - base;
        if (cu >= 0 && cu < 64)
          used |= 1ull << cu;
      }
    }
    if (~used) {
      c = base + __ffsll((long long)~used) - 1;
      break;
    }
  }
  color[v] = c;
  atomicMax(max_color, c);
}

// 2-colour balancing, one CTA per partition (:2094-2200)
__global__ void __launch_bounds__(kThreads)
    balance_kernel(PartInfo pi, int ncolors, const int *__restrict__ weight,
                   const int *__restrict__ adj_ptr,
                   const int *__restrict__ adj, int *color) {
  typedef cub::BlockScan<int, kThreads> Scan;
  typedef cub::BlockReduce<int, kThreads> Reduce;
  __shared__ union {
    typename Scan::TempStorage scan;
    typename Reduce::TempStorage reduce;
  } tmp;
  __shared__ int sh_load[2];
  __shared__ int sh_moved;
  const int t = blockIdx.x;
  const int off = split_of(pi, t);
  const int nrows_t = split_of(pi, t + 1) - off;
  const int nb = (nrows_t + kBlkFactor - 1) / kBlkFactor;
  const int b0 = off >> kBlkBits;
  // loads of colours 0 and 1
  int l0 = 0, l1 = 0;
  for (int i = threadIdx.x; i < nb; i += kThreads) {
    const int c = color[b0 + i];
    const int w = weight[b0 + i];
    l0 += c == 0 ? w : 0;
    l1 += c == 1 ? w : 0;
  }
  l0 = Reduce(tmp.reduce).Sum(l0);
  __syncthreads();
  l1 = Reduce(tmp.reduce).Sum(l1);
  if (threadIdx.x == 0) {
    sh_load[0] = l0;
    sh_load[1] = l1;
  }
  __syncthreads();
  const int mean = (sh_load[0] + sh_load[1]) / 2;
  for (int step = 0; step < ncolors - 1; ++step) {
    const int load0 = sh_load[0], load1 = sh_load[1];
    const int max_c = (load1 - mean) > (load0 - mean) ? 1 : 0;
    const int target = 1 - max_c;
    const int load_max = max_c ? load1 : load0;
    __syncthreads();
    if (threadIdx.x == 0)
      sh_moved = 0;
    __syncthreads();
    if (load_max - mean > 0) {
      int carry = 0;
      for (int base = 0; base < nb; base += kThreads) {
        const int i = base + threadIdx.x;
        int w = 0;
        bool eligible = false;
        int v = b0 + i;
        if (i < nb && color[v] == max_c) {
          eligible = true;
          for (int k = adj_ptr[v]; k < adj_ptr[v + 1]; ++k)
            if (color[adj[k]] == target) {
              eligible = false;
              break;
            }
          if (eligible)
            w = weight[v];
        }
        int excl, total;
        Scan(tmp.scan).ExclusiveSum(w, excl, total);
        const bool move = eligible && (load_max - (carry + excl) - mean > 0);
        __syncthreads();
        if (move) {
          color[v] = target;
          if (w)
            atomicAdd(&sh_moved, w);
        }
        carry += total;
        if (load_max - carry - mean <= 0)
          break; // uniform: every later vertex fails the loop condition
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      sh_load[max_c] -= sh_moved;
      sh_load[target] += sh_moved;
    }
    __syncthreads();
  }
}

// run heads over 16-row blocks
__global__ void run_head_kernel(PartInfo pi, int nblk, int ncolors,
                                const int *__restrict__ color,
                                unsigned char *__restrict__ is_head,
                                int *__restrict__ head_key) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblk)
    return;
  const int row = b << kBlkBits;
  const int t = part_of(pi, row);
  const bool head = (row == split_of(pi, t)) || (color[b - 1] != color[b]);
  is_head[b] = head;
  head_key[b] = t * ncolors + color[b];
}

__global__ void run_extent_kernel(PartInfo pi, int nruns, int nblk, int ncolors,
                                  const int *__restrict__ head_blk,
                                  const int *__restrict__ head_key_in,
                                  int *__restrict__ key_out,
                                  int *__restrict__ start_row,
                                  int *__restrict__ end_row,
                                  int *__restrict__ counts) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nruns)
    return;
  const int b = head_blk[r];
  const int key = head_key_in[b];
  const int t = key / ncolors;
  const int off = split_of(pi, t);
  const int part_end = split_of(pi, t + 1);
  const int last_blk = (r + 1 < nruns ? head_blk[r + 1] : nblk) - 1;
  key_out[r] = key;
  start_row[r] = (b << kBlkBits) - off;
  end_row[r] = min((last_blk << kBlkBits) + kBlkFactor - 1, part_end - 1) - off;
  atomicAdd(&counts[key], 1);
}

__global__ void range_ptr_kernel(int P, int ncolors,
                                 const int *__restrict__ counts,
                                 int *__restrict__ range_ptr,
                                 int *__restrict__ part_nranges) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P)
    return;
  int acc = 0;
  int *rp = range_ptr + (size_t)t * (ncolors + 1);
  rp[0] = 0;
  for (int c = 0; c < ncolors; ++c) {
    acc += counts[t * ncolors + c];
    rp[c + 1] = acc;
  }
  part_nranges[t] = acc;
}

__global__ void gather_kernel(int n, const int *__restrict__ perm,
                              const int *__restrict__ a_in,
                              const int *__restrict__ b_in,
                              int *__restrict__ a_out,
                              int *__restrict__ b_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  a_out[i] = a_in[perm[i]];
  b_out[i] = b_in[perm[i]];
}

__global__ void iota_kernel(int n, int *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = i;
}

} // namespace

int build_refmeta(cfs_matrix_s *m, cudaStream_t s) {
  const int N = m->nrows, P = m->nparts;
  PartInfo pi;
  pi.P = P;
  pi.nrows = N;
  pi.per_split = ((N / P - 1) | (kBlkFactor - 1)) + 1;
  const int V = (N + kBlkFactor - 1) / kBlkFactor;
  m->nblk = V;
  m->refmeta = false;

  // ---- vertex weights
  CFS_TRY(m->weight.alloc(V));
  block_weight_kernel<<<blocks_for(V), kThreads, 0, s>>>(V, N, m->low_rowptr.p,
                                                         m->weight.p);
  CFS_CUDA_TRY(cudaGetLastError());

  // ---- conflict edges: count, scan, emit, sort, unique
  DevArray<unsigned long long> cnt, off;
  CFS_TRY(cnt.alloc((size_t)N + 1));
  CFS_TRY(off.alloc((size_t)N + 1));
  CFS_CUDA_TRY(cudaMemsetAsync(cnt.p, 0, ((size_t)N + 1) * 8, s));
  conflict_edges_kernel<false><<<blocks_for(N), kThreads, 0, s>>>(
      pi, m->csr_rowptr, m->csr_colind, m->low_rowptr.p, m->low_colind.p,
      cnt.p, nullptr, nullptr);
  CFS_CUDA_TRY(cudaGetLastError());
  {
    size_t tb = 0;
    CFS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, off.p,
                                               (long long)N + 1, s));
    DevArray<char> tmp;
    CFS_TRY(tmp.alloc(tb));
    CFS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.p, off.p,
                                               (long long)N + 1, s));
    CFS_CUDA_TRY(cudaStreamSynchronize(s));
  }
  unsigned long long nkeys_raw = 0;
  CFS_CUDA_TRY(cudaMemcpy(&nkeys_raw, off.p + N, 8, cudaMemcpyDeviceToHost));
  if (nkeys_raw > (1ull << 31)) {
    set_error("conflict graph needs %llu edge keys: the reference's "
              "preprocessing is infeasible for this input",
              nkeys_raw);
    return CFS_ERR_TOO_LARGE;
  }
  DevArray<unsigned long long> keys, keys_sorted;
  DevArray<long long> nunique_dev;
  CFS_TRY(keys.alloc(nkeys_raw));
  CFS_TRY(keys_sorted.alloc(nkeys_raw));
  CFS_TRY(nunique_dev.alloc(1));
  unsigned long long nkeys = 0;
  if (nkeys_raw) {
    conflict_edges_kernel<true><<<blocks_for(N), kThreads, 0, s>>>(
        pi, m->csr_rowptr, m->csr_colind, m->low_rowptr.p, m->low_colind.p,
        nullptr, off.p, keys.p);
    CFS_CUDA_TRY(cudaGetLastError());
    size_t tb = 0;
    CFS_CUDA_TRY(cub::DeviceRadixSort::SortKeys(
        nullptr, tb, keys.p, keys_sorted.p, (long long)nkeys_raw, 0, 64, s));
    DevArray<char> tmp;
    CFS_TRY(tmp.alloc(tb));
    CFS_CUDA_TRY(cub::DeviceRadixSort::SortKeys(
        tmp.p, tb, keys.p, keys_sorted.p, (long long)nkeys_raw, 0, 64, s));
    size_t tb2 = 0;
    CFS_CUDA_TRY(cub::DeviceSelect::Unique(nullptr, tb2, keys_sorted.p, keys.p,
                                           nunique_dev.p, (long long)nkeys_raw,
                                           s));
    DevArray<char> tmp2;
    CFS_TRY(tmp2.alloc(tb2));
    CFS_CUDA_TRY(cub::DeviceSelect::Unique(tmp2.p, tb2, keys_sorted.p, keys.p,
                                           nunique_dev.p, (long long)nkeys_raw,
                                           s));
    CFS_CUDA_TRY(cudaStreamSynchronize(s));
    long long nu = 0;
    CFS_CUDA_TRY(cudaMemcpy(&nu, nunique_dev.p, 8, cudaMemcpyDeviceToHost));
    nkeys = (unsigned long long)nu;
  }
  m->nedges = (int64_t)nkeys;
  CFS_TRY(m->adj_ptr.alloc((size_t)V + 1));
  CFS_TRY(m->adj.alloc(nkeys));
  {
    const unsigned long long work = nkeys > (unsigned long long)V + 1
                                        ? nkeys
                                        : (unsigned long long)V + 1;
    adjacency_kernel<<<blocks_for(work), kThreads, 0, s>>>(V, nkeys, keys.p,
                                                           m->adj_ptr.p, m->adj.p);
    CFS_CUDA_TRY(cudaGetLastError());
  }

  // ---- first-fit colouring, partition by partition
  CFS_TRY(m->color.alloc(V));
  CFS_TRY(m->color_first.alloc(V));
  DevArray<int> max_color;
  CFS_TRY(max_color.alloc(1));
  CFS_CUDA_TRY(cudaMemsetAsync(max_color.p, 0xff, 4, s)); // -1
  for (int t = 0; t < P; ++t) {
    const int r0 = t * pi.per_split;
    const int r1 = t + 1 >= P ? N : (t + 1) * pi.per_split;
    const int b0 = r0 >> kBlkBits;
    const int b1 = (r1 + kBlkFactor - 1) >> kBlkBits;
    if (b1 > b0)
      first_fit_round_kernel<<<blocks_for(b1 - b0), kThreads, 0, s>>>(
          b0, b1, m->adj_ptr.p, m->adj.p, m->color.p, max_color.p);
  }
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_CUDA_TRY(cudaMemcpyAsync(m->color_first.p, m->color.p, (size_t)V * 4,
                               cudaMemcpyDeviceToDevice, s));
  int hmax = -1;
  CFS_CUDA_TRY(cudaMemcpyAsync(&hmax, max_color.p, 4, cudaMemcpyDeviceToHost,
                               s));
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  const int nc = hmax + 1;
  m->ncolors = nc;

  // ---- balancing
  if (nc > 1) {
    balance_kernel<<<P, kThreads, 0, s>>>(pi, nc, m->weight.p, m->adj_ptr.p,
                                          m->adj.p, m->color.p);
    CFS_CUDA_TRY(cudaGetLastError());
  }

  // ---- ranges
  DevArray<unsigned char> is_head;
  DevArray<int> head_key, blk_ids, head_blk, nruns_dev;
  CFS_TRY(is_head.alloc(V));
  CFS_TRY(head_key.alloc(V));
  CFS_TRY(blk_ids.alloc(V));
  CFS_TRY(head_blk.alloc(V));
  CFS_TRY(nruns_dev.alloc(1));
  run_head_kernel<<<blocks_for(V), kThreads, 0, s>>>(pi, V, nc, m->color.p,
                                                     is_head.p, head_key.p);
  iota_kernel<<<blocks_for(V), kThreads, 0, s>>>(V, blk_ids.p);
  CFS_CUDA_TRY(cudaGetLastError());
  {
    size_t tb = 0;
    CFS_CUDA_TRY(cub::DeviceSelect::Flagged(nullptr, tb, blk_ids.p, is_head.p,
                                            head_blk.p, nruns_dev.p, V, s));
    DevArray<char> tmp;
    CFS_TRY(tmp.alloc(tb));
    CFS_CUDA_TRY(cub::DeviceSelect::Flagged(tmp.p, tb, blk_ids.p, is_head.p,
                                            head_blk.p, nruns_dev.p, V, s));
    CFS_CUDA_TRY(cudaStreamSynchronize(s));
  }
  int nruns = 0;
  CFS_CUDA_TRY(cudaMemcpy(&nruns, nruns_dev.p, 4, cudaMemcpyDeviceToHost));
  m->nranges = nruns;
  DevArray<int> run_key, run_key_sorted, run_start, run_end, perm, perm_sorted,
      counts;
  CFS_TRY(run_key.alloc(nruns));
  CFS_TRY(run_key_sorted.alloc(nruns));
  CFS_TRY(run_start.alloc(nruns));
  CFS_TRY(run_end.alloc(nruns));
  CFS_TRY(perm.alloc(nruns));
  CFS_TRY(perm_sorted.alloc(nruns));
  CFS_TRY(counts.alloc((size_t)P * nc));
  CFS_CUDA_TRY(cudaMemsetAsync(counts.p, 0, (size_t)P * nc * 4, s));
  CFS_TRY(m->range_ptr.alloc((size_t)P * (nc + 1)));
  CFS_TRY(m->part_nranges.alloc(P));
  CFS_TRY(m->range_start.alloc(nruns));
  CFS_TRY(m->range_end.alloc(nruns));
  if (nruns) {
    run_extent_kernel<<<blocks_for(nruns), kThreads, 0, s>>>(
        pi, nruns, V, nc, head_blk.p, head_key.p, run_key.p, run_start.p,
        run_end.p, counts.p);
    iota_kernel<<<blocks_for(nruns), kThreads, 0, s>>>(nruns, perm.p);
    CFS_CUDA_TRY(cudaGetLastError());
    size_t tb = 0;
    CFS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(
        nullptr, tb, run_key.p, run_key_sorted.p, perm.p, perm_sorted.p, nruns,
        0, 32, s));
    DevArray<char> tmp;
    CFS_TRY(tmp.alloc(tb));
    CFS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(
        tmp.p, tb, run_key.p, run_key_sorted.p, perm.p, perm_sorted.p, nruns,
        0, 32, s));
    gather_kernel<<<blocks_for(nruns), kThreads, 0, s>>>(
        nruns, perm_sorted.p, run_start.p, run_end.p, m->range_start.p,
        m->range_end.p);
    CFS_CUDA_TRY(cudaGetLastError());
    CFS_CUDA_TRY(cudaStreamSynchronize(s));
  }
  range_ptr_kernel<<<blocks_for(P), kThreads, 0, s>>>(P, nc, counts.p,
                                                      m->range_ptr.p,
                                                      m->part_nranges.p);
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  m->refmeta = true;
  return CFS_OK;
}

} // namespace cfsb
