// tiles6.cu -- preprocessing of variant 6 (spmv_tile.cuh): for every tile of
// kT6Slices slices, the column window it touches, the column-major slot of
// every entry and the per-column slot offsets.
//
// This is the GPU counterpart of the reference's conflict analysis
// (conflict_free_aposteriori, csr_matrix.tpp:1364-1477: which rows write which
// y entries) done per tile instead of per thread partition: after it, every
// transposed-term product has a private shared-memory slot and every column of
// a tile's window is summed by exactly one thread.
//
// Applies when every tile's window [min column, max column] spans at most
// kT6MaxCols columns and its products fit in kT6MaxSmemBytes of shared memory
// (banded / FEM-like orderings). Otherwise nt6 stays 0 and the other variants
// run.
#include <cub/cub.cuh>

#include <vector>

#include "common.cuh"

namespace cfsb {
namespace {

constexpr int kBoundsThreads = 256;
constexpr int kSlotThreads = 1024;
constexpr int kBinsPerThread = kT6MaxCols / kSlotThreads;

__device__ __forceinline__ void tile_range(long long tile, long long nslices,
                                           const int *slice_ptr, size_t *begin,
                                           size_t *end) {
  const long long s0 = tile * kT6Slices;
  const long long s1 = s0 + kT6Slices < nslices ? s0 + kT6Slices : nslices;
  *begin = (size_t)slice_ptr[s0] * kSliceRows;
  *end = (size_t)slice_ptr[s1] * kSliceRows;
}

// per tile: smallest / largest column and the number of real entries
__global__ void __launch_bounds__(kBoundsThreads)
    tile_bounds_kernel(long long nslices, const int *__restrict__ slice_ptr,
                       const int *__restrict__ sell_col, int *__restrict__ lo,
                       int *__restrict__ hi, int *__restrict__ count) {
  size_t begin, end;
  tile_range(blockIdx.x, nslices, slice_ptr, &begin, &end);
  int my_lo = INT_MAX, my_hi = -1, my_n = 0;
  for (size_t e = begin + threadIdx.x; e < end; e += kBoundsThreads) {
    const int c = sell_col[e];
    if (c >= 0) {
      my_lo = min(my_lo, c);
      my_hi = max(my_hi, c);
      ++my_n;
    }
  }
  typedef cub::BlockReduce<int, kBoundsThreads> Reduce;
  __shared__ typename Reduce::TempStorage tmp;
  my_lo = Reduce(tmp).Reduce(my_lo, cub::Min());
  __syncthreads();
  my_hi = Reduce(tmp).Reduce(my_hi, cub::Max());
  __syncthreads();
  my_n = Reduce(tmp).Sum(my_n);
  if (threadIdx.x == 0) {
    lo[blockIdx.x] = my_n ? my_lo : 0;
    hi[blockIdx.x] = my_hi;
    count[blockIdx.x] = my_n;
  }
}

// per tile: histogram over the window -> slot offsets (cptr) -> a slot for
// every entry. Which of its column's slots an entry gets is free, and it
// decides the shared-memory bank the SpMV kernel's product store hits: ncu had
// those stores at 6.3-way bank conflicts (10.6 M of 15.4 M wavefronts) when the
// slots were handed out in the order the atomics landed. Here warp w walks the
// steps of slice w exactly like the SpMV kernel will, and in every step its
// lanes pick, one after the other, a free slot of their column whose bank no
// earlier lane of the same store wavefront has taken (`banks` = 16 bank pairs
// per half warp for 8-byte products, 32 banks per warp for 4-byte ones).
// Columns with more than 32 entries in the tile keep the first-come order.
__global__ void __launch_bounds__(kSlotThreads)
    tile_slots_kernel(long long nslices, const int *__restrict__ slice_ptr,
                      const int *__restrict__ sell_col,
                      const int *__restrict__ lo, const int *__restrict__ ncols,
                      const long long *__restrict__ cptr_off,
                      unsigned short *__restrict__ cptr,
                      unsigned *__restrict__ pack, int banks) {
  extern __shared__ int window[];
  int *start = window;               // histogram, then exclusive offsets
  int *cursor = window + kT6MaxCols;
  typedef cub::BlockScan<int, kSlotThreads> Scan;
  __shared__ typename Scan::TempStorage scan_tmp;
  const long long tile = blockIdx.x;
  size_t begin, end;
  tile_range(tile, nslices, slice_ptr, &begin, &end);
  const int base = lo[tile], W = ncols[tile];
  for (int j = threadIdx.x; j < kT6MaxCols; j += kSlotThreads) {
    start[j] = 0;
    cursor[j] = 0;
  }
  __syncthreads();
  for (size_t e = begin + threadIdx.x; e < end; e += kSlotThreads) {
    const int c = sell_col[e];
    if (c >= 0)
      atomicAdd(&start[c - base], 1);
  }
  __syncthreads();
  // exclusive scan over the window: kBinsPerThread consecutive bins per thread
  int local[kBinsPerThread], sum = 0;
#pragma unroll
  for (int k = 0; k < kBinsPerThread; ++k) {
    local[k] = start[threadIdx.x * kBinsPerThread + k];
    sum += local[k];
  }
  int offset = 0;
  Scan(scan_tmp).ExclusiveSum(sum, offset);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kBinsPerThread; ++k) {
    start[threadIdx.x * kBinsPerThread + k] = offset;
    offset += local[k];
  }
  __syncthreads();
  unsigned short *out = cptr + cptr_off[tile];
  // W < kT6MaxCols and the bins beyond the window are empty: start[W] = total
  for (int j = threadIdx.x; j <= W; j += kSlotThreads)
    out[j] = (unsigned short)start[j];
  __syncthreads();
  // cursor[j]: bit mask of the used slots of column j (<= 32 entries), or the
  // number of slots handed out (longer columns)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long sl = tile * kT6Slices + warp;
  if (sl < nslices) {
    const int p0 = slice_ptr[sl], width = slice_ptr[sl + 1] - p0;
    const int group = banks == 16 ? lane >> 4 : 0; // store wavefront of the lane
    for (int k = 0; k < width; ++k) {
      const size_t e = ((size_t)p0 + k) * kSliceRows + lane;
      const int c = sell_col[e];
      const int j = c >= 0 ? c - base : 0;
      const int first = c >= 0 ? start[j] : 0;
      const int n = c >= 0 ? start[j + 1] - first : 0;
      int slot = -1;
      unsigned taken[2] = {0u, 0u}; // banks claimed per store wavefront
      for (int turn = 0; turn < 32; ++turn) {
        int bank = -1;
        if (lane == turn && c >= 0) {
          if (n > 32) {
            slot = first + atomicAdd(&cursor[j], 1);
          } else {
            const unsigned all = n == 32 ? 0xffffffffu : ((1u << n) - 1u);
            while (slot < 0) {
              const unsigned free_bits = ~(unsigned)cursor[j] & all;
              // first free slot on a bank nobody in this wavefront holds, else
              // the first free one
              int pick = -1;
              for (unsigned f = free_bits; f; f &= f - 1) {
                const int b = __ffs(f) - 1;
                if (!((taken[group] >> ((first + b) & (banks - 1))) & 1u)) {
                  pick = b;
                  break;
                }
              }
              if (pick < 0)
                pick = __ffs(free_bits) - 1;
              const unsigned bit = 1u << pick;
              if (!((unsigned)atomicOr(&cursor[j], (int)bit) & bit))
                slot = first + pick; // else another warp took it: look again
            }
          }
          bank = slot & (banks - 1);
        }
        bank = __shfl_sync(0xffffffffu, bank, turn);
        if (bank >= 0)
          taken[banks == 16 ? turn >> 4 : 0] |= 1u << bank;
      }
      pack[e] = c >= 0 ? ((unsigned)j | ((unsigned)slot << 16)) : 0xffffffffu;
    }
  }
}

} // namespace

int build_tiles6(cfs_matrix_s *m, cudaStream_t s) {
  m->nt6 = 0;
  m->t6_smem_entries = 0;
  // regular (stencil) matrices run variant 5; tiny ones gain nothing
  if (!g_options.tile6 || m->nslices < kT6Slices ||
      m->nregular * 8 >= m->nslices)
    return CFS_OK;
  const long long nt = (m->nslices + kT6Slices - 1) / kT6Slices;
  DevArray<int> lo, hi, count;
  CFS_TRY(lo.alloc((size_t)nt));
  CFS_TRY(hi.alloc((size_t)nt));
  CFS_TRY(count.alloc((size_t)nt));
  tile_bounds_kernel<<<(unsigned)nt, kBoundsThreads, 0, s>>>(
      m->nslices, m->slice_ptr.p, m->sell_col.p, lo.p, hi.p, count.p);
  CFS_CUDA_TRY(cudaGetLastError());
  std::vector<int> hlo((size_t)nt), hhi((size_t)nt), hcount((size_t)nt);
  CFS_CUDA_TRY(cudaMemcpyAsync(hlo.data(), lo.p, (size_t)nt * 4,
                               cudaMemcpyDeviceToHost, s));
  CFS_CUDA_TRY(cudaMemcpyAsync(hhi.data(), hi.p, (size_t)nt * 4,
                               cudaMemcpyDeviceToHost, s));
  CFS_CUDA_TRY(cudaMemcpyAsync(hcount.data(), count.p, (size_t)nt * 4,
                               cudaMemcpyDeviceToHost, s));
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  const int max_entries = (int)(kT6MaxSmemBytes / m->vsize());
  std::vector<int> hncols((size_t)nt);
  std::vector<long long> hoff((size_t)nt + 1, 0);
  int largest = 0, widest = 0;
  for (long long t = 0; t < nt; ++t) {
    const int W = hcount[t] ? hhi[t] - hlo[t] + 1 : 0;
    if (W > kT6MaxCols - 1 || hcount[t] > max_entries || hcount[t] > 65535)
      return CFS_OK; // not a bounded-window matrix: the other variants run
    hncols[t] = W;
    hoff[t + 1] = hoff[t] + W + 1;
    largest = hcount[t] > largest ? hcount[t] : largest;
    widest = W > widest ? W : widest;
  }
  CFS_TRY(m->t6_lo.alloc((size_t)nt));
  CFS_TRY(m->t6_ncols.alloc((size_t)nt));
  CFS_TRY(m->t6_cptr_off.alloc((size_t)nt + 1));
  CFS_TRY(m->t6_cptr.alloc((size_t)hoff[nt]));
  CFS_TRY(m->t6_pack.alloc((size_t)m->padded_entries));
  CFS_CUDA_TRY(cudaMemcpyAsync(m->t6_lo.p, hlo.data(), (size_t)nt * 4,
                               cudaMemcpyHostToDevice, s));
  CFS_CUDA_TRY(cudaMemcpyAsync(m->t6_ncols.p, hncols.data(), (size_t)nt * 4,
                               cudaMemcpyHostToDevice, s));
  CFS_CUDA_TRY(cudaMemcpyAsync(m->t6_cptr_off.p, hoff.data(),
                               ((size_t)nt + 1) * 8, cudaMemcpyHostToDevice,
                               s));
  const int slot_smem = 2 * kT6MaxCols * (int)sizeof(int);
  CFS_CUDA_TRY(cudaFuncSetAttribute(tile_slots_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    slot_smem));
  tile_slots_kernel<<<(unsigned)nt, kSlotThreads, slot_smem, s>>>(
      m->nslices, m->slice_ptr.p, m->sell_col.p, m->t6_lo.p, m->t6_ncols.p,
      m->t6_cptr_off.p, m->t6_cptr.p, m->t6_pack.p,
      g_options.slot_banks ? (m->is_double ? 16 : 32) : 1);
  CFS_CUDA_TRY(cudaGetLastError());
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  m->nt6 = nt;
  m->t6_smem_entries = largest;
  m->t6_max_cols = widest;
  m->t6_cptr_entries = hoff[nt];
  return CFS_OK;
}

} // namespace cfsb
