// spmv_tile.cuh -- variant 6: the transposed term goes through shared memory,
// tile by tile, and reaches y as ONE coalesced reduction per column.
//
// Why: on a banded / FEM-like matrix the 32 lanes of a warp hit 32 unrelated
// columns in every step. The gathers of x survive that (L1/L2), the REDs of
// the transposed term y[col] += a*x[row] do not: 76 M scattered REDs cost
// 240 of the 475 us of the generic kernel on BASELINE configs[3] (8 M rows,
// tools/banded_diag.py) -- the atomics bound the kernel. The reference avoids
// conflicting writes with colours and barriers between threads
// (csr_matrix.tpp:2966-3028); here the same goal is reached inside a CTA:
//
//   tile    = 32 slices = 1024 (length-sorted) rows, one warp per slice
//   phase 1 = the row-major pass of every other variant, except that the
//             product a*x[row] of entry e is STORED to shared memory at
//             slot(e), the entry's position in the COLUMN-major order of the
//             tile (precomputed, tiles6.cu): every entry has its own slot,
//             so plain stores, no atomics, no conflicts
//   phase 2 = one thread per column of the tile's window adds the (contiguous)
//             slots of its column and issues one RED: consecutive threads,
//             consecutive columns -> full-sector reductions, and a column that
//             received k entries from this tile costs 1 RED lane instead of k
//
// Streamed per entry: 8 (value) + 4 (16-bit tile-local column | 16-bit slot)
// bytes, i.e. the algorithmic 12 bytes, + 2 bytes per window column and tile.
#pragma once

#include "common.cuh"
#include "spmv_tma.cuh"

namespace cfsb {
namespace tile6 {

constexpr int kThreads = kT6Slices * 32;

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
__device__ __forceinline__ unsigned ld_stream(const unsigned *p) {
  unsigned v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// Measured on B200 (ncu, profiles/): the kernel is bound by the LSU data pipe
// (wavefronts), not by HBM: per slice-step ~13 wavefronts of global loads (3 of
// stream, ~10 for the 32 scattered x gathers), ~6 for the scattered 64-bit
// product stores, ~8 for phase 2. Dealing the slice-steps of a tile out evenly
// over its warps (the slices of a length-sorted tile are not equally long) was
// tried and lost to its own register spills (485 us vs 421 us).
template <typename T, bool HALO, bool DOT>
__global__ void __launch_bounds__(kThreads, 2)
    sym_spmv_tile_kernel(long long tile_begin, long long nslices, int row_begin,
                         const int *__restrict__ slice_ptr,
                         const int *__restrict__ vrow_row,
                         const unsigned *__restrict__ pack,
                         const T *__restrict__ sell_val,
                         const T *__restrict__ diagonal,
                         const int *__restrict__ tile_lo,
                         const int *__restrict__ tile_ncols,
                         const long long *__restrict__ tile_cptr_off,
                         const unsigned short *__restrict__ cptr,
                         int prod_entries,
                         const T *__restrict__ x, T *__restrict__ y,
                         T *__restrict__ y_lower, double *__restrict__ dot,
                         const T *__restrict__ x_lower,
                         T *__restrict__ y_clear) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *prod = reinterpret_cast<T *>(smem_raw);
  // the slot offsets of the tile's columns, staged before phase 1 so that
  // phase 2 does not start with a round trip to global memory
  unsigned short *cs = reinterpret_cast<unsigned short *>(prod + prod_entries);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long tile = tile_begin + blockIdx.x;
  const long long s = tile * kT6Slices + warp;
  const int lo = tile_lo[tile];
  const int W = tile_ncols[tile];
  {
    const unsigned short *cp = cptr + tile_cptr_off[tile];
    for (int j = threadIdx.x; j <= W; j += kThreads)
      cs[j] = cp[j];
  }

  // ---- phase 1: row-major, products to their column-major slots
  if (s < nslices) {
    const int tag = vrow_row[s * kSliceRows + lane];
    const int p0 = slice_ptr[s], p1 = slice_ptr[s + 1];
    const bool active = tag >= 0;
    const int row = tag & kVrowRowMask;
    T xr = 0, acc = 0;
    if (active) {
      xr = x[row];
      if (!(tag & kVrowCont))
        acc = diagonal[row - row_begin] * xr;
    }
    const T dterm = acc;
    const unsigned *pp = pack + (size_t)p0 * kSliceRows + lane;
    const T *vp = sell_val + (size_t)p0 * kSliceRows + lane;
    int w = p1 - p0;
    for (; w >= 4; w -= 4) {
      unsigned p[4];
      T a[4], xc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        p[u] = ld_stream(pp + u * kSliceRows);
        a[u] = ld_stream(vp + u * kSliceRows);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        xc[u] = p[u] != 0xffffffffu
                    ? tma::x_at<HALO>(x, x_lower, row_begin,
                                      lo + (int)(p[u] & 0xffffu))
                    : T(0);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (p[u] != 0xffffffffu) {
          acc += a[u] * xc[u];
          prod[p[u] >> 16] = a[u] * xr;
        }
      }
      pp += 4 * kSliceRows;
      vp += 4 * kSliceRows;
    }
    for (; w > 0; --w) {
      const unsigned p = ld_stream(pp);
      const T a = ld_stream(vp);
      if (p != 0xffffffffu) {
        acc += a * tma::x_at<HALO>(x, x_lower, row_begin,
                                   lo + (int)(p & 0xffffu));
        prod[p >> 16] = a * xr;
      }
      pp += kSliceRows;
      vp += kSliceRows;
    }
    if (active) {
      tma::red_add(y + row, acc);
      if (y_clear && !(tag & kVrowCont)) // see spmv_reg.cuh
        y_clear[row] = T(0);
    }
    if (DOT) { // x'(A x), see spmv_reg.cuh
      double c = (double)xr * (2.0 * (double)acc - (double)dterm);
#pragma unroll
      for (int o = 16; o; o >>= 1)
        c += __shfl_xor_sync(0xffffffffu, c, o);
      if (lane == 0)
        tma::red_add(dot + (s % kDotSlots) * kDotStride, c);
    }
  }
  __syncthreads();

  // ---- phase 2: column-major, one coalesced RED per column
  for (int j = threadIdx.x; j < W; j += kThreads) {
    const int b = cs[j], e = cs[j + 1];
    if (e > b) {
      T sum = prod[b];
      for (int k = b + 1; k < e; ++k)
        sum += prod[k];
      tma::y_add<HALO>(y, y_lower, row_begin, lo + j, sum);
    }
  }
}

} // namespace tile6
} // namespace cfsb
