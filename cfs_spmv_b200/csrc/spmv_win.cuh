// spmv_win.cuh -- variant 4: windowed symmetric SpMV, one CTA per tile.
//
// The matrix streams straight from HBM into registers (coalesced,
// L1-bypassing loads, exactly like variant 1, which reaches the copy-bandwidth
// roofline when nothing else is in the way); what changes is where x is read
// and where the transposed term goes:
//   * the columns a tile touches are covered by a few 32-column-aligned
//     windows (windows.cu). The x windows are fetched into shared memory with
//     cp.async.bulk (TMA engine; one copy per window, issued by different warps
//     so that they overlap), completion on an mbarrier;
//   * y[col] += a*x[row] goes to a shared-memory accumulator with a plain
//     read-modify-write: every slot has ONE owner warp, fixed by the
//     preprocessing -- the reference's conflict-free idea (no two concurrent
//     workers write the same y entry, csr_matrix.tpp:1427-1477) applied inside a
//     CTA. Entries whose slot belongs to another warp are "far": gather + RED;
//   * after the tile the accumulator is flushed with coalesced REDs: one per
//     window slot instead of one per entry (the compressed local vector of the
//     reference's methods 2/3, kept in shared memory).
// A CTA needs 2 * window bytes of shared memory (14 KB for f64), so a dozen
// CTAs share an SM and hide each other's latencies; no persistent pipeline.
#pragma once

#include "common.cuh"
#include "spmv_tma.cuh"

namespace cfsb {
namespace win {

using tma::IssueInfo;
using tma::win_lo_of;
using tma::win_nblk_of;

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

constexpr int kWinSlots = kWindowBlocks * 32;
using tma::kTileRows;
constexpr int kBatch = 7; // entries per lane in flight

template <typename T>
__global__ void __launch_bounds__(kTileRows)
    sym_spmv_win_kernel(int row_begin, const TileRec *__restrict__ tile_rec,
                        const int *__restrict__ vrow_row,
                        const int *__restrict__ sell_slot,
                        const T *__restrict__ sell_val,
                        const T *__restrict__ diagonal,
                        const T *__restrict__ x, T *__restrict__ y) {
  __shared__ __align__(128) T xw[kWinSlots];
  __shared__ __align__(128) T yw[kWinSlots];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const TileRec *rec = &tile_rec[blockIdx.x];
  const IssueInfo ii = tma::fetch_issue_info<true>(rec);

  if (tid == 0) {
    tma::mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  int nslots = 0;
#pragma unroll
  for (int j = 0; j < kMaxWindows; ++j)
    if (j < ii.nwin)
      nslots += win_nblk_of(ii, j) * 32;
  for (int i = tid; i < nslots; i += kTileRows)
    yw[i] = T(0);
  __syncthreads(); // barrier initialised, accumulator clean

  // x windows: one bulk copy per window, window j issued by warp j % 4
  if (lane == 0 && nslots) {
    if (warp == 0)
      tma::mbar_expect_tx(&bar, (uint32_t)nslots * (uint32_t)sizeof(T));
    int off = 0;
#pragma unroll
    for (int j = 0; j < kMaxWindows; ++j) {
      if (j < ii.nwin) {
        const int n = win_nblk_of(ii, j) * 32;
        if (warp == j % kTileSlices)
          tma::load_1d_keep(xw + off, x + win_lo_of(ii, j),
                            (uint32_t)n * sizeof(T), &bar);
        off += n;
      }
    }
  }

  const bool has_slice = warp < ii.head.y;
  int tag = -1;
  if (has_slice)
    tag = vrow_row[((size_t)ii.head.x + warp) * kSliceRows + lane];
  const bool active = tag >= 0;
  const int row = tag & kVrowRowMask;
  T xr = 0, acc = 0;
  if (active) {
    xr = x[row];
    if (!(tag & kVrowCont))
      acc = diagonal[row - row_begin] * xr;
  }
  int k0 = 0, k1 = 0;
  if (has_slice) {
    k0 = rec->slice_step[warp];
    k1 = rec->slice_step[warp + 1];
  }
  const size_t e0 = (size_t)ii.head.z * kSliceRows + lane;
  const int *cp = sell_slot + e0 + (size_t)k0 * kSliceRows;
  const T *vp = sell_val + e0 + (size_t)k0 * kSliceRows;
  int w = k1 - k0;

  // first batch of the stream is in flight while the x windows arrive
  int c[kBatch];
  T a[kBatch];
#pragma unroll
  for (int u = 0; u < kBatch; ++u) {
    c[u] = -1;
    a[u] = T(0);
    if (u < w) {
      c[u] = ld_stream(cp + u * kSliceRows);
      a[u] = ld_stream(vp + u * kSliceRows);
    }
  }
  if (nslots)
    tma::mbar_wait(&bar, 0);

  while (w > 0) {
    // prefetch the next batch before working on this one
    int cn[kBatch];
    T an[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      cn[u] = -1;
      an[u] = T(0);
      if (kBatch + u < w) {
        cn[u] = ld_stream(cp + (kBatch + u) * kSliceRows);
        an[u] = ld_stream(vp + (kBatch + u) * kSliceRows);
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      if (c[u] >= 0) {
        // slot owned by this warp: shared memory, no atomics
        acc += a[u] * xw[c[u]];
        yw[c[u]] += a[u] * xr;
      } else if (c[u] < -1) {
        const int gc = -(c[u] + 2);
        acc += a[u] * x[gc];
        tma::red_add(y + gc, a[u] * xr);
      }
      __syncwarp(); // orders the steps of the warp on shared slots
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      c[u] = cn[u];
      a[u] = an[u];
    }
    cp += kBatch * kSliceRows;
    vp += kBatch * kSliceRows;
    w -= kBatch;
  }
  if (active)
    tma::red_add(y + row, acc);
  __syncthreads(); // all window updates of the tile are in shared memory

  // flush: consecutive threads hold consecutive columns
  int off = 0;
#pragma unroll
  for (int j = 0; j < kMaxWindows; ++j) {
    if (j < ii.nwin) {
      const int n = win_nblk_of(ii, j) * 32;
      T *dst = y + win_lo_of(ii, j);
      for (int i = tid; i < n; i += kTileRows)
        tma::red_add(dst + i, yw[off + i]);
      off += n;
    }
  }
}

} // namespace win
} // namespace cfsb
