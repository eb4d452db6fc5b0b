// spmv_tma.cuh -- PTX helpers shared by the kernels (TMA bulk copies,
// mbarriers, L2 reductions) and variant 2, the persistent TMA-staged kernel.
//
// Variant 2: a tile is <= kTileSlices consecutive slices (<= kTileSteps
// slice-steps); its values, indices, row tags and its 128-byte TileRec are
// contiguous pieces of global memory. Tiles stream through a ring of
// shared-memory stages with cp.async.bulk (TMA engine, SASS UBLKCP), completion
// on an mbarrier; the consumer warps, one slice each, read entries from shared
// memory, gather x through L1/L2 and issue the transposed-term REDs. Every HBM
// request is a large contiguous burst with an L2 evict-first hint, so the
// streamed matrix does not push x and y out of the 126 MB L2.
// Measured (tools/tma_bench.cu): bulk copies issued by ONE warp are serviced one
// after the other (~0.25 us each whatever their size); copies issued by
// different warps overlap. Hence lane 0 of every warp issues a share.
#pragma once

#include "common.cuh"

namespace cfsb {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\t"
               "WAIT_LOOP:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra WAIT_DONE;\n\t"
               "bra WAIT_LOOP;\n\t"
               "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
// streamed once: L2 evict-first
__device__ __forceinline__ void load_1d(void *dst, const void *src,
                                        uint32_t bytes, uint64_t *bar,
                                        uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::"
               "bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
// reused data (x windows): default L2 policy
__device__ __forceinline__ void load_1d_keep(void *dst, const void *src,
                                             uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::"
               "bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void red_add(double *p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void red_add(float *p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// Remote reduction: the target lives in a peer GPU's memory (NVLink, mapped
// through symmetric memory), so the reduction is issued at system scope.
__device__ __forceinline__ void red_add_sys(double *p, double v) {
  asm volatile("red.relaxed.sys.global.add.f64 [%0], %1;" ::"l"(p), "d"(v)
               : "memory");
}
__device__ __forceinline__ void red_add_sys(float *p, float v) {
  asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v)
               : "memory");
}
// y[col] += v. With HALO the columns below row_begin belong to the row block of
// the GPU below: they are reduced straight into ITS y over NVLink (y_lower is
// that vector's virtual base pointer), fused into the SpMV kernel instead of a
// separate exchange step.
template <bool HALO, typename T>
__device__ __forceinline__ void y_add(T *y, T *y_lower, int row_begin, int col,
                                      T v) {
  if (HALO && col < row_begin)
    red_add_sys(y_lower + col, v);
  else
    red_add(y + col, v);
}

// Deterministic mode: a contribution is rounded ONCE to a multiple of
// 1 / *scale (a power of two chosen from |A|, |x| and the longest row so that
// no partial sum can overflow 63 bits) and added with an INTEGER reduction.
// Integer addition is associative, so the value of q[col] does not depend on
// the order in which the warps of the grid arrive -- y is bitwise reproducible
// from run to run, which the order of floating-point REDs is not.
__device__ __forceinline__ void det_add(long long *q, double v, double scale) {
  const long long fixed = __double2ll_rn(v * scale);
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(q), "l"(fixed) : "memory");
}

// x[col]. With HALO the columns below row_begin are rows of the GPU below: they
// are read straight from ITS x over NVLink (x_lower is that vector's virtual
// base pointer; a caller that has pulled the halo into its own vector passes
// its own base) -- no copy step before the kernel.
template <bool HALO, typename T>
__device__ __forceinline__ T x_at(const T *x, const T *x_lower, int row_begin,
                                  int col) {
  if (HALO && col < row_begin)
    return x_lower[col];
  return x[col];
}

constexpr int kTileRows = kTileSlices * kSliceRows;

// The part of a TileRec an issuing lane keeps in registers.
struct IssueInfo {
  int4 head;        // slice_begin, nslices, step_begin, nsteps
  int nwin, row_lo;
  int4 lo0, lo1;    // win_lo[0..7]
  int4 nblk;        // win_nblk[0..7] as 8 x u16
};

template <bool WINDOWS>
__device__ __forceinline__ IssueInfo fetch_issue_info(const TileRec *r) {
  IssueInfo ii;
  const int4 *p = reinterpret_cast<const int4 *>(r);
  ii.head = p[0];
  ii.nwin = 0;
  ii.row_lo = -1;
  if (WINDOWS) {
    ii.nwin = r->nwin;
    ii.row_lo = r->row_lo;
    ii.lo0 = p[4];
    ii.lo1 = p[5];
    ii.nblk = p[6];
  }
  return ii;
}
__device__ __forceinline__ int win_lo_of(const IssueInfo &ii, int j) {
  const int v[8] = {ii.lo0.x, ii.lo0.y, ii.lo0.z, ii.lo0.w,
                    ii.lo1.x, ii.lo1.y, ii.lo1.z, ii.lo1.w};
  return v[j];
}
__device__ __forceinline__ int win_nblk_of(const IssueInfo &ii, int j) {
  const int v[4] = {ii.nblk.x, ii.nblk.y, ii.nblk.z, ii.nblk.w};
  return (v[j >> 1] >> ((j & 1) * 16)) & 0xffff;
}

// Shared-memory stage of variant 2.
template <typename T> struct Stage {
  static constexpr int kVals = kTileSteps * kSliceRows * (int)sizeof(T);
  static constexpr int kCols = kTileSteps * kSliceRows * 4;
  static constexpr int kTags = kTileRows * 4;
  static constexpr int oVals = 0;
  static constexpr int oCols = oVals + kVals;
  static constexpr int oTags = oCols + kCols;
  static constexpr int oRec = oTags + kTags;
  static constexpr int kBytes = oRec + (int)sizeof(TileRec);
  static constexpr int smem_bytes(int stages) { return stages * kBytes + 128; }
};

// MODE: measurement aid, see spmv.cu (non-zero modes compute wrong results).
template <typename T, int STAGES, int MODE>
__global__ void __launch_bounds__(kTileRows)
    sym_spmv_tma_kernel(int ntiles, int row_begin,
                        const TileRec *__restrict__ tile_rec,
                        const int *__restrict__ vrow_row,
                        const int *__restrict__ sell_col,
                        const T *__restrict__ sell_val,
                        const T *__restrict__ diagonal,
                        const T *__restrict__ x, T *__restrict__ y) {
  typedef Stage<T> L;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * L::kBytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int stride = gridDim.x;

  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;"
               : "=l"(policy));
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s)
      mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // the four copies of a tile, one per warp; warp 0 also arms the barrier
  auto issue = [&](int tile, int stage, const int4 head) {
    unsigned char *base = smem + stage * L::kBytes;
    const uint32_t nent = (uint32_t)head.w * kSliceRows;
    const uint32_t vb = nent * (uint32_t)sizeof(T), cb = nent * 4u,
                   tb = (uint32_t)head.y * kSliceRows * 4u;
    if (warp == 0)
      mbar_expect_tx(&full[stage], vb + cb + tb + (uint32_t)sizeof(TileRec));
    const size_t e0 = (size_t)head.z * kSliceRows;
    if (warp == 0 % kTileSlices && nent)
      load_1d(base + L::oVals, sell_val + e0, vb, &full[stage], policy);
    if (warp == 1 % kTileSlices && nent)
      load_1d(base + L::oCols, sell_col + e0, cb, &full[stage], policy);
    if (warp == 2 % kTileSlices)
      load_1d(base + L::oTags, vrow_row + (size_t)head.x * kSliceRows, tb,
              &full[stage], policy);
    if (warp == 3 % kTileSlices)
      load_1d(base + L::oRec, &tile_rec[tile], (uint32_t)sizeof(TileRec),
              &full[stage], policy);
  };

  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      const int t = blockIdx.x + s * stride;
      if (t < ntiles)
        issue(t, s, *reinterpret_cast<const int4 *>(&tile_rec[t]));
    }
  }

  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += stride, ++it) {
    const int stage = it % STAGES;
    const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
    // issuing lanes prefetch the header of the tile they will issue next
    const int t_next = tile + STAGES * stride;
    int4 head_next = make_int4(0, 0, 0, 0);
    if (lane == 0 && t_next < ntiles)
      head_next = *reinterpret_cast<const int4 *>(&tile_rec[t_next]);

    unsigned char *base = smem + stage * L::kBytes;
    mbar_wait(&full[stage], parity);
    const TileRec *rec = reinterpret_cast<const TileRec *>(base + L::oRec);
    if (warp < rec->nslices) {
      const T *vals = reinterpret_cast<const T *>(base + L::oVals);
      const int *cols = reinterpret_cast<const int *>(base + L::oCols);
      const int *tags = reinterpret_cast<const int *>(base + L::oTags);
      const int tag = tags[warp * kSliceRows + lane];
      const bool active = tag >= 0;
      const int row = tag & kVrowRowMask;
      T xr = 0, acc = 0;
      if (active) {
        xr = x[row];
        if (!(tag & kVrowCont))
          acc = diagonal[row - row_begin] * xr;
      }
      int k = rec->slice_step[warp] * kSliceRows + lane;
      const int kend = rec->slice_step[warp + 1] * kSliceRows + lane;
      constexpr int U = 8;
      for (; k + (U - 1) * kSliceRows < kend; k += U * kSliceRows) {
        int c[U];
        T xc[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          c[u] = cols[k + u * kSliceRows];
#pragma unroll
        for (int u = 0; u < U; ++u)
          xc[u] = c[u] >= 0 ? ((MODE & 2) ? xr : x[c[u]]) : T(0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (c[u] >= 0) {
            const T a = vals[k + u * kSliceRows];
            acc += a * xc[u];
            if (!(MODE & 1))
              red_add(y + c[u], a * xr);
            else
              acc += a * xr;
          }
        }
      }
      for (; k < kend; k += kSliceRows) {
        const int c = cols[k];
        if (c >= 0) {
          const T a = vals[k];
          acc += a * ((MODE & 2) ? xr : x[c]);
          if (!(MODE & 1))
            red_add(y + c, a * xr);
        }
      }
      if (active)
        red_add(y + row, acc);
    }
    __syncthreads(); // every consumer is done with this stage
    if (lane == 0 && t_next < ntiles)
      issue(t_next, stage, head_next);
  }
}

} // namespace tma
} // namespace cfsb
