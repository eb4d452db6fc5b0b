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
// MODE is a measurement aid (tools/sweep.py): 0 = the real kernel; bit 0 drops
// the transposed-term REDs, bit 1 replaces the x gathers by a register value.
// Non-zero modes compute WRONG results and are never selected by default.
template <typename T, int MODE = 0>
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
      xc[u] = c[u] >= 0 ? ((MODE & 2) ? xr : x[c[u]]) : T(0);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (c[u] >= 0) {
        acc += a[u] * xc[u];
        if (!(MODE & 1))
          red_add(y + c[u], a[u] * xr);
        else
          acc += a[u] * xr;
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

// ---------------------------------------------------------------------------
// Variant 2: persistent, TMA-staged.
//
// The value / index streams of a tile (<= kTileSlices slices, <= kTileSteps
// slice-steps) are three contiguous pieces of global memory. One elected
// thread fetches them with cp.async.bulk (the TMA engine, SASS UBLKCP) into a
// ring of shared-memory stages, completion signalled on an mbarrier; the
// 8 consumer warps (one slice each) read entries from shared memory, gather x
// through L1/L2 and issue the transposed-term REDs. DRAM latency is hidden by
// the ring depth instead of by occupancy, and every HBM request is a large
// contiguous burst marked L2 evict-first so that x and y stay L2-resident.
// ---------------------------------------------------------------------------
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
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src,
                                            uint32_t bytes, uint64_t *bar,
                                            uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::"
               "bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

template <typename T> struct StageLayout {
  static constexpr int kVals = kTileSteps * kSliceRows * (int)sizeof(T);
  static constexpr int kCols = kTileSteps * kSliceRows * 4;
  static constexpr int kTags = kTileSlices * kSliceRows * 4;
  static constexpr int kBytes = kVals + kCols + kTags;
};

template <typename T, int STAGES, int MODE = 0>
__global__ void __launch_bounds__(kTileSlices * 32)
    sym_spmv_tma_kernel(int ntiles, int row_begin,
                        const int4 *__restrict__ tile_info,
                        const int *__restrict__ slice_ptr,
                        const int *__restrict__ vrow_row,
                        const int *__restrict__ sell_col,
                        const T *__restrict__ sell_val,
                        const T *__restrict__ diagonal,
                        const T *__restrict__ x, T *__restrict__ y) {
  typedef StageLayout<T> L;
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

  auto issue = [&](int tile, int stage) {
    const int4 ti = tile_info[tile];
    unsigned char *base = smem + stage * L::kBytes;
    const uint32_t nent = (uint32_t)ti.w * kSliceRows;
    const uint32_t vb = nent * (uint32_t)sizeof(T), cb = nent * 4u,
                   tb = (uint32_t)ti.y * kSliceRows * 4u;
    mbar_expect_tx(&full[stage], vb + cb + tb);
    const size_t e0 = (size_t)ti.z * kSliceRows;
    if (nent) {
      tma_load_1d(base, sell_val + e0, vb, &full[stage], policy);
      tma_load_1d(base + L::kVals, sell_col + e0, cb, &full[stage], policy);
    }
    tma_load_1d(base + L::kVals + L::kCols,
                vrow_row + (size_t)ti.x * kSliceRows, tb, &full[stage], policy);
  };

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      const int t = blockIdx.x + s * stride;
      if (t < ntiles)
        issue(t, s);
    }
  }

  // per-warp slice bounds are prefetched one tile ahead
  int4 ti = make_int4(0, 0, 0, 0);
  int p0 = 0, p1 = 0;
  if ((int)blockIdx.x < ntiles) {
    ti = tile_info[blockIdx.x];
    if (warp < ti.y) {
      p0 = slice_ptr[ti.x + warp];
      p1 = slice_ptr[ti.x + warp + 1];
    }
  }
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += stride, ++it) {
    const int stage = it % STAGES;
    const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
    // prefetch the next tile's bounds
    int4 ti_n = make_int4(0, 0, 0, 0);
    int p0_n = 0, p1_n = 0;
    if (tile + stride < ntiles) {
      ti_n = tile_info[tile + stride];
      if (warp < ti_n.y) {
        p0_n = slice_ptr[ti_n.x + warp];
        p1_n = slice_ptr[ti_n.x + warp + 1];
      }
    }
    mbar_wait(&full[stage], parity);
    if (warp < ti.y) {
      const unsigned char *base = smem + stage * L::kBytes;
      const T *vals = reinterpret_cast<const T *>(base);
      const int *cols = reinterpret_cast<const int *>(base + L::kVals);
      const int *tags =
          reinterpret_cast<const int *>(base + L::kVals + L::kCols);
      const int tag = tags[warp * kSliceRows + lane];
      const bool active = tag >= 0;
      const int row = tag & kVrowRowMask;
      T xr = 0, acc = 0;
      if (active) {
        xr = x[row];
        if (!(tag & kVrowCont))
          acc = diagonal[row - row_begin] * xr;
      }
      int k = (p0 - ti.z) * kSliceRows + lane;
      const int kend = (p1 - ti.z) * kSliceRows + lane;
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
          acc += a * x[c];
          red_add(y + c, a * xr);
        }
      }
      if (active)
        red_add(y + row, acc);
    }
    __syncthreads(); // every consumer is done with this stage
    if (tid == 0) {
      const int t = tile + STAGES * stride;
      if (t < ntiles)
        issue(t, stage);
    }
    ti = ti_n;
    p0 = p0_n;
    p1 = p1_n;
  }
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

namespace {
constexpr int kStages = 2;

template <typename T>
int launch_tma(const cfs_matrix_s *m, void *y_ext, const void *x_ext,
               cudaStream_t s) {
  static int num_sms = 0;
  static bool configured = false;
  const int smem_bytes = kStages * StageLayout<T>::kBytes + kStages * 8;
  if (!configured) {
    int dev = 0;
    CFS_CUDA_TRY(cudaGetDevice(&dev));
    CFS_CUDA_TRY(cudaDeviceGetAttribute(&num_sms,
                                        cudaDevAttrMultiProcessorCount, dev));
    CFS_CUDA_TRY(cudaFuncSetAttribute(
        sym_spmv_tma_kernel<T, kStages, 0>,
        cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CFS_CUDA_TRY(cudaFuncSetAttribute(
        sym_spmv_tma_kernel<T, kStages, 1>,
        cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CFS_CUDA_TRY(cudaFuncSetAttribute(
        sym_spmv_tma_kernel<T, kStages, 2>,
        cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CFS_CUDA_TRY(cudaFuncSetAttribute(
        sym_spmv_tma_kernel<T, kStages, 3>,
        cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  const long long want = (long long)num_sms * g_options.ctas_per_sm;
  const int grid = (int)(m->ntiles < want ? m->ntiles : want);
  const T *xb = (const T *)x_ext - m->halo_begin;
  T *yb = (T *)y_ext - m->halo_begin;
#define CFS_LAUNCH_TMA(MODE)                                                    \
  sym_spmv_tma_kernel<T, kStages, MODE>                                        \
      <<<grid, kTileSlices * 32, smem_bytes, s>>>(                             \
          (int)m->ntiles, m->row_begin, m->tile_info.p, m->slice_ptr.p,        \
          m->vrow_row.p, m->sell_col.p, (const T *)m->sell_val.p,              \
          (const T *)m->diagonal.p, xb, yb)
  switch (g_options.diag_mode) {
  case 1: CFS_LAUNCH_TMA(1); break;
  case 2: CFS_LAUNCH_TMA(2); break;
  case 3: CFS_LAUNCH_TMA(3); break;
  default: CFS_LAUNCH_TMA(0); break;
  }
#undef CFS_LAUNCH_TMA
  return CFS_OK;
}
} // namespace

int launch_sym_spmv(const cfs_matrix_s *m, void *y_ext, const void *x_ext,
                    cudaStream_t s, cudaEvent_t ev0, cudaEvent_t ev1) {
  const size_t vs = m->vsize();
  const size_t ext_len = (size_t)(m->row_begin + m->nrows - m->halo_begin);
  CFS_CUDA_TRY(cudaMemsetAsync(y_ext, 0, ext_len * vs, s));
  if (m->nslices == 0)
    return CFS_OK;
  if (ev0)
    CFS_CUDA_TRY(cudaEventRecord(ev0, s));
  if (g_options.spmv_variant == 2 && m->ntiles > 0) {
    CFS_TRY(m->is_double ? launch_tma<double>(m, y_ext, x_ext, s)
                         : launch_tma<float>(m, y_ext, x_ext, s));
  } else {
    const unsigned grid = (unsigned)((m->nslices * 32 + kSpmvThreads - 1) /
                                     kSpmvThreads);
    if (m->is_double && g_options.diag_mode) {
      const double *xb = (const double *)x_ext - m->halo_begin;
      double *yb = (double *)y_ext - m->halo_begin;
#define CFS_LAUNCH_SELL(MODE)                                                  \
  sym_spmv_sell_kernel<double, MODE><<<grid, kSpmvThreads, 0, s>>>(            \
      m->nslices, m->row_begin, m->slice_ptr.p, m->vrow_row.p, m->sell_col.p,  \
      (const double *)m->sell_val.p, (const double *)m->diagonal.p, xb, yb)
      switch (g_options.diag_mode) {
      case 1: CFS_LAUNCH_SELL(1); break;
      case 2: CFS_LAUNCH_SELL(2); break;
      default: CFS_LAUNCH_SELL(3); break;
      }
#undef CFS_LAUNCH_SELL
    } else if (m->is_double) {
      const double *xb = (const double *)x_ext - m->halo_begin;
      double *yb = (double *)y_ext - m->halo_begin;
      sym_spmv_sell_kernel<double><<<grid, kSpmvThreads, 0, s>>>(
          m->nslices, m->row_begin, m->slice_ptr.p, m->vrow_row.p,
          m->sell_col.p, (const double *)m->sell_val.p,
          (const double *)m->diagonal.p, xb, yb);
    } else {
      const float *xb = (const float *)x_ext - m->halo_begin;
      float *yb = (float *)y_ext - m->halo_begin;
      sym_spmv_sell_kernel<float><<<grid, kSpmvThreads, 0, s>>>(
          m->nslices, m->row_begin, m->slice_ptr.p, m->vrow_row.p,
          m->sell_col.p, (const float *)m->sell_val.p,
          (const float *)m->diagonal.p, xb, yb);
    }
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
