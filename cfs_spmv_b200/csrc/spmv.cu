// spmv.cu -- the sm_100a SpMV kernels and their launchers.
//
// The symmetric kernels replace the reference's hot loop,
// cpu_mv_sym_conflict_free_v2 (include/matrix/csr_matrix.tpp:2966-3028) and
// cpu_mv_sym_serial (:2707-2729):  y = A*x with only the lower triangle stored;
// every stored (i, j, a) contributes  y[i] += a*x[j]  and  y[j] += a*x[i].
//
// Common to all variants:
//   * one warp per slice of 32 virtual rows, one lane per virtual row; entry k
//     of all 32 rows is contiguous in memory, so values (256 B for f64) and
//     column ids (128 B) stream from HBM exactly once in full cache lines;
//   * the direct term y[i] += a*x[j] accumulates in a register of the lane that
//     owns the row: no reduction tree for unsplit rows;
//   * y is zeroed by a memset on the same stream before the kernel; the
//     diagonal term is folded into the direct-term accumulator and every
//     contribution reaches y through an L2 reduction (RED / bulk reduce-add).
//
// Variant 1 (this file): one warp per slice, direct L1-bypassing loads.
// Variant 2 (spmv_tma.cuh): persistent, tiles staged in shared memory by TMA.
// Variant 4 (spmv_win.cuh): x / y windows in shared memory, owner warp per slot.
// Variant 5 (spmv_reg.cuh, default): compressed index stream for regular
//   slices, shuffle-merged REDs, optional fused NVLink halo reduction.
#include "common.cuh"
#include "spmv_tma.cuh"
#include "spmv_win.cuh"
#include "spmv_reg.cuh"
#include "spmv_tile.cuh"

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

constexpr int kSpmvThreads = 256;
constexpr int kUnroll = 4;

// MODE is a measurement aid (tools/sweep.py): 0 = the real kernel; bit 0 drops
// the transposed-term REDs, bit 1 replaces the x gathers by a register value.
// Non-zero modes compute WRONG results and are never selected by default.
//
// x and y are "virtual base" pointers: indexed by GLOBAL row/column id (for a
// shard they point halo_begin elements before the extended local vector).
template <typename T, int MODE, bool HALO, bool HUBS, bool DOT = false,
          bool DET = false>
__global__ void __launch_bounds__(kSpmvThreads)
    sym_spmv_sell_kernel(long long slice_begin, long long slice_end,
                         int row_begin,
                         const int *__restrict__ slice_ptr,
                         const int *__restrict__ vrow_row,
                         const int *__restrict__ sell_col,
                         const T *__restrict__ sell_val,
                         const T *__restrict__ diagonal,
                         const T *__restrict__ x, T *__restrict__ y,
                         T *__restrict__ y_lower, double *__restrict__ dot,
                         const T *__restrict__ x_lower,
                         T *__restrict__ y_clear, long long *__restrict__ yq,
                         const double *__restrict__ qscale) {
  const int lane = threadIdx.x & 31;
  const long long s =
      slice_begin + ((blockIdx.x * (long long)kSpmvThreads + threadIdx.x) >> 5);
  if (s >= slice_end)
    return;
  const double scale = DET ? *qscale : 0.0; // deterministic mode, spmv_tma.cuh
  auto emit = [&](int col, T v) {
    if (DET)
      tma::det_add(yq + col, (double)v, scale);
    else
      tma::y_add<HALO>(y, y_lower, row_begin, col, v);
  };
  const int tag = vrow_row[s * kSliceRows + lane];
  const bool active = tag >= 0;
  const int row = tag & kVrowRowMask;
  T xr = 0, acc = 0;
  if (active) {
    xr = x[row];
    if (!(tag & kVrowCont))
      acc = diagonal[row - row_begin] * xr;
  }
  const T dterm = acc;
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
    bool hub[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      // hub columns: the transposed term is done column-wise by
      // hub_spmv_kernel (hubs.cu); only the direct term stays here
      hub[u] = HUBS && c[u] >= 0 && (c[u] & kHubFlag);
      if (HUBS && c[u] >= 0)
        c[u] &= ~kHubFlag;
      xc[u] = c[u] >= 0 ? ((MODE & 2) ? xr
                                      : tma::x_at<HALO>(x, x_lower, row_begin,
                                                        c[u]))
                        : T(0);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (c[u] >= 0) {
        acc += a[u] * xc[u];
        if (MODE & 1)
          acc += a[u] * xr;
        else if (!hub[u])
          emit(c[u], a[u] * xr);
      }
    }
    cp += kUnroll * kSliceRows;
    vp += kUnroll * kSliceRows;
  }
  for (; w > 0; --w) {
    int c = ld_stream(cp);
    const T a = ld_stream(vp);
    if (c >= 0) {
      const bool hub = HUBS && (c & kHubFlag);
      if (HUBS)
        c &= ~kHubFlag;
      acc += a * ((MODE & 2) ? xr
                             : tma::x_at<HALO>(x, x_lower, row_begin, c));
      if (!(MODE & 1) && !hub)
        emit(c, a * xr);
    }
    cp += kSliceRows;
    vp += kSliceRows;
  }
  if (active) {
    if (DET)
      tma::det_add(yq + row, (double)acc, scale);
    else
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
      tma::red_add(dot + (s % reg::kDotSlots) * reg::kDotStride, c);
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

constexpr int kStages = 2; // ring depth of the persistent kernel

// function attributes (dynamic shared memory limits) belong to a device: one
// process may drive several (multi.cu), so what has been granted is remembered
// per device
constexpr int kMaxDevices = 64;
struct PerDevice {
  int v[kMaxDevices] = {};
  int &here() {
    int dev = 0;
    cudaGetDevice(&dev);
    return v[dev < 0 || dev >= kMaxDevices ? 0 : dev];
  }
};

// what only the sharded / ping-pong entry points set: the x vector of the GPU
// below (virtual base, read over NVLink by the halo kernels) and the vector the
// row owners clear for the next SpMV
struct Extras {
  const void *x_lower = nullptr;
  void *y_clear = nullptr;
  long long *yq = nullptr;        // deterministic mode: 64-bit fixed-point y
  const double *qscale = nullptr; // ... and its scale (device)
};

template <typename T, int MODE>
void launch_sell(const cfs_matrix_s *m, const T *xb, T *yb, T *y_lower,
                 cudaStream_t s, long long s0, long long s1, const Extras &ex,
                 double *dot = nullptr) {
  const unsigned grid =
      (unsigned)(((s1 - s0) * 32 + kSpmvThreads - 1) / kSpmvThreads);
  // hub columns: flagged column stream + a second, column-wise kernel
  const bool hubs = MODE == 0 && !y_lower && m->nhubs > 0 && g_options.hubs &&
                    s0 == 0 && s1 == m->nslices;
  const T *xl = ex.x_lower ? (const T *)ex.x_lower : xb;
  T *yc = (T *)ex.y_clear;
#define CFS_LAUNCH_SELL(M, HALO, HUBS, DOT, COLS, YL, DOTP)                    \
  sym_spmv_sell_kernel<T, M, HALO, HUBS, DOT><<<grid, kSpmvThreads, 0, s>>>(   \
      s0, s1, m->row_begin, m->slice_ptr.p, m->vrow_row.p, COLS,               \
      (const T *)m->sell_val.p, (const T *)m->diagonal.p, xb, yb, YL, DOTP,    \
      xl, yc, nullptr, nullptr)
  if (ex.yq) { // deterministic mode: one kernel, integer reductions
    sym_spmv_sell_kernel<T, 0, false, false, false, true>
        <<<grid, kSpmvThreads, 0, s>>>(
            s0, s1, m->row_begin, m->slice_ptr.p, m->vrow_row.p, m->sell_col.p,
            (const T *)m->sell_val.p, (const T *)m->diagonal.p, xb, yb, nullptr,
            nullptr, xl, yc, ex.yq, ex.qscale);
    return;
  }
  if (y_lower && dot)
    CFS_LAUNCH_SELL(0, true, false, true, m->sell_col.p, y_lower, dot);
  else if (y_lower)
    CFS_LAUNCH_SELL(MODE, true, false, false, m->sell_col.p, y_lower, nullptr);
  else if (hubs && dot)
    CFS_LAUNCH_SELL(0, false, true, true, m->hub_colstream.p, nullptr, dot);
  else if (hubs)
    CFS_LAUNCH_SELL(MODE, false, true, false, m->hub_colstream.p, nullptr,
                    nullptr);
  else if (dot)
    CFS_LAUNCH_SELL(0, false, false, true, m->sell_col.p, nullptr, dot);
  else
    CFS_LAUNCH_SELL(MODE, false, false, false, m->sell_col.p, nullptr, nullptr);
#undef CFS_LAUNCH_SELL
  if (hubs)
    launch_hub_spmv(m, yb, xb, s);
}

template <typename T, int MODE>
int launch_tma(const cfs_matrix_s *m, const T *xb, T *yb, cudaStream_t s) {
  static PerDevice sms_of, ctas_of;
  int &num_sms = sms_of.here(), &max_ctas = ctas_of.here();
  const int smem_bytes = tma::Stage<T>::smem_bytes(kStages);
  auto kernel = tma::sym_spmv_tma_kernel<T, kStages, MODE>;
  if (!num_sms) {
    int dev = 0;
    CFS_CUDA_TRY(cudaGetDevice(&dev));
    CFS_CUDA_TRY(cudaFuncSetAttribute(
        kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CFS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &max_ctas, kernel, kTileSlices * 32, smem_bytes));
    if (max_ctas < 1) {
      set_error("persistent kernel does not fit on an SM (%d B smem)",
                smem_bytes);
      return CFS_ERR_CUDA;
    }
    CFS_CUDA_TRY(cudaDeviceGetAttribute(&num_sms,
                                        cudaDevAttrMultiProcessorCount, dev));
  }
  const int per_sm =
      g_options.ctas_per_sm < max_ctas ? g_options.ctas_per_sm : max_ctas;
  const long long want = (long long)num_sms * per_sm;
  const int grid = (int)(m->ntiles < want ? m->ntiles : want);
  kernel<<<grid, kTileSlices * 32, smem_bytes, s>>>(
      (int)m->ntiles, m->row_begin, m->tile_rec.p, m->vrow_row.p,
      m->sell_col.p, (const T *)m->sell_val.p, (const T *)m->diagonal.p, xb,
      yb);
  return CFS_OK;
}

template <typename T>
int launch_reg(const cfs_matrix_s *m, const T *xb, T *yb, T *y_lower,
               cudaStream_t s, long long s0, long long s1, const Extras &ex,
               double *dot = nullptr) {
  const T *xl = ex.x_lower ? (const T *)ex.x_lower : xb;
  T *yc = (T *)ex.y_clear;
  const unsigned grid =
      (unsigned)(((s1 - s0) * 32 + reg::kThreads - 1) / reg::kThreads);
  // Row chunks of the host-vector pipeline run next to PCIe copies: a kernel
  // that saturates HBM starves the copy engines (tools/e2e_probe.py), so these
  // launches ask for unused shared memory to cap the resident CTAs per SM.
  int smem = 0;
  if (!y_lower && (s0 != 0 || s1 != m->nslices) && g_options.pipeline_smem) {
    smem = g_options.pipeline_smem;
    static PerDevice granted_of;
    int &granted = granted_of.here();
    if (granted < smem) {
      CFS_CUDA_TRY(cudaFuncSetAttribute(
          reg::sym_spmv_reg_kernel<T, false>,
          cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      granted = smem;
    }
  }
  // value indexing (valindex.cu): 1 = dictionary + one-byte codes, 2 = one value
  const int vi = !g_options.value_index || m->ndict == 0 ? 0
                 : m->ndict == 1                        ? 2
                                                        : 1;
#define CFS_LAUNCH_REG(HALO, DOT, VI, PF, SMEM, YL, DOTP)                      \
  reg::sym_spmv_reg_kernel<T, HALO, DOT, VI, PF>                               \
      <<<grid, reg::kThreads, SMEM, s>>>(                                      \
          s0, s1, m->row_begin, m->slice_ptr.p, m->slice_cptr.p,               \
          m->vrow_row.p, m->ccol.p, (const T *)m->sell_val.p,                  \
          (const T *)m->diagonal.p, xb, yb, YL, DOTP, m->vcode.p,              \
          (const T *)m->vdict.p, m->ndict, 0, xl, yc, nullptr, nullptr)
#define CFS_LAUNCH_BULK(HALO, DOT, YL, DOTP)                                   \
  do {                                                                         \
    auto kernel = reg::sym_spmv_reg_kernel<T, HALO, DOT, 0, false, true>;      \
    static PerDevice granted_of; /* per instantiation */                       \
    int &granted_bulk = granted_of.here();                                     \
    if (granted_bulk < bulk_smem) {                                            \
      CFS_CUDA_TRY(cudaFuncSetAttribute(                                       \
          kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bulk_smem));    \
      granted_bulk = bulk_smem;                                                \
    }                                                                          \
    kernel<<<grid, reg::kThreads, bulk_smem, s>>>(                             \
        s0, s1, m->row_begin, m->slice_ptr.p, m->slice_cptr.p, m->vrow_row.p,  \
        m->ccol.p, (const T *)m->sell_val.p, (const T *)m->diagonal.p, xb, yb, \
        YL, DOTP, m->vcode.p, (const T *)m->vdict.p, m->ndict,                 \
        m->max_slice_steps, xl, yc, nullptr, nullptr);                         \
  } while (0)
#define CFS_LAUNCH_REG_VI(HALO, DOT, SMEM, YL, DOTP)                           \
  do {                                                                         \
    if (vi == 2)                                                               \
      CFS_LAUNCH_REG(HALO, DOT, 2, false, SMEM, YL, DOTP);                     \
    else if (vi == 1)                                                          \
      CFS_LAUNCH_REG(HALO, DOT, 1, false, SMEM, YL, DOTP);                     \
    else if (g_options.l2_prefetch && g_options.reg_blocks == 12)              \
      reg::sym_spmv_reg_kernel<T, HALO, DOT, 0, true, false, 12>               \
          <<<grid, reg::kThreads, SMEM, s>>>(                                  \
              s0, s1, m->row_begin, m->slice_ptr.p, m->slice_cptr.p,           \
              m->vrow_row.p, m->ccol.p, (const T *)m->sell_val.p,              \
              (const T *)m->diagonal.p, xb, yb, YL, DOTP, m->vcode.p,          \
              (const T *)m->vdict.p, m->ndict, 0, xl, yc, nullptr, nullptr);   \
    else if (g_options.l2_prefetch)                                            \
      CFS_LAUNCH_REG(HALO, DOT, 0, true, SMEM, YL, DOTP);                      \
    else                                                                       \
      CFS_LAUNCH_REG(HALO, DOT, 0, false, SMEM, YL, DOTP);                     \
  } while (0)
  if (ex.yq) { // deterministic mode (no halo fusion, no x'Ax: the caller checks)
#define CFS_LAUNCH_DET(VI)                                                     \
  reg::sym_spmv_reg_kernel<T, false, false, VI, false, false, 16, true>        \
      <<<grid, reg::kThreads, 0, s>>>(                                         \
          s0, s1, m->row_begin, m->slice_ptr.p, m->slice_cptr.p,               \
          m->vrow_row.p, m->ccol.p, (const T *)m->sell_val.p,                  \
          (const T *)m->diagonal.p, xb, yb, nullptr, nullptr, m->vcode.p,      \
          (const T *)m->vdict.p, m->ndict, 0, xl, yc, ex.yq, ex.qscale)
    if (vi == 2)
      CFS_LAUNCH_DET(2);
    else if (vi == 1)
      CFS_LAUNCH_DET(1);
    else
      CFS_LAUNCH_DET(0);
#undef CFS_LAUNCH_DET
    return CFS_OK;
  }
  // variant 7: value blocks staged by the TMA engine (streamed values only)
  const int bulk_smem =
      128 + (reg::kThreads / 32) * m->max_slice_steps * kSliceRows * (int)sizeof(T);
  if (g_options.spmv_variant == 7 && vi == 0 && !smem && m->max_slice_steps > 0 &&
      bulk_smem <= 200 * 1024) {
    if (y_lower && dot)
      CFS_LAUNCH_BULK(true, true, y_lower, dot);
    else if (y_lower)
      CFS_LAUNCH_BULK(true, false, y_lower, nullptr);
    else if (dot)
      CFS_LAUNCH_BULK(false, true, nullptr, dot);
    else
      CFS_LAUNCH_BULK(false, false, nullptr, nullptr);
    return CFS_OK;
  }
  if (y_lower && dot)
    CFS_LAUNCH_REG_VI(true, true, 0, y_lower, dot);
  else if (y_lower)
    CFS_LAUNCH_REG_VI(true, false, 0, y_lower, nullptr);
  else if (dot)
    CFS_LAUNCH_REG_VI(false, true, 0, nullptr, dot);
  else if (smem) // occupancy-capped pipeline launches (measurement knob)
    CFS_LAUNCH_REG(false, false, 0, false, smem, nullptr, nullptr);
  else
    CFS_LAUNCH_REG_VI(false, false, 0, nullptr, nullptr);
#undef CFS_LAUNCH_REG_VI
#undef CFS_LAUNCH_REG
#undef CFS_LAUNCH_BULK
  return CFS_OK;
}

template <typename T, bool HALO, bool DOT>
int launch_tile6_as(const cfs_matrix_s *m, const T *xb, T *yb, T *y_lower,
                    double *dot, cudaStream_t s, long long s0, long long s1,
                    const Extras &ex) {
  // whole tiles only: [s0, s1) starts on a tile boundary (the caller checks)
  const long long tile0 = s0 / kT6Slices;
  const long long ntiles = (s1 - s0 + kT6Slices - 1) / kT6Slices;
  auto kernel = tile6::sym_spmv_tile_kernel<T, HALO, DOT>;
  // products of the largest tile (rounded to 8 entries) + its slot offsets
  const int prod_entries = (m->t6_smem_entries + 7) & ~7;
  const int smem = prod_entries * (int)sizeof(T) + (m->t6_max_cols + 2) * 2;
  static PerDevice granted_of; // per instantiation
  int &granted = granted_of.here();
  if (granted < smem) {
    const int cap = kT6MaxSmemBytes + (kT6MaxCols + 2) * 2;
    CFS_CUDA_TRY(cudaFuncSetAttribute(
        kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    granted = cap;
  }
  kernel<<<(unsigned)ntiles, tile6::kThreads, smem, s>>>(
      tile0, s1, m->row_begin, m->slice_ptr.p, m->vrow_row.p, m->t6_pack.p,
      (const T *)m->sell_val.p, (const T *)m->diagonal.p, m->t6_lo.p,
      m->t6_ncols.p, m->t6_cptr_off.p, m->t6_cptr.p, prod_entries, xb, yb,
      y_lower, dot, ex.x_lower ? (const T *)ex.x_lower : xb, (T *)ex.y_clear);
  return CFS_OK;
}

template <typename T>
int launch_tile6(const cfs_matrix_s *m, const T *xb, T *yb, T *y_lower,
                 double *dot, cudaStream_t s, long long s0, long long s1,
                 const Extras &ex) {
  if (y_lower && dot)
    return launch_tile6_as<T, true, true>(m, xb, yb, y_lower, dot, s, s0, s1,
                                          ex);
  if (y_lower)
    return launch_tile6_as<T, true, false>(m, xb, yb, y_lower, nullptr, s, s0,
                                           s1, ex);
  if (dot)
    return launch_tile6_as<T, false, true>(m, xb, yb, nullptr, dot, s, s0, s1,
                                           ex);
  return launch_tile6_as<T, false, false>(m, xb, yb, nullptr, nullptr, s, s0,
                                          s1, ex);
}

template <typename T>
int launch_win(const cfs_matrix_s *m, const T *xb, T *yb, cudaStream_t s) {
  win::sym_spmv_win_kernel<T><<<(unsigned)m->ntiles, kTileSlices * 32, 0, s>>>(
      m->row_begin, m->tile_rec.p, m->vrow_row.p, m->sell_slot.p,
      (const T *)m->sell_val.p, (const T *)m->diagonal.p, xb, yb);
  return CFS_OK;
}

template <typename T>
int launch_sym_typed(const cfs_matrix_s *m, void *y_ext, const void *x_ext,
                     void *y_lower_base, cudaStream_t s, long long s0,
                     long long s1, double *dot, Extras ex) {
  // A shard with a fused halo: only its first slices reach below row_begin, the
  // rest runs the plain instantiation (no per-column test against row_begin).
  // Only where that test sits in the inner loop -- the tile kernel of banded
  // matrices (4-GPU step 0.45 -> 0.41 ms); in the register kernel it is one
  // warp-uniform compare per chain, and two launches in a row cost more there
  // (27-pt, 8 GPUs: step minus kernel 24 -> 38 us with the split).
  if (y_lower_base && !dot && m->halo_slice_end < s1 && m->nt6 > 0 &&
      g_options.tile6) {
    const long long hs = m->halo_slice_end > s0 ? m->halo_slice_end : s0;
    if (hs > s0)
      CFS_TRY(launch_sym_typed<T>(m, y_ext, x_ext, y_lower_base, s, s0, hs, dot,
                                  ex));
    Extras rest = ex;
    rest.x_lower = nullptr;
    return launch_sym_typed<T>(m, y_ext, x_ext, nullptr, s, hs, s1, dot, rest);
  }
  // the row owners index y_clear by global row id like y
  if (ex.y_clear)
    ex.y_clear = (T *)ex.y_clear - m->halo_begin;
  const T *xb = (const T *)x_ext - m->halo_begin;
  T *yb = (T *)y_ext - m->halo_begin;
  T *yl = (T *)y_lower_base;
  const int mode = g_options.diag_mode;
  int variant = g_options.spmv_variant;
  if (variant == 7)
    variant = 5; // the same kernel with TMA-staged value blocks (launch_reg)
  const bool partial = s0 != 0 || s1 != m->nslices;
  // bounded column windows (banded / FEM orderings): products transposed
  // through shared memory, one coalesced RED per column and tile
  // (slice ranges of the host-vector pipeline: whole tiles only)
  if (ex.yq) { // deterministic mode: the register kernels with integer REDs
    ex.yq -= m->halo_begin;
    if (m->ccol.p && m->nregular * 8 >= m->nslices)
      return launch_reg<T>(m, xb, yb, nullptr, s, s0, s1, ex, nullptr);
    if (!m->sell_col.p && m->padded_entries) {
      set_error("deterministic mode needs the uncompressed column stream here: "
                "set option keep_layouts=1 before cfs_cuda_matrix_tune");
      return CFS_ERR_STATE;
    }
    launch_sell<T, 0>(m, xb, yb, nullptr, s, s0, s1, ex, nullptr);
    return CFS_OK;
  }
  if ((variant == 5 || variant == 6) && m->nt6 > 0 && g_options.tile6 &&
      mode == 0 && s0 % kT6Slices == 0 &&
      (s1 % kT6Slices == 0 || s1 == m->nslices))
    return launch_tile6<T>(m, xb, yb, yl, dot, s, s0, s1, ex);
  if ((yl || partial || dot || ex.y_clear) && variant != 1)
    variant = 5; // halo fusion / slice ranges / x'Ax exist in the register kernels
  // bulk copies of the x / y windows need 16-byte aligned vectors
  // the compressed-index kernel pays off when slices are regular; ragged
  // matrices run the generic warp-per-slice kernel (more registers, no spills)
  if (variant == 5 && m->ccol.p && mode == 0 &&
      m->nregular * 8 >= m->nslices)
    return launch_reg<T>(m, xb, yb, yl, s, s0, s1, ex, dot);
  if (variant == 5 || yl || partial || dot || ex.y_clear)
    variant = 1;
  if (!m->sell_col.p && m->padded_entries) {
    set_error("SpMV variant %d needs the uncompressed column stream, which was "
              "released after tune: set option keep_layouts=1 before "
              "cfs_cuda_matrix_tune", variant);
    return CFS_ERR_STATE;
  }
  if (dot) {
    launch_sell<T, 0>(m, xb, yb, yl, s, s0, s1, ex, dot);
    return CFS_OK;
  }
  if (variant >= 3 &&
      (!m->sell_slot.p || (((uintptr_t)x_ext | (uintptr_t)y_ext) & 15)))
    variant = 2;
  if (variant >= 2 && m->ntiles == 0)
    variant = 1;
  if (variant == 4)
    return launch_win<T>(m, xb, yb, s);
  if (variant == 3) // the persistent windowed kernel was retired for 4
    return launch_win<T>(m, xb, yb, s);
  if (variant == 2) {
    switch (mode) {
    case 1:
      return launch_tma<T, 1>(m, xb, yb, s);
    case 2:
      return launch_tma<T, 2>(m, xb, yb, s);
    case 3:
      return launch_tma<T, 3>(m, xb, yb, s);
    default:
      return launch_tma<T, 0>(m, xb, yb, s);
    }
  }
  switch (mode) {
  case 1:
    launch_sell<T, 1>(m, xb, yb, yl, s, s0, s1, ex);
    break;
  case 2:
    launch_sell<T, 2>(m, xb, yb, yl, s, s0, s1, ex);
    break;
  case 3:
    launch_sell<T, 3>(m, xb, yb, yl, s, s0, s1, ex);
    break;
  default:
    launch_sell<T, 0>(m, xb, yb, yl, s, s0, s1, ex);
    break;
  }
  return CFS_OK;
}

} // namespace

int launch_sym_spmv(const cfs_matrix_s *m, void *y_ext, const void *x_ext,
                    cudaStream_t s, cudaEvent_t ev0, cudaEvent_t ev1,
                    void *y_lower_base, bool y_is_zero, long long slice0,
                    long long slice1, double *xdoty, const void *x_lower_base,
                    void *y_clear) {
  Extras ex;
  ex.x_lower = x_lower_base;
  ex.y_clear = y_clear;
  // deterministic mode (det.cu): whole-matrix launches on one GPU
  const bool whole = slice0 == 0 && (slice1 < 0 || slice1 == m->nslices);
  if (m->far && (!whole || y_lower_base || xdoty || g_options.deterministic)) {
    set_error("a Format::hyb matrix runs whole, on one GPU, without x'Ax and "
              "not in deterministic mode (its far part is a second kernel)");
    return CFS_ERR_STATE;
  }
  const bool det = g_options.deterministic && !y_lower_base && !xdoty &&
                   slice0 == 0 && (slice1 < 0 || slice1 == m->nslices) &&
                   m->nslices > 0;
  if (g_options.deterministic && !det && m->nslices > 0) {
    set_error("deterministic mode covers cfs_cuda_spmv / cfs_cuda_spmv_async on "
              "one GPU (no fused halo, no x'Ax, no host-vector pipeline: set "
              "option pipeline=0)");
    return CFS_ERR_STATE;
  }
  if (det) {
    CFS_TRY(det_prepare(m, x_ext, s));
    ex.yq = m->det_acc.p;
    ex.qscale = m->det_scale.p;
    y_is_zero = true; // every row of y is written by det_finish
  }
  if (slice1 < 0)
    slice1 = m->nslices;
  const size_t vs = m->vsize();
  const size_t ext_len = (size_t)(m->row_begin + m->nrows - m->halo_begin);
  if (!y_is_zero)
    CFS_CUDA_TRY(cudaMemsetAsync(y_ext, 0, ext_len * vs, s));
  if (m->nslices == 0 || slice1 <= slice0) {
    if (m->far)
      CFS_TRY(launch_csr_sell(m->far, y_ext, x_ext, s, true));
    return CFS_OK;
  }
  if (ev0)
    CFS_CUDA_TRY(cudaEventRecord(ev0, s));
  CFS_TRY(m->is_double
              ? launch_sym_typed<double>(m, y_ext, x_ext, y_lower_base, s,
                                         slice0, slice1, xdoty, ex)
              : launch_sym_typed<float>(m, y_ext, x_ext, y_lower_base, s,
                                        slice0, slice1, xdoty, ex));
  CFS_CUDA_TRY(cudaGetLastError());
  if (det)
    CFS_TRY(det_finish(m, y_ext, s));
  if (m->far) // Format::hyb: the far part adds onto y (hyb.cu)
    CFS_TRY(launch_csr_sell(m->far, y_ext, x_ext, s, true));
  if (ev1)
    CFS_CUDA_TRY(cudaEventRecord(ev1, s));
  return CFS_OK;
}

int launch_csr_spmv(const cfs_matrix_s *m, void *y, const void *x,
                    cudaStream_t s) {
  if (m->nrows == 0)
    return CFS_OK;
  if (m->sell_val.p && g_options.csr_layout) // sliced layout (csr_path.cu)
    return launch_csr_sell(m, y, x, s);
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
