// spmv_reg.cuh -- variant 5: one warp per slice, compressed column stream,
// shuffle-merged transposed updates for regular slices.
//
// Regular slice (compress.cu): lane l owns row r0+l, step k hits column
// base_k + l. Only the w bases are read (one 128-byte row instead of w), the
// value stream is the only thing that still costs 8 bytes/entry of HBM traffic.
// Steps whose bases are consecutive (base_{k+1} = base_k + 1 -- the three
// x-neighbours of a stencil) form a CHAIN of length L <= kMaxChain. All
// transposed updates of a chain land in y[base .. base+32+L-1): lane m collects
//     S[m] = sum_j p_j[m - j]          (p_j = a_j * x[row], one rotate-shuffle
//                                       per j)
// and the L-1 wrapped values go to lanes 0..L-2 as E. One full-warp RED plus one
// RED with L-1 active lanes replace L full-warp REDs: the GPU restatement of the
// reference's goal -- fewer, conflict-free writes of the transposed term
// (csr_matrix.tpp:3005-3013) -- without any shared memory.
// Irregular slices run the generic loop of variant 1 from the same stream.
#pragma once

#include "common.cuh"
#include "spmv_tma.cuh"

namespace cfsb {
namespace reg {

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

constexpr int kThreads = 128;
constexpr int kMaxChain = 3;
constexpr int kUnroll = 4;

// 8 CTAs of 256 threads per SM: the kernel is latency-bound, occupancy wins over
// registers (measured: 275 us at 40 registers, 239 us at 32, 231 us with 128-thread CTAs)
// DOT: the kernel also returns x'(A x) -- for a symmetric matrix that is
// sum_i x_i * (2 * direct_i - d_i * x_i), and direct_i (diagonal + lower part
// of row i times x) is exactly what the lane that owns row i holds in `acc`, so
// the p'Ap of a conjugate-gradient step costs one warp reduction and one RED
// per warp instead of a pass over two vectors (cg.cu). dot[] has kDotSlots
// partial sums, 32 bytes apart, to keep same-address reductions rare.
using cfsb::kDotSlots;
using cfsb::kDotStride;

// Tried and dropped for the value-indexed instantiations: software-pipelining
// the x gathers of the next chain behind the shuffles / REDs of the current one
// (the kernel waits on those gathers once the value stream is gone). At the 32
// registers that 64 warps/SM allow it spills, and lost: 152 -> 189 us; at 40
// registers (48 warps/SM, no spills) it lost too: 172 us. More warps beat more
// loads in flight per warp, again.
//
// VI (value indexing, valindex.cu): 0 = values are streamed (8 bytes/entry),
// 1 = one byte per entry names the value in a dictionary of <= 256 distinct
// values kept in shared memory, 2 = the lower triangle holds ONE distinct value,
// nothing but the indices is streamed. Same bits multiplied either way.
__device__ __forceinline__ unsigned ld_stream(const unsigned char *p) {
  unsigned v;
  asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

template <typename T, bool HALO, bool DOT = false, int VI = 0>
__global__ void __launch_bounds__(kThreads, 16)
    sym_spmv_reg_kernel(long long slice_begin, long long slice_end,
                        int row_begin,
                        const int *__restrict__ slice_ptr,
                        const int *__restrict__ slice_cptr,
                        const int *__restrict__ vrow_row,
                        const int *__restrict__ ccol,
                        const T *__restrict__ sell_val,
                        const T *__restrict__ diagonal,
                        const T *__restrict__ x, T *__restrict__ y,
                        T *__restrict__ y_lower, double *__restrict__ dot,
                        const unsigned char *__restrict__ vcode,
                        const T *__restrict__ vdict, int ndict) {
  __shared__ T sdict[VI == 1 ? kMaxDict : 1];
  T one_value = T(0);
  if (VI == 1) {
    for (int k = threadIdx.x; k < ndict; k += kThreads)
      sdict[k] = vdict[k];
    __syncthreads();
  } else if (VI == 2) {
    one_value = vdict[0];
  }
  const int lane = threadIdx.x & 31;
  const long long s =
      slice_begin + ((blockIdx.x * (long long)kThreads + threadIdx.x) >> 5);
  if (s >= slice_end)
    return;
  const int tag = vrow_row[s * kSliceRows + lane];
  const int p0 = slice_ptr[s], p1 = slice_ptr[s + 1];
  const int cptr = slice_cptr[s];
  const bool active = tag >= 0;
  const int row = tag & kVrowRowMask;
  T xr = 0, acc = 0;
  if (active) {
    xr = x[row];
    if (!(tag & kVrowCont))
      acc = diagonal[row - row_begin] * xr;
  }
  const T dterm = acc; // d_i * x_i (0 for a continuation chunk)
  const T *vp = sell_val + (size_t)p0 * kSliceRows + lane;
  const unsigned char *kp = vcode + (size_t)p0 * kSliceRows + lane;
  const int *cp = ccol + (size_t)(cptr & ~kSliceRegular) * kSliceRows + lane;
  int w = p1 - p0;
  // the value of the entry `step` steps further down this lane's row
  auto value_at = [&](int step) -> T {
    if (VI == 2)
      return one_value;
    if (VI == 1)
      return sdict[ld_stream(kp + (size_t)step * kSliceRows)];
    return ld_stream(vp + (size_t)step * kSliceRows);
  };

  if (cptr & kSliceRegular) {
    // lane k holds the base column of step k
    const int base = ld_stream(cp);
    const int nb = __shfl_down_sync(0xffffffffu, base, 1);
    const unsigned cont =
        __ballot_sync(0xffffffffu, lane + 1 < w && nb == base + 1);
    for (int k = 0; k < w;) {
      int L = 1;
      while (L < kMaxChain && ((cont >> (k + L - 1)) & 1u))
        ++L;
      const int cbase = __shfl_sync(0xffffffffu, base, k);
      T a[kMaxChain], xc[kMaxChain];
#pragma unroll
      for (int j = 0; j < kMaxChain; ++j) {
        a[j] = T(0);
        xc[j] = T(0);
        if (j < L) {
          a[j] = value_at(k + j);
          xc[j] = x[cbase + lane + j];
        }
      }
      T S = T(0), E = T(0);
#pragma unroll
      for (int j = 0; j < kMaxChain; ++j) {
        if (j < L) { // warp-uniform
          acc += a[j] * xc[j];
          const T p = a[j] * xr;
          const T r = j == 0 ? p : __shfl_sync(0xffffffffu, p, (lane - j) & 31);
          if (lane >= j)
            S += r;
          else
            E += r;
        }
      }
      if (HALO && cbase < row_begin) { // warp-uniform: chain reaches the halo
        tma::y_add<true>(y, y_lower, row_begin, cbase + lane, S);
        if (lane < L - 1)
          tma::y_add<true>(y, y_lower, row_begin, cbase + kSliceRows + lane, E);
      } else {
        tma::red_add(y + cbase + lane, S);
        if (lane < L - 1)
          tma::red_add(y + cbase + kSliceRows + lane, E);
      }
      k += L;
    }
  } else {
    for (; w >= kUnroll; w -= kUnroll) {
      int c[kUnroll];
      T a[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        c[u] = ld_stream(cp + u * kSliceRows);
        a[u] = value_at(u);
      }
      T xc[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        xc[u] = c[u] >= 0 ? x[c[u]] : T(0);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (c[u] >= 0) {
          acc += a[u] * xc[u];
          tma::y_add<HALO>(y, y_lower, row_begin, c[u], a[u] * xr);
        }
      }
      cp += kUnroll * kSliceRows;
      vp += kUnroll * kSliceRows;
      kp += kUnroll * kSliceRows;
    }
    for (; w > 0; --w) {
      const int c = ld_stream(cp);
      const T a = value_at(0);
      if (c >= 0) {
        acc += a * x[c];
        tma::y_add<HALO>(y, y_lower, row_begin, c, a * xr);
      }
      cp += kSliceRows;
      vp += kSliceRows;
      kp += kSliceRows;
    }
  }
  if (active)
    tma::red_add(y + row, acc);
  if (DOT) {
    double c = (double)xr * (2.0 * (double)acc - (double)dterm);
#pragma unroll
    for (int o = 16; o; o >>= 1)
      c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0)
      tma::red_add(dot + (s % kDotSlots) * kDotStride, c);
  }
}

} // namespace reg
} // namespace cfsb
