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

// S += r on the lanes where `upper` holds, E += r on the others -- as two
// predicated adds (the compiler's if-conversion makes two adds and four selects)
__device__ __forceinline__ void split_add(double &S, double &E, bool upper,
                                          double r) {
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t"
      "@p add.f64 %0, %0, %3;\n\t@!p add.f64 %1, %1, %3;\n\t}"
      : "+d"(S), "+d"(E)
      : "r"((int)upper), "d"(r));
}
__device__ __forceinline__ void split_add(float &S, float &E, bool upper,
                                          float r) {
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t"
      "@p add.f32 %0, %0, %3;\n\t@!p add.f32 %1, %1, %3;\n\t}"
      : "+f"(S), "+f"(E)
      : "r"((int)upper), "f"(r));
}

// PF: the warp asks L2 for its slice's whole value block (w x 256 B for f64)
// before it touches the first entry -- one prefetch.global.L2 per 128-byte line,
// dealt out over the lanes. The kernel is bound by bytes in flight, not by
// DRAM: at 32 registers a warp can hold one chain (<= 768 B) of value loads,
// and the prefetches put the other ~2.5 KB of the slice on their way without a
// single register. The loads behind them then meet L2 latency instead of DRAM
// latency.
//
// BULK (variant 7): instead, lane 0 hands the whole value block of the slice to
// the TMA engine -- ONE cp.async.bulk (SASS UBLKCP) into this warp's piece of
// shared memory, completion on a per-warp mbarrier -- and the warp goes on
// loading its row tags, x[row], the diagonal and the column bases while the
// block is in flight. bulk_steps = the widest slice of the matrix (the stride
// of a warp's piece). Copies of different warps overlap (tools/tma_bench.cu).
template <typename T, bool HALO, bool DOT = false, int VI = 0, bool PF = false,
          bool BULK = false, int MINB = 16, bool DET = false>
__global__ void __launch_bounds__(kThreads, MINB)
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
                        const T *__restrict__ vdict, int ndict,
                        int bulk_steps, const T *__restrict__ x_lower,
                        T *__restrict__ y_clear, long long *__restrict__ yq,
                        const double *__restrict__ qscale) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
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
  // where a contribution to y[col] goes: an L2 reduction (of the GPU below for
  // halo columns), or in deterministic mode an integer reduction into yq
  const double scale = DET ? *qscale : 0.0;
  auto emit = [&](int col, T v) {
    if (DET)
      tma::det_add(yq + col, (double)v, scale);
    else
      tma::y_add<HALO>(y, y_lower, row_begin, col, v);
  };
  const int p0 = slice_ptr[s], p1 = slice_ptr[s + 1];
  const T *sval = nullptr; // BULK: this warp's value block in shared memory
  uint64_t *bar = nullptr;
  if (BULK) {
    const int warp = threadIdx.x >> 5;
    bar = reinterpret_cast<uint64_t *>(dyn_smem) + warp;
    T *dst = reinterpret_cast<T *>(dyn_smem + 128) +
             (size_t)warp * bulk_steps * kSliceRows;
    sval = dst + lane;
    if (lane == 0) {
      tma::mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      if (p1 > p0) {
        uint64_t policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;"
                     : "=l"(policy));
        const uint32_t bytes =
            (uint32_t)(p1 - p0) * kSliceRows * (uint32_t)sizeof(T);
        tma::mbar_expect_tx(bar, bytes);
        tma::load_1d(dst, sell_val + (size_t)p0 * kSliceRows, bytes, bar,
                     policy);
      }
    }
    __syncwarp();
  }
  if (PF && VI == 0 && !BULK) {
    const char *blk = reinterpret_cast<const char *>(sell_val + (size_t)p0 * kSliceRows);
    const int bytes = (p1 - p0) * kSliceRows * (int)sizeof(T);
    for (int off = lane * 128; off < bytes; off += 32 * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + off));
  }
  const int tag = vrow_row[s * kSliceRows + lane];
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
    if (BULK)
      return sval[step * kSliceRows];
    return ld_stream(vp + (size_t)step * kSliceRows);
  };

  if (cptr & kSliceRegular) {
    // lane k holds the base column of step k
    const int base = ld_stream(cp);
    const int nb = __shfl_down_sync(0xffffffffu, base, 1);
    const unsigned cont =
        __ballot_sync(0xffffffffu, lane + 1 < w && nb == base + 1);
    if (BULK && w > 0)
      tma::mbar_wait(bar, 0);
    // The kernel is bound by instruction issue once the value stream is
    // compressed or prefetched (ncu, round 1: issue slots 72 %), so the chain
    // lengths are separate straight-line paths -- no predicated slots for the
    // entries a shorter chain does not have -- and the split of a rotated
    // product into S (lanes >= j) and E (the j lanes that wrapped) is two
    // predicated adds instead of adds + selects.
    for (int k = 0; k < w;) {
      const int cbase = __shfl_sync(0xffffffffu, base, k);
      const unsigned bits = cont >> k;
      const T *xp = x + cbase + lane;
      const bool halo = HALO && cbase < row_begin; // warp-uniform
      // x of the chain's window; a window that reaches below row_begin reads
      // those entries from the GPU below
      auto xv = [&](int j) -> T {
        if (HALO && halo)
          return tma::x_at<true>(x, x_lower, row_begin, cbase + lane + j);
        return xp[j];
      };
      T S, E = T(0);
      int L;
      if ((bits & 3u) == 3u) {
        L = 3;
        const T a0 = value_at(0), a1 = value_at(1), a2 = value_at(2);
        const T x0 = xv(0), x1 = xv(1), x2 = xv(2);
        acc += a0 * x0;
        acc += a1 * x1;
        acc += a2 * x2;
        S = a0 * xr;
        const T r1 = __shfl_sync(0xffffffffu, a1 * xr, (lane - 1) & 31);
        const T r2 = __shfl_sync(0xffffffffu, a2 * xr, (lane - 2) & 31);
        split_add(S, E, lane >= 1, r1);
        split_add(S, E, lane >= 2, r2);
      } else if (bits & 1u) {
        L = 2;
        const T a0 = value_at(0), a1 = value_at(1);
        const T x0 = xv(0), x1 = xv(1);
        acc += a0 * x0;
        acc += a1 * x1;
        S = a0 * xr;
        const T r1 = __shfl_sync(0xffffffffu, a1 * xr, (lane - 1) & 31);
        split_add(S, E, lane >= 1, r1);
      } else {
        L = 1;
        const T a0 = value_at(0);
        acc += a0 * xv(0);
        S = a0 * xr;
      }
      if (halo || DET) { // chain reaches the rows of the GPU below
        emit(cbase + lane, S);
        if (lane < L - 1)
          emit(cbase + kSliceRows + lane, E);
      } else {
        T *yp = y + cbase + lane;
        tma::red_add(yp, S);
        if (lane < L - 1)
          tma::red_add(yp + kSliceRows, E);
      }
      k += L;
      vp += L * kSliceRows; // running pointers: value_at(j) is one add away
      kp += L * kSliceRows;
      if (BULK)
        sval += L * kSliceRows;
    }
  } else {
    if (BULK && w > 0)
      tma::mbar_wait(bar, 0);
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
        xc[u] = c[u] >= 0 ? tma::x_at<HALO>(x, x_lower, row_begin, c[u]) : T(0);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (c[u] >= 0) {
          acc += a[u] * xc[u];
          emit(c[u], a[u] * xr);
        }
      }
      cp += kUnroll * kSliceRows;
      vp += kUnroll * kSliceRows;
      kp += kUnroll * kSliceRows;
      if (BULK)
        sval += kUnroll * kSliceRows;
    }
    for (; w > 0; --w) {
      const int c = ld_stream(cp);
      const T a = value_at(0);
      if (c >= 0) {
        acc += a * tma::x_at<HALO>(x, x_lower, row_begin, c);
        emit(c, a * xr);
      }
      cp += kSliceRows;
      vp += kSliceRows;
      kp += kSliceRows;
      if (BULK)
        sval += kSliceRows;
    }
  }
  if (active) {
    if (DET)
      tma::det_add(yq + row, (double)acc, scale);
    else
      tma::red_add(y + row, acc);
    // ping-pong result vectors: the row's owner clears the OTHER vector for the
    // next SpMV, which then needs no separate y initialisation
    if (y_clear && !(tag & kVrowCont))
      y_clear[row] = T(0);
  }
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
