// det.cu -- deterministic mode of the symmetric SpMV (option "deterministic").
//
// The reference's kernel is bitwise reproducible for a fixed thread count: the
// colouring gives every y entry one writer per phase
// (csr_matrix.tpp:2988-3018). The default GPU kernels are not: the transposed
// term reaches y through L2 reductions whose ORDER varies from run to run, and
// floating-point addition is not associative. Colours do not carry over to
// 10^5 concurrent warps (the reference's graph has one vertex per 16 rows and
// one sequential sweep per partition), so reproducibility is restored where
// the order enters instead: every contribution is rounded once to a multiple of
// 2^-k and accumulated with INTEGER reductions, which commute exactly.
//
//   k = 61 - ceil(log2(|A|max * |x|max * longest full row)): no partial sum of
//   any y entry can leave 63 bits, and each term is off by <= 2^-(k+1) -- for
//   the 27-point matrices ~1e-17 absolute, below the rounding of a double sum.
//   |A|max and the longest row are properties of the matrix (max is order
//   independent); |x|max is one pass over x per SpMV.
//
// Steps of one SpMV: clear the 64-bit accumulators, |x|max, scale, the SpMV
// kernel (DET instantiation: same streams, same shuffle merging, integer REDs),
// y = acc * 2^-k. Cost and counters: RESULTS.md.
#include "common.cuh"

namespace cfsb {
namespace {

constexpr int kThreads = 256;

// max |v| as the bit pattern of a non-negative double (those order like u64)
template <typename T>
__global__ void absmax_kernel(long long n, const T *__restrict__ v,
                              unsigned long long *__restrict__ out) {
  double best = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double a = fabs((double)v[i]);
    best = a > best ? a : best; // NaN never wins: a NaN in x gives garbage anyway
  }
  for (int o = 16; o; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if ((threadIdx.x & 31) == 0 && best > 0.0)
    atomicMax(out, (unsigned long long)__double_as_longlong(best));
}

__global__ void scale_kernel(const unsigned long long *__restrict__ maxes,
                             int longest_row, double *__restrict__ scale) {
  const double amax = __longlong_as_double((long long)maxes[0]);
  const double xmax = __longlong_as_double((long long)maxes[1]);
  const double bound = amax * xmax * (double)(longest_row > 0 ? longest_row : 1);
  int k = 0;
  if (bound > 0.0 && bound < 1e300) {
    int e;
    frexp(bound, &e); // bound < 2^e
    k = 61 - e;
  }
  k = k > 1000 ? 1000 : (k < -1000 ? -1000 : k);
  scale[0] = ldexp(1.0, k);
  scale[1] = ldexp(1.0, -k);
}

template <typename T>
__global__ void from_fixed_kernel(long long n, const long long *__restrict__ acc,
                                  const double *__restrict__ scale,
                                  T *__restrict__ y) {
  const double inv = scale[1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = (T)((double)acc[i] * inv);
}

int grid_of(long long n) {
  const long long want = (n + kThreads - 1) / kThreads;
  return (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
}

} // namespace

int det_prepare(const cfs_matrix_s *m, const void *x_ext, cudaStream_t s) {
  const long long len = (long long)m->row_begin + m->nrows - m->halo_begin;
  if (!m->det_acc.p) {
    CFS_TRY(m->det_acc.alloc((size_t)len));
    CFS_TRY(m->det_max.alloc(2));
    CFS_TRY(m->det_scale.alloc(2));
    CFS_CUDA_TRY(cudaMemsetAsync(m->det_max.p, 0, 16, s));
  }
  if (!m->det_amax_done) { // |A|max over the stored values and the diagonal
    if (m->is_double) {
      absmax_kernel<double><<<grid_of(m->padded_entries), kThreads, 0, s>>>(
          m->padded_entries, (const double *)m->sell_val.p, m->det_max.p);
      absmax_kernel<double><<<grid_of(m->nrows), kThreads, 0, s>>>(
          m->nrows, (const double *)m->diagonal.p, m->det_max.p);
    } else {
      absmax_kernel<float><<<grid_of(m->padded_entries), kThreads, 0, s>>>(
          m->padded_entries, (const float *)m->sell_val.p, m->det_max.p);
      absmax_kernel<float><<<grid_of(m->nrows), kThreads, 0, s>>>(
          m->nrows, (const float *)m->diagonal.p, m->det_max.p);
    }
    m->det_amax_done = true;
  }
  CFS_CUDA_TRY(cudaMemsetAsync(m->det_acc.p, 0, (size_t)len * 8, s));
  CFS_CUDA_TRY(cudaMemsetAsync(m->det_max.p + 1, 0, 8, s));
  if (m->is_double)
    absmax_kernel<double><<<grid_of(len), kThreads, 0, s>>>(
        len, (const double *)x_ext, m->det_max.p + 1);
  else
    absmax_kernel<float><<<grid_of(len), kThreads, 0, s>>>(
        len, (const float *)x_ext, m->det_max.p + 1);
  scale_kernel<<<1, 1, 0, s>>>(m->det_max.p, m->max_row_nnz_full,
                               m->det_scale.p);
  CFS_CUDA_TRY(cudaGetLastError());
  return CFS_OK;
}

int det_finish(const cfs_matrix_s *m, void *y_ext, cudaStream_t s) {
  const long long len = (long long)m->row_begin + m->nrows - m->halo_begin;
  if (m->is_double)
    from_fixed_kernel<double><<<grid_of(len), kThreads, 0, s>>>(
        len, m->det_acc.p, m->det_scale.p, (double *)y_ext);
  else
    from_fixed_kernel<float><<<grid_of(len), kThreads, 0, s>>>(
        len, m->det_acc.p, m->det_scale.p, (float *)y_ext);
  CFS_CUDA_TRY(cudaGetLastError());
  return CFS_OK;
}

} // namespace cfsb
