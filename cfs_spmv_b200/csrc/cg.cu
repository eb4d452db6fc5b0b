// cg.cu -- conjugate gradients on the device: the caller-side loop a symmetric
// SpMV exists for. The reference stops at SpDMV
// (include/kernel/sparse_kernel.hpp:17-27): a solver built on it would move x
// and y across the host boundary every iteration (bench_spmv_mmf.cpp:162-167
// is exactly that loop). Here the vectors never leave HBM and an iteration is
// four kernels replayed as one CUDA graph:
//
//   q = A p, and p'q from the SAME kernel    sym_spmv_*_kernel<DOT>: for a
//                                            symmetric matrix p'Ap falls out of
//                                            the per-row accumulators
//   x += a p;  r -= a q;  r'r                cg_update_xr_kernel (a = rr / p'q)
//   p = r + b p;  q = 0                      cg_update_p_kernel  (b = rr' / rr)
//   scalars rotate, history, stop flag       cg_rotate_kernel
//
// No host synchronisation inside a batch of iterations: the step lengths are
// computed on the device from device scalars; once ||r|| <= tol * ||r0|| the
// update kernels become no-ops, so x is exactly the first converged iterate no
// matter how many iterations were enqueued after it.
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "spmv_reg.cuh"

namespace cfsb {
namespace {

constexpr int kThreads = 256;
constexpr int kDotWords = reg::kDotSlots * reg::kDotStride;

// device scalars of one solve
struct CgState {
  double rr;       // r'r of the current iterate
  double rr_next;  // accumulated by cg_update_xr_kernel
  double rr0;      // r'r at the start
  double tol2;     // (rel_tol)^2
  int iteration;   // completed iterations
  int done;        // 1: converged, 2: breakdown (p'Ap <= 0)
  int done_at;     // iteration count when `done` was raised
  int pad;
  double pq[kDotWords]; // partial sums of p'Ap written by the SpMV kernel
};

__device__ __forceinline__ double block_sum(double v, double *scratch) {
#pragma unroll
  for (int o = 16; o; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    scratch[warp] = v;
  __syncthreads();
  v = threadIdx.x < kThreads / 32 ? scratch[threadIdx.x] : 0.0;
  if (warp == 0) {
#pragma unroll
    for (int o = 16; o; o >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  __syncthreads();
  return v; // valid in warp 0
}

// r = b - q (q = A x0), p = r, q = 0, rr_next += r'r
template <typename T>
__global__ void __launch_bounds__(kThreads)
    cg_init_kernel(long long n, const T *__restrict__ b, T *__restrict__ q,
                   T *__restrict__ r, T *__restrict__ p, CgState *st) {
  __shared__ double scratch[kThreads / 32];
  double local = 0;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads) {
    const T ri = b[i] - q[i];
    r[i] = ri;
    p[i] = ri;
    q[i] = T(0);
    local += (double)ri * (double)ri;
  }
  local = block_sum(local, scratch);
  if (threadIdx.x == 0)
    atomicAdd(&st->rr_next, local);
}

__global__ void cg_init_scalars_kernel(CgState *st, double *history) {
  st->rr = st->rr0 = st->rr_next;
  st->rr_next = 0;
  st->iteration = 0;
  st->done = st->rr0 == 0.0 ? 1 : 0; // b = A x0 already
  st->done_at = 0;
  history[0] = st->rr0;
}

// alpha = rr / p'q;  x += alpha p;  r -= alpha q;  rr_next += r'r
template <typename T>
__global__ void __launch_bounds__(kThreads)
    cg_update_xr_kernel(long long n, const T *__restrict__ p,
                        const T *__restrict__ q, T *__restrict__ x,
                        T *__restrict__ r, CgState *st) {
  __shared__ double scratch[kThreads / 32];
  __shared__ double s_alpha;
  if (st->done)
    return;
  // p'q: sum of the partial sums the SpMV kernel left
  double part = 0;
  for (int k = threadIdx.x; k < reg::kDotSlots; k += kThreads)
    part += st->pq[k * reg::kDotStride];
  part = block_sum(part, scratch);
  if (threadIdx.x == 0)
    s_alpha = part > 0.0 ? st->rr / part : 0.0; // <= 0: not positive definite
  __syncthreads();
  const double alpha = s_alpha;
  double local = 0;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads) {
    x[i] = (T)((double)x[i] + alpha * (double)p[i]);
    const T ri = (T)((double)r[i] - alpha * (double)q[i]);
    r[i] = ri;
    local += (double)ri * (double)ri;
  }
  local = block_sum(local, scratch);
  if (threadIdx.x == 0)
    atomicAdd(&st->rr_next, local);
}

// beta = rr_next / rr;  p = r + beta p;  q = 0 (the SpMV reduces into it)
template <typename T>
__global__ void __launch_bounds__(kThreads)
    cg_update_p_kernel(long long n, const T *__restrict__ r, T *__restrict__ p,
                       T *__restrict__ q, const CgState *st) {
  const bool frozen = st->done != 0;
  const double beta = frozen ? 0.0 : st->rr_next / st->rr;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads) {
    if (!frozen)
      p[i] = (T)((double)r[i] + beta * (double)p[i]);
    q[i] = T(0);
  }
}

__global__ void cg_rotate_kernel(CgState *st, double *history,
                                 int history_len) {
  const int t = threadIdx.x;
  double part = 0;
  for (int k = t; k < reg::kDotSlots; k += 32)
    part += st->pq[k * reg::kDotStride];
#pragma unroll
  for (int o = 16; o; o >>= 1)
    part += __shfl_xor_sync(0xffffffffu, part, o);
  __syncwarp();
  for (int k = t; k < reg::kDotSlots; k += 32)
    st->pq[k * reg::kDotStride] = 0.0;
  if (t != 0 || st->done)
    return;
  st->iteration += 1;
  if (!(part > 0.0)) { // p'Ap <= 0 (or NaN): A is not positive definite
    st->done = 2;
    st->done_at = st->iteration;
    st->rr_next = 0;
    return;
  }
  st->rr = st->rr_next;
  st->rr_next = 0;
  if (st->iteration < history_len)
    history[st->iteration] = st->rr;
  if (st->rr <= st->tol2 * st->rr0) {
    st->done = 1;
    st->done_at = st->iteration;
  }
}

// ---- the same updates for a row shard: the scalars are global sums the
// caller has reduced across the ranks (scal = {r'r, p'Ap, next r'r})
template <typename T>
__global__ void __launch_bounds__(kThreads)
    shard_update_xr_kernel(long long n, const double *__restrict__ scal,
                           const T *__restrict__ p, const T *__restrict__ q,
                           T *__restrict__ x, T *__restrict__ r,
                           double *__restrict__ rr_next) {
  __shared__ double scratch[kThreads / 32];
  const double pq = scal[1];
  const double alpha = pq > 0.0 ? scal[0] / pq : 0.0;
  double local = 0;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads) {
    x[i] = (T)((double)x[i] + alpha * (double)p[i]);
    const T ri = (T)((double)r[i] - alpha * (double)q[i]);
    r[i] = ri;
    local += (double)ri * (double)ri;
  }
  local = block_sum(local, scratch);
  if (threadIdx.x == 0)
    atomicAdd(rr_next, local);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    shard_update_p_kernel(long long n, const double *__restrict__ scal,
                          const T *__restrict__ r, T *__restrict__ p) {
  const double beta = scal[0] > 0.0 ? scal[2] / scal[0] : 0.0;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads)
    p[i] = (T)((double)r[i] + beta * (double)p[i]);
}

// *out += sum of the partial sums the SpMV kernel left; the slots are cleared
__global__ void dot_collect_kernel(double *__restrict__ slots,
                                   double *__restrict__ out) {
  const int t = threadIdx.x;
  double part = 0;
  for (int k = t; k < reg::kDotSlots; k += 32) {
    part += slots[k * reg::kDotStride];
    slots[k * reg::kDotStride] = 0.0;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1)
    part += __shfl_xor_sync(0xffffffffu, part, o);
  if (t == 0)
    *out += part;
}

int grid_for(long long n) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0)
      sms = 148;
  }
  const long long want = (n + kThreads - 1) / kThreads;
  const long long cap = (long long)sms * 8; // 8 CTAs of 256 threads per SM
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

template <typename T>
int solve(cfs_matrix_s *m, T *x, const T *b, int max_iters, double rel_tol,
          int batch, cfs_cg_result *res, double *history_out,
          int history_capacity) {
  const long long n = m->nrows;
  cudaStream_t s = m->stream;
  DevArray<T> r, p, q;
  DevArray<CgState> st;
  DevArray<double> history;
  CFS_TRY(r.alloc((size_t)n));
  CFS_TRY(p.alloc((size_t)n));
  CFS_TRY(q.alloc((size_t)n));
  CFS_TRY(st.alloc(1));
  const int history_len = max_iters + 1;
  CFS_TRY(history.alloc((size_t)history_len));
  CFS_CUDA_TRY(cudaMemsetAsync(st.p, 0, sizeof(CgState), s));
  CFS_CUDA_TRY(cudaMemsetAsync(history.p, 0, (size_t)history_len * 8, s));
  double *pq = (double *)((char *)st.p + offsetof(CgState, pq));
  const int grid = grid_for(n);

  // released on every way out, the early error returns included
  struct Events {
    cudaEvent_t a = nullptr, b = nullptr;
    ~Events() {
      if (a)
        cudaEventDestroy(a);
      if (b)
        cudaEventDestroy(b);
    }
  } events;
  CFS_CUDA_TRY(cudaEventCreate(&events.a));
  CFS_CUDA_TRY(cudaEventCreate(&events.b));
  const cudaEvent_t e0 = events.a, e1 = events.b;
  CFS_CUDA_TRY(cudaEventRecord(e0, s));

  // r = b - A x0
  CFS_TRY(launch_sym_spmv(m, q.p, x, s));
  cg_init_kernel<T><<<grid, kThreads, 0, s>>>(n, b, q.p, r.p, p.p, st.p);
  cg_init_scalars_kernel<<<1, 1, 0, s>>>(st.p, history.p);
  {
    const double tol2 = rel_tol * rel_tol;
    CFS_CUDA_TRY(cudaMemcpyAsync((char *)st.p + offsetof(CgState, tol2), &tol2,
                                 8, cudaMemcpyHostToDevice, s));
  }
  CFS_CUDA_TRY(cudaGetLastError());

  // one iteration, captured once
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  // (in deterministic mode launch_sym_spmv refuses x'Ax: say so before a
  // capture is open)
  if (g_options.deterministic) {
    set_error("cfs_cuda_cg_solve: not available in deterministic mode (its "
              "SpMV also returns p'Ap)");
    return CFS_ERR_STATE;
  }
  CFS_CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int status = launch_sym_spmv(m, q.p, p.p, s, nullptr, nullptr, nullptr, true,
                               0, -1, pq);
  cg_update_xr_kernel<T><<<grid, kThreads, 0, s>>>(n, p.p, q.p, x, r.p, st.p);
  cg_update_p_kernel<T><<<grid, kThreads, 0, s>>>(n, r.p, p.p, q.p, st.p);
  cg_rotate_kernel<<<1, 32, 0, s>>>(st.p, history.p, history_len);
  {
    const cudaError_t e = cudaStreamEndCapture(s, &graph);
    if (status == CFS_OK && e != cudaSuccess)
      status = cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
  }
  if (status == CFS_OK &&
      cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess)
    status = cuda_fail(cudaGetLastError(), "cudaGraphInstantiate", __FILE__,
                       __LINE__);
  if (graph)
    cudaGraphDestroy(graph);

  CgState host;
  memset(&host, 0, sizeof(host));
  int enqueued = 0;
  while (status == CFS_OK && enqueued < max_iters) {
    const int todo = max_iters - enqueued < batch ? max_iters - enqueued : batch;
    for (int k = 0; k < todo && status == CFS_OK; ++k)
      if (cudaGraphLaunch(exec, s) != cudaSuccess)
        status = cuda_fail(cudaGetLastError(), "cudaGraphLaunch", __FILE__,
                           __LINE__);
    enqueued += todo;
    // everything before pq is what the host looks at
    if (status == CFS_OK &&
        (cudaMemcpyAsync(&host, st.p, offsetof(CgState, pq),
                         cudaMemcpyDeviceToHost, s) != cudaSuccess ||
         cudaStreamSynchronize(s) != cudaSuccess))
      status = cuda_fail(cudaGetLastError(), "read solver state", __FILE__,
                         __LINE__);
    if (host.done)
      break;
  }
  if (status == CFS_OK && max_iters == 0) {
    if (cudaMemcpyAsync(&host, st.p, offsetof(CgState, pq),
                        cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
      status = cuda_fail(cudaGetLastError(), "read solver state", __FILE__,
                         __LINE__);
  }
  cudaEventRecord(e1, s);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  if (exec)
    cudaGraphExecDestroy(exec);
  if (status != CFS_OK)
    return status;
  res->iterations = host.done ? host.done_at : host.iteration;
  res->converged = host.done == 1;
  res->breakdown = host.done == 2;
  res->executed = enqueued;
  res->initial_residual_norm = sqrt(host.rr0);
  res->residual_norm = sqrt(host.rr);
  res->ms_total = ms;
  if (history_out && history_capacity > 0) {
    const int nh = res->iterations + 1 < history_capacity
                       ? res->iterations + 1
                       : history_capacity;
    CFS_CUDA_TRY(cudaMemcpy(history_out, history.p, (size_t)nh * 8,
                            cudaMemcpyDeviceToHost));
    for (int k = 0; k < nh; ++k)
      history_out[k] = sqrt(history_out[k]);
  }
  return CFS_OK;
}

bool on_device(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

} // namespace
} // namespace cfsb

using namespace cfsb;

extern "C" int cfs_cuda_cg_solve(cfs_mat_t m, void *x, const void *b,
                                 int max_iters, double rel_tol,
                                 cfs_cg_result *result, double *history,
                                 int history_capacity) {
  if (!m || !x || !b || !result || max_iters < 0 || !(rel_tol >= 0.0)) {
    set_error("cfs_cuda_cg_solve: bad arguments");
    return CFS_ERR_INVALID;
  }
  if (!m->tuned || !m->symmetric || m->sharded || m->nrows != m->ncols) {
    set_error("cfs_cuda_cg_solve: needs a tuned, unsharded symmetric matrix");
    return CFS_ERR_STATE;
  }
  CFS_CUDA_TRY(cudaSetDevice(m->device));
  memset(result, 0, sizeof(*result));
  const size_t bytes = (size_t)m->nrows * m->vsize();
  // host vectors are staged once; the iterations run on device copies
  DevArray<char> xd, bd;
  void *xp = x;
  const void *bp = b;
  const bool x_host = !on_device(x), b_host = !on_device(b);
  if (x_host) {
    CFS_TRY(xd.alloc(bytes));
    CFS_CUDA_TRY(cudaMemcpyAsync(xd.p, x, bytes, cudaMemcpyHostToDevice,
                                 m->stream));
    xp = xd.p;
  }
  if (b_host) {
    CFS_TRY(bd.alloc(bytes));
    CFS_CUDA_TRY(cudaMemcpyAsync(bd.p, b, bytes, cudaMemcpyHostToDevice,
                                 m->stream));
    bp = bd.p;
  }
  const int batch = g_options.cg_batch;
  const int status =
      m->is_double
          ? solve<double>(m, (double *)xp, (const double *)bp, max_iters,
                          rel_tol, batch, result, history, history_capacity)
          : solve<float>(m, (float *)xp, (const float *)bp, max_iters, rel_tol,
                         batch, result, history, history_capacity);
  if (status != CFS_OK)
    return status;
  if (x_host)
    CFS_CUDA_TRY(cudaMemcpy(x, xd.p, bytes, cudaMemcpyDeviceToHost));
  return CFS_OK;
}


extern "C" int cfs_cuda_spmv_halo_dot_async(cfs_mat_t m, void *y_dev,
                                            const void *x_dev,
                                            void *y_lower_base, int y_is_zero,
                                            double *dot_dev, void *stream) {
  if (!m || !y_dev || !x_dev || !dot_dev)
    return CFS_ERR_INVALID;
  if (!m->tuned || !m->symmetric) {
    set_error("cfs_cuda_spmv_halo_dot_async: needs a tuned symmetric matrix");
    return CFS_ERR_STATE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (!m->dot_slots.p) {
    CFS_TRY(m->dot_slots.alloc(kDotWords));
    CFS_CUDA_TRY(cudaMemsetAsync(m->dot_slots.p, 0, kDotWords * 8, s));
  }
  CFS_TRY(launch_sym_spmv(m, y_dev, x_dev, s, nullptr, nullptr, y_lower_base,
                          y_is_zero != 0, 0, -1, m->dot_slots.p));
  dot_collect_kernel<<<1, 32, 0, s>>>(m->dot_slots.p, dot_dev);
  CFS_CUDA_TRY(cudaGetLastError());
  return CFS_OK;
}

extern "C" int cfs_cuda_cg_update_xr(int64_t n, int is_double,
                                     const double *scal, const void *p,
                                     const void *q, void *x, void *r,
                                     double *rr_next_dev, void *stream) {
  if (n < 0 || !scal || !p || !q || !x || !r || !rr_next_dev)
    return CFS_ERR_INVALID;
  CFS_TRY(require_device());
  if (n == 0)
    return CFS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = grid_for(n);
  if (is_double)
    shard_update_xr_kernel<double><<<grid, kThreads, 0, s>>>(
        n, scal, (const double *)p, (const double *)q, (double *)x,
        (double *)r, rr_next_dev);
  else
    shard_update_xr_kernel<float><<<grid, kThreads, 0, s>>>(
        n, scal, (const float *)p, (const float *)q, (float *)x, (float *)r,
        rr_next_dev);
  CFS_CUDA_TRY(cudaGetLastError());
  return CFS_OK;
}

extern "C" int cfs_cuda_cg_update_p(int64_t n, int is_double,
                                    const double *scal, const void *r, void *p,
                                    void *stream) {
  if (n < 0 || !scal || !r || !p)
    return CFS_ERR_INVALID;
  CFS_TRY(require_device());
  if (n == 0)
    return CFS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = grid_for(n);
  if (is_double)
    shard_update_p_kernel<double><<<grid, kThreads, 0, s>>>(
        n, scal, (const double *)r, (double *)p);
  else
    shard_update_p_kernel<float><<<grid, kThreads, 0, s>>>(
        n, scal, (const float *)r, (float *)p);
  CFS_CUDA_TRY(cudaGetLastError());
  return CFS_OK;
}
