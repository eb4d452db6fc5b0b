// managed_probe.cu -- what unified (managed) memory costs the drop-in vectors
// (development aid; decides the policy of cfs_cuda_host_alloc / cfs_cuda_spmv).
// The reference's bench / test obtain x and y from internal_alloc, fill x on the
// host, call y = A x in a loop and (the test) read y on the host afterwards.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/managed_probe tools/managed_probe.cu
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      printf("%s: %s\n", #x, cudaGetErrorString(e));                           \
      return 1;                                                                \
    }                                                                          \
  } while (0)

__global__ void axpy_kernel(double *__restrict__ y, const double *__restrict__ x,
                            size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    y[i] = 2.0 * x[i];
}

static double now_us() {
  return std::chrono::duration<double, std::micro>(
             std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

int main(int argc, char **argv) {
  const int advise = argc > 1 ? atoi(argv[1]) : 0; // 1: preferred location GPU
  for (size_t N : {(size_t)1000000, (size_t)8000000}) {
    const size_t B = N * 8;
    double *x, *y;
    CK(cudaMallocManaged(&x, B));
    CK(cudaMallocManaged(&y, B));
    if (advise) {
      CK(cudaMemAdvise(x, B, cudaMemAdviseSetPreferredLocation, 0));
      CK(cudaMemAdvise(y, B, cudaMemAdviseSetPreferredLocation, 0));
    }
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    double t0 = now_us();
    for (size_t i = 0; i < N; ++i) {
      x[i] = 0.25 + (double)(i & 1023);
      y[i] = 0;
    }
    printf("N=%zu advise=%d: host first touch of x,y %.0f us\n", N, advise,
           now_us() - t0);
    auto run = [&](const char *what) {
      const double a = now_us();
      axpy_kernel<<<148 * 8, 256, 0, s>>>(y, x, N);
      cudaStreamSynchronize(s);
      printf("  %-52s %9.1f us\n", what, now_us() - a);
      return 0;
    };
    run("kernel, x/y on the host (fault-driven migration)");
    run("kernel again (resident)");
    run("kernel again (resident)");
    // prefetch when already resident: API cost + stream cost
    double a = now_us();
    for (int k = 0; k < 100; ++k) {
      CK(cudaMemPrefetchAsync(x, B, 0, s));
      CK(cudaMemPrefetchAsync(y, B, 0, s));
    }
    const double api = (now_us() - a) / 100;
    CK(cudaStreamSynchronize(s));
    printf("  2 prefetches of resident ranges: API %.1f us per pair, drained "
           "after %.1f us per pair\n",
           api, (now_us() - a) / 100);
    a = now_us();
    for (int k = 0; k < 100; ++k) {
      CK(cudaMemPrefetchAsync(x, B, 0, s));
      CK(cudaMemPrefetchAsync(y, B, 0, s));
      axpy_kernel<<<148 * 8, 256, 0, s>>>(y, x, N);
      CK(cudaStreamSynchronize(s));
    }
    printf("  prefetch x, prefetch y, kernel, sync (resident): %.1f us per "
           "call\n",
           (now_us() - a) / 100);
    a = now_us();
    for (int k = 0; k < 100; ++k) {
      axpy_kernel<<<148 * 8, 256, 0, s>>>(y, x, N);
      CK(cudaStreamSynchronize(s));
    }
    printf("  kernel, sync (resident): %.1f us per call\n",
           (now_us() - a) / 100);
    // pointer classification cost
    a = now_us();
    for (int k = 0; k < 1000; ++k) {
      cudaPointerAttributes at;
      cudaPointerGetAttributes(&at, x);
      cudaPointerGetAttributes(&at, y);
    }
    printf("  2 x cudaPointerGetAttributes: %.2f us\n", (now_us() - a) / 1000);
    // the host reads y (test_spmv_mmf.cpp:94-104)
    a = now_us();
    double sum = 0;
    for (size_t i = 0; i < N; ++i)
      sum += y[i];
    printf("  host reads y after the kernel (CPU faults): %.0f us (sum %g)\n",
           now_us() - a, sum);
    run("kernel after the host READ y (y back by GPU faults)");
    // the host rewrites x between calls
    a = now_us();
    for (size_t i = 0; i < N; ++i)
      x[i] += 1.0;
    printf("  host rewrites x (CPU faults): %.0f us\n", now_us() - a);
    run("kernel after a host write of x, no prefetch");
    a = now_us();
    for (size_t i = 0; i < N; ++i)
      x[i] += 1.0;
    printf("  host rewrites x (CPU faults): %.0f us\n", now_us() - a);
    a = now_us();
    CK(cudaMemPrefetchAsync(x, B, 0, s));
    axpy_kernel<<<148 * 8, 256, 0, s>>>(y, x, N);
    CK(cudaStreamSynchronize(s));
    printf("  prefetch + kernel after a host write of x: %.1f us\n",
           now_us() - a);
    // prefetch y to the host instead of faulting
    a = now_us();
    CK(cudaMemPrefetchAsync(y, B, cudaCpuDeviceId, s));
    CK(cudaStreamSynchronize(s));
    sum = 0;
    for (size_t i = 0; i < N; ++i)
      sum += y[i];
    printf("  prefetch y to the host + host read: %.0f us (sum %g)\n",
           now_us() - a, sum);
    // explicit copies for comparison (what the pinned path pays)
    double *hx;
    CK(cudaHostAlloc(&hx, B, 0));
    double *dx;
    CK(cudaMalloc(&dx, B));
    a = now_us();
    CK(cudaMemcpyAsync(dx, hx, B, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    printf("  pinned H2D of the same size: %.1f us\n", now_us() - a);
    cudaFree(dx);
    cudaFreeHost(hx);
    cudaFree(x);
    cudaFree(y);
    cudaStreamDestroy(s);
  }
  return 0;
}
