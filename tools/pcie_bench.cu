// pcie_bench.cu -- what the host link gives a pipelined host-vector SpMV
// (development aid). Pinned 64 MB vectors; copy engines vs SM-issued copies,
// whole vs chunked, one direction vs both.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pcie_bench tools/pcie_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      printf("%s: %s\n", #x, cudaGetErrorString(e));                           \
      return 1;                                                                \
    }                                                                          \
  } while (0)

__global__ void copy_kernel(double2 *__restrict__ dst,
                            const double2 *__restrict__ src, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

int main() {
  const size_t N = 8000000, B = N * 8;
  double *hx, *hy, *dx, *dy;
  CK(cudaHostAlloc(&hx, B, cudaHostAllocPortable));
  CK(cudaHostAlloc(&hy, B, cudaHostAllocPortable));
  CK(cudaMalloc(&dx, B));
  CK(cudaMalloc(&dy, B));
  for (size_t i = 0; i < N; ++i)
    hx[i] = (double)i;
  cudaStream_t s1, s2;
  CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  cudaEvent_t e0, e1, e2;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventCreate(&e2));
  const int iters = 10;
  auto report = [&](const char *name, float ms) {
    printf("%-58s %.3f ms  (%.1f GB/s per direction)\n", name, ms / iters,
           B / 1e6 / (ms / iters));
  };
  for (int K : {1, 8, 16}) {
    const size_t c = B / K;
    for (int mode = 0; mode < 5; ++mode) {
      // 0: H2D only (CE); 1: D2H only (CE); 2: both CE; 3: H2D CE + D2H by SMs;
      // 4: both by SMs
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s1));
        CK(cudaStreamWaitEvent(s2, e0, 0));
        for (int it = 0; it < iters; ++it)
          for (int k = 0; k < K; ++k) {
            char *hxp = (char *)hx + k * c, *dxp = (char *)dx + k * c;
            char *hyp = (char *)hy + k * c, *dyp = (char *)dy + k * c;
            if (mode == 0 || mode == 2 || mode == 3)
              CK(cudaMemcpyAsync(dxp, hxp, c, cudaMemcpyHostToDevice, s1));
            if (mode == 4)
              copy_kernel<<<64, 256, 0, s1>>>((double2 *)dxp, (double2 *)hxp,
                                               c / 16);
            if (mode == 1 || mode == 2)
              CK(cudaMemcpyAsync(hyp, dyp, c, cudaMemcpyDeviceToHost, s2));
            if (mode == 3 || mode == 4)
              copy_kernel<<<64, 256, 0, s2>>>((double2 *)hyp, (double2 *)dyp,
                                               c / 16);
          }
        CK(cudaEventRecord(e2, s2));
        CK(cudaStreamWaitEvent(s1, e2, 0));
        CK(cudaEventRecord(e1, s1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      char name[128];
      const char *what[] = {"H2D (copy engine)", "D2H (copy engine)",
                            "H2D + D2H (copy engines)",
                            "H2D copy engine + D2H SM stores",
                            "H2D SM loads + D2H SM stores"};
      snprintf(name, sizeof(name), "%s, %d chunk(s)", what[mode], K);
      report(name, ms);
    }
  }
  // SM copy grid-size sensitivity, D2H only
  for (int grid : {16, 32, 64, 148, 296}) {
    float ms = 0;
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, s2));
    for (int it = 0; it < iters; ++it)
      copy_kernel<<<grid, 256, 0, s2>>>((double2 *)hy, (double2 *)dy, B / 16);
    CK(cudaEventRecord(e1, s2));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    char name[128];
    snprintf(name, sizeof(name), "D2H SM stores alone, grid %d", grid);
    report(name, ms);
  }
  return 0;
}
