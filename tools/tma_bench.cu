// tma_bench.cu -- micro-benchmark of cp.async.bulk (1-D TMA) on one GPU:
// how do op count and op size per tile affect the time of a 2-stage ring?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_bench tools/tma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int STAGES>
__global__ void ring(const char *src, size_t src_bytes, int ntiles, int nops, int op_bytes, int spread, int par, unsigned long long *sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int stage_bytes = nops * op_bytes;
  uint64_t *full = (uint64_t *)(smem + STAGES * stage_bytes);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // par == 0: thread 0 issues every op; 1: lane o of warp 0 issues op o;
  // 2: thread 32*(o%4) + o/4 issues op o (spread over the 4 warps)
  auto issue = [&](int tile, int stage) {
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[stage])), "r"(stage_bytes) : "memory");
    for (int o = 0; o < nops; ++o) {
      const int who = par == 0 ? 0 : par == 1 ? o : 32 * (o % 4) + o / 4;
      if (tid != who) continue;
      // spread: ops of one tile come from far-apart regions (like vals/cols/x)
      size_t off = ((size_t)tile * stage_bytes + (size_t)o * (spread ? (src_bytes / nops) : op_bytes)) % (src_bytes - op_bytes);
      off &= ~(size_t)127;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + stage * stage_bytes + o * op_bytes)), "l"(src + off), "r"(op_bytes), "r"(smem_u32(&full[stage])) : "memory");
    }
  };
  for (int s = 0; s < STAGES; ++s) {
    int t = blockIdx.x + s * gridDim.x;
    if (t < ntiles) issue(t, s);
  }
  unsigned long long acc = 0;
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int stage = it % STAGES;
    const uint32_t parity = (it / STAGES) & 1;
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&full[stage])), "r"(parity) : "memory");
    // touch the data lightly
    const unsigned long long *p = (const unsigned long long *)(smem + stage * stage_bytes);
    for (int i = tid; i < stage_bytes / 8; i += blockDim.x * 8) acc += p[i];
    __syncthreads();
    {
      int t = tile + STAGES * gridDim.x;
      if (t < ntiles) issue(t, stage);
    }
  }
  if (acc == 0x1234567) sink[0] = acc;
}

int main() {
  const size_t src_bytes = 2ull << 30;
  char *src; unsigned long long *sink;
  cudaMalloc(&src, src_bytes); cudaMalloc(&sink, 8);
  cudaMemset(src, 1, src_bytes);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t total = 1400ull << 20; // bytes per launch
  printf("%6s %8s %5s %6s %6s %10s %10s %10s\n", "nops", "opbytes", "ctas", "stages", "spread", "us", "GB/s", "us/tile/cta");
  for (int stages : {2})
  for (int spread : {1})
  for (int par : {0, 1, 2})
  for (int ctas : {1, 3})
  for (int nops : {1, 3, 9})
  for (int tile_kb : {8, 32}) {
    const int op_bytes = (tile_kb * 1024 / nops) & ~127;
    const int stage_bytes = nops * op_bytes;
    const int smem = stages * stage_bytes + 64;
    if ((size_t)smem * ctas > 220 * 1024) continue;
    const int ntiles = (int)(total / stage_bytes);
    auto kern = stages == 2 ? ring<2> : ring<4>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int grid = sms * ctas;
    kern<<<grid, 128, smem>>>(src, src_bytes, ntiles, nops, op_bytes, spread, par, sink);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) kern<<<grid, 128, smem>>>(src, src_bytes, ntiles, nops, op_bytes, spread, par, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    if (cudaGetLastError() != cudaSuccess) { printf("error\n"); return 1; }
    printf("par%d %6d %8d %5d %6d %6d %10.1f %10.1f %10.2f\n", par, nops, op_bytes, ctas, stages, spread, ms * 1e3, (double)ntiles * stage_bytes / ms / 1e6, ms * 1e3 / ((double)ntiles / grid));
  }
  return 0;
}
