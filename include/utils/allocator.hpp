// allocator.hpp -- API of reference include/utils/allocator.hpp:11-12.
// Memory comes from cfs_cuda_host_alloc: 64-byte aligned and, by default,
// UNIFIED memory, so that the x / y vectors of bench_spmv_mmf and test_spmv_mmf
// live in HBM while the GPU works on them and the host can still read and write
// them (CFS_GPU_ALLOC=managed|pinned|plain, see cfs_cuda.h).
#ifndef ALLOCATOR_HPP
#define ALLOCATOR_HPP

#include <cstddef>

#include "cfs_config.hpp"
#include "platform.hpp"

namespace cfs {
namespace util {
namespace memory {

void *internal_alloc(size_t bytes, Platform platform = Platform::cpu);
void internal_free(void *pointer, Platform platform = Platform::cpu);
// extension: plain host memory for arrays that are uploaded once (the CSR of
// the host loader); freed with internal_free like everything else
void *internal_alloc_host(size_t bytes);

} // namespace memory
} // namespace util
} // namespace cfs

#endif
