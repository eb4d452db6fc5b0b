// allocator.hpp -- API of reference include/utils/allocator.hpp:11-12.
// Memory comes from cfs_cuda_host_alloc: 64-byte aligned and page-locked, so
// the x / y vectors of bench_spmv_mmf and test_spmv_mmf are DMA targets.
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

} // namespace memory
} // namespace util
} // namespace cfs

#endif
