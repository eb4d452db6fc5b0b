// allocator.cpp -- internal_alloc / internal_free over the C ABI
// (replaces reference src/allocator.cpp:8-43).
#include <cstdlib>
#include <iostream>

#include "cfs_cuda.h"
#include "utils/allocator.hpp"

namespace cfs {
namespace util {
namespace memory {

void *internal_alloc(size_t bytes, Platform) {
  void *p = cfs_cuda_host_alloc(bytes);
  if (!p) {
    // the reference's convention: message on stdout, exit(1)
    std::cout << "[ERROR]: cfs_cuda_host_alloc() failed!" << std::endl;
    exit(1);
  }
  return p;
}

void *internal_alloc_host(size_t bytes) {
  void *p = cfs_cuda_host_alloc_kind(bytes, CFS_ALLOC_PLAIN);
  if (!p) {
    std::cout << "[ERROR]: cfs_cuda_host_alloc_kind() failed!" << std::endl;
    exit(1);
  }
  return p;
}

void internal_free(void *pointer, Platform) { cfs_cuda_host_free(pointer); }

} // namespace memory
} // namespace util
} // namespace cfs
