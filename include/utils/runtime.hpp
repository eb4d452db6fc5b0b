// runtime.hpp -- process-level knobs (API of reference
// include/utils/runtime.hpp:15-23).
#ifndef RUNTIME_HPP
#define RUNTIME_HPP

#include <cstddef>
#include <iostream>

#include "cfs_config.hpp"

namespace cfs {
namespace util {
namespace runtime {

const int MaxThreads = 96;

// CFS_NUM_THREADS, default 1 (reference src/runtime.cpp:10-21). On the GPU it
// no longer counts threads: it is the partition count P of the
// reference-compatible preprocessing metadata.
size_t get_num_threads();
// Pins the calling host thread (reference src/runtime.cpp:23-35).
void setaffinity_oncpu(unsigned int cpu);
// CFS_GPU_DEVICE, default 0: the (first) GPU this process drives.
int get_gpu_device();
// CFS_NUM_GPUS, default 1: GPUs a symmetric matrix is spread over -- what
// CFS_NUM_THREADS is to the reference's host threads (src/runtime.cpp:10-21):
// rows are cut into that many nnz-balanced blocks, one per GPU, from
// CFS_GPU_DEVICE upwards.
int get_num_gpus();

} // namespace runtime
} // namespace util
} // namespace cfs

#endif
