// runtime.cpp -- environment knobs (replaces reference src/runtime.cpp:10-35).
#include <sched.h>
#include <stdlib.h>

#include "utils/runtime.hpp"

namespace cfs {
namespace util {
namespace runtime {

static int env_int(const char *name, int fallback) {
  const char *text = getenv(name);
  if (!text)
    return fallback;
  const int v = atoi(text);
  return v < 0 ? fallback : v;
}

size_t get_num_threads() { return (size_t)env_int("CFS_NUM_THREADS", 1); }

int get_gpu_device() { return env_int("CFS_GPU_DEVICE", 0); }

int get_num_gpus() {
  const int n = env_int("CFS_NUM_GPUS", 1);
  return n < 1 ? 1 : n;
}

void setaffinity_oncpu(unsigned int cpu) {
  cpu_set_t mask;
  CPU_ZERO(&mask);
  CPU_SET(cpu, &mask);
  if (sched_setaffinity(0, sizeof(mask), &mask) != 0) {
    std::cout << "sched_setaffinity() failed" << std::endl;
    exit(1);
  }
}

} // namespace runtime
} // namespace util
} // namespace cfs
