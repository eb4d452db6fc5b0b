// cfs.hpp -- umbrella header of the B200-native library. Same header names,
// namespaces, classes and signatures as the reference's include/cfs.hpp
// (reference include/cfs.hpp:4-10), so bench_spmv_mmf.cpp / test_spmv_mmf.cpp
// compile unchanged; everything behind the classes is a call into the C ABI of
// include/cfs_cuda.h.
#ifndef CFS_HPP
#define CFS_HPP

#include "cfs_config.hpp"
#include "utils/platform.hpp"
#include "utils/runtime.hpp"
#include "utils/allocator.hpp"
#include "io/mmf.hpp"
#include "matrix/sparse_matrix.hpp"
#include "matrix/csr_matrix.hpp"
#include "kernel/sparse_kernel.hpp"
#include "kernel/cg_solver.hpp"

#endif
