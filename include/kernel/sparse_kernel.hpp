// sparse_kernel.hpp -- the SpDMV operator (API of reference
// include/kernel/sparse_kernel.hpp:17-27).
#ifndef SPARSE_KERNEL_HPP
#define SPARSE_KERNEL_HPP

#include <cassert>
#include <iostream>

#include "cfs_config.hpp"
#include "matrix/sparse_matrix.hpp"

namespace cfs {

using namespace matrix::sparse;

namespace kernel {
namespace sparse {

template <typename IndexType, typename ValueType> struct SpDMV {
public:
  SpDMV() = delete;
  // Runs A->tune(Kernel::SpDMV, t): all (GPU) preprocessing is paid here, which
  // is what bench_spmv_mmf reports as preproc(sec).
  SpDMV(SparseMatrix<IndexType, ValueType> *A, Tuning t = Tuning::Aggressive);
  // y = A * x, M == A->nrows(), N == A->ncols() (asserted).
  void operator()(ValueType *__restrict y, const int M,
                  const ValueType *__restrict x, const int N);

private:
  SparseMatrix<IndexType, ValueType> *A_; // not owned
};

} // namespace sparse
} // namespace kernel
} // namespace cfs

#endif
