// cg_solver.hpp -- conjugate gradients for symmetric positive definite A, the
// caller-side loop of SpDMV kept on the GPU. Not part
// of the reference's API (it stops at SpDMV, include/kernel/sparse_kernel.hpp:
// 17-27); it follows the same conventions: a functor built on a SparseMatrix
// the caller owns, construction runs tune(), host or device vectors of the
// matrix precision, fatal problems print and exit(1).
#ifndef CG_SOLVER_HPP
#define CG_SOLVER_HPP

#include "cfs_config.hpp"
#include "matrix/sparse_matrix.hpp"

namespace cfs {

using namespace matrix::sparse;

namespace kernel {
namespace sparse {

template <typename IndexType, typename ValueType> struct ConjugateGradient {
public:
  ConjugateGradient() = delete;
  // A must be a symmetric format (Format::sss); runs A->tune(Kernel::SpDMV, t).
  ConjugateGradient(SparseMatrix<IndexType, ValueType> *A,
                    Tuning t = Tuning::Aggressive);
  // Solves A x = b. x: initial guess on entry, solution on return; N ==
  // A->nrows() (asserted). Stops when ||r_k|| <= rel_tol * ||r_0|| or after
  // max_iters iterations; returns the number of iterations taken.
  int operator()(ValueType *__restrict x, const ValueType *__restrict b,
                 const int N, const int max_iters, const double rel_tol);

  bool converged() const { return converged_; }
  bool breakdown() const { return breakdown_; } // p'Ap <= 0: A is not SPD
  double residual_norm() const { return residual_norm_; }
  double initial_residual_norm() const { return initial_residual_norm_; }
  double solve_milliseconds() const { return ms_; } // device time

private:
  SparseMatrix<IndexType, ValueType> *A_; // not owned
  bool converged_, breakdown_;
  double residual_norm_, initial_residual_norm_, ms_;
};

} // namespace sparse
} // namespace kernel
} // namespace cfs

#endif
