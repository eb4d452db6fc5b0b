// sparse_matrix.hpp -- the abstract matrix interface and its factory
// (API of reference include/matrix/sparse_matrix.hpp:23-41).
#ifndef SPARSE_MATRIX_HPP
#define SPARSE_MATRIX_HPP

#include <cmath>
#include <iostream>
#include <random>
#include <string>

#include "cfs_config.hpp"
#include "utils/platform.hpp"

// Consumers of the reference rely on this leaking out of the headers
// (reference sparse_matrix.hpp:10, csr_matrix.hpp:34).
using namespace std;

namespace cfs {

using namespace util;

namespace matrix {
namespace sparse {

template <typename IndexT, typename ValueT> class SparseMatrix {
public:
  virtual ~SparseMatrix() {}
  virtual int nrows() const = 0;
  virtual int ncols() const = 0;
  virtual int nnz() const = 0; // expanded (full-matrix) count
  virtual bool symmetric() const = 0;
  virtual size_t size() const = 0; // footprint of the format in bytes
  virtual Platform platform() const = 0;
  // all preprocessing happens here (GPU); always returns true
  virtual bool tune(Kernel k, Tuning t = Tuning::Aggressive) = 0;
  // y = A * x; y is fully overwritten; synchronous
  virtual void dense_vector_multiply(ValueT *__restrict y,
                                     const ValueT *__restrict x) = 0;

  // Loads a Matrix Market file. Format::sss / Format::hyb keep only the lower
  // triangle when the file is symmetric; anything else is plain CSR.
  static SparseMatrix<IndexT, ValueT> *
  create(const string &filename, Format format = Format::csr,
         Platform platform = Platform::cpu);
};

} // namespace sparse
} // namespace matrix
} // namespace cfs

#endif
