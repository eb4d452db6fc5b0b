// csr_matrix.hpp -- the one concrete matrix class (public API of reference
// include/matrix/csr_matrix.hpp:49-75). The object keeps the caller-visible
// full CSR on the host; tune() hands it to the GPU through the C ABI
// (include/cfs_cuda.h), where the lower triangle is extracted, partitioned,
// coloured and laid out for the sm_100a kernels.
// A matrix built from a Matrix Market file is ingested ON the GPU when one is
// present (cfs_cuda_matrix_create_from_mmf): its full CSR then lives in HBM
// and the host copy behind rowptr() / colind() / values() is only made when
// one of them is called.
// With CFS_NUM_GPUS > 1 a symmetric matrix is cut into nnz-balanced row blocks,
// one per GPU of the box (cfs_cuda_multi_*, the role of get_num_threads +
// partition_by_nnz in the reference), behind the same two calls.
#ifndef CSR_MATRIX_HPP
#define CSR_MATRIX_HPP

#include <cassert>
#include <cstring>
#include <iostream>
#include <string>

#include "cfs_config.hpp"
#include "io/mmf.hpp"
#include "matrix/sparse_matrix.hpp"
#include "utils/allocator.hpp"
#include "utils/platform.hpp"
#include "utils/runtime.hpp"

using namespace std;

struct cfs_matrix_s; // opaque handles of the C ABI
struct cfs_multi_s;

namespace cfs {

using namespace io;
using namespace util::memory;
using namespace util::runtime;

namespace matrix {
namespace sparse {

template <typename IndexT, typename ValueT>
class CSRMatrix : public SparseMatrix<IndexT, ValueT> {
public:
  CSRMatrix() = delete;
  CSRMatrix(const CSRMatrix &) = delete;
  CSRMatrix &operator=(const CSRMatrix &) = delete;
  // From a Matrix Market file; owns its arrays.
  CSRMatrix(const string &filename, Platform platform = Platform::cpu,
            bool symmetric = false, bool hybrid = false);
  // Wraps the caller's FULL CSR (0-based); never owns or frees it.
  CSRMatrix(IndexT *rowptr, IndexT *colind, ValueT *values, IndexT nrows,
            IndexT ncols, bool symmetric = false, bool hybrid = false,
            Platform platform = Platform::cpu);
  virtual ~CSRMatrix();

  virtual int nrows() const override { return nrows_; }
  virtual int ncols() const override { return ncols_; }
  virtual int nnz() const override { return nnz_; }
  virtual bool symmetric() const override { return symmetric_; }
  virtual size_t size() const override;
  virtual Platform platform() const override { return platform_; }
  virtual bool tune(Kernel k, Tuning t) override;
  virtual void dense_vector_multiply(ValueT *__restrict y,
                                     const ValueT *__restrict x) override;

  // The host CSR. Like the reference, a file-constructed symmetric matrix
  // releases it during tune(): these return nullptr afterwards.
  IndexT *rowptr() const {
    fetch_host_csr();
    return rowptr_;
  }
  IndexT *colind() const {
    fetch_host_csr();
    return colind_;
  }
  ValueT *values() const {
    fetch_host_csr();
    return values_;
  }

  // The C ABI handle behind this matrix (nullptr before tune()); lets a
  // device-side caller (a solver loop) use cfs_cuda_spmv_async directly.
  cfs_matrix_s *device_handle() const { return device_; }

  // Format::hyb: entries at least this far from the diagonal go to the
  // non-symmetric part (reference csr_matrix.hpp:92)
  static constexpr int HybBwThreshold = 10000;

private:
  Platform platform_;
  int nrows_, ncols_, nnz_;
  bool symmetric_, hybrid_, owns_data_, tuned_;
  int nparts_; // CFS_NUM_THREADS when the object was built
  mutable IndexT *rowptr_;
  mutable IndexT *colind_;
  mutable ValueT *values_;
  cfs_matrix_s *device_;
  cfs_multi_s *multi_; // CFS_NUM_GPUS > 1: row shards over the GPUs of the box
  int ngpus_;          // CFS_NUM_GPUS when the object was built
  mutable bool host_csr_pending_; // the full CSR is in HBM only (GPU ingest)

  void release_host_csr();
  void fetch_host_csr() const;
  bool ingest_on_gpu(const string &filename, bool symmetric);
};

} // namespace sparse
} // namespace matrix
} // namespace cfs

#endif
