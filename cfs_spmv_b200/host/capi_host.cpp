// capi_host.cpp -- include/cfs_host.h over the C++ classes.
#include "cfs.hpp"
#include "cfs_host.h"

using cfs::matrix::sparse::CSRMatrix;

extern "C" {

int cfs_host_load_mmf(const char *filename, int want_symmetric,
                      cfs_host_csr *out) {
  if (!filename || !out)
    return 2;
  CSRMatrix<int, double> *m = new CSRMatrix<int, double>(
      filename, cfs::util::Platform::cpu, want_symmetric != 0);
  out->nrows = m->nrows();
  out->ncols = m->ncols();
  out->nnz = m->nnz();
  out->symmetric = m->symmetric() ? 1 : 0;
  out->rowptr = m->rowptr();
  out->colind = m->colind();
  out->values = m->values();
  out->handle = m;
  return 0;
}

void cfs_host_free_csr(cfs_host_csr *m) {
  if (!m || !m->handle)
    return;
  delete static_cast<CSRMatrix<int, double> *>(m->handle);
  m->handle = nullptr;
}

} // extern "C"
