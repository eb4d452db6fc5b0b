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

int cfs_host_scan_mmf_header(const char *image, size_t bytes,
                             cfs_host_mmf_header *out) {
  if (!image || !out)
    return 2;
  cfs::io::detail::MmfHeader h;
  cfs::io::detail::scan_matrix_market_header(image, bytes, h);
  out->nrows = h.nr_rows;
  out->ncols = h.nr_cols;
  out->declared = h.nr_declared;
  out->symmetric = h.symmetric ? 1 : 0;
  out->col_wise = h.col_wise ? 1 : 0;
  out->zero_based = h.zero_based ? 1 : 0;
  out->entries_offset = h.entries_offset;
  return 0;
}

void cfs_host_free_csr(cfs_host_csr *m) {
  if (!m || !m->handle)
    return;
  delete static_cast<CSRMatrix<int, double> *>(m->handle);
  m->handle = nullptr;
}

} // extern "C"
