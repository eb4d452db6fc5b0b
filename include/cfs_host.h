/*
 * cfs_host.h -- C entry points of the HOST side of the library (libsparse.so):
 * the Matrix Market loader + CSR construction that precede the GPU path
 * (reference include/io/mmf.hpp:179-343, include/matrix/csr_matrix.tpp:9-111).
 * Pure host code: usable without a GPU. Fatal file problems follow the
 * reference's convention (message on stdout, exit(1)).
 */
#ifndef CFS_HOST_H
#define CFS_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cfs_host_csr {
  int32_t nrows, ncols, nnz; /* nnz: expanded (full) count, CSRMatrix::nnz()  */
  int32_t symmetric;         /* CSRMatrix::symmetric() after the fallback     */
  const int32_t *rowptr;     /* 0-based full CSR, owned by the handle         */
  const int32_t *colind;
  const double *values;
  void *handle;
} cfs_host_csr;

/* CSRMatrix<int,double>(filename, Platform::cpu, want_symmetric). */
int cfs_host_load_mmf(const char *filename, int want_symmetric,
                      cfs_host_csr *out);
void cfs_host_free_csr(cfs_host_csr *m);

/* Banner, comment and size lines of a Matrix Market file image
 * (include/io/mmf.hpp:203-273 of the reference: ParseMmfHeaderLine,
 * ParseMmfSizeLine); entries_offset is where the entry lines begin -- what
 * cfs_cuda_matrix_create_from_mmf (cfs_cuda.h) takes over from. Fatal header
 * problems print the reference's message and exit(1). */
typedef struct cfs_host_mmf_header {
  int64_t nrows, ncols, declared;
  int32_t symmetric, col_wise, zero_based;
  uint64_t entries_offset;
} cfs_host_mmf_header;
int cfs_host_scan_mmf_header(const char *image, size_t bytes,
                             cfs_host_mmf_header *out);

#ifdef __cplusplus
}
#endif
#endif
