// load_timer.cpp -- times SparseMatrix::create(file, format), i.e. the Matrix
// Market loader + CSR construction that precede the SpMV path (reference
// include/io/mmf.hpp:179-343, include/matrix/csr_matrix.tpp:9-111), and prints
// a checksum of the resulting full CSR. Uses the public API only, so the SAME
// source builds against the unmodified reference (oracle/_ref/load_timer) and
// against this repo's library (build/dropin/load_timer): equal checksums =
// equal CSR, bit for bit.
//
//   load_timer <mmf_file> <format 0:CSR 1:SSS> [warm]
// warm: one internal_alloc / internal_free before the clock starts (on a GPU
// box that creates the CUDA context, which is otherwise part of the first call)
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "cfs.hpp"

using namespace std;
using namespace cfs::util;
using namespace cfs::matrix::sparse;

static uint64_t fnv(uint64_t h, const void *p, size_t bytes) {
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < bytes; ++i) {
    h ^= b[i];
    h *= 1099511628211ULL;
  }
  return h;
}

int main(int argc, char **argv) {
  if (argc < 3) {
    cout << "usage: " << argv[0] << " <mmf_file> <format 0|1>" << endl;
    return 1;
  }
  const Format fmt = atoi(argv[2]) == 0 ? Format::csr : Format::sss;
  typedef chrono::steady_clock clk;
  if (argc > 3)
    cfs::util::memory::internal_free(cfs::util::memory::internal_alloc(64));
  const clk::time_point t0 = clk::now();
  SparseMatrix<int, double> *A = SparseMatrix<int, double>::create(argv[1], fmt);
  const clk::time_point t1 = clk::now();
  CSRMatrix<int, double> *C = static_cast<CSRMatrix<int, double> *>(A);
  const int n = A->nrows(), nnz = A->nnz();
  uint64_t h = 14695981039346656037ULL;
  h = fnv(h, C->rowptr(), ((size_t)n + 1) * sizeof(int));
  h = fnv(h, C->colind(), (size_t)nnz * sizeof(int));
  h = fnv(h, C->values(), (size_t)nnz * sizeof(double));
  const clk::time_point t2 = clk::now();
  printf("load(sec) %.4f nrows %d ncols %d nnz %d symmetric %d csr_fnv %016llx "
         "checksum(sec) %.4f\n",
         chrono::duration<double>(t1 - t0).count(), n, A->ncols(), nnz,
         (int)A->symmetric(), (unsigned long long)h,
         chrono::duration<double>(t2 - t1).count());
  delete A;
  return 0;
}
