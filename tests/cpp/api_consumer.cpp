// api_consumer.cpp -- exercises the C++ API exactly the way the reference's
// bench and test programs do (create -> SpDMV -> operator() twice -> compare
// with the plain-CSR path using isEqual), plus the in-memory array constructor
// the reference's drivers never touch. Prints one PASSED!/FAILED! per check.
//
//   api_consumer <mmf_file> <format 0|1|2>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "cfs.hpp"

using namespace std;
using namespace cfs::util;
using namespace cfs::util::memory;
using namespace cfs::util::runtime;
using namespace cfs::matrix::sparse;
using namespace cfs::kernel::sparse;

// double: the reference's element-wise isEqual (test_spmv_mmf.cpp:94-104, which
// is hard-wired to double). float: normwise 1e-5, the north-star tolerance --
// element-wise relative checks are meaningless for single precision where
// y[i] cancels to ~0.
static bool same(const double *a, const double *b, int n) {
  for (int i = 0; i < n; ++i)
    if (!isEqual(a[i], b[i])) {
      cout << "element " << i << " differs: " << a[i] << " vs " << b[i] << endl;
      return false;
    }
  return true;
}
static bool same(const float *a, const float *b, int n) {
  double num = 0, den = 0;
  for (int i = 0; i < n; ++i) {
    num += ((double)a[i] - b[i]) * ((double)a[i] - b[i]);
    den += (double)b[i] * b[i];
  }
  if (num > 1e-10 * den) {
    cout << "normwise error " << sqrt(num / den) << endl;
    return false;
  }
  return true;
}

template <typename V> static int run(const string &file, Format fmt) {
  int failures = 0;
  // reference CSR result (Format::csr, Tuning::None), like test_spmv_mmf
  SparseMatrix<int, V> *C = SparseMatrix<int, V>::create(file, Format::csr);
  const int M = C->nrows(), N = C->ncols();
  V *x = (V *)internal_alloc(N * sizeof(V));
  V *y = (V *)internal_alloc(M * sizeof(V));
  V *y_csr = (V *)internal_alloc(M * sizeof(V));
  mt19937 gen(12345);
  uniform_real_distribution<> dis(10.01, 20.42);
  for (int i = 0; i < N; ++i)
    x[i] = (V)dis(gen);
  // keep a copy of the full CSR for the array constructor below
  CSRMatrix<int, V> *Cc = static_cast<CSRMatrix<int, V> *>(C);
  vector<int> rp(Cc->rowptr(), Cc->rowptr() + M + 1);
  vector<int> ci(Cc->colind(), Cc->colind() + C->nnz());
  vector<V> va(Cc->values(), Cc->values() + C->nnz());
  SpDMV<int, V> csr(C, Tuning::None);
  csr(y_csr, M, x, N);

  // 1. file constructor in the requested format, called twice on a dirty y
  SparseMatrix<int, V> *A = SparseMatrix<int, V>::create(file, fmt);
  const bool was_symmetric = A->symmetric();
  SpDMV<int, V> fn(A, Tuning::Aggressive);
  for (int i = 0; i < M; ++i)
    y[i] = (V)-7;
  fn(y, M, x, N);
  fn(y, M, x, N);
  bool ok = same(y, y_csr, M) && A->nnz() == C->nnz() && A->size() > 0;
  cout << "file ctor: " << (ok ? "PASSED!" : "FAILED!") << endl;
  failures += !ok;
  // 1b. the host rewrites x between two calls and reads y after each (the
  // vectors come from internal_alloc: unified memory by default, so this is
  // the page-migration path in both directions)
  for (int i = 0; i < N; ++i)
    x[i] *= (V)2;
  fn(y, M, x, N);
  for (int i = 0; i < M; ++i)
    y[i] *= (V)0.5;
  ok = same(y, y_csr, M);
  for (int i = 0; i < N; ++i)
    x[i] *= (V)0.5;
  fn(y, M, x, N);
  ok = ok && same(y, y_csr, M);
  cout << "host-mutated x: " << (ok ? "PASSED!" : "FAILED!") << endl;
  failures += !ok;
  // a symmetric file matrix gives its full CSR back during tune()
  if (was_symmetric) {
    CSRMatrix<int, V> *Ac = static_cast<CSRMatrix<int, V> *>(A);
    ok = Ac->rowptr() == nullptr && Ac->colind() == nullptr &&
         Ac->values() == nullptr;
    cout << "ownership: " << (ok ? "PASSED!" : "FAILED!") << endl;
    failures += !ok;
  }

  // 2. array constructor: never owns, never frees the caller's arrays
  {
    CSRMatrix<int, V> B(rp.data(), ci.data(), va.data(), M, N, was_symmetric);
    SpDMV<int, V> fb(&B);
    fb(y, M, x, N);
    ok = same(y, y_csr, M) && B.rowptr() == rp.data() && B.nnz() == C->nnz();
    cout << "array ctor: " << (ok ? "PASSED!" : "FAILED!") << endl;
    failures += !ok;
  }
  // 3. the solver loop on top of the same matrix: A z = (A x) gives z = x
  if (was_symmetric) {
    const bool dp = sizeof(V) == 8;
    V *z = (V *)internal_alloc(N * sizeof(V));
    for (int i = 0; i < N; ++i)
      z[i] = (V)0;
    ConjugateGradient<int, V> cg(A);
    const int its = cg(z, y_csr, N, 5000, dp ? 1e-12 : 1e-6);
    if (cg.breakdown()) {
      cout << "cg: SKIPPED (matrix is not positive definite)" << endl;
    } else {
      double num = 0, den = 0;
      for (int i = 0; i < N; ++i) {
        num += ((double)z[i] - x[i]) * ((double)z[i] - x[i]);
        den += (double)x[i] * x[i];
      }
      ok = cg.converged() && sqrt(num / den) <= (dp ? 1e-8 : 1e-3);
      cout << "cg (" << its << " iterations, error " << sqrt(num / den)
           << "): " << (ok ? "PASSED!" : "FAILED!") << endl;
      failures += !ok;
    }
    internal_free(z);
  }
  cout << "size(MB): " << A->size() / (float)(1024 * 1024)
       << " threads: " << get_num_threads() << endl;
  delete A;
  delete C;
  internal_free(x);
  internal_free(y);
  internal_free(y_csr);
  return failures;
}

int main(int argc, char **argv) {
  if (argc < 3) {
    cerr << "usage: " << argv[0] << " <mmf_file> <format 0|1|2>" << endl;
    return 2;
  }
  const int f = atoi(argv[2]);
  const Format fmt = f == 1 ? Format::sss : f == 2 ? Format::hyb : Format::csr;
  int failures = run<double>(argv[1], fmt);
  failures += run<float>(argv[1], fmt);
  cout << (failures ? "FAILED!" : "ALL PASSED!") << endl;
  return failures ? 1 : 0;
}
