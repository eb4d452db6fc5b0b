// ref_tool -- drives the UNMODIFIED reference (/root/reference) so that its
// results can be used as the parity oracle and as the CPU baseline.
//
// TEST / BENCH INFRASTRUCTURE ONLY. Built into oracle/_ref/ by
// oracle/ref_build/Makefile, from the reference sources where they lie.
// Nothing in the product (cfs_spmv_b200/, include/) links or runs this.
//
//   ref_tool dump  <input> <P> <d|s> <xseed> <out.bin>
//       runs SparseMatrix/CSRMatrix + SpDMV exactly as bench/test do
//       (bench/bench_spmv_mmf.cpp:100-167, test/test_spmv_mmf.cpp:54-89) and
//       writes the private preprocessing metadata plus y to <out.bin>.
//   ref_tool dumpcsr <input> <P> <d|s> <xseed> <out.bin> <A|N>
//       the NON-symmetric path (Format::csr): array ctor with symmetric=false,
//       tune(Aggressive -> partition_by_nnz, csr_matrix.tpp:438-541 | None ->
//       partition_by_nrows, :404-435), cpu_mv / cpu_mv_serial (:2665-2704);
//       writes row_split_ and y.
//   ref_tool bench <input> <P> <d|s> <xseed> <loops>
//       times the CFS SpMV like bench_spmv_mmf.cpp:145-168 and prints one
//       JSON line.
//
//   <input> is  mtx:<file.mtx>              (file ctor, csr_matrix.tpp:9)
//           or  csr:<file.bin>              (array ctor, csr_matrix.tpp:114)
//           or  gen:lap7:nx:ny:nz[:seed] | gen:lap27:nx:ny:nz[:seed]
//               (seed != 0: one coefficient per edge, see cfs_gen.h)
//               | gen:banded:nrows:bw:per_row_x16:seed      (array ctor)
//
// The metadata of the reference is private (csr_matrix.hpp:77-124); this tool
// reads it by compiling the reference headers with `private` spelled `public`.
#include <algorithm>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <omp.h>
#include <sched.h>

#include "cfs_gen.h" // cfs_spmv_b200/csrc: shared synthetic generators

#define private public
#include "cfs.hpp"
#include "kernel/sparse_kernel.tpp"
#include "matrix/csr_matrix.tpp"
#include "matrix/sparse_matrix.tpp"
#undef private

using namespace cfs::util;
using namespace cfs::util::memory;
using namespace cfs::matrix::sparse;
using namespace cfs::kernel::sparse;

namespace {

struct Writer {
  FILE *f;
  explicit Writer(const char *path) : f(fopen(path, "wb")) {
    if (!f) {
      perror(path);
      exit(2);
    }
  }
  ~Writer() { fclose(f); }
  void put(const char *name, int dtype, uint64_t count, const void *data,
           size_t elem) {
    uint32_t nl = (uint32_t)strlen(name);
    uint8_t dt = (uint8_t)dtype;
    fwrite(&nl, 4, 1, f);
    fwrite(name, 1, nl, f);
    fwrite(&dt, 1, 1, f);
    fwrite(&count, 8, 1, f);
    if (count)
      fwrite(data, elem, count, f);
  }
  void i32(const char *n, const std::vector<int> &v) {
    put(n, 0, v.size(), v.data(), 4);
  }
  void i32(const char *n, const int *p, size_t c) { put(n, 0, c, p, 4); }
  void scalar(const char *n, long long v) {
    int64_t x = v;
    put(n, 1, 1, &x, 8);
  }
  void real(const char *n, const std::vector<float> &v) {
    put(n, 2, v.size(), v.data(), 4);
  }
  void real(const char *n, const std::vector<double> &v) {
    put(n, 3, v.size(), v.data(), 8);
  }
};

struct HostCsr {
  int nrows = 0, ncols = 0;
  std::vector<int> rowptr, colind;
  std::vector<double> values;
};

std::vector<std::string> split_colon(const std::string &s) {
  std::vector<std::string> out;
  std::stringstream ss(s);
  std::string tok;
  while (std::getline(ss, tok, ':'))
    out.push_back(tok);
  return out;
}

void load_csr_bin(const std::string &path, HostCsr &m) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) {
    perror(path.c_str());
    exit(2);
  }
  int64_t hdr[3];
  if (fread(hdr, 8, 3, f) != 3)
    exit(2);
  m.nrows = (int)hdr[0];
  m.ncols = (int)hdr[1];
  m.rowptr.resize(m.nrows + 1);
  m.colind.resize(hdr[2]);
  m.values.resize(hdr[2]);
  if (fread(m.rowptr.data(), 4, m.nrows + 1, f) != (size_t)m.nrows + 1 ||
      fread(m.colind.data(), 4, hdr[2], f) != (size_t)hdr[2] ||
      fread(m.values.data(), 8, hdr[2], f) != (size_t)hdr[2])
    exit(2);
  fclose(f);
}

void generate(const std::vector<std::string> &a, HostCsr &m) {
  cfs_gen_spec g;
  if (a[1] == "lap7" || a[1] == "lap27")
    g = cfs_gen_laplacian_seeded(
        a[1] == "lap7" ? 7 : 27, atoi(a[2].c_str()), atoi(a[3].c_str()),
        atoi(a[4].c_str()), a.size() > 5 ? strtoull(a[5].c_str(), 0, 10) : 0);
  else if (a[1] == "banded")
    g = cfs_gen_banded(atoll(a[2].c_str()), atoi(a[3].c_str()),
                       atoi(a[4].c_str()), strtoull(a[5].c_str(), 0, 10));
  else {
    fprintf(stderr, "unknown generator %s\n", a[1].c_str());
    exit(2);
  }
  const int64_t n = g.nrows;
  m.nrows = m.ncols = (int)n;
  m.rowptr.assign(n + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    m.rowptr[i + 1] = cfs_gen_row(&g, i, nullptr, nullptr);
  for (int64_t i = 0; i < n; ++i)
    m.rowptr[i + 1] += m.rowptr[i];
  m.colind.resize(m.rowptr[n]);
  m.values.resize(m.rowptr[n]);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    cfs_gen_row(&g, i, &m.colind[m.rowptr[i]], &m.values[m.rowptr[i]]);
}

template <typename V> struct Run {
  SparseMatrix<int, V> *A = nullptr;
  CSRMatrix<int, V> *csr = nullptr;
  HostCsr host;          // backing store for the array ctor
  std::vector<V> values; // values in the matrix precision

  void open(const std::string &input) {
    std::vector<std::string> a = split_colon(input);
    if (a[0] == "mtx") {
      A = SparseMatrix<int, V>::create(input.substr(4), Format::sss);
    } else {
      if (a[0] == "csr")
        load_csr_bin(input.substr(4), host);
      else
        generate(a, host);
      values.assign(host.values.begin(), host.values.end());
      std::vector<double>().swap(host.values);
      A = new CSRMatrix<int, V>(host.rowptr.data(), host.colind.data(),
                                values.data(), host.nrows, host.ncols,
                                /*symmetric=*/true);
    }
    csr = static_cast<CSRMatrix<int, V> *>(A);
  }
};

template <typename V>
int do_dump(const std::string &input, int P, uint64_t xseed, const char *out) {
  Run<V> r;
  r.open(input);
  const int M = r.A->nrows(), N = r.A->ncols();
  const long long nnz_full = r.A->nnz();
  V *x = (V *)internal_alloc(N * sizeof(V));
  V *y = (V *)internal_alloc(M * sizeof(V));
  for (int i = 0; i < N; ++i)
    x[i] = (V)cfs_gen_x(xseed, i);
  for (int i = 0; i < M; ++i)
    y[i] = 0; // the P=1 kernel accumulates into not-yet-assigned rows only

  // plain CSR result first (the test's comparator, test_spmv_mmf.cpp:85-89):
  // must run before tune() because the file ctor frees the full CSR there.
  std::vector<V> y_csr(M);
  std::vector<int> full_rowptr, full_colind;
  std::vector<V> full_values;
  {
    CSRMatrix<int, V> *c = r.csr;
    if (nnz_full < 2000000) { // the loaded full CSR, for loader parity
      full_rowptr.assign(c->rowptr_, c->rowptr_ + M + 1);
      full_colind.assign(c->colind_, c->colind_ + nnz_full);
      full_values.assign(c->values_, c->values_ + nnz_full);
    }
    for (int i = 0; i < M; ++i) {
      V t = 0;
      for (int j = c->rowptr_[i]; j < c->rowptr_[i + 1]; ++j)
        t += c->values_[j] * x[c->colind_[j]];
      y_csr[i] = t;
    }
  }

  SpDMV<int, V> fn(r.A, Tuning::Aggressive);
  for (int rep = 0; rep < 2; ++rep) // twice, like test_spmv_mmf.cpp:80-83
    fn(y, M, x, N);

  CSRMatrix<int, V> *c = r.csr;
  Writer w(out);
  w.scalar("nrows", M);
  w.scalar("ncols", N);
  w.scalar("nnz_full", nnz_full);
  w.scalar("symmetric", c->symmetric_ ? 1 : 0);
  w.scalar("P", P);
  w.scalar("ncolors", c->ncolors_);
  w.scalar("nranges", c->nranges_);
  w.scalar("nnz_low", c->nnz_low_);
  w.scalar("nnz_diag", c->nnz_diag_);
  w.scalar("size_bytes", (long long)c->size());
  w.scalar("is_double", sizeof(V) == 8);
  if (c->row_split_)
    w.i32("row_split", c->row_split_, P + 1);
  std::vector<int> part_nrows, part_offset, part_nnz_low, part_nranges;
  std::vector<int> lrowptr, lcolind, rptr, rstart, rend;
  std::vector<V> lvalues, diagonal;
  if (c->cmp_symmetry_) {
    for (int t = 0; t < P; ++t) {
      auto *d = c->sym_thread_data_[t];
      part_nrows.push_back(d->nrows_);
      part_offset.push_back(P == 1 ? 0 : d->row_offset_);
      part_nnz_low.push_back(d->nnz_low_);
      lrowptr.insert(lrowptr.end(), d->rowptr_, d->rowptr_ + d->nrows_ + 1);
      lcolind.insert(lcolind.end(), d->colind_, d->colind_ + d->nnz_low_);
      lvalues.insert(lvalues.end(), d->values_, d->values_ + d->nnz_low_);
      diagonal.insert(diagonal.end(), d->diagonal_, d->diagonal_ + d->nrows_);
      if (P > 1) {
        part_nranges.push_back(d->nranges_);
        rptr.insert(rptr.end(), d->range_ptr_,
                    d->range_ptr_ + c->ncolors_ + 1);
        rstart.insert(rstart.end(), d->range_start_,
                      d->range_start_ + d->nranges_);
        rend.insert(rend.end(), d->range_end_, d->range_end_ + d->nranges_);
      }
    }
  }
  w.i32("part_nrows", part_nrows);
  w.i32("part_offset", part_offset);
  w.i32("part_nnz_low", part_nnz_low);
  w.i32("part_nranges", part_nranges);
  w.i32("lower_rowptr", lrowptr);
  w.i32("lower_colind", lcolind);
  w.real("lower_values", lvalues);
  w.real("diagonal", diagonal);
  w.i32("range_ptr", rptr);
  w.i32("range_start", rstart);
  w.i32("range_end", rend);
  w.real("x", std::vector<V>(x, x + N));
  w.real("y", std::vector<V>(y, y + M));
  w.real("y_csr", y_csr);
  w.i32("csr_rowptr", full_rowptr);
  w.i32("csr_colind", full_colind);
  w.real("csr_values", full_values);
  delete r.A;
  internal_free(x);
  internal_free(y);
  return 0;
}

template <typename V>
int do_dump_csr(const std::string &input, int P, uint64_t xseed,
                const char *out, bool aggressive) {
  std::vector<std::string> a = split_colon(input);
  HostCsr host;
  if (a[0] == "csr")
    load_csr_bin(input.substr(4), host);
  else
    generate(a, host);
  std::vector<V> values(host.values.begin(), host.values.end());
  CSRMatrix<int, V> A(host.rowptr.data(), host.colind.data(), values.data(),
                      host.nrows, host.ncols, /*symmetric=*/false);
  const int M = A.nrows(), N = A.ncols();
  V *x = (V *)internal_alloc(N * sizeof(V));
  V *y = (V *)internal_alloc(M * sizeof(V));
  for (int i = 0; i < N; ++i)
    x[i] = (V)cfs_gen_x(xseed, i);
  for (int i = 0; i < M; ++i)
    y[i] = (V)-7;
  SpDMV<int, V> fn(&A, aggressive ? Tuning::Aggressive : Tuning::None);
  fn(y, M, x, N);
  fn(y, M, x, N);
  Writer w(out);
  w.scalar("nrows", M);
  w.scalar("ncols", N);
  w.scalar("nnz_full", A.nnz());
  w.scalar("P", P);
  w.scalar("aggressive", aggressive ? 1 : 0);
  w.scalar("size_bytes", (long long)A.size());
  w.scalar("is_double", sizeof(V) == 8);
  if (A.row_split_)
    w.i32("row_split", A.row_split_, P + 1);
  w.real("y", std::vector<V>(y, y + M));
  internal_free(x);
  internal_free(y);
  return 0;
}

template <typename V>
int do_bench(const std::string &input, int P, uint64_t xseed, size_t loops,
             long warmup, const char *y_out) {
  double t0 = omp_get_wtime();
  Run<V> r;
  r.open(input);
  double t_load = omp_get_wtime() - t0;
  const int M = r.A->nrows(), N = r.A->ncols();
  const long long nnz_full = r.A->nnz();
  V *x = (V *)internal_alloc(N * sizeof(V));
  V *y = (V *)internal_alloc(M * sizeof(V));
#pragma omp parallel for schedule(static) num_threads(P)
  for (int i = 0; i < M; ++i)
    y[i] = 0;
#pragma omp parallel for schedule(static) num_threads(P)
  for (int i = 0; i < N; ++i)
    x[i] = (V)cfs_gen_x(xseed, i);

  t0 = omp_get_wtime();
  SpDMV<int, V> spdmv(r.A);
  double preproc = omp_get_wtime() - t0;
  // warm-up: loops / 2 like bench_spmv_mmf.cpp:154 unless the caller says
  const size_t nwarm = warmup >= 0 ? (size_t)warmup : loops / 2;
  for (size_t i = 0; i < nwarm; ++i)
    spdmv(y, M, x, N);
  std::vector<double> per_step(loops);
  t0 = omp_get_wtime();
  for (size_t i = 0; i < loops; ++i) {
    double s = omp_get_wtime();
    spdmv(y, M, x, N);
    per_step[i] = omp_get_wtime() - s;
  }
  double total = omp_get_wtime() - t0;
  double ysum = 0;
  for (int i = 0; i < M; ++i)
    ysum += y[i];
  std::sort(per_step.begin(), per_step.end());
  printf("{\"impl\": \"reference\", \"nrows\": %d, \"nnz_full\": %lld, "
         "\"threads\": %d, \"loops\": %zu, \"load_s\": %.6f, "
         "\"preproc_s\": %.6f, \"t_spmv_s\": %.9f, \"t_spmv_min_s\": %.9f, "
         "\"gflops\": %.6f, \"ncolors\": %d, \"size_bytes\": %zu, "
         "\"dtype\": \"%s\", \"ysum\": %.17g, \"warmup\": %zu}\n",
         M, nnz_full, P, loops, t_load, preproc, total / loops, per_step[0],
         (double)loops * 2.0 * nnz_full * 1e-9 / total, r.csr->ncolors_,
         r.A->size(), sizeof(V) == 8 ? "f64" : "f32", ysum, nwarm);
  if (y_out && y_out[0]) { // the reference's y, raw, for full-size parity
    FILE *f = fopen(y_out, "wb");
    if (!f || fwrite(y, sizeof(V), M, f) != (size_t)M) {
      perror(y_out);
      return 2;
    }
    fclose(f);
  }
  delete r.A;
  internal_free(x);
  internal_free(y);
  return 0;
}

} // namespace

int main(int argc, char **argv) {
  if (argc < 7) {
    fprintf(stderr,
            "usage: %s dump|bench <input> <P> <d|s> <xseed> <out.bin|loops> "
            "[bench: warmup (-1 = loops/2)] [bench: y_out.bin]\n",
            argv[0]);
    return 2;
  }
  const std::string cmd = argv[1], input = argv[2];
  const int P = atoi(argv[3]);
  const bool dp = argv[4][0] == 'd';
  const uint64_t xseed = strtoull(argv[5], 0, 10);
  // the reference reads P from the environment in its ctors (runtime.cpp:10)
  setenv("CFS_NUM_THREADS", argv[3], 1);
  if (cmd == "dump")
    return dp ? do_dump<double>(input, P, xseed, argv[6])
              : do_dump<float>(input, P, xseed, argv[6]);
  if (cmd == "dumpcsr") {
    const bool aggressive = argc < 8 || argv[7][0] == 'A';
    return dp ? do_dump_csr<double>(input, P, xseed, argv[6], aggressive)
              : do_dump_csr<float>(input, P, xseed, argv[6], aggressive);
  }
  if (cmd == "bench") {
    const long warmup = argc > 7 ? atol(argv[7]) : -1;
    const char *y_out = argc > 8 ? argv[8] : nullptr;
    return dp ? do_bench<double>(input, P, xseed, (size_t)atoll(argv[6]),
                                 warmup, y_out)
              : do_bench<float>(input, P, xseed, (size_t)atoll(argv[6]),
                                warmup, y_out);
  }
  return 2;
}
