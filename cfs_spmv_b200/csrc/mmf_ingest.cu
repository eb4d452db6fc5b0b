// mmf_ingest.cu -- Matrix Market ingest on the GPU.
//
// Replaces, for the file constructor of CSRMatrix, the reference's loader
// MMF<I,V> (include/io/mmf.hpp:179-343, src/mmf.cpp:6-44: getline / substr /
// atoi / atof per line, mirror, std::sort) and the serial CSR fill
// (include/matrix/csr_matrix.tpp:74-107). The banner, comment and size lines
// are read by the host (cfs_spmv_b200/host/mmf.cpp, a few lines of text); the
// entry lines -- all the work -- are processed here:
//
//   text image -> HBM            one copy
//   line ends                    cub::DeviceSelect over the bytes
//   one thread per line          trim / split / atoi / atoi / atof, bit-exact
//                                (decfloat.cuh); lines this parser must not
//                                decide are listed and patched in from the
//                                host's strtol / strtod
//   mirror + order by (row,col)  scan + 64-bit radix sort (stable). One
//                                (row, col) twice with DIFFERENT values: the
//                                reference's order is std::sort's, the file
//                                goes to the host loader (same algorithm)
//   full CSR                     row boundaries of the sorted keys
//
// Anything the reference treats as fatal or undefined (short lines, too few
// lines, indices out of range) returns CFS_ERR_NEEDS_HOST: the caller runs the
// host loader, which prints the reference's message and exits like it.
#include <stdlib.h>
#include <string.h>

#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "common.cuh"
#include "decfloat.cuh"

namespace cfsb {
namespace {

constexpr int kThreads = 256;

__device__ const uint64_t g_pow5[CFS_POW5_TABLE_WORDS] = CFS_POW5_TABLE_INIT;

struct IsNewline {
  const char *text;
  __host__ __device__ bool operator()(unsigned int i) const {
    return text[i] == '\n';
  }
};

struct NewlineFlag {
  const char *text;
  __host__ __device__ int operator()(unsigned int i) const {
    return text[i] == '\n' ? 1 : 0;
  }
};

enum { kErrShortLine = 1, kErrRange = 2 };

// line i is text[begin_i, nl[i]) with begin_0 = first and begin_i = nl[i-1]+1
// (nl holds positions relative to `text`)
__global__ void __launch_bounds__(kThreads)
    parse_lines_kernel(const char *__restrict__ text, unsigned int first,
                       const unsigned int *__restrict__ nl, long long nlines,
                       int nrows, int ncols, int zero_based, int mirror,
                       int *__restrict__ row, int *__restrict__ col,
                       double *__restrict__ val, int *__restrict__ count,
                       unsigned int *__restrict__ host_list,
                       unsigned int *__restrict__ host_count,
                       unsigned int host_capacity,
                       unsigned long long *__restrict__ first_error) {
  const long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
  if (i >= nlines)
    return;
  const char *b = text + (i == 0 ? first : nl[i - 1] + 1);
  const char *e = text + nl[i];
  dec::LineTokens lt;
  dec::split_line(b, e, &lt);
  int r = 0, c = 0;
  double v = 0.42; // value of a two-token line (include/io/mmf.hpp:331-334)
  int status = dec::kParsed;
  if (lt.ntokens < 2) {
    atomicMin(first_error, ((unsigned long long)i << 2) | kErrShortLine);
    status = -1;
  } else {
    status |= dec::parse_int(lt.tok[0], lt.tok_end[0], &r);
    status |= dec::parse_int(lt.tok[1], lt.tok_end[1], &c);
    if (lt.ntokens >= 3)
      status |= dec::parse_double(lt.tok[2], lt.tok_end[2], g_pow5, &v);
  }
  if (status == dec::kNeedHost) {
    const unsigned int slot = atomicAdd(host_count, 1u);
    if (slot < host_capacity)
      host_list[slot] = (unsigned int)i;
    r = c = 1; // placeholder until the host's answer is patched in
  } else if (status == dec::kParsed) {
    if (zero_based) {
      ++r;
      ++c;
    }
    if (r < 1 || r > nrows || c < 1 || c > ncols) {
      atomicMin(first_error, ((unsigned long long)i << 2) | kErrRange);
      r = c = 1;
    }
  } else {
    r = c = 1;
  }
  row[i] = r;
  col[i] = c;
  val[i] = v;
  count[i] = (mirror && r != c) ? 2 : 1;
}

__global__ void __launch_bounds__(kThreads)
    patch_lines_kernel(long long npatch, const unsigned int *__restrict__ line,
                       const int *__restrict__ prow,
                       const int *__restrict__ pcol,
                       const double *__restrict__ pval, int mirror,
                       int *__restrict__ row, int *__restrict__ col,
                       double *__restrict__ val, int *__restrict__ count) {
  const long long k = blockIdx.x * (long long)kThreads + threadIdx.x;
  if (k >= npatch)
    return;
  const unsigned int i = line[k];
  row[i] = prow[k];
  col[i] = pcol[k];
  val[i] = pval[k];
  count[i] = (mirror && prow[k] != pcol[k]) ? 2 : 1;
}

// entries (1-based row, col) -> sort keys; a mirrored entry follows its
// original, as in the reference's insertion order (include/io/mmf.hpp:293-300)
__global__ void __launch_bounds__(kThreads)
    expand_kernel(long long nlines, const int *__restrict__ row,
                  const int *__restrict__ col, const double *__restrict__ val,
                  const int *__restrict__ count,
                  const long long *__restrict__ offset,
                  unsigned long long *__restrict__ key,
                  double *__restrict__ out_val) {
  const long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
  if (i >= nlines)
    return;
  const long long o = offset ? offset[i] : i;
  const unsigned long long r = (unsigned int)row[i], c = (unsigned int)col[i];
  key[o] = (r << 32) | c;
  out_val[o] = val[i];
  if (count[i] == 2) {
    key[o + 1] = (c << 32) | r;
    out_val[o + 1] = val[i];
  }
}

// *ambiguous is raised when one (row, col) occurs twice with different values:
// the reference leaves the order of such entries to std::sort (unstable), so
// only the host loader -- which runs the same algorithm -- can reproduce it.
template <typename T>
__global__ void __launch_bounds__(kThreads)
    csr_from_sorted_kernel(long long nnz, int nrows,
                           const unsigned long long *__restrict__ key,
                           const double *__restrict__ val,
                           int *__restrict__ rowptr, int *__restrict__ colind,
                           T *__restrict__ values, int *__restrict__ ambiguous) {
  const long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
  if (i > nnz)
    return;
  if (i > 0 && i < nnz && key[i] == key[i - 1] &&
      __double_as_longlong(val[i]) != __double_as_longlong(val[i - 1]))
    *ambiguous = 1;
  // rows (prev, cur] start at entry i; rows without entries repeat the offset
  const int prev = i == 0 ? -1 : (int)(key[i - 1] >> 32) - 1;
  const int cur = i == nnz ? nrows : (int)(key[i] >> 32) - 1;
  for (int r = prev + 1; r <= cur; ++r)
    rowptr[r] = (int)i;
  if (i < nnz) {
    colind[i] = (int)(unsigned int)key[i] - 1;
    values[i] = (T)val[i];
  }
}

// what include/io/mmf.hpp:309-343 does with one line, on the host
struct HostLine {
  int ntokens, row, col;
  double val;
};

HostLine host_parse_line(const char *b, const char *e) {
  dec::LineTokens lt;
  dec::split_line(b, e, &lt);
  HostLine h;
  h.ntokens = lt.ntokens;
  h.row = h.col = 0;
  h.val = 0.42;
  if (lt.ntokens >= 2) {
    h.row = atoi(std::string(lt.tok[0], lt.tok_end[0]).c_str());
    h.col = atoi(std::string(lt.tok[1], lt.tok_end[1]).c_str());
  }
  if (lt.ntokens >= 3)
    h.val = atof(std::string(lt.tok[2], lt.tok_end[2]).c_str());
  return h;
}

// Temporaries come from the stream-ordered pool (cudaMallocAsync): a buffer
// released early (the text, the per-line arrays) is handed to a later one (the
// sort buffers) without another trip to the driver -- cudaMalloc / cudaFree of
// hundreds of MB each were most of the wall time of an ingest, and its jitter.
template <typename T> struct PoolArray {
  T *p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  PoolArray() {}
  PoolArray(const PoolArray &) = delete;
  PoolArray &operator=(const PoolArray &) = delete;
  ~PoolArray() { release(); }
  int alloc(size_t count, cudaStream_t stream) {
    release();
    s = stream;
    n = count;
    CFS_CUDA_TRY(cudaMallocAsync((void **)&p, (count ? count : 1) * sizeof(T),
                                 stream));
    return CFS_OK;
  }
  void release() {
    if (p)
      cudaFreeAsync(p, s);
    p = nullptr;
    n = 0;
  }
};

int keep_pool_memory() {
  static bool done = false;
  if (done)
    return CFS_OK;
  int dev = 0;
  cudaMemPool_t pool;
  CFS_CUDA_TRY(cudaGetDevice(&dev));
  CFS_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
  unsigned long long keep = ~0ULL; // do not hand freed blocks back at every sync
  CFS_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold,
                                       &keep));
  done = true;
  return CFS_OK;
}

struct Timer {
  cudaEvent_t e[2] = {nullptr, nullptr};
  cudaStream_t s;
  explicit Timer(cudaStream_t stream) : s(stream) {
    cudaEventCreate(&e[0]);
    cudaEventCreate(&e[1]);
    cudaEventRecord(e[0], s);
  }
  float stop() {
    float ms = 0;
    cudaEventRecord(e[1], s);
    cudaEventSynchronize(e[1]);
    cudaEventElapsedTime(&ms, e[0], e[1]);
    cudaEventRecord(e[0], s);
    return ms;
  }
  ~Timer() {
    cudaEventDestroy(e[0]);
    cudaEventDestroy(e[1]);
  }
};

int needs_host(const char *why) {
  set_error("Matrix Market ingest: %s -- the host loader decides", why);
  return CFS_ERR_NEEDS_HOST;
}

template <typename T>
int build_csr(cfs_matrix_s *m, long long nnz, const unsigned long long *key,
              const double *val, cudaStream_t s) {
  CFS_TRY(m->own_rowptr.alloc((size_t)m->nrows + 1));
  CFS_TRY(m->own_colind.alloc((size_t)nnz));
  CFS_TRY(m->own_values.alloc((size_t)nnz * sizeof(T)));
  DevArray<int> ambiguous;
  CFS_TRY(ambiguous.alloc(1));
  CFS_CUDA_TRY(cudaMemsetAsync(ambiguous.p, 0, 4, s));
  const unsigned grid = (unsigned)((nnz + 1 + kThreads - 1) / kThreads);
  csr_from_sorted_kernel<T><<<grid, kThreads, 0, s>>>(
      nnz, m->nrows, key, val, m->own_rowptr.p, m->own_colind.p,
      (T *)m->own_values.p, ambiguous.p);
  CFS_CUDA_TRY(cudaGetLastError());
  int flag = 0;
  CFS_CUDA_TRY(cudaMemcpyAsync(&flag, ambiguous.p, 4, cudaMemcpyDeviceToHost,
                               s));
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  if (flag)
    return needs_host("duplicate entries with different values (their order "
                      "is std::sort's)");
  return CFS_OK;
}

int ingest(cfs_matrix_s *m, const cfs_mmf_text *in, cfs_mmf_report *rep) {
  cudaStream_t s = m->stream;
  const long long L = in->declared;
  const size_t body = in->bytes - in->entries_offset;
  CFS_TRY(keep_pool_memory());
  Timer timer(s);

  // ---- text -> HBM (only the entry lines)
  PoolArray<char> text;
  CFS_TRY(text.alloc(body + 1, s));
  if (body)
    CFS_CUDA_TRY(cudaMemcpyAsync(text.p, in->text + in->entries_offset, body,
                                 cudaMemcpyHostToDevice, s));
  rep->ms_upload = timer.stop();

  // ---- line ends
  PoolArray<int> d_total;
  CFS_TRY(d_total.alloc(1, s));
  cub::CountingInputIterator<unsigned int> byte_ids(0);
  PoolArray<char> tmp;
  size_t tmp_bytes = 0;
  {
    cub::TransformInputIterator<int, NewlineFlag,
                                cub::CountingInputIterator<unsigned int>>
        flags(byte_ids, NewlineFlag{text.p});
    CFS_CUDA_TRY(cub::DeviceReduce::Sum(nullptr, tmp_bytes, flags, d_total.p,
                                        (long long)body, s));
    CFS_TRY(tmp.alloc(tmp_bytes, s));
    CFS_CUDA_TRY(cub::DeviceReduce::Sum(tmp.p, tmp_bytes, flags, d_total.p,
                                        (long long)body, s));
  }
  int total_lines = 0;
  CFS_CUDA_TRY(cudaMemcpyAsync(&total_lines, d_total.p, 4,
                               cudaMemcpyDeviceToHost, s));
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  if (total_lines < L) // "Requesting dereference, but mmf ended."
    return needs_host("fewer entry lines than the size line announces");
  PoolArray<unsigned int> nl;
  CFS_TRY(nl.alloc((size_t)total_lines, s));
  CFS_CUDA_TRY(cub::DeviceSelect::If(nullptr, tmp_bytes, byte_ids, nl.p,
                                     d_total.p, (long long)body,
                                     IsNewline{text.p}, s));
  CFS_TRY(tmp.alloc(tmp_bytes, s));
  CFS_CUDA_TRY(cub::DeviceSelect::If(tmp.p, tmp_bytes, byte_ids, nl.p,
                                     d_total.p, (long long)body,
                                     IsNewline{text.p}, s));

  // ---- one thread per line
  PoolArray<int> row, col, count;
  PoolArray<double> val;
  CFS_TRY(row.alloc((size_t)L, s));
  CFS_TRY(col.alloc((size_t)L, s));
  CFS_TRY(count.alloc((size_t)L, s));
  CFS_TRY(val.alloc((size_t)L, s));
  const unsigned int host_capacity = (unsigned int)L;
  PoolArray<unsigned int> host_list, host_count;
  PoolArray<unsigned long long> first_error;
  CFS_TRY(host_list.alloc((size_t)L, s));
  CFS_TRY(host_count.alloc(1, s));
  CFS_TRY(first_error.alloc(1, s));
  CFS_CUDA_TRY(cudaMemsetAsync(host_count.p, 0, 4, s));
  CFS_CUDA_TRY(cudaMemsetAsync(first_error.p, 0xff, 8, s));
  const int mirror = in->file_symmetric ? 1 : 0;
  const unsigned grid = (unsigned)((L + kThreads - 1) / kThreads);
  if (L)
    parse_lines_kernel<<<grid, kThreads, 0, s>>>(
        text.p, 0u, nl.p, L, in->nrows, in->ncols, in->zero_based, mirror,
        row.p, col.p, val.p, count.p, host_list.p, host_count.p, host_capacity,
        first_error.p);
  CFS_CUDA_TRY(cudaGetLastError());
  unsigned int nhost = 0;
  unsigned long long err = 0;
  CFS_CUDA_TRY(cudaMemcpyAsync(&nhost, host_count.p, 4, cudaMemcpyDeviceToHost,
                               s));
  CFS_CUDA_TRY(cudaMemcpyAsync(&err, first_error.p, 8, cudaMemcpyDeviceToHost,
                               s));
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  if (err != ~0ULL)
    return needs_host((err & 3) == kErrShortLine
                          ? "an entry line with fewer than two tokens"
                          : "an index outside the matrix");

  // ---- lines only strtol / strtod may decide
  rep->host_lines = nhost;
  if (nhost) {
    std::vector<unsigned int> lines(nhost), ends((size_t)L);
    CFS_CUDA_TRY(cudaMemcpy(lines.data(), host_list.p, (size_t)nhost * 4,
                            cudaMemcpyDeviceToHost));
    CFS_CUDA_TRY(cudaMemcpy(ends.data(), nl.p, (size_t)L * 4,
                            cudaMemcpyDeviceToHost));
    std::vector<int> prow(nhost), pcol(nhost);
    std::vector<double> pval(nhost);
    const char *base = in->text + in->entries_offset;
    for (unsigned int k = 0; k < nhost; ++k) {
      const unsigned int i = lines[k];
      const HostLine h = host_parse_line(base + (i ? ends[i - 1] + 1 : 0),
                                         base + ends[i]);
      int r = h.row, c = h.col;
      if (in->zero_based) {
        ++r;
        ++c;
      }
      if (h.ntokens < 2 || r < 1 || r > in->nrows || c < 1 || c > in->ncols)
        return needs_host("an unusable entry line");
      prow[k] = r;
      pcol[k] = c;
      pval[k] = h.val;
    }
    PoolArray<int> d_prow, d_pcol;
    PoolArray<double> d_pval;
    CFS_TRY(d_prow.alloc(nhost, s));
    CFS_TRY(d_pcol.alloc(nhost, s));
    CFS_TRY(d_pval.alloc(nhost, s));
    CFS_CUDA_TRY(cudaMemcpyAsync(d_prow.p, prow.data(), (size_t)nhost * 4,
                                 cudaMemcpyHostToDevice, s));
    CFS_CUDA_TRY(cudaMemcpyAsync(d_pcol.p, pcol.data(), (size_t)nhost * 4,
                                 cudaMemcpyHostToDevice, s));
    CFS_CUDA_TRY(cudaMemcpyAsync(d_pval.p, pval.data(), (size_t)nhost * 8,
                                 cudaMemcpyHostToDevice, s));
    patch_lines_kernel<<<(nhost + kThreads - 1) / kThreads, kThreads, 0, s>>>(
        nhost, host_list.p, d_prow.p, d_pcol.p, d_pval.p, mirror, row.p, col.p,
        val.p, count.p);
    CFS_CUDA_TRY(cudaGetLastError());
    CFS_CUDA_TRY(cudaStreamSynchronize(s));
  }
  text.release();
  nl.release();
  host_list.release();
  rep->ms_parse = timer.stop();

  // ---- mirror the off-diagonal entries of a symmetric file
  long long nnz = L;
  PoolArray<long long> offset;
  if (mirror && L) {
    CFS_TRY(offset.alloc((size_t)L, s));
    CFS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, count.p,
                                               offset.p, L, s));
    CFS_TRY(tmp.alloc(tmp_bytes, s));
    CFS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, count.p,
                                               offset.p, L, s));
    long long last_off = 0;
    int last_cnt = 0;
    CFS_CUDA_TRY(cudaMemcpyAsync(&last_off, offset.p + (L - 1), 8,
                                 cudaMemcpyDeviceToHost, s));
    CFS_CUDA_TRY(cudaMemcpyAsync(&last_cnt, count.p + (L - 1), 4,
                                 cudaMemcpyDeviceToHost, s));
    CFS_CUDA_TRY(cudaStreamSynchronize(s));
    nnz = last_off + last_cnt;
  }
  if (nnz > 0x7fffffffLL)
    return needs_host("more entries than a 32-bit index holds");
  PoolArray<unsigned long long> key, key_alt;
  PoolArray<double> xval, xval_alt;
  CFS_TRY(key.alloc((size_t)nnz, s));
  CFS_TRY(key_alt.alloc((size_t)nnz, s));
  CFS_TRY(xval.alloc((size_t)nnz, s));
  CFS_TRY(xval_alt.alloc((size_t)nnz, s));
  if (L)
    expand_kernel<<<grid, kThreads, 0, s>>>(L, row.p, col.p, val.p, count.p,
                                            mirror ? offset.p : nullptr, key.p,
                                            xval.p);
  CFS_CUDA_TRY(cudaGetLastError());
  row.release();
  col.release();
  val.release();
  count.release();
  offset.release();

  // ---- order by (row, col)
  int row_bits = 1;
  while (row_bits < 31 && (1LL << row_bits) <= in->nrows)
    ++row_bits;
  cub::DoubleBuffer<unsigned long long> keys(key.p, key_alt.p);
  cub::DoubleBuffer<double> vals(xval.p, xval_alt.p);
  CFS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, vals,
                                               nnz, 0, 32 + row_bits, s));
  CFS_TRY(tmp.alloc(tmp_bytes, s));
  if (nnz)
    CFS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys, vals,
                                                 nnz, 0, 32 + row_bits, s));
  rep->ms_sort = timer.stop();

  // ---- full CSR
  CFS_TRY(m->is_double
              ? build_csr<double>(m, nnz, keys.Current(), vals.Current(), s)
              : build_csr<float>(m, nnz, keys.Current(), vals.Current(), s));
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  rep->ms_build = timer.stop();
  m->nnz_full = nnz;
  m->csr_rowptr = m->own_rowptr.p;
  m->csr_colind = m->own_colind.p;
  m->csr_values = m->own_values.p;
  rep->nnz = nnz;
  return CFS_OK;
}

} // namespace
} // namespace cfsb

using namespace cfsb;

extern "C" {

int cfs_cuda_matrix_create_from_mmf(cfs_mat_t *out, const cfs_mmf_text *in,
                                    int is_double, int symmetric,
                                    cfs_mmf_report *report) {
  if (!out || !in || !in->text || in->entries_offset > in->bytes ||
      in->declared < 0 || in->nrows < 0 || in->ncols < 0) {
    set_error("cfs_cuda_matrix_create_from_mmf: bad arguments");
    return CFS_ERR_INVALID;
  }
  cfs_mmf_report local;
  if (!report)
    report = &local;
  memset(report, 0, sizeof(*report));
  if (in->bytes - in->entries_offset >= 0xffffffffULL)
    return needs_host("file image of 4 GiB or more");
  CFS_TRY(require_device());
  const int device = current_device();
  cfs_matrix_s *m = new cfs_matrix_s;
  m->device = device;
  m->is_double = is_double != 0;
  // asking for a symmetric format on a general file quietly gives plain CSR
  // (include/matrix/csr_matrix.tpp:13-15)
  m->symmetric = symmetric != 0 && in->file_symmetric != 0;
  m->nrows = in->nrows;
  m->ncols = in->ncols;
  m->global_nrows = in->nrows;
  int status = CFS_OK;
  if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) !=
      cudaSuccess)
    status = cuda_fail(cudaGetLastError(), "cudaStreamCreate", __FILE__,
                       __LINE__);
  if (status == CFS_OK)
    status = ingest(m, in, report);
  { // hand the pool's blocks back: tune() allocates with cudaMalloc
    cudaMemPool_t pool;
    if (m->stream)
      cudaStreamSynchronize(m->stream);
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
      cudaMemPoolTrimTo(pool, 0);
    cudaGetLastError();
  }
  if (status != CFS_OK) {
    cfs_cuda_matrix_destroy(m);
    return status;
  }
  *out = m;
  return CFS_OK;
}

int cfs_cuda_matrix_download_csr(cfs_mat_t m, int32_t *rowptr, int32_t *colind,
                                 void *values) {
  if (!m)
    return CFS_ERR_INVALID;
  if (!m->csr_rowptr) {
    set_error("cfs_cuda_matrix_download_csr: the full CSR was released by "
              "tune() (compress_symmetry, csr_matrix.tpp:1700-1706)");
    return CFS_ERR_STATE;
  }
  CFS_CUDA_TRY(cudaSetDevice(m->device));
  if (rowptr)
    CFS_CUDA_TRY(cudaMemcpy(rowptr, m->csr_rowptr, ((size_t)m->nrows + 1) * 4,
                            cudaMemcpyDeviceToHost));
  if (colind && m->nnz_full)
    CFS_CUDA_TRY(cudaMemcpy(colind, m->csr_colind, (size_t)m->nnz_full * 4,
                            cudaMemcpyDeviceToHost));
  if (values && m->nnz_full)
    CFS_CUDA_TRY(cudaMemcpy(values, m->csr_values,
                            (size_t)m->nnz_full * m->vsize(),
                            cudaMemcpyDeviceToHost));
  return CFS_OK;
}

} // extern "C"
