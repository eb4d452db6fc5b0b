// cfs_cuda.cu -- implementation of the C ABI declared in include/cfs_cuda.h.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <mutex>
#include <unordered_map>

#include "cfs_gen.h"
#include "common.cuh"

namespace cfsb {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e),
            what, file, line);
  cudaGetLastError(); // clear the sticky non-fatal error state
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
    return CFS_ERR_NO_DEVICE;
  return CFS_ERR_CUDA;
}

static int g_device = -1;
Options g_options;

int require_device() {
  if (g_device >= 0) {
    CFS_CUDA_TRY(cudaSetDevice(g_device));
    return CFS_OK;
  }
  return cfs_cuda_init(0);
}

int current_device() { return g_device; }

enum PtrKind { kPtrHost, kPtrDevice };

static int classify(const void *p, PtrKind *kind, bool *pinned = nullptr,
                    bool *managed = nullptr) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (pinned)
    *pinned = false;
  if (managed)
    *managed = false;
  if (e != cudaSuccess) {
    cudaGetLastError();
    *kind = kPtrHost;
    return CFS_OK;
  }
  *kind = (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged)
              ? kPtrDevice
              : kPtrHost;
  if (pinned)
    *pinned = a.type == cudaMemoryTypeHost;
  if (managed)
    *managed = a.type == cudaMemoryTypeManaged;
  return CFS_OK;
}

// allocations handed out by cfs_cuda_host_alloc: the kind picks the right free;
// `seen` marks a managed vector that cfs_cuda_spmv has moved to the GPU once
struct AllocRec {
  int kind;
  size_t bytes;
  bool seen;
};
static std::mutex g_alloc_mutex;
static std::unordered_map<const void *, AllocRec> g_allocs;

static int upload_or_borrow(const void *src, size_t bytes, DevArray<char> &own,
                            const void **out) {
  PtrKind k;
  bool managed = false;
  CFS_TRY(classify(src, &k, nullptr, &managed));
  // managed matrix arrays are copied like host arrays: borrowing them would
  // page-fault them over one 64 KB block at a time
  if (k == kPtrDevice && !managed) {
    *out = src;
    return CFS_OK;
  }
  CFS_TRY(own.alloc(bytes));
  CFS_CUDA_TRY(cudaMemcpy(own.p, src, bytes, cudaMemcpyHostToDevice));
  *out = own.p;
  return CFS_OK;
}

} // namespace cfsb

using namespace cfsb;

extern "C" {

const char *cfs_cuda_last_error(void) { return g_error; }
const char *cfs_cuda_version(void) { return "cfs-b200 0.1 (sm_100a)"; }

// bumped by every accepted cfs_cuda_set_option: captured graphs bake the kernel
// choice in and are re-captured when it has moved on
static unsigned long long g_options_generation = 1;

static int set_option_checked(const char *key, long long value);

int cfs_cuda_set_option(const char *key, long long value) {
  if (!key)
    return CFS_ERR_INVALID;
  const int status = set_option_checked(key, value);
  if (status == CFS_OK)
    ++g_options_generation;
  return status;
}

static int set_option_checked(const char *key, long long value) {
  if (!strcmp(key, "value_index") && (value == 0 || value == 1)) {
    g_options.value_index = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "rechunk_pct") && value >= 100 && value <= 400) {
    g_options.rechunk_pct = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "rechunk") && (value == 0 || value == 1)) {
    g_options.rechunk = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "tile6") && (value == 0 || value == 1)) {
    g_options.tile6 = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "spmv_variant") && value >= 1 && value <= 7) {
    g_options.spmv_variant = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "ctas_per_sm") && value >= 1 && value <= 8) {
    g_options.ctas_per_sm = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "sort_rows") && (value == 0 || value == 1)) {
    g_options.sort_rows = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "hubs") && (value == 0 || value == 1)) {
    g_options.hubs = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "pipeline") && (value == 0 || value == 1)) {
    g_options.pipeline = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "pipeline_adaptive") && (value == 0 || value == 1)) {
    g_options.pipeline_adaptive = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "pipeline_taper") && (value == 0 || value == 1)) {
    g_options.pipeline_taper = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "pipeline_split") && (value == 0 || value == 1)) {
    g_options.pipeline_split = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "pipeline_smem") && value >= 0 && value <= 200 * 1024) {
    g_options.pipeline_smem = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "pipeline_skip") && value >= 0 && value <= 7) {
    g_options.pipeline_skip = (int)value; // measurement aid: wrong results
    return CFS_OK;
  }
  if (!strcmp(key, "pipeline_trace") && (value == 0 || value == 1)) {
    g_options.pipeline_trace = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "pipeline_graph") && (value == 0 || value == 1)) {
    g_options.pipeline_graph = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "pipeline_chunks") && value >= 2 && value <= 256) {
    g_options.pipeline_chunks = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "csr_layout") && (value == 0 || value == 1)) {
    g_options.csr_layout = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "multi_graph") && (value == 0 || value == 1)) {
    g_options.multi_graph = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "multi_zero_copy") && (value == 0 || value == 1)) {
    g_options.multi_zero_copy = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "slot_banks") && (value == 0 || value == 1)) {
    g_options.slot_banks = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "deterministic") && (value == 0 || value == 1)) {
    g_options.deterministic = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "keep_layouts") && (value == 0 || value == 1)) {
    g_options.keep_layouts = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "reg_blocks") && (value == 16 || value == 12)) {
    g_options.reg_blocks = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "l2_prefetch") && (value == 0 || value == 1)) {
    g_options.l2_prefetch = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "managed_prefetch") && value >= 0 && value <= 2) {
    g_options.managed_prefetch = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "managed_advise") && (value == 0 || value == 1)) {
    g_options.managed_advise = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "cg_batch") && value >= 1 && value <= 4096) {
    g_options.cg_batch = (int)value;
    return CFS_OK;
  }
  if (!strcmp(key, "diag_mode") && value >= 0 && value <= 3) {
    g_options.diag_mode = (int)value;
    return CFS_OK;
  }
  set_error("cfs_cuda_set_option: unknown key or bad value: %s=%lld", key,
            value);
  return CFS_ERR_INVALID;
}

int cfs_cuda_device_count(int *count) {
  if (!count)
    return CFS_ERR_INVALID;
  *count = 0;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
  }
  return CFS_OK;
}

int cfs_cuda_init(int device) {
  int n = 0;
  CFS_TRY(cfs_cuda_device_count(&n));
  if (n == 0) {
    set_error("no CUDA device visible: the B200 path has no CPU fallback");
    return CFS_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (0..%d)", device, n - 1);
    return CFS_ERR_INVALID;
  }
  CFS_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CFS_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only",
              device, prop.major, prop.minor);
    return CFS_ERR_NO_DEVICE;
  }
  g_device = device;
  // CFS_GPU_OPTIONS="key=value,key=value": cfs_cuda_set_option for callers
  // that only have the environment (the reference's own bench / test binaries)
  static bool env_options_read = false;
  if (!env_options_read) {
    env_options_read = true;
    if (const char *env = getenv("CFS_GPU_OPTIONS")) {
      std::string all(env);
      size_t at = 0;
      while (at < all.size()) {
        size_t end = all.find(',', at);
        if (end == std::string::npos)
          end = all.size();
        const std::string item = all.substr(at, end - at);
        const size_t eq = item.find('=');
        if (eq != std::string::npos &&
            cfs_cuda_set_option(item.substr(0, eq).c_str(),
                                atoll(item.c_str() + eq + 1)) != CFS_OK)
          fprintf(stderr, "[cfs] CFS_GPU_OPTIONS: ignored '%s'\n",
                  item.c_str());
        at = end + 1;
      }
    }
  }
  return CFS_OK;
}

// CFS_GPU_ALLOC=managed|pinned|plain (cfs_cuda.h); read once
static int default_alloc_kind() {
  static int kind = 0;
  if (kind)
    return kind;
  kind = CFS_ALLOC_MANAGED;
  if (const char *env = getenv("CFS_GPU_ALLOC")) {
    if (!strcmp(env, "pinned"))
      kind = CFS_ALLOC_PINNED;
    else if (!strcmp(env, "plain"))
      kind = CFS_ALLOC_PLAIN;
    else if (strcmp(env, "managed"))
      fprintf(stderr, "[cfs] CFS_GPU_ALLOC: unknown kind '%s', using managed\n",
              env);
  }
  return kind;
}

// The allocator can be the first thing a process calls (bench_spmv_mmf.cpp:127
// comes before any matrix work in some callers): bind the GPU the process was
// told to use, not GPU 0, before the first CUDA allocation creates a context.
static bool bind_for_alloc() {
  if (g_device >= 0)
    return cudaSetDevice(g_device) == cudaSuccess;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return false;
  }
  int dev = 0;
  if (const char *env = getenv("CFS_GPU_DEVICE")) {
    const int v = atoi(env);
    if (v >= 0 && v < n)
      dev = v;
  }
  return cfs_cuda_init(dev) == CFS_OK;
}

void *cfs_cuda_host_alloc_kind(size_t bytes, int kind) {
  void *p = nullptr;
  const size_t want = bytes ? bytes : 64;
  if (kind == CFS_ALLOC_DEFAULT) {
    kind = default_alloc_kind();
    // small buffers gain nothing from living in HBM (a 64 KB vector is copied
    // in 2 us) and unified allocations come in 2 MB steps of address space
    if (kind == CFS_ALLOC_MANAGED && want < (64u << 10))
      kind = CFS_ALLOC_PLAIN;
  }
  if (kind != CFS_ALLOC_PLAIN && bind_for_alloc()) {
    if (kind == CFS_ALLOC_MANAGED) {
      if (cudaMallocManaged(&p, want, cudaMemAttachGlobal) == cudaSuccess) {
        // the vectors belong in HBM whenever the GPU works on them
        if (g_options.managed_advise)
          cudaMemAdvise(p, want, cudaMemAdviseSetPreferredLocation, g_device);
        cudaGetLastError();
        std::lock_guard<std::mutex> hold(g_alloc_mutex);
        g_allocs[p] = AllocRec{CFS_ALLOC_MANAGED, want, false};
        return p;
      }
    } else if (cudaHostAlloc(&p, want, cudaHostAllocPortable) == cudaSuccess) {
      std::lock_guard<std::mutex> hold(g_alloc_mutex);
      g_allocs[p] = AllocRec{CFS_ALLOC_PINNED, want, false};
      return p;
    }
  }
  cudaGetLastError();
  // no GPU (MMF loading on a host-only box) or arrays that are uploaded once:
  // plain 64-byte aligned memory, exactly what the reference returns
  // (allocator.cpp:33)
  if (posix_memalign(&p, 64, want) != 0)
    return nullptr;
  std::lock_guard<std::mutex> hold(g_alloc_mutex);
  g_allocs[p] = AllocRec{CFS_ALLOC_PLAIN, want, false};
  return p;
}

void *cfs_cuda_host_alloc(size_t bytes) {
  return cfs_cuda_host_alloc_kind(bytes, CFS_ALLOC_DEFAULT);
}

void cfs_cuda_host_free(void *ptr) {
  if (!ptr)
    return;
  int kind = CFS_ALLOC_PLAIN;
  {
    std::lock_guard<std::mutex> hold(g_alloc_mutex);
    auto it = g_allocs.find(ptr);
    if (it == g_allocs.end()) {
      free(ptr); // not ours: behave like the reference's free()
      return;
    }
    kind = it->second.kind;
    g_allocs.erase(it);
  }
  if (kind == CFS_ALLOC_PINNED)
    cudaFreeHost(ptr);
  else if (kind == CFS_ALLOC_MANAGED)
    cudaFree(ptr);
  else
    free(ptr);
}

int cfs_cuda_vector_prefetch(const void *ptr, size_t bytes, int to_device) {
  if (!ptr)
    return CFS_ERR_INVALID;
  PtrKind k;
  bool managed = false;
  classify(ptr, &k, nullptr, &managed);
  if (!managed || bytes == 0)
    return CFS_OK;
  CFS_TRY(require_device());
  CFS_CUDA_TRY(cudaMemPrefetchAsync(ptr, bytes,
                                    to_device ? g_device : cudaCpuDeviceId, 0));
  CFS_CUDA_TRY(cudaStreamSynchronize(0));
  if (!to_device) {
    std::lock_guard<std::mutex> hold(g_alloc_mutex);
    auto it = g_allocs.find(ptr);
    if (it != g_allocs.end())
      it->second.seen = false;
  }
  return CFS_OK;
}

static int create_common(cfs_mat_t *out, int32_t nrows, int32_t ncols,
                         int32_t row_begin, int32_t global_nrows, bool sharded,
                         const int32_t *rowptr, const int32_t *colind,
                         const void *values, int is_double, int symmetric) {
  if (!out || !rowptr || nrows < 0 || ncols < 0) {
    set_error("cfs_cuda_matrix_create: bad arguments");
    return CFS_ERR_INVALID;
  }
  CFS_TRY(require_device());
  cfs_matrix_s *m = new cfs_matrix_s;
  m->device = g_device;
  m->is_double = is_double != 0;
  m->symmetric = symmetric != 0;
  m->nrows = nrows;
  m->ncols = ncols;
  m->row_begin = row_begin;
  m->global_nrows = global_nrows;
  m->sharded = sharded;
  int status = CFS_OK;
  do {
    // rowptr first: its last entry is nnz (csr_matrix.tpp:122)
    DevArray<char> tmp;
    const void *p = nullptr;
    PtrKind k;
    bool managed = false;
    classify(rowptr, &k, nullptr, &managed);
    if (managed) // copied like a host array (see upload_or_borrow)
      k = kPtrHost;
    int32_t nnz = 0;
    if (k == kPtrDevice) {
      if (cudaMemcpy(&nnz, rowptr + nrows, 4, cudaMemcpyDeviceToHost) !=
          cudaSuccess) {
        status = cuda_fail(cudaGetLastError(), "read nnz", __FILE__, __LINE__);
        break;
      }
      m->csr_rowptr = rowptr;
    } else {
      nnz = rowptr[nrows];
      if ((status = m->own_rowptr.alloc((size_t)nrows + 1)) != CFS_OK)
        break;
      if (cudaMemcpy(m->own_rowptr.p, rowptr, ((size_t)nrows + 1) * 4,
                     cudaMemcpyHostToDevice) != cudaSuccess) {
        status = cuda_fail(cudaGetLastError(), "upload rowptr", __FILE__,
                           __LINE__);
        break;
      }
      m->csr_rowptr = m->own_rowptr.p;
    }
    m->nnz_full = nnz;
    if (nnz > 0 && (!colind || !values)) {
      set_error("cfs_cuda_matrix_create: null colind/values");
      status = CFS_ERR_INVALID;
      break;
    }
    classify(colind, &k, nullptr, &managed);
    if (managed)
      k = kPtrHost;
    if (k == kPtrDevice || nnz == 0) {
      m->csr_colind = colind;
    } else {
      if ((status = m->own_colind.alloc((size_t)nnz)) != CFS_OK)
        break;
      if (cudaMemcpy(m->own_colind.p, colind, (size_t)nnz * 4,
                     cudaMemcpyHostToDevice) != cudaSuccess) {
        status = cuda_fail(cudaGetLastError(), "upload colind", __FILE__,
                           __LINE__);
        break;
      }
      m->csr_colind = m->own_colind.p;
    }
    if (nnz == 0)
      m->csr_values = values;
    else if ((status = upload_or_borrow(values, (size_t)nnz * m->vsize(),
                                        m->own_values, &p)) != CFS_OK)
      break;
    else
      m->csr_values = p;
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) !=
        cudaSuccess) {
      status = cuda_fail(cudaGetLastError(), "cudaStreamCreate", __FILE__,
                         __LINE__);
      break;
    }
  } while (0);
  if (status != CFS_OK) {
    delete m;
    return status;
  }
  *out = m;
  return CFS_OK;
}

int cfs_cuda_matrix_create(cfs_mat_t *out, int32_t nrows, int32_t ncols,
                           const int32_t *rowptr, const int32_t *colind,
                           const void *values, int is_double, int symmetric) {
  return create_common(out, nrows, ncols, 0, nrows, false, rowptr, colind,
                       values, is_double, symmetric);
}

int cfs_cuda_matrix_create_shard(cfs_mat_t *out, int32_t global_nrows,
                                 int32_t row_begin, int32_t row_end,
                                 const int32_t *rowptr, const int32_t *colind,
                                 const void *values, int is_double) {
  if (row_begin < 0 || row_end < row_begin || row_end > global_nrows) {
    set_error("cfs_cuda_matrix_create_shard: bad row range");
    return CFS_ERR_INVALID;
  }
  return create_common(out, row_end - row_begin, global_nrows, row_begin,
                       global_nrows, true, rowptr, colind, values, is_double,
                       1);
}

int cfs_cuda_matrix_set_hybrid(cfs_mat_t m, int32_t bandwidth_threshold) {
  if (!m || bandwidth_threshold < 0)
    return CFS_ERR_INVALID;
  if (m->tuned) {
    set_error("cfs_cuda_matrix_set_hybrid: call it before cfs_cuda_matrix_tune");
    return CFS_ERR_STATE;
  }
  if (!m->symmetric || m->sharded) {
    set_error("cfs_cuda_matrix_set_hybrid: an unsharded symmetric matrix only "
              "(csr_matrix.tpp:13-15: a general file quietly gives plain CSR)");
    return CFS_ERR_INVALID;
  }
  m->hyb_threshold = bandwidth_threshold;
  return CFS_OK;
}

int cfs_cuda_matrix_tune(cfs_mat_t m, int nparts, int tuning) {
  // Tuning only selects the partitioner of a NON-symmetric matrix (:250-254)
  if (!m)
    return CFS_ERR_INVALID;
  if (m->tuned) {
    set_error("cfs_cuda_matrix_tune: matrix already tuned");
    return CFS_ERR_STATE;
  }
  CFS_CUDA_TRY(cudaSetDevice(m->device));
  if (nparts < 1)
    nparts = 1;
  m->nparts = nparts;
  int meta_status = CFS_OK;
  if (!m->symmetric) {
    // plain CSR (cpu_mv): keep a device copy that outlives the caller's arrays
    if (!m->own_rowptr.p) {
      CFS_TRY(m->own_rowptr.alloc((size_t)m->nrows + 1));
      CFS_CUDA_TRY(cudaMemcpy(m->own_rowptr.p, m->csr_rowptr,
                              ((size_t)m->nrows + 1) * 4,
                              cudaMemcpyDeviceToDevice));
      m->csr_rowptr = m->own_rowptr.p;
    }
    if (!m->own_colind.p && m->nnz_full) {
      CFS_TRY(m->own_colind.alloc((size_t)m->nnz_full));
      CFS_CUDA_TRY(cudaMemcpy(m->own_colind.p, m->csr_colind,
                              (size_t)m->nnz_full * 4,
                              cudaMemcpyDeviceToDevice));
      m->csr_colind = m->own_colind.p;
    }
    if (!m->own_values.p && m->nnz_full) {
      CFS_TRY(m->own_values.alloc((size_t)m->nnz_full * m->vsize()));
      CFS_CUDA_TRY(cudaMemcpy(m->own_values.p, m->csr_values,
                              (size_t)m->nnz_full * m->vsize(),
                              cudaMemcpyDeviceToDevice));
      m->csr_values = m->own_values.p;
    }
    CFS_TRY(tune_csr(m, nparts, tuning, m->stream));
  } else {
    if (!m->sharded && m->nrows != m->ncols) {
      set_error("symmetric matrix must be square");
      return CFS_ERR_INVALID;
    }
    // CFS_GPU_TUNE_REPORT=1: wall time of every preprocessing step on stderr
    const bool report = getenv("CFS_GPU_TUNE_REPORT") != nullptr;
    auto now = [&]() {
      if (report)
        cudaStreamSynchronize(m->stream);
      timespec ts;
      clock_gettime(CLOCK_MONOTONIC, &ts);
      return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    };
    // partition count first: an invalid one must not cost a layout build
    if (nparts > 1) {
      if (m->sharded) {
        set_error("reference-compatible metadata (nparts > 1) is defined on "
                  "the unsharded matrix only");
        return CFS_ERR_INVALID;
      }
      const long long S = ((m->nrows / nparts - 1) | (kBlkFactor - 1)) + 1;
      if (m->nrows / nparts < 1 || (long long)(nparts - 1) * S > m->nrows) {
        // the reference overshoots row_split_ and crashes here
        set_error("nparts=%d is not a valid partition count for %d rows",
                  nparts, m->nrows);
        return CFS_ERR_INVALID;
      }
    }
    // Layouts of kernels that are not the selected one (variants 1 / 2 / 4: the
    // uncompressed column stream, window slots, tile records) are built and
    // kept only for small matrices -- tests switch variants after tune -- or on
    // request (option keep_layouts): on BASELINE configs[1] they are 0.85 GB
    // of HBM and 15 % of the tune time for kernels that lost by 1.5-2.8x.
    const int v_sel = g_options.spmv_variant;
    const bool keep_all =
        g_options.keep_layouts || m->nnz_full < (4ll << 20);
    const bool want_windows = keep_all || v_sel == 2 || v_sel == 3 || v_sel == 4;
    double t0 = now();
    auto lap = [&](const char *what) {
      if (!report)
        return;
      const double t1 = now();
      fprintf(stderr, "[tune] %-22s %8.2f ms\n", what, t1 - t0);
      t0 = t1;
    };
    if (m->hyb_threshold > 0) { // Format::hyb: near part here, far part aside
      CFS_TRY(split_hybrid(m, m->stream));
      lap("hybrid split");
    }
    CFS_TRY(build_lower(m, m->stream));
    lap("lower triangle");
    CFS_TRY(build_layout(m, m->stream));
    lap("sliced layout");
    if (want_windows) {
      CFS_TRY(build_windows(m, m->stream));
      lap("windows (variant 4)");
    }
    CFS_TRY(build_compressed_cols(m, m->stream));
    lap("index compression");
    CFS_TRY(build_tiles6(m, m->stream));
    lap("transposed tiles");
    CFS_TRY(build_value_index(m, m->stream));
    lap("value index");
    CFS_TRY(build_hubs(m, m->stream));
    lap("hub columns");
    CFS_TRY(build_pipeline_reach(m, m->stream));
    lap("pipeline reach");
    CFS_TRY(build_halo_extent(m, m->stream));
    // row_split_ (partition_by_nrows, csr_matrix.tpp:418-423)
    m->row_split.assign((size_t)nparts + 1, 0);
    if (nparts > 1) {
      const long long S = ((m->nrows / nparts - 1) | (kBlkFactor - 1)) + 1;
      for (int t = 0; t < nparts; ++t)
        m->row_split[t] = (int32_t)(t * S);
      m->row_split[nparts] = m->nrows;
      meta_status = build_refmeta(m, m->stream);
      if (meta_status != CFS_OK && meta_status != CFS_ERR_TOO_LARGE)
        return meta_status;
    } else {
      m->row_split[1] = m->nrows;
    }
    if (!keep_all) {
      // the P = 1 lower CSR (the reference's SymThreadData content) has done
      // its work: layout, hub columns and metadata are built. Its row pointer
      // (4 bytes per row) stays for the per-partition counts.
      m->low_colind.release();
      m->low_values.release();
      if (!want_windows) {
        m->tile_rec.release();
        m->ntiles = 0;
      }
      // regular matrices run from the compressed column stream
      if (v_sel != 1 && !want_windows && m->ccol.p &&
          m->nregular * 8 >= m->nslices && m->nhubs == 0)
        m->sell_col.release();
    }
    // compress_symmetry() frees the full CSR when it owns it (:1700-1706)
    m->own_rowptr.release();
    m->own_colind.release();
    m->own_values.release();
    m->csr_rowptr = m->csr_colind = nullptr;
    m->csr_values = nullptr;
  }
  m->tuned = true;
  return meta_status;
}

void cfs_cuda_matrix_destroy(cfs_mat_t m) {
  if (!m)
    return;
  cudaSetDevice(m->device);
  for (cudaEvent_t e : m->ev_x)
    cudaEventDestroy(e);
  for (cudaEvent_t e : m->ev_k)
    cudaEventDestroy(e);
  for (cudaEvent_t e : m->ev_d)
    cudaEventDestroy(e);
  if (m->ev_fork)
    cudaEventDestroy(m->ev_fork);
  if (m->ev_join)
    cudaEventDestroy(m->ev_join);
  if (m->pipe_graph)
    cudaGraphExecDestroy(m->pipe_graph);
  if (m->h2d_stream)
    cudaStreamDestroy(m->h2d_stream);
  if (m->d2h_stream)
    cudaStreamDestroy(m->d2h_stream);
  if (m->stream)
    cudaStreamDestroy(m->stream);
  if (m->far)
    cfs_cuda_matrix_destroy(m->far);
  delete m;
}

int cfs_cuda_matrix_info(cfs_mat_t m, cfs_matrix_info *info) {
  if (!m || !info)
    return CFS_ERR_INVALID;
  memset(info, 0, sizeof(*info));
  const int64_t vs = (int64_t)m->vsize();
  info->nrows = m->nrows;
  info->ncols = m->ncols;
  info->row_begin = m->row_begin;
  info->halo_begin = m->halo_begin;
  info->nnz_full = m->nnz_full;
  info->nnz_low = m->nnz_low;
  info->nnz_diag = m->nnz_diag;
  info->nparts = m->nparts;
  info->ncolors = m->ncolors;
  info->nranges = m->nranges;
  info->symmetric = m->symmetric;
  info->is_double = m->is_double;
  info->tuned = m->tuned;
  info->refmeta = m->refmeta;
  info->nvrows = m->nvrows;
  info->nslices = m->nslices;
  info->padded_entries = m->padded_entries;
  info->nconflict_edges = m->nedges;
  info->ntiles = m->ntiles;
  info->far_entries = m->far_entries;
  info->regular_slices = m->nregular;
  info->hub_columns = m->nhubs;
  info->hub_entries = m->hub_entries;
  info->hyb_far_entries = m->far ? m->hyb_far_entries : 0;
  info->sort_window = m->sort_window;
  info->index_rows = m->ccol_rows;
  info->value_dictionary = m->ndict;
  info->transposed_tiles = m->nt6;
  info->tile_smem_bytes = (int64_t)((m->t6_smem_entries + 7) & ~7) * vs +
                          (m->nt6 ? (m->t6_max_cols + 2) * 2 : 0);
  if (m->symmetric && m->tuned) {
    // size(), csr_matrix.tpp:191-228 (including its (nrows + 1*nthreads) term)
    int64_t s = ((int64_t)m->nrows + 1LL * m->nparts) * 4;
    s += m->nnz_low * 4 + m->nnz_low * vs + m->nnz_diag * vs;
    if (m->nparts > 1) {
      s += ((int64_t)m->ncolors + 1) * 4;
      s += 2LL * m->nranges * 4;
    }
    if (m->far) // size(), csr_matrix.tpp:211-216: colind_high + values_high
      s += m->hyb_far_entries * (4 + vs);
    info->size_bytes = s;
    // SURVEY.md 8(d)
    info->algorithmic_bytes = m->nnz_low * (vs + 4) + (int64_t)m->nrows * vs +
                              ((int64_t)m->nrows + 1) * 4 +
                              2 * (int64_t)m->nrows * vs;
    if (m->far) // + the far part's CSR
      info->algorithmic_bytes +=
          m->hyb_far_entries * (vs + 4) + ((int64_t)m->nrows + 1) * 4;
  } else {
    info->size_bytes = ((int64_t)m->nrows + 1) * 4 + m->nnz_full * (4 + vs);
    if (m->part_by_nnz)
      info->size_bytes += ((int64_t)m->nparts + 1) * 4; // row_split_, :223-224
    info->algorithmic_bytes = info->size_bytes + 2 * (int64_t)m->nrows * vs;
  }
  info->device_bytes =
      (int64_t)(m->own_rowptr.bytes() + m->own_colind.bytes() +
                m->own_values.bytes() + m->low_rowptr.bytes() +
                m->low_colind.bytes() + m->low_values.bytes() +
                m->diagonal.bytes() + m->slice_ptr.bytes() +
                m->vrow_row.bytes() + m->sell_col.bytes() +
                m->sell_val.bytes() + m->tile_rec.bytes() + m->hub_ptr.bytes() + m->hub_row.bytes() +
                m->hub_val.bytes() + m->hub_colstream.bytes() +
                m->hub_chunks.bytes() + m->vcode.bytes() + m->vdict.bytes() +
                m->t6_pack.bytes() +
                m->t6_cptr.bytes() + m->t6_lo.bytes() + m->t6_ncols.bytes() +
                m->t6_cptr_off.bytes() + m->sell_slot.bytes() + m->ccol.bytes() + m->slice_cptr.bytes() + m->weight.bytes() + m->adj_ptr.bytes() +
                m->adj.bytes() + m->color.bytes() + m->color_first.bytes() +
                m->range_ptr.bytes() + m->part_nranges.bytes() +
                m->range_start.bytes() + m->range_end.bytes() +
                m->stage_x.bytes() + m->stage_y.bytes() +
                m->reach_min.bytes() + m->reach_rlo.bytes() +
                m->reach_rhi.bytes());
  if (m->far) {
    cfs_matrix_info fi;
    cfs_cuda_matrix_info(m->far, &fi);
    info->device_bytes += fi.device_bytes;
  }
  return CFS_OK;
}

int cfs_cuda_spmv_async(cfs_mat_t m, void *y_dev, const void *x_dev,
                        void *stream) {
  if (!m || !y_dev || !x_dev)
    return CFS_ERR_INVALID;
  if (!m->tuned) {
    set_error("cfs_cuda_spmv: call cfs_cuda_matrix_tune first");
    return CFS_ERR_STATE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (m->symmetric)
    return launch_sym_spmv(m, y_dev, x_dev, s);
  return launch_csr_spmv(m, y_dev, x_dev, s);
}

int cfs_cuda_spmv_halo_async(cfs_mat_t m, void *y_dev, const void *x_dev,
                             void *y_lower_base, int y_is_zero, void *stream) {
  if (!m || !y_dev || !x_dev)
    return CFS_ERR_INVALID;
  if (!m->tuned || !m->symmetric) {
    set_error("cfs_cuda_spmv_halo_async: needs a tuned symmetric matrix");
    return CFS_ERR_STATE;
  }
  return launch_sym_spmv(m, y_dev, x_dev, (cudaStream_t)stream, nullptr,
                         nullptr, y_lower_base, y_is_zero != 0);
}

int cfs_cuda_spmv_shard_async(cfs_mat_t m, void *y_dev, const void *x_dev,
                              void *y_lower_base, const void *x_lower_base,
                              void *y_clear, int y_is_zero, void *stream) {
  if (!m || !y_dev || !x_dev)
    return CFS_ERR_INVALID;
  if (!m->tuned || !m->symmetric) {
    set_error("cfs_cuda_spmv_shard_async: needs a tuned symmetric matrix");
    return CFS_ERR_STATE;
  }
  if (x_lower_base && !y_lower_base) {
    set_error("cfs_cuda_spmv_shard_async: x_lower_base needs y_lower_base (the "
              "halo kernels read and reduce through the same test)");
    return CFS_ERR_INVALID;
  }
  return launch_sym_spmv(m, y_dev, x_dev, (cudaStream_t)stream, nullptr,
                         nullptr, y_lower_base, y_is_zero != 0, 0, -1, nullptr,
                         x_lower_base, y_clear);
}

int cfs_cuda_spmv_shard_part_async(cfs_mat_t m, void *y_dev, const void *x_dev,
                                   void *y_lower_base, const void *x_lower_base,
                                   void *y_clear, int part, void *stream) {
  if (!m || !y_dev || !x_dev || part < 1 || part > 2)
    return CFS_ERR_INVALID;
  if (!m->tuned || !m->symmetric) {
    set_error("cfs_cuda_spmv_shard_part_async: needs a tuned symmetric matrix");
    return CFS_ERR_STATE;
  }
  const long long hs = m->halo_slice_end < m->nslices ? m->halo_slice_end
                                                      : m->nslices;
  if (part == 1) { // the slices that reach below row_begin
    if (hs == 0 || !y_lower_base)
      return CFS_OK;
    return launch_sym_spmv(m, y_dev, x_dev, (cudaStream_t)stream, nullptr,
                           nullptr, y_lower_base, true, 0, hs, nullptr,
                           x_lower_base, y_clear);
  }
  const long long from = y_lower_base ? hs : 0; // no halo: everything is interior
  if (from >= m->nslices)
    return CFS_OK;
  return launch_sym_spmv(m, y_dev, x_dev, (cudaStream_t)stream, nullptr,
                         nullptr, nullptr, true, from, m->nslices, nullptr,
                         nullptr, y_clear);
}

// Host vectors in, host vectors out: H2D of x, the kernel and D2H of y overlap
// chunk by chunk (see cfs_matrix_s::Chunk). Three streams, events between them.
// enqueue_pipeline issues one whole step; with pinned vectors the step is
// captured ONCE per (x, y) pair into a CUDA graph (about 110 runtime calls
// become one launch -- the bench loop of the reference, bench_spmv_mmf.cpp:
// 162-167, calls with the same two vectors every time).
static int enqueue_pipeline(cfs_mat_t m, void *y, const void *x, bool fork) {
  const size_t vs = m->vsize();
  const size_t S = m->stages.size();
  char *xd = m->stage_x.p, *yd = m->stage_y.p;
  const bool trace = g_options.pipeline_trace && !fork;
  if (fork) { // bring the copy streams into the capture
    CFS_CUDA_TRY(cudaEventRecord(m->ev_fork, m->stream));
    CFS_CUDA_TRY(cudaStreamWaitEvent(m->h2d_stream, m->ev_fork, 0));
  }
  if (trace)
    CFS_CUDA_TRY(cudaEventRecord(m->ev_fork, m->stream));
  CFS_CUDA_TRY(cudaMemsetAsync(yd, 0, (size_t)m->nrows * vs, m->stream));
  for (size_t j = 0; j < S; ++j) {
    const cfs_matrix_s::Stage &st = m->stages[j];
    const size_t xo = (size_t)st.x_row0 * vs, xe = (size_t)st.x_row1 * vs;
    if (xe > xo && !(g_options.pipeline_skip & 4))
      CFS_CUDA_TRY(cudaMemcpyAsync(xd + xo, (const char *)x + xo, xe - xo,
                                   cudaMemcpyHostToDevice, m->h2d_stream));
    CFS_CUDA_TRY(cudaEventRecord(m->ev_x[j], m->h2d_stream));
    CFS_CUDA_TRY(cudaStreamWaitEvent(m->stream, m->ev_x[j], 0));
    if (!(g_options.pipeline_skip & 1))
      CFS_TRY(launch_sym_spmv(m, yd, xd, m->stream, nullptr, nullptr, nullptr,
                              true, st.slice0, st.slice1));
    CFS_CUDA_TRY(cudaEventRecord(m->ev_k[j], m->stream));
    if (!st.y_ready.empty())
      CFS_CUDA_TRY(cudaStreamWaitEvent(m->d2h_stream, m->ev_k[j], 0));
    for (const std::pair<int, int> &r : st.y_ready) {
      const size_t yo = (size_t)r.first * vs, ye = (size_t)r.second * vs;
      if (ye > yo && !(g_options.pipeline_skip & 2))
        CFS_CUDA_TRY(cudaMemcpyAsync((char *)y + yo, yd + yo, ye - yo,
                                     cudaMemcpyDeviceToHost, m->d2h_stream));
    }
    if (trace)
      CFS_CUDA_TRY(cudaEventRecord(m->ev_d[j], m->d2h_stream));
  }
  if (fork) { // join: the capture ends on m->stream
    CFS_CUDA_TRY(cudaEventRecord(m->ev_join, m->d2h_stream));
    CFS_CUDA_TRY(cudaStreamWaitEvent(m->stream, m->ev_join, 0));
  }
  return CFS_OK;
}

static int spmv_host_pipelined(cfs_mat_t m, void *y, const void *x,
                               bool pinned) {
  const size_t vs = m->vsize();
  const size_t K = m->stages.size();
  if (!m->h2d_stream) {
    CFS_CUDA_TRY(cudaStreamCreateWithFlags(&m->h2d_stream,
                                           cudaStreamNonBlocking));
    CFS_CUDA_TRY(cudaStreamCreateWithFlags(&m->d2h_stream,
                                           cudaStreamNonBlocking));
    m->ev_x.resize(K);
    m->ev_k.resize(K);
    m->ev_d.resize(K);
    const unsigned flags =
        g_options.pipeline_trace ? cudaEventDefault : cudaEventDisableTiming;
    for (size_t c = 0; c < K; ++c) {
      CFS_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_x[c], flags));
      CFS_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_k[c], flags));
      CFS_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_d[c], flags));
    }
    CFS_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_fork, flags));
    CFS_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
  }
  if (!m->stage_x.p)
    CFS_TRY(m->stage_x.alloc((size_t)m->ncols * vs));
  if (!m->stage_y.p)
    CFS_TRY(m->stage_y.alloc((size_t)m->nrows * vs));
  if (pinned && g_options.pipeline_graph) {
    if (!m->pipe_graph || m->pipe_x != x || m->pipe_y != y ||
        m->pipe_generation != g_options_generation) {
      if (m->pipe_graph) {
        cudaGraphExecDestroy(m->pipe_graph);
        m->pipe_graph = nullptr;
      }
      cudaGraph_t graph = nullptr;
      CFS_CUDA_TRY(cudaStreamBeginCapture(m->stream,
                                          cudaStreamCaptureModeThreadLocal));
      const int status = enqueue_pipeline(m, y, x, true);
      const cudaError_t e = cudaStreamEndCapture(m->stream, &graph);
      if (status != CFS_OK) {
        if (graph)
          cudaGraphDestroy(graph);
        return status;
      }
      CFS_CUDA_TRY(e);
      const cudaError_t ei = cudaGraphInstantiate(&m->pipe_graph, graph, 0);
      cudaGraphDestroy(graph);
      CFS_CUDA_TRY(ei);
      m->pipe_x = x;
      m->pipe_y = y;
      m->pipe_generation = g_options_generation;
    }
    CFS_CUDA_TRY(cudaGraphLaunch(m->pipe_graph, m->stream));
    CFS_CUDA_TRY(cudaStreamSynchronize(m->stream));
    return CFS_OK;
  }
  CFS_TRY(enqueue_pipeline(m, y, x, false));
  CFS_CUDA_TRY(cudaStreamSynchronize(m->d2h_stream));
  CFS_CUDA_TRY(cudaStreamSynchronize(m->stream));
  if (g_options.pipeline_trace) { // development aid: where the step's time goes
    for (size_t c = 0; c < K; ++c) {
      float tx = 0, tk = 0, td = 0;
      cudaEventElapsedTime(&tx, m->ev_fork, m->ev_x[c]);
      cudaEventElapsedTime(&tk, m->ev_fork, m->ev_k[c]);
      cudaEventElapsedTime(&td, m->ev_fork, m->ev_d[c]);
      printf("pipeline stage %2zu (slices %lld..%lld): x rows %d..%d in at "
             "%.3f ms, kernel done %.3f, %zu y range(s) out by %.3f\n", c,
             m->stages[c].slice0, m->stages[c].slice1, m->stages[c].x_row0,
             m->stages[c].x_row1, tx, tk, m->stages[c].y_ready.size(), td);
    }
    fflush(stdout);
  }
  return CFS_OK;
}

// Unified-memory vectors (cfs_cuda_host_alloc, kind managed): the kernel runs
// straight on them. A vector the host has just written sits in host pages; one
// bulk prefetch moves it 3x faster than a page fault per 64 KB block, but a
// prefetch of pages that are resident already costs 55-95 us (measured,
// tools/managed_probe.cu) -- more than the whole SpMV of BASELINE configs[0].
// Policy (option managed_prefetch):
//   0 = never;
//   1 = (default) when this library has not yet brought the allocation to the
//       GPU, and for a while after a call that evidently ran into page faults
//       (see cfs_cuda_spmv): the host loop of the reference never touches the
//       vectors between calls (bench_spmv_mmf.cpp:162-167) and pays nothing, a
//       caller that rewrites x between calls gets bulk transfers;
//   2 = on every call.
// Correct in all three: pages the host touched come over by GPU page faults.
static int managed_to_device(const void *p, size_t bytes, cudaStream_t s,
                             bool force, bool *moved) {
  const int mode = g_options.managed_prefetch;
  if (mode == 0 || bytes == 0)
    return CFS_OK;
  if (mode == 1) {
    std::lock_guard<std::mutex> hold(g_alloc_mutex);
    auto it = g_allocs.find(p);
    if (it != g_allocs.end()) {
      if (it->second.seen && !force)
        return CFS_OK;
      it->second.seen = true;
    }
  }
  CFS_CUDA_TRY(cudaMemPrefetchAsync(p, bytes, g_device, s));
  *moved = true;
  return CFS_OK;
}

static double wall_us() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

int cfs_cuda_spmv(cfs_mat_t m, void *y, const void *x) {
  if (!m || !y || !x)
    return CFS_ERR_INVALID;
  if (!m->tuned) {
    set_error("cfs_cuda_spmv: call cfs_cuda_matrix_tune first");
    return CFS_ERR_STATE;
  }
  CFS_CUDA_TRY(cudaSetDevice(m->device));
  const size_t vs = m->vsize();
  const size_t xlen = m->sharded
                          ? (size_t)(m->row_begin + m->nrows - m->halo_begin)
                          : (size_t)m->ncols;
  const size_t ylen = m->sharded ? xlen : (size_t)m->nrows;
  PtrKind kx, ky;
  bool px = false, py = false, mx = false, my = false;
  classify(x, &kx, &px, &mx);
  classify(y, &ky, &py, &my);
  if (kx == kPtrHost && ky == kPtrHost && m->symmetric && m->plan_state == 1 &&
      g_options.pipeline && !g_options.deterministic)
    CFS_TRY(build_pipeline_plan(m, m->stream)); // first call with host vectors
  if (kx == kPtrHost && ky == kPtrHost && m->symmetric && !m->stages.empty() &&
      g_options.pipeline && !g_options.deterministic)
    return spmv_host_pipelined(m, y, x, px && py);
  // a call on unified memory that took far longer than the matrix can explain
  // ran into page faults: the host touches the vectors between calls, so the
  // next `managed_hot` calls prefetch; when they are used up one call goes
  // without, and if that faults again the window doubles
  const bool hot = m->managed_hot > 0;
  const double t_call = (mx || my) ? wall_us() : 0.0;
  bool prefetched = false;
  if (mx)
    CFS_TRY(managed_to_device(x, xlen * vs, m->stream, hot, &prefetched));
  if (my)
    CFS_TRY(managed_to_device(y, ylen * vs, m->stream, hot, &prefetched));
  const void *xd = x;
  void *yd = y;
  if (kx == kPtrHost) {
    if (!m->stage_x.p)
      CFS_TRY(m->stage_x.alloc(xlen * vs));
    CFS_CUDA_TRY(cudaMemcpyAsync(m->stage_x.p, x, xlen * vs,
                                 cudaMemcpyHostToDevice, m->stream));
    xd = m->stage_x.p;
  }
  if (ky == kPtrHost) {
    if (!m->stage_y.p)
      CFS_TRY(m->stage_y.alloc(ylen * vs));
    yd = m->stage_y.p;
  }
  CFS_TRY(cfs_cuda_spmv_async(m, yd, xd, m->stream));
  if (ky == kPtrHost)
    CFS_CUDA_TRY(cudaMemcpyAsync(y, yd, ylen * vs, cudaMemcpyDeviceToHost,
                                 m->stream));
  CFS_CUDA_TRY(cudaStreamSynchronize(m->stream));
  if ((mx || my) && g_options.managed_prefetch == 1) {
    if (hot) {
      --m->managed_hot;
    } else if (!prefetched) {
      // generous bound for a fault-free call: the algorithmic bytes at a
      // quarter of a slow kernel's rate plus launch and wake-up latency
      cfs_matrix_info info;
      cfs_cuda_matrix_info(m, &info);
      const double expect_us = 100.0 + 4.0 * (double)info.algorithmic_bytes / 3e6;
      if (wall_us() - t_call > expect_us) {
        m->managed_window = m->managed_window ? 2 * m->managed_window : 8;
        if (m->managed_window > 4096)
          m->managed_window = 4096;
        m->managed_hot = m->managed_window;
      } else {
        m->managed_window = 0;
      }
    }
  }
  return CFS_OK;
}

int cfs_cuda_spmv_timed(cfs_mat_t m, void *y_dev, const void *x_dev,
                        void *stream, int iters, float *total_ms,
                        float *kernel_ms) {
  if (!m || !y_dev || !x_dev || iters < 1 || iters > 65536)
    return CFS_ERR_INVALID;
  if (!m->tuned || !m->symmetric) {
    set_error("cfs_cuda_spmv_timed: needs a tuned symmetric matrix");
    return CFS_ERR_STATE;
  }
  CFS_CUDA_TRY(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  // Two events around the whole loop -- events around every single launch put
  // their own latency into a 150 us kernel (round 1 reported a kernel time
  // above the step time that way) -- then the same loop with the y
  // initialisation alone; the kernel's share is the difference.
  cudaEvent_t ev[4];
  for (auto &e : ev)
    CFS_CUDA_TRY(cudaEventCreate(&e));
  const size_t ext_bytes =
      (size_t)(m->row_begin + m->nrows - m->halo_begin) * m->vsize();
  int status = CFS_OK;
  cudaEventRecord(ev[0], s);
  for (int i = 0; i < iters && status == CFS_OK; ++i)
    status = launch_sym_spmv(m, y_dev, x_dev, s);
  cudaEventRecord(ev[1], s);
  cudaEventRecord(ev[2], s);
  for (int i = 0; i < iters; ++i)
    cudaMemsetAsync(y_dev, 0, ext_bytes, s);
  cudaEventRecord(ev[3], s);
  // leave y = A x behind, like every other entry point
  if (status == CFS_OK)
    status = launch_sym_spmv(m, y_dev, x_dev, s);
  if (status == CFS_OK && cudaStreamSynchronize(s) != cudaSuccess)
    status = cuda_fail(cudaGetLastError(), "sync", __FILE__, __LINE__);
  if (status == CFS_OK) {
    float t_all = 0, t_clear = 0;
    cudaEventElapsedTime(&t_all, ev[0], ev[1]);
    cudaEventElapsedTime(&t_clear, ev[2], ev[3]);
    if (total_ms)
      *total_ms = t_all;
    if (kernel_ms)
      *kernel_ms = t_all - t_clear;
  }
  for (auto &e : ev)
    cudaEventDestroy(e);
  return status;
}

// ---- metadata export ------------------------------------------------------
static int copy_out(const void *dev, size_t n, size_t elem, void *dst,
                    size_t cap, size_t *count) {
  if (count)
    *count = n;
  if (!dst)
    return CFS_OK;
  if (cap < n) {
    set_error("export buffer too small: need %zu elements", n);
    return CFS_ERR_INVALID;
  }
  if (n)
    CFS_CUDA_TRY(cudaMemcpy(dst, dev, n * elem, cudaMemcpyDeviceToHost));
  return CFS_OK;
}

int cfs_cuda_matrix_export(cfs_mat_t m, int what, void *dst, size_t cap,
                           size_t *count) {
  if (!m || !m->tuned || (!m->symmetric && what != CFS_META_ROW_SPLIT)) {
    set_error("cfs_cuda_matrix_export: needs a tuned symmetric matrix (a "
              "non-symmetric one only has row_split)");
    return CFS_ERR_STATE;
  }
  CFS_CUDA_TRY(cudaSetDevice(m->device));
  const size_t vs = m->vsize();
  const int P = m->nparts;
  const size_t N = (size_t)m->nrows;
  switch (what) {
  case CFS_META_ROW_SPLIT: {
    const size_t n = (size_t)P + 1;
    if (count)
      *count = n;
    if (dst) {
      if (cap < n)
        return CFS_ERR_INVALID;
      memcpy(dst, m->row_split.data(), n * 4);
    }
    return CFS_OK;
  }
  case CFS_META_PART_NNZ_LOW:
  case CFS_META_LOWER_ROWPTR: {
    // per-partition local rowptr_ (csr_matrix.tpp:1233-1234, 1294-1296) from
    // the global lower rowptr
    std::vector<int32_t> g(N + 1);
    if (N + 1)
      CFS_CUDA_TRY(cudaMemcpy(g.data(), m->low_rowptr.p, (N + 1) * 4,
                              cudaMemcpyDeviceToHost));
    std::vector<int32_t> outv;
    if (what == CFS_META_PART_NNZ_LOW) {
      for (int t = 0; t < P; ++t)
        outv.push_back(g[m->row_split[t + 1]] - g[m->row_split[t]]);
    } else {
      outv.reserve(N + P);
      for (int t = 0; t < P; ++t)
        for (int i = m->row_split[t]; i <= m->row_split[t + 1]; ++i)
          outv.push_back(g[i] - g[m->row_split[t]]);
    }
    if (count)
      *count = outv.size();
    if (dst) {
      if (cap < outv.size())
        return CFS_ERR_INVALID;
      memcpy(dst, outv.data(), outv.size() * 4);
    }
    return CFS_OK;
  }
  case CFS_META_LOWER_COLIND:
  case CFS_META_LOWER_VALUES:
    if (!m->low_colind.p && m->nnz_low) {
      set_error("the lower CSR was released after tune (the kernels read the "
                "sliced layout): set option keep_layouts=1 before "
                "cfs_cuda_matrix_tune to export it");
      return CFS_ERR_STATE;
    }
    if (what == CFS_META_LOWER_COLIND)
      return copy_out(m->low_colind.p, (size_t)m->nnz_low, 4, dst, cap, count);
    return copy_out(m->low_values.p, (size_t)m->nnz_low, vs, dst, cap, count);
  case CFS_META_DIAGONAL:
    return copy_out(m->diagonal.p, N, vs, dst, cap, count);
  case CFS_META_SELL_SLICE_PTR:
    return copy_out(m->slice_ptr.p, (size_t)m->nslices + 1, 4, dst, cap, count);
  case CFS_META_SELL_VROW:
    return copy_out(m->vrow_row.p, (size_t)m->nslices * kSliceRows, 4, dst, cap,
                    count);
  case CFS_META_SELL_COL:
    if (!m->sell_col.p && m->padded_entries) {
      set_error("the uncompressed column stream was released after tune (the "
                "selected kernel reads the compressed one): set option "
                "keep_layouts=1 before cfs_cuda_matrix_tune");
      return CFS_ERR_STATE;
    }
    return copy_out(m->sell_col.p, (size_t)m->padded_entries, 4, dst, cap,
                    count);
  case CFS_META_SELL_VAL:
    return copy_out(m->sell_val.p, (size_t)m->padded_entries, vs, dst, cap,
                    count);
  default:
    break;
  }
  if (!m->refmeta) {
    // P == 1 (serial(), csr_matrix.tpp:642) has no graph / colours / ranges
    if (count)
      *count = 0;
    return CFS_OK;
  }
  const size_t V = (size_t)m->nblk;
  switch (what) {
  case CFS_META_WEIGHT:
    return copy_out(m->weight.p, V, 4, dst, cap, count);
  case CFS_META_ADJ_PTR:
    return copy_out(m->adj_ptr.p, V + 1, 4, dst, cap, count);
  case CFS_META_ADJ:
    return copy_out(m->adj.p, (size_t)m->nedges, 4, dst, cap, count);
  case CFS_META_COLOR_FIRST:
    return copy_out(m->color_first.p, V, 4, dst, cap, count);
  case CFS_META_COLOR:
    return copy_out(m->color.p, V, 4, dst, cap, count);
  case CFS_META_RANGE_PTR:
    return copy_out(m->range_ptr.p, (size_t)P * (m->ncolors + 1), 4, dst, cap,
                    count);
  case CFS_META_PART_NRANGES:
    return copy_out(m->part_nranges.p, (size_t)P, 4, dst, cap, count);
  case CFS_META_RANGE_START:
    return copy_out(m->range_start.p, (size_t)m->nranges, 4, dst, cap, count);
  case CFS_META_RANGE_END:
    return copy_out(m->range_end.p, (size_t)m->nranges, 4, dst, cap, count);
  default:
    set_error("cfs_cuda_matrix_export: unknown selector %d", what);
    return CFS_ERR_INVALID;
  }
}

} // extern "C"
