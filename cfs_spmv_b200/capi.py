"""ctypes binding of the C ABI (include/cfs_cuda.h) -- the same entry points the
C++ drop-in classes call. Used by tests/, bench.py and __graft_entry__.

There is no fallback: if libcfs_cuda.so is missing this module raises, and
every compute call raises CfsError when no B200 is usable.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libcfs_cuda.so")

CFS_OK = 0
CFS_ERR_CUDA, CFS_ERR_INVALID, CFS_ERR_NO_DEVICE, CFS_ERR_STATE, \
    CFS_ERR_TOO_LARGE = 1, 2, 3, 4, 5
CFS_ERR_NEEDS_HOST = 6
CFS_ALLOC_DEFAULT, CFS_ALLOC_PLAIN, CFS_ALLOC_PINNED, CFS_ALLOC_MANAGED = 0, 1, 2, 3

META = {
    "row_split": 1, "part_nnz_low": 2, "lower_rowptr": 3, "lower_colind": 4,
    "lower_values": 5, "diagonal": 6, "weight": 7, "adj_ptr": 8, "adj": 9,
    "color_first": 10, "color": 11, "range_ptr": 12, "part_nranges": 13,
    "range_start": 14, "range_end": 15,
    "sell_slice_ptr": 100, "sell_vrow": 101, "sell_col": 102, "sell_val": 103,
}
_VALUE_META = ("lower_values", "diagonal", "sell_val")

# every symbol include/cfs_cuda.h declares (tests check the library exports all)
DECLARED_SYMBOLS = (
    "cfs_cuda_device_count", "cfs_cuda_init", "cfs_cuda_last_error",
    "cfs_cuda_version", "cfs_cuda_set_option", "cfs_cuda_host_alloc", "cfs_cuda_host_free",
    "cfs_cuda_host_alloc_kind", "cfs_cuda_vector_prefetch",
    "cfs_cuda_matrix_create", "cfs_cuda_matrix_create_shard",
    "cfs_cuda_matrix_create_from_mmf", "cfs_cuda_matrix_download_csr",
    "cfs_cuda_matrix_tune", "cfs_cuda_matrix_destroy", "cfs_cuda_matrix_info",
    "cfs_cuda_matrix_set_hybrid",
    "cfs_cuda_spmv", "cfs_cuda_spmv_async", "cfs_cuda_spmv_halo_async",
    "cfs_cuda_spmv_shard_async", "cfs_cuda_spmv_shard_part_async",
    "cfs_cuda_spmv_timed", "cfs_cuda_cg_solve",
    "cfs_cuda_spmv_halo_dot_async", "cfs_cuda_cg_update_xr",
    "cfs_cuda_cg_update_p",
    "cfs_cuda_matrix_export",
    "cfs_cuda_multi_create", "cfs_cuda_multi_tune", "cfs_cuda_multi_spmv",
    "cfs_cuda_multi_info", "cfs_cuda_multi_destroy",
    "cfs_gen_host_count", "cfs_gen_host_fill", "cfs_gen_host_x",
    "cfs_cuda_gen_count", "cfs_cuda_gen_fill", "cfs_cuda_gen_x",
)


class CfsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cfs_cuda error %d: %s" % (code, msg))
        self.code = code


class GenSpec(ctypes.Structure):
    """struct cfs_gen_spec (cfs_spmv_b200/csrc/cfs_gen.h)"""
    _fields_ = [("kind", ctypes.c_int32), ("nx", ctypes.c_int32),
                ("ny", ctypes.c_int32), ("nz", ctypes.c_int32),
                ("nrows", ctypes.c_int64), ("bw", ctypes.c_int32),
                ("per_row", ctypes.c_int32), ("seed", ctypes.c_uint64)]

    @staticmethod
    def laplacian(points, nx, ny, nz, seed=0):
        """seed != 0: one coefficient per edge (cfs_gen.h) instead of -1"""
        assert points in (7, 27)
        return GenSpec(1 if points == 7 else 2, nx, ny, nz, nx * ny * nz, 0, 0,
                       seed)

    @staticmethod
    def banded(nrows, bw, per_row_x16, seed):
        return GenSpec(3, 0, 0, 0, nrows, bw, per_row_x16, seed)

    def ref_tool_spec(self):
        """the same matrix, spelled for oracle/_ref/ref_tool"""
        if self.kind in (1, 2):
            return "gen:lap%d:%d:%d:%d%s" % (
                7 if self.kind == 1 else 27, self.nx, self.ny, self.nz,
                ":%d" % self.seed if self.seed else "")
        return "gen:banded:%d:%d:%d:%d" % (self.nrows, self.bw, self.per_row,
                                           self.seed)


class MatrixInfo(ctypes.Structure):
    _fields_ = [("nrows", ctypes.c_int32), ("ncols", ctypes.c_int32),
                ("row_begin", ctypes.c_int32), ("halo_begin", ctypes.c_int32),
                ("nnz_full", ctypes.c_int64), ("nnz_low", ctypes.c_int64),
                ("nnz_diag", ctypes.c_int64), ("nparts", ctypes.c_int32),
                ("ncolors", ctypes.c_int32), ("nranges", ctypes.c_int32),
                ("symmetric", ctypes.c_int32), ("is_double", ctypes.c_int32),
                ("tuned", ctypes.c_int32), ("refmeta", ctypes.c_int32),
                ("size_bytes", ctypes.c_int64),
                ("device_bytes", ctypes.c_int64),
                ("algorithmic_bytes", ctypes.c_int64),
                ("nvrows", ctypes.c_int64), ("nslices", ctypes.c_int64),
                ("padded_entries", ctypes.c_int64),
                ("nconflict_edges", ctypes.c_int64),
                ("ntiles", ctypes.c_int64), ("far_entries", ctypes.c_int64),
                ("regular_slices", ctypes.c_int64),
                ("index_rows", ctypes.c_int64),
                ("hub_columns", ctypes.c_int64),
                ("hub_entries", ctypes.c_int64),
                ("sort_window", ctypes.c_int64),
                ("transposed_tiles", ctypes.c_int64),
                ("tile_smem_bytes", ctypes.c_int64),
                ("value_dictionary", ctypes.c_int64),
                ("hyb_far_entries", ctypes.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None
MULTI_MAX_GPUS = 16


class MultiInfo(ctypes.Structure):
    """cfs_multi_info of include/cfs_cuda.h"""
    _fields_ = [("ngpus", ctypes.c_int32), ("fused_halo", ctypes.c_int32),
                ("nrows", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("nnz_full", ctypes.c_int64), ("nnz_low", ctypes.c_int64),
                ("device", ctypes.c_int32 * MULTI_MAX_GPUS),
                ("row_begin", ctypes.c_int32 * MULTI_MAX_GPUS),
                ("row_end", ctypes.c_int32 * MULTI_MAX_GPUS),
                ("halo_begin", ctypes.c_int32 * MULTI_MAX_GPUS),
                ("shard_nnz_low", ctypes.c_int64 * MULTI_MAX_GPUS)]


class MmfText(ctypes.Structure):
    """cfs_mmf_text of include/cfs_cuda.h"""
    _fields_ = [("text", ctypes.c_void_p), ("bytes", ctypes.c_size_t),
                ("entries_offset", ctypes.c_size_t),
                ("declared", ctypes.c_int64),
                ("nrows", ctypes.c_int32), ("ncols", ctypes.c_int32),
                ("file_symmetric", ctypes.c_int32),
                ("zero_based", ctypes.c_int32)]


class MmfReport(ctypes.Structure):
    """cfs_mmf_report of include/cfs_cuda.h"""
    _fields_ = [("nnz", ctypes.c_int64), ("host_lines", ctypes.c_int64),
                ("ms_upload", ctypes.c_float), ("ms_parse", ctypes.c_float),
                ("ms_sort", ctypes.c_float), ("ms_build", ctypes.c_float)]


class CgResult(ctypes.Structure):
    """cfs_cg_result of include/cfs_cuda.h"""
    _fields_ = [("iterations", ctypes.c_int32), ("executed", ctypes.c_int32),
                ("converged", ctypes.c_int32), ("breakdown", ctypes.c_int32),
                ("initial_residual_norm", ctypes.c_double),
                ("residual_norm", ctypes.c_double),
                ("ms_total", ctypes.c_float)]


class HostMmfHeader(ctypes.Structure):
    """cfs_host_mmf_header of include/cfs_host.h"""
    _fields_ = [("nrows", ctypes.c_int64), ("ncols", ctypes.c_int64),
                ("declared", ctypes.c_int64), ("symmetric", ctypes.c_int32),
                ("col_wise", ctypes.c_int32), ("zero_based", ctypes.c_int32),
                ("entries_offset", ctypes.c_uint64)]


_host_lib = None


def host_lib():
    """libsparse.so: the host side (Matrix Market header scan, C++ API)"""
    global _host_lib
    if _host_lib is None:
        path = os.path.join(os.path.dirname(LIB_PATH), "libsparse.so")
        L = ctypes.CDLL(path)
        L.cfs_host_scan_mmf_header.argtypes = [
            ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(HostMmfHeader)]
        _host_lib = L
    return _host_lib


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "%s is missing: build it with `python cfs_spmv_b200/build.py` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u64, sz = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64,
                             ctypes.c_uint64, ctypes.c_size_t)
    L.cfs_cuda_last_error.restype = ctypes.c_char_p
    L.cfs_cuda_version.restype = ctypes.c_char_p
    L.cfs_cuda_device_count.argtypes = [ctypes.POINTER(ctypes.c_int)]
    L.cfs_cuda_init.argtypes = [ctypes.c_int]
    L.cfs_cuda_set_option.argtypes = [ctypes.c_char_p, ctypes.c_longlong]
    L.cfs_cuda_host_alloc.restype = vp
    L.cfs_cuda_host_alloc.argtypes = [sz]
    L.cfs_cuda_host_free.argtypes = [vp]
    L.cfs_cuda_host_alloc_kind.restype = vp
    L.cfs_cuda_host_alloc_kind.argtypes = [sz, ctypes.c_int]
    L.cfs_cuda_vector_prefetch.argtypes = [vp, sz, ctypes.c_int]
    L.cfs_cuda_matrix_create.argtypes = [ctypes.POINTER(vp), i32, i32, vp, vp,
                                         vp, ctypes.c_int, ctypes.c_int]
    L.cfs_cuda_matrix_create_shard.argtypes = [ctypes.POINTER(vp), i32, i32,
                                               i32, vp, vp, vp, ctypes.c_int]
    L.cfs_cuda_matrix_tune.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    L.cfs_cuda_matrix_set_hybrid.argtypes = [vp, i32]
    L.cfs_cuda_matrix_create_from_mmf.argtypes = [
        ctypes.POINTER(vp), ctypes.POINTER(MmfText), ctypes.c_int,
        ctypes.c_int, ctypes.POINTER(MmfReport)]
    L.cfs_cuda_matrix_download_csr.argtypes = [vp, vp, vp, vp]
    L.cfs_cuda_matrix_destroy.argtypes = [vp]
    L.cfs_cuda_matrix_destroy.restype = None
    L.cfs_cuda_matrix_info.argtypes = [vp, ctypes.POINTER(MatrixInfo)]
    L.cfs_cuda_spmv.argtypes = [vp, vp, vp]
    L.cfs_cuda_spmv_async.argtypes = [vp, vp, vp, vp]
    L.cfs_cuda_spmv_halo_async.argtypes = [vp, vp, vp, vp, ctypes.c_int, vp]
    L.cfs_cuda_spmv_shard_async.argtypes = [vp, vp, vp, vp, vp, vp,
                                            ctypes.c_int, vp]
    L.cfs_cuda_spmv_shard_part_async.argtypes = [vp, vp, vp, vp, vp, vp,
                                                 ctypes.c_int, vp]
    L.cfs_cuda_spmv_timed.argtypes = [vp, vp, vp, vp, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_float),
                                      ctypes.POINTER(ctypes.c_float)]
    L.cfs_cuda_matrix_export.argtypes = [vp, ctypes.c_int, vp, sz,
                                         ctypes.POINTER(sz)]
    L.cfs_cuda_multi_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int,
                                        ctypes.c_int, i32, vp, vp, vp,
                                        ctypes.c_int]
    L.cfs_cuda_multi_tune.argtypes = [vp]
    L.cfs_cuda_multi_spmv.argtypes = [vp, vp, vp]
    L.cfs_cuda_multi_info.argtypes = [vp, ctypes.POINTER(MultiInfo)]
    L.cfs_cuda_multi_destroy.argtypes = [vp]
    L.cfs_cuda_multi_destroy.restype = None
    L.cfs_cuda_cg_solve.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_double,
                                    ctypes.POINTER(CgResult), vp, ctypes.c_int]
    L.cfs_cuda_spmv_halo_dot_async.argtypes = [vp, vp, vp, vp, ctypes.c_int,
                                               vp, vp]
    L.cfs_cuda_cg_update_xr.argtypes = [i64, ctypes.c_int, vp, vp, vp, vp, vp,
                                        vp, vp]
    L.cfs_cuda_cg_update_p.argtypes = [i64, ctypes.c_int, vp, vp, vp, vp]
    gs = ctypes.POINTER(GenSpec)
    L.cfs_gen_host_count.argtypes = [gs, i64, i64, vp]
    L.cfs_gen_host_fill.argtypes = [gs, i64, i64, vp, vp, vp, ctypes.c_int]
    L.cfs_gen_host_x.argtypes = [u64, i64, i64, vp, ctypes.c_int]
    L.cfs_cuda_gen_count.argtypes = [gs, i64, i64, vp, ctypes.POINTER(i64)]
    L.cfs_cuda_gen_fill.argtypes = [gs, i64, i64, vp, vp, vp, ctypes.c_int]
    L.cfs_cuda_gen_x.argtypes = [u64, i64, i64, vp, ctypes.c_int]
    _lib = L
    return L


def check(code):
    if code != CFS_OK:
        raise CfsError(code, lib().cfs_cuda_last_error().decode())


def device_count():
    n = ctypes.c_int(0)
    code = lib().cfs_cuda_device_count(ctypes.byref(n))
    return n.value if code == CFS_OK else 0


def init(device=0):
    check(lib().cfs_cuda_init(device))


def set_option(key, value):
    check(lib().cfs_cuda_set_option(key.encode(), int(value)))


def _ptr(a):
    """numpy array / torch tensor / int -> raw address"""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return a.data_ptr()  # torch tensor (device or pinned host)


class Matrix:
    """cfs_mat_t: create -> tune -> spmv / export (mirrors CSRMatrix + SpDMV)."""

    def __init__(self, nrows, ncols, rowptr, colind, values, is_double=True,
                 symmetric=True, shard=None):
        self._h = ctypes.c_void_p()
        self.is_double = bool(is_double)
        self.dtype = np.float64 if is_double else np.float32
        self._keep = (rowptr, colind, values)  # borrowed until tune()
        if shard is None:
            check(lib().cfs_cuda_matrix_create(
                ctypes.byref(self._h), nrows, ncols, _ptr(rowptr),
                _ptr(colind), _ptr(values), int(is_double), int(symmetric)))
        else:
            global_nrows, row_begin, row_end = shard
            check(lib().cfs_cuda_matrix_create_shard(
                ctypes.byref(self._h), global_nrows, row_begin, row_end,
                _ptr(rowptr), _ptr(colind), _ptr(values), int(is_double)))

    @classmethod
    def from_csr(cls, rowptr, colind, values, symmetric=True):
        rowptr = np.ascontiguousarray(rowptr, np.int32)
        colind = np.ascontiguousarray(colind, np.int32)
        values = np.ascontiguousarray(values)
        assert values.dtype in (np.float32, np.float64)
        n = len(rowptr) - 1
        return cls(n, n, rowptr, colind, values, values.dtype == np.float64,
                   symmetric)

    @classmethod
    def from_mmf(cls, path, is_double=True, symmetric=True):
        """CSRMatrix(filename) with the entries parsed on the GPU: header on
        the host (cfs_host_scan_mmf_header), the rest by
        cfs_cuda_matrix_create_from_mmf. Returns (matrix, header, report);
        raises CfsError(CFS_ERR_NEEDS_HOST) where the host loader must decide."""
        image = np.fromfile(path, dtype=np.uint8)
        h = HostMmfHeader()
        assert host_lib().cfs_host_scan_mmf_header(
            image.ctypes.data, image.size, ctypes.byref(h)) == 0
        t = MmfText(image.ctypes.data, image.size, h.entries_offset,
                    h.declared, h.nrows, h.ncols, h.symmetric, h.zero_based)
        self = cls.__new__(cls)
        self._h = ctypes.c_void_p()
        self.is_double = bool(is_double)
        self.dtype = np.float64 if is_double else np.float32
        self._keep = None
        rep = MmfReport()
        check(lib().cfs_cuda_matrix_create_from_mmf(
            ctypes.byref(self._h), ctypes.byref(t), int(is_double),
            int(symmetric), ctypes.byref(rep)))
        self.nrows, self.ncols = int(h.nrows), int(h.ncols)
        report = {k: getattr(rep, k) for k, _ in MmfReport._fields_}
        return self, h, report

    def download_csr(self, nrows, nnz):
        """the full CSR held in HBM (before tune() releases it)"""
        rp = np.empty(nrows + 1, np.int32)
        ci = np.empty(nnz, np.int32)
        v = np.empty(nnz, self.dtype)
        check(lib().cfs_cuda_matrix_download_csr(self._h, _ptr(rp), _ptr(ci),
                                                 _ptr(v)))
        return rp, ci, v

    def set_hybrid(self, threshold=10000):
        """Format::hyb: entries at least `threshold` off the diagonal go to a
        non-symmetric far part (call before tune)"""
        check(lib().cfs_cuda_matrix_set_hybrid(self._h, threshold))

    def tune(self, nparts=1, tuning=1, allow_too_large=False):
        code = lib().cfs_cuda_matrix_tune(self._h, nparts, tuning)
        self._keep = None
        if code == CFS_ERR_TOO_LARGE and allow_too_large:
            return code
        check(code)
        return code

    def info(self):
        mi = MatrixInfo()
        check(lib().cfs_cuda_matrix_info(self._h, ctypes.byref(mi)))
        return mi.as_dict()

    def spmv(self, y, x):
        """synchronous; host numpy arrays or device tensors"""
        check(lib().cfs_cuda_spmv(self._h, _ptr(y), _ptr(x)))
        return y

    def spmv_async(self, y_dev, x_dev, stream=0):
        check(lib().cfs_cuda_spmv_async(self._h, _ptr(y_dev), _ptr(x_dev),
                                        stream))

    def spmv_halo_async(self, y_dev, x_dev, y_lower_base, y_is_zero, stream=0):
        """SpMV fused with the halo reduction into the GPU below (NVLink)"""
        check(lib().cfs_cuda_spmv_halo_async(self._h, _ptr(y_dev), _ptr(x_dev),
                                             y_lower_base, int(y_is_zero),
                                             stream))

    def spmv_shard_async(self, y_dev, x_dev, y_lower_base=None,
                         x_lower_base=None, y_clear=None, y_is_zero=True,
                         stream=0):
        """spmv_halo_async that reads the x halo from the GPU below and clears
        the owned rows of y_clear for the next SpMV (ping-pong results)"""
        check(lib().cfs_cuda_spmv_shard_async(
            self._h, _ptr(y_dev), _ptr(x_dev), y_lower_base, x_lower_base,
            _ptr(y_clear), int(y_is_zero), stream))

    def spmv_shard_part_async(self, y_dev, x_dev, y_lower_base, x_lower_base,
                              y_clear, part, stream=0):
        """part 1: the slices that reach below row_begin (remote x / y);
        part 2: all the others"""
        check(lib().cfs_cuda_spmv_shard_part_async(
            self._h, _ptr(y_dev), _ptr(x_dev), y_lower_base, x_lower_base,
            _ptr(y_clear), part, stream))

    def cg_solve(self, x, b, max_iters, rel_tol, want_history=False):
        """A x = b by conjugate gradients on the device (cfs_cuda_cg_solve).
        x: initial guess in, solution out (numpy or torch, host or device).
        Returns a dict of cfs_cg_result (+ 'history' of residual norms)."""
        res = CgResult()
        hist = np.zeros(max_iters + 1) if want_history else None
        check(lib().cfs_cuda_cg_solve(
            self._h, _ptr(x), _ptr(b), int(max_iters), float(rel_tol),
            ctypes.byref(res), _ptr(hist), max_iters + 1 if want_history else 0))
        out = {k: getattr(res, k) for k, _ in CgResult._fields_}
        if want_history:
            out["history"] = hist[:res.iterations + 1]
        return out

    def spmv_halo_dot_async(self, y_dev, x_dev, y_lower_base, y_is_zero,
                            dot_dev, stream=0):
        """spmv_halo_async that also adds this shard's x'(A x) to *dot_dev"""
        check(lib().cfs_cuda_spmv_halo_dot_async(
            self._h, _ptr(y_dev), _ptr(x_dev), y_lower_base, int(y_is_zero),
            _ptr(dot_dev), stream))

    def spmv_timed(self, y_dev, x_dev, iters, stream=0):
        """-> (total_ms, kernel_ms) summed over `iters` SpMVs"""
        total, kern = ctypes.c_float(0), ctypes.c_float(0)
        check(lib().cfs_cuda_spmv_timed(self._h, _ptr(y_dev), _ptr(x_dev),
                                        stream, iters, ctypes.byref(total),
                                        ctypes.byref(kern)))
        return total.value, kern.value

    def export(self, name):
        sel = META[name]
        n = ctypes.c_size_t(0)
        check(lib().cfs_cuda_matrix_export(self._h, sel, None, 0,
                                           ctypes.byref(n)))
        dtype = self.dtype if name in _VALUE_META else np.int32
        out = np.zeros(n.value, dtype=dtype)
        if n.value:
            check(lib().cfs_cuda_matrix_export(self._h, sel, out.ctypes.data,
                                               n.value, ctypes.byref(n)))
        return out

    def metadata(self):
        """same keys as oracle.Oracle.metadata() / a ref_tool dump"""
        inf = self.info()
        P = inf["nparts"]
        row_split = self.export("row_split")
        md = {
            "nrows": inf["nrows"], "P": P, "ncolors": inf["ncolors"],
            "nranges": inf["nranges"], "nnz_low": inf["nnz_low"],
            "nnz_diag": inf["nnz_diag"], "size_bytes": inf["size_bytes"],
            "nnz_full": inf["nnz_full"],
            "row_split": row_split if P > 1 else np.zeros(0, np.int32),
            "part_nrows": np.diff(row_split).astype(np.int32),
            "part_offset": row_split[:-1].copy(),
        }
        for k in ("part_nnz_low", "lower_rowptr", "lower_colind",
                  "lower_values", "diagonal", "range_ptr", "part_nranges",
                  "range_start", "range_end"):
            md[k] = self.export(k)
        return md

    def close(self):
        if self._h:
            lib().cfs_cuda_matrix_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiMatrix:
    """cfs_multi_t: a symmetric matrix cut into row shards over the GPUs of
    this process (what the C++ layer does under CFS_NUM_GPUS > 1)"""

    def __init__(self, rowptr, colind, values, ngpus, first_device=0):
        self._h = ctypes.c_void_p()
        rowptr = np.ascontiguousarray(rowptr, np.int32)
        colind = np.ascontiguousarray(colind, np.int32)
        values = np.ascontiguousarray(values)
        assert values.dtype in (np.float32, np.float64)
        self.dtype = values.dtype
        check(lib().cfs_cuda_multi_create(
            ctypes.byref(self._h), ngpus, first_device, len(rowptr) - 1,
            _ptr(rowptr), _ptr(colind), _ptr(values),
            int(values.dtype == np.float64)))
        check(lib().cfs_cuda_multi_tune(self._h))

    def info(self):
        mi = MultiInfo()
        check(lib().cfs_cuda_multi_info(self._h, ctypes.byref(mi)))
        g = mi.ngpus
        return {"ngpus": g, "fused_halo": mi.fused_halo, "nrows": mi.nrows,
                "nnz_full": mi.nnz_full, "nnz_low": mi.nnz_low,
                "device": list(mi.device[:g]),
                "row_begin": list(mi.row_begin[:g]),
                "row_end": list(mi.row_end[:g]),
                "halo_begin": list(mi.halo_begin[:g]),
                "shard_nnz_low": list(mi.shard_nnz_low[:g])}

    def spmv(self, y, x):
        check(lib().cfs_cuda_multi_spmv(self._h, _ptr(y), _ptr(x)))
        return y

    def close(self):
        if self._h:
            lib().cfs_cuda_multi_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- synthetic inputs ------------------------------------------------------
def cg_update_xr(n, is_double, scal, p, q, x, r, rr_next, stream=0):
    check(lib().cfs_cuda_cg_update_xr(n, int(is_double), _ptr(scal), _ptr(p),
                                      _ptr(q), _ptr(x), _ptr(r), _ptr(rr_next),
                                      stream))


def cg_update_p(n, is_double, scal, r, p, stream=0):
    check(lib().cfs_cuda_cg_update_p(n, int(is_double), _ptr(scal), _ptr(r),
                                     _ptr(p), stream))


def gen_host_csr(spec, row_begin=0, row_end=None, dtype=np.float64):
    """full CSR of rows [row_begin,row_end) on the host (no GPU needed)"""
    row_end = spec.nrows if row_end is None else row_end
    n = row_end - row_begin
    rowptr = np.zeros(n + 1, np.int32)
    check(lib().cfs_gen_host_count(ctypes.byref(spec), row_begin, row_end,
                                   rowptr.ctypes.data))
    nnz = int(rowptr[-1])
    colind = np.zeros(nnz, np.int32)
    values = np.zeros(nnz, dtype)
    check(lib().cfs_gen_host_fill(ctypes.byref(spec), row_begin, row_end,
                                  rowptr.ctypes.data, colind.ctypes.data,
                                  values.ctypes.data,
                                  int(dtype == np.float64)))
    return rowptr, colind, values


def gen_host_x(seed, n, dtype=np.float64, begin=0):
    x = np.zeros(n, dtype)
    check(lib().cfs_gen_host_x(seed, begin, begin + n, x.ctypes.data,
                               int(dtype == np.float64)))
    return x


def gen_device_csr(spec, row_begin=0, row_end=None, is_double=True,
                   device="cuda"):
    """full CSR of rows [row_begin,row_end) built directly in HBM; returns torch
    tensors (torch only supplies the device memory)."""
    import torch
    row_end = spec.nrows if row_end is None else row_end
    n = row_end - row_begin
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=device)
    nnz = ctypes.c_int64(0)
    check(lib().cfs_cuda_gen_count(ctypes.byref(spec), row_begin, row_end,
                                   rowptr.data_ptr(), ctypes.byref(nnz)))
    colind = torch.empty(nnz.value, dtype=torch.int32, device=device)
    values = torch.empty(nnz.value, device=device,
                         dtype=torch.float64 if is_double else torch.float32)
    check(lib().cfs_cuda_gen_fill(ctypes.byref(spec), row_begin, row_end,
                                  rowptr.data_ptr(), colind.data_ptr(),
                                  values.data_ptr(), int(is_double)))
    return rowptr, colind, values


def gen_device_x(seed, begin, end, is_double=True, device="cuda"):
    import torch
    x = torch.empty(end - begin, device=device,
                    dtype=torch.float64 if is_double else torch.float32)
    check(lib().cfs_cuda_gen_x(seed, begin, end, x.data_ptr(), int(is_double)))
    return x
