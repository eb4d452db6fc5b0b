"""ctypes front end of the CPU oracle (oracle/cfs_oracle.c) and helpers to run
the compiled reference (oracle/_ref/ref_tool) and read its dumps.

TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module. The product
(cfs_spmv_b200/) never does.
"""
import ctypes
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(BUILD_DIR, "libcfs_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_TOOL = os.path.join(REF_DIR, "ref_tool")
REFERENCE_ROOT = "/root/reference"
HOST_CXX = "/usr/bin/g++"
HOST_CC = "/usr/bin/gcc"


def build(force=False):
    """Compile the C restatement. -ffp-contract=off and no -march: the same
    mul-then-add arithmetic the reference gets from `g++ -O2` on x86-64."""
    src = os.path.join(HERE, "cfs_oracle.c")
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= os.path.getmtime(src)):
        return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    cc = HOST_CC if os.path.exists(HOST_CC) else "gcc"
    subprocess.check_call([cc, "-O2", "-std=c99", "-fPIC", "-shared",
                           "-ffp-contract=off", "-w", src, "-o", LIB_PATH])
    return LIB_PATH


def build_reference(force=False):
    """Compile the unmodified reference into oracle/_ref (only possible where
    /root/reference exists, i.e. in the build container)."""
    if not os.path.isdir(REFERENCE_ROOT):
        return os.path.exists(REF_TOOL)
    # make decides what is stale (it is quick when nothing is)
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "ref_build"),
                           "-j8"])
    return os.path.exists(REF_TOOL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.cfs_oracle_build.restype = ctypes.c_void_p
        L.cfs_oracle_build.argtypes = [ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_int]
        L.cfs_oracle_free.argtypes = [ctypes.c_void_p]
        L.cfs_oracle_size_bytes.restype = ctypes.c_longlong
        L.cfs_oracle_size_bytes.argtypes = [ctypes.c_void_p]
        L.cfs_oracle_spmv.argtypes = [ctypes.c_void_p] * 3
        L.cfs_oracle_csr_spmv.argtypes = [ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p]
        L.cfs_oracle_partition_by_nrows.argtypes = [ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_void_p]
        L.cfs_oracle_partition_by_nnz.argtypes = [ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_void_p,
                                                  ctypes.c_void_p]
        for name in ("nrows", "P", "nnz_low", "nnz_diag", "ncolors", "nranges",
                     "nblk"):
            f = getattr(L, "cfs_oracle_" + name)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_void_p]
        for name in ("row_split", "part_nnz_low", "lower_rowptr",
                     "lower_colind", "lower_values", "diagonal", "weight",
                     "adj_ptr", "adj", "color_first", "color", "range_ptr",
                     "part_nranges", "range_start", "range_end"):
            f = getattr(L, "cfs_oracle_" + name)
            f.restype = ctypes.c_void_p
            f.argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


def _arr(ptr, count, dtype):
    if count == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    nbytes = count * np.dtype(dtype).itemsize
    buf = (ctypes.c_char * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count).copy()


def partition_by_nrows(nrows, P):
    out = np.zeros(P + 1, dtype=np.int32)
    lib().cfs_oracle_partition_by_nrows(nrows, P, out.ctypes.data)
    return out


def partition_by_nnz(rowptr, P):
    """row_split_ of a NON-symmetric matrix under Tuning::Aggressive
    (csr_matrix.tpp:438-541)"""
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    out = np.zeros(P + 1, dtype=np.int32)
    lib().cfs_oracle_partition_by_nnz(len(rowptr) - 1, P, rowptr.ctypes.data,
                                      out.ctypes.data)
    return out


def csr_spmv(rowptr, colind, values, x):
    """cpu_mv / cpu_mv_serial (csr_matrix.tpp:2665-2704): y[i] summed in CSR
    order with one accumulator"""
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    colind = np.ascontiguousarray(colind, np.int32)
    values = np.ascontiguousarray(values)
    x = np.ascontiguousarray(x, values.dtype)
    n = len(rowptr) - 1
    y = np.zeros(n, values.dtype)
    lib().cfs_oracle_csr_spmv(n, rowptr.ctypes.data, colind.ctypes.data,
                              values.ctypes.data,
                              int(values.dtype == np.float64), y.ctypes.data,
                              x.ctypes.data)
    return y


def valid_partition_count(nrows, P):
    """the reference overshoots (and crashes) unless this holds."""
    if P == 1:
        return True
    s = ((nrows // P - 1) | 15) + 1
    return nrows // P >= 1 and (P - 1) * s <= nrows


class Oracle:
    """The reference's tune() + SpMV on a full CSR matrix, restated on the CPU."""

    def __init__(self, rowptr, colind, values, P=1):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        self.colind = np.ascontiguousarray(colind, dtype=np.int32)
        assert values.dtype in (np.float32, np.float64)
        self.values = np.ascontiguousarray(values)
        self.dtype = self.values.dtype
        self.nrows = len(self.rowptr) - 1
        self.P = int(P)
        assert valid_partition_count(self.nrows, self.P)
        L = lib()
        self._h = L.cfs_oracle_build(self.nrows, self.rowptr.ctypes.data,
                                     self.colind.ctypes.data,
                                     self.values.ctypes.data,
                                     int(self.dtype == np.float64), self.P)
        h = self._h
        N, P = self.nrows, self.P
        self.nnz_full = int(self.rowptr[-1])
        self.nnz_low = L.cfs_oracle_nnz_low(h)
        self.nnz_diag = L.cfs_oracle_nnz_diag(h)
        self.ncolors = L.cfs_oracle_ncolors(h)
        self.nranges = L.cfs_oracle_nranges(h)
        self.size_bytes = L.cfs_oracle_size_bytes(h)
        self.row_split = _arr(L.cfs_oracle_row_split(h), P + 1, np.int32)
        self.part_nnz_low = _arr(L.cfs_oracle_part_nnz_low(h), P, np.int32)
        self.lower_rowptr = _arr(L.cfs_oracle_lower_rowptr(h), N + P, np.int32)
        self.lower_colind = _arr(L.cfs_oracle_lower_colind(h), self.nnz_low,
                                 np.int32)
        self.lower_values = _arr(L.cfs_oracle_lower_values(h), self.nnz_low,
                                 self.dtype)
        self.diagonal = _arr(L.cfs_oracle_diagonal(h), N, self.dtype)
        if P > 1:
            V = L.cfs_oracle_nblk(h)
            self.nblk = V
            self.weight = _arr(L.cfs_oracle_weight(h), V, np.int32)
            self.adj_ptr = _arr(L.cfs_oracle_adj_ptr(h), V + 1, np.int32)
            self.adj = _arr(L.cfs_oracle_adj(h), int(self.adj_ptr[-1]),
                            np.int32)
            self.color_first = _arr(L.cfs_oracle_color_first(h), V, np.int32)
            self.color = _arr(L.cfs_oracle_color(h), V, np.int32)
            self.range_ptr = _arr(L.cfs_oracle_range_ptr(h),
                                  P * (self.ncolors + 1), np.int32)
            self.part_nranges = _arr(L.cfs_oracle_part_nranges(h), P, np.int32)
            self.range_start = _arr(L.cfs_oracle_range_start(h), self.nranges,
                                    np.int32)
            self.range_end = _arr(L.cfs_oracle_range_end(h), self.nranges,
                                  np.int32)
        else:
            self.range_ptr = np.zeros(0, np.int32)
            self.part_nranges = np.zeros(0, np.int32)
            self.range_start = np.zeros(0, np.int32)
            self.range_end = np.zeros(0, np.int32)

    def spmv(self, x, y0=None):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        y = (np.zeros(self.nrows, dtype=self.dtype) if y0 is None
             else np.array(y0, dtype=self.dtype))
        lib().cfs_oracle_spmv(self._h, y.ctypes.data, x.ctypes.data)
        return y

    def csr_spmv(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        y = np.zeros(self.nrows, dtype=self.dtype)
        lib().cfs_oracle_csr_spmv(self.nrows, self.rowptr.ctypes.data,
                                  self.colind.ctypes.data,
                                  self.values.ctypes.data,
                                  int(self.dtype == np.float64),
                                  y.ctypes.data, x.ctypes.data)
        return y

    def metadata(self):
        """Same keys as a ref_tool dump (read_dump), for bit-exact comparison."""
        P = self.P
        part_nrows = np.diff(self.row_split).astype(np.int32)
        return {
            "nrows": self.nrows, "P": P, "ncolors": self.ncolors,
            "nranges": self.nranges, "nnz_low": self.nnz_low,
            "nnz_diag": self.nnz_diag, "size_bytes": self.size_bytes,
            "nnz_full": self.nnz_full,
            "row_split": self.row_split if P > 1 else np.zeros(0, np.int32),
            "part_nrows": part_nrows,
            "part_offset": self.row_split[:-1].copy(),
            "part_nnz_low": self.part_nnz_low,
            "part_nranges": self.part_nranges,
            "lower_rowptr": self.lower_rowptr,
            "lower_colind": self.lower_colind,
            "lower_values": self.lower_values,
            "diagonal": self.diagonal,
            "range_ptr": self.range_ptr,
            "range_start": self.range_start,
            "range_end": self.range_end,
        }

    def __del__(self):
        try:
            if self._h:
                lib().cfs_oracle_free(self._h)
                self._h = None
        except Exception:
            pass


METADATA_KEYS = ("row_split", "part_nrows", "part_offset", "part_nnz_low",
                 "part_nranges", "lower_rowptr", "lower_colind", "lower_values",
                 "diagonal", "range_ptr", "range_start", "range_end")
SCALAR_KEYS = ("nrows", "P", "ncolors", "nranges", "nnz_low", "nnz_diag",
               "size_bytes", "nnz_full")


# ---------------------------------------------------------------------------
# compiled reference: run + read
# ---------------------------------------------------------------------------
_DTYPES = {0: np.int32, 1: np.int64, 2: np.float32, 3: np.float64}


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        (nl,) = struct.unpack_from("<I", data, pos)
        pos += 4
        name = data[pos:pos + nl].decode()
        pos += nl
        dt = data[pos]
        pos += 1
        (count,) = struct.unpack_from("<Q", data, pos)
        pos += 8
        dtype = np.dtype(_DTYPES[dt])
        arr = np.frombuffer(data, dtype=dtype, count=count, offset=pos).copy()
        pos += count * dtype.itemsize
        out[name] = int(arr[0]) if (dt == 1 and count == 1) else arr
    return out


def write_csr_bin(path, rowptr, colind, values, ncols=None):
    nrows = len(rowptr) - 1
    with open(path, "wb") as f:
        f.write(struct.pack("<qqq", nrows, ncols or nrows, len(colind)))
        f.write(np.ascontiguousarray(rowptr, np.int32).tobytes())
        f.write(np.ascontiguousarray(colind, np.int32).tobytes())
        f.write(np.ascontiguousarray(values, np.float64).tobytes())


def ref_available():
    return os.path.exists(REF_TOOL)


def run_ref_dump(input_spec, P, precision, xseed, out_path):
    """input_spec: 'mtx:<file>' | 'csr:<file>' | 'gen:...' (see ref_tool.cpp)."""
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(P)
    subprocess.check_call([REF_TOOL, "dump", input_spec, str(P), precision,
                           str(xseed), out_path], env=env)
    return read_dump(out_path)


def run_ref_dump_csr(input_spec, P, precision, xseed, out_path, tuning="A"):
    """the reference's NON-symmetric path: row_split_ + y (ref_tool dumpcsr)"""
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(P)
    subprocess.check_call([REF_TOOL, "dumpcsr", input_spec, str(P), precision,
                           str(xseed), out_path, tuning], env=env)
    return read_dump(out_path)


def run_ref_bench(input_spec, P, precision, xseed, loops, timeout=None,
                  warmup=-1, y_out=None):
    """times the compiled reference's SpDMV (ref_tool bench); warmup = -1 is
    the reference's own loops / 2; y_out: file that receives its y (raw)"""
    import json
    env = dict(os.environ)
    env.setdefault("OMP_PROC_BIND", "close")
    cmd = [REF_TOOL, "bench", input_spec, str(P), precision, str(xseed),
           str(loops), str(warmup)]
    if y_out:
        cmd.append(y_out)
    out = subprocess.run(cmd, env=env, check=True,
                         capture_output=True, text=True, timeout=timeout)
    for line in out.stdout.splitlines():
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("ref_tool bench printed no JSON: " + out.stdout)
