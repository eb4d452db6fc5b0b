"""The host-side Matrix Market loader + CSR construction of libsparse.so
against dumps of the reference's own loader (tests/golden/mtx/*.npz hold the
full CSR the compiled reference built from each fixture)."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from cfs_spmv_b200 import capi, gen

MTX_DIR = os.path.join(cases.GOLDEN_DIR, "mtx")
LIBSPARSE = os.path.join(os.path.dirname(capi.LIB_PATH), "libsparse.so")
GOOD = sorted(f[:-4] for f in os.listdir(MTX_DIR)
              if f.endswith(".mtx") and not f.startswith("err_"))
BAD = sorted(f for f in os.listdir(MTX_DIR) if f.startswith("err_"))


class HostCsr(ctypes.Structure):
    _fields_ = [("nrows", ctypes.c_int32), ("ncols", ctypes.c_int32),
                ("nnz", ctypes.c_int32), ("symmetric", ctypes.c_int32),
                ("rowptr", ctypes.POINTER(ctypes.c_int32)),
                ("colind", ctypes.POINTER(ctypes.c_int32)),
                ("values", ctypes.POINTER(ctypes.c_double)),
                ("handle", ctypes.c_void_p)]


def load(path, want_symmetric=True):
    L = ctypes.CDLL(LIBSPARSE)
    m = HostCsr()
    assert L.cfs_host_load_mmf(path.encode(), int(want_symmetric),
                               ctypes.byref(m)) == 0
    out = {
        "nrows": m.nrows, "ncols": m.ncols, "nnz": m.nnz,
        "symmetric": m.symmetric,
        "rowptr": np.ctypeslib.as_array(m.rowptr, (m.nrows + 1,)).copy(),
        "colind": np.ctypeslib.as_array(m.colind, (max(m.nnz, 1),))[:m.nnz].copy(),
        "values": np.ctypeslib.as_array(m.values, (max(m.nnz, 1),))[:m.nnz].copy(),
    }
    L.cfs_host_free_csr(ctypes.byref(m))
    return out


@pytest.mark.parametrize("name", GOOD)
def test_loader_matches_reference(name):
    gold = np.load(os.path.join(MTX_DIR, name + "-P1.npz"))
    got = load(os.path.join(MTX_DIR, name + ".mtx"), want_symmetric=True)
    assert got["nrows"] == int(gold["nrows"])
    assert got["ncols"] == int(gold["ncols"])
    assert got["nnz"] == int(gold["nnz_full"])
    assert got["symmetric"] == int(gold["symmetric"])
    assert np.array_equal(got["rowptr"], gold["csr_rowptr"])
    assert np.array_equal(got["colind"], gold["csr_colind"])
    assert got["values"].tobytes() == gold["csr_values"].tobytes()


def test_format_csr_never_compresses():
    got = load(os.path.join(MTX_DIR, "sym_lower.mtx"), want_symmetric=False)
    assert got["symmetric"] == 0 and got["nnz"] == 160


@pytest.mark.parametrize("fname", BAD)
def test_loader_errors_like_the_reference(fname):
    """message on stdout + exit(1), same text as the reference prints"""
    expect = json.load(open(os.path.join(MTX_DIR, "errors.json")))[fname]
    code = ("import ctypes,sys; L=ctypes.CDLL(%r); b=ctypes.create_string_buffer(64);"
            "L.cfs_host_load_mmf(%r, 1, b)" % (
                LIBSPARSE, os.path.join(MTX_DIR, fname).encode()))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True,
                       text=True)
    assert r.returncode == expect["exit"] == 1
    assert r.stdout == expect["stdout"]


def test_missing_file_is_fatal():
    code = ("import ctypes; L=ctypes.CDLL(%r); b=ctypes.create_string_buffer(64);"
            "L.cfs_host_load_mmf(b'/nonexistent/x.mtx', 1, b)" % LIBSPARSE)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True,
                       text=True)
    assert r.returncode == 1 and r.stdout == "MMF file error.\n"


def test_written_laplacian_round_trips(tmp_path):
    """gen.write_mtx (lower triangle, diagonal first) -> loader -> full CSR"""
    rp, ci, v = cases.matrix("lap7_9x7x5")
    path = str(tmp_path / "lap.mtx")
    gen.write_mtx(path, rp, ci, v)
    got = load(path)
    assert got["symmetric"] == 1
    assert np.array_equal(got["rowptr"], rp)
    assert np.array_equal(got["colind"], ci)
    assert np.array_equal(got["values"], v)


def test_libsparse_exports_the_cxx_api():
    """explicit instantiations for <int,float> and <int,double>
    (reference src/cfs.cpp:11-21, src/csr.cpp:10-11, src/mmf.cpp)"""
    out = subprocess.run(["nm", "-DC", "--defined-only", LIBSPARSE],
                         capture_output=True, text=True).stdout
    for sym in ("cfs::matrix::sparse::SparseMatrix<int, double>::create(",
                "cfs::matrix::sparse::SparseMatrix<int, float>::create(",
                "cfs::matrix::sparse::CSRMatrix<int, double>::tune(",
                "cfs::matrix::sparse::CSRMatrix<int, float>::tune(",
                "cfs::kernel::sparse::SpDMV<int, double>::operator()(",
                "cfs::kernel::sparse::SpDMV<int, float>::operator()(",
                "cfs::util::memory::internal_alloc(",
                "cfs::util::memory::internal_free(",
                "cfs::util::runtime::get_num_threads()",
                "cfs::util::runtime::setaffinity_oncpu(",
                "cfs::io::DoRead(", "cfs_host_load_mmf"):
        assert sym in out, sym
