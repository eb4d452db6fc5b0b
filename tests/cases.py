"""Parity cases shared by the golden generator, the CPU oracle tests and the
GPU parity tests. Every case is a small symmetric matrix (FULL CSR, full
diagonal) + a partition count P + a precision."""
import os

import numpy as np

from cfs_spmv_b200 import capi, gen

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
XSEED = 20261018


def _lap(points, nx, ny=None, nz=None):
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    return capi.GenSpec.laplacian(points, nx, ny, nz)


# name -> (builder returning (rowptr, colind, values[f64]) , ref_tool input or None)
MATRICES = {
    "lap7_12": lambda: capi.gen_host_csr(_lap(7, 12)),
    "lap7_9x7x5": lambda: capi.gen_host_csr(_lap(7, 9, 7, 5)),
    "lap27_10": lambda: capi.gen_host_csr(_lap(27, 10)),
    "lap27_14": lambda: capi.gen_host_csr(_lap(27, 14)),
    "banded_3000": lambda: capi.gen_host_csr(
        capi.GenSpec.banded(3000, 100, 152, 7)),
    "rmat_9": lambda: gen.rmat(9, 8, seed=1),
    "ragged_333": lambda: gen.random_symmetric(333, 3.0, seed=5),
    "ragged_64": lambda: gen.random_symmetric(64, 1.0, seed=11),
    "diag_only_48": lambda: (np.arange(49, dtype=np.int32),
                             np.arange(48, dtype=np.int32),
                             np.linspace(1.0, 2.0, 48)),
}

GEN_SPECS = {
    "lap7_12": "gen:lap7:12:12:12", "lap7_9x7x5": "gen:lap7:9:7:5",
    "lap27_10": "gen:lap27:10:10:10", "lap27_14": "gen:lap27:14:14:14",
    "banded_3000": "gen:banded:3000:100:152:7",
}

# (matrix, P, precision)
CASES = [
    ("lap7_12", 1, "d"), ("lap7_12", 2, "d"), ("lap7_12", 4, "d"),
    ("lap7_12", 8, "d"), ("lap7_12", 4, "s"), ("lap7_12", 27, "d"),
    ("lap7_9x7x5", 3, "d"),
    ("lap27_10", 1, "d"), ("lap27_10", 4, "d"), ("lap27_10", 8, "d"),
    ("lap27_10", 16, "d"), ("lap27_10", 8, "s"),
    ("lap27_14", 16, "d"), ("lap27_14", 43, "d"),
    ("banded_3000", 1, "s"), ("banded_3000", 4, "d"), ("banded_3000", 6, "d"),
    ("banded_3000", 31, "d"),
    ("rmat_9", 2, "d"), ("rmat_9", 4, "d"), ("rmat_9", 8, "d"),
    ("rmat_9", 8, "s"),
    ("ragged_333", 1, "d"), ("ragged_333", 3, "d"), ("ragged_333", 5, "d"),
    ("ragged_64", 2, "d"), ("ragged_64", 4, "s"),
    ("diag_only_48", 1, "d"), ("diag_only_48", 3, "d"),
]


def case_id(case):
    return "%s-P%d-%s" % case


def golden_path(case):
    return os.path.join(GOLDEN_DIR, case_id(case) + ".npz")


def dtype_of(prec):
    return np.float64 if prec == "d" else np.float32


_cache = {}


def matrix(name):
    if name not in _cache:
        rp, ci, v = MATRICES[name]()
        _cache[name] = (np.ascontiguousarray(rp, np.int32),
                        np.ascontiguousarray(ci, np.int32),
                        np.ascontiguousarray(v, np.float64))
    return _cache[name]


def case_inputs(case):
    name, P, prec = case
    rp, ci, v = matrix(name)
    dt = dtype_of(prec)
    x = gen.gen_x(XSEED, len(rp) - 1, dtype=dt)
    return rp, ci, v.astype(dt), x


def normwise_rel_err(y, y_ref):
    y = np.asarray(y, np.float64)
    y_ref = np.asarray(y_ref, np.float64)
    d = np.linalg.norm(y_ref)
    return np.linalg.norm(y - y_ref) / (d if d > 0 else 1.0)


# north_star tolerances (BASELINE.json): normwise relative error vs the
# reference's CFS kernel
TOL = {"d": 1e-12, "s": 1e-5}


def gold_array(gold, key):
    """a dump omits arrays the reference never allocated (row_split_ at P=1)"""
    return gold[key] if key in gold.files else np.zeros(0, np.int32)


# ---- the non-symmetric path (Format::csr): reference partition_by_nnz /
# partition_by_nrows + cpu_mv (csr_matrix.tpp:404-541, 2665-2704) ----------
def general_matrix(name):
    """a NON-symmetric matrix derived from MATRICES[name]: every fifth
    off-diagonal entry dropped (pattern no longer symmetric), values skewed"""
    key = "general:" + name
    if key not in _cache:
        rp, ci, v = matrix(name)
        n = len(rp) - 1
        rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
        h = (rows * 2654435761 + ci.astype(np.int64) * 40503) % 1000003
        keep = (rows == ci) | (h % 5 != 0)
        vals = v * (1.0 + (h % 17) / 16.0)
        counts = np.bincount(rows[keep], minlength=n)
        rp2 = np.zeros(n + 1, np.int32)
        np.cumsum(counts, out=rp2[1:])
        _cache[key] = (rp2, np.ascontiguousarray(ci[keep], np.int32),
                       np.ascontiguousarray(vals[keep], np.float64))
    return _cache[key]


# (matrix, P, precision, tuning): A = Tuning::Aggressive (partition_by_nnz),
# N = Tuning::None (partition_by_nrows)
CSR_CASES = [
    ("lap7_12", 1, "d", "A"), ("lap7_12", 4, "d", "A"), ("lap7_12", 4, "d", "N"),
    ("lap7_12", 27, "d", "A"), ("lap27_10", 8, "d", "A"),
    ("lap27_10", 16, "s", "A"), ("lap27_14", 43, "d", "A"),
    ("banded_3000", 6, "d", "A"), ("banded_3000", 31, "d", "N"),
    ("rmat_9", 2, "d", "A"), ("rmat_9", 8, "d", "A"), ("rmat_9", 8, "s", "A"),
    ("rmat_9", 16, "d", "A"), ("ragged_333", 3, "d", "A"),
    ("ragged_333", 5, "d", "N"), ("ragged_64", 2, "d", "A"),
    ("ragged_64", 4, "d", "A"), ("diag_only_48", 3, "d", "A"),
]


def csr_case_id(case):
    return "csr-%s-P%d-%s-%s" % case


def csr_golden_path(case):
    return os.path.join(GOLDEN_DIR, "csr", csr_case_id(case) + ".npz")
