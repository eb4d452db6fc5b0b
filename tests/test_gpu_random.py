"""Seeded random matrices through every execution path, against the oracle:
structures the named cases do not cover -- random bandwidths (the tile kernel's
applicability boundary), mixed regular / irregular slices, quantised values
(value dictionary of every size), long rows, empty rows -- in double and
single, symmetric and Format::csr."""
import numpy as np
import pytest
import scipy.sparse as sp

import cases
from cfs_spmv_b200 import capi, gen
from oracle import oracle

pytestmark = pytest.mark.gpu


def random_banded(rng, n, bw, per_row, nvalues):
    """symmetric, full diagonal, lower entries at random offsets <= bw, values
    drawn from `nvalues` levels (0 = continuous)"""
    rows, cols = [], []
    for i in range(1, n):
        k = min(rng.poisson(per_row), i, bw)
        if k:
            off = rng.choice(np.arange(1, min(bw, i) + 1), size=k, replace=False)
            rows.append(np.full(k, i))
            cols.append(i - off)
    hi = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    lo = np.concatenate(cols) if cols else np.zeros(0, np.int64)
    if nvalues:
        val = -(1.0 + rng.integers(0, nvalues, len(hi)) / 64.0)
    else:
        val = rng.uniform(-1, 1, len(hi))
    A = sp.coo_matrix((np.concatenate([val, val, np.full(n, 40.0)]),
                       (np.concatenate([hi, lo, np.arange(n)]),
                        np.concatenate([lo, hi, np.arange(n)]))),
                      shape=(n, n)).tocsr()
    A.sort_indices()
    return (A.indptr.astype(np.int32), A.indices.astype(np.int32),
            A.data.astype(np.float64))


SETTINGS = [  # (tile6, value_index, hubs)
    (1, 1, 1), (0, 1, 1), (1, 0, 0), (0, 0, 0)]


@pytest.mark.parametrize("seed", range(12))
def test_random_symmetric_matrices(gpu, seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1500, 40000))
    bw = int(rng.choice([3, 40, 700, 2500, 9000, n - 1]))
    per_row = float(rng.choice([0.5, 3, 9, 40]))
    nvalues = int(rng.choice([0, 1, 2, 17, 256, 300]))
    rp, ci, v = random_banded(rng, n, min(bw, n - 1), per_row, nvalues)
    prec = "d" if seed % 3 else "s"
    dt = cases.dtype_of(prec)
    v = v.astype(dt)
    x = gen.gen_x(seed, n, dtype=dt)
    ref = oracle.Oracle(rp, ci, v, 1).spmv(x)
    for tile6, vi, hubs in SETTINGS:
        capi.set_option("tile6", tile6)
        capi.set_option("value_index", vi)
        capi.set_option("hubs", hubs)
        try:
            A = capi.Matrix.from_csr(rp, ci, v)
            A.tune(1)
            y = np.full(n, 2.5, dtype=dt)
            A.spmv(y, x)
            A.spmv(y, x)
            err = cases.normwise_rel_err(y, ref)
            assert err <= cases.TOL[prec], (seed, tile6, vi, hubs, A.info())
            A.close()
        finally:
            capi.set_option("tile6", 1)
            capi.set_option("value_index", 1)
            capi.set_option("hubs", 1)


@pytest.mark.parametrize("seed", range(6))
def test_random_general_matrices(gpu, seed):
    """Format::csr on rectangular random matrices with empty and long rows:
    bit-identical to cpu_mv where no row exceeds 32 entries"""
    rng = np.random.default_rng(2000 + seed)
    nrows, ncols = int(rng.integers(50, 30000)), int(rng.integers(50, 30000))
    lam = float(rng.choice([0.3, 4, 12]))
    counts = np.minimum(rng.poisson(lam, nrows), ncols)
    if seed % 2:
        counts[rng.integers(0, nrows, 3)] = min(ncols, 700)   # long rows
    rp = np.zeros(nrows + 1, np.int32)
    np.cumsum(counts, out=rp[1:])
    ci = np.concatenate([np.sort(rng.choice(ncols, c, replace=False))
                         for c in counts] + [np.zeros(0, np.int64)]).astype(np.int32)
    v = rng.standard_normal(int(rp[-1]))
    x = rng.standard_normal(ncols)
    P = int(rng.choice([1, 2, 7, 16]))
    A = capi.Matrix(nrows, ncols, rp, ci, v, True, False)
    A.tune(P, tuning=1)
    if P > 1:
        assert np.array_equal(A.export("row_split"),
                              oracle.partition_by_nnz(rp, P))
    y = np.full(nrows, 9.0)
    A.spmv(y, x)
    ref = oracle.csr_spmv(rp, ci, v, x)
    assert cases.normwise_rel_err(y, ref) <= 1e-12
    if counts.max() <= 32:
        assert y.tobytes() == ref.tobytes()
    A.close()
