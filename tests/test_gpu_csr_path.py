"""The NON-symmetric path (Format::csr) on the GPU against the reference's own
results (tests/golden/csr, dumps of the compiled reference) and the oracle:
row_split_ of partition_by_nnz / partition_by_nrows bit-exact
(csr_matrix.tpp:404-541), y of cpu_mv (:2665-2704) BIT-IDENTICAL for rows of at
most 32 entries, normwise 1e-12 / 1e-5 otherwise."""
import numpy as np
import pytest

import cases
from cfs_spmv_b200 import capi, gen
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", cases.CSR_CASES, ids=cases.csr_case_id)
def test_csr_path_matches_reference(gpu, case):
    name, P, prec, tuning = case
    gold = np.load(cases.csr_golden_path(case))
    rp, ci, v = cases.general_matrix(name)
    dt = cases.dtype_of(prec)
    n = len(rp) - 1
    x = gen.gen_x(cases.XSEED, n, dtype=dt)
    A = capi.Matrix.from_csr(rp, ci, v.astype(dt), symmetric=False)
    A.tune(P, tuning=1 if tuning == "A" else 0)
    if P > 1:
        assert np.array_equal(A.export("row_split"), gold["row_split"])
    y = np.full(n, -7, dtype=dt)
    A.spmv(y, x)
    A.spmv(y, x)  # twice on a dirty y, like test_spmv_mmf.cpp:80-83
    assert cases.normwise_rel_err(y, gold["y"]) <= cases.TOL[prec]
    if np.diff(rp).max() <= 32:
        assert y.tobytes() == gold["y"].tobytes()
    inf = A.info()
    assert inf["size_bytes"] == int(gold["size_bytes"])
    # the warp-per-row comparator kernel agrees too
    capi.set_option("csr_layout", 0)
    y2 = np.zeros(n, dtype=dt)
    A.spmv(y2, x)
    capi.set_option("csr_layout", 1)
    assert cases.normwise_rel_err(y2, gold["y"]) <= cases.TOL[prec]
    A.close()


@pytest.mark.parametrize("P", [2, 3, 7, 16, 64, 148, 1000])
def test_partition_by_nnz_matches_oracle_on_larger_inputs(gpu, P):
    for spec in (capi.GenSpec.laplacian(27, 40, 40, 40),
                 capi.GenSpec.banded(50000, 500, 152, 3)):
        rp, ci, v = capi.gen_host_csr(spec)
        A = capi.Matrix.from_csr(rp, ci, v, symmetric=False)
        A.tune(P, tuning=1)
        assert np.array_equal(A.export("row_split"),
                              oracle.partition_by_nnz(rp, P))
        A.close()
    rp, ci, v = gen.rmat(13, 8, seed=2)
    A = capi.Matrix.from_csr(rp, ci, v, symmetric=False)
    A.tune(P, tuning=1)
    assert np.array_equal(A.export("row_split"), oracle.partition_by_nnz(rp, P))
    x = gen.gen_x(1, len(rp) - 1, np.float64)
    y = np.zeros(len(rp) - 1)
    A.spmv(y, x)
    assert cases.normwise_rel_err(y, oracle.csr_spmv(rp, ci, v, x)) <= 1e-12
    A.close()


def test_partition_by_nnz_degenerate_inputs(gpu):
    # fewer nonzeros than partitions (nnz_per_split == 0), rows without entries
    n = 100
    rp = np.zeros(n + 1, np.int32)
    rp[51:] = 3
    ci = np.array([0, 5, 7], np.int32)
    v = np.array([1.0, 2.0, 3.0])
    for P in (2, 5, 8):
        A = capi.Matrix.from_csr(rp, ci, v, symmetric=False)
        A.tune(P, tuning=1)
        assert np.array_equal(A.export("row_split"),
                              oracle.partition_by_nnz(rp, P))
        y = np.zeros(n)
        A.spmv(y, np.arange(1.0, n + 1))
        assert y[50] == 1 * 1 + 2 * 6 + 3 * 8 and np.count_nonzero(y) == 1
        A.close()


def test_csr_path_rectangular(gpu):
    rng = np.random.default_rng(4)
    nrows, ncols = 700, 300
    counts = rng.integers(0, 9, nrows)
    rp = np.zeros(nrows + 1, np.int32)
    np.cumsum(counts, out=rp[1:])
    ci = np.concatenate([np.sort(rng.choice(ncols, c, replace=False))
                         for c in counts]).astype(np.int32)
    v = rng.standard_normal(rp[-1])
    A = capi.Matrix(nrows, ncols, rp, ci, v, True, False)
    A.tune(4, tuning=1)
    assert np.array_equal(A.export("row_split"), oracle.partition_by_nnz(rp, 4))
    x = rng.standard_normal(ncols)
    y = np.zeros(nrows)
    A.spmv(y, x)
    assert y.tobytes() == oracle.csr_spmv(rp, ci, v, x).tobytes()
    A.close()
