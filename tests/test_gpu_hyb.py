"""Format::hyb (csrc/hyb.cu): the reference's split by bandwidth
(csr_matrix.tpp:314-401) carried through -- the band inside the threshold runs
symmetrically, everything else is kept in both triangles and only gathered. The
reference cannot run its own HYB (tune() aborts for P > 1), so the
oracle here is y = A x of the whole matrix and, for the metadata, the reference
pipeline applied to the near part."""
import numpy as np
import pytest

import cases
from cfs_spmv_b200 import capi, gen
from oracle import oracle

pytestmark = pytest.mark.gpu


def _split(rp, ci, v, thr):
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    near = np.abs(ci.astype(np.int64) - rows) < thr
    nrp = np.zeros(len(rp), np.int64)
    np.add.at(nrp, rows[near] + 1, 1)
    return (np.cumsum(nrp).astype(np.int32), ci[near], v[near]), int((~near).sum())


def _matrices():
    yield "lap27_wide", capi.gen_host_csr(capi.GenSpec.laplacian(27, 110, 100, 5, 3)), 10000
    yield "lap27_small_threshold", capi.gen_host_csr(capi.GenSpec.laplacian(27, 20, 20, 20)), 50
    yield "rmat_13", gen.rmat(13, 8, 1), 1000
    yield "ragged", gen.random_symmetric(4000, 6, 9), 300
    yield "all_near", capi.gen_host_csr(capi.GenSpec.laplacian(7, 12, 12, 12)), 10000


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("P", [1, 4])
def test_hybrid_split_gives_the_same_y(gpu, P, dtype):
    tol = 1e-12 if dtype == np.float64 else 1e-5
    for name, (rp, ci, v), thr in _matrices():
        v = v.astype(dtype)
        n = len(rp) - 1
        (nrp, nci, nv), nfar = _split(rp, ci, v, thr)
        A = capi.Matrix.from_csr(rp, ci, v)
        A.set_hybrid(thr)
        A.tune(P)
        inf = A.info()
        assert inf["hyb_far_entries"] == nfar, name
        assert inf["nnz_full"] == len(ci), name       # nnz() stays the full count
        near = oracle.Oracle(nrp, nci, nv, P)
        assert inf["nnz_low"] == near.nnz_low, name
        # size(): the near part's formula + colind_high / values_high
        assert inf["size_bytes"] == near.size_bytes + nfar * (4 + v.itemsize), name
        if P > 1:  # metadata = the reference pipeline on the near part
            md, omd = A.metadata(), near.metadata()
            for k in ("row_split", "range_ptr", "range_start", "range_end",
                      "lower_rowptr", "lower_colind"):
                assert np.array_equal(md[k], omd[k]), (name, k)
        x = gen.gen_x(4, n, dtype)
        ref = oracle.Oracle(rp, ci, v, 1).spmv(x)
        y = np.full(n, 9.0, dtype)
        for _ in range(2):
            A.spmv(y, x)
            assert cases.normwise_rel_err(y, ref) <= tol, (name, P)
        A.close()


def test_hybrid_is_refused_where_it_does_not_apply(gpu):
    rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(7, 8, 8, 8))
    A = capi.Matrix.from_csr(rp, ci, v, symmetric=False)
    with pytest.raises(capi.CfsError):
        A.set_hybrid(100)
    A.tune(1)
    A.close()
    B = capi.Matrix.from_csr(rp, ci, v)
    B.tune(1)
    with pytest.raises(capi.CfsError):
        B.set_hybrid(100)  # after tune
    B.close()
