"""Conjugate gradients on the device (cfs_cuda_cg_solve)
against the same algorithm on the CPU built on the ORACLE's SpMV
(cpu_mv_sym_serial restated, reference csr_matrix.tpp:2707-2729): same
iterates, same iteration count, true residual below the tolerance."""
import numpy as np
import pytest
import scipy.sparse as sp

import cases
from cfs_spmv_b200 import capi, gen
from oracle import oracle

pytestmark = pytest.mark.gpu


def cpu_cg(spmv, b, x0, max_iters, tol):
    """textbook CG, float64 scalars, vectors in the matrix precision"""
    dt = b.dtype
    x = x0.copy()
    r = (b - spmv(x)).astype(dt)
    p = r.copy()
    rr = float(r.astype(np.float64) @ r.astype(np.float64))
    rr0 = rr
    hist = [np.sqrt(rr)]
    it = 0
    while it < max_iters and rr > tol * tol * rr0:
        q = spmv(p)
        alpha = rr / float(p.astype(np.float64) @ q.astype(np.float64))
        x = (x.astype(np.float64) + alpha * p.astype(np.float64)).astype(dt)
        r = (r.astype(np.float64) - alpha * q.astype(np.float64)).astype(dt)
        rr_new = float(r.astype(np.float64) @ r.astype(np.float64))
        p = (r.astype(np.float64) + (rr_new / rr) * p.astype(np.float64)).astype(dt)
        rr = rr_new
        it += 1
        hist.append(np.sqrt(rr))
    return x, it, np.array(hist)


def spd_cases():
    # 27-point Laplacian (diag 26, boundary rows strictly dominant) and the
    # banded SPD generator (diag = 1 + sum |offdiag|)
    yield "lap27_24", capi.gen_host_csr(capi.GenSpec.laplacian(27, 24, 24, 24))
    yield "lap7_30", capi.gen_host_csr(capi.GenSpec.laplacian(7, 30, 30, 30))
    yield "banded", capi.gen_host_csr(capi.GenSpec.banded(20000, 300, 152, 3))


@pytest.mark.parametrize("name,csr", list(spd_cases()), ids=lambda v: v if isinstance(v, str) else "")
@pytest.mark.parametrize("dt,tol,xtol", [(np.float64, 1e-10, 1e-8),
                                         (np.float32, 1e-5, 2e-3)])
def test_cg_matches_cpu_cg_on_the_oracle(gpu, name, csr, dt, tol, xtol):
    rp, ci, v = csr
    v = v.astype(dt)
    n = len(rp) - 1
    A = capi.Matrix.from_csr(rp, ci, v)
    A.tune(1)
    o = oracle.Oracle(rp, ci, v, 1)
    b = gen.gen_x(3, n, dt)
    x = np.zeros(n, dt)
    res = A.cg_solve(x, b, 2000, tol, want_history=True)
    xc, itc, hist = cpu_cg(lambda z: o.spmv(z), b, np.zeros(n, dt), 2000, tol)
    assert res["converged"] == 1 and res["breakdown"] == 0
    assert abs(res["iterations"] - itc) <= max(2, itc // 20)
    # the first iterations follow the CPU recurrence closely (p'Ap from the
    # fused kernel, r'r, the step lengths)
    k = min(10, len(hist), len(res["history"]))
    assert np.allclose(res["history"][:k], hist[:k],
                       rtol=1e-9 if dt == np.float64 else 1e-3)
    assert np.linalg.norm(x.astype(np.float64) - xc) <= xtol * np.linalg.norm(xc)
    # the TRUE residual, by an independent SpMV
    M = sp.csr_matrix((v.astype(np.float64), ci, rp), shape=(n, n))
    true_res = np.linalg.norm(b - M @ x.astype(np.float64)) / np.linalg.norm(b)
    assert true_res <= (100 * tol if dt == np.float64 else 1e-4)
    assert res["executed"] >= res["iterations"]
    A.close()


def test_cg_device_vectors_initial_guess_and_cap(gpu):
    import torch
    rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(27, 20, 20, 20))
    n = len(rp) - 1
    A = capi.Matrix.from_csr(rp, ci, v)
    A.tune(1)
    M = sp.csr_matrix((v, ci, rp), shape=(n, n))
    xs = gen.gen_x(9, n, np.float64)
    b = M @ xs
    # device vectors, non-zero initial guess
    xd = torch.from_numpy(gen.gen_x(4, n, np.float64)).cuda()
    bd = torch.from_numpy(b).cuda()
    res = A.cg_solve(xd, bd, 3000, 1e-12)
    assert res["converged"] == 1
    assert np.linalg.norm(xd.cpu().numpy() - xs) <= 1e-8 * np.linalg.norm(xs)
    # r0 = 0: nothing to do
    x = np.zeros(n)
    res = A.cg_solve(x, np.zeros(n), 100, 1e-6)
    assert res["iterations"] == 0 and res["converged"] == 1
    assert not x.any()
    # iteration cap
    x = np.zeros(n)
    res = A.cg_solve(x, b, 5, 1e-14)
    assert res["iterations"] == 5 and res["converged"] == 0
    # the iterate is frozen at convergence: more enqueued batches change nothing
    capi.set_option("cg_batch", 1)
    x1 = np.zeros(n)
    r1 = A.cg_solve(x1, b, 3000, 1e-8)
    capi.set_option("cg_batch", 64)
    x2 = np.zeros(n)
    r2 = A.cg_solve(x2, b, 3000, 1e-8)
    capi.set_option("cg_batch", 16)
    assert r1["iterations"] == r2["iterations"]
    assert r2["executed"] >= r2["iterations"]
    assert np.allclose(x1, x2, rtol=0, atol=1e-12 * np.abs(xs).max())
    A.close()


def test_cg_reports_an_indefinite_matrix(gpu):
    rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(7, 12, 12, 12))
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    v = np.where(rows == ci, -1.0, v)   # negative diagonal: not SPD
    n = len(rp) - 1
    A = capi.Matrix.from_csr(rp, ci, v)
    A.tune(1)
    x = np.zeros(n)
    res = A.cg_solve(x, np.ones(n), 50, 1e-10)
    assert res["breakdown"] == 1 and res["converged"] == 0
    A.close()


def test_cg_on_a_ragged_matrix(gpu):
    """R-MAT pattern (hub columns, length-sorted rows: the generic row kernel
    and the column-wise hub kernel) made SPD by its diagonal"""
    rp, ci, v = gen.rmat(12, 8, 1, dtype=np.float64)
    n = len(rp) - 1
    M = sp.csr_matrix((v, ci, rp), shape=(n, n))
    d = np.asarray(abs(M).sum(axis=1)).ravel() + 1.0
    rows = np.repeat(np.arange(n), np.diff(rp))
    v = np.where(rows == ci, d[rows], v)
    M = sp.csr_matrix((v, ci, rp), shape=(n, n))
    A = capi.Matrix.from_csr(rp, ci, v)
    A.tune(1)
    xs = gen.gen_x(5, n, np.float64)
    b = M @ xs
    x = np.zeros(n)
    res = A.cg_solve(x, b, 500, 1e-12)
    assert res["converged"] == 1
    assert np.linalg.norm(x - xs) <= 1e-9 * np.linalg.norm(xs)
    A.close()


def test_distributed_cg_building_blocks_on_one_gpu(gpu):
    """the shard loop (DistributedCG: spmv_halo_dot + update_xr + update_p, the
    scalars reduced by the caller) at world size 1 gives what cfs_cuda_cg_solve
    gives; the multi-GPU run of the same loop is tests/test_gpu_multi.py"""
    import torch
    from cfs_spmv_b200.dist import DistributedCG, ShardedSpMV
    spec = capi.GenSpec.laplacian(27, 40, 40, 40)
    op = ShardedSpMV(spec, 0, 1, is_double=True, xseed=3)
    n = spec.nrows
    xs = op.x_ext.clone()
    op.step()
    b = op.y_owned().clone()
    cg = DistributedCG(op)
    res = cg.solve(b, 3000, 1e-10, check_every=1)
    assert res["converged"] and not res["breakdown"]
    err = (torch.linalg.norm(cg.x - xs) / torch.linalg.norm(xs)).item()
    assert err <= 1e-7
    x2 = torch.zeros(n, dtype=torch.float64, device="cuda")
    ref = op.matrix.cg_solve(x2, b, 3000, 1e-10)
    assert abs(res["iterations"] - ref["iterations"]) <= 2
    assert (torch.linalg.norm(cg.x - x2) / torch.linalg.norm(x2)).item() <= 1e-8
