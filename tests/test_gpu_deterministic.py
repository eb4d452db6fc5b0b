"""Deterministic mode (option "deterministic", csrc/det.cu): y is bitwise equal
from run to run -- what the reference guarantees for a fixed thread count
(csr_matrix.tpp:2988-3018: one writer per y entry and phase) and what L2
floating-point reductions in arbitrary order do not -- and still within the
tolerance of the reference's y."""
import numpy as np
import pytest

import cases
from cfs_spmv_b200 import capi, gen
from oracle import oracle

pytestmark = pytest.mark.gpu


def _cases():
    yield "lap27_64_distinct", capi.gen_host_csr(
        capi.GenSpec.laplacian(27, 64, 64, 64, 7))
    yield "lap27_40_constant", capi.gen_host_csr(
        capi.GenSpec.laplacian(27, 40, 40, 40))
    yield "banded_300k", capi.gen_host_csr(
        capi.GenSpec.banded(300000, 2000, 152, 7))
    yield "rmat_13", gen.rmat(13, 8, 1)
    yield "ragged", gen.random_symmetric(3000, 7, 5)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_two_runs_are_bitwise_equal_and_close_to_the_reference(gpu, dtype):
    import torch
    tol = 1e-12 if dtype == np.float64 else 1e-5
    capi.set_option("keep_layouts", 1)
    try:
        for name, (rp, ci, v) in _cases():
            v = v.astype(dtype)
            n = len(rp) - 1
            A = capi.Matrix.from_csr(rp, ci, v)
            A.tune(1)
            x = gen.gen_x(3, n, dtype)
            ref = oracle.Oracle(rp, ci, v, 1).spmv(x)
            xd = torch.from_numpy(x).cuda()
            capi.set_option("deterministic", 1)
            runs = []
            for _ in range(6):
                yd = torch.full_like(xd, 7.0)   # dirty: fully overwritten
                A.spmv_async(yd, xd, 0)
                torch.cuda.synchronize()
                runs.append(yd.cpu().numpy())
            capi.set_option("deterministic", 0)
            for r in runs[1:]:
                assert np.array_equal(r.view(np.uint8), runs[0].view(np.uint8)), \
                    name
            assert cases.normwise_rel_err(runs[0], ref) <= tol, name
            # host vectors take the same route (no pipeline in this mode)
            capi.set_option("deterministic", 1)
            y = np.zeros(n, dtype)
            A.spmv(y, x)
            capi.set_option("deterministic", 0)
            assert np.array_equal(y.view(np.uint8), runs[0].view(np.uint8)), name
            # a different x scale picks a different fixed-point exponent
            capi.set_option("deterministic", 1)
            y2 = np.zeros(n, dtype)
            A.spmv(y2, (x * dtype(1024.0)).astype(dtype))
            capi.set_option("deterministic", 0)
            assert cases.normwise_rel_err(y2, ref * 1024.0) <= tol, name
            A.close()
    finally:
        capi.set_option("deterministic", 0)
        capi.set_option("keep_layouts", 0)


def test_what_the_mode_does_not_cover_fails_loudly(gpu):
    import torch
    rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(7, 12, 12, 12))
    A = capi.Matrix.from_csr(rp, ci, v)
    A.tune(1)
    n = len(rp) - 1
    x = torch.ones(n, dtype=torch.float64, device="cuda")
    b = torch.ones(n, dtype=torch.float64, device="cuda")
    capi.set_option("deterministic", 1)
    try:
        with pytest.raises(capi.CfsError):
            A.cg_solve(x, b, 10, 1e-8)  # its SpMV also returns p'Ap
    finally:
        capi.set_option("deterministic", 0)
        A.close()
