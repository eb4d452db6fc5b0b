"""y against the REFERENCE at the full sizes of BASELINE.json, not only through
size-independent properties: the unmodified reference compiled into oracle/_ref
runs the same matrix and the same x on the host cores (ref_tool bench ... y_out)
and the GPU's y must agree within the tolerance BASELINE.json states -- normwise
1e-12 (double) / 1e-5 (single).

  configs[1]  27-point stencil 200^3, double, per-edge coefficients and the
              constant-coefficient Laplacian (the value-indexed kernel)
  configs[3]  banded SPD matrix, 8 M rows (one GPU's share of the 32 M), double
  configs[2]  symmetric R-MAT scale 24, single (the reference at P = 1: its
              conflict-graph preprocessing is infeasible on a power-law matrix)
"""
import os
import tempfile

import numpy as np
import pytest

from cfs_spmv_b200 import capi, gen
from oracle import oracle

pytestmark = pytest.mark.gpu
XSEED = 1


def _reference_y(spec, n, P, prec, dtype):
    path = os.path.join(tempfile.gettempdir(), "cfs_full_y_%d.bin" % os.getpid())
    try:
        oracle.run_ref_bench(spec, P, prec, XSEED, 1, timeout=1800, warmup=1,
                             y_out=path)
        y = np.fromfile(path, dtype=dtype)
    finally:
        if os.path.exists(path):
            os.remove(path)
    assert len(y) == n
    return y.astype(np.float64)


def _partitions(n):
    P = min(os.cpu_count() or 1, 96)
    while P > 1 and not oracle.valid_partition_count(n, P):
        P -= 1
    return P


def _gpu_y(A, n, is_double):
    import torch
    x = capi.gen_device_x(XSEED, 0, n, is_double)
    y = torch.full_like(x, 3.0)
    A.spmv_async(y, x, 0)
    torch.cuda.synchronize()
    return y.cpu().numpy().astype(np.float64)


def _err(y, ref):
    return float(np.linalg.norm(y - ref) / np.linalg.norm(ref))


@pytest.mark.parametrize("seed", [7, 0], ids=["per_edge_values", "constant"])
def test_config2_27pt_200_cubed_against_the_reference(gpu, seed):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref was not built (no reference tree at build time)")
    import torch
    spec = capi.GenSpec.laplacian(27, 200, 200, 200, seed)
    n = spec.nrows
    rp, ci, v = capi.gen_device_csr(spec)
    A = capi.Matrix(n, n, rp, ci, v, True, True)
    A.tune(1)
    del rp, ci, v
    torch.cuda.empty_cache()
    assert (A.info()["value_dictionary"] == 1) == (seed == 0)
    y = _gpu_y(A, n, True)
    A.close()
    ref = _reference_y(spec.ref_tool_spec(), n, _partitions(n), "d", np.float64)
    assert _err(y, ref) <= 1e-12


def test_config4_banded_8m_rows_against_the_reference(gpu):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref was not built (no reference tree at build time)")
    import torch
    n = 8000000
    spec = capi.GenSpec.banded(n, 2000, 152, 7)
    rp, ci, v = capi.gen_device_csr(spec)
    # the generator tests 2000 offsets per row: the reference gets the matrix
    # the GPU built (array constructor) instead of generating it again
    path = os.path.join(tempfile.gettempdir(), "cfs_banded_8m.bin")
    oracle.write_csr_bin(path, rp.cpu().numpy(), ci.cpu().numpy(),
                         v.cpu().numpy())
    try:
        A = capi.Matrix(n, n, rp, ci, v, True, True)
        A.tune(1)
        del rp, ci, v
        torch.cuda.empty_cache()
        assert A.info()["transposed_tiles"] > 0  # variant 6 is what runs
        y = _gpu_y(A, n, True)
        A.close()
        ref = _reference_y("csr:" + path, n, _partitions(n), "d", np.float64)
    finally:
        os.remove(path)
    assert _err(y, ref) <= 1e-12


def test_config3_rmat_scale_24_single_against_the_reference(gpu):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref was not built (no reference tree at build time)")
    import torch
    rp, ci, v = gen.rmat_torch(24, 8, 1, is_double=False)
    n = rp.numel() - 1
    path = os.path.join(tempfile.gettempdir(), "cfs_rmat_24.bin")
    oracle.write_csr_bin(path, rp.cpu().numpy(), ci.cpu().numpy(),
                         v.cpu().numpy())
    try:
        A = capi.Matrix(n, n, rp, ci, v, False, True)
        A.tune(1)
        del rp, ci, v
        torch.cuda.empty_cache()
        assert A.info()["hub_columns"] > 0
        y = _gpu_y(A, n, False)
        A.close()
        ref = _reference_y("csr:" + path, n, 1, "s", np.float32)
    finally:
        os.remove(path)
    assert _err(y, ref) <= 1e-5
