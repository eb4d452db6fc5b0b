"""The vectors of the drop-in: memory from cfs_cuda_host_alloc (what
internal_alloc hands the reference's bench / test, bench_spmv_mmf.cpp:127-133,
test_spmv_mmf.cpp:70-71) in all three kinds. cfs_cuda_spmv must give the
reference's y whether the vectors are unified memory that lives in HBM (default),
page-locked or plain host memory -- also when the host rewrites x between calls
and reads y after each, which is the page-migration path in both directions."""
import ctypes

import numpy as np
import pytest

import cases
from cfs_spmv_b200 import capi, gen
from oracle import oracle

pytestmark = pytest.mark.gpu

KINDS = {"managed": capi.CFS_ALLOC_MANAGED, "pinned": capi.CFS_ALLOC_PINNED,
         "plain": capi.CFS_ALLOC_PLAIN, "default": capi.CFS_ALLOC_DEFAULT}


def _vector(n, dtype, kind):
    nbytes = n * np.dtype(dtype).itemsize
    p = capi.lib().cfs_cuda_host_alloc_kind(nbytes, kind)
    assert p and p % 64 == 0
    ctype = ctypes.c_double if dtype == np.float64 else ctypes.c_float
    a = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctype)), shape=(n,))
    return p, a


@pytest.mark.parametrize("kind", sorted(KINDS))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("prefetch", [0, 1, 2])
def test_allocator_vectors(gpu, kind, dtype, prefetch):
    if prefetch != 1 and kind not in ("managed", "default"):
        pytest.skip("managed_prefetch only concerns unified memory")
    capi.set_option("managed_prefetch", prefetch)
    try:
        rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(27, 40, 30, 20))
        v = v.astype(dtype)
        n = len(rp) - 1
        tol = 1e-12 if dtype == np.float64 else 1e-5
        A = capi.Matrix.from_csr(rp, ci, v)
        A.tune(4)
        o = oracle.Oracle(rp, ci, v, 4)
        px, x = _vector(n, dtype, KINDS[kind])
        py, y = _vector(n, dtype, KINDS[kind])
        x[:] = gen.gen_x(5, n, dtype)
        y[:] = -3.0                      # dirty y: fully overwritten
        for rep in range(4):
            A.spmv(py, px)               # raw addresses, like the C++ layer
            ref = o.spmv(np.array(x))
            assert cases.normwise_rel_err(np.array(y), ref) <= tol, (kind, rep)
            x *= dtype(1.5)              # the host rewrites x ...
            x[rep::7] += dtype(0.25)
            y[::3] = 99.0                # ... and scribbles over y
        A.close()
        capi.lib().cfs_cuda_host_free(px)
        capi.lib().cfs_cuda_host_free(py)
    finally:
        capi.set_option("managed_prefetch", 1)


def test_vector_prefetch_is_optional_and_harmless(gpu):
    n = 300000
    p, a = _vector(n, np.float64, capi.CFS_ALLOC_MANAGED)
    a[:] = np.arange(n)
    capi.check(capi.lib().cfs_cuda_vector_prefetch(p, n * 8, 1))
    capi.check(capi.lib().cfs_cuda_vector_prefetch(p, n * 8, 0))
    assert a[-1] == n - 1 and a.sum() == n * (n - 1) / 2
    q, b = _vector(n, np.float64, capi.CFS_ALLOC_PINNED)
    capi.check(capi.lib().cfs_cuda_vector_prefetch(q, n * 8, 1))  # no-op
    capi.lib().cfs_cuda_host_free(p)
    capi.lib().cfs_cuda_host_free(q)


def test_managed_matrix_arrays_are_copied_not_borrowed(gpu):
    """a caller may take its CSR arrays from internal_alloc too"""
    rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(7, 12, 11, 10))
    n, nnz = len(rp) - 1, len(ci)
    L = capi.lib()
    prp = L.cfs_cuda_host_alloc_kind((n + 1) * 4, capi.CFS_ALLOC_MANAGED)
    pci = L.cfs_cuda_host_alloc_kind(nnz * 4, capi.CFS_ALLOC_MANAGED)
    pv = L.cfs_cuda_host_alloc_kind(nnz * 8, capi.CFS_ALLOC_MANAGED)
    ctypes.memmove(prp, rp.ctypes.data, (n + 1) * 4)
    ctypes.memmove(pci, ci.ctypes.data, nnz * 4)
    ctypes.memmove(pv, v.ctypes.data, nnz * 8)
    A = capi.Matrix(n, n, prp, pci, pv, True, True)
    A.tune(2)
    x = gen.gen_x(3, n, np.float64)
    y = np.zeros(n)
    A.spmv(y, x)
    ref = oracle.Oracle(rp, ci, v, 2).spmv(x)
    assert cases.normwise_rel_err(y, ref) <= 1e-12
    A.close()
    for p in (prp, pci, pv):
        L.cfs_cuda_host_free(p)
