"""CPU-side checks of the C ABI boundary: the shared library loads without a
GPU, exports exactly what include/cfs_cuda.h declares, its host-only entry
points work, and every compute entry point FAILS LOUDLY (no CPU fallback)
when there is no device."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from cfs_spmv_b200 import capi, gen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_in_header():
    text = open(os.path.join(ROOT, "include", "cfs_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cfs_(?:cuda|gen)_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_in_header() == sorted(capi.DECLARED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    for name in _declared_in_header():
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH],
                         capture_output=True, text=True).stdout
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    assert set(_declared_in_header()) <= exported


def test_library_is_built_for_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_torch_or_oracle_in_the_product_library():
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True,
                         text=True).stdout
    assert "torch" not in out and "oracle" not in out


def test_host_generators_match_numpy_restatement():
    x_c = capi.gen_host_x(7, 1000)
    assert np.array_equal(x_c, gen.gen_x(7, 1000))
    assert x_c.min() >= 0.01 and x_c.max() < 0.42
    xf = capi.gen_host_x(7, 100, np.float32)
    assert np.array_equal(xf, gen.gen_x(7, 100, np.float32))


@pytest.mark.parametrize("points,dims", [(7, (5, 4, 3)), (27, (4, 5, 6))])
def test_laplacian_generator_properties(points, dims):
    spec = capi.GenSpec.laplacian(points, *dims)
    rp, ci, v = capi.gen_host_csr(spec)
    n = dims[0] * dims[1] * dims[2]
    assert len(rp) == n + 1
    import scipy.sparse as sp
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    assert abs(A - A.T).max() == 0
    assert np.all(A.diagonal() == (6.0 if points == 7 else 26.0))
    assert np.all(np.diff(ci)[np.diff(ci) <= 0].size <= n)  # sorted per row
    for i in range(n):
        assert np.all(np.diff(ci[rp[i]:rp[i + 1]]) > 0)
    nx, ny, nz = dims
    if points == 7:
        assert rp[-1] == 7 * n - 2 * (nx * ny + ny * nz + nx * nz)
    else:
        assert rp[-1] == (3 * nx - 2) * (3 * ny - 2) * (3 * nz - 2)
    # row shards reproduce the rows of the whole
    r2, c2, v2 = capi.gen_host_csr(spec, 7, 31)
    assert np.array_equal(c2, ci[rp[7]:rp[31]])
    assert np.array_equal(v2, v[rp[7]:rp[31]])


def test_banded_generator_properties():
    spec = capi.GenSpec.banded(2000, 50, 152, 3)
    rp, ci, v = capi.gen_host_csr(spec)
    import scipy.sparse as sp
    A = sp.csr_matrix((v, ci, rp), shape=(2000, 2000))
    assert abs(A - A.T).max() == 0
    coo = A.tocoo()
    assert np.abs(coo.row - coo.col).max() <= 50
    off = A - sp.diags(A.diagonal())
    # strictly diagonally dominant with positive diagonal => SPD
    assert np.all(A.diagonal() > np.abs(off).sum(axis=1).A1)
    low = sp.tril(A, -1).nnz / 2000.0
    assert 8.0 < low < 11.0  # ~9.5 lower entries per row


def test_compute_entry_points_fail_loudly_without_a_gpu():
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.CfsError) as e:
        capi.init(0)
    assert e.value.code == capi.CFS_ERR_NO_DEVICE
    rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(7, 3, 3, 3))
    with pytest.raises(capi.CfsError) as e:
        capi.Matrix.from_csr(rp, ci, v)
    assert e.value.code == capi.CFS_ERR_NO_DEVICE
    assert b"no CPU fallback" in capi.lib().cfs_cuda_last_error() or \
        b"CUDA" in capi.lib().cfs_cuda_last_error()


def test_ingest_and_solver_fail_loudly_without_a_gpu(tmp_path):
    """the entry points next to the path have no CPU fallback either; only the
    header scan (host code, like the reference's loader) works anywhere"""
    path = str(tmp_path / "m.mtx")
    open(path, "w").write("%%MatrixMarket matrix coordinate real symmetric\n"
                          "% comment\n3 3 2\n1 1 2.0\n2 1 1.0\n")
    image = np.fromfile(path, dtype=np.uint8)
    h = capi.HostMmfHeader()
    assert capi.host_lib().cfs_host_scan_mmf_header(
        image.ctypes.data, image.size, ctypes.byref(h)) == 0
    assert (h.nrows, h.ncols, h.declared, h.symmetric, h.zero_based) == \
        (3, 3, 2, 1, 0)
    assert bytes(image[h.entries_offset:h.entries_offset + 3]) == b"1 1"
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.CfsError) as e:
        capi.Matrix.from_mmf(path)
    assert e.value.code == capi.CFS_ERR_NO_DEVICE
    res = capi.CgResult()
    assert capi.lib().cfs_cuda_cg_solve(None, None, None, 1, 1e-6,
                                        ctypes.byref(res), None, 0) == \
        capi.CFS_ERR_INVALID


def test_host_alloc_is_64_byte_aligned_and_freeable():
    L = capi.lib()
    for nbytes in (1, 64, 4096, 1 << 20):
        p = L.cfs_cuda_host_alloc(nbytes)
        assert p and p % 64 == 0
        ctypes.memset(p, 0xAB, nbytes)
        L.cfs_cuda_host_free(p)
    L.cfs_cuda_host_free(None)


@pytest.mark.parametrize("points", [7, 27])
def test_per_edge_coefficient_stencil_is_symmetric_and_dominant(points):
    """seed != 0 (cfs_gen.h): one coefficient per edge -- the general-values
    matrix bench.py measures by default"""
    spec = capi.GenSpec.laplacian(points, 6, 5, 4, 7)
    rp, ci, v = capi.gen_host_csr(spec)
    n = len(rp) - 1
    import scipy.sparse as sp
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    assert abs(A - A.T).max() == 0.0
    off = A - sp.diags(A.diagonal())
    assert np.all(off.data < -0.499) and np.all(off.data >= -1.0)
    rowabs = np.asarray(abs(off).sum(axis=1)).ravel()
    assert np.allclose(A.diagonal(), 0.0625 + rowabs, rtol=1e-15, atol=0)
    assert len(np.unique(off.data)) == off.nnz // 2     # nothing to dictionary-code
    # same pattern as the constant-coefficient stencil
    rp0, ci0, _ = capi.gen_host_csr(capi.GenSpec.laplacian(points, 6, 5, 4))
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0)
    assert spec.ref_tool_spec().endswith(":7")


def test_allocator_kinds_without_a_gpu():
    """every kind degrades to plain 64-byte aligned memory on a host-only box"""
    L = capi.lib()
    for kind in (capi.CFS_ALLOC_DEFAULT, capi.CFS_ALLOC_PLAIN,
                 capi.CFS_ALLOC_PINNED, capi.CFS_ALLOC_MANAGED):
        p = L.cfs_cuda_host_alloc_kind(4096, kind)
        assert p and p % 64 == 0
        ctypes.memset(p, 0x5A, 4096)
        assert L.cfs_cuda_vector_prefetch(p, 4096, 1) == 0   # optional, harmless
        L.cfs_cuda_host_free(p)


def test_new_entry_points_fail_loudly_without_a_gpu():
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(7, 4, 4, 4))
    with pytest.raises(capi.CfsError) as e:
        capi.MultiMatrix(rp, ci, v, 2)
    assert e.value.code == capi.CFS_ERR_NO_DEVICE
    L = capi.lib()
    assert L.cfs_cuda_matrix_set_hybrid(None, 100) == capi.CFS_ERR_INVALID
    assert L.cfs_cuda_spmv_shard_async(None, None, None, None, None, None, 1,
                                       None) == capi.CFS_ERR_INVALID
    assert L.cfs_cuda_multi_spmv(None, None, None) == capi.CFS_ERR_INVALID
    assert L.cfs_cuda_spmv_shard_part_async(None, None, None, None, None, None,
                                            1, None) == capi.CFS_ERR_INVALID
    for key, val in (("deterministic", 1), ("deterministic", 0),
                     ("keep_layouts", 0), ("l2_prefetch", 1),
                     ("managed_prefetch", 1), ("spmv_variant", 7),
                     ("spmv_variant", 5)):
        capi.set_option(key, val)
    with pytest.raises(capi.CfsError):
        capi.set_option("deterministic", 2)


def test_nnz_balanced_row_blocks():
    """the split ShardedSpMV and cfs_cuda_multi_create use (partition_by_nnz
    semantics lifted to GPUs): contiguous, aligned, about equal nnz"""
    from cfs_spmv_b200.dist import nnz_balanced_blocks
    rng = np.random.default_rng(0)
    per_row = rng.integers(1, 60, size=10000)
    per_row[:500] = 400                     # a heavy head
    prefix = np.concatenate([[0], np.cumsum(per_row)])
    for world in (2, 3, 8):
        b = nnz_balanced_blocks(prefix, world)
        assert b[0] == 0 and b[-1] == 10000 and len(b) == world + 1
        assert all(x % 16 == 0 for x in b[:-1]) and b == sorted(b)
        loads = [prefix[b[g + 1]] - prefix[b[g]] for g in range(world)]
        assert max(loads) <= 1.1 * prefix[-1] / world + 400 * 16
