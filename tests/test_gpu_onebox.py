"""One process, several GPUs (cfs_cuda_multi_*, CFS_NUM_GPUS): nnz-balanced row
blocks behind the reference's own API. y against the CPU oracle for the fused
NVLink halo (stencil, banded) and for the strip reduction (R-MAT, whose halos
reach down to row 0), and the reference's unmodified test / bench programs under
CFS_NUM_GPUS. Needs >= 2 GPUs on the box (skipped otherwise)."""
import os
import subprocess

import numpy as np
import pytest

import cases
from cfs_spmv_b200 import capi, gen
from oracle import oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "build", "dropin")


def _matrices():
    yield "lap27", capi.gen_host_csr(capi.GenSpec.laplacian(27, 40, 40, 48, 7)), 1
    yield "lap7", capi.gen_host_csr(capi.GenSpec.laplacian(7, 33, 17, 60)), 1
    yield "banded", capi.gen_host_csr(capi.GenSpec.banded(60000, 700, 152, 3)), 1
    yield "rmat", gen.rmat(14, 8, 1), 0
    yield "rmat_18", gen.rmat(18, 8, 1), 0
    yield "ragged", gen.random_symmetric(5000, 6, 11), 0


@pytest.mark.parametrize("ngpus", [2, 3, 4, 8])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_row_blocks_over_the_gpus_of_one_process(gpu, ngpus, dtype):
    if gpu.device_count() < ngpus:
        pytest.skip("needs %d GPUs" % ngpus)
    tol = 1e-12 if dtype == np.float64 else 1e-5
    for name, (rp, ci, v), fused in _matrices():
        v = v.astype(dtype)
        n = len(rp) - 1
        A = capi.MultiMatrix(rp, ci, v, ngpus)
        inf = A.info()
        # with two GPUs every halo lies inside the one block below
        assert inf["ngpus"] == ngpus, (name, inf)
        assert inf["fused_halo"] == (1 if ngpus == 2 else fused), (name, inf)
        # contiguous, 16-row aligned, nnz-balanced blocks
        assert inf["row_begin"][0] == 0 and inf["row_end"][-1] == n
        assert inf["row_begin"][1:] == inf["row_end"][:-1]
        assert all(b % 16 == 0 for b in inf["row_begin"])
        assert sum(inf["shard_nnz_low"]) == inf["nnz_low"]
        if name in ("lap27", "banded"):
            mean = inf["nnz_low"] / ngpus
            assert max(inf["shard_nnz_low"]) <= 1.05 * mean + 2000, (name, inf)
        x = gen.gen_x(5, n, dtype)
        ref = oracle.Oracle(rp, ci, v, 1).spmv(x)
        y = np.full(n, -1.0, dtype)
        for _ in range(3):  # the buffers are reused
            A.spmv(y, x)
            assert cases.normwise_rel_err(y, ref) <= tol, (name, ngpus)
        A.close()
    capi.init(0)


def _run(binary, args, env):
    return subprocess.run([os.path.join(DROPIN, binary)] + args,
                          env=dict(os.environ, **env), capture_output=True,
                          text=True, timeout=600)


@pytest.mark.parametrize("ngpus", [2, 4, 8])
def test_reference_programs_on_several_gpus(gpu, ngpus, tmp_path):
    if gpu.device_count() < ngpus:
        pytest.skip("needs %d GPUs" % ngpus)
    if not os.path.exists(os.path.join(DROPIN, "test_spmv_mmf")):
        pytest.skip("reference sources were not available at build time")
    rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(27, 30, 30, 30, 3))
    path = str(tmp_path / "lap27_30.mtx")
    gen.write_mtx(path, rp, ci, v)
    for alloc in ("managed", "pinned"):
        env = {"CFS_NUM_GPUS": str(ngpus), "CFS_NUM_THREADS": "4",
               "CFS_GPU_ALLOC": alloc}
        for fmt in ("0", "1", "2"):
            r = _run("test_spmv_mmf", [path, fmt], env)
            assert r.returncode == 0 and "PASSED!" in r.stdout, \
                (alloc, fmt, r.stdout[-2000:], r.stderr[-2000:])
        r = _run("bench_spmv_mmf_dp", [path, "1", "8"], env)
        assert r.returncode == 0 and "gflops/s" in r.stdout, r.stdout + r.stderr
        r = _run("api_consumer", [path, "0"], env)
        assert r.returncode == 0 and "ALL PASSED!" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("ngpus", [2, 4])
def test_unified_memory_vectors_are_used_in_place(gpu, ngpus):
    """managed x / y (what internal_alloc hands out) + fused halos: the GPUs
    read x and write y where they are; the host rewrites x and reads y between
    calls"""
    import ctypes
    if gpu.device_count() < ngpus:
        pytest.skip("needs %d GPUs" % ngpus)
    L = capi.lib()
    for name, spec in (("lap27", capi.GenSpec.laplacian(27, 48, 40, 64, 7)),
                       ("banded", capi.GenSpec.banded(200000, 700, 152, 3))):
        rp, ci, v = capi.gen_host_csr(spec)
        n = len(rp) - 1
        o = oracle.Oracle(rp, ci, v, 1)
        A = capi.MultiMatrix(rp, ci, v, ngpus)
        assert A.info()["fused_halo"] == 1
        px = L.cfs_cuda_host_alloc_kind(n * 8, capi.CFS_ALLOC_MANAGED)
        py = L.cfs_cuda_host_alloc_kind(n * 8, capi.CFS_ALLOC_MANAGED)
        x = np.ctypeslib.as_array(ctypes.cast(px, ctypes.POINTER(ctypes.c_double)), (n,))
        y = np.ctypeslib.as_array(ctypes.cast(py, ctypes.POINTER(ctypes.c_double)), (n,))
        x[:] = gen.gen_x(2, n)
        y[:] = 5.0
        for rep in range(5):
            A.spmv(py, px)
            ref = o.spmv(np.array(x))
            assert cases.normwise_rel_err(np.array(y), ref) <= 1e-12, (name, rep)
            if rep % 2 == 0:        # the host rewrites x and scribbles over y
                x *= 1.25
                y[::5] = -1.0
        for opt in (0, 1):          # the copying path gives the same
            capi.set_option("multi_zero_copy", opt)
            A.spmv(py, px)
            assert cases.normwise_rel_err(np.array(y), o.spmv(np.array(x))) <= 1e-12
        A.close()
        L.cfs_cuda_host_free(px)
        L.cfs_cuda_host_free(py)
    capi.init(0)
