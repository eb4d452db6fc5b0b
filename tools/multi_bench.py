"""cfs_cuda_multi_spmv (one process, N GPUs, the path behind CFS_NUM_GPUS) on
BASELINE configs[1] with unified-memory vectors, per call, for N = 1, 2, 4, 8
(development aid and source of the numbers in RESULTS.md):
python tools/multi_bench.py [max_gpus] [grid=200]"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from cfs_spmv_b200 import capi, gen  # noqa: E402


def main():
    max_g = int(sys.argv[1]) if len(sys.argv) > 1 else capi.device_count()
    n1 = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    capi.init(0)
    L = capi.lib()
    for name, spec in (("27-pt %d^3 per-edge values" % n1,
                        capi.GenSpec.laplacian(27, n1, n1, n1, 7)),
                       ("banded 8 M rows", capi.GenSpec.banded(8000000, 2000, 152, 7))):
        if name.startswith("banded"):
            import torch
            rp, ci, v = (t.cpu().numpy() for t in capi.gen_device_csr(spec))
        else:
            rp, ci, v = capi.gen_host_csr(spec)
        n = len(rp) - 1
        px = L.cfs_cuda_host_alloc_kind(n * 8, capi.CFS_ALLOC_MANAGED)
        py = L.cfs_cuda_host_alloc_kind(n * 8, capi.CFS_ALLOC_MANAGED)
        x = np.ctypeslib.as_array(ctypes.cast(px, ctypes.POINTER(ctypes.c_double)), (n,))
        y = np.ctypeslib.as_array(ctypes.cast(py, ctypes.POINTER(ctypes.c_double)), (n,))
        x[:] = gen.gen_x(1, n)
        ref = None
        g = 1
        while g <= max_g:
            for zc in (1, 0):
                capi.set_option("multi_zero_copy", zc)
                t0 = time.time()
                A = capi.MultiMatrix(rp, ci, v, g)
                setup = time.time() - t0
                for _ in range(5):
                    A.spmv(py, px)
                iters = 200
                t0 = time.perf_counter()
                for _ in range(iters):
                    A.spmv(py, px)
                us = (time.perf_counter() - t0) / iters * 1e6
                yy = np.array(y)
                if ref is None:
                    ref = yy
                err = np.linalg.norm(yy - ref) / np.linalg.norm(ref)
                print("%-28s %d GPU(s) %-9s %8.1f us per SpMV  %7.1f GFLOP/s  "
                      "setup %.1f s  vs 1 GPU %.1e" % (
                          name, g, "in place" if zc else "copies", us,
                          2.0 * rp[-1] / us / 1e3, setup, err), flush=True)
                A.close()
            g *= 2
        L.cfs_cuda_host_free(px)
        L.cfs_cuda_host_free(py)
    capi.set_option("multi_zero_copy", 1)


if __name__ == "__main__":
    main()
