"""Host-vector SpMV (cfs_cuda_spmv with pinned x, y) on config 2 under different
pipeline settings (development aid): chunk count x graph replay."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cfs_spmv_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    capi.init(0)
    spec = capi.GenSpec.laplacian(27, n, n, n)
    N = spec.nrows
    x_dev = capi.gen_device_x(1, 0, N, True)
    x = x_dev.cpu().pin_memory()
    y = torch.empty_like(x).pin_memory()
    ref = None
    capi.set_option("pipeline_ramp", 0)
    capi.set_option("pipeline_chunks", 8)
    capi.set_option("pipeline_graph", 0)
    rp, ci, v = capi.gen_device_csr(spec, is_double=True)
    A = capi.Matrix(N, N, rp, ci, v, True, True)
    A.tune(1)
    del rp, ci, v
    nnz = A.info()["nnz_full"]
    A.spmv(y, x)
    ref = y.clone()
    for skip, name in ((0, "all"), (1, "no kernels"), (2, "no D2H"),
                       (4, "no H2D"), (3, "H2D only"), (5, "D2H only"),
                       (6, "kernels only")):
        capi.set_option("pipeline_skip", skip)
        for _ in range(3):
            A.spmv(y, x)
        steps = 30
        t0 = time.perf_counter()
        for _ in range(steps):
            A.spmv(y, x)
        ms = (time.perf_counter() - t0) * 1e3 / steps
        print("8 chunks, %-12s: %.3f ms/step" % (name, ms), flush=True)
    capi.set_option("pipeline_skip", 0)
    A.close()
    capi.set_option("pipeline", 0)
    rp, ci, v = capi.gen_device_csr(spec, is_double=True)
    A = capi.Matrix(N, N, rp, ci, v, True, True)
    A.tune(1)
    for _ in range(3):
        A.spmv(y, x)
    t0 = time.perf_counter()
    for _ in range(30):
        A.spmv(y, x)
    ms = (time.perf_counter() - t0) * 1e3 / 30
    print("unpipelined: %.3f ms/step, maxdiff %.1e" %
          (ms, (y - ref).abs().max().item()))


if __name__ == "__main__":
    main()
