"""Host-vector SpMV (cfs_cuda_spmv with pinned x, y) on config 2 under different
pipeline settings (development aid): chunk count x head/rest split x graph
replay, and the pipeline with its components switched off."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cfs_spmv_b200 import capi  # noqa: E402


def timed(A, y, x, steps=30):
    for _ in range(3):
        A.spmv(y, x)
    t0 = time.perf_counter()
    for _ in range(steps):
        A.spmv(y, x)
    return (time.perf_counter() - t0) * 1e3 / steps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    capi.init(0)
    spec = capi.GenSpec.laplacian(27, n, n, n)
    N = spec.nrows
    x = capi.gen_device_x(1, 0, N, True).cpu().pin_memory()
    y = torch.empty_like(x).pin_memory()
    ref = None
    for chunks, split, taper in ((6, 1, 0), (6, 1, 1), (8, 1, 0), (8, 1, 1),
                                 (10, 1, 1), (12, 1, 1), (8, 0, 1)):
        capi.set_option("pipeline_chunks", chunks)
        capi.set_option("pipeline_split", split)
        capi.set_option("pipeline_taper", taper)
        rp, ci, v = capi.gen_device_csr(spec, is_double=True)
        A = capi.Matrix(N, N, rp, ci, v, True, True)
        A.tune(1)
        del rp, ci, v
        nnz = A.info()["nnz_full"]
        ms = timed(A, y, x)
        if ref is None:
            ref = y.clone()
        err = (y - ref).abs().max().item()
        print("chunks %3d split %d taper %d: %.3f ms/step  %.1f GFLOP/s  "
              "maxdiff %.1e" % (chunks, split, taper, ms, 2 * nnz / ms / 1e6,
                                err), flush=True)
        A.close()
    capi.set_option("pipeline_taper", 0)
    capi.set_option("pipeline_chunks", 8)
    capi.set_option("pipeline_split", 1)
    capi.set_option("pipeline_graph", 0)
    rp, ci, v = capi.gen_device_csr(spec, is_double=True)
    A = capi.Matrix(N, N, rp, ci, v, True, True)
    A.tune(1)
    del rp, ci, v
    for skip, name in ((0, "all"), (1, "no kernels"), (2, "no D2H"),
                       (4, "no H2D"), (3, "H2D only"), (5, "D2H only"),
                       (6, "kernels only")):
        capi.set_option("pipeline_skip", skip)
        print("8 chunks split, no graph, %-12s: %.3f ms/step"
              % (name, timed(A, y, x)), flush=True)
    capi.set_option("pipeline_skip", 0)
    capi.set_option("pipeline", 0)
    print("unpipelined: %.3f ms/step, maxdiff %.1e" %
          (timed(A, y, x), (y - ref).abs().max().item()))
    A.close()


if __name__ == "__main__":
    main()
