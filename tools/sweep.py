"""Times the SpMV kernel variants on one GPU (development aid).
usage: python tools/sweep.py [n=200] [points=27] [dtype=d]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cfs_spmv_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    points = int(sys.argv[2]) if len(sys.argv) > 2 else 27
    is_double = (sys.argv[3] if len(sys.argv) > 3 else "d") == "d"
    capi.init(0)
    capi.set_option("keep_layouts", 1)  # variants are switched after tune
    spec = capi.GenSpec.laplacian(points, n, n, n)
    N = spec.nrows
    rp, ci, v = capi.gen_device_csr(spec, is_double=is_double)
    A = capi.Matrix(N, N, rp, ci, v, is_double, True)
    A.tune(1)
    del rp, ci, v
    torch.cuda.empty_cache()
    inf = A.info()
    x = capi.gen_device_x(1, 0, N, is_double)
    y = torch.zeros_like(x)
    ref = None
    configs = [("v1", 1, 2, 0), ("v2 ctas=4", 2, 4, 0), ("v3 ctas=2", 3, 2, 0),
               ("v3 ctas=3", 3, 3, 0), ("v4", 4, 1, 0), ("v5", 5, 1, 0),
               ("v5 streamed values", 5, 1, 0)]
    print("value dictionary: %d distinct values" % inf["value_dictionary"])
    print("regular slices %d of %d, index rows %d vs %d" % (
        inf["regular_slices"], inf["nslices"], inf["index_rows"],
        inf["padded_entries"] // 32))
    print("far entries: %d of %d (%.2f%%), tiles %d" % (
        inf["far_entries"], inf["nnz_low"],
        100.0 * inf["far_entries"] / max(inf["nnz_low"], 1), inf["ntiles"]))
    if os.environ.get("SWEEP_DIAG"):
        configs += [("v%d noRED" % v, v, 2, 1) for v in (1, 2)]
        configs += [("v%d noGATHER" % v, v, 2, 2) for v in (1, 2)]
        configs += [("v%d stream-only" % v, v, 2, 3) for v in (1, 2)]
    only = os.environ.get("SWEEP_ONLY")
    if only:
        configs = [c for c in configs if c[0].startswith(only)]
    iters = int(os.environ.get("SWEEP_ITERS", "50"))
    for name, variant, ctas, mode in configs:
        capi.set_option("spmv_variant", variant)
        capi.set_option("value_index", 0 if "streamed" in name else 1)
        capi.set_option("ctas_per_sm", ctas)
        capi.set_option("diag_mode", mode)
        A.spmv_timed(y, x, 3)
        tot, kern = A.spmv_timed(y, x, iters)
        if ref is None:
            ref = y.clone()
        err = (torch.linalg.norm(y - ref) / torch.linalg.norm(ref)).item()
        us = kern / iters * 1e3
        print("%-12s kernel %8.1f us  step %8.1f us  %7.1f GB/s alg  "
              "%6.1f GFLOP/s  relerr vs v1 %.2e" % (
                  name, us, tot / iters * 1e3,
                  inf["algorithmic_bytes"] / us / 1e3,
                  2 * inf["nnz_full"] / us / 1e3, err), flush=True)


if __name__ == "__main__":
    main()
