"""Runs the non-headline BASELINE.json configs on one GPU (development aid and
source of the numbers in RESULTS.md): R-MAT (config 3) and banded (config 4,
single-GPU share)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cfs_spmv_b200 import capi, gen  # noqa: E402


def time_matrix(name, n, rp, ci, v, is_double, iters=50):
    t0 = time.time()
    A = capi.Matrix(n, n, rp, ci, v, is_double, True)
    A.tune(1)
    torch.cuda.synchronize()
    tune_s = time.time() - t0
    del rp, ci, v
    torch.cuda.empty_cache()
    inf = A.info()
    x = capi.gen_device_x(1, 0, n, is_double)
    y = torch.zeros_like(x)
    out = {"config": name, "rows": n, "nnz_full": inf["nnz_full"],
           "nnz_low": inf["nnz_low"], "dtype": "f64" if is_double else "f32",
           "padded_entries": inf["padded_entries"],
           "padding": inf["padded_entries"] / max(inf["nnz_low"], 1),
           "regular_slices": inf["regular_slices"],
           "sort_window": inf["sort_window"], "hub_columns": inf["hub_columns"],
           "transposed_tiles": inf["transposed_tiles"],
           "tile_smem_bytes": inf["tile_smem_bytes"],
           "hub_entries": inf["hub_entries"], "nslices": inf["nslices"],
           "tune_s": round(tune_s, 3)}
    variants = [int(v) for v in os.environ.get("RUN_VARIANTS", "1,5").split(",")]
    for variant in variants:
        capi.set_option("spmv_variant", 5 if variant == 50 else variant)
        capi.set_option("tile6", 0 if variant == 50 else 1)
        A.spmv_timed(y, x, 3)
        tot, kern = A.spmv_timed(y, x, iters)
        us = kern / iters * 1e3
        out["v%d_us" % variant] = round(us, 1)
        out["v%d_gflops" % variant] = round(2 * inf["nnz_full"] / us / 1e3, 1)
        out["v%d_alg_gbs" % variant] = round(inf["algorithmic_bytes"] / us / 1e3, 1)
    # symmetry property: x2'(A x1) == x1'(A x2)
    x2 = capi.gen_device_x(2, 0, n, is_double)
    y2 = torch.zeros_like(x)
    A.spmv_async(y, x, 0)
    A.spmv_async(y2, x2, 0)
    torch.cuda.synchronize()
    a = torch.dot(x2.double(), y.double()).item()
    b = torch.dot(x.double(), y2.double()).item()
    out["symmetry_rel"] = abs(a - b) / abs(a)
    A.close()
    print(json.dumps(out), flush=True)


def main():
    capi.init(0)
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("rmat", "all"):
        scale = int(sys.argv[2]) if len(sys.argv) > 2 else 24
        t0 = time.time()
        rp, ci, v = gen.rmat_torch(scale, 8, 1, is_double=False)
        torch.cuda.synchronize()
        print("rmat scale %d generated in %.1f s" % (scale, time.time() - t0),
              flush=True)
        time_matrix("config3 R-MAT scale %d ef 8 single" % scale, 1 << scale,
                    rp, ci, v, False)
    if what in ("banded", "all"):
        n = int(sys.argv[3]) if len(sys.argv) > 3 else 32000000
        spec = capi.GenSpec.banded(n, 2000, 152, 7)
        t0 = time.time()
        rp, ci, v = capi.gen_device_csr(spec)
        torch.cuda.synchronize()
        print("banded %d generated in %.1f s" % (n, time.time() - t0), flush=True)
        time_matrix("config4 banded %d bw 2000 double (1 GPU)" % n, n, rp, ci,
                    v, True)


if __name__ == "__main__":
    main()
