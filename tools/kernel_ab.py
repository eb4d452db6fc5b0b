"""A/B of the SpMV kernel variants on the general-values 27-point matrix
(development aid, round 2): python tools/kernel_ab.py [n=200] [seed=7]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cfs_spmv_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    is_double = (sys.argv[3] if len(sys.argv) > 3 else "d") == "d"
    capi.init(0)
    capi.set_option("keep_layouts", 1)  # variants are switched after tune
    spec = capi.GenSpec.laplacian(27, n, n, n, seed)
    N = spec.nrows
    rp, ci, v = capi.gen_device_csr(spec, is_double=is_double)
    A = capi.Matrix(N, N, rp, ci, v, is_double, True)
    A.tune(1)
    del rp, ci, v
    torch.cuda.empty_cache()
    inf = A.info()
    print("value dictionary %d, device bytes %.3f GB, algorithmic %.3f GB" % (
        inf["value_dictionary"], inf["device_bytes"] / 1e9,
        inf["algorithmic_bytes"] / 1e9))
    x = capi.gen_device_x(1, 0, N, is_double)
    y = torch.zeros_like(x)
    ref = None
    configs = [("v1", {"spmv_variant": 1}),
               ("v5 no prefetch", {"spmv_variant": 5, "l2_prefetch": 0}),
               ("v5 L2 prefetch", {"spmv_variant": 5, "l2_prefetch": 1}),
               ("v5 pf, 12 CTAs/SM", {"spmv_variant": 5, "l2_prefetch": 1,
                                      "reg_blocks": 12}),
               ("v7 TMA bulk", {"spmv_variant": 7, "reg_blocks": 16})]
    iters = int(os.environ.get("SWEEP_ITERS", "100"))
    for rep in range(2):
        for name, opts in configs:
            for k, val in opts.items():
                capi.set_option(k, val)
            A.spmv_timed(y, x, 5)
            tot, kern = A.spmv_timed(y, x, iters)
            if ref is None:
                ref = y.clone()
            err = (torch.linalg.norm(y - ref) / torch.linalg.norm(ref)).item()
            us = kern / iters * 1e3
            print("%-18s kernel %8.1f us  step %8.1f us  %7.1f GB/s alg  "
                  "%6.1f GFLOP/s  relerr vs v1 %.2e" % (
                      name, us, tot / iters * 1e3,
                      inf["algorithmic_bytes"] / us / 1e3,
                      2 * inf["nnz_full"] / us / 1e3, err), flush=True)


if __name__ == "__main__":
    main()
