"""Conjugate gradients on the BASELINE configs[1] matrix (27-point Laplacian
200^3, double), device vectors: time per iteration next to the SpMV alone.
usage: python tools/cg_bench.py [n=200] [iters=300]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cfs_spmv_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    capi.init(0)
    spec = capi.GenSpec.laplacian(27, n, n, n)
    N = spec.nrows
    rp, ci, v = capi.gen_device_csr(spec, is_double=True)
    A = capi.Matrix(N, N, rp, ci, v, True, True)
    A.tune(1)
    del rp, ci, v
    torch.cuda.empty_cache()
    inf = A.info()
    xs = capi.gen_device_x(1, 0, N, True)
    b = torch.zeros_like(xs)
    A.spmv_async(b, xs, 0)
    torch.cuda.synchronize()
    A.spmv_timed(b.clone(), xs, 5)
    tot, kern = A.spmv_timed(torch.zeros_like(xs), xs, 100)
    spmv_us = tot / 100 * 1e3
    capi.set_option("cg_batch", 64)
    x = torch.zeros_like(xs)
    A.cg_solve(x, b, 20, 0.0)                       # warm-up
    x.zero_()
    res = A.cg_solve(x, b, iters, 0.0, want_history=True)
    it_us = res["ms_total"] * 1e3 / (res["executed"] + 1)  # + the initial SpMV
    # bytes one iteration must move: the SpMV's algorithmic bytes + the memset
    # of q folded into the p update: xr kernel 4 reads + 2 writes, p kernel 2
    # reads + 2 writes of N-vectors
    vec = N * 8
    alg = inf["algorithmic_bytes"] + 10 * vec
    out = {
        "workload": "CG on 27-pt Laplacian %d^3, double" % n,
        "iterations": res["executed"], "us_per_iteration": round(it_us, 1),
        "spmv_step_us": round(spmv_us, 1),
        "iteration_over_spmv": round(it_us / spmv_us, 3),
        "algorithmic_gbs": round(alg / it_us / 1e3, 1),
        "residual_drop": res["history"][-1] / res["history"][0],
        "true_error": (torch.linalg.norm(x - xs) / torch.linalg.norm(xs)).item(),
    }
    print(json.dumps(out))
    # solve to 1e-10
    x.zero_()
    res = A.cg_solve(x, b, 5000, 1e-10)
    print(json.dumps({"solve_to_1e-10": {k: res[k] for k in
                                          ("iterations", "executed", "converged",
                                           "ms_total")},
                      "true_error": (torch.linalg.norm(x - xs) /
                                     torch.linalg.norm(xs)).item()}))


if __name__ == "__main__":
    main()
