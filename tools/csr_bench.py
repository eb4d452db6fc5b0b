"""Format::csr (non-symmetric path) on the FULL 27-point Laplacian: sliced
layout (csr_path.cu) vs the warp-per-row comparator kernel, device vectors.
usage: python tools/csr_bench.py [n=200]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cfs_spmv_b200 import capi  # noqa: E402


def timed(A, y, x, iters=30):
    for _ in range(3):
        A.spmv_async(y, x, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        A.spmv_async(y, x, 0)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    capi.init(0)
    spec = capi.GenSpec.laplacian(27, n, n, n)
    N = spec.nrows
    rp, ci, v = capi.gen_device_csr(spec, is_double=True)
    A = capi.Matrix(N, N, rp, ci, v, True, False)
    A.tune(16, 1)
    inf = A.info()
    x = capi.gen_device_x(1, 0, N, True)
    y = torch.zeros_like(x)
    out = {"workload": "Format::csr, FULL 27-pt Laplacian %d^3, double" % n,
           "nnz": inf["nnz_full"], "bytes": inf["algorithmic_bytes"]}
    for layout, name in ((1, "sliced"), (0, "warp_per_row")):
        capi.set_option("csr_layout", layout)
        us = timed(A, y, x)
        out[name + "_us"] = round(us, 1)
        out[name + "_gflops"] = round(2 * inf["nnz_full"] / us / 1e3, 1)
        out[name + "_gbs"] = round(inf["algorithmic_bytes"] / us / 1e3, 1)
        if layout == 1:
            ref = y.clone()
        else:
            out["max_abs_diff"] = (y - ref).abs().max().item()
    capi.set_option("csr_layout", 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
