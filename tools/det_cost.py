"""What deterministic mode costs next to the default kernels (development aid
and source of the numbers in RESULTS.md): python tools/det_cost.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cfs_spmv_b200 import capi  # noqa: E402


def time_steps(A, y, x, iters):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    for _ in range(5):
        A.spmv_async(y, x, 0)
    e0.record()
    for _ in range(iters):
        A.spmv_async(y, x, 0)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    capi.init(0)
    capi.set_option("keep_layouts", 1)
    for name, spec in (
            ("27-pt 200^3 per-edge values f64", capi.GenSpec.laplacian(27, 200, 200, 200, 7)),
            ("27-pt 200^3 constant f64", capi.GenSpec.laplacian(27, 200, 200, 200)),
            ("banded 8 M rows f64", capi.GenSpec.banded(8000000, 2000, 152, 7))):
        n = spec.nrows
        rp, ci, v = capi.gen_device_csr(spec)
        A = capi.Matrix(n, n, rp, ci, v, True, True)
        A.tune(1)
        del rp, ci, v
        torch.cuda.empty_cache()
        x = capi.gen_device_x(1, 0, n)
        y = torch.zeros_like(x)
        t_def = time_steps(A, y, x, 100)
        y_def = y.clone()
        capi.set_option("deterministic", 1)
        t_det = time_steps(A, y, x, 100)
        y1 = y.clone()
        A.spmv_async(y, x, 0)
        torch.cuda.synchronize()
        capi.set_option("deterministic", 0)
        same = torch.equal(y1.view(torch.int64), y.view(torch.int64))
        err = (torch.linalg.norm(y1 - y_def) / torch.linalg.norm(y_def)).item()
        print("%-34s default (memset + kernel) %7.1f us   deterministic %7.1f us"
              "  (x%.2f)  bitwise repeatable: %s  vs default: %.2e" % (
                  name, t_def, t_det, t_det / t_def, same, err), flush=True)
        A.close()


if __name__ == "__main__":
    main()
