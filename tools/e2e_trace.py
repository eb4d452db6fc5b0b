"""Prints the per-chunk timeline of one host-vector SpMV (development aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cfs_spmv_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
capi.init(0)
capi.set_option("pipeline_chunks", chunks)
capi.set_option("pipeline_graph", 0)
capi.set_option("pipeline_trace", 1)
spec = capi.GenSpec.laplacian(27, n, n, n)
N = spec.nrows
rp, ci, v = capi.gen_device_csr(spec, is_double=True)
A = capi.Matrix(N, N, rp, ci, v, True, True)
A.tune(1)
x = capi.gen_device_x(1, 0, N, True).cpu().pin_memory()
y = torch.empty_like(x).pin_memory()
for i in range(3):
    print("--- step", i, flush=True)
    A.spmv(y, x)
