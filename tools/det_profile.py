"""Default and deterministic SpMV launches of the same matrix for ncu (atomic /
reduction counters of both; development aid): python tools/det_profile.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cfs_spmv_b200 import capi  # noqa: E402

capi.init(0)
spec = capi.GenSpec.laplacian(27, 200, 200, 200, 7)
n = spec.nrows
rp, ci, v = capi.gen_device_csr(spec)
A = capi.Matrix(n, n, rp, ci, v, True, True)
A.tune(1)
del rp, ci, v
x = capi.gen_device_x(1, 0, n)
y = torch.zeros_like(x)
for det in (0, 1, 0, 1):
    capi.set_option("deterministic", det)
    A.spmv_async(y, x, 0)
    torch.cuda.synchronize()
capi.set_option("deterministic", 0)
print("done")
