"""single GPU: cost of the HALO-enabled kernel when no column is remote"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cfs_spmv_b200 import capi
capi.init(0)
spec = capi.GenSpec.laplacian(27, 200, 200, 200)
N = spec.nrows
rp, ci, v = capi.gen_device_csr(spec)
A = capi.Matrix(N, N, rp, ci, v, True, True); A.tune(1)
del rp, ci, v
x = capi.gen_device_x(1, 0, N); y = torch.zeros_like(x)
s = torch.cuda.current_stream().cuda_stream
def timeit(fn, n=100):
    for _ in range(5): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("plain async   %.1f us" % timeit(lambda: A.spmv_async(y, x, s)))
print("halo kernel   %.1f us" % timeit(lambda: A.spmv_halo_async(y, x, y.data_ptr(), False, s)))
print("zero_ + halo(y_is_zero) %.1f us" % timeit(lambda: (y.zero_(), A.spmv_halo_async(y, x, y.data_ptr(), True, s))))
print("zero_ only    %.1f us" % timeit(lambda: y.zero_()))
