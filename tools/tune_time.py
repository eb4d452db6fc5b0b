import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cfs_spmv_b200 import capi
capi.init(0)
for pts, n in ((7, 100), (27, 100), (27, 200)):
    spec = capi.GenSpec.laplacian(pts, n, n, n)
    rp, ci, v = capi.gen_device_csr(spec)
    for P in (1, 16, 1, 16, 148):
        A = capi.Matrix(spec.nrows, spec.nrows, rp, ci, v, True, True)
        torch.cuda.synchronize(); t0 = time.time()
        A.tune(P)
        torch.cuda.synchronize(); dt = time.time() - t0
        inf = A.info()
        print("lap%d %d^3 P=%3d tune %.4f s ncolors=%d nranges=%d edges=%d" % (pts, n, P, dt, inf["ncolors"], inf["nranges"], inf["nconflict_edges"]), flush=True)
        A.close()
