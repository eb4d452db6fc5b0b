"""R-MAT on several GPUs of one process (strip reduction) against the oracle and
the single-GPU result: python tools/rmat_multi_check.py [scale=22] [ngpus=4]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cfs_spmv_b200 import capi, gen  # noqa: E402
from oracle import oracle  # noqa: E402  (checker)


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    ngpus = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    capi.init(0)
    for dtype in (np.float32, np.float64):
        rp, ci, v = (t.cpu().numpy() for t in gen.rmat_torch(
            scale, 8, 1, is_double=dtype == np.float64))
        n = len(rp) - 1
        x = gen.gen_x(1, n, dtype)
        ref = oracle.Oracle(rp, ci, v, 1).spmv(x).astype(np.float64)
        A1 = capi.Matrix.from_csr(rp, ci, v)
        A1.tune(1)
        y1 = np.zeros(n, dtype)
        A1.spmv(y1, x)
        A1.close()
        t0 = time.time()
        A = capi.MultiMatrix(rp, ci, v, ngpus)
        inf = A.info()
        y = np.zeros(n, dtype)
        for _ in range(3):
            A.spmv(y, x)
        t1 = time.perf_counter()
        for _ in range(10):
            A.spmv(y, x)
        ms = (time.perf_counter() - t1) / 10 * 1e3
        A.close()
        e_ref = np.linalg.norm(y.astype(np.float64) - ref) / np.linalg.norm(ref)
        e_one = np.linalg.norm(y.astype(np.float64) - y1) / np.linalg.norm(ref)
        tol = 1e-12 if dtype == np.float64 else 1e-5
        print("R-MAT scale %d %s on %d GPUs: fused_halo %d, rows %s, shard nnz_low "
              "%s, vs oracle %.2e, vs one GPU %.2e (%s), %.2f ms per call with "
              "host vectors" % (scale, np.dtype(dtype).name, ngpus,
                                inf["fused_halo"], inf["row_end"],
                                inf["shard_nnz_low"], e_ref, e_one,
                                "OK" if e_ref <= tol else "FAIL", ms), flush=True)
    capi.init(0)


if __name__ == "__main__":
    main()
