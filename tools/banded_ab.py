"""Banded matrix (BASELINE configs[3] family), variant 6 with and without the
bank-aware slot assignment (development aid): python tools/banded_ab.py [rows]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cfs_spmv_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000000
    capi.init(0)
    spec = capi.GenSpec.banded(n, 2000, 152, 7)
    rp, ci, v = capi.gen_device_csr(spec)
    x = capi.gen_device_x(1, 0, n)
    ref = None
    for is_double in (True, False):
        vv = v if is_double else v.float()
        xx = x if is_double else x.float()
        for banks in (0, 1, 0, 1):
            capi.set_option("slot_banks", banks)
            A = capi.Matrix(n, n, rp, ci, vv, is_double, True)
            A.tune(1)
            y = torch.zeros_like(xx)
            A.spmv_timed(y, xx, 5)
            tot, kern = A.spmv_timed(y, xx, 100)
            if ref is None:
                ref = y.double().clone()
            err = (torch.linalg.norm(y.double() - ref) / torch.linalg.norm(ref)).item()
            inf = A.info()
            us = kern / 100 * 1e3
            print("%s slot_banks=%d  kernel %7.1f us  %6.1f GB/s alg  tiles %d  "
                  "relerr %.1e" % ("f64" if is_double else "f32", banks, us,
                                   inf["algorithmic_bytes"] / us / 1e3,
                                   inf["transposed_tiles"], err), flush=True)
            A.close()
    capi.set_option("slot_banks", 1)


if __name__ == "__main__":
    main()
