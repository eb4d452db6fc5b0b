"""Where the time of the generic row kernel goes on the banded config
(development aid): the kernel with its REDs and/or gathers removed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cfs_spmv_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000000
capi.init(0)
spec = capi.GenSpec.banded(n, 2000, 152, 7)
rp, ci, v = capi.gen_device_csr(spec)
A = capi.Matrix(n, n, rp, ci, v, True, True)
A.tune(1)
del rp, ci, v
torch.cuda.empty_cache()
x = capi.gen_device_x(1, 0, n, True)
y = torch.zeros_like(x)
capi.set_option("spmv_variant", 1)
for mode, name in ((0, "full"), (1, "no REDs"), (2, "no gathers"),
                   (3, "stream only")):
    capi.set_option("diag_mode", mode)
    A.spmv_timed(y, x, 3)
    tot, kern = A.spmv_timed(y, x, 30)
    print("banded %d, variant 1, %-12s: %.1f us" % (n, name, kern / 30 * 1e3),
          flush=True)
capi.set_option("diag_mode", 0)
