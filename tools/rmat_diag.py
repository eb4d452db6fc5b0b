"""Where the time goes on the R-MAT config (development aid): the row kernel
with its REDs and/or gathers removed, with and without the hub columns."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cfs_spmv_b200 import capi, gen  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
capi.init(0)
rp, ci, v = gen.rmat_torch(scale, 8, 1, is_double=False)
n = 1 << scale
for hubs in (1, 0):
    capi.set_option("hubs", hubs)
    A = capi.Matrix(n, n, rp, ci, v, False, True)
    A.tune(1)
    inf = A.info()
    x = capi.gen_device_x(1, 0, n, False)
    y = torch.zeros_like(x)
    capi.set_option("spmv_variant", 1)
    for mode, name in ((0, "full"), (1, "no REDs"), (2, "no gathers"),
                       (3, "stream only")):
        capi.set_option("diag_mode", mode)
        A.spmv_timed(y, x, 3)
        tot, kern = A.spmv_timed(y, x, 20)
        print("rmat %d hubs=%d (%d hub cols, %d hub entries of %d) %-12s: "
              "%.1f us (kernel events) %.1f us (step)" % (
                  scale, hubs, inf["hub_columns"], inf["hub_entries"],
                  inf["nnz_low"], name, kern / 20 * 1e3, tot / 20 * 1e3),
              flush=True)
    capi.set_option("diag_mode", 0)
    A.close()
    torch.cuda.empty_cache()
