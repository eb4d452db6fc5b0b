"""Virtual-row chunk length vs time of the tile kernel on the banded config
(development aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cfs_spmv_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000000
capi.init(0)
spec = capi.GenSpec.banded(n, 2000, 152, 7)
x = capi.gen_device_x(1, 0, n, True)
y = torch.zeros_like(x)
for pct in (100, 110, 120, 130, 150, 180):
    capi.set_option("rechunk_pct", pct)
    rp, ci, v = capi.gen_device_csr(spec)
    A = capi.Matrix(n, n, rp, ci, v, True, True)
    A.tune(1)
    del rp, ci, v
    inf = A.info()
    A.spmv_timed(y, x, 3)
    tot, kern = A.spmv_timed(y, x, 30)
    print("chunk %3d%% of mean: nvrows %d padding %.3f tiles %d smem %d B: "
          "%.1f us" % (pct, inf["nvrows"],
                       inf["padded_entries"] / inf["nnz_low"],
                       inf["transposed_tiles"], inf["tile_smem_bytes"],
                       kern / 30 * 1e3), flush=True)
    A.close()
    torch.cuda.empty_cache()
