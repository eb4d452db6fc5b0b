import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cfs_spmv_b200 import gen
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
rp, ci, v = gen.rmat_torch(scale, 8, 1, is_double=False)
n = 1 << scale
rows = torch.repeat_interleave(torch.arange(n, device="cuda"), (rp[1:] - rp[:-1]).long())
low = ci.long() < rows
cols = ci.long()[low]
cnt = torch.bincount(cols, minlength=n)
tot = cols.numel()
print("scale", scale, "nnz_low", tot, "max col count", cnt.max().item())
for H in (128, 256, 512, 1024, 2048, 4096, 16384):
    hub = cnt >= H
    print("H=%6d hubs=%8d entries=%5.1f%%" % (H, hub.sum().item(), 100.0 * cnt[hub].sum().item() / tot))
