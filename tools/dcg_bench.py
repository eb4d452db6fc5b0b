"""Conjugate gradients over row shards (DistributedCG), one process per GPU:
the weak-scaling family of bench.py (8 M rows per GPU, 27-point Laplacian;
N=8 is the 400^3 matrix of BASELINE configs[4]). Launch with torchrun.
usage: torchrun --nproc-per-node N tools/dcg_bench.py [iters=100]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from cfs_spmv_b200 import capi  # noqa: E402
from cfs_spmv_b200.dist import DistributedCG, ShardedSpMV  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    capi.init(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx, ny, nz = bench.grid_for(world)
    spec = capi.GenSpec.laplacian(27, nx, ny, nz)
    op = ShardedSpMV(spec, rank, world, is_double=True, xseed=1)
    xs = op.x_ext[op.b - op.h:].clone()
    for _ in range(3):
        op.step()
    b = op.y_owned().clone()
    cg = DistributedCG(op)
    cg.solve(b, 10, 0.0, check_every=1000)            # warm-up
    res = cg.solve(b, iters, 0.0, check_every=1000)   # fixed iteration count
    t = torch.tensor([res["ms_total"]], dtype=torch.float64, device="cuda")
    num = ((cg.x - xs) ** 2).sum().reshape(1)
    den = (xs ** 2).sum().reshape(1)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(num)
        dist.all_reduce(den)
    if rank == 0:
        us = t.item() * 1e3 / iters
        print(json.dumps({
            "workload": "CG, 27-pt Laplacian %dx%dx%d double, %d GPU(s)" % (
                nx, ny, nz, world),
            "rows": nx * ny * nz, "iterations": iters,
            "us_per_iteration": round(us, 1),
            "gflops_spmv_part": round(2.0 * bench.lap27_nnz_full(nx, ny, nz) /
                                      us / 1e3, 1),
            "residual_drop": res["residual_norm"] / res["initial_residual_norm"],
            "error_vs_known_solution": float((num / den).sqrt().item())}),
            flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
