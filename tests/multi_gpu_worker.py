"""Worker of tests/test_gpu_multi.py: launched with torchrun, one rank per GPU.
Builds a sharded matrix, runs the multi-GPU SpMV step and checks the assembled
y against the CPU oracle on rank 0."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cfs_spmv_b200 import capi, gen  # noqa: E402
from cfs_spmv_b200.dist import DistributedCG, ShardedSpMV  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    capi.init(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures = 0
    mode = os.environ.get("CFS_GPU_HALO", "p2p")
    for spec, dbl in ((capi.GenSpec.laplacian(27, 40, 40, 16 * world), True),
                      (capi.GenSpec.laplacian(7, 33, 17, 24 * world), False),
                      (capi.GenSpec.banded(20000 * world, 700, 152, 3), True)):
        op = ShardedSpMV(spec, rank, world, is_double=dbl, xseed=5)
        for _ in range(3):  # repeated steps reuse every buffer
            op.step()
        torch.cuda.synchronize()
        y_own = op.y_owned().cpu().numpy()
        parts = [None] * world
        dist.all_gather_object(parts, (op.b, y_own))
        if rank == 0:
            from oracle import oracle
            dt = np.float64 if dbl else np.float32
            rp, ci, v = capi.gen_host_csr(spec, dtype=dt)
            x = gen.gen_x(5, spec.nrows, dt)
            ref = oracle.Oracle(rp, ci, v, 1).spmv(x)
            y = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])])
            err = np.linalg.norm(y.astype(np.float64) - ref) / np.linalg.norm(ref)
            ok = err <= (1e-12 if dbl else 1e-5)
            print("multi-gpu world=%d halo=%s(%s) kind=%d %s err=%.3e %s" % (
                world, mode, "p2p" if op.p2p is not None else "nccl",
                spec.kind, "f64" if dbl else "f32", err,
                "OK" if ok else "FAIL"), flush=True)
            failures += not ok
    # conjugate gradients over the shards (fused halo only): A x = A xs
    if mode == "p2p":
        for spec, dbl, tol, xtol in (
                (capi.GenSpec.laplacian(27, 40, 40, 16 * world), True, 1e-10, 1e-7),
                (capi.GenSpec.banded(20000 * world, 700, 152, 3), False, 1e-5, 1e-3)):
            op = ShardedSpMV(spec, rank, world, is_double=dbl, xseed=9)
            if op.p2p is None:
                continue
            xs = op.x_ext[op.b - op.h:].clone()   # the solution to find
            op.step()
            b = op.y_owned().clone()               # b = A xs
            cg = DistributedCG(op)
            res = cg.solve(b, 3000, tol)
            num = ((cg.x - xs).double() ** 2).sum().reshape(1)
            den = (xs.double() ** 2).sum().reshape(1)
            dist.all_reduce(num)
            dist.all_reduce(den)
            err = float((num / den).sqrt().item())
            ok = res["converged"] and err <= xtol
            if rank == 0:
                print("multi-gpu cg world=%d kind=%d %s: %d iterations, error "
                      "%.3e %s" % (world, spec.kind, "f64" if dbl else "f32",
                                   res["iterations"], err,
                                   "OK" if ok else "FAIL"), flush=True)
            failures += not ok
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
