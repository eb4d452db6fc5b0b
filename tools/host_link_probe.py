"""What the host side of the box gives N GPUs at once (development aid): every
active rank copies a pinned 64 MB vector in and another one out at the same
time, like the end-to-end leg of bench.py, with 1, 2, 4, ... ranks active.
torchrun --nproc-per-node 8 tools/host_link_probe.py"""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 8000000
    hx = torch.ones(n, dtype=torch.float64).pin_memory()
    hy = torch.empty(n, dtype=torch.float64).pin_memory()
    dx = torch.empty(n, dtype=torch.float64, device="cuda")
    dy = torch.ones(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    active = 1
    while active <= world:
        for mode in ("h2d", "d2h", "both"):
            dist.barrier()
            torch.cuda.synchronize()
            iters = 10
            t0 = time.perf_counter()
            if rank < active:
                for _ in range(iters):
                    if mode in ("h2d", "both"):
                        with torch.cuda.stream(s1):
                            dx.copy_(hx, non_blocking=True)
                    if mode in ("d2h", "both"):
                        with torch.cuda.stream(s2):
                            hy.copy_(dy, non_blocking=True)
                torch.cuda.synchronize()
            dt = torch.tensor([(time.perf_counter() - t0) / iters],
                              dtype=torch.float64, device="cuda")
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if rank == 0:
                per_dir = n * 8 / dt.item() / 1e9
                dirs = 2 if mode == "both" else 1
                print("%d rank(s) active, %-4s: %.3f ms per 64 MB vector -> %.1f "
                      "GB/s per direction per GPU, %.1f GB/s aggregate" % (
                          active, mode, dt.item() * 1e3, per_dir,
                          per_dir * dirs * active), flush=True)
        active *= 2
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
