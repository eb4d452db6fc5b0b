"""Host logic of the multi-GPU path, exercised on CPU with the gloo backend
(world_size 2 and 3): row-block partitioning, the halo exchange plan and the
x-gather / y-scatter-add exchange. The per-rank 'kernel' here is a numpy
restatement of the symmetric update on the rank's row block (the CUDA kernel
itself is covered by tests/test_gpu_parity.py::test_row_shards_reproduce_the_whole)."""
import os
import socket

import numpy as np
import pytest

import cases
from cfs_spmv_b200 import capi, dist as cdist, gen


def test_row_blocks_are_aligned_and_cover():
    for n, w in ((8000000, 8), (1000, 3), (17, 2), (64000000, 8), (48, 5)):
        b = cdist.row_blocks(n, w)
        assert b[0] == 0 and b[-1] == n and len(b) == w + 1
        assert all(x <= y for x, y in zip(b, b[1:]))
        assert all(x % 16 == 0 for x in b[:-1])


def test_nnz_balanced_blocks():
    rp, ci, v = cases.matrix("rmat_9")
    low = np.array([(ci[rp[i]:rp[i + 1]] < i).sum() for i in range(len(rp) - 1)])
    prefix = np.concatenate([[0], np.cumsum(low)])
    b = cdist.nnz_balanced_blocks(prefix, 4)
    assert b[0] == 0 and b[-1] == len(rp) - 1
    assert all(x % 16 == 0 for x in b[:-1])
    per = [prefix[b[g + 1]] - prefix[b[g]] for g in range(4)]
    assert max(per) <= 1.5 * (prefix[-1] / 4) + low.max() * 16


def test_plan_exchange_is_consistent():
    # what g receives from r is exactly what r sends to g
    ranges = [(0, 0, 160), (90, 160, 320), (100, 320, 480), (470, 480, 500)]
    plans = [cdist.plan_exchange(ranges, g) for g in range(4)]
    for g in range(4):
        for peer, lo, hi in plans[g][0]:
            assert (g, lo, hi) in plans[peer][1]
            assert ranges[peer][1] <= lo < hi <= ranges[peer][2]
        covered = sorted((lo, hi) for _, lo, hi in plans[g][0])
        h, b, _ = ranges[g]
        pos = h
        for lo, hi in covered:
            assert lo == pos
            pos = hi
        assert pos == b


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rp, ci, v = cases.matrix(name)
        n = len(rp) - 1
        x = gen.gen_x(3, n)
        bounds = cdist.row_blocks(n, world)
        b, e = bounds[rank], bounds[rank + 1]
        cols = ci[rp[b]:rp[e]]
        h = int(min(cols.min(), b)) if e > b else b
        mine = torch.tensor([h, b, e], dtype=torch.int64)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        ranges = [tuple(int(t) for t in r.tolist()) for r in allr]
        x_ext = torch.zeros(e - h, dtype=torch.float64)
        x_ext[b - h:] = torch.from_numpy(x[b:e])       # only the owned part
        y_ext = torch.zeros(e - h, dtype=torch.float64)
        halo = cdist.HaloExchanger(ranges, rank, x_ext)
        for _ in range(2):                             # twice: buffers reusable
            halo.exchange_x(x_ext)
            assert np.array_equal(x_ext.numpy(), x[h:e]), "x halo wrong"
            y = np.zeros(e - h)
            for i in range(b, e):                      # symmetric update
                for j in range(rp[i], rp[i + 1]):
                    c = ci[j]
                    if c < i:
                        y[i - h] += v[j] * x[c]
                        y[c - h] += v[j] * x[i]
                    elif c == i:
                        y[i - h] += v[j] * x[i]
            y_ext.copy_(torch.from_numpy(y))
            halo.reduce_y(y_ext)
        import scipy.sparse as sp
        A = sp.csr_matrix((v, ci, rp), shape=(n, n))
        y_ref = A @ x
        err = np.linalg.norm(y_ext.numpy()[b - h:] - y_ref[b:e]) / np.linalg.norm(y_ref)
        out.put((rank, float(err)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "lap27_10"), (3, "banded_3000"),
                                        (2, "rmat_9"), (3, "lap7_9x7x5")])
def test_halo_exchange_gloo(world, name):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, out))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    errs = dict(out.get(timeout=5) for _ in range(world))
    assert len(errs) == world and max(errs.values()) <= 1e-13
