"""Multi-GPU runtime: the reference's 1-D contiguous row partitioning
(partition_by_nrows / partition_by_nnz, include/matrix/csr_matrix.tpp:404-541)
lifted from OpenMP threads to GPUs, one process per GPU.

GPU g owns global rows [R_g, R_{g+1}). It stores the lower-triangle entries of
its rows; their columns reach down to halo_begin_g <= R_g, so

  * it needs x[halo_begin_g : R_g) from the owners below it   (x halo), and
  * it produces transposed contributions to y[halo_begin_g : R_g), the
    reference's "direct conflicts" (csr_matrix.tpp:1447), which the owners must
    add to their y                                               (y halo).

Both are point-to-point exchanges with the few ranks directly below (one for
stencil / banded matrices), done over NCCL (NVLink) with torch.distributed as
plumbing. No dense reduce-scatter of y: only the touched strip moves.
"""
import numpy as np


def row_blocks(nrows, world, align=16):
    """contiguous, `align`-row aligned (the reference's BlkFactor) equal-row
    blocks: bounds[g] .. bounds[g+1]"""
    bounds = [min(nrows, (nrows * g // world) // align * align)
              for g in range(world)] + [nrows]
    return bounds


def nnz_balanced_blocks(row_nnz_prefix, world, align=16):
    """contiguous blocks with ~equal nnz (partition_by_nnz semantics,
    csr_matrix.tpp:438-541): row_nnz_prefix[i] = nnz of rows < i"""
    nrows = len(row_nnz_prefix) - 1
    total = int(row_nnz_prefix[-1])
    bounds = [0]
    for g in range(1, world):
        target = total * g // world
        r = int(np.searchsorted(row_nnz_prefix, target, side="left"))
        r = min(nrows, (r + align // 2) // align * align)
        bounds.append(max(r, bounds[-1]))
    bounds.append(nrows)
    return bounds


def plan_exchange(ranges, rank):
    """ranges[g] = (halo_begin, row_begin, row_end) of every rank.
    Returns (recv_x, send_x): lists of (peer, global_lo, global_hi).
      recv_x: x rows this rank must receive from `peer` (its halo, cut by owner)
      send_x: x rows this rank owns that `peer` needs
    The y halo travels the same segments in the opposite direction."""
    h, b, e = ranges[rank]
    recv_x, send_x = [], []
    for peer, (ph, pb, pe) in enumerate(ranges):
        if peer == rank:
            continue
        lo, hi = max(h, pb), min(b, pe)  # my halo inside peer's rows
        if lo < hi:
            recv_x.append((peer, lo, hi))
        lo, hi = max(ph, b), min(pb, e)  # peer's halo inside my rows
        if lo < hi:
            send_x.append((peer, lo, hi))
    return recv_x, send_x


class HaloExchanger:
    """x-halo gather and y-halo scatter-add between row-block owners.
    Works on any torch.distributed backend (NCCL on GPUs, gloo in the CPU
    tests); vectors are the extended local vectors covering [halo_begin,
    row_end)."""

    def __init__(self, ranges, rank, like):
        import torch
        self.rank = rank
        self.h, self.b, self.e = ranges[rank]
        self.recv_x, self.send_x = plan_exchange(ranges, rank)
        self.world = len(ranges)
        # staging for incoming y contributions
        self.y_in = [torch.empty(hi - lo, dtype=like.dtype, device=like.device)
                     for (_, lo, hi) in self.send_x]
        self.bytes_per_step = sum((hi - lo) for _, lo, hi in
                                  self.recv_x + self.send_x) * like.element_size()

    def _sl(self, lo, hi):
        return slice(lo - self.h, hi - self.h)

    def exchange_x(self, x_ext):
        import torch.distributed as dist
        if self.world == 1 or not (self.recv_x or self.send_x):
            return
        ops = [dist.P2POp(dist.isend, x_ext[self._sl(lo, hi)], peer)
               for peer, lo, hi in self.send_x]
        ops += [dist.P2POp(dist.irecv, x_ext[self._sl(lo, hi)], peer)
                for peer, lo, hi in self.recv_x]
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def reduce_y(self, y_ext):
        import torch.distributed as dist
        if self.world == 1 or not (self.recv_x or self.send_x):
            return
        ops = [dist.P2POp(dist.isend, y_ext[self._sl(lo, hi)], peer)
               for peer, lo, hi in self.recv_x]
        ops += [dist.P2POp(dist.irecv, buf, peer)
                for buf, (peer, lo, hi) in zip(self.y_in, self.send_x)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for buf, (peer, lo, hi) in zip(self.y_in, self.send_x):
            y_ext[self._sl(lo, hi)] += buf


class P2PHalo:
    """Halo handling over peer memory (NVLink) instead of NCCL messages.

    x_ext and TWO y_ext vectors of every rank live in symmetric memory
    (torch.distributed._symmetric_memory: plumbing that maps each rank's
    buffers into the others and provides a device-side barrier). One step =

      ONE kernel (cfs_cuda_spmv_shard_async): reads its x halo straight from
      the x of the GPU below, reduces its halo contributions straight into the
      y of the GPU below (RED.sys over NVLink), and its row owners clear the
      OTHER y vector, which the next step reduces into;
      ONE barrier: all reductions have landed, every "other" vector is clear.

    Round 1 needed a y memset, a barrier, an x-halo copy, the kernel and a
    second barrier per step (45-60 us around a 150 us kernel). A caller that
    rewrites x between two steps synchronises once more before the kernel
    (`x_changed=True`): the GPU above reads this GPU's x.
    Needs every halo to lie inside the row block of ONE lower neighbour (true
    for stencil and banded matrices)."""

    def __init__(self, ranges, rank, dtype, device):
        import os
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.rank, self.world = rank, len(ranges)
        self.h, self.b, self.e = ranges[rank]
        for g in range(1, self.world):
            if ranges[g][0] < ranges[g - 1][1]:
                raise RuntimeError("halo of rank %d spans more than one "
                                   "neighbour" % g)
        max_len = max(e - h for h, b, e in ranges)
        self.x_sym = symm_mem.empty(max_len, dtype=dtype, device=device)
        self.y_sym = symm_mem.empty(2 * max_len, dtype=dtype, device=device)
        self.hx = symm_mem.rendezvous(self.x_sym, dist.group.WORLD)
        self.hy = symm_mem.rendezvous(self.y_sym, dist.group.WORLD)
        n = self.e - self.h
        item = self.x_sym.element_size()
        self.x_ext = self.x_sym[:n]
        self.y_bufs = [self.y_sym[:n], self.y_sym[max_len:max_len + n]]
        self.y_lower_bases = [None, None]
        self.x_lower_base = None
        self.x_peer = None
        self.nhalo = self.b - self.h
        if rank > 0 and self.nhalo > 0:
            ph = ranges[rank - 1][0]
            y_peer = self.hy.buffer_ptrs[rank - 1]
            self.y_lower_bases = [y_peer - ph * item,
                                  y_peer + (max_len - ph) * item]
            self.x_lower_base = self.hx.buffer_ptrs[rank - 1] - ph * item
            peer_x = self.hx.get_buffer(rank - 1, (max_len,), dtype)
            self.x_peer = peer_x[self.h - ph:self.b - ph]
        self.bytes_per_step = 2 * self.nhalo * item
        # the rows the GPU above reduces into: [its halo_begin, my row_end)
        self.band = None
        if rank + 1 < self.world and ranges[rank + 1][0] < ranges[rank + 1][1]:
            lo = max(ranges[rank + 1][0], self.b) - self.h
            self.band = (lo, self.e - self.h)
        self.sync = os.environ.get("CFS_GPU_HALO_SYNC", "barrier")
        self.overlap = os.environ.get("CFS_GPU_HALO_OVERLAP", "1") != "0"
        self.side = torch.cuda.Stream(device=device)
        self.ev_part1 = None
        self.ev_main = None
        self.dirty = False
        self.t = 0
        self.reset()

    # the older, bulk-synchronous pieces (DistributedCG uses them: its SpMV
    # also returns p'Ap): result in y_bufs[0], x halo pulled by a peer copy
    @property
    def y_ext(self):
        return self.y_bufs[0]

    @property
    def y_lower_base(self):
        return self.y_lower_bases[0]

    def reset(self):
        """both result vectors clear on every rank"""
        self.join()
        self.y_sym.zero_()
        self.t = 0
        self.hy.barrier()

    @property
    def y_current(self):
        """the vector the last step() left y in"""
        return self.y_bufs[(self.t - 1) % 2] if self.t else self.y_bufs[0]

    # Synchronisation of the bulk-synchronous pieces. "barrier": device
    # barriers over all ranks. "neighbour" (CFS_GPU_HALO_SYNC=neighbour): a
    # rank signals / waits for its two neighbours through the signal pads of
    # the symmetric memory instead.
    _TIMEOUT_MS = 20000

    def sync_before(self):
        if self.sync != "neighbour":
            self.hy.barrier()
            return
        if self.rank + 1 < self.world:
            self.hy.put_signal(self.rank + 1, 0, self._TIMEOUT_MS)
        if self.rank > 0:
            self.hy.wait_signal(self.rank - 1, 0, self._TIMEOUT_MS)

    def sync_after(self):
        if self.sync != "neighbour":
            self.hy.barrier()
            return
        if self.rank > 0:
            self.hy.put_signal(self.rank - 1, 1, self._TIMEOUT_MS)
        if self.rank + 1 < self.world:
            self.hy.wait_signal(self.rank + 1, 1, self._TIMEOUT_MS)

    def step(self, matrix, stream, x_changed=False):
        if x_changed or not self.overlap:
            if x_changed:
                self.join()
                self.hy.barrier()      # x below is final
            k = self.t % 2
            matrix.spmv_shard_async(self.y_bufs[k], self.x_ext,
                                    self.y_lower_bases[k], self.x_lower_base,
                                    self.y_bufs[1 - k], True, stream)
            self.hy.barrier()          # reductions landed, other vector clear
            self.t += 1
            return self.y_bufs[k]
        return self._step_overlapped(matrix)

    # The barrier off the critical path (CFS_GPU_HALO_OVERLAP=0 switches it
    # off). Only the leading slices of a shard touch the GPU below (part 1:
    # 0.1-2 % of them); all the others (part 2) need nothing from anybody else.
    #   main stream:  ... part2(t) ................. part2(t+1) ...
    #   side stream:  barrier(t-1) -> part1(t) -> [part2(t) done] barrier(t) -> part1(t+1)
    # part 1 of step t runs behind barrier t-1 (every GPU has finished step
    # t-1: the vector it reduces into is clear, x below is final) and next to
    # part 2 of the same step; barrier t waits for both parts; part 2 of step
    # t+1 only waits for part 1 of step t (rows that part cleared / reduced).
    # One more thing follows from not waiting: part 2 of step t clears the OTHER
    # vector while the GPU above may still be reducing its step t-1 halo into
    # the top rows of that vector (the band), so the band is cleared once more
    # behind barrier t-1, when those reductions have certainly landed.
    def _step_overlapped(self, matrix):
        import torch
        main = torch.cuda.current_stream()
        side = self.side
        k = self.t % 2
        if self.ev_part1 is not None:
            main.wait_event(self.ev_part1)        # part1(t-1) has finished
        matrix.spmv_shard_part_async(self.y_bufs[k], self.x_ext,
                                     self.y_lower_bases[k], self.x_lower_base,
                                     self.y_bufs[1 - k], 2, main.cuda_stream)
        ev_main = torch.cuda.Event()
        ev_main.record(main)
        if self.ev_main is not None:
            side.wait_event(self.ev_main)         # part2(t-1): before barrier(t-1)
        else:
            side.wait_stream(main)                # first step: set-up is done
        with torch.cuda.stream(side):
            if self.band is not None:             # behind barrier(t-1)
                self.y_bufs[1 - k][self.band[0]:self.band[1]].zero_()
            matrix.spmv_shard_part_async(self.y_bufs[k], self.x_ext,
                                         self.y_lower_bases[k],
                                         self.x_lower_base, self.y_bufs[1 - k],
                                         1, side.cuda_stream)
            self.ev_part1 = torch.cuda.Event()
            self.ev_part1.record(side)
            side.wait_event(ev_main)              # part2(t) has finished
            self.hy.barrier()                     # barrier(t)
        self.ev_main = ev_main
        self.dirty = True
        self.t += 1
        return self.y_bufs[k]

    def join(self):
        """the main stream waits for everything the side stream has been given
        (a consumer of y, a step that is not overlapped)"""
        import torch
        if getattr(self, "dirty", False):
            torch.cuda.current_stream().wait_stream(self.side)
            self.dirty = False
            self.ev_part1 = None
            self.ev_main = None


class ShardedSpMV:
    """One rank's share of y = A*x for a generated matrix (bench / tests).
    spec: a capi.GenSpec; or arrays=(n, rowptr, colind, values) device tensors
    of the whole matrix (world 1: R-MAT, which needs a global sort to build)."""

    def __init__(self, spec, rank, world, is_double=True, xseed=1, arrays=None):
        import os
        import torch
        import torch.distributed as dist
        from . import capi
        self.rank, self.world = rank, world
        if arrays is not None:
            assert world == 1, "array input: one GPU"
            n, rp, ci, v = arrays
            b, e = 0, n
        else:
            n = spec.nrows
            b, e, rp, ci, v = self._balanced_shard(spec, rank, world, is_double)
        if world == 1:
            self.matrix = capi.Matrix(n, n, rp, ci, v, is_double, True)
        else:
            self.matrix = capi.Matrix(0, 0, rp, ci, v, is_double, True,
                                      shard=(n, b, e))
        self.matrix.tune(1)
        del rp, ci, v
        torch.cuda.empty_cache()
        self.info = self.matrix.info()
        h = self.info["halo_begin"]
        self.h, self.b, self.e = h, b, e
        if world > 1:
            mine = torch.tensor([h, b, e], dtype=torch.int64, device="cuda")
            allr = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            ranges = [tuple(int(v) for v in t.tolist()) for t in allr]
        else:
            ranges = [(h, b, e)]
        self.ranges = ranges
        x_gen = capi.gen_device_x(xseed, h, e, is_double)
        self.p2p = None
        mode = os.environ.get("CFS_GPU_HALO", "p2p" if world > 1 else "none")
        if world > 1 and mode == "p2p":
            try:
                self.p2p = P2PHalo(ranges, rank, x_gen.dtype, x_gen.device)
            except Exception as exc:  # no peer access / halo too wide
                if rank == 0:
                    print("P2P halo unavailable (%s): using NCCL messages" % exc,
                          flush=True)
                self.p2p = None
            ok = torch.tensor([1 if self.p2p is not None else 0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                self.p2p = None
        # single GPU: the same ping-pong of two result vectors drops the y
        # initialisation from a loop of SpMVs (CFS_GPU_PINGPONG=0: memset + kernel)
        self.pingpong = (world == 1 and self.info["symmetric"] == 1 and
                         os.environ.get("CFS_GPU_PINGPONG", "1") != "0")
        self.t = 0
        if self.p2p is not None:
            self.x_ext = self.p2p.x_ext
            self.x_ext.copy_(x_gen)
            self.y_bufs = self.p2p.y_bufs
            dist.barrier()
        else:
            self.x_ext = x_gen
            # the halo part of x is (re)filled by exchange_x every step
            self.y_bufs = [torch.zeros_like(self.x_ext),
                           torch.zeros_like(self.x_ext)]
        self.y_ext = self.y_bufs[0]
        self.halo = HaloExchanger(ranges, rank, self.x_ext)
        nh = len(self.halo.recv_x) + len(self.halo.send_x)
        hubs = 1 if self.info["hub_columns"] else 0
        if world == 1:
            self.exchange_desc = "none (single GPU)" + (
                "; y initialisation fused into the previous SpMV (two result "
                "vectors used alternately)" if self.pingpong else "")
            self.kernels_per_step = 1 + hubs
        elif self.p2p is not None:
            self.exchange_desc = (
                "fused over NVLink peer memory: ONE kernel per step reads its x "
                "halo from the GPU below, reduces its y halo contributions "
                "straight into that GPU's y (RED.sys) and clears the other of "
                "two result vectors; ONE device barrier per step%s; %d bytes on "
                "rank %d" % (", off the critical path: the slices that touch "
                             "the GPU below run on a side stream behind the "
                             "previous step's barrier, next to all the others"
                             if self.p2p.overlap else "",
                             self.p2p.bytes_per_step, rank))
            # SpMV (interior + halo part when overlapped) + the symmetric-memory
            # barrier (+ the band clear)
            self.kernels_per_step = 2 + (
                (1 if self.p2p.nhalo > 0 else 0) +
                (1 if self.p2p.band is not None else 0)
                if self.p2p.overlap else 0)
        else:
            self.exchange_desc = (
                "NCCL P2P per step: x halo down-up, y halo strip add; "
                "%d neighbour segments, %d bytes on rank %d" % (
                    nh, self.halo.bytes_per_step, rank))
            self.kernels_per_step = 1 + hubs + len(self.halo.send_x)
        self._host = None

    @staticmethod
    def _balanced_shard(spec, rank, world, is_double):
        """this rank's rows: contiguous, 16-row aligned blocks holding about the
        same number of stored entries (partition_by_nnz semantics,
        csr_matrix.tpp:438-541, lifted to GPUs). Every rank counts the entries
        of an equal-rows block, the counts per 1024-row chunk are exchanged and
        the boundaries moved to the chunk where the prefix sum crosses
        g/world; a rank whose block moved generates its rows again."""
        import torch
        import torch.distributed as dist
        from . import capi
        n = spec.nrows
        eq = row_blocks(n, world)
        b, e = eq[rank], eq[rank + 1]
        rp, ci, v = capi.gen_device_csr(spec, b, e, is_double)
        if world == 1:
            return b, e, rp, ci, v
        chunk = 1024
        nchunks = (n + chunk - 1) // chunk
        mine = torch.zeros(nchunks, dtype=torch.int64, device="cuda")
        # equal-rows blocks are 16-row aligned, chunks start at multiples of
        # 1024: a chunk can straddle two blocks, both add their part
        rows = torch.arange(b, e, device="cuda")
        cnt = (rp[1:] - rp[:-1]).to(torch.int64)
        mine.index_add_(0, rows // chunk, cnt)
        dist.all_reduce(mine)
        prefix = torch.zeros(nchunks + 1, dtype=torch.int64)
        prefix[1:] = torch.cumsum(mine.cpu(), 0)
        # rows of chunk boundaries -> nnz_balanced_blocks on the coarse prefix
        cb = nnz_balanced_blocks(prefix.numpy(), world, align=1)
        bounds = [min(n, c * chunk) for c in cb]
        bounds[-1] = n
        if any(bounds[g + 1] <= bounds[g] for g in range(world)):
            bounds = eq  # too few rows for chunk-wise balancing
        nb, ne = bounds[rank], bounds[rank + 1]
        if (nb, ne) != (b, e):
            del rp, ci, v, rows, cnt
            torch.cuda.empty_cache()
            rp, ci, v = capi.gen_device_csr(spec, nb, ne, is_double)
        return nb, ne, rp, ci, v

    def kernel_desc(self):
        inf = self.info
        t = "double" if inf["is_double"] else "float"
        if inf.get("transposed_tiles"):
            return ("sym_spmv_tile_kernel<%s> (variant 6: transposed term "
                    "transposed through shared memory, one coalesced RED per "
                    "column and tile)" % t)
        if inf["regular_slices"] * 8 >= inf["nslices"]:
            return ("sym_spmv_reg_kernel<%s> (variant 5: compressed index "
                    "stream, shuffle-merged REDs, %s)" % (
                        t, "values dictionary-coded: %d distinct value(s), "
                           "lossless" % inf["value_dictionary"]
                        if inf.get("value_dictionary") else
                        "8-byte values streamed behind L2 prefetches"))
        return ("sym_spmv_sell_kernel<%s> (variant 1: one warp per slice of 32 "
                "length-sorted virtual rows%s)" % (
                    t, " + hub_spmv_kernel: %d hub columns run column-wise" %
                    inf["hub_columns"] if inf["hub_columns"] else ""))

    def step(self, x_changed=False):
        import torch
        s = torch.cuda.current_stream().cuda_stream
        if self.p2p is not None:
            self.y_ext = self.p2p.step(self.matrix, s, x_changed)
            return
        if self.pingpong:
            k = self.t % 2
            self.matrix.spmv_shard_async(self.y_bufs[k], self.x_ext, None, None,
                                         self.y_bufs[1 - k], True, s)
            self.y_ext = self.y_bufs[k]
            self.t += 1
            return
        self.halo.exchange_x(self.x_ext)
        self.matrix.spmv_async(self.y_ext, self.x_ext, s)
        self.halo.reduce_y(self.y_ext)

    def time_kernel(self, iters):
        """average device time (ms) of the SpMV kernel alone: `iters` launches
        back to back between two CUDA events, the two result vectors used
        alternately (each launch clears the vector the next one reduces into),
        no halo exchange and no barrier -- a shard reduces its halo
        contributions into the halo part of its own vector"""
        import torch
        self.sync()
        s = torch.cuda.current_stream()
        bufs = self.y_bufs
        if self.p2p is not None:  # private vectors: nothing of a neighbour's
            bufs = [torch.zeros_like(self.x_ext), torch.zeros_like(self.x_ext)]
        if self.info["symmetric"] != 1:
            raise RuntimeError("time_kernel: symmetric matrices only")
        for b in bufs:
            b.zero_()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        for k in range(2):  # warm
            self.matrix.spmv_shard_async(bufs[k], self.x_ext, None, None,
                                         bufs[1 - k], True, s.cuda_stream)
        e0.record(s)
        for k in range(iters):
            self.matrix.spmv_shard_async(bufs[k % 2], self.x_ext, None, None,
                                         bufs[1 - k % 2], True, s.cuda_stream)
        e1.record(s)
        e1.synchronize()
        if self.p2p is None:  # leave the step's invariant behind: both clear
            for b in self.y_bufs:
                b.zero_()
            self.t = 0
        return e0.elapsed_time(e1) / iters

    def sync(self):
        """the current stream waits for whatever the overlapped step left on
        its side stream (the last barrier and halo part)"""
        if self.p2p is not None:
            self.p2p.join()

    def y_owned(self):
        self.sync()
        return self.y_ext[self.b - self.h:]

    def checksum(self):
        import torch
        import torch.distributed as dist
        t = self.y_owned().sum().reshape(1)
        if self.world > 1:
            dist.all_reduce(t)
        return float(t.item())

    def e2e(self, steps):
        """the same SpMV with HOST vectors: pinned H2D of x, kernel (+ y halo
        exchange), D2H of y, every step. Returns (total_ms, h2d, d2h bytes)."""
        import time
        import torch
        n_own = self.e - self.b
        if self.world == 1:
            y_dev = self.y_owned().clone()
            x_host = self.x_ext.cpu().pin_memory()
            y_host = torch.empty_like(x_host).pin_memory()
            self.matrix.spmv(y_host, x_host)  # allocates the staging buffers
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                self.matrix.spmv(y_host, x_host)  # synchronous C ABI call
            ms = (time.perf_counter() - t0) * 1e3
            err = (y_host - y_dev.cpu()).abs().max().item()
            scale = y_dev.abs().max().item()
            tol = 1e-9 if y_dev.dtype == torch.float64 else 1e-3
            assert err <= tol * max(scale, 1.0), "e2e result differs"
            return ms, x_host.numel() * x_host.element_size(), \
                y_host.numel() * y_host.element_size()
        x_host = self.x_ext.cpu().pin_memory()       # owned rows + halo
        y_host = torch.empty(n_own, dtype=self.x_ext.dtype).pin_memory()
        s = torch.cuda.current_stream()

        def one():
            self.x_ext.copy_(x_host, non_blocking=True)
            if self.p2p is not None:
                # every rank's x has just been rewritten: synchronise first
                self.y_ext = self.p2p.step(self.matrix, s.cuda_stream, True)
            else:
                self.halo.exchange_x(self.x_ext)
                self.matrix.spmv_async(self.y_ext, self.x_ext, s.cuda_stream)
                self.halo.reduce_y(self.y_ext)
            y_host.copy_(self.y_owned(), non_blocking=True)
            s.synchronize()

        one()
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        ms = (time.perf_counter() - t0) * 1e3
        return ms, x_host.numel() * x_host.element_size(), \
            y_host.numel() * y_host.element_size()


class DistributedCG:
    """Conjugate gradients over the row shards of a ShardedSpMV (fused NVLink
    halo only): A x = b with x, r, p, q sharded like the rows.

    Per iteration: q = A p with the halo reduction fused into the kernel AND
    this shard's part of p'Ap coming out of it (cfs_cuda_spmv_halo_dot_async);
    a 1-double all-reduce; x += a p, r -= a q, r'r (cfs_cuda_cg_update_xr); a
    1-double all-reduce; p = r + b p (cfs_cuda_cg_update_p). The all-reduces are
    the exchange step conjugate gradients really have; torch.distributed is
    the plumbing for them. p lives in the symmetric x buffer of the halo
    exchange, so the neighbours pull its halo like any x."""

    def __init__(self, op):
        import torch
        if op.world > 1 and op.p2p is None:
            raise RuntimeError("DistributedCG needs the fused NVLink halo")
        self.op = op
        self.n = op.e - op.b
        self.off = op.b - op.h          # halo rows in front of the owned ones
        dt = op.x_ext.dtype
        dev = op.x_ext.device
        self.is_double = dt == torch.float64
        self.x = torch.zeros(self.n, dtype=dt, device=dev)
        self.r = torch.zeros(self.n, dtype=dt, device=dev)
        # {r'r, p'Ap, next r'r} as global sums, + the local parts before the
        # all-reduce
        self.scal = torch.zeros(3, dtype=torch.float64, device=dev)

    def _allreduce(self, t):
        import torch.distributed as dist
        if self.op.world > 1:
            dist.all_reduce(t)

    def _spmv_dot(self, stream):
        """q = A p (result in op.y_ext), scal[1] = p'Ap (global)"""
        op = self.op
        self.scal[1:2].zero_()
        if op.p2p is not None:
            p2p = op.p2p
            p2p.y_ext.zero_()
            p2p.sync_before()
            if p2p.x_peer is not None:
                p2p.x_ext[:p2p.nhalo].copy_(p2p.x_peer)
            op.matrix.spmv_halo_dot_async(p2p.y_ext, p2p.x_ext,
                                          p2p.y_lower_base, True,
                                          self.scal[1:2], stream)
            p2p.sync_after()
        else:
            op.y_ext.zero_()
            op.matrix.spmv_halo_dot_async(op.y_ext, op.x_ext, None, True,
                                          self.scal[1:2], stream)
        self._allreduce(self.scal[1:2])

    def solve(self, b, max_iters, rel_tol, x0=None, check_every=10):
        """b, x0: this rank's OWNED rows (device tensors). Returns a dict;
        the solution stays in self.x."""
        import torch
        from . import capi
        op = self.op
        stream = torch.cuda.current_stream().cuda_stream
        p_own = op.x_ext[self.off:]
        self.x.zero_()
        if x0 is not None:
            self.x.copy_(x0)
        # r = b - A x0
        p_own.copy_(self.x)
        self._spmv_dot(stream)
        self.r.copy_(b - op.y_ext[self.off:])
        p_own.copy_(self.r)
        self.scal[0:1].copy_((self.r.double() * self.r.double()).sum().reshape(1))
        self._allreduce(self.scal[0:1])
        rr0 = float(self.scal[0].item())
        tol2 = rel_tol * rel_tol * rr0
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        it, rr, converged, breakdown = 0, rr0, rr0 == 0.0, False
        while it < max_iters and not converged and not breakdown:
            self._spmv_dot(stream)
            self.scal[2:3].zero_()
            capi.cg_update_xr(self.n, self.is_double, self.scal, p_own,
                              op.y_ext[self.off:], self.x, self.r,
                              self.scal[2:3], stream)
            self._allreduce(self.scal[2:3])
            capi.cg_update_p(self.n, self.is_double, self.scal, self.r, p_own,
                             stream)
            self.scal[0:1].copy_(self.scal[2:3])
            it += 1
            if it % check_every == 0 or it == max_iters:
                vals = self.scal.tolist()
                rr = vals[0]
                breakdown = not (vals[1] > 0.0)
                converged = rr <= tol2
        e1.record()
        torch.cuda.synchronize()
        return {"iterations": it, "converged": bool(converged),
                "breakdown": bool(breakdown),
                "initial_residual_norm": rr0 ** 0.5,
                "residual_norm": max(rr, 0.0) ** 0.5,
                "ms_total": e0.elapsed_time(e1)}
