#!/usr/bin/env python
"""bench.py -- symmetric SpMV throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--workload lap27|banded|rmat] [--values distinct|constant]
                  [--scaling weak|strong]

A "step" is ONE SpMV y = A*x over the whole (sharded) matrix with x, y and the
matrix resident in HBM.

  lap27 (default)  N=1 is BASELINE.json configs[1]: the 27-point stencil on a
                   200^3 grid, double precision. N>1 is the weak-scaling family
                   that ends in configs[4] (400^3 on 8 GPUs): 8 M rows per GPU.
                   --values distinct (default): every edge has its own
                   coefficient, so the 8-byte value stream is really streamed;
                   --values constant: the constant-coefficient Laplacian (ONE
                   distinct off-diagonal value, which this library
                   dictionary-codes away: reported as roofline.compressed).
  banded           configs[3]: banded SPD matrix, half bandwidth 2000, ~9.5
                   lower entries per row at random offsets, double.
                   --scaling strong: 32 M rows in all (configs[3] as stated);
                   --scaling weak (default): 8 M rows per GPU.
  rmat             configs[2]: symmetric R-MAT, scale 24 (16 M rows, ~256 M
                   nnz), single precision; N=1 only.

Rows are split in contiguous blocks (the reference's row partitioning lifted to
GPUs); x halos and the transposed y contributions that cross a block boundary
are exchanged every step.

`--impl reference` times the reference's own OpenMP CFS path (the unmodified
reference compiled into oracle/_ref) on the host cores, rank 0 only.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "symmetric SpMV GFLOP/s (2*nnz_full per SpMV)"
UNIT = "GFLOP/s"
XSEED = 1
VALUE_SEED = 7                  # --values distinct: per-edge coefficients
# weak-scaling family: 8 M rows per GPU; N=1 is configs[1], N=8 is configs[4]
GRIDS = {1: (200, 200, 200), 2: (200, 200, 400), 4: (200, 400, 400),
         8: (400, 400, 400)}
BANDED_ROWS_PER_GPU = 8000000   # weak family; configs[3] = 32 M rows
BANDED_ROWS_STRONG = 32000000
BANDED = (2000, 152, 7)         # half bandwidth, lower entries per row x16, seed
RMAT = (24, 8, 1)               # scale, edge factor, seed


def grid_for(n_gpus):
    if n_gpus in GRIDS:
        return GRIDS[n_gpus]
    return (200, 200, 200 * n_gpus)


def lap27_nnz_full(nx, ny, nz):
    return (3 * nx - 2) * (3 * ny - 2) * (3 * nz - 2)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(key, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant
    kernel from an ncu --set full capture of THIS workload at N=1
    (profiles/traffic.json); None where it was not profiled"""
    if world != 1:
        return None, "not profiled at N > 1"
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            rec = json.load(f).get(key)
        if rec:
            return rec["dram_bytes_per_launch"], rec.get("source")
    except Exception:
        pass
    return None, "not profiled"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the GPU is under load"""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,"
              "clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    REASONS = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
               "sw_power_cap")

    def __init__(self, device_index):
        self.device_index = device_index
        self.samples = []  # (t, sm, sm_max, power, [reasons])
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.device_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm, smax = float(parts[1]), float(parts[2])
                power = float(parts[3]) if parts[3][0].isdigit() else 0.0
            except Exception:
                continue
            reasons = [r for r, p in zip(self.REASONS, parts[4:8])
                       if p.lower().startswith("active")]
            self.samples.append((time.time(), sm, smax, power, reasons))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1, fallback_window=None):
        def pick(a, b):
            return [s for s in self.samples if a <= s[0] <= b]
        inside = pick(t0, t1)
        where = "during the timed region"
        if not inside and fallback_window is not None:
            inside = pick(*fallback_window)
            where = ("timed region shorter than the sampling period; sampled "
                     "during an identical untimed loop right after it")
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [],
                    "samples": 0, "sampled": "nvidia-smi gave no samples"}
        sm = sorted(s[1] for s in inside)
        reasons = sorted(set(r for s in inside for r in s[4]))
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": inside[0][2],
                "power_w_max": max(s[3] for s in inside),
                "reasons": reasons, "samples": len(inside), "sampled": where}


# ---------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------
def workload_spec(args, world):
    """-> (GenSpec or None for rmat, is_double, traffic key)"""
    from cfs_spmv_b200 import capi
    if args.workload == "banded":
        n = (BANDED_ROWS_STRONG if args.scaling == "strong"
             else BANDED_ROWS_PER_GPU * world)
        return capi.GenSpec.banded(n, *BANDED), True, "banded_f64"
    if args.workload == "rmat":
        return None, False, "rmat_f32"
    nx, ny, nz = grid_for(world)
    seed = VALUE_SEED if args.values == "distinct" else 0
    return (capi.GenSpec.laplacian(27, nx, ny, nz, seed), True,
            "lap27_%s_f64" % args.values)


def workload_config(args, n_gpus, nnz_full=None, rows=None):
    if args.workload == "banded":
        n = (BANDED_ROWS_STRONG if args.scaling == "strong"
             else BANDED_ROWS_PER_GPU * n_gpus)
        return {
            "workload": "banded SPD matrix, %d rows, half bandwidth %d, ~9.5 "
                        "lower entries per row at random offsets, double "
                        "(BASELINE.json configs[3]%s)" % (
                            n, BANDED[0],
                            "" if n == BANDED_ROWS_STRONG else
                            " weak-scaling family, 8 M rows per GPU"),
            "rows": n, "nnz_full": nnz_full, "rows_per_gpu": n // n_gpus,
            "partition": "contiguous row blocks, one per GPU",
            "cache": "inputs larger than L2 (>= 1.1 GB streamed per GPU per "
                     "step vs 126 MB L2); no flush needed",
        }
    if args.workload == "rmat":
        return {
            "workload": "symmetric R-MAT scale %d, edge factor %d, (a,b,c)="
                        "(0.57,0.19,0.19), deduplicated + full diagonal, single "
                        "precision (BASELINE.json configs[2])" % RMAT[:2],
            "rows": rows or (1 << RMAT[0]), "nnz_full": nnz_full,
            "rows_per_gpu": rows or (1 << RMAT[0]),
            "partition": "one GPU",
            "cache": "inputs larger than L2 (1.1 GB streamed per step vs "
                     "126 MB L2); no flush needed",
        }
    nx, ny, nz = grid_for(n_gpus)
    return {
        "workload": "27-point stencil %dx%dx%d, double, lower triangle "
                    "stored (BASELINE.json configs[%d]%s)" % (
                        nx, ny, nz, 1 if n_gpus == 1 else 4,
                        "" if n_gpus in (1, 8) else " weak-scaling family"),
        "values": "one coefficient per edge in [-1,-0.5), diagonal 1/16 + sum "
                  "|a|: ~nnz/2 distinct values, nothing to dictionary-code"
                  if args.values == "distinct" else
                  "constant-coefficient Laplacian (26 / -1): ONE distinct "
                  "off-diagonal value, dictionary-coded (lossless)",
        "rows": nx * ny * nz, "nnz_full": lap27_nnz_full(nx, ny, nz),
        "rows_per_gpu": nx * ny * nz // n_gpus,
        "partition": "contiguous row blocks, one per GPU",
        "cache": "inputs larger than L2 (1.46 GB streamed per GPU per step vs "
                 "126 MB L2); no flush needed",
    }


# ---------------------------------------------------------------------------
# the reference arm / CPU baseline (the ONLY places that execute oracle/)
# ---------------------------------------------------------------------------
def partition_count(n, cores):
    from oracle import oracle
    P = min(cores, 96)  # MaxThreads of the reference (runtime.hpp:15)
    while P > 1 and not oracle.valid_partition_count(n, P):
        P -= 1
    return P


def reference_sample(args):
    """the bounded sample of the workload the CPU runs: (ref_tool input spec,
    rows, precision letter, description, max partitions)"""
    if args.workload == "banded":
        n = BANDED_ROWS_PER_GPU
        return ("gen:banded:%d:%d:%d:%d" % ((n,) + BANDED), n, "d",
                "banded SPD matrix, %d rows, half bandwidth %d, double" % (
                    n, BANDED[0]), None)
    if args.workload == "rmat":
        return None, 0, "s", "", 1
    seed = VALUE_SEED if args.values == "distinct" else 0
    g = (200, 200, 200)
    return ("gen:lap27:%d:%d:%d%s" % (g + (":%d" % seed if seed else "",)),
            g[0] * g[1] * g[2], "d",
            "27-pt stencil 200x200x200 double, %s values" % args.values, None)


def rmat_sample_file(scale):
    """R-MAT for the reference: built with numpy (cfs_spmv_b200/gen.py, the same
    definition the GPU generator follows) and handed over as a CSR file"""
    from cfs_spmv_b200 import gen
    from oracle import oracle
    import numpy as np
    rp, ci, v = gen.rmat(scale, RMAT[1], RMAT[2], dtype=np.float64)
    path = os.path.join(tempfile.gettempdir(), "cfs_rmat_%d.bin" % scale)
    oracle.write_csr_bin(path, rp, ci, v)
    return "csr:" + path, len(rp) - 1


def run_cpu(args, loops, warmup, y_out=None, spec_override=None):
    """times the compiled reference (oracle/_ref/ref_tool bench) on the host
    cores; falls back to the C restatement on one thread when it is missing.
    -> dict(value, ms, cores, kind, sample, ...)"""
    from oracle import oracle
    cores = os.cpu_count() or 1
    spec, n, prec, desc, pmax = reference_sample(args)
    if spec_override is not None:
        spec, n, desc = spec_override
    if args.workload == "rmat" and spec is None:
        scale = 22
        spec, n = rmat_sample_file(scale)
        desc = ("symmetric R-MAT scale %d (%d rows: a bounded sample of the "
                "scale-%d workload), single" % (scale, n, RMAT[0]))
    P = partition_count(n, cores if pmax is None else pmax)
    if not oracle.ref_available():
        out = time_oracle_port(loops)
        out["sample"] = ("oracle C restatement (oracle/_ref missing), 27-pt "
                         "100^3 double, %d SpMVs, 1 thread" % loops)
        return out
    r = oracle.run_ref_bench(spec, P, prec, XSEED, loops, timeout=3000,
                             warmup=warmup, y_out=y_out)
    why_p = ("" if pmax is None else
             " (P = 1: the reference's conflict-graph preprocessing is "
             "O(sum of squared column counts) on a power-law matrix, "
             "SURVEY.md 7.3)")
    return {"value": r["gflops"], "ms": r["t_spmv_s"] * 1e3, "cores": P,
            "kind": "reference", "preproc_s": r["preproc_s"],
            "ncolors": r["ncolors"], "nnz_full": r["nnz_full"],
            "sample": "%s; %d timed SpMVs after %d warm-up; unmodified "
                      "reference, OpenMP CFS, CFS_NUM_THREADS=%d of %d host "
                      "cores%s; preproc %.1f s, %d colours" % (
                          desc, loops, r["warmup"], P, cores, why_p,
                          r["preproc_s"], r["ncolors"])}


def time_oracle_port(loops):
    import numpy as np
    from cfs_spmv_b200 import capi, gen
    from oracle import oracle
    spec = capi.GenSpec.laplacian(27, 100, 100, 100)
    rp, ci, v = capi.gen_host_csr(spec)
    o = oracle.Oracle(rp, ci, v, 1)
    x = gen.gen_x(XSEED, len(rp) - 1)
    o.spmv(x)
    t0 = time.perf_counter()
    for _ in range(loops):
        o.spmv(x)
    dt = (time.perf_counter() - t0) / loops
    return {"value": 2.0 * int(rp[-1]) * 1e-9 / dt, "ms": dt * 1e3, "cores": 1,
            "kind": "port"}


def run_reference(args, rank, world):
    """the reference's own CPU implementation of the path, host cores"""
    if rank != 0:
        return
    loops = max(2, min(args.steps, 400))
    out = run_cpu(args, loops, args.warmup)
    share = ""
    if args.gpus > 1 and args.scaling == "weak" and args.workload != "rmat":
        share = (" -- ONE GPU's share of the N=%d workload (the rate is "
                 "size-independent to first order; as the reference's host CSR "
                 "the whole matrix grows to > 60 GB at N = 8)" % args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": out["value"],
        "unit": UNIT, "n_gpus": args.gpus, "steps": loops,
        "warmup": args.warmup, "ms_per_step": out["ms"],
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32" if args.workload == "rmat" else "f64",
        "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": out["value"], "unit": UNIT,
                         "cores": out["cores"], "kind": out["kind"],
                         "sample": out["sample"] + share},
        "e2e": {"value": out["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "host CPU run: a bounded sample of the workload named in "
                "config" + share,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_and_parity(args, y_gpu):
    """N=1: ONE run of the unmodified reference on the full workload gives the
    CPU baseline and its y; parity = normwise error of the GPU's y against it"""
    import numpy as np
    tol = 1e-5 if args.workload == "rmat" else 1e-12
    y_path = os.path.join(tempfile.gettempdir(), "cfs_ref_y_%d.bin" % os.getpid())
    spec_override = None
    try:
        if args.workload == "rmat":
            spec_override = y_gpu.pop("ref_spec")
        out = run_cpu(args, 10, 5, y_out=y_path, spec_override=spec_override)
        base = {"value": out["value"], "unit": UNIT, "cores": out["cores"],
                "kind": out["kind"], "sample": out["sample"]}
        parity = None
        if out["kind"] == "reference" and os.path.exists(y_path):
            y_ref = np.fromfile(y_path, dtype=np.float32 if args.workload == "rmat"
                                else np.float64).astype(np.float64)
            y = y_gpu["y"].astype(np.float64)
            err = float(np.linalg.norm(y - y_ref) / np.linalg.norm(y_ref))
            parity = {"normwise_rel_err": err, "tolerance": tol,
                      "ok": bool(err <= tol),
                      "against": "y of the unmodified reference's CFS kernel "
                                 "(oracle/_ref/ref_tool, P=%d) on the same "
                                 "full-size matrix and x" % out["cores"],
                      "rows": int(len(y_ref))}
        return base, parity
    except Exception as e:  # a baseline failure must not void the GPU number
        return ({"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                 "sample": "failed: %r" % (e,)}, None)
    finally:
        if os.path.exists(y_path):
            os.remove(y_path)


def info_rows_total(op, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([op.e - op.b], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t)
    return int(t.item())


def multi_gpu_parity(rank, world):
    """N>1: the multi-GPU step on a small sharded matrix, assembled on rank 0
    and checked against the CPU oracle (the C restatement, pinned bitwise to
    the reference by tests/test_oracle_golden.py)"""
    import numpy as np
    import torch
    import torch.distributed as dist
    from cfs_spmv_b200 import capi, gen
    from cfs_spmv_b200.dist import ShardedSpMV
    spec = capi.GenSpec.laplacian(27, 40, 40, 16 * world, VALUE_SEED)
    op = ShardedSpMV(spec, rank, world, is_double=True, xseed=5)
    for _ in range(3):
        op.step()
    torch.cuda.synchronize()
    parts = [None] * world
    dist.all_gather_object(parts, (op.b, op.y_owned().cpu().numpy()))
    desc = op.exchange_desc
    del op
    if rank != 0:
        return None
    from oracle import oracle
    rp, ci, v = capi.gen_host_csr(spec)
    ref = oracle.Oracle(rp, ci, v, 1).spmv(gen.gen_x(5, spec.nrows))
    y = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])])
    err = float(np.linalg.norm(y - ref) / np.linalg.norm(ref))
    return {"normwise_rel_err": err, "tolerance": 1e-12, "ok": bool(err <= 1e-12),
            "against": "CPU oracle (C restatement pinned bitwise to the "
                       "reference) on a 27-pt 40x40x%d matrix sharded over the "
                       "same %d GPUs with the same exchange (%s)" % (
                           16 * world, world, desc.split(":")[0]),
            "rows": int(spec.nrows)}


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true",
                    help="skip the CPU baseline + full-size parity leg")
    ap.add_argument("--workload", default="lap27",
                    choices=["lap27", "banded", "rmat"])
    ap.add_argument("--values", default="distinct",
                    choices=["distinct", "constant"],
                    help="lap27 only: per-edge coefficients (default) or the "
                         "constant-coefficient stencil")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="banded only: strong = 32 M rows in all (configs[3])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload != "banded":
        args.scaling = "weak"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from cfs_spmv_b200 import capi, gen
    from cfs_spmv_b200.dist import ShardedSpMV

    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torchrun "
                         "--nproc-per-node %d" % (args.gpus, world, args.gpus))
    if args.workload == "rmat" and world != 1:
        raise SystemExit("--workload rmat is BASELINE configs[2]: one GPU")
    torch.cuda.set_device(local_rank)
    capi.init(local_rank)  # fails loudly without a B200 / the native library
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    spec, is_double, traffic_key = workload_spec(args, world)
    t_setup = time.time()
    ref_spec = None
    if spec is None:  # R-MAT: needs a global sort, built with torch on the GPU
        rp, ci, v = gen.rmat_torch(RMAT[0], RMAT[1], RMAT[2], is_double=False)
        n = rp.numel() - 1
        if not args.no_cpu_baseline:  # the reference gets the same matrix
            from oracle import oracle
            path = os.path.join(tempfile.gettempdir(), "cfs_rmat_full.bin")
            oracle.write_csr_bin(path, rp.cpu().numpy(), ci.cpu().numpy(),
                                 v.cpu().numpy())
            ref_spec = ("csr:" + path, n,
                        "symmetric R-MAT scale %d (the full workload), "
                        "single" % RMAT[0])
        op = ShardedSpMV(None, rank, world, is_double=False, xseed=XSEED,
                         arrays=(n, rp, ci, v))
        del rp, ci, v
    else:
        op = ShardedSpMV(spec, rank, world, is_double=is_double, xseed=XSEED)
    info = op.info
    setup_s = time.time() - t_setup
    cnt = torch.tensor([info["nnz_full"]], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(cnt)
    nnz_full = int(cnt.item())
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        op.step()
    op.sync()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        op.step()
    op.sync()  # the overlapped multi-GPU step leaves its tail on a side stream
    e1.record(stream)
    barrier()
    w1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = ms.item()

    # dominant kernel alone, right behind the timed region (same clocks, same
    # temperature): K back-to-back launches of the SpMV kernel between two CUDA
    # events -- no exchange, no barrier; the two result vectors alternate like
    # in the step, so there is no y initialisation to subtract either
    k_iters = max(10, min(args.steps, 200))
    kernel_ms_avg = op.time_kernel(k_iters)
    kernel_how = ("%d back-to-back launches of the kernel between two CUDA "
                  "events right behind the timed region (two result vectors "
                  "used alternately: no y initialisation, no exchange, no "
                  "barrier)" % k_iters)
    if world == 1 and op.pingpong and op.kernels_per_step == 1:
        # at N=1 a step IS one launch of this kernel: the timed region is the
        # measurement (a second loop minutes of load later runs at other clocks)
        kernel_ms_avg = min(kernel_ms_avg, total_ms / args.steps)
        kernel_how = ("at N=1 a step is ONE launch of this kernel: the smaller "
                      "of the timed region itself (%d launches between two "
                      "CUDA events) and a loop of %d launches right behind it"
                      % (args.steps, k_iters))
    barrier()

    # identical untimed loop (>= 1 s) so that nvidia-smi (50 ms period) is
    # guaranteed samples under this load even for a short timed region
    fb0 = time.time()
    while time.time() - fb0 < 1.0 and (w1 - w0) < 0.5:
        for _ in range(50):
            op.step()
        op.sync()
        torch.cuda.synchronize()
    fb1 = time.time()
    barrier()
    op.step()  # y = A x with the exchange, for the checks below
    barrier()
    checksum = op.checksum()
    y_host = op.y_owned().cpu().numpy() if world == 1 else None

    # end to end through the host-pointer entry point
    e2e_steps = max(5, min(args.steps, 30))
    e2e_ms, h2d, d2h = op.e2e(e2e_steps)
    e2e_t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    barrier()
    if sampler:
        sampler.stop()
    kernels_per_step = op.kernels_per_step
    exchange_desc = op.exchange_desc
    kernel_name = op.kernel_desc()

    # the constant-coefficient matrix beside the headline (N=1 lap27 distinct)
    compressed = None
    if world == 1 and args.workload == "lap27" and args.values == "distinct":
        op = None
        torch.cuda.empty_cache()
        nx, ny, nz = grid_for(1)
        op2 = ShardedSpMV(capi.GenSpec.laplacian(27, nx, ny, nz, 0), 0, 1,
                          is_double=True, xseed=XSEED)
        op2.time_kernel(5)
        kk = op2.time_kernel(k_iters) * k_iters
        ab = op2.info["algorithmic_bytes"]
        tr, tr_src = profiled_traffic("lap27_constant_f64", 1)
        peak, _ = measured_peak()
        compressed = {
            "values": "constant-coefficient Laplacian (26 / -1): ONE distinct "
                      "off-diagonal value, the value stream is dictionary-"
                      "coded away (lossless: the same bits are multiplied)",
            "kernel": op2.kernel_desc(), "kernel_ms": kk / k_iters,
            "achieved": ab / (kk / k_iters * 1e-3) / 1e9,
            "frac": ab / (kk / k_iters * 1e-3) / 1e9 / peak,
            "traffic": tr, "traffic_source": tr_src,
            "note": "frac > 1 here means fewer bytes moved, not a faster "
                    "memory system: compare traffic with "
                    "algorithmic_bytes_per_launch",
        }
        del op2
        torch.cuda.empty_cache()

    # a size-independent property at the FULL size of the N-GPU workload: with
    # per-edge coefficients every row of A sums to exactly 1/16 (diagonal =
    # 1/16 + sum |a|, off-diagonals -a), so A*1 = 1/16 up to the rounding of a
    # 27-term sum -- through the same exchange the timed steps used
    full_size_property = None
    if (world > 1 and args.workload == "lap27" and args.values == "distinct"
            and op is not None):
        x_keep = op.x_ext.clone()
        op.x_ext.fill_(1.0)
        barrier()
        op.step(x_changed=True)
        barrier()
        dev = (op.y_owned() - 0.0625).abs().max().reshape(1)
        dist.all_reduce(dev, op=dist.ReduceOp.MAX)
        op.x_ext.copy_(x_keep)
        barrier()
        full_size_property = {
            "property": "A*1 == 1/16 in every row (per-edge coefficients: "
                        "diagonal = 1/16 + sum |a|), all %d rows, through the "
                        "multi-GPU step" % info_rows_total(op, world),
            "max_abs_deviation": float(dev.item()),
            "tolerance": 1e-13, "ok": bool(dev.item() <= 1e-13)}
    parity = None
    if world > 1:
        parity = multi_gpu_parity(rank, world)
        if parity is not None and full_size_property is not None:
            parity["full_size_property"] = full_size_property

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_bytes = info["algorithmic_bytes"]
        achieved = alg_bytes / (kernel_ms_avg * 1e-3) / 1e9
        traffic, traffic_src = profiled_traffic(traffic_key, world)
        ms_per_step = total_ms / args.steps
        line = {
            "metric": METRIC,
            "value": 2.0 * nnz_full * 1e-9 / (ms_per_step * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None,
            "dtype": "f64" if is_double else "f32", "data": "synthetic",
            # the same keys and values as the reference arm's line
            "config": workload_config(args, world),
            "details": {"nnz_full_counted": nnz_full,
                        "setup_s": round(setup_s, 2),
                        "exchange": exchange_desc,
                        "layout": {k: info[k] for k in
                                   ("nnz_low", "nvrows", "nslices",
                                    "padded_entries", "device_bytes",
                                    "value_dictionary")}},
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "frac_of_traffic": (traffic / (kernel_ms_avg * 1e-3) / 1e9 / peak
                                    if traffic else None),
                "peak_source": peak_src,
                "frac_of_8000": achieved / 8000.0,
                "kernel": kernel_name,
                "kernel_ms": kernel_ms_avg,
                "kernel_ms_how": kernel_how,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "achieved = ALGORITHMIC bytes (SURVEY.md 8d: 12 B per "
                        "stored entry for f64 + vectors) / kernel time; one "
                        "launch = the whole (shard of the) matrix. traffic = "
                        "measured DRAM bytes of the same launch: below the "
                        "algorithmic bytes where index / value streams are "
                        "compressed",
                "hbm_gbs_whole_step": alg_bytes / (ms_per_step * 1e-3) / 1e9,
            },
            "e2e": {
                "value": 2.0 * nnz_full * 1e-9 / (e2e_t.item() / e2e_steps * 1e-3),
                "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": d2h * world, "steps": e2e_steps,
                "ms_per_step": e2e_t.item() / e2e_steps,
                "path": "cfs_cuda_spmv(host y, host x): pinned H2D, kernel and D2H "
                        "overlapped in row chunks (each run as head + rest so "
                        "that y of the previous chunk can leave early), the step "
                        "replayed as one CUDA graph"
                        if world == 1 else
                        "per GPU: pinned H2D of its x rows + halo, fused-halo "
                        "kernel, D2H of its y rows",
            },
            "gpu_launches": args.steps * world * kernels_per_step,
            "clocks": sampler.summary(w0, w1, (fb0, fb1)),
            "checksum": checksum,
        }
        if compressed:
            line["roofline"]["compressed"] = compressed
        if world == 1 and not args.no_cpu_baseline:
            base, parity = cpu_baseline_and_parity(
                args, {"y": y_host, "ref_spec": ref_spec})
            line["cpu_baseline"] = base
        if parity is not None:
            line["parity"] = parity
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
