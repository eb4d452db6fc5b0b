#!/bin/bash
# single-GPU numbers quoted in RESULTS.md (round 2)
cd "$(dirname "$0")/.."
O=gpurun_out
python bench.py --workload banded --steps 300 > $O/bench_banded_n1.json 2> $O/bench_banded_n1.err
python bench.py --steps 2000 > $O/bench_n1_2000.json 2> $O/bench_n1_2000.err
python tools/tune_time.py > $O/tune_time.log 2>&1
python tools/cg_bench.py > $O/cg_bench.log 2>&1
CFS_GPU_TUNE_REPORT=1 python - > $O/tune_report.log 2>&1 <<'PY'
import sys
sys.path.insert(0, ".")
import torch
from cfs_spmv_b200 import capi
capi.init(0)
for seed in (7, 0):
    spec = capi.GenSpec.laplacian(27, 200, 200, 200, seed)
    rp, ci, v = capi.gen_device_csr(spec)
    A = capi.Matrix(spec.nrows, spec.nrows, rp, ci, v, True, True)
    A.tune(1)
    inf = A.info()
    print("seed", seed, "device_bytes", inf["device_bytes"], "size_bytes", inf["size_bytes"], "dict", inf["value_dictionary"], flush=True)
    A.close()
PY
