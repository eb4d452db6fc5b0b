#!/bin/bash
# Where the preproc(sec) of the drop-in bench goes, run to run (development aid)
cd "$(dirname "$0")/.."
MTX=/tmp/lap7_100.mtx
[ -f $MTX ] || python - <<'PY'
import sys
sys.path.insert(0, ".")
from cfs_spmv_b200 import capi, gen
rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(7, 100, 100, 100))
gen.write_mtx("/tmp/lap7_100.mtx", rp, ci, v)
PY
lscpu | grep -iE "numa|socket|model name|^cpu\(s\)" 
nvidia-smi topo -m 2>/dev/null | head -12
for P in 1 16 16 16; do
  echo "== CFS_NUM_THREADS=$P"
  CFS_GPU_TUNE_REPORT=1 CFS_GPU_INGEST_REPORT=1 CFS_NUM_THREADS=$P build/dropin/bench_spmv_mmf_dp $MTX 1 128 2>&1
done
for i in 1 2 3 4 5; do
  echo "== pinned, CFS_NUM_THREADS=16, run $i"
  CFS_GPU_ALLOC=pinned CFS_NUM_THREADS=16 build/dropin/bench_spmv_mmf_dp $MTX 1 128 2>&1 | tail -1
done
for i in 1 2 3; do
  echo "== pinned, CFS_NUM_THREADS=16, pinned to cpu 0 (taskset), run $i"
  CFS_GPU_ALLOC=pinned CFS_NUM_THREADS=16 taskset -c 0 build/dropin/bench_spmv_mmf_dp $MTX 1 128 2>&1 | tail -1
done
