#!/bin/bash
# the drop-in bench on BASELINE configs[0] under pipeline settings (development aid)
cd "$(dirname "$0")/.."
MTX=/tmp/lap7_100.mtx
python - <<'PY'
import sys
sys.path.insert(0, ".")
from cfs_spmv_b200 import capi, gen
rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(7, 100, 100, 100))
gen.write_mtx("/tmp/lap7_100.mtx", rp, ci, v)
PY
for O in "" "pipeline_adaptive=0" "pipeline_adaptive=0,pipeline_chunks=3" "pipeline_adaptive=0,pipeline_chunks=4" "pipeline=0"; do
  echo "== CFS_GPU_OPTIONS=$O"; CFS_GPU_OPTIONS=$O CFS_NUM_THREADS=1 build/dropin/bench_spmv_mmf_dp $MTX 1 256
done
