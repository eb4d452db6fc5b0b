#!/bin/bash
# The reference's unmodified bench on 1..N GPUs of one process (CFS_NUM_GPUS),
# 27-pt 200^3 per-edge values through the array constructor is not available to
# the stock binary, so the file route: 27-pt 100^3 and 7-pt 160^3 .mtx
cd "$(dirname "$0")/.."
python - <<'PY'
import sys, os
sys.path.insert(0, ".")
from cfs_spmv_b200 import capi, gen
for name, spec in (("lap27_100", capi.GenSpec.laplacian(27, 100, 100, 100, 7)),):
    path = "/tmp/%s.mtx" % name
    if not os.path.exists(path):
        rp, ci, v = capi.gen_host_csr(spec)
        gen.write_mtx(path, rp, ci, v)
PY
NG=${1:-4}
for G in 1 2 4 8; do
  [ $G -le $NG ] || continue
  for Z in 1 0; do
    echo "== CFS_NUM_GPUS=$G multi_zero_copy=$Z"
    CFS_GPU_OPTIONS=multi_zero_copy=$Z CFS_NUM_GPUS=$G CFS_NUM_THREADS=1 build/dropin/bench_spmv_mmf_dp /tmp/lap27_100.mtx 1 256
  done
done
