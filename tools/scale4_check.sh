#!/bin/bash
# 4-GPU measurements of round 2 (development aid)
cd "$(dirname "$0")/.."
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
$RUN --master-port 29521 bench.py --gpus 4 --steps 300 --warmup 10 > $O/scale_4.json 2> $O/scale_4.err
$RUN --master-port 29522 bench.py --gpus 4 --workload banded --scaling strong --steps 300 --warmup 10 > $O/banded_strong_4.json 2> $O/banded_strong_4.err
python - <<'PY'
import sys
sys.path.insert(0, ".")
from cfs_spmv_b200 import capi, gen
rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(27, 100, 100, 100, 7))
gen.write_mtx("/tmp/lap27_100.mtx", rp, ci, v)
PY
for G in 1 2 4; do
  echo "== bench_spmv_mmf (unmodified) 27-pt 100^3 per-edge values, CFS_NUM_GPUS=$G" >> $O/dropin_multi.log
  CFS_NUM_GPUS=$G CFS_NUM_THREADS=1 build/dropin/bench_spmv_mmf_dp /tmp/lap27_100.mtx 1 128 >> $O/dropin_multi.log 2>&1
done
CFS_NUM_GPUS=4 build/dropin/test_spmv_mmf /tmp/lap27_100.mtx 1 >> $O/dropin_multi.log 2>&1
