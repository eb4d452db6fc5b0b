#!/bin/bash
# N-GPU benches (development aid): tools/scale_check.sh <N>
cd "$(dirname "$0")/.."
N=$1
O=gpurun_out
if [ "$N" -ge 2 ]; then
  python -m pytest tests/test_gpu_multi.py -x -q > $O/tests_multi_$N.log 2>&1
fi
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$RUN --master-port 29521 bench.py --gpus $N --steps 300 --warmup 10 > $O/scale_$N.json 2> $O/scale_$N.err
$RUN --master-port 29522 bench.py --gpus $N --workload banded --steps 300 --warmup 10 > $O/banded_$N.json 2> $O/banded_$N.err
