#!/bin/bash
# ncu captures of the banded (variant 6) and the value-indexed kernel, round 2
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --workload banded --steps 5 --warmup 3 --no-cpu-baseline"
$B > $O/plain_banded.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/launches_r2_banded.csv $B > $O/ncu_l_banded.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sym_spmv_tile -s 4 -c 2 \
    -o $O/prof_r2_banded $B > $O/ncu_p_banded.log 2>&1
B="python bench.py --values constant --steps 5 --warmup 3 --no-cpu-baseline"
$B > $O/plain_const.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/launches_r2_const.csv $B > $O/ncu_l_const.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sym_spmv_reg -s 4 -c 2 \
    -o $O/prof_r2_const $B > $O/ncu_p_const.log 2>&1
