#!/bin/bash
# ncu evidence of round 2 (development aid): launch lists and full captures of the
# dominant kernels, each after the same command has exited 0 without ncu.
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
$B > $O/plain_lap27.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/launches_r2_lap27.csv $B > $O/ncu_l_lap27.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sym_spmv_reg -s 4 -c 2 \
    -o $O/prof_r2_lap27 $B > $O/ncu_p_lap27.log 2>&1
B="python bench.py --workload rmat --steps 5 --warmup 3 --no-cpu-baseline"
$B > $O/plain_rmat.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $O/launches_r2_rmat.csv $B > $O/ncu_l_rmat.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'sym_spmv_sell|hub_spmv' -s 4 -c 4 \
    -o $O/prof_r2_rmat $B > $O/ncu_p_rmat.log 2>&1
