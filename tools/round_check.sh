#!/bin/bash
# One GPU call that re-validates and re-measures everything (development aid):
# gpu tests, the headline bench + reference arm, the banded (configs[3]) bench,
# ncu launch lists / captures for profiles/, the ingest and CG measurements.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/tests_gpu.log 2>&1
echo "pytest rc=$?" >> $O/tests_gpu.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 20 --warmup 10 > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py --workload banded --steps 500 > $O/bench_banded_n1.json 2> $O/bench_banded_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/launches_r1b.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/launches_banded.csv python bench.py --workload banded --steps 5 --warmup 3 > $O/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sym_spmv_tile -c 2 \
    -o $O/prof_banded python bench.py --workload banded --steps 5 --warmup 3 > $O/ncu_p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sym_spmv_reg -c 2 \
    -o $O/prof_r1b python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/ncu_p1.log 2>&1
python tools/cg_bench.py > $O/cg_bench.log 2>&1
tools/ingest_bench.sh > $O/ingest_bench.log 2>&1
python tools/run_configs.py rmat > $O/rmat.log 2>&1
