#!/bin/bash
# The unmodified bench_spmv_mmf on BASELINE configs[0] (7-pt Laplacian 100^3,
# --enable-dp) against this library under the vector-allocation kinds, partition
# counts and OpenMP wait policies (development aid: the P=16 question of round 1).
cd "$(dirname "$0")/.."
MTX=/tmp/lap7_100.mtx
python - <<'PY'
import sys
sys.path.insert(0, ".")
from cfs_spmv_b200 import capi, gen
rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(7, 100, 100, 100))
gen.write_mtx("/tmp/lap7_100.mtx", rp, ci, v)
PY
NP=$(nproc)
LOOPS=${LOOPS:-128}
echo "== host: $NP hardware threads; loops=$LOOPS"
echo "== reference (OpenMP CFS, $NP host threads)"
CFS_NUM_THREADS=$NP OMP_PROC_BIND=close oracle/_ref/bench_spmv_mmf_dp $MTX 1 $LOOPS
for ALLOC in managed pinned; do
  for P in 1 $NP; do
    echo "== this library, CFS_GPU_ALLOC=$ALLOC CFS_NUM_THREADS=$P"
    CFS_GPU_ALLOC=$ALLOC CFS_NUM_THREADS=$P build/dropin/bench_spmv_mmf_dp $MTX 1 $LOOPS
    echo "== this library, CFS_GPU_ALLOC=$ALLOC CFS_NUM_THREADS=$P OMP_WAIT_POLICY=passive"
    OMP_WAIT_POLICY=passive CFS_GPU_ALLOC=$ALLOC CFS_NUM_THREADS=$P build/dropin/bench_spmv_mmf_dp $MTX 1 $LOOPS
  done
done
echo "== CFS_GPU_ALLOC=pinned CFS_NUM_THREADS=$NP, long loop (1024)"
CFS_GPU_ALLOC=pinned CFS_NUM_THREADS=$NP build/dropin/bench_spmv_mmf_dp $MTX 1 1024
echo "== CFS_GPU_ALLOC=managed CFS_NUM_THREADS=$NP, long loop (1024)"
CFS_GPU_ALLOC=managed CFS_NUM_THREADS=$NP build/dropin/bench_spmv_mmf_dp $MTX 1 1024
echo "== managed, prefetch on every call"
CFS_GPU_OPTIONS=managed_prefetch=2 CFS_NUM_THREADS=1 build/dropin/bench_spmv_mmf_dp $MTX 1 $LOOPS
echo "== managed, no advice"
CFS_GPU_OPTIONS=managed_advise=0 CFS_NUM_THREADS=1 build/dropin/bench_spmv_mmf_dp $MTX 1 $LOOPS
echo "== single precision, managed"
CFS_NUM_THREADS=1 build/dropin/bench_spmv_mmf_sp $MTX 1 $LOOPS
echo "== test_spmv_mmf (unmodified) against this library, fmt 0 1 2"
for F in 0 1 2; do CFS_NUM_THREADS=$NP build/dropin/test_spmv_mmf $MTX $F; done
for F in 0 1 2; do CFS_GPU_ALLOC=pinned CFS_NUM_THREADS=$NP build/dropin/test_spmv_mmf $MTX $F; done
