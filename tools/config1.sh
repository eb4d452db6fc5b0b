#!/bin/bash
# BASELINE.json configs[0]: bench_spmv_mmf on a 7-point Laplacian 100^3 .mtx,
# --enable-dp: the reference's OpenMP CFS on the host cores next to the SAME
# unmodified bench source linked against this library (B200).
set -e
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
echo "== reference (OpenMP CFS, $NP host threads)"
CFS_NUM_THREADS=$NP OMP_PROC_BIND=close oracle/_ref/bench_spmv_mmf_dp $MTX 1 128
echo "== this library (B200), same bench source, CFS_NUM_THREADS=$NP (metadata partitions)"
CFS_NUM_THREADS=$NP build/dropin/bench_spmv_mmf_dp $MTX 1 128
echo "== this library (B200), CFS_NUM_THREADS=1"
CFS_NUM_THREADS=1 build/dropin/bench_spmv_mmf_dp $MTX 1 128
echo "== test_spmv_mmf (unmodified) against this library"
CFS_NUM_THREADS=$NP build/dropin/test_spmv_mmf $MTX 1
