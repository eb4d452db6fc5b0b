#!/bin/bash
# Matrix Market ingest. The same consumer source
# (tests/cpp/load_timer.cpp: SparseMatrix::create + CSR checksum) linked against
# the unmodified reference and against this library, on the BASELINE configs[0]
# file (7-point Laplacian 100^3, 3.97 M lines) and a 27-point 60^3 file.
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import sys, time
sys.path.insert(0, ".")
from cfs_spmv_b200 import capi, gen
for name, spec in (("lap7_100", capi.GenSpec.laplacian(7, 100, 100, 100)),
                   ("lap27_60", capi.GenSpec.laplacian(27, 60, 60, 60))):
    rp, ci, v = capi.gen_host_csr(spec)
    t0 = time.time()
    gen.write_mtx("/tmp/%s.mtx" % name, rp, ci, v)
    print("wrote /tmp/%s.mtx in %.1f s" % (name, time.time() - t0), flush=True)
PY
for M in /tmp/lap7_100.mtx /tmp/lap27_60.mtx; do
  ls -l $M
  cat $M > /dev/null   # page cache warm for everybody
  echo "== reference loader (unmodified, host)"
  oracle/_ref/load_timer $M 1
  echo "== this library, host loader (CFS_GPU_INGEST=0)"
  CFS_GPU_INGEST=0 build/dropin/load_timer $M 1 warm
  echo "== this library, GPU ingest, CUDA context created before the clock"
  for k in 1 2 3 4; do CFS_GPU_INGEST_REPORT=1 build/dropin/load_timer $M 1 warm; done
  echo "== this library, GPU ingest, cold process (context creation inside)"
  build/dropin/load_timer $M 1
done
