"""Drop-in check: the reference's UNMODIFIED test_spmv_mmf.cpp and
bench_spmv_mmf.cpp (compiled against this repo's include/ + libsparse.so by
tools/build_dropin.py, where the reference tree was available) and this repo's
own API consumer run on the B200 and report like the reference binaries do."""
import os
import re
import subprocess

import pytest

import cases
from cfs_spmv_b200 import capi, gen

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "build", "dropin")


def _run(binary, args, threads):
    env = dict(os.environ, CFS_NUM_THREADS=str(threads))
    return subprocess.run([os.path.join(DROPIN, binary)] + args, env=env,
                          capture_output=True, text=True, timeout=600)


@pytest.fixture(scope="module")
def mtx(tmp_path_factory):
    d = tmp_path_factory.mktemp("mtx")
    rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(7, 20, 20, 20))
    path = str(d / "lap7_20.mtx")
    gen.write_mtx(path, rp, ci, v)
    rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(27, 16, 16, 16))
    path27 = str(d / "lap27_16.mtx")
    gen.write_mtx(path27, rp, ci, v)
    return {"lap7": path, "lap27": path27,
            "general": os.path.join(cases.GOLDEN_DIR, "mtx", "general.mtx")}


@pytest.mark.parametrize("threads", [1, 4, 8])
@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_own_consumer(gpu, mtx, fmt, threads):
    assert os.path.exists(os.path.join(DROPIN, "api_consumer")), \
        "build/dropin/api_consumer missing: run __graft_entry__.build()"
    for key in ("lap7", "lap27"):
        r = _run("api_consumer", [mtx[key], str(fmt)], threads)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "ALL PASSED!" in r.stdout and "FAILED" not in r.stdout


def test_own_consumer_on_a_general_file(gpu, mtx):
    r = _run("api_consumer", [mtx["general"], "1"], 3)
    assert r.returncode == 0 and "ALL PASSED!" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("threads", [1, 2, 8, 27])
@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_reference_test_program_unmodified(gpu, mtx, fmt, threads):
    if not os.path.exists(os.path.join(DROPIN, "test_spmv_mmf")):
        pytest.skip("reference sources were not available at build time")
    r = _run("test_spmv_mmf", [mtx["lap7"], str(fmt)], threads)
    # the reference's test always exits 0 and reports on stdout
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASSED!" in r.stdout and "FAILED!" not in r.stdout, r.stdout


@pytest.mark.parametrize("binary,dtype", [("bench_spmv_mmf_dp", "double"),
                                          ("bench_spmv_mmf_sp", "float")])
def test_reference_bench_program_unmodified(gpu, mtx, binary, dtype):
    if not os.path.exists(os.path.join(DROPIN, binary)):
        pytest.skip("reference sources were not available at build time")
    r = _run(binary, [mtx["lap27"], "1", "16"], 8)
    assert r.returncode == 0, r.stdout + r.stderr
    # bench_spmv_mmf.cpp:169-173
    m = re.search(r"matrix: (\S+) format: SSS preproc\(sec\): (\S+) "
                  r"t\(sec\): (\S+) gflops/s: (\S+) threads: 8 "
                  r"size\(MB\): (\S+)", r.stdout)
    assert m, r.stdout
    assert m.group(1) == "lap27_16.mtx" and float(m.group(4)) > 0
    # size(MB) follows the reference's formula (csr_matrix.tpp:191-228)
    rp, ci, v = capi.gen_host_csr(capi.GenSpec.laplacian(27, 16, 16, 16))
    from oracle import oracle
    o = oracle.Oracle(rp, ci, v.astype("float64" if dtype == "double"
                                       else "float32"), 8)
    assert abs(float(m.group(5)) - o.size_bytes / 2.0 ** 20) < 1e-2
