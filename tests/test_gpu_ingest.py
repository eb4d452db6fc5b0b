"""Matrix Market ingest on the GPU (cfs_cuda_matrix_create_from_mmf, the file
constructor of CSRMatrix on a GPU box) against (1) dumps of the reference's own
loader (tests/golden/mtx), (2) the reference's error messages and (3) the host
loader of libsparse.so on generated files, bit for bit."""
import json
import os
import random
import subprocess
import sys

import numpy as np
import pytest

import cases
from cfs_spmv_b200 import capi, gen
from test_mmf_loader import BAD, GOOD, LIBSPARSE, MTX_DIR, load

pytestmark = pytest.mark.gpu


def load_with(path, gpu_ingest, want_symmetric=True):
    old = os.environ.get("CFS_GPU_INGEST")
    os.environ["CFS_GPU_INGEST"] = "1" if gpu_ingest else "0"
    try:
        return load(path, want_symmetric)
    finally:
        if old is None:
            del os.environ["CFS_GPU_INGEST"]
        else:
            os.environ["CFS_GPU_INGEST"] = old


def same_csr(a, b):
    for k in ("nrows", "ncols", "nnz", "symmetric"):
        assert a[k] == b[k], k
    assert np.array_equal(a["rowptr"], b["rowptr"])
    assert np.array_equal(a["colind"], b["colind"])
    assert a["values"].tobytes() == b["values"].tobytes()


@pytest.mark.parametrize("name", GOOD)
def test_gpu_ingest_matches_reference_loader(gpu, name):
    gold = np.load(os.path.join(MTX_DIR, name + "-P1.npz"))
    path = os.path.join(MTX_DIR, name + ".mtx")
    # ... through CSRMatrix(filename): what bench_spmv_mmf / test_spmv_mmf do
    got = load_with(path, True)
    assert got["symmetric"] == int(gold["symmetric"])
    assert np.array_equal(got["rowptr"], gold["csr_rowptr"])
    assert np.array_equal(got["colind"], gold["csr_colind"])
    assert got["values"].tobytes() == gold["csr_values"].tobytes()
    # ... and through the C ABI directly: proves the GPU path ran
    try:
        A, h, rep = capi.Matrix.from_mmf(path, True, True)
    except capi.CfsError as e:
        # only where one (row, col) occurs twice with different values: the
        # reference's order is std::sort's, the file goes to the host loader
        assert name == "duplicates" and e.code == capi.CFS_ERR_NEEDS_HOST
        return
    assert rep["nnz"] == int(gold["nnz_full"])
    rp, ci, v = A.download_csr(int(h.nrows), rep["nnz"])
    assert np.array_equal(rp, gold["csr_rowptr"])
    assert np.array_equal(ci, gold["csr_colind"])
    assert v.tobytes() == gold["csr_values"].tobytes()
    A.close()
    # single precision: atof, then the cast of MMF<int,float>
    A, h, rep = capi.Matrix.from_mmf(path, False, True)
    rp, ci, v = A.download_csr(int(h.nrows), rep["nnz"])
    assert v.tobytes() == gold["csr_values"].astype(np.float32).tobytes()
    A.close()


@pytest.mark.parametrize("fname", BAD)
def test_gpu_ingest_errors_like_the_reference(gpu, fname):
    expect = json.load(open(os.path.join(MTX_DIR, "errors.json")))[fname]
    code = ("import ctypes,sys; L=ctypes.CDLL(%r); b=ctypes.create_string_buffer(64);"
            "L.cfs_host_load_mmf(%r, 1, b)" % (
                LIBSPARSE, os.path.join(MTX_DIR, fname).encode()))
    env = dict(os.environ, CFS_GPU_INGEST="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True,
                       text=True, env=env)
    assert r.returncode == expect["exit"] == 1
    assert r.stdout == expect["stdout"]


def test_short_line_and_bad_index_go_to_the_host_loader(gpu, tmp_path):
    p = str(tmp_path / "short.mtx")
    open(p, "w").write("%%MatrixMarket matrix coordinate real symmetric\n"
                       "3 3 3\n1 1 2.0\n2\n3 3 1.0\n")
    with pytest.raises(capi.CfsError) as e:
        capi.Matrix.from_mmf(p)
    assert e.value.code == capi.CFS_ERR_NEEDS_HOST
    p = str(tmp_path / "range.mtx")
    open(p, "w").write("%%MatrixMarket matrix coordinate real symmetric\n"
                       "3 3 2\n1 1 2.0\n4 1 1.0\n")
    with pytest.raises(capi.CfsError) as e:
        capi.Matrix.from_mmf(p)
    assert e.value.code == capi.CFS_ERR_NEEDS_HOST


def tricky_file(path, n, nlines, seed):
    """every irregularity the tokeniser and the number parser know about"""
    rng = random.Random(seed)
    vals = ["1", "-2.5", "+3e2", "1e-3", ".5", "5.", "0x1p3", "1e400", "-1e-400",
            "4.9406564584124654e-324", "1.7976931348623157e308", "nan", "inf",
            "123456789012345678901234567890e-20", "0.1", "1.5abc", "7\r",
            "9007199254740993", "2.2250738585072011e-308", "1e22", "1e23",
            "8.5e-5", "3.0000000000000004"]
    lines = []
    # distinct positions of the lower triangle (duplicates are another test)
    cells = rng.sample(range(n * (n + 1) // 2), nlines)
    for k in range(nlines):
        r = int((np.sqrt(8.0 * cells[k] + 1) - 1) / 2)
        while r * (r + 1) // 2 > cells[k]:
            r -= 1
        while (r + 1) * (r + 2) // 2 <= cells[k]:
            r += 1
        c = cells[k] - r * (r + 1) // 2
        r, c = r + 1, c + 1
        kind = rng.randrange(10)
        v = rng.choice(vals) if kind < 3 else "%.17g" % rng.uniform(-9, 9)
        if kind == 3:
            lines.append("%d %d" % (r, c))                 # two tokens -> 0.42
        elif kind == 4:
            lines.append("  %d   %d  %s  extra tokens" % (r, c, v))
        elif kind == 5:
            lines.append("\t%d %d %s\t" % (r, c, v))
        elif kind == 6:
            lines.append("%d \t%d %s" % (r, c, v))           # tab inside a token
        elif kind == 7:
            lines.append("%05d +%d %s" % (r, c, v))
        else:
            lines.append("%d %d %s" % (r, c, v))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real symmetric\n% comment\n")
        f.write("%d %d %d\n" % (n, n, len(lines)))
        f.write("\n".join(lines) + "\n")
        f.write("9 9 9 trailing lines beyond the declared count are ignored\n")


def test_gpu_ingest_equals_host_loader_on_tricky_input(gpu, tmp_path):
    p = str(tmp_path / "tricky.mtx")
    tricky_file(p, 500, 20000, 5)
    host = load_with(p, False)
    dev = load_with(p, True)
    same_csr(dev, host)
    A, h, rep = capi.Matrix.from_mmf(p)
    assert 0 < rep["host_lines"] < 0.2 * 20000  # hex / inf / nan / long digits
    A.close()
    same_csr(load_with(p, True, False), load_with(p, False, False))


def test_duplicates_keep_the_reference_order(gpu, tmp_path):
    """equal values: any order is the same CSR, the GPU path keeps the file;
    different values: the host loader (same std::sort as the reference)"""
    head = "%%MatrixMarket matrix coordinate real general\n3 3 4\n"
    p = str(tmp_path / "dup_same.mtx")
    open(p, "w").write(head + "1 1 2.0\n2 2 3.0\n1 1 2.0\n3 3 1.0\n")
    A, h, rep = capi.Matrix.from_mmf(p)
    assert rep["nnz"] == 4
    A.close()
    p = str(tmp_path / "dup_diff.mtx")
    open(p, "w").write(head + "1 1 2.0\n2 2 3.0\n1 1 2.5\n3 3 1.0\n")
    with pytest.raises(capi.CfsError) as e:
        capi.Matrix.from_mmf(p)
    assert e.value.code == capi.CFS_ERR_NEEDS_HOST
    same_csr(load_with(p, True), load_with(p, False))


def test_gpu_ingest_general_and_base0(gpu, tmp_path):
    p = str(tmp_path / "general.mtx")
    rng = np.random.default_rng(3)
    n, m = 300, 5000
    cells = rng.choice(n * n, m, replace=False)
    r, c = cells // n, cells % n
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general base-0\n")
        f.write("%d %d %d\n" % (n, n, m))
        for i in range(m):
            f.write("%d %d %r\n" % (r[i], c[i], float(rng.standard_normal())))
    same_csr(load_with(p, True), load_with(p, False))


def test_gpu_ingest_large_file_and_spmv(gpu, tmp_path):
    """7-point Laplacian 40^3 with random SPD-ish values through the file:
    GPU ingest == host loader, then tune + SpMV on the ingested matrix"""
    rp, ci, v = cases.matrix("lap7_9x7x5")
    spec = capi.GenSpec.laplacian(7, 40, 40, 40)
    rp, ci, v = capi.gen_host_csr(spec)
    rng = np.random.default_rng(11)
    # symmetric random values: value depends on the unordered pair
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    lo, hi = np.minimum(rows, ci), np.maximum(rows, ci)
    v = np.sin(lo * 12.9898 + hi * 78.233) * 43758.5453
    v = v - np.floor(v) + (rows == ci) * 8.0
    p = str(tmp_path / "lap.mtx")
    gen.write_mtx(p, rp, ci, v)
    host = load_with(p, False)
    A, h, rep = capi.Matrix.from_mmf(p)
    n, nnz = int(h.nrows), rep["nnz"]
    drp, dci, dv = A.download_csr(n, nnz)
    assert np.array_equal(drp, host["rowptr"])
    assert np.array_equal(dci, host["colind"])
    assert dv.tobytes() == host["values"].tobytes()
    assert rep["host_lines"] == 0
    A.tune(1)
    x = gen.gen_x(1, n, np.float64)
    y = np.zeros(n)
    A.spmv(y, x)
    import scipy.sparse as sp
    ref = sp.csr_matrix((host["values"], host["colind"], host["rowptr"]),
                        shape=(n, n)) @ x
    assert np.linalg.norm(y - ref) <= 1e-12 * np.linalg.norm(ref)
    A.close()


def test_gpu_ingest_degenerate_files(gpu, tmp_path):
    """no entries at all, diagonal only, CR LF line ends, 1 x 1"""
    banner = "%%MatrixMarket matrix coordinate real symmetric\n"
    files = {
        "empty.mtx": banner + "4 4 0\n",
        "diag.mtx": banner + "3 3 3\n1 1 1.5\n2 2 2.5\n3 3 3.5\n",
        "crlf.mtx": banner + "3 3 3\r\n1 1 1.5\r\n2 1 -2.5\r\n3 3 3.5\r\n",
        "one.mtx": banner + "1 1 1\n1 1 7\n",
    }
    for name, text in files.items():
        p = str(tmp_path / name)
        open(p, "w", newline="").write(text)
        same_csr(load_with(p, True), load_with(p, False))
        A, h, rep = capi.Matrix.from_mmf(p)   # the GPU path took it
        A.close()
    A, h, rep = capi.Matrix.from_mmf(str(tmp_path / "empty.mtx"))
    assert rep["nnz"] == 0
    rp, ci, v = A.download_csr(4, 0)
    assert not rp.any()
    A.close()
