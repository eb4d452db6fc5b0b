"""Parity of the CUDA path (through the C ABI) with the reference.

  * preprocessing metadata: BIT-EXACT against the golden dumps of the compiled
    reference (tests/golden) and against the CPU oracle on fresh inputs;
  * y: normwise relative error <= 1e-12 (double) / 1e-5 (single) against the
    reference's CFS kernel output (BASELINE.json north_star) -- the reduction
    order differs, so not bitwise.
"""
import numpy as np
import pytest

import cases
from cfs_spmv_b200 import capi, gen
from oracle import oracle

pytestmark = pytest.mark.gpu

GRAPH_KEYS = ("weight", "adj_ptr", "adj", "color_first", "color")


def _check_metadata(md, ref, who):
    for k in oracle.SCALAR_KEYS:
        assert int(md[k]) == int(ref[k]), (who, k, int(md[k]), int(ref[k]))
    for k in oracle.METADATA_KEYS:
        r = ref[k] if not hasattr(ref, "files") else cases.gold_array(ref, k)
        assert np.array_equal(md[k], r), (who, k)
        assert md[k].size == 0 or md[k].dtype == r.dtype, (who, k)


@pytest.mark.parametrize("case", cases.CASES, ids=cases.case_id)
def test_golden_parity(gpu, case):
    name, P, prec = case
    gold = np.load(cases.golden_path(case))
    rp, ci, v, x = cases.case_inputs(case)
    A = capi.Matrix.from_csr(rp, ci, v)
    A.tune(P)
    _check_metadata(A.metadata(), gold, "golden")
    # conflict graph and colours against the (pinned) oracle
    o = oracle.Oracle(rp, ci, v, P)
    if P > 1:
        for k in GRAPH_KEYS:
            assert np.array_equal(A.export(k), getattr(o, k)), k
    # y, twice on the same buffer (test_spmv_mmf.cpp:80-83), dirty start
    y = np.full(len(x), 123.0, dtype=x.dtype)
    for _ in range(2):
        A.spmv(y, x)
        assert cases.normwise_rel_err(y, gold["y"]) <= cases.TOL[prec]
    # the reference's own acceptance test against plain CSR (isEqual,
    # platform.hpp:33-37; test_spmv_mmf.cpp is hard-wired to double)
    if prec == "d":
        assert np.all(np.abs(y - gold["y_csr"]) <= 1e-8 * np.abs(y) + 1e-300)
    A.close()


@pytest.mark.parametrize("case", [c for c in cases.CASES if c[1] in (1, 4)],
                         ids=cases.case_id)
def test_csr_comparator_kernel(gpu, case):
    """Format::csr path (cpu_mv), the comparator of test_spmv_mmf.cpp:85-89"""
    name, P, prec = case
    gold = np.load(cases.golden_path(case))
    rp, ci, v, x = cases.case_inputs(case)
    A = capi.Matrix.from_csr(rp, ci, v, symmetric=False)
    A.tune(P, tuning=0)
    y = np.zeros(len(x), dtype=x.dtype)
    A.spmv(y, x)
    assert cases.normwise_rel_err(y, gold["y_csr"]) <= cases.TOL[prec]
    inf = A.info()
    assert inf["nnz_full"] == int(gold["nnz_full"])
    A.close()


def _decode_layout(A):
    """execution layout -> list of (row, col, val), for the property test"""
    sp_ = A.export("sell_slice_ptr")
    vrow = A.export("sell_vrow")
    col = A.export("sell_col")
    val = A.export("sell_val")
    rows, cols, vals = [], [], []
    first_seen = {}
    for s in range(len(sp_) - 1):
        w = sp_[s + 1] - sp_[s]
        blk_c = col[sp_[s] * 32:(sp_[s] + w) * 32].reshape(w, 32)
        blk_v = val[sp_[s] * 32:(sp_[s] + w) * 32].reshape(w, 32)
        for lane in range(32):
            tag = vrow[s * 32 + lane]
            if tag < 0:
                assert np.all(blk_c[:, lane] == -1)
                continue
            r = tag & ((1 << 30) - 1)
            cont = bool(tag & (1 << 30))
            assert cont == (r in first_seen), "exactly one first chunk per row"
            first_seen[r] = True
            m = blk_c[:, lane] >= 0
            # real entries first, padding after, never interleaved
            k = int(m.sum())
            assert np.all(m[:k]) and not np.any(m[k:])
            assert k <= 32
            rows += [r] * k
            cols += list(blk_c[:k, lane])
            vals += list(blk_v[:k, lane])
    return np.array(rows), np.array(cols), np.array(vals), first_seen


@pytest.mark.parametrize("name", ["lap27_10", "rmat_9", "ragged_333",
                                  "diag_only_48", "banded_3000"])
def test_layout_holds_exactly_the_lower_triangle(gpu, name):
    rp, ci, v = cases.matrix(name)
    A = capi.Matrix.from_csr(rp, ci, v)
    A.tune(1)
    rows, cols, vals, seen = _decode_layout(A)
    n = len(rp) - 1
    assert sorted(seen) == list(range(n)), "every row owns a first chunk"
    o = oracle.Oracle(rp, ci, v, 1)
    lrp = o.lower_rowptr
    exp_rows = np.repeat(np.arange(n), np.diff(lrp))
    # same multiset, and same order inside every row (chunks are in order)
    assert len(rows) == o.nnz_low
    order = np.argsort(rows, kind="stable")
    assert np.array_equal(rows[order], exp_rows)
    assert np.array_equal(cols[order], o.lower_colind)
    assert np.array_equal(vals[order], o.lower_values)
    inf = A.info()
    assert inf["padded_entries"] >= inf["nnz_low"]
    A.close()


def test_long_rows_are_split(gpu):
    """an arrow matrix: last row is dense (n-1 lower entries) -> many chunks"""
    n = 1000
    rows = np.concatenate([np.arange(n), np.full(n - 1, n - 1), np.arange(n - 1)])
    cols = np.concatenate([np.arange(n), np.arange(n - 1), np.full(n - 1, n - 1)])
    vals = np.concatenate([np.full(n, 8.0), np.linspace(-1, 1, n - 1),
                           np.linspace(-1, 1, n - 1)])
    order = np.lexsort((cols, rows))
    rp = np.zeros(n + 1, np.int64)
    np.add.at(rp, rows + 1, 1)
    rp = np.cumsum(rp).astype(np.int32)
    ci, v = cols[order].astype(np.int32), vals[order]
    x = gen.gen_x(5, n)
    for dt, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
        capi.set_option("rechunk", 0)
        A = capi.Matrix.from_csr(rp, ci, v.astype(dt))
        A.tune(1)
        capi.set_option("rechunk", 1)
        assert A.info()["nvrows"] == (n - 1) + 32  # ceil(999/32) chunks
        A.close()
        # ragged + small bandwidth: shorter chunks (8 = the minimum) for the
        # tile kernel
        A = capi.Matrix.from_csr(rp, ci, v.astype(dt))
        A.tune(1)
        assert A.info()["nvrows"] == (n - 1) + 125  # ceil(999/8) chunks
        y = np.zeros(n, dt)
        A.spmv(y, x.astype(dt))
        o = oracle.Oracle(rp, ci, v.astype(dt), 1)
        assert cases.normwise_rel_err(y, o.spmv(x.astype(dt))) <= tol
        A.close()


@pytest.mark.parametrize("points,n,P", [(7, 40, 8), (27, 32, 37), (27, 48, 148)])
def test_fresh_laplacians_against_oracle(gpu, points, n, P):
    """non-golden sizes, partition counts up to one per SM (148)"""
    spec = capi.GenSpec.laplacian(points, n, n, n)
    rp, ci, v = capi.gen_host_csr(spec)
    if not oracle.valid_partition_count(len(rp) - 1, P):
        pytest.skip("P not valid for the reference")
    x = gen.gen_x(11, len(rp) - 1)
    o = oracle.Oracle(rp, ci, v, P)
    A = capi.Matrix.from_csr(rp, ci, v)
    A.tune(P)
    _check_metadata(A.metadata(), o.metadata(), "oracle")
    for k in GRAPH_KEYS:
        assert np.array_equal(A.export(k), getattr(o, k)), k
    y = np.zeros(len(x))
    A.spmv(y, x)
    assert cases.normwise_rel_err(y, o.spmv(x)) <= 1e-12
    A.close()


def test_device_generated_matrix_equals_host_generated(gpu):
    import torch
    for spec in (capi.GenSpec.laplacian(27, 9, 8, 7),
                 capi.GenSpec.laplacian(7, 6, 5, 11),
                 capi.GenSpec.banded(4000, 300, 152, 9)):
        rp, ci, v = capi.gen_host_csr(spec)
        drp, dci, dv = capi.gen_device_csr(spec)
        assert np.array_equal(drp.cpu().numpy(), rp)
        assert np.array_equal(dci.cpu().numpy(), ci)
        assert np.array_equal(dv.cpu().numpy(), v)
        # shard
        drp, dci, dv = capi.gen_device_csr(spec, 100, 300, is_double=False)
        assert np.array_equal(dci.cpu().numpy(), ci[rp[100]:rp[300]])
        assert np.array_equal(dv.cpu().numpy(),
                              v[rp[100]:rp[300]].astype(np.float32))
    x = capi.gen_device_x(3, 10, 1000)
    assert np.array_equal(x.cpu().numpy(), gen.gen_x(3, 990, begin=10))
    torch.cuda.synchronize()


def test_device_pointers_and_streams(gpu):
    """device-resident CSR in, device vectors, caller's stream (bench path)"""
    import torch
    spec = capi.GenSpec.laplacian(27, 24, 24, 24)
    n = spec.nrows
    drp, dci, dv = capi.gen_device_csr(spec)
    A = capi.Matrix(n, n, drp, dci, dv, True, True)
    A.tune(1)
    del drp, dci, dv  # borrowed only until tune() returns
    torch.cuda.empty_cache()
    x = capi.gen_device_x(4, 0, n)
    y = torch.full((n,), 7.0, dtype=torch.float64, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            A.spmv_async(y, x, s.cuda_stream)
    s.synchronize()
    rp, ci, v = capi.gen_host_csr(spec)
    o = oracle.Oracle(rp, ci, v, 1)
    assert cases.normwise_rel_err(y.cpu().numpy(),
                                  o.spmv(x.cpu().numpy())) <= 1e-12
    A.close()


def test_row_shards_reproduce_the_whole(gpu):
    """multi-GPU layout on one GPU: G row shards, extended local vectors;
    summing the halo contributions reproduces the unsharded y"""
    import torch
    spec = capi.GenSpec.laplacian(27, 12, 12, 20)
    n = spec.nrows
    rp, ci, v = capi.gen_host_csr(spec)
    x = gen.gen_x(8, n)
    o = oracle.Oracle(rp, ci, v, 1)
    y_ref = o.spmv(x)
    for G in (2, 3, 5):
        bounds = [(n * g // G) // 16 * 16 for g in range(G)] + [n]
        y = np.zeros(n)
        for g in range(G):
            b, e = bounds[g], bounds[g + 1]
            srp = (rp[b:e + 1] - rp[b]).astype(np.int32)
            A = capi.Matrix(0, 0, srp, ci[rp[b]:rp[e]].copy(),
                            v[rp[b]:rp[e]].copy(), True, True,
                            shard=(n, b, e))
            A.tune(1)
            inf = A.info()
            h = inf["halo_begin"]
            assert inf["row_begin"] == b and h <= b
            assert h == ((ci[rp[b]:rp[e]].min() if e > b else b) & ~31)
            x_ext = torch.from_numpy(x[h:e].copy()).cuda()
            y_ext = torch.empty(e - h, dtype=torch.float64, device="cuda")
            A.spmv_async(y_ext, x_ext, 0)
            torch.cuda.synchronize()
            y[h:e] += y_ext.cpu().numpy()
            A.close()
        assert cases.normwise_rel_err(y, y_ref) <= 1e-12


def test_call_order_and_argument_errors(gpu):
    rp, ci, v = cases.matrix("lap7_12")
    A = capi.Matrix.from_csr(rp, ci, v)
    y = np.zeros(len(rp) - 1)
    with pytest.raises(capi.CfsError) as e:
        A.spmv(y, y)
    assert e.value.code == capi.CFS_ERR_STATE
    with pytest.raises(capi.CfsError) as e:
        A.tune(512)  # the reference overshoots row_split_ here
    assert e.value.code == capi.CFS_ERR_INVALID
    A.close()
    A = capi.Matrix.from_csr(rp, ci, v)
    A.tune(2)
    with pytest.raises(capi.CfsError) as e:
        A.tune(2)
    assert e.value.code == capi.CFS_ERR_STATE
    A.close()


def test_full_size_config2_properties(gpu):
    """BASELINE config 2 (27-pt 200^3, double) at FULL size, through
    size-independent properties: A*1 is known in closed form, the operator is
    linear, and x2'(A x1) == x1'(A x2)."""
    import torch
    nx = 200
    spec = capi.GenSpec.laplacian(27, nx, nx, nx)
    n = spec.nrows
    drp, dci, dv = capi.gen_device_csr(spec)
    A = capi.Matrix(n, n, drp, dci, dv, True, True)
    A.tune(1)
    del drp, dci, dv
    torch.cuda.empty_cache()
    inf = A.info()
    assert inf["nnz_full"] == (3 * nx - 2) ** 3 == 213847192
    assert inf["nnz_low"] == 102923596 and inf["nnz_diag"] == n
    assert inf["algorithmic_bytes"] == 1459083156
    ones = torch.ones(n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(ones)
    A.spmv_async(y, ones, 0)
    # A*1 = 26 - (#neighbours) = 27 - (count per axis product)
    ax = torch.full((nx,), 3.0, dtype=torch.float64, device="cuda")
    ax[0] = ax[-1] = 2.0
    cnt = ax.view(nx, 1, 1) * ax.view(1, nx, 1) * ax.view(1, 1, nx)
    expect = (27.0 - cnt).reshape(-1)
    assert torch.equal(y, expect)  # small integers: exact in any order
    x1 = capi.gen_device_x(1, 0, n)
    x2 = capi.gen_device_x(2, 0, n)
    y1, y2, y12 = (torch.empty_like(x1) for _ in range(3))
    A.spmv_async(y1, x1, 0)
    A.spmv_async(y2, x2, 0)
    A.spmv_async(y12, 2.0 * x1 - 3.0 * x2, 0)
    torch.cuda.synchronize()
    lin = torch.linalg.norm(y12 - (2.0 * y1 - 3.0 * y2)) / torch.linalg.norm(y12)
    assert lin.item() <= 1e-12
    a, b = torch.dot(x2, y1).item(), torch.dot(x1, y2).item()
    assert abs(a - b) <= 1e-12 * abs(a)
    A.close()


@pytest.mark.parametrize("variant", [1, 2, 4, 5, 6, 50])
@pytest.mark.parametrize("name,prec", [("lap27_14", "d"), ("lap7_12", "s"),
                                       ("banded_3000", "d"), ("rmat_9", "d"),
                                       ("ragged_333", "d"), ("lap27_10", "s")])
def test_every_kernel_variant_matches_the_reference(gpu, variant, name, prec):
    """1: warp per slice, 2: persistent TMA-staged, 4: shared-memory x/y
    windows, 5 (default): compressed index stream + shuffle-merged REDs,
    handing over to 6 (products transposed through shared memory) on irregular
    matrices with bounded column windows; 50 = variant 5 with that switched
    off"""
    rp, ci, v = cases.matrix(name)
    dt = cases.dtype_of(prec)
    x = gen.gen_x(cases.XSEED, len(rp) - 1, dtype=dt)
    o = oracle.Oracle(rp, ci, v.astype(dt), 1)
    ref = o.spmv(x)
    A = capi.Matrix.from_csr(rp, ci, v.astype(dt))
    A.tune(1)
    try:
        capi.set_option("spmv_variant", 5 if variant == 50 else variant)
        capi.set_option("tile6", 0 if variant == 50 else 1)
        y = np.full(len(x), -3.0, dtype=dt)
        for _ in range(2):
            A.spmv(y, x)
            assert cases.normwise_rel_err(y, ref) <= cases.TOL[prec]
    finally:
        capi.set_option("spmv_variant", 5)
        capi.set_option("tile6", 1)
        A.close()


@pytest.mark.parametrize("prec", ["d", "s"])
def test_transposed_tiles_on_a_banded_matrix(gpu, prec):
    """variant 6 on the matrix family it exists for (BASELINE configs[3],
    scaled down): applicable, same y as the oracle, same y with it off"""
    spec = capi.GenSpec.banded(150000, 2000, 152, 7)
    rp, ci, v = capi.gen_host_csr(spec)
    dt = cases.dtype_of(prec)
    n = len(rp) - 1
    x = gen.gen_x(cases.XSEED, n, dtype=dt)
    ref = oracle.Oracle(rp, ci, v.astype(dt), 1).spmv(x)
    A = capi.Matrix.from_csr(rp, ci, v.astype(dt))
    A.tune(1)
    inf = A.info()
    assert inf["transposed_tiles"] == (inf["nslices"] + 31) // 32
    assert 0 < inf["tile_smem_bytes"] <= 96 * 1024
    y = np.full(n, 5.0, dtype=dt)
    A.spmv(y, x)
    A.spmv(y, x)
    assert cases.normwise_rel_err(y, ref) <= cases.TOL[prec]
    capi.set_option("tile6", 0)
    y0 = np.zeros(n, dtype=dt)
    A.spmv(y0, x)
    capi.set_option("tile6", 1)
    assert cases.normwise_rel_err(y0, ref) <= cases.TOL[prec]
    A.close()
    # a matrix whose tiles span too many columns keeps the other variants
    rp, ci, v = gen.rmat(15, 8, seed=1)
    A = capi.Matrix.from_csr(rp, ci, v)
    A.tune(1)
    assert A.info()["transposed_tiles"] == 0
    A.close()


def test_regular_slices_and_windows_are_found(gpu):
    """layout statistics on a stencil: most slices regular (index stream
    compressed to bases), few entries outside the windows"""
    # x-lines of 200 rows: 5 of every 6.25 slices see no grid boundary
    spec = capi.GenSpec.laplacian(27, 200, 12, 8)
    rp, ci, v = capi.gen_host_csr(spec)
    A = capi.Matrix.from_csr(rp, ci, v)
    A.tune(1)
    inf = A.info()
    assert inf["regular_slices"] >= 0.6 * inf["nslices"]
    assert inf["index_rows"] < 0.6 * (inf["padded_entries"] // 32)
    assert inf["ntiles"] == (inf["nslices"] + 3) // 4
    # an irregular matrix has none, and still multiplies correctly (above)
    rp, ci, v = cases.matrix("rmat_9")
    B = capi.Matrix.from_csr(rp, ci, v)
    B.tune(1)
    assert B.info()["regular_slices"] == 0
    A.close()
    B.close()


def test_host_vector_pipeline(gpu):
    """cfs_cuda_spmv with host x / y overlaps H2D, kernel and D2H chunk by
    chunk on matrices in natural row order; same result as the bulk path"""
    spec = capi.GenSpec.laplacian(27, 64, 64, 48)
    rp, ci, v = capi.gen_host_csr(spec)
    n = spec.nrows
    x = gen.gen_x(21, n)
    ref = oracle.Oracle(rp, ci, v, 1).spmv(x)
    A = capi.Matrix.from_csr(rp, ci, v)
    A.tune(1)
    assert A.info()["sort_window"] == 0
    try:
        ys = {}
        for pipeline in (1, 0):
            capi.set_option("pipeline", pipeline)
            y = np.full(n, 9.0)
            for _ in range(3):
                A.spmv(y, x)
            assert cases.normwise_rel_err(y, ref) <= 1e-12
            ys[pipeline] = y
        assert cases.normwise_rel_err(ys[1], ys[0]) <= 1e-13
    finally:
        capi.set_option("pipeline", 1)
        A.close()
    # banded, single precision: rows reach 700 columns down, chunks overlap
    spec = capi.GenSpec.banded(200000, 700, 152, 5)
    rp, ci, v = capi.gen_host_csr(spec, dtype=np.float32)
    x = gen.gen_x(22, spec.nrows, np.float32)
    ref = oracle.Oracle(rp, ci, v, 1).spmv(x)
    capi.set_option("sort_rows", 0)   # keep natural order => pipelined path
    try:
        B = capi.Matrix.from_csr(rp, ci, v)
        B.tune(1)
        y = np.zeros(spec.nrows, np.float32)
        B.spmv(y, x)
        assert cases.normwise_rel_err(y, ref) <= 1e-5
        B.close()
    finally:
        capi.set_option("sort_rows", 1)


def test_hub_columns_run_column_wise(gpu):
    """power-law style input: a few columns hold most of the lower triangle.
    Their transposed term runs column-wise (hubs.cu); result unchanged."""
    rng = np.random.default_rng(3)
    n = 6000
    hubs = [0, 1, 5, 17]
    pairs = set()
    for h in hubs:                       # dense columns below the diagonal
        for i in rng.choice(np.arange(h + 1, n), size=n // 2, replace=False):
            pairs.add((int(i), h))
    for _ in range(4 * n):               # sparse background
        i, j = int(rng.integers(1, n)), int(rng.integers(0, n))
        if j < i:
            pairs.add((i, j))
    hi = np.array([p[0] for p in pairs])
    lo = np.array([p[1] for p in pairs])
    val = rng.uniform(-1, 1, size=len(hi))
    rows = np.concatenate([hi, lo, np.arange(n)])
    cols = np.concatenate([lo, hi, np.arange(n)])
    vals = np.concatenate([val, val, np.full(n, 40.0)])
    order = np.lexsort((cols, rows))
    rp = np.zeros(n + 1, np.int64)
    np.add.at(rp, rows + 1, 1)
    rp = np.cumsum(rp).astype(np.int32)
    ci, v = cols[order].astype(np.int32), vals[order]
    x = gen.gen_x(9, n)
    for dt, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
        ref = oracle.Oracle(rp, ci, v.astype(dt), 1).spmv(x.astype(dt))
        A = capi.Matrix.from_csr(rp, ci, v.astype(dt))
        A.tune(1)
        inf = A.info()
        assert inf["hub_columns"] == len(hubs)
        assert inf["hub_entries"] >= len(hubs) * (n // 2)
        try:
            for use_hubs in (1, 0):
                capi.set_option("hubs", use_hubs)
                y = np.full(n, 5.0, dt)
                for _ in range(2):
                    A.spmv(y, x.astype(dt))
                    assert cases.normwise_rel_err(y, ref) <= tol
        finally:
            capi.set_option("hubs", 1)
            A.close()


@pytest.mark.parametrize("prec", ["d", "s"])
@pytest.mark.parametrize("ndistinct", [1, 2, 3, 200, 256, 257, 100000])
def test_value_indexing_is_lossless(gpu, prec, ndistinct):
    """valindex.cu: <= 256 distinct lower-triangle values (compared bit for
    bit) -> dictionary + one-byte codes; 1 value -> no value stream at all;
    more -> values stay streamed. Same y as the oracle either way."""
    spec = capi.GenSpec.laplacian(27, 200, 12, 8)   # regular slices
    rp, ci, v = capi.gen_host_csr(spec)
    n = len(rp) - 1
    rows = np.repeat(np.arange(n), np.diff(rp))
    lo, hi = np.minimum(rows, ci), np.maximum(rows, ci)
    if ndistinct > 1:   # symmetric: the value depends on the unordered pair
        k = (lo * 7919 + hi * 104729) % ndistinct
        table = -(1.0 + np.arange(ndistinct) / 1024.0)
        if ndistinct == 2:
            table = np.array([0.0, -0.0])   # equal as numbers, different bits
        v = np.where(rows == ci, v, table[k])
    dt = cases.dtype_of(prec)
    v = v.astype(dt)
    x = gen.gen_x(cases.XSEED, n, dtype=dt)
    ref = oracle.Oracle(rp, ci, v, 1).spmv(x)
    A = capi.Matrix.from_csr(rp, ci, v)
    A.tune(1)
    inf = A.info()
    lower = v[ci < rows]
    expect = len(np.unique(lower.view(np.uint64 if dt == np.float64
                                      else np.uint32)))
    assert inf["value_dictionary"] == (expect if expect <= 256 else 0)
    y = np.full(n, 3.0, dtype=dt)
    A.spmv(y, x)
    A.spmv(y, x)
    assert cases.normwise_rel_err(y, ref) <= cases.TOL[prec]
    capi.set_option("value_index", 0)
    y0 = np.zeros(n, dtype=dt)
    A.spmv(y0, x)
    capi.set_option("value_index", 1)
    assert cases.normwise_rel_err(y0, ref) <= cases.TOL[prec]
    # the dictionary changes which bytes are read, not what is computed
    assert cases.normwise_rel_err(y, y0) <= (1e-15 if prec == "d" else 1e-6)
    A.close()


@pytest.mark.parametrize("kind", ["lap27", "banded"])
@pytest.mark.parametrize("chunks,split", [(8, 1), (8, 0), (3, 1), (16, 1),
                                          (0, 1)])
def test_host_vector_pipeline(gpu, chunks, split, kind):
    """cfs_cuda_spmv with HOST x and y on a matrix large enough for the staged
    H2D / kernel / D2H pipeline (>= 4096 slices): pageable vectors (plain
    enqueue) and pinned ones (the step replayed as a CUDA graph, re-captured
    when the pointers change), against the oracle"""
    import torch
    # lap27: natural row order, variant 5; banded: rows length-sorted inside
    # 1024-row windows, split virtual rows, the tile kernel on slice ranges
    spec = (capi.GenSpec.laplacian(27, 64, 64, 64) if kind == "lap27"
            else capi.GenSpec.banded(300000, 2000, 152, 7))
    rp, ci, v = capi.gen_host_csr(spec)
    n = len(rp) - 1
    o = oracle.Oracle(rp, ci, v, 1)
    # chunks == 0: the default, chunk count adapted to the vector size
    capi.set_option("pipeline_adaptive", 1 if chunks == 0 else 0)
    capi.set_option("pipeline_chunks", chunks if chunks else 6)
    capi.set_option("pipeline_split", split)
    try:
        A = capi.Matrix.from_csr(rp, ci, v)
        A.tune(1)
    finally:
        capi.set_option("pipeline_adaptive", 1)
        capi.set_option("pipeline_chunks", 6)
        capi.set_option("pipeline_split", 1)
    for seed in (1, 2):
        x = gen.gen_x(seed, n, np.float64)
        ref = o.spmv(x)
        y = np.full(n, -1.0)
        A.spmv(y, x)                              # pageable
        assert cases.normwise_rel_err(y, ref) <= 1e-12
        xp = torch.from_numpy(x).pin_memory()
        for _ in range(2):                        # new pinned buffers each time
            yp = torch.full((n,), 7.0, dtype=torch.float64).pin_memory()
            A.spmv(yp, xp)
            A.spmv(yp, xp)                        # replay of the captured graph
            assert cases.normwise_rel_err(yp.numpy(), ref) <= 1e-12
    # the unpipelined path agrees
    capi.set_option("pipeline", 0)
    y0 = np.zeros(n)
    A.spmv(y0, x)
    capi.set_option("pipeline", 1)
    assert cases.normwise_rel_err(y0, ref) <= 1e-12
    A.close()
