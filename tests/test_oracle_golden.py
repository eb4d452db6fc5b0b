"""Pins the CPU oracle (oracle/cfs_oracle.c) to the UNMODIFIED reference:
every golden fixture under tests/golden/ is a dump of the compiled reference
(tests/golden/make_golden.py). Metadata and y must match BIT FOR BIT."""
import os

import numpy as np
import pytest

import cases
from cfs_spmv_b200 import gen
from oracle import oracle


@pytest.mark.parametrize("case", cases.CASES, ids=cases.case_id)
def test_oracle_matches_reference_dump(case):
    name, P, prec = case
    gold = np.load(cases.golden_path(case))
    rp, ci, v, x = cases.case_inputs(case)
    assert np.array_equal(gold["x"], x)
    o = oracle.Oracle(rp, ci, v, P)
    md = o.metadata()
    for k in oracle.SCALAR_KEYS:
        assert int(gold[k]) == int(md[k]), k
    for k in oracle.METADATA_KEYS:
        g = cases.gold_array(gold, k)
        assert g.size == 0 or g.dtype == md[k].dtype, k
        assert np.array_equal(g, md[k]), k
    # y: bitwise (same operation order, no FMA contraction on either side)
    y = o.spmv(x)
    assert y.tobytes() == gold["y"].tobytes()
    # second call on a dirty y, like test_spmv_mmf.cpp:80-83
    y2 = o.spmv(x, y0=y)
    assert y2.tobytes() == gold["y"].tobytes()
    assert o.csr_spmv(x).tobytes() == gold["y_csr"].tobytes()
    # and the reference's own acceptance check (isEqual, platform.hpp:27-37)
    eps = 1e-8 if prec == "d" else 1e-4
    assert np.all(np.abs(y - gold["y_csr"]) <= eps * np.abs(y))


def test_partition_formula():
    # row_split_ = t*S, S = ((N/P-1)|15)+1 (csr_matrix.tpp:418-423); the probe
    # value quoted in SURVEY.md 8(a5)
    rs = oracle.partition_by_nrows(1000000, 8)
    assert list(rs) == [0, 125008, 250016, 375024, 500032, 625040, 750048,
                        875056, 1000000]
    assert not oracle.valid_partition_count(1000000, 512)


def test_oracle_against_live_reference_when_present(tmp_path):
    """where the compiled reference exists (build container or a box that
    received oracle/_ref), cross-check one fresh, non-golden case"""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built here")
    from cfs_spmv_b200 import capi, gen
    spec = capi.GenSpec.laplacian(27, 11, 9, 13)
    rp, ci, v = capi.gen_host_csr(spec)
    for P in (5, 12):
        d = oracle.run_ref_dump(spec.ref_tool_spec(), P, "d", 99,
                                str(tmp_path / "d.bin"))
        x = gen.gen_x(99, len(rp) - 1)
        o = oracle.Oracle(rp, ci, v, P)
        md = o.metadata()
        for k in oracle.METADATA_KEYS:
            assert np.array_equal(d[k], md[k]), k
        assert int(d["ncolors"]) == o.ncolors
        assert o.spmv(x).tobytes() == d["y"].tobytes()


@pytest.mark.parametrize("case", cases.CSR_CASES, ids=cases.csr_case_id)
def test_oracle_csr_path_matches_reference(case):
    """the NON-symmetric path: row_split_ of partition_by_nnz (Aggressive) /
    partition_by_nrows (None) and y of cpu_mv, bitwise"""
    name, P, prec, tuning = case
    gold = np.load(cases.csr_golden_path(case))
    rp, ci, v = cases.general_matrix(name)
    dt = cases.dtype_of(prec)
    n = len(rp) - 1
    assert int(gold["nnz_full"]) == rp[-1]
    if P > 1:
        split = (oracle.partition_by_nnz(rp, P) if tuning == "A"
                 else oracle.partition_by_nrows(n, P))
        assert np.array_equal(split, gold["row_split"])
    x = gen.gen_x(cases.XSEED, n, dtype=dt)
    y = oracle.csr_spmv(rp, ci, v.astype(dt), x)
    assert y.tobytes() == gold["y"].tobytes()
