"""Generates the golden fixtures in this directory by running the UNMODIFIED
reference (compiled into oracle/_ref by oracle/ref_build/Makefile) on every
case of tests/cases.py. Only runnable where /root/reference exists (the build
container); the fixtures it writes are committed and travel everywhere.

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from cfs_spmv_b200 import gen  # noqa: E402
from oracle import oracle  # noqa: E402

KEEP = oracle.METADATA_KEYS + ("x", "y", "y_csr")


def main():
    assert oracle.build_reference(), "reference not buildable here"
    with tempfile.TemporaryDirectory() as tmp:
        for case in cases.CASES:
            name, P, prec = case
            rp, ci, v = cases.matrix(name)
            # half of the generator-backed cases go through the reference's
            # in-memory array ctor from a CSR file, the others through its own
            # call of the shared generator; both must agree with `matrix()`.
            csr_path = os.path.join(tmp, name + ".bin")
            if not os.path.exists(csr_path):
                oracle.write_csr_bin(csr_path, rp, ci, v)
            spec = cases.GEN_SPECS.get(name) if P % 2 == 0 else None
            spec = spec or ("csr:" + csr_path)
            d = oracle.run_ref_dump(spec, P, prec, cases.XSEED,
                                    os.path.join(tmp, "dump.bin"))
            x = gen.gen_x(cases.XSEED, len(rp) - 1, cases.dtype_of(prec))
            assert np.array_equal(d["x"], x), "x generator mismatch"
            out = {k: d[k] for k in KEEP if k in d}
            for k in oracle.SCALAR_KEYS:
                out[k] = np.int64(d[k])
            np.savez_compressed(cases.golden_path(case), **out)
            print("%-22s ncolors=%d nranges=%d nnz_low=%d" % (
                cases.case_id(case), d["ncolors"], d["nranges"], d["nnz_low"]))
    # Matrix Market fixtures: the reference's FILE constructor + loader
    mtx_dir = os.path.join(cases.GOLDEN_DIR, "mtx")
    os.makedirs(mtx_dir, exist_ok=True)
    errors = {}
    with tempfile.TemporaryDirectory() as tmp:
        for fname in sorted(os.listdir(mtx_dir)):
            if not fname.endswith(".mtx"):
                continue
            if fname.startswith("err_"):
                # error fixtures: keep what the reference prints (stdout) and
                # its exit status
                import subprocess
                r = subprocess.run([os.path.join(oracle.REF_DIR, "test_spmv_mmf"),
                                    os.path.join(mtx_dir, fname), "1"],
                                   capture_output=True, text=True,
                                   env=dict(os.environ, CFS_NUM_THREADS="1"))
                errors[fname] = {"exit": r.returncode, "stdout": r.stdout}
                print("%-30s exit=%d %r" % (fname, r.returncode, r.stdout))
                continue
            for P in (1, 3):
                d = oracle.run_ref_dump("mtx:" + os.path.join(mtx_dir, fname),
                                        P, "d", cases.XSEED,
                                        os.path.join(tmp, "dump.bin"))
                out = {k: d[k] for k in KEEP + ("csr_rowptr", "csr_colind",
                                                "csr_values") if k in d}
                for k in oracle.SCALAR_KEYS + ("symmetric", "ncols"):
                    out[k] = np.int64(d[k])
                np.savez_compressed(
                    os.path.join(mtx_dir, "%s-P%d.npz" % (fname[:-4], P)), **out)
                print("%-22s P=%d symmetric=%d nnz_full=%d" % (
                    fname, P, d["symmetric"], d["nnz_full"]))
    import json
    with open(os.path.join(mtx_dir, "errors.json"), "w") as f:
        json.dump(errors, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
