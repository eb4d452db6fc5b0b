"""Golden fixtures of the reference's NON-symmetric path (Format::csr):
row_split_ of partition_by_nnz / partition_by_nrows and y of cpu_mv, from the
UNMODIFIED reference (oracle/_ref/ref_tool dumpcsr) on tests/cases.CSR_CASES.
Only runnable where /root/reference exists.

    python tests/golden/make_golden_csr.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    assert oracle.build_reference(), "reference not buildable here"
    os.makedirs(os.path.join(cases.GOLDEN_DIR, "csr"), exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        for case in cases.CSR_CASES:
            name, P, prec, tuning = case
            rp, ci, v = cases.general_matrix(name)
            path = os.path.join(tmp, name + ".bin")
            oracle.write_csr_bin(path, rp, ci, v)
            d = oracle.run_ref_dump_csr("csr:" + path, P, prec, cases.XSEED,
                                        os.path.join(tmp, "dump.bin"), tuning)
            out = {"y": d["y"], "nnz_full": np.int64(d["nnz_full"]),
                   "size_bytes": np.int64(d["size_bytes"])}
            if "row_split" in d:
                out["row_split"] = d["row_split"]
            np.savez_compressed(cases.csr_golden_path(case), **out)
            print("%-30s nnz=%d row_split=%s" % (
                cases.csr_case_id(case), d["nnz_full"],
                list(d.get("row_split", []))[:6]))


if __name__ == "__main__":
    main()
