"""Generates tests/golden/reader_129.npz by running the REFERENCE's own reader (Interface.cc, compiled unmodified by
oracle/build_ref.sh into oracle/_ref/ref_reader_dump) on a deterministic synthetic Pfile pair.  Runs only where
/root/reference exists (the build container); the fixture it writes is committed and travels.

    bash oracle/build_ref.sh && python tests/golden/make_reader_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from reader_case import CASES, make_inputs, parse_dump, reader_args  # noqa: E402


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_reader_dump")
    for name, case in CASES.items():
        with tempfile.TemporaryDirectory() as d:
            make_inputs(d, case)
            out = os.path.join(d, "dump.bin")
            subprocess.check_call([exe, out] + reader_args(d, case), cwd=d)
            chunks = parse_dump(out, case)
            arrs = {}
            for i, (kind, cid, x, t) in enumerate(chunks):
                arrs[f"c{i}_kind"], arrs[f"c{i}_id"], arrs[f"c{i}_x"], arrs[f"c{i}_t"] = kind, cid, x, t
            np.savez_compressed(os.path.join(HERE, f"reader_{name}.npz"), n=len(chunks), **arrs)
            print(name, [(k, c, x.shape) for k, c, x, t in chunks])


if __name__ == "__main__":
    main()
