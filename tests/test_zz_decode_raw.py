"""Decode straight from raw Pfile records without targets (BP_GPU.decode_raw = bp_crossvalid_raw with a null score
pointer and no target records): the enhanced frames must equal, bit for bit, bp_forward on the rows the host reader
assembles.  Added after the last GPU call of round 1 — it runs last in the suite so that `-x` cannot mask the proven
tests behind it."""
import importlib

import numpy as np
import pytest

from reader_case import CASES
from test_raw_reader import _net, _raw, run_raw_dump, splice_numpy

pytestmark = pytest.mark.gpu


def test_decode_raw_equals_forward_on_host_assembled_rows():
    bp = importlib.import_module("dnn-for-speech-enhancement_b200")
    case = CASES["129b"]
    h, chunks = run_raw_dump(case)
    g, _, _ = _net(bp, case, 16)
    for c in chunks:
        if c["n_samples"] == 0:
            continue
        x, _t = splice_numpy(h, c)
        out = g.decode_raw(_raw(bp, h, c, with_targ=False))
        assert out.shape == (c["n_samples"], case["out"])
        assert np.array_equal(out, g.forward(x.shape[0], x))
    with pytest.raises(bp.BpError):   # a score without target records is refused
        g.crossvalid_raw(_raw(bp, h, [c for c in chunks if c["n_samples"] > 0][0], with_targ=False))
    g.close()
