"""GPU tests added after the last GPU call of round 1 (they have not run on a B200 yet) — this file sorts last so that
`-x` cannot mask the proven tests behind them.
  * decode straight from raw Pfile records without targets (BP_GPU.decode_raw = bp_crossvalid_raw with a null score
    pointer and no target records): bit-identical with bp_forward on the rows the host reader assembles;
  * the pipelined decode (bp_decode_raw_submit / _wait, two chunks in flight) against the synchronous one;
  * BPtrain's chunk prefetch thread (prefetch=1, the default) trains the very same epoch as the serial loop
    (prefetch=0), for both readers: byte-identical .wts, identical CV score."""
import importlib
import os
import subprocess
import tempfile

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


def test_pipelined_decode_two_chunks_in_flight():
    """bp_decode_raw_submit / bp_decode_raw_wait: results of a pipelined stream equal the synchronous decode, a third
    chunk in flight and a wait without a chunk are refused."""
    bp = importlib.import_module("dnn-for-speech-enhancement_b200")
    case = CASES["129b"]
    h, chunks = run_raw_dump(case)
    chunks = [c for c in chunks if c["n_samples"] > 0]
    g, _, _ = _net(bp, case, 16)
    want = [g.decode_raw(_raw(bp, h, c, with_targ=False)) for c in chunks]
    stream = [chunks[i % len(chunks)] for i in range(7)]
    outs = [bp.PinnedArray((c["n_samples"], case["out"])) for c in stream]
    with pytest.raises(bp.BpError):
        g.decode_raw_wait()
    g.decode_raw_submit(_raw(bp, h, stream[0], with_targ=False), outs[0].array)
    for k in range(1, len(stream)):
        g.decode_raw_submit(_raw(bp, h, stream[k], with_targ=False), outs[k].array)
        if k == 1:
            with pytest.raises(bp.BpError):   # two in flight already
                g.decode_raw_submit(_raw(bp, h, stream[k], with_targ=False), outs[k].array)
        g.decode_raw_wait()
        assert np.array_equal(outs[k - 1].array, want[(k - 1) % len(chunks)])
    g.decode_raw_wait()
    assert np.array_equal(outs[-1].array, want[(len(stream) - 1) % len(chunks)])
    g.close()


@pytest.mark.parametrize("reader", ["host", "gpu"])
def test_bptrain_prefetch_thread_trains_the_same_epoch(reader):
    from test_cli_gpu import OURS, _args, _cv
    if not os.path.exists(OURS):
        pytest.skip("BPtrain not built")
    T = importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")
    with tempfile.TemporaryDirectory() as d:
        feas, targs, mu, ivar = T.synth_corpus(40, 129, 129, seed=5, min_len=30, max_len=90)
        T.write_pfile(f"{d}/fea.pfile", feas)
        T.write_pfile(f"{d}/targ.pfile", targs)
        T.write_norm(f"{d}/fea.norm", mu, ivar)
        for pf in (0, 1):
            extra = [f"reader={reader}", f"prefetch={pf}", "traincache=400", "dropoutflag=1", "visible_omit=0.1",
                     "hid_omit=0.2"]
            a = [x for x in _args(d, f"pf{pf}", extra)
                 if x not in ("traincache=3000", "dropoutflag=0", "visible_omit=0", "hid_omit=0")]
            o = subprocess.run([OURS] + a, cwd=d, capture_output=True, text=True, timeout=600)
            assert o.returncode == 1, o.stdout + o.stderr
        log = open(f"{d}/pf1.log").read()
        assert "Starting chunk 4 of" in log          # several chunks went through the two slots
        assert open(f"{d}/pf0.wts", "rb").read() == open(f"{d}/pf1.wts", "rb").read()
        assert _cv(f"{d}/pf0.log") == _cv(f"{d}/pf1.log")
