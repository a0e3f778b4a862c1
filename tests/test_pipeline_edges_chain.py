"""GPU tests of the decode pipeline, the chunk prefetch thread, edge shapes and the chained launches:
  * decode straight from raw Pfile records without targets (BP_GPU.decode_raw = bp_crossvalid_raw with a null score
    pointer and no target records): bit-identical with bp_forward on the rows the host reader assembles;
  * the pipelined decode (bp_decode_raw_submit / _wait, two chunks in flight) against the synchronous one;
  * edge shapes against the oracle, with one launch per product and with the chained launches; the chained launches
    bit-identical with the per-product kernels at the BASELINE shapes;
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


def test_pipelined_forward_of_spliced_rows():
    """bp_forward_submit / bp_forward_wait (the pipeline for callers that keep the host reader) == bp_forward."""
    bp = importlib.import_module("dnn-for-speech-enhancement_b200")
    case = CASES["129b"]
    h, chunks = run_raw_dump(case)
    chunks = [c for c in chunks if c["n_samples"] > 0]
    g, _, _ = _net(bp, case, 16)
    xs = [np.ascontiguousarray(splice_numpy(h, c)[0]) for c in chunks]
    want = [g.forward(x.shape[0], x) for x in xs]
    outs = [bp.PinnedArray((x.shape[0], case["out"])) for x in xs]
    g.forward_submit(xs[0].shape[0], xs[0], outs[0].array)
    for k in range(1, len(xs)):
        g.forward_submit(xs[k].shape[0], xs[k], outs[k].array)
        g.forward_wait()
        assert np.array_equal(outs[k - 1].array, want[k - 1])
    g.forward_wait()
    assert np.array_equal(outs[-1].array, want[-1])
    with pytest.raises(bp.BpError):
        g.forward_wait()
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


EDGE_CASES = [("bunch 37, ragged tail", [75, 96, 33], 37, 3 * 37 + 5, dict(lrate=0.7, momentum=0.9)),
              ("bunch 1", [40, 24, 8], 1, 5, dict(lrate=0.1, momentum=0.5)),
              ("one weight layer", [129, 65], 32, 96, dict(momentum=0.9)),
              ("nine weight layers", [64, 72, 40, 96, 33, 80, 48, 56, 64, 20], 32, 64, dict(lrate=0.5)),
              ("one output unit", [90, 70, 1], 32, 64, dict(momentum=0.9)),
              ("chunk shorter than a bunch", [75, 96, 33], 64, 20, dict()),
              ("weight cost, odd sizes", [257, 131, 67, 3], 40, 120, dict(lrate=0.5, momentum=0.9, weightcost=1e-3)),
              ("sigmoid, odd sizes", [61, 45, 29], 24, 72, dict(activation=1, momentum=0.9))]


@pytest.mark.parametrize("chain", [0, 1], ids=["per-product", "chained"])
@pytest.mark.parametrize("case", EDGE_CASES, ids=[c[0] for c in EDGE_CASES])
def test_edge_shapes_against_the_oracle(case, chain):
    """Odd and unit bunches, 1 and 9 weight layers, a 1-unit output, a chunk shorter than a bunch (= no-op like
    BP_GPU.cu:297-318), forward of 1 frame, CV of a ragged tail — against the tf32-conditioned oracle, with one launch
    per product and with the chained launches (csrc/bp_chain.cuh)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import gpu_edge_cases as E
    name, sizes, bunch, n_frames, kw = case
    assert E.case(name, sizes, bunch, n_frames, chain, **kw)


CHAIN_CASES = [("tiny ragged", [75, 96, 33], 37, 4, {}),
               ("one layer", [129, 65], 32, 3, {}),
               ("nine layers", [64, 72, 40, 96, 33, 80, 48, 56, 64, 20], 32, 2, {}),
               ("odd sizes, weight cost", [300, 261, 131, 33], 96, 3, dict(weightcost=1e-3)),
               ("dropout", [129, 70, 50, 20], 48, 4, dict(dropoutflag=1, visible_omit=0.1, hid_omit=0.3)),
               ("sigmoid", [61, 45, 29], 24, 3, dict(activation=1)),
               ("3xTF32", [300, 512, 129], 256, 2, dict(math_mode=1)),
               ("C2, 3 bunches", [2827, 2048, 2048, 2048, 257], 1024, 3, {}),
               ("C3 dropout", [3084, 2048, 2048, 2048, 257], 2048, 2, dict(dropoutflag=1, visible_omit=0.2, hid_omit=0.2)),
               ("C4 per-GPU share", [2827, 2048, 2048, 2048, 2048, 2048, 257], 512, 2, {})]


@pytest.mark.parametrize("case", CHAIN_CASES, ids=[c[0] for c in CHAIN_CASES])
def test_chained_launches_are_bit_identical_with_one_launch_per_product(case):
    """bp_chain_kernel (the forward products of a bunch in one launch, the dX chain + all dW products in a second one)
    runs the same tiles with the same k order and the same epilogues as the per-product kernels: the trained weights
    must be IDENTICAL (with the output layer's split-K off on the per-product side, which the chain does not use)."""
    import oracle_py as O
    bp = importlib.import_module("dnn-for-speech-enhancement_b200")
    _name, sizes, bunch, nb, kw = case
    w, b = O.glorot_init(sizes, seed=3)
    x, t = O.synth_data(bunch * nb + 3, sizes[0], sizes[-1], seed=11)
    res = []
    for chain in (0, 1):
        g = bp.BP_GPU(1, len(sizes), sizes, bunch, 1.0, 0.9, kw.get("weightcost", 0.0), w, b, kw.get("dropoutflag", 0),
                      kw.get("visible_omit", 0.0), kw.get("hid_omit", 0.0), device=0, seed=777,
                      activation=kw.get("activation", 0), math_mode=kw.get("math_mode", 0))
        g.set_option("chain", chain)
        g.set_option("splitk", 0)
        assert g.get_option("chain") == chain
        g.train(x.shape[0], x, t)
        res.append(g.returnWeights())
        g.close()
    for l in range(1, len(sizes)):
        assert np.array_equal(res[0][0][l], res[1][0][l]) and np.array_equal(res[0][1][l], res[1][1][l]), f"layer {l}"
        assert np.isfinite(res[1][0][l]).all() and not np.array_equal(res[1][0][l], w[l])

