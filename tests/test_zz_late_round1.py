"""GPU tests added after the last GPU call of round 1 (they have not run on a B200 yet) — this file sorts last so that
`-x` cannot mask the proven tests behind them.
  * decode straight from raw Pfile records without targets (BP_GPU.decode_raw = bp_crossvalid_raw with a null score
    pointer and no target records): bit-identical with bp_forward on the rows the host reader assembles;
  * the pipelined decode (bp_decode_raw_submit / _wait, two chunks in flight) against the synchronous one;
  * the ReLU bit mask (EPI_FWD_HID_MASK / EPI_DX_MASK): the two epilogues in isolation, and training with
    bp_set_option("relu_mask", 1) bit-identical with the default;
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


@pytest.mark.parametrize("shape", [(128, 128, 64), (257, 100, 96), (64, 37, 40), (300, 1030, 515), (2048, 1024, 257)])
def test_relu_mask_epilogues_in_isolation(shape):
    """kind 6 (dX product reading the bit mask) == kind 1 (reading Y) bit for bit; kind 7 (forward product leaving
    the mask) == kind 0 bit for bit, with the library checking every mask bit against (y > 0)."""
    import ctypes as C
    bp = importlib.import_module("dnn-for-speech-enhancement_b200")
    lib = bp.load_library()
    M, N, K = shape
    rng = np.random.default_rng(M + 3 * N + 7 * K)
    fp = C.POINTER(C.c_float)

    def run(kind, A, B, bias, aux):
        out = np.full((N, M), np.nan, dtype=np.float32)
        rc = lib.bp_debug_gemm(kind, M, N, K, A.ctypes.data_as(fp), A.shape[1], B.ctypes.data_as(fp), B.shape[1],
                               out.ctypes.data_as(fp), M, None if bias is None else bias.ctypes.data_as(fp),
                               None if aux is None else aux.ctypes.data_as(fp), M, 1.0, 0, 0, None)
        assert rc == 0, lib.bp_last_error().decode()
        return out

    # dX: A = W (M x K), B = D (N x K), aux = Y with zeros, negatives and -0.0 in it
    A = rng.standard_normal((M, K), dtype=np.float32)
    B = rng.standard_normal((N, K), dtype=np.float32)
    Y = rng.standard_normal((N, M), dtype=np.float32)
    Y[rng.random((N, M)) < 0.3] = 0.0
    Y[0, 0] = -0.0
    with_y, with_mask = run(1, A, B, None, Y), run(6, A, B, None, Y)
    assert np.array_equal(with_y.view(np.uint32), with_mask.view(np.uint32))
    assert np.all(with_mask[Y <= 0] == 0)
    # forward: A = W (K x M), B = X (N x K)
    A = rng.standard_normal((K, M), dtype=np.float32)
    bias = rng.standard_normal(M).astype(np.float32)
    plain, masked = run(0, A, B, bias, None), run(7, A, B, bias, None)
    assert np.array_equal(plain.view(np.uint32), masked.view(np.uint32))


@pytest.mark.parametrize("case", ["ragged", "pairs", "dropout"])
def test_relu_mask_training_is_bit_identical(case):
    bp = importlib.import_module("dnn-for-speech-enhancement_b200")
    import oracle_py as O
    sizes, bunch, nb, kw = {"ragged": ([75, 96, 70, 33], 37, 4, {}),
                            "pairs": ([300, 2048, 512, 33], 1024, 2, {}),
                            "dropout": ([129, 70, 50, 20], 48, 5, dict(dropoutflag=1, visible_omit=0.1, hid_omit=0.3))}[case]
    w, b = O.glorot_init(sizes, seed=3)
    x, t = O.synth_data(bunch * nb, sizes[0], sizes[-1], seed=11)
    res = []
    for mask in (0, 1):
        g = bp.BP_GPU(1, len(sizes), sizes, bunch, 1.0, 0.9, 0.0, w, b, kw.get("dropoutflag", 0),
                      kw.get("visible_omit", 0.0), kw.get("hid_omit", 0.0), device=0, seed=777)
        g.set_option("relu_mask", mask)
        g.train(bunch * nb, x, t)
        res.append(g.returnWeights())
        g.close()
    for l in range(1, len(sizes)):
        assert np.array_equal(res[0][0][l], res[1][0][l]) and np.array_equal(res[0][1][l], res[1][1][l]), f"layer {l}"
        assert not np.array_equal(res[0][0][l], w[l])   # it did train


EDGE_CASES = [("bunch 37, ragged tail", [75, 96, 33], 37, 3 * 37 + 5, dict(lrate=0.7, momentum=0.9)),
              ("bunch 1", [40, 24, 8], 1, 5, dict(lrate=0.1, momentum=0.5)),
              ("one weight layer", [129, 65], 32, 96, dict(momentum=0.9)),
              ("nine weight layers", [64, 72, 40, 96, 33, 80, 48, 56, 64, 20], 32, 64, dict(lrate=0.5)),
              ("one output unit", [90, 70, 1], 32, 64, dict(momentum=0.9)),
              ("chunk shorter than a bunch", [75, 96, 33], 64, 20, dict()),
              ("weight cost, odd sizes", [257, 131, 67, 3], 40, 120, dict(lrate=0.5, momentum=0.9, weightcost=1e-3)),
              ("sigmoid, odd sizes", [61, 45, 29], 24, 72, dict(activation=1, momentum=0.9))]


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("case", EDGE_CASES, ids=[c[0] for c in EDGE_CASES])
def test_edge_shapes_against_the_oracle(case, fused):
    """The shapes of scripts/gpu_edge_cases.py (odd and unit bunches, 1 and 9 weight layers, a 1-unit output, a chunk
    shorter than a bunch = no-op like BP_GPU.cu:297-318, forward of 1 frame, CV of a ragged tail) against the
    tf32-conditioned oracle, with the separate and the fused update."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import gpu_edge_cases as E
    name, sizes, bunch, n_frames, kw = case
    assert E.case(name, sizes, bunch, n_frames, fused, **kw)
