"""CPU tests of the re-written host data/IO surface (dnn-for-speech-enhancement_b200/host/Interface.cc) — bit-exact
against golden chunks produced by the REFERENCE's own reader (tests/golden/reader_*.npz, generated in the build
container by tests/golden/make_reader_golden.py from unmodified /root/reference/Interface.cc), plus the .wts format."""
import importlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

from reader_case import CASES, layersizes, make_inputs, parse_dump, reader_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = os.path.join(ROOT, "dnn-for-speech-enhancement_b200")
EXE = os.path.join(PKG, "bin", "reader_dump")


def _build():
    if not os.path.exists(EXE) or not os.path.exists(EXE.replace("reader_dump", "prefetch_dump")):
        subprocess.check_call(["make", "-C", os.path.join(PKG, "csrc"), "-s"])
        subprocess.check_call(["make", "-C", os.path.join(PKG, "host"), "-s"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_reader_bit_exact_vs_reference_golden(name):
    _build()
    case = CASES[name]
    gold = np.load(os.path.join(GOLDEN, f"reader_{name}.npz"))
    with tempfile.TemporaryDirectory() as d:
        make_inputs(d, case)
        out = os.path.join(d, "dump.bin")
        subprocess.check_call([EXE, out] + reader_args(d, case), cwd=d)
        chunks = parse_dump(out, case)
    assert len(chunks) == int(gold["n"])
    for i, (kind, cid, x, t) in enumerate(chunks):
        assert kind == int(gold[f"c{i}_kind"]) and cid == int(gold[f"c{i}_id"])
        assert np.array_equal(x.view(np.uint32), gold[f"c{i}_x"].view(np.uint32)), f"chunk {i} inputs differ"
        assert np.array_equal(t.view(np.uint32), gold[f"c{i}_t"].view(np.uint32)), f"chunk {i} targets differ"


@pytest.mark.parametrize("traincache", [7, 16, 40, 1000])
def test_host_reader_prefetch_thread_equals_serial_loop(traincache):
    """BPtrain prefetch=1 reads the host reader's chunks one ahead on a second thread, alternating between two buffer
    pairs (host/ChunkPrefetch.h, Interface::Readchunk(index, slot)): every row of every chunk must be byte-identical
    with the serial loop's, whose rows are pinned against the reference reader above."""
    _build()
    case = dict(CASES["129"], traincache=traincache)
    with tempfile.TemporaryDirectory() as d:
        make_inputs(d, case)
        a, b = os.path.join(d, "serial.bin"), os.path.join(d, "prefetch.bin")
        subprocess.check_call([EXE, a] + reader_args(d, case), cwd=d, stdout=subprocess.DEVNULL)
        subprocess.check_call([EXE.replace("reader_dump", "prefetch_dump"), b] + reader_args(d, case), cwd=d,
                              stdout=subprocess.DEVNULL)
        assert open(a, "rb").read() == open(b, "rb").read()


def test_reader_no_nat_257_matches_numpy_restatement():
    """C2-style input (11 x 257, no NAT block): rows = normalised 11-frame windows, targets = centre frame."""
    _build()
    case = dict(dim=257, out=257, ctx=11, off=5, nat=0, seed=3, lens=[40, 9, 27], traincache=500, train="0-1",
                cv="2-2", rseed=1, hidden=2)
    with tempfile.TemporaryDirectory() as d:
        feas, targs, mu, ivar = make_inputs(d, case)
        out = os.path.join(d, "dump.bin")
        subprocess.check_call([EXE, out] + reader_args(d, case), cwd=d)
        chunks = parse_dump(out, case)
    norm = [((f - mu) * ivar).astype(np.float32) for f in feas]
    # CV chunk is in order: sentence 2 -> 27-10 = 17 samples
    kind, cid, x, t = chunks[-1]
    assert kind == 1 and x.shape == (17, 11 * 257)
    for j in range(17):
        assert np.array_equal(x[j], norm[2][j:j + 11].reshape(-1))
        assert np.array_equal(t[j], targs[2][j + 5])
    # train chunk: 30 samples of sentence 0 (sentence 1 is shorter than the context), shuffled
    kind, cid, x, t = chunks[0]
    assert kind == 0 and x.shape == (30, 11 * 257)
    want = {norm[0][j:j + 11].reshape(-1).tobytes(): targs[0][j + 5] for j in range(30)}
    for j in range(30):
        assert np.array_equal(want.pop(x[j].tobytes()), t[j])
    assert not want


def test_wts_roundtrip_and_layout():
    T = importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")
    rng = np.random.default_rng(0)
    ls = [7, 5, 3]
    w = [None] + [rng.standard_normal((ls[i - 1], ls[i])).astype(np.float32) for i in (1, 2)]
    b = [None] + [rng.standard_normal(ls[i]).astype(np.float32) for i in (1, 2)]
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "a.wts")
        T.write_wts(p, w, b)
        raw = open(p, "rb").read()
        # MAT-v4 header of the first matrix: {10, n_out, n_in, 0, len("weights12")+1}  (Interface.cc:422-431)
        assert np.frombuffer(raw[:20], "<i4").tolist() == [10, 5, 7, 0, 10] and raw[20:30] == b"weights12\0"
        w2, b2 = T.read_wts(p, ls)
    for i in (1, 2):
        assert np.array_equal(w[i], w2[i]) and np.array_equal(b[i], b2[i])
