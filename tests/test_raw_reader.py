"""Device-side chunk reader (include/bp_gpu.h bp_raw_chunk).

The host planner (Interface::ReadchunkRaw, dnn-for-speech-enhancement_b200/host/Interface.cc) emits raw Pfile records +
a sample table; bp_splice_kernel (csrc/bp_splice.cuh) assembles the rows.  Both halves are checked bit for bit against
the golden chunks produced by the REFERENCE's own reader (tests/golden/reader_*.npz, see make_reader_golden.py):
  * CPU: the tables, replayed by a numpy restatement of the splice (this file), reproduce the golden rows;
  * GPU: the same tables through bp_upload_raw_chunk + bp_download_chunk reproduce them on the device."""
import importlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

from reader_case import CASES, layersizes, make_inputs, reader_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = os.path.join(ROOT, "dnn-for-speech-enhancement_b200")
EXE = os.path.join(PKG, "bin", "raw_dump")


def _build():
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", os.path.join(PKG, "csrc"), "-s"])
        subprocess.check_call(["make", "-C", os.path.join(PKG, "host"), "-s"])


def parse_raw_dump(path):
    raw = np.fromfile(path, dtype=np.uint8)
    pos = 0

    def take(n, dtype):
        nonlocal pos
        a = raw[pos:pos + 4 * n].view(dtype).copy()
        pos += 4 * n
        return a

    dim, ctx, off, nat, out = (int(v) for v in take(5, np.int32))
    mean, ivar = take(dim, np.float32), take(dim, np.float32)
    chunks = []
    for kind in (0, 1):
        for _ in range(int(take(1, np.int32)[0])):
            cid, nrec, ns = (int(v) for v in take(3, np.int32))
            if ns == 0:
                chunks.append(dict(kind=kind, id=cid, n_records=0, n_samples=0))
                continue
            fea = take(nrec * (dim + 2), np.uint32).reshape(nrec, dim + 2)
            targ = take(nrec * (out + 2), np.uint32).reshape(nrec, out + 2)
            chunks.append(dict(kind=kind, id=cid, n_records=nrec, n_samples=ns, fea=fea, targ=targ,
                               frame=take(ns, np.int32), seg=take(ns, np.int32), row=take(ns, np.int32)))
    assert pos == raw.size
    return dict(dim=dim, ctx=ctx, off=off, nat=nat, out=out, mean=mean, ivar=ivar), chunks


def splice_numpy(h, c):
    """numpy restatement of bp_splice_kernel (= the reader's arithmetic, Interface.cc:745-746, 769-779, 844-846)."""
    dim, ctx = h["dim"], h["ctx"]
    fea = c["fea"][:, 2:].byteswap().view(np.float32)
    norm = ((fea - h["mean"]).astype(np.float32) * h["ivar"]).astype(np.float32)
    pad = np.vstack([norm, np.zeros((6, dim), np.float32)])       # records past the block count as 0
    targ = c["targ"][:, 2:].byteswap().view(np.float32)
    x = np.zeros((c["n_samples"], dim * (ctx + h["nat"])), np.float32)
    t = np.zeros((c["n_samples"], h["out"]), np.float32)
    for i in range(c["n_samples"]):
        f, r = int(c["frame"][i]), int(c["row"][i])
        x[r, :dim * ctx] = norm[f:f + ctx].reshape(-1)
        if h["nat"]:
            s = pad[c["seg"][i]].copy()
            for q in range(1, 6):
                s = (s + pad[c["seg"][i] + q]).astype(np.float32)
            x[r, dim * ctx:] = s / np.float32(6.0)
        t[r] = targ[f + h["off"]]
    return x, t


def run_raw_dump(case):
    _build()
    with tempfile.TemporaryDirectory() as d:
        make_inputs(d, case)
        out = os.path.join(d, "raw.bin")
        subprocess.check_call([EXE, out] + reader_args(d, case), cwd=d, stdout=subprocess.DEVNULL)
        return parse_raw_dump(out)


@pytest.mark.parametrize("name", sorted(CASES))
def test_raw_tables_reproduce_reference_reader(name):
    case = CASES[name]
    gold = np.load(os.path.join(GOLDEN, f"reader_{name}.npz"))
    h, chunks = run_raw_dump(case)
    assert (h["dim"], h["ctx"], h["off"], h["nat"], h["out"]) == (case["dim"], case["ctx"], case["off"],
                                                                 case["nat"], case["out"])
    assert len(chunks) == int(gold["n"])
    for i, c in enumerate(chunks):
        assert c["kind"] == int(gold[f"c{i}_kind"]) and c["id"] == int(gold[f"c{i}_id"])
        assert c["n_samples"] == gold[f"c{i}_x"].shape[0]
        if c["n_samples"] == 0:
            continue
        assert sorted(c["row"].tolist()) == list(range(c["n_samples"]))      # a permutation
        x, t = splice_numpy(h, c)
        assert np.array_equal(x.view(np.uint32), gold[f"c{i}_x"].view(np.uint32)), f"chunk {i} inputs differ"
        assert np.array_equal(t.view(np.uint32), gold[f"c{i}_t"].view(np.uint32)), f"chunk {i} targets differ"


@pytest.mark.parametrize("traincache", [7, 16, 40])
def test_prefetch_thread_reads_the_same_chunks_as_the_serial_loop(traincache):
    """ChunkPrefetcher (host/ChunkPrefetch.h, BPtrain reader=gpu prefetch=1) runs the ReadchunkRaw calls one chunk ahead on
    a second thread: records, tables and chunk order must be byte-identical with the serial loop (prefetch=0), also
    when there are many more chunks than the two slots."""
    _build()
    case = dict(CASES["129"], traincache=traincache)
    with tempfile.TemporaryDirectory() as d:
        make_inputs(d, case)
        dumps = []
        for pf in (0, 1, 1):
            out = os.path.join(d, f"raw{len(dumps)}.bin")
            subprocess.check_call([EXE, out] + reader_args(d, case) + [f"prefetch={pf}"], cwd=d,
                                  stdout=subprocess.DEVNULL)
            dumps.append(open(out, "rb").read())
        n_train = sum(1 for c in parse_raw_dump(os.path.join(d, "raw0.bin"))[1] if c["kind"] == 0)
    assert n_train >= {7: 5, 16: 4, 40: 2}[traincache]
    assert dumps[0] == dumps[1] == dumps[2]


def _net(bp, case, bunch, **kw):
    ls = layersizes(case)
    rng = np.random.default_rng(1)
    w = [None] + [(rng.standard_normal((ls[i - 1], ls[i])) * 0.05).astype(np.float32) for i in range(1, len(ls))]
    b = [None] + [np.zeros(ls[i], np.float32) for i in range(1, len(ls))]
    return bp.BP_GPU(1, len(ls), ls, bunch, 0.5, 0.9, 0.0, w, b, device=0, **kw), w, b


def _raw(bp, h, c, with_targ=True):
    return bp.RawChunk(h["dim"], h["ctx"], h["off"], h["nat"], c["fea"], c["targ"] if with_targ else None, h["mean"],
                       h["ivar"], c["frame"], c["seg"], c["row"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_reader_bit_exact_vs_reference_golden(name):
    bp = importlib.import_module("dnn-for-speech-enhancement_b200")
    case = CASES[name]
    gold = np.load(os.path.join(GOLDEN, f"reader_{name}.npz"))
    h, chunks = run_raw_dump(case)
    g, _, _ = _net(bp, case, 8)
    for i, c in enumerate(chunks):
        if c["n_samples"] == 0:
            continue
        g.upload_raw_chunk(_raw(bp, h, c))
        x, t = g.download_chunk(0, c["n_samples"])
        assert np.array_equal(x.view(np.uint32), gold[f"c{i}_x"].view(np.uint32)), f"chunk {i} inputs differ"
        assert np.array_equal(t.view(np.uint32), gold[f"c{i}_t"].view(np.uint32)), f"chunk {i} targets differ"
    g.close()


@pytest.mark.gpu
def test_train_raw_equals_train_on_host_assembled_rows():
    """bp_train_raw(records, table) == bp_train(rows the host reader assembles), weight for weight, bit for bit;
    same for bp_crossvalid_raw.  Also the 3xTF32 mode (the splice feeds the lo-part split)."""
    bp = importlib.import_module("dnn-for-speech-enhancement_b200")
    case = CASES["129b"]
    h, chunks = run_raw_dump(case)
    for math in (0, 1):
        a, _, _ = _net(bp, case, 16, math_mode=math)
        b, _, _ = _net(bp, case, 16, math_mode=math)
        for c in chunks:
            if c["n_samples"] == 0 or c["kind"] != 0:
                continue
            x, t = splice_numpy(h, c)
            a.train(x.shape[0], x, t)
            b.train_raw(_raw(bp, h, c))
        wa, ba = a.returnWeights()
        wb, bb = b.returnWeights()
        for l in range(1, len(wa)):
            assert np.array_equal(wa[l], wb[l]) and np.array_equal(ba[l], bb[l]), f"layer {l} (math {math})"
        cv = [c for c in chunks if c["kind"] == 1 and c["n_samples"] > 0][0]
        x, t = splice_numpy(h, cv)
        s_raw, out_raw = b.crossvalid_raw(_raw(bp, h, cv), want_out=True)
        assert s_raw == a.CrossValid(x.shape[0], x, t)
        assert np.array_equal(out_raw, a.forward(x.shape[0], x))
        a.close()
        b.close()


@pytest.mark.gpu
def test_raw_chunk_argument_checks():
    bp = importlib.import_module("dnn-for-speech-enhancement_b200")
    case = CASES["129"]
    h, chunks = run_raw_dump(case)
    c = [c for c in chunks if c["n_samples"] > 0][0]
    g, _, _ = _net(bp, case, 8)
    bad = dict(c)
    bad["frame"] = c["frame"].copy()
    bad["frame"][0] = c["n_records"] - 3          # window runs past the record block
    with pytest.raises(bp.BpError):
        g.upload_raw_chunk(_raw(bp, h, bad))
    h2 = dict(h)
    h2["ctx"] = h["ctx"] - 1                       # layersizes[0] no longer matches
    with pytest.raises(bp.BpError):
        g.upload_raw_chunk(_raw(bp, h2, c))
    g.close()
