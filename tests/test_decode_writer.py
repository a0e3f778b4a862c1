"""Decode output writer (host/DecodeWriter.cc): raw and Pfile formats, optional de-normalisation.
The Pfile it writes must be readable by the same rules the reader applies to its inputs (Interface.cc:468-555)."""
import importlib
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "dnn-for-speech-enhancement_b200")
EXE = os.path.join(PKG, "bin", "decode_dump")
T = importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")


def _expected(dim, n_sent):
    rows, sent, frame = [], [], []
    for s in range(n_sent):
        if s == 1:
            continue
        for f in range(3 + s):
            rows.append(100.0 * s + f + 0.25 * np.arange(dim))
            sent.append(s)
            frame.append(f + 5)
    return np.asarray(rows, np.float32), np.asarray(sent), np.asarray(frame)


def _run(d, fmt, norm, dim, n_sent):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", os.path.join(PKG, "host"), "-s", "../bin/decode_dump"])
    out = os.path.join(d, "dec." + fmt)
    subprocess.check_call([EXE, out, fmt, norm or "-", str(dim), str(n_sent)])
    return out


def test_pfile_output_round_trips_with_sentence_table():
    dim, n_sent = 7, 4
    rows, sent, frame = _expected(dim, n_sent)
    with tempfile.TemporaryDirectory() as d:
        sents, sid, fid = T.read_pfile(_run(d, "pfile", None, dim, n_sent))
    assert len(sents) == n_sent and sents[1].shape[0] == 0          # the sample-less sentence is an empty sentence
    assert np.array_equal(np.vstack([s for s in sents if len(s)]), rows)
    assert np.array_equal(sid, sent) and np.array_equal(fid, frame)


def test_raw_output_and_denormalisation():
    dim, n_sent = 5, 3
    rows, _, _ = _expected(dim, n_sent)
    mean = np.linspace(-2, 2, dim).astype(np.float32)
    inv_std = np.linspace(0.5, 1.5, dim).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        T.write_norm(os.path.join(d, "out.norm"), mean, inv_std)
        plain = np.fromfile(_run(d, "raw", None, dim, n_sent), dtype="<f4").reshape(-1, dim)
        den = np.fromfile(_run(d, "raw", os.path.join(d, "out.norm"), dim, n_sent), dtype="<f4").reshape(-1, dim)
    assert np.array_equal(plain, rows)
    assert np.array_equal(den, (rows / inv_std + mean).astype(np.float32))


def test_reference_parser_reads_the_decode_pfile():
    """The Pfile the decode writer produces goes through the REFERENCE's own Pfile parser and reader (unmodified
    Interface.cc behind oracle/_ref/ref_reader_dump): same frame / sentence counts — the empty sentence included — and
    the rows it hands out are the rows our reader hands out, byte for byte."""
    import pytest
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_reader_dump")
    ours = os.path.join(PKG, "bin", "reader_dump")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/ref_reader_dump not built (needs /root/reference)")
    dim, n_sent = 129, 24   # sentences of 3 .. 26 frames: with an 11-frame context some yield samples, some none
    with tempfile.TemporaryDirectory() as d:
        pf = _run(d, "pfile", None, dim, n_sent)
        T.write_norm(os.path.join(d, "n.norm"), np.zeros(dim, np.float32), np.ones(dim, np.float32))
        args = [f"fea_file={pf}", f"norm_file={d}/n.norm", f"targ_file={pf}", f"outwts_file={d}/o.wts", "initwts_file=",
                "train_sent_range=0-19", "cv_sent_range=20-23", f"fea_dim={dim}", "fea_context=11", "targ_offset=5",
                "traincache=1000", "bunchsize=8", f"layersizes={12 * dim},3,{dim}", "gpu_used=1", "init_randem_seed=1",
                "momentum=0.5", "weightcost=0", "lrate=1", "dropoutflag=0", "visible_omit=0", "hid_omit=0"]
        out = {}
        for tag, exe in (("ref", ref), ("ours", ours)):
            subprocess.run([exe, f"{d}/{tag}.bin"] + args + [f"log_file={d}/{tag}.log"], check=True, timeout=60,
                           stdout=subprocess.DEVNULL)
            out[tag] = (open(f"{d}/{tag}.bin", "rb").read(), [l for l in open(f"{d}/{tag}.log") if l.startswith("Get")])
    frames = sum(3 + s for s in range(n_sent) if s != 1)
    assert f"Get pfile info over: Training data has {frames} frames, {n_sent} sentences.\n" in out["ref"][1]
    assert out["ours"] == out["ref"]
