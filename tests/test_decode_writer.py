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
