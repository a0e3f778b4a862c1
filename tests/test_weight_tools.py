"""Weight tooling: gen / extend / tomat against the behaviour of the reference's Gen_rand_net.cpp,
Extend_rand_net.cpp and change_cudaSavedModels2matlabWeigths_4layers.m (cited in tools/weights.py)."""
import importlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W = importlib.import_module("dnn-for-speech-enhancement_b200.tools.weights")
T = importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")


def test_gen_ranges_and_file_round_trip():
    ls = [40, 64, 32, 9]
    ws, bs = W.gen_rand_net(ls, flag=1, beta=0.5, seed=3)
    for i in range(1, len(ls)):
        r = 0.5 * np.sqrt(6.0) / np.sqrt(ls[i - 1] + ls[i])
        assert ws[i].shape == (ls[i - 1], ls[i]) and np.abs(ws[i]).max() <= r * (1 + 1e-6)
        assert np.abs(ws[i]).max() > 0.9 * r and abs(float(ws[i].mean())) < 0.1 * r   # fills the interval, centred
        assert not bs[i].any()
    w0, _ = W.gen_rand_net(ls, flag=0, beta=1.0, seed=3)
    assert np.abs(w0[1]).max() <= 1.0 / np.sqrt(ls[0]) * (1 + 1e-6)
    with tempfile.TemporaryDirectory() as d:
        T.write_wts(f"{d}/a.wts", ws, bs)
        rw, rb = T.read_wts(f"{d}/a.wts", ls)
    assert all(np.array_equal(ws[i], rw[i]) and np.array_equal(bs[i], rb[i]) for i in range(1, len(ls)))


def test_extend_keeps_old_block_and_fills_new_parts():
    ori, add = [10, 16, 12, 5], [0, 8, 4, 0]
    ws, bs = W.gen_rand_net(ori, seed=1)
    bs = [None] + [np.arange(n, dtype=np.float32) for n in ori[1:]]
    nws, nbs, new = W.extend_rand_net(ws, bs, ori, add, beta=0.5, seed=2)
    assert new == [10, 24, 16, 5]
    for i in range(1, len(ori)):
        oi, oo = ori[i - 1], ori[i]
        assert nws[i].shape == (new[i - 1], new[i])
        assert np.array_equal(nws[i][:oi, :oo], ws[i])                      # w[m*new_n_out + n] = old[m*old_n_out + n]
        assert np.array_equal(nbs[i][:oo], bs[i]) and not nbs[i][oo:].any()
        r = 0.5 * np.sqrt(6.0) / np.sqrt(new[i - 1] + new[i])
        fresh = np.ones_like(nws[i], dtype=bool)
        fresh[:oi, :oo] = False
        if fresh.any():
            assert np.all(nws[i][fresh] != 0) and np.abs(nws[i][fresh]).max() <= r * (1 + 1e-6)


def test_cli_gen_extend_tomat():
    from scipy.io import loadmat
    mod = "dnn-for-speech-enhancement_b200.tools.weights"
    with tempfile.TemporaryDirectory() as d:
        run = lambda *a: subprocess.check_call([sys.executable, "-m", mod, *a], cwd=ROOT, stdout=subprocess.DEVNULL)
        run("gen", "12,20,7", f"{d}/a.wts", "--seed", "5")
        run("extend", f"{d}/a.wts", "12,20,7", "0,4,0", f"{d}/b.wts")
        run("tomat", f"{d}/b.wts", "12,24,7", f"{d}/se.mat")
        ws, bs = T.read_wts(f"{d}/b.wts", [12, 24, 7])
        m = loadmat(f"{d}/se.mat")
    for i in (1, 2):
        assert m[f"w{i}"].shape == (ws[i].shape[0] + 1, ws[i].shape[1])
        assert np.array_equal(m[f"w{i}"][:-1], ws[i]) and np.array_equal(m[f"w{i}"][-1], bs[i])
