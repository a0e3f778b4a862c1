"""BASELINE.json configs[0] — "1x512-sigmoid DNN, 257-d in/out, batch 128, synthetic Pfile, CPU forward only (plumbing,
no GPU)": synthetic Pfile -> the re-written host reader (Interface::Readchunk_cv, fea_context=1, no NAT block) -> the
CPU oracle's forward / CrossValid in sigmoid mode, checked against a float64 NumPy restatement of BP_GPU.cu:709-759 +
:458-467 on the rows the reader produced.  Exercises every CPU-side piece of the drop-in surface without a GPU."""
import os
import subprocess
import tempfile

import numpy as np

from reader_case import make_inputs, parse_dump, reader_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "bin", "reader_dump")
C1 = dict(dim=257, out=257, ctx=1, off=0, nat=0, seed=17, lens=[300, 420, 180, 260, 350, 490], traincache=2000,
          train="0-3", cv="4-5", rseed=3, hidden=512)


def test_c1_sigmoid_forward_on_reader_rows(oracle):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "host"), "-s"])
    with tempfile.TemporaryDirectory() as d:
        feas, targs, mu, ivar = make_inputs(d, C1)
        out = os.path.join(d, "dump.bin")
        subprocess.check_call([EXE, out] + reader_args(d, C1), cwd=d)
        chunks = parse_dump(out, C1)
    cv = [c for c in chunks if c[0] == 1]
    assert len(cv) == 1
    _, _, x, t = cv[0]
    # context 1: every frame of the CV sentences is a sample, in file order, normalised; targets untouched
    want_x = np.concatenate([((f - mu) * ivar).astype(np.float32) for f in feas[4:6]])
    assert x.shape == (350 + 490, 257) and np.array_equal(x, want_x)
    assert np.array_equal(t, np.concatenate(targs[4:6]))

    sizes = [257, 512, 257]
    w, b = oracle.glorot_init(sizes, seed=3)
    net = oracle.Net(sizes, 128, activation=1, weights=w, bias=b)
    got = net.forward(x)                                    # 6 full bunches of 128 + a partial one of 72 (BP_GPU.cu:450)
    h = 1.0 / (1.0 + np.exp(-(x.astype(np.float64) @ w[1].astype(np.float64) + b[1])))
    ref = h @ w[2].astype(np.float64) + b[2]
    scale = float(np.sqrt(np.mean(ref ** 2)))
    assert float(np.max(np.abs(got - ref))) <= 2e-5 * scale   # fp32 arithmetic against float64
    sq = net.crossvalid(x, t)
    want_sq = float(np.sum((ref - t) ** 2))
    assert abs(sq - want_sq) <= 1e-4 * want_sq
