"""Differential test of the weight side of the host surface against the REFERENCE's own Interface.cc, live
(oracle/_ref/ref_weights_dump = tests/native/weights_dump.cc linked against the unmodified reference object):
  * random initialisation (srand48(init_randem_seed) -> drand48 per weight, Interface.cc:338-352, 1036-1042): the very
    same floats, for several seeds, ranges and shapes;
  * Writeweights (Interface.cc:411-465): byte-identical MAT-v4 .wts files;
  * initwts_file: a .wts written by either side loads to the same arrays on both sides, and our Python tools
    (tools/pfile.py read_wts / write_wts) agree with both.
Skipped where the reference binary is absent."""
import importlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

from reader_case import make_inputs, reader_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_weights_dump")
OURS = os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "bin", "weights_dump")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ref_weights_dump not built")

CASE = dict(dim=129, out=129, ctx=11, off=5, nat=1, seed=1, lens=[30, 20, 15], traincache=100, train="0-1", cv="2-2",
            rseed=7, hidden=5)


LOGS = {}   # log text of the last run per side, with the side's file names neutralised


def _run(exe, d, tag, ls, extra):
    args = [a for a in reader_args(d, CASE) if not a.startswith(("layersizes=", "outwts_file=", "log_file="))]
    args += ["layersizes=" + ",".join(map(str, ls)), f"outwts_file={d}/{tag}.wts", f"log_file={d}/{tag}.log"] + extra
    subprocess.run([exe, f"{d}/{tag}.bin"] + args, cwd=d, stdout=subprocess.DEVNULL, check=True, timeout=60)
    LOGS[tag] = open(f"{d}/{tag}.log").read().replace(f"{d}/{tag}.", f"{d}/X.")
    return open(f"{d}/{tag}.bin", "rb").read(), open(f"{d}/{tag}.wts", "rb").read()


def _build():
    if not os.path.exists(OURS):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "host"), "-s"])


@pytest.mark.parametrize("seed,ls,rng", [(7, [1548, 5, 129], []), (12345, [1548, 33, 17, 129], []),
                                         (0, [1548, 64, 8, 3, 129], ["init_randem_weight_min=-0.5",
                                                                      "init_randem_weight_max=0.25",
                                                                      "init_randem_bias_min=0.0",
                                                                      "init_randem_bias_max=0.3"])])
def test_random_init_and_wts_bytes_equal_reference(seed, ls, rng):
    _build()
    with tempfile.TemporaryDirectory() as d:
        make_inputs(d, CASE)
        extra = [f"init_randem_seed={seed}"] + rng
        ref_bin, ref_wts = _run(REF, d, "ref", ls, extra)
        our_bin, our_wts = _run(OURS, d, "ours", ls, extra)
    assert len(ref_bin) == 4 * sum((ls[i - 1] + 1) * ls[i] for i in range(1, len(ls)))
    assert our_bin == ref_bin, "initial weights differ from the reference's"
    assert our_wts == ref_wts, ".wts bytes differ from the reference's"
    assert LOGS["ours"] == LOGS["ref"] and "Saving weights to file..." in LOGS["ours"], "log lines differ"


def test_wts_files_load_identically_on_both_sides():
    _build()
    T = importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")
    ls = [1548, 9, 4, 129]
    rng = np.random.default_rng(3)
    w = [None] + [rng.standard_normal((ls[i - 1], ls[i])).astype(np.float32) for i in range(1, len(ls))]
    b = [None] + [rng.standard_normal(ls[i]).astype(np.float32) for i in range(1, len(ls))]
    with tempfile.TemporaryDirectory() as d:
        make_inputs(d, CASE)
        T.write_wts(f"{d}/init.wts", w, b)                       # our Python writer
        extra = [f"initwts_file={d}/init.wts"]
        ref_bin, ref_wts = _run(REF, d, "ref", ls, extra)
        our_bin, our_wts = _run(OURS, d, "ours", ls, extra)
        assert our_bin == ref_bin and our_wts == ref_wts
        want = b"".join(w[i].tobytes() + b[i].tobytes() for i in range(1, len(ls)))
        assert ref_bin == want, "the reference loads our Python-written .wts to different arrays"
        assert ref_wts == open(f"{d}/init.wts", "rb").read(), "Python writer's bytes differ from the reference writer's"
        w2, b2 = T.read_wts(f"{d}/ref.wts", ls)                  # our Python reader on the reference's file
        for i in range(1, len(ls)):
            assert np.array_equal(w2[i], w[i]) and np.array_equal(b2[i], b[i])


def test_truncated_weight_and_norm_files_are_refused():
    """Two places where we are deliberately stricter than the reference (both still "message + exit(0)"): a truncated
    initwts_file is reported as such (the reference reads on into garbage and complains about bias node counts), and a
    norm file that ends early is an error (the reference carries on with whatever its buffers held)."""
    _build()
    T = importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")
    ls = [1548, 5, 129]
    rng = np.random.default_rng(0)
    w = [None] + [rng.standard_normal((ls[i - 1], ls[i])).astype(np.float32) for i in (1, 2)]
    b = [None] + [rng.standard_normal(ls[i]).astype(np.float32) for i in (1, 2)]
    with tempfile.TemporaryDirectory() as d:
        make_inputs(d, CASE)
        T.write_wts(f"{d}/good.wts", w, b)
        good = open(f"{d}/good.wts", "rb").read()
        open(f"{d}/short.wts", "wb").write(good[: len(good) // 2])
        norm = open(f"{d}/fea.norm").read().split("\n")
        open(f"{d}/short.norm", "w").write("\n".join(norm[:100]))
        for extra, msg in (([f"initwts_file={d}/short.wts"], "init weights file truncated"),
                           ([f"norm_file={d}/short.norm"], "norm file too short")):
            args = [a for a in reader_args(d, CASE) if not a.startswith(("layersizes=", "outwts_file=", "log_file="))]
            args += ["layersizes=" + ",".join(map(str, ls)), f"outwts_file={d}/o.wts", f"log_file={d}/o.log"] + extra
            p = subprocess.run([OURS, f"{d}/o.bin"] + args, cwd=d, capture_output=True, text=True, timeout=60)
            assert p.returncode == 0 and msg in open(f"{d}/o.log").read()
            assert os.path.getsize(f"{d}/o.bin") == 0      # stopped before any weights were produced
