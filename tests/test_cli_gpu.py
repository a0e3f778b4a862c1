"""GPU end-to-end drop-in test: our `BPtrain` CLI against the reference's own `BPtrain` (oracle/_ref/BPtrain_ref,
compiled from /root/reference by oracle/build_ref.sh) on identical synthetic Pfile / norm inputs and identical
command lines, at the reference's own operating shape (129-dim LPS, 11-frame context + NAT block = 1548 inputs)."""
import importlib
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "bin", "BPtrain")
REF = os.path.join(ROOT, "oracle", "_ref", "BPtrain_ref")
LS = [1548, 256, 192, 129]


def _args(d, tag, extra=()):
    return [f"fea_file={d}/fea.pfile", f"norm_file={d}/fea.norm", f"targ_file={d}/targ.pfile",
            f"outwts_file={d}/{tag}.wts", f"log_file={d}/{tag}.log", "initwts_file=", "train_sent_range=0-29",
            "cv_sent_range=30-39", "fea_dim=129", "fea_context=11", "targ_offset=5", "traincache=3000",
            "bunchsize=64", "layersizes=" + ",".join(map(str, LS)), "numlayers=4", "gpu_used=1",
            "init_randem_seed=7", "momentum=0.5", "weightcost=0", "lrate=1", "dropoutflag=0", "visible_omit=0",
            "hid_omit=0", "init_randem_weight_min=-0.05", "init_randem_weight_max=0.05",
            "init_randem_bias_min=-0.05", "init_randem_bias_max=0.05"] + list(extra)


def _cv(path):
    m = re.search(r"CV over\. squared error: ([-0-9.eE+naninf]+)", open(path).read())
    assert m, open(path).read()[-500:]
    return float(m.group(1))


@pytest.mark.skipif(not (os.path.exists(OURS) and os.path.exists(REF)), reason="BPtrain / BPtrain_ref not built")
def test_bptrain_cli_matches_reference_binary():
    T = importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")
    with tempfile.TemporaryDirectory() as d:
        feas, targs, mu, ivar = T.synth_corpus(40, 129, 129, seed=1, min_len=40, max_len=120)
        T.write_pfile(f"{d}/fea.pfile", feas)
        T.write_pfile(f"{d}/targ.pfile", targs)
        T.write_norm(f"{d}/fea.norm", mu, ivar)
        r = subprocess.run([REF] + _args(d, "ref"), cwd=d, capture_output=True, text=True, timeout=600)
        assert os.path.exists(f"{d}/ref.wts") and os.path.getsize(f"{d}/ref.wts") > 0, r.stdout + r.stderr
        o = subprocess.run([OURS] + _args(d, "ours"), cwd=d, capture_output=True, text=True, timeout=600)
        assert o.returncode == 1, o.stdout + o.stderr          # success exit status is 1 (BPtrain.cc:100)
        rw, rb = T.read_wts(f"{d}/ref.wts", LS)
        ow, ob = T.read_wts(f"{d}/ours.wts", LS)
        cv_ref, cv_ours = _cv(f"{d}/ref.log"), _cv(f"{d}/ours.log")
        log = open(f"{d}/ours.log").read()
    for line in ("parameters input:", "Get pfile info over:", "Get chunk info over:", "Starting chunk 1 of",
                 "Saving weights to file...", "Starting CV.", "Total cost time:"):
        assert line in log
    fro = lambda a: float(np.linalg.norm(np.asarray(a, np.float64).ravel()))
    for l in range(1, len(LS)):
        assert fro(ow[l] - rw[l]) <= 1e-2 * fro(rw[l]), f"W{l}: {fro(ow[l] - rw[l]) / fro(rw[l]):.3e}"
        assert fro(ob[l] - rb[l]) <= 2e-2 * fro(rb[l]) + 1e-6, f"b{l}"
    assert abs(cv_ours - cv_ref) <= 1e-2 * abs(cv_ref), (cv_ours, cv_ref)


@pytest.mark.skipif(not os.path.exists(OURS), reason="BPtrain not built")
@pytest.mark.parametrize("gpus", [1])
def test_bptrain_device_reader_equals_host_reader(gpus):
    """reader=gpu (records spliced on the device) must give the same epoch as reader=host: identical
    .wts bytes, identical CV score and identical decode output, because the assembled rows are bit-identical and the
    shuffle consumes the same lrand48 stream."""
    T = importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")
    with tempfile.TemporaryDirectory() as d:
        feas, targs, mu, ivar = T.synth_corpus(40, 129, 129, seed=2, min_len=30, max_len=90)
        T.write_pfile(f"{d}/fea.pfile", feas)
        T.write_pfile(f"{d}/targ.pfile", targs)
        T.write_norm(f"{d}/fea.norm", mu, ivar)
        for tag in ("host", "gpu"):
            extra = [f"reader={tag}", f"decode_file={d}/{tag}.dec", "traincache=1000"]
            o = subprocess.run([OURS] + [a for a in _args(d, tag, extra) if not a.startswith("traincache=3000")],
                               cwd=d, capture_output=True, text=True, timeout=600)
            assert o.returncode == 1, o.stdout + o.stderr
        assert open(f"{d}/host.wts", "rb").read() == open(f"{d}/gpu.wts", "rb").read()
        assert _cv(f"{d}/host.log") == _cv(f"{d}/gpu.log")
        assert open(f"{d}/host.dec", "rb").read() == open(f"{d}/gpu.dec", "rb").read()
        assert os.path.getsize(f"{d}/gpu.dec") > 0


@pytest.mark.skipif(not os.path.exists(OURS), reason="BPtrain not built")
def test_bptrain_decode_pfile_lines_up_with_targets():
    """decode_format=pfile: the enhanced frames come out as a Pfile whose (sentence, frame) ids are
    those of the clean-target frames they estimate, holding the same values as the raw decode output."""
    T = importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")
    with tempfile.TemporaryDirectory() as d:
        feas, targs, mu, ivar = T.synth_corpus(40, 129, 129, seed=3, min_len=8, max_len=60)
        T.write_pfile(f"{d}/fea.pfile", feas)
        T.write_pfile(f"{d}/targ.pfile", targs)
        T.write_norm(f"{d}/fea.norm", mu, ivar)
        for tag, extra in (("raw", [f"decode_file={d}/raw.dec", "reader=gpu"]),
                           ("pf", [f"decode_file={d}/pf.dec", "decode_format=pfile", "reader=host"])):
            o = subprocess.run([OURS] + _args(d, tag, extra), cwd=d, capture_output=True, text=True, timeout=600)
            assert o.returncode == 1, o.stdout + o.stderr
        raw = np.fromfile(f"{d}/raw.dec", dtype="<f4").reshape(-1, 129)
        sents, sid, fid = T.read_pfile(f"{d}/pf.dec")
    assert np.array_equal(np.vstack([s for s in sents if len(s)]), raw)
    # CV range is sentences 30-39 (see _args): every sentence of T >= 11 frames yields frames 5 .. T-6
    want_sid, want_fid = [], []
    for k, s in enumerate(range(30, 40)):
        n = feas[s].shape[0]
        for j in range(max(0, n - 10)):
            want_sid.append(k)
            want_fid.append(j + 5)
    assert sid.tolist() == want_sid and fid.tolist() == want_fid


@pytest.mark.skipif(not os.path.exists(OURS), reason="BPtrain not built")
@pytest.mark.parametrize("reader", ["host", "gpu"])
def test_bptrain_in_process_epochs_equal_chained_invocations(reader):
    """epochs=3 (the Perl driver's loop folded into one process) must write, for every epoch, the very
    bytes that three chained single-epoch invocations write when they follow the Perl schedule
    (finetune_DNN_speech_enhancement_dropout_NAT.pl:131-165: initwts = previous epoch's file, momentum += 0.04,
    init_randem_seed += 345), dropout on, so that masks, shuffles, momentum reset and hyper-parameters are all covered."""
    T = importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")
    drop = ["dropoutflag=1", "visible_omit=0.1", "hid_omit=0.2", f"reader={reader}", "traincache=1500"]

    def args(d, out, log, init, extra):
        a = [x for x in _args(d, "x", drop + list(extra))
             if not x.startswith(("outwts_file=", "log_file=", "initwts_file="))
             and x not in ("traincache=3000", "dropoutflag=0", "visible_omit=0", "hid_omit=0")]
        return a + [f"outwts_file={out}", f"log_file={log}", f"initwts_file={init}"]

    with tempfile.TemporaryDirectory() as d:
        feas, targs, mu, ivar = T.synth_corpus(40, 129, 129, seed=4, min_len=30, max_len=90)
        T.write_pfile(f"{d}/fea.pfile", feas)
        T.write_pfile(f"{d}/targ.pfile", targs)
        T.write_norm(f"{d}/fea.norm", mu, ivar)
        # (a) chained processes, as the Perl script drives them
        momentum, seed, init = 0.5, 7, ""
        for e in (1, 2, 3):
            if e > 1:
                momentum, seed, init = momentum + 0.04, seed + 345, f"{d}/chain.{e - 1}.wts"
            # Perl interpolates numbers with %.15g
            extra = [f"momentum={momentum:.15g}", f"init_randem_seed={seed}"]
            a = args(d, f"{d}/chain.{e}.wts", f"{d}/chain.{e}.log", init, extra)  # later keys win (argv order)
            o = subprocess.run([OURS] + a, cwd=d, capture_output=True, text=True, timeout=600)
            assert o.returncode == 1, o.stdout + o.stderr
        # (b) one process
        o = subprocess.run([OURS] + args(d, f"{d}/one.%d.wts", f"{d}/one.%d.log", "", ["epochs=3"]), cwd=d,
                           capture_output=True, text=True, timeout=600)
        assert o.returncode == 1, o.stdout + o.stderr
        for e in (1, 2, 3):
            assert open(f"{d}/one.{e}.wts", "rb").read() == open(f"{d}/chain.{e}.wts", "rb").read(), f"epoch {e}"
            assert _cv(f"{d}/one.{e}.log") == _cv(f"{d}/chain.{e}.log")
            assert "Total cost time:" in open(f"{d}/one.{e}.log").read()
        assert "momentum:		                0.580000" in open(f"{d}/one.3.log").read()
