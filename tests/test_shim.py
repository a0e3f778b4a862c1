"""The drop-in boundary, literally: the reference's UNMODIFIED BPtrain.cc and Interface.cc linked with
host/BP_GPU_shim.cc (the binding of INTEGRATION.md §B, which implements the reference's class BP_GPU on libbpgpu's
C-ABI) instead of BP_GPU.cu / DevFunc.cu — built by oracle/build_ref.sh as oracle/_ref/BPtrain_shim.
  CPU: the binary exists (so the five members the reference's main() needs are all bound) and, without a GPU, fails
       the reference's way: message, exit(0);
  GPU: it trains the very same epoch as our own CLI — byte-identical .wts, identical CV line — on identical inputs
       (gated until it has run on a B200 once)."""
import importlib
import os
import subprocess
import tempfile

import pytest

from reader_case import CASES, make_inputs, reader_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "_ref", "BPtrain_shim")
OURS = os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "bin", "BPtrain")
needs_shim = pytest.mark.skipif(not os.path.exists(SHIM), reason="oracle/_ref/BPtrain_shim not built (needs /root/reference)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@needs_shim
def test_reference_main_links_against_our_trainer_and_fails_loudly_without_a_gpu():
    undefined = subprocess.run(["nm", "-D", "--undefined-only", SHIM], capture_output=True, text=True).stdout
    for sym in ("bp_create", "bp_train", "bp_crossvalid", "bp_return_weights", "bp_destroy", "bp_last_error"):
        assert sym in undefined, f"{sym} is not what the shim binds"
    assert "cublas" not in undefined.lower() and "curand" not in undefined.lower()
    if _has_gpu():
        pytest.skip("GPU present")
    case = CASES["129"]
    with tempfile.TemporaryDirectory() as d:
        make_inputs(d, case)
        p = subprocess.run([SHIM] + reader_args(d, case), cwd=d, capture_output=True, text=True, timeout=60)
    assert p.returncode == 0 and "no CUDA device" in p.stdout   # reference convention: print + exit(0)


@pytest.mark.gpu
@needs_shim
@pytest.mark.parametrize("dropout", [0, 1])
def test_reference_main_with_our_trainer_equals_our_cli(dropout):
    from test_cli_gpu import LS, _args, _cv
    T = importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")
    with tempfile.TemporaryDirectory() as d:
        feas, targs, mu, ivar = T.synth_corpus(40, 129, 129, seed=6, min_len=30, max_len=90)
        T.write_pfile(f"{d}/fea.pfile", feas)
        T.write_pfile(f"{d}/targ.pfile", targs)
        T.write_norm(f"{d}/fea.norm", mu, ivar)
        extra = ["traincache=700"] + (["dropoutflag=1", "visible_omit=0.1", "hid_omit=0.2"] if dropout else [])
        drop = ("traincache=3000",) + (("dropoutflag=0", "visible_omit=0", "hid_omit=0") if dropout else ())
        # Our CLI keys the dropout masks on seed= mixed with init_randem_seed (host/BPtrain.cc: epoch_dropout_seed); the
        # reference's main knows nothing of seeds, so the library takes BP_SEED from the environment: same key here.
        z = (0x5EED5EED + 0x9E3779B97F4A7C15 * 7) & (2 ** 64 - 1)          # seed default, init_randem_seed=7 (_args)
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
        env = dict(os.environ, BP_SEED=str(z ^ (z >> 31)))
        for tag, exe in (("shim", SHIM), ("ours", OURS)):
            a = [x for x in _args(d, tag, extra) if x not in drop]
            o = subprocess.run([exe] + a, cwd=d, capture_output=True, text=True, timeout=600,
                               env=env if tag == "shim" else None)
            assert o.returncode == 1, o.stdout + o.stderr          # success exit status is 1 (BPtrain.cc:100)
        assert open(f"{d}/shim.wts", "rb").read() == open(f"{d}/ours.wts", "rb").read()
        assert _cv(f"{d}/shim.log") == _cv(f"{d}/ours.log")
