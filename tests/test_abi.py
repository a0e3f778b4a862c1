"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol include/bp_gpu.h
declares, fails loudly (no CPU fallback) when asked to compute, and its host-side mask function equals the oracle's."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(bp):
    if not os.path.exists(bp.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "csrc"), "-s"])
    lib = bp.load_library()
    names = bp.declared_symbols()
    assert {"bp_create", "bp_create_ex", "bp_train", "bp_crossvalid", "bp_forward", "bp_return_weights",
            "bp_destroy", "bp_last_error", "bp_upload_chunk", "bp_train_resident", "bp_comm_init"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bp_gpu.h but not exported"


def test_built_for_sm100a_with_tcgen05_and_tma(bp):
    """The shipped cubin is sm_100a and contains the Blackwell tensor / TMA / TMEM instructions (SASS names)."""
    lib = bp.LIB_PATH
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    sass = out.stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "SYNCS"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS"


def test_no_cpu_fallback(bp):
    """Without a GPU the product must refuse, not compute on the host."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except Exception:
        pass
    w = [None, np.zeros((4, 3), np.float32)]
    b = [None, np.zeros(3, np.float32)]
    with pytest.raises(bp.BpError) as e:
        bp.BP_GPU(1, 2, [4, 3], 8, 1.0, 0.0, 0.0, w, b, device=0)
    assert "no CUDA device" in str(e.value) or "rc=-2" in str(e.value) or "rc=-3" in str(e.value)


def test_bad_arguments_are_rejected(bp):
    lib = bp.load_library()
    h = C.c_void_p()
    assert lib.bp_create_ex(C.byref(h), None, None, None) != 0
    assert b"null" in lib.bp_last_error()
    assert lib.bp_train(None, 1, None, None) != 0
    assert lib.bp_set_option(None, b"chain", 1) != 0
    assert lib.bp_debug_gemm(9, 1, 1, 1, None, 1, None, 1, None, 1, None, None, 1, 1.0, 0, 0, None) != 0


def test_dropout_mask_host_function_matches_oracle(bp, oracle):
    rng = np.random.default_rng(0)
    for _ in range(2000):
        seed = int(rng.integers(0, 2**63))
        step, layer, f, u = (int(v) for v in rng.integers(0, 5000, size=4))
        p = float(rng.uniform(0, 1))
        assert bp.dropout_mask(seed, step, layer, f, u, p) == oracle.dropout_mask(seed, step, layer, f, u, p)


def test_cli_fails_loudly_without_a_gpu(tmp_path):
    """On a box without a CUDA device the BPtrain CLI must not compute anything on the CPU: the reference's error
    convention (message into the log, exit status 0 — e.g. BP_GPU.cu:20-24) with the library's ENODEV message."""
    import importlib
    import subprocess
    try:
        import ctypes
        if ctypes.CDLL("libcuda.so.1").cuInit(0) == 0:
            pytest.skip("a CUDA device is present")
    except OSError:
        pass
    exe = os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "bin", "BPtrain")
    if not os.path.exists(exe):
        pytest.skip("BPtrain not built")
    T = importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")
    d = str(tmp_path)
    feas, targs, mu, ivar = T.synth_corpus(6, 129, 129, seed=1, min_len=20, max_len=40)
    T.write_pfile(f"{d}/fea.pfile", feas)
    T.write_pfile(f"{d}/targ.pfile", targs)
    T.write_norm(f"{d}/fea.norm", mu, ivar)
    args = [f"fea_file={d}/fea.pfile", f"norm_file={d}/fea.norm", f"targ_file={d}/targ.pfile", f"outwts_file={d}/o.wts",
            f"log_file={d}/o.log", "initwts_file=", "train_sent_range=0-3", "cv_sent_range=4-5", "fea_dim=129",
            "fea_context=11", "targ_offset=5", "traincache=500", "bunchsize=32", "layersizes=1548,64,129",
            "gpu_used=1", "init_randem_seed=3", "momentum=0.5", "weightcost=0", "lrate=1", "dropoutflag=0",
            "visible_omit=0", "hid_omit=0"]
    r = subprocess.run([exe] + args, cwd=d, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0                                   # every error path is exit(0), success is 1
    log = open(f"{d}/o.log").read()
    assert "GPU trainer creation failed" in log and "no CPU fallback" in log
    assert os.path.getsize(f"{d}/o.wts") == 0                  # nothing was trained, nothing was written


def test_product_path_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import, load, link or exec it (comments aside)."""
    import re
    pkg = os.path.join(ROOT, "dnn-for-speech-enhancement_b200")
    pat = re.compile(r"libbporacle|oracle_py|bp_oracle|oracle/|import\s+oracle|from\s+oracle")
    hits = []
    for d, _dirs, files in os.walk(pkg):
        if os.sep + "lib" in d or os.sep + "bin" in d or "__pycache__" in d:
            continue
        for f in files:
            if f.endswith((".py", ".cc", ".h", ".cu", ".cuh")) or f == "Makefile":
                for i, line in enumerate(open(os.path.join(d, f), errors="ignore"), 1):
                    if pat.search(line):
                        hits.append(f"{f}:{i}: {line.strip()}")
    assert not hits, hits
    out = subprocess.run(["ldd", bp_lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out


def bp_lib_path():
    import importlib
    return importlib.import_module("dnn-for-speech-enhancement_b200").LIB_PATH
