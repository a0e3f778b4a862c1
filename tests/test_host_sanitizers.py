"""The host data/IO code (Interface.cc: argv, Pfile parsing, chunk planner, all-core assembler, raw-chunk planner, the
prefetch thread) under AddressSanitizer + UndefinedBehaviorSanitizer on random corpora — the reference has no
sanitizer coverage at all and several latent out-of-bounds reads (App. D).  20 corpora x 4 harnesses
were run clean when this was written; a small subset runs here.  Skipped where the sanitizer runtimes are missing."""
import os
import shutil
import subprocess
import tempfile

import pytest

from reader_case import make_inputs, reader_args
from test_reader_fuzz import random_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "host")


@pytest.fixture(scope="module")
def san_bins():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    d = tempfile.mkdtemp(prefix="bp_asan_")
    bins = {}
    for h in ("reader_dump", "raw_dump", "prefetch_dump"):
        exe = os.path.join(d, h)
        c = subprocess.run(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-std=c++17",
                            "-pthread", "-I", HOST, "-o", exe, os.path.join(ROOT, "tests", "native", h + ".cc"),
                            os.path.join(HOST, "Interface.cc")], capture_output=True, text=True)
        if c.returncode != 0:
            shutil.rmtree(d, ignore_errors=True)
            pytest.skip("sanitizer runtimes not available: " + c.stderr[-200:])
        bins[h] = exe
    yield bins
    shutil.rmtree(d, ignore_errors=True)


@pytest.mark.parametrize("seed", [100, 107, 203, 47])
def test_host_reader_is_clean_under_asan_and_ubsan(san_bins, seed):
    case = random_case(seed)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1")
    with tempfile.TemporaryDirectory() as d:
        make_inputs(d, case)
        for h, exe in san_bins.items():
            p = subprocess.run([exe, f"{d}/{h}.bin"] + reader_args(d, case), cwd=d, capture_output=True, text=True,
                               timeout=300, env=env)
            assert p.returncode == 0 and "ERROR: AddressSanitizer" not in p.stderr and "runtime error" not in p.stderr, \
                f"{h}, seed {seed}:\n{p.stderr[-3000:]}"
