"""Hand-over protocol of the chunk prefetch thread (host/ChunkPrefetch.h) under random reader / consumer timing:
tests/native/prefetch_stress.cc checks order, slot alternation, that a slot is never written while the consumer holds
it, the end marker, and destruction with the reader blocked or mid-read — plainly and, where the toolchain has it,
under ThreadSanitizer."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "host")
EXE = os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "bin", "prefetch_stress")


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_prefetcher_protocol_under_random_timing(seed):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", HOST, "-s", "../bin/prefetch_stress"])
    p = subprocess.run([EXE, "400", str(seed)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr


def test_prefetcher_is_clean_under_thread_sanitizer():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "stress_tsan")
        c = subprocess.run(["g++", "-O1", "-g", "-fsanitize=thread", "-std=c++17", "-pthread", "-I", HOST, "-o", exe,
                            os.path.join(ROOT, "tests", "native", "prefetch_stress.cc")], capture_output=True, text=True)
        if c.returncode != 0:
            pytest.skip("ThreadSanitizer runtime not available: " + c.stderr[-200:])
        p = subprocess.run([exe, "300", "7"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "WARNING: ThreadSanitizer" not in p.stderr, p.stdout + p.stderr[-3000:]
