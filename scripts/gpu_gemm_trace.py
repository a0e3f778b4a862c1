"""Prints CTA 0's pipeline timeline (clock64 per k-block) for one hidden-layer GEMM in several modes."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "lib", "libbpgpu.so"))
lib.bp_last_error.restype = C.c_char_p
fp = C.POINTER(C.c_float)
lib.bp_debug_gemm.argtypes = [C.c_int]*4 + [fp, C.c_int, fp, C.c_int, fp, C.c_int, fp, fp, C.c_int, C.c_float, C.c_int, C.c_int, fp]
rng = np.random.default_rng(0)
M, N, K = 2048, 1024, 2048
A = rng.standard_normal((K, M), dtype=np.float32); B = rng.standard_normal((N, K), dtype=np.float32)
out = np.zeros((N, M), np.float32)
os.environ["BP_DBG_TRACE"] = "1"
os.environ["BP_DBG_REPS"] = "5"
for flags, name in [(0, "full"), (1, "TMA only"), (2, "MMA only")]:
    os.environ["BP_DBG_FLAGS"] = str(flags)
    ms = C.c_float(0)
    print(f"==== {name} (fwd A=MN B=K, {M}x{N}x{K})", flush=True)
    rc = lib.bp_debug_gemm(3, M, N, K, A.ctypes.data_as(fp), M, B.ctypes.data_as(fp), K, out.ctypes.data_as(fp), M, None, None, M, 1.0, 0, 0, C.byref(ms))
    if rc: print(lib.bp_last_error().decode()); sys.exit(1)
    print(f"   {ms.value*1e3:.1f} us per launch", flush=True)
