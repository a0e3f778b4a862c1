"""clock64 timeline of the leader CTA of pair 0 (traced instantiation of bp_gemm2_kernel) for the forward and the dX
product of a hidden layer — same FLOPs, 16.7 vs 24.6 us alone in round 1d — to see where dX loses its time: slot-free ->
loads issued -> stage full -> MMAs issued per k-block, accumulator ready, epilogue done.
    python scripts/gpu_pair_trace.py            (default kernel choice: 128-wide pairs at this shape)"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "lib", "libbpgpu.so"))
lib.bp_last_error.restype = C.c_char_p
fp = C.POINTER(C.c_float)
lib.bp_debug_gemm.argtypes = [C.c_int]*4 + [fp, C.c_int, fp, C.c_int, fp, C.c_int, fp, fp, C.c_int, C.c_float, C.c_int, C.c_int, fp]
rng = np.random.default_rng(0)
M, N, K = 2048, 1024, 2048
os.environ["BP_DBG_TRACE"] = "1"
os.environ["BP_DBG_REPS"] = "50"     # the timeline printed is that of the last launch (clocks and L2 warm)
for kind, name in [(3, "fwd A=MN B=K"), (1, "dX  A=K  B=K"), (2, "dW  A=MN B=MN")]:
    if kind == 3: A = rng.standard_normal((K, M), dtype=np.float32); B = rng.standard_normal((N, K), dtype=np.float32)
    elif kind == 1: A = rng.standard_normal((M, K), dtype=np.float32); B = rng.standard_normal((N, K), dtype=np.float32)
    else: A = rng.standard_normal((K, M), dtype=np.float32); B = rng.standard_normal((K, N), dtype=np.float32)
    out = np.zeros((N, M), np.float32); aux = np.ones((N, M), np.float32)
    ms = C.c_float(0)
    print(f"==== {name} ({M}x{N}x{K})", flush=True)
    rc = lib.bp_debug_gemm(kind, M, N, K, A.ctypes.data_as(fp), A.shape[1], B.ctypes.data_as(fp), B.shape[1],
                           out.ctypes.data_as(fp), M, None, aux.ctypes.data_as(fp) if kind == 1 else None, M, 1.0, 0, 0,
                           C.byref(ms))
    if rc: print(lib.bp_last_error().decode()); sys.exit(1)
    print(f"   {ms.value*1e3:.1f} us per launch (traced instantiation)", flush=True)
