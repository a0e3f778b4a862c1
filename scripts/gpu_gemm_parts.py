"""Isolates the GEMM pipeline's components on the B200: full kernel vs TMA-only (no MMA) vs MMA-only (no loads),
steady-state (20 launches), for the hidden-layer shapes.  BP_CLUSTER must be 1x1 for the TMA-only variant."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "lib", "libbpgpu.so"))
lib.bp_last_error.restype = C.c_char_p
fp = C.POINTER(C.c_float)
lib.bp_debug_gemm.argtypes = [C.c_int]*4 + [fp, C.c_int, fp, C.c_int, fp, C.c_int, fp, fp, C.c_int, C.c_float, C.c_int, C.c_int, fp]
rng = np.random.default_rng(0)
os.environ["BP_DBG_REPS"] = "20"
for (kind, name) in [(3, "fwd"), (1, "dX"), (2, "dW")]:
    for (M, N, K) in [(2048, 1024, 2048), (257, 1024, 2048), (2048, 2049, 1024), (2048, 8192, 2048)]:
        if kind == 3: A = rng.standard_normal((K, M), dtype=np.float32); B = rng.standard_normal((N, K), dtype=np.float32)
        elif kind == 1: A = rng.standard_normal((M, K), dtype=np.float32); B = rng.standard_normal((N, K), dtype=np.float32)
        else: A = rng.standard_normal((K, M), dtype=np.float32); B = rng.standard_normal((K, N), dtype=np.float32)
        out = np.zeros((N, M), np.float32); aux = np.ones((N, M), np.float32)
        res = []
        for flags in (0, 1, 2):
            os.environ["BP_DBG_FLAGS"] = str(flags)
            ms = C.c_float(0)
            rc = lib.bp_debug_gemm(kind, M, N, K, A.ctypes.data_as(fp), A.shape[1], B.ctypes.data_as(fp), B.shape[1], out.ctypes.data_as(fp), M, None, aux.ctypes.data_as(fp) if kind == 1 else None, M, 1.0, 0, 0, C.byref(ms))
            if rc: print(lib.bp_last_error().decode()); sys.exit(1)
            res.append(ms.value * 1e3)
        tf = 2.0 * M * N * K / (res[0] * 1e-6) / 1e12
        print(f"{name} M={M} N={N} K={K}: full {res[0]:.1f} us ({tf:.0f} TF/s) | TMA-only {res[1]:.1f} us | MMA-only {res[2]:.1f} us")
