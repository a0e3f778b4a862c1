"""Parity + timing of one GEMM kernel choice (BP_PAIRS / BP_MC from the environment) on the products of the bench
workloads and on ragged shapes whose cluster tiles hang over every edge.  numpy + ctypes only.
    BP_PAIRS=3 BP_MC=2 python scripts/gpu_mc_probe.py [quick]
Reference: float32 BLAS product of TF32-truncated operands (the tensor core truncates)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "lib", "libbpgpu.so")


def tr(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def main():
    os.environ.setdefault("BP_DBG_REPS", "20")
    lib = C.CDLL(LIB)
    lib.bp_last_error.restype = C.c_char_p
    fp = C.POINTER(C.c_float)
    lib.bp_debug_gemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, fp, C.c_int, fp, C.c_int, fp, C.c_int, fp, fp,
                                  C.c_int, C.c_float, C.c_int, C.c_int, fp]
    rng = np.random.default_rng(0)
    cases = [(3, "fwd", 300, 700, 515), (1, "dX ", 300, 700, 515), (2, "dW ", 300, 700, 515),
             (3, "fwd", 2048, 1024, 2048), (3, "fwd", 2048, 1024, 2827), (1, "dX ", 2048, 1024, 2048),
             (1, "dX ", 2048, 1024, 257), (2, "dW ", 2048, 2049, 1024), (2, "dW ", 2048, 2828, 1024),
             (3, "fwd", 2048, 2048, 2048), (3, "fwd", 2048, 8192, 2048), (2, "dW ", 2048, 3085, 2048)]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        cases = cases[:9]
    if len(sys.argv) > 1 and sys.argv[1] == "dxgap":
        # which half of dX costs the extra ~6.5 us per launch: its operand majors (kind 4 = K-major A, plain epilogue)
        # or its epilogue (kind 5 = forward's MN-major A with the dX epilogue)?
        cases = [(k, n, M, N, K) for (M, N, K) in ((300, 700, 515), (2048, 1024, 257), (2048, 1024, 2048))
                 for (k, n) in ((3, "fwd        "), (1, "dX         "), (4, "dX-A, plain"), (5, "fwd-A, dXep"),
                                (6, "dX, bitmask"))]
    tag = f"PAIRS={os.environ.get('BP_PAIRS', '-')} MC={os.environ.get('BP_MC', '-')}"
    ok_all = True
    for kind, name, M, N, K in cases:
        if kind in (3, 5):
            A = rng.standard_normal((K, M), dtype=np.float32); B = rng.standard_normal((N, K), dtype=np.float32)
            ref = tr(B) @ tr(A)
        elif kind in (1, 4, 6):
            A = rng.standard_normal((M, K), dtype=np.float32); B = rng.standard_normal((N, K), dtype=np.float32)
            ref = tr(B) @ tr(A).T
        else:
            A = rng.standard_normal((K, M), dtype=np.float32); B = rng.standard_normal((K, N), dtype=np.float32)
            ref = tr(B).T @ tr(A)
        out = np.full((N, M), np.nan, dtype=np.float32)
        aux = np.ones((N, M), dtype=np.float32)
        ms = C.c_float(0)
        rc = lib.bp_debug_gemm(kind, M, N, K, A.ctypes.data_as(fp), A.shape[1], B.ctypes.data_as(fp), B.shape[1],
                               out.ctypes.data_as(fp), M, None, aux.ctypes.data_as(fp) if kind in (1, 5, 6) else None, M,
                               1.0, 0, 0, C.byref(ms))
        if rc != 0:
            print(f"[{tag}] {name} {M}x{N}x{K}: rc={rc} {lib.bp_last_error().decode()}", flush=True)
            return 2
        err = float(np.nanmax(np.abs(out - ref))) / np.sqrt(K)
        nnan = int(np.isnan(out).sum())
        good = nnan == 0 and err < 1e-4
        ok_all &= good
        tf = 2.0 * M * N * K / (ms.value * 1e-3) / 1e12 if ms.value > 0 else 0.0
        print(f"[{tag}] {name} M={M:5d} N={N:5d} K={K:5d}: err/sqrtK={err:.2e} nan={nnan} {ms.value * 1e3:8.1f} us "
              f"{tf:6.1f} TF/s {'OK' if good else 'FAIL'}", flush=True)
        if not good and os.environ.get("MC_DIAG") and M <= 512:
            # coarse error map: one character per 32 x 32 block of the output (rows = N blocks, columns = M blocks)
            bad = ~(np.abs(out - ref) < 1e-3 * np.sqrt(K))
            for nb in range(0, N, 32):
                row = ""
                for mb in range(0, M, 32):
                    blk, o = bad[nb:nb + 32, mb:mb + 32], out[nb:nb + 32, mb:mb + 32]
                    row += "N" if np.isnan(o).all() else ("n" if np.isnan(o).any() else ("x" if blk.all() else ("+" if blk.any() else ".")))
                print(f"   n={nb:4d} {row}", flush=True)
    print(f"[{tag}]", "ALL OK" if ok_all else "SOME FAILED", flush=True)
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
