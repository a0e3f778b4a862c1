"""GPU bring-up probe for the tcgen05 GEMM (run on the B200 box; numpy + ctypes only, no torch).

Prints, for every operand-major combination and a set of awkward shapes, the max error against a float64 product
of (a) the raw fp32 operands, (b) operands truncated to TF32, (c) operands rounded to TF32 — which tells us both
whether the descriptors/layouts are right and how the tensor core treats the low 13 mantissa bits.
Usage:  python tests/gpu_probe_gemm.py            (exit code 0 = all shapes within tolerance)
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "lib", "libbpgpu.so")


def tf32_trunc(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_round(x):
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def main():
    lib = C.CDLL(LIB)
    lib.bp_last_error.restype = C.c_char_p
    fp = C.POINTER(C.c_float)
    lib.bp_debug_gemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, fp, C.c_int, fp, C.c_int, fp, C.c_int, fp, fp,
                                  C.c_int, C.c_float, C.c_int, C.c_int, fp]
    rng = np.random.default_rng(0)
    shapes = [(128, 128, 32), (128, 128, 64), (128, 128, 256), (256, 256, 128), (257, 100, 96), (64, 36, 40),
              (2048, 1024, 2048), (2049, 257, 1024), (257, 1024, 2827), (300, 4100, 515)]
    ok_all = True
    for kind, name in [(3, "fwd  A=MN B=K "), (1, "dX   A=K  B=K "), (2, "dW   A=MN B=MN")]:
        for (M, N, K) in shapes:
            if kind in (3,):
                A = rng.standard_normal((K, M), dtype=np.float32)  # W[k][m]
                B = rng.standard_normal((N, K), dtype=np.float32)  # X[n][k]
                ref = lambda a, b: (b.astype(np.float64) @ a.astype(np.float64))  # [N][M]
            elif kind == 1:
                A = rng.standard_normal((M, K), dtype=np.float32)  # W[m][k]
                B = rng.standard_normal((N, K), dtype=np.float32)  # D[n][k]
                ref = lambda a, b: (b.astype(np.float64) @ a.astype(np.float64).T)
            else:
                A = rng.standard_normal((K, M), dtype=np.float32)  # D[k][m]
                B = rng.standard_normal((K, N), dtype=np.float32)  # X[k][n]
                ref = lambda a, b: (b.astype(np.float64).T @ a.astype(np.float64))
            out = np.full((N, M), np.nan, dtype=np.float32)
            aux = np.ones((N, M), dtype=np.float32)  # Y > 0 -> act' = 1 for kind 1
            ms = C.c_float(0)
            rc = lib.bp_debug_gemm(kind, M, N, K, A.ctypes.data_as(fp), A.shape[1], B.ctypes.data_as(fp), B.shape[1],
                                   out.ctypes.data_as(fp), M, None, aux.ctypes.data_as(fp) if kind == 1 else None, M,
                                   1.0, 0, 0, C.byref(ms))
            if rc != 0:
                print(f"{name} M={M} N={N} K={K}: rc={rc} {lib.bp_last_error().decode()}")
                ok_all = False
                if "cuda" in lib.bp_last_error().decode().lower():
                    return 2
                continue
            r_raw, r_tr, r_rn = ref(A, B), ref(tf32_trunc(A), tf32_trunc(B)), ref(tf32_round(A), tf32_round(B))
            scale = np.sqrt(K)
            e_raw = np.nanmax(np.abs(out - r_raw)) / scale
            e_tr = np.nanmax(np.abs(out - r_tr)) / scale
            e_rn = np.nanmax(np.abs(out - r_rn)) / scale
            nnan = int(np.isnan(out).sum())
            good = nnan == 0 and min(e_tr, e_rn) < 1e-4 and e_raw < 1e-2
            ok_all &= good
            gfl = 2.0 * M * N * K / (ms.value * 1e-3) / 1e12 if ms.value > 0 else 0
            print(f"{name} M={M:5d} N={N:5d} K={K:5d}: err/sqrtK raw={e_raw:.2e} trunc={e_tr:.2e} round={e_rn:.2e} "
                  f"nan={nnan} {ms.value:.3f} ms {gfl:.1f} TF/s {'OK' if good else 'FAIL'}")
            if not good:
                bad = np.argwhere(~(np.abs(out - r_tr) < 1e-2 * scale))
                print("   first bad (n,m):", bad[:6].tolist(), " count", len(bad), " out[0,:4]", out[0, :4],
                      " ref[0,:4]", r_tr[0, :4])
    # fused epilogue check (kind 0: bias + relu)
    M, N, K = 300, 200, 130
    A = rng.standard_normal((K, M), dtype=np.float32)
    B = rng.standard_normal((N, K), dtype=np.float32)
    bias = rng.standard_normal(M, dtype=np.float32)
    out = np.zeros((N, M), dtype=np.float32)
    rc = lib.bp_debug_gemm(0, M, N, K, A.ctypes.data_as(fp), M, B.ctypes.data_as(fp), K, out.ctypes.data_as(fp), M,
                           bias.ctypes.data_as(fp), None, 0, 1.0, 0, 0, None)
    r = np.maximum(tf32_trunc(B).astype(np.float64) @ tf32_trunc(A).astype(np.float64) + bias, 0)
    e = np.abs(out - r).max()
    print(f"fwd bias+relu rc={rc} err={e:.2e} {'OK' if rc == 0 and e < 1e-3 else 'FAIL'}")
    ok_all &= (rc == 0 and e < 1e-3)
    print("ALL OK" if ok_all else "SOME FAILED")
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
