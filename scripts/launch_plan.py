#!/usr/bin/env python
"""Launch plan of one training bunch / decode batch per BASELINE config, from the library's own selection rule
(bp_debug_plan, host-only: runs without a GPU).  For every product: M x N x K, the kernel picked, output tiles, CTAs of
the persistent launch on a 148-SM B200, 64-deep k-blocks per tile, waves (tiles per CTA or CTA pair), algorithmic GFLOP.
   python scripts/launch_plan.py > profiles/r1_launch_plan.md"""
import ctypes as C
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

bp = importlib.import_module("dnn-for-speech-enhancement_b200")
lib = bp.load_library()
lib.bp_debug_plan.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_int)] * 4
SMS = 148


def plan(M, N, K, b64):
    v = [C.c_int(0) for _ in range(4)]
    assert lib.bp_debug_plan(M, N, K, int(b64), SMS, 0, *[C.byref(x) for x in v]) == 0
    return [x.value for x in v]


def products(sizes, frames, train):
    L = len(sizes) - 1
    for l in range(1, L + 1):
        yield f"fwd {l}" + (" (output)" if l == L else ""), sizes[l], frames, sizes[l - 1], True
    if train:
        for l in range(L, 0, -1):
            yield f"dW {l}", sizes[l], sizes[l - 1] + 1, frames, False
            if l >= 2:
                yield f"dX {l} -> {l - 1}", sizes[l - 1], frames, sizes[l], True


def main():
    print("# Launch plan per BASELINE config (148 SMs; `python scripts/launch_plan.py`, host-only)\n")
    print("Kernel: `lone` = 128 x 128 tiles on single CTAs (`bp_gemm_kernel`), `pair128` / `pair256` = 256 x 128 / "
          "256 x 256 tiles on CTA pairs (`bp_gemm2_kernel`, `cta_group::2`).  Every launch is persistent: CTAs = "
          "min(tiles, SMs) or 2 x min(pair tiles, 74); waves = tiles per CTA (pair), rounded up.  The 257-wide output "
          "layer's forward product is additionally cut into K slices (split-K) so that it covers the machine.\n")
    for name in ("C2", "C3", "C4", "C5"):
        sizes, frames, _d, _v, _h, train = bench.WORKLOADS[name]
        print(f"## {name}: {'-'.join(map(str, sizes))}, {frames} frames per GPU, {'train' if train else 'decode'}\n")
        print("| product | M (units) x N x K | kernel | tiles | CTAs | k-blocks / tile | waves | GFLOP |")
        print("|---|---|---|---|---|---|---|---|")
        tot = 0.0
        for label, M, N, K, b64 in products(sizes, frames, train):
            pair_n, tiles, ctas, kb = plan(M, N, K, b64)
            units = ctas // 2 if pair_n else ctas
            waves = -(-tiles // units)
            gf = 2.0 * M * N * K / 1e9
            tot += gf
            kern = "lone" if pair_n == 0 else f"pair{pair_n}"
            print(f"| {label} | {M} x {N} x {K} | {kern} | {tiles} | {ctas} | {kb} | {waves} | {gf:.2f} |")
        print(f"\nSum {tot:.1f} GFLOP per {'bunch' if train else 'batch'} "
              f"(SURVEY.md §8d: {bench.flops_per_frame(sizes, train) * frames / 1e9:.1f}).\n")


if __name__ == "__main__":
    main()
