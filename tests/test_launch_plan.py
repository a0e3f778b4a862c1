"""Kernel choice per product, on the CPU (bp_debug_plan is host-only): the plans of the BASELINE configs on a 148-SM
B200 are the ones DESIGN.md section 3 / 5 describe and the r1d measurements were taken with.  A change of the selection
rule (pick_kernel, csrc/bp_launch.cu) has to show up here."""
import ctypes as C

import pytest

SMS = 148


def plan(bp, M, N, K, b64):
    lib = bp.load_library()
    lib.bp_debug_plan.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_int)] * 4
    v = [C.c_int(0) for _ in range(4)]
    assert lib.bp_debug_plan(M, N, K, int(b64), SMS, 0, *[C.byref(x) for x in v]) == 0
    return tuple(x.value for x in v)   # pair_n, tiles, ctas, k_blocks


@pytest.mark.parametrize("name,M,N,K,b64,want", [
    # C2: 2827 -> 2048 x 3 -> 257, bunch 1024
    ("C2 fwd layer 1", 2048, 1024, 2827, True, (128, 64, 128, 45)),
    ("C2 fwd hidden", 2048, 1024, 2048, True, (128, 64, 128, 32)),
    ("C2 fwd output (split-K lone CTAs)", 257, 1024, 2048, True, (0, 24, 24, 32)),
    ("C2 dX hidden", 2048, 1024, 2048, True, (128, 64, 128, 32)),
    ("C2 dX from the output layer", 2048, 1024, 257, True, (128, 64, 128, 5)),
    ("C2 dW hidden (+ bias column)", 2048, 2049, 1024, False, (256, 72, 144, 16)),
    ("C2 dW layer 1", 2048, 2828, 1024, False, (256, 96, 148, 16)),
    ("C2 dW output", 257, 2049, 1024, False, (0, 51, 51, 16)),
    # C3: bunch 2048 -> 256-wide pairs everywhere they fit
    ("C3 fwd hidden", 2048, 2048, 2048, True, (256, 64, 128, 32)),
    ("C3 dW layer 1", 2048, 3085, 2048, False, (256, 104, 148, 32)),
    # C5: decode, 8192 frames per batch
    ("C5 fwd hidden", 2048, 8192, 2048, True, (256, 256, 148, 32)),
    ("C5 fwd output", 257, 8192, 2048, True, (256, 64, 128, 32)),
    # C4: 512 frames per GPU -> too few pair tiles, lone CTAs (small bunches)
    ("C4 fwd hidden, 512 local frames", 2048, 512, 2048, True, (0, 64, 64, 32)),
    # the reference script's own bunch 128
    ("bunch 128 fwd hidden", 2048, 128, 2048, True, (0, 16, 16, 32)),
])
def test_plans_of_the_baseline_configs(bp, name, M, N, K, b64, want):
    assert plan(bp, M, N, K, b64) == want, name


def test_plan_rejects_bad_arguments(bp):
    lib = bp.load_library()
    assert lib.bp_debug_plan(0, 1, 1, 0, SMS, 0, None, None, None, None) != 0
