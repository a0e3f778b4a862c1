#!/usr/bin/env python
"""Round-2 first call: the fused update (EPI_DW_SGD, bp_set_option("fused_update", 1)) against the separate update.

1. Parity, bit for bit: two trainers with identical initial state train the same bunches, one with the momentum-SGD
   update applied by the dW GEMM's epilogue, one with bp_sgd_kernel; weights, biases and the next forward must be
   IDENTICAL (the gradient values and the update arithmetic are the same, only where they are applied differs).
   Cases cover: every kernel choice (lone CTAs / 256-wide pairs), a bunch that is not a power of two (division path),
   weight cost (bias row exempt), momentum carried over several bunches, dropout, sigmoid, 3xTF32 (w_lo written).
2. Timing at C2 (2827-2048x3-257, bunch 1024): ms per bunch fused / unfused, prefetch on / off, in this one box.

Not a pytest test until it has passed on a B200 once (the round-end suite runs with -x).
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
bp = importlib.import_module("dnn-for-speech-enhancement_b200")
import oracle_py as O  # noqa: E402  (checker only: initial weights / synthetic frames)


def make(sizes, bunch, fused, **kw):
    w, b = O.glorot_init(sizes, seed=3)
    g = bp.BP_GPU(1, len(sizes), sizes, bunch, kw.pop("lrate", 1.0), kw.pop("momentum", 0.9),
                  kw.pop("weightcost", 0.0), w, b, kw.pop("dropoutflag", 0), kw.pop("visible_omit", 0.0),
                  kw.pop("hid_omit", 0.0), device=0, seed=777, **kw)
    g.set_option("fused_update", 1 if fused else 0)
    return g


def parity_case(name, sizes, bunch, n_bunches, **kw):
    x, t = O.synth_data(bunch * n_bunches, sizes[0], sizes[-1], seed=11)
    res = []
    for fused in (0, 1):
        g = make(sizes, bunch, fused, **dict(kw))
        g.train(bunch * n_bunches, x, t)
        ws, bs = g.returnWeights()
        out = g.forward(min(bunch, 64), x[: min(bunch, 64)])
        res.append((ws, bs, out))
        g.close()
    ok = True
    for l in range(1, len(sizes)):
        ok &= np.array_equal(res[0][0][l], res[1][0][l]) and np.array_equal(res[0][1][l], res[1][1][l])
    ok &= np.array_equal(res[0][2], res[1][2])
    moved = max(float(np.abs(res[0][0][l] - O.glorot_init(sizes, seed=3)[0][l]).max()) for l in range(1, len(sizes)))
    worst = max(float(np.abs(res[0][0][l] - res[1][0][l]).max()) for l in range(1, len(sizes)))
    print(f"[{'OK ' if ok else 'BAD'}] {name}: weights moved by up to {moved:.3e}, fused vs separate max |diff| = {worst:.3e}")
    return ok


def timing(sizes, bunch, steps=200):
    x, t = O.synth_data(bunch * 16, sizes[0], sizes[-1], seed=5)
    for label, fused, pre in (("separate update", 0, 1), ("fused update", 1, 1), ("fused, no L2 prefetch", 1, 0),
                              ("separate update", 0, 1), ("fused update", 1, 1)):
        g = make(sizes, bunch, fused)
        g.set_option("fused_prefetch", pre)
        g.upload_chunk(bunch * 16, x, t)
        for _ in range(3):
            g.train_resident(0, 16)
        g.sync()
        g.timer_start()
        done = 0
        while done < steps:
            g.train_resident(0, 16)
            done += 16
        ms = g.timer_stop()
        g.set_profiling(True)
        g.train_resident(0, 16)
        prof, n = g.profile()
        g.set_profiling(False)
        print(f"{label:24s} {ms / done:.4f} ms/bunch  {done * bunch / ms * 1e3 / 1e6:.3f} M frames/s  "
              + " ".join(f"{k}={v / max(n, 1):.4f}" for k, v in prof.items()))
        g.close()


def main():
    ok = True
    ok &= parity_case("lone CTAs, 1 layer", [96, 40], 64, 3)
    ok &= parity_case("small 3-layer, bunch 96 (division path), weight cost", [300, 260, 130, 33], 96, 4, weightcost=0.01)
    ok &= parity_case("pairs: 2827-2048-2048-257, bunch 1024", [2827, 2048, 2048, 257], 1024, 3)
    ok &= parity_case("C2 net, 4 bunches", [2827, 2048, 2048, 2048, 257], 1024, 4)
    ok &= parity_case("dropout 0.2/0.2, bunch 2048", [3084, 2048, 2048, 257], 2048, 2, dropoutflag=1, visible_omit=0.2,
                      hid_omit=0.2)
    ok &= parity_case("sigmoid + weight cost", [257, 512, 257], 128, 5, activation=bp.BP_ACT_SIGMOID, weightcost=1e-4)
    ok &= parity_case("3xTF32", [300, 260, 130, 33], 128, 3, math_mode=bp.BP_MATH_3XTF32)
    print("PARITY", "ALL OK" if ok else "FAILED")
    if ok and (len(sys.argv) < 2 or sys.argv[1] != "parity"):
        t0 = time.time()
        timing([2827, 2048, 2048, 2048, 257], 1024)
        print(f"timing took {time.time() - t0:.1f} s")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
