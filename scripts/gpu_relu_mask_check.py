#!/usr/bin/env python
"""ReLU bit mask (EPI_FWD_HID_MASK / EPI_DX_MASK, bp_set_option("relu_mask", 1)) against the default dX epilogue.

1. Parity, bit for bit: two trainers with identical initial state train the same bunches, one whose dX epilogues read
   Y (default), one whose forward epilogues leave a bit mask of Y > 0 that the dX epilogues read instead.  Weights,
   biases and the next forward must be IDENTICAL (y > 0 ? e : 0 either way).  Cases: lone CTAs and both pair widths,
   unit counts and bunches that are not multiples of 32 / 128 (ragged mask words), dropout (the mask must see the
   values as stored, after the dropout zeros), 3xTF32, the fused update on top, and a sigmoid net (switch ignored).
2. Timing at C2 and C3: ms per bunch and per-class ms, mask off / on alternating in this one box; then the isolated
   dX product (bp_debug_gemm kind 1) for reference.

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


def make(sizes, bunch, mask, fused=0, **kw):
    w, b = O.glorot_init(sizes, seed=3)
    g = bp.BP_GPU(1, len(sizes), sizes, bunch, kw.pop("lrate", 1.0), kw.pop("momentum", 0.9),
                  kw.pop("weightcost", 0.0), w, b, kw.pop("dropoutflag", 0), kw.pop("visible_omit", 0.0),
                  kw.pop("hid_omit", 0.0), device=0, seed=777, **kw)
    g.set_option("relu_mask", 1 if mask else 0)
    g.set_option("fused_update", fused)
    return g


def parity_case(name, sizes, bunch, n_bunches, fused=0, **kw):
    x, t = O.synth_data(bunch * n_bunches, sizes[0], sizes[-1], seed=11)
    res = []
    for mask in (0, 1):
        g = make(sizes, bunch, mask, fused, **dict(kw))
        g.train(bunch * n_bunches, x, t)
        ws, bs = g.returnWeights()
        out = g.forward(min(bunch, 64), x[: min(bunch, 64)])
        res.append((ws, bs, out))
        g.close()
    ok = True
    for l in range(1, len(sizes)):
        ok &= np.array_equal(res[0][0][l], res[1][0][l]) and np.array_equal(res[0][1][l], res[1][1][l])
    ok &= np.array_equal(res[0][2], res[1][2])
    w0 = O.glorot_init(sizes, seed=3)[0]
    moved = max(float(np.abs(res[0][0][l] - w0[l]).max()) for l in range(1, len(sizes)))
    worst = max(float(np.abs(res[0][0][l] - res[1][0][l]).max()) for l in range(1, len(sizes)))
    print(f"[{'OK ' if ok else 'BAD'}] {name}: weights moved by up to {moved:.3e}, mask vs Y max |diff| = {worst:.3e}")
    return ok


def timing(tag, sizes, bunch, steps=200, **kw):
    x, t = O.synth_data(bunch * 16, sizes[0], sizes[-1], seed=5)
    for mask in (0, 1, 0, 1):
        g = make(sizes, bunch, mask, **dict(kw))
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
        print(f"{tag} relu_mask={mask}  {ms / done:.4f} ms/bunch  {done * bunch / ms * 1e3 / 1e6:.3f} M frames/s  "
              + " ".join(f"{k}={v / max(n, 1):.4f}" for k, v in prof.items()))
        g.close()


def main():
    ok = True
    ok &= parity_case("lone CTAs, ragged words: 75-96-33, bunch 37", [75, 96, 33], 37, 4)
    ok &= parity_case("odd units 300-261-131-33, bunch 96, weight cost", [300, 261, 131, 33], 96, 4, weightcost=0.01)
    ok &= parity_case("128-wide pairs: 2827-2048-2048-257, bunch 1024", [2827, 2048, 2048, 257], 1024, 3)
    ok &= parity_case("C2 net, 4 bunches", [2827, 2048, 2048, 2048, 257], 1024, 4)
    ok &= parity_case("256-wide pairs + dropout 0.2/0.2, bunch 2048", [3084, 2048, 2048, 257], 2048, 2, dropoutflag=1,
                      visible_omit=0.2, hid_omit=0.2)
    ok &= parity_case("dropout, small", [129, 70, 50, 20], 48, 5, dropoutflag=1, visible_omit=0.1, hid_omit=0.3)
    ok &= parity_case("3xTF32", [300, 260, 130, 33], 128, 3, math_mode=bp.BP_MATH_3XTF32)
    ok &= parity_case("with the fused update", [300, 260, 130, 33], 128, 3, fused=1)
    ok &= parity_case("sigmoid net (switch ignored)", [257, 512, 257], 128, 3, activation=bp.BP_ACT_SIGMOID)
    print("PARITY", "ALL OK" if ok else "FAILED")
    if ok and (len(sys.argv) < 2 or sys.argv[1] != "parity"):
        t0 = time.time()
        timing("C2", [2827, 2048, 2048, 2048, 257], 1024)
        timing("C3", [3084, 2048, 2048, 2048, 257], 2048, dropoutflag=1, visible_omit=0.2, hid_omit=0.2)
        print(f"timing took {time.time() - t0:.1f} s")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
