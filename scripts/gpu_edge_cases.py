#!/usr/bin/env python
"""Edge cases of the C-ABI path against the oracle, to be run once on a B200 and then promoted into
tests/test_gpu_parity.py (the round-end suite runs with -x, so unproven cases wait here):
  odd bunch sizes (37, 1), a one-layer net, the deepest net the ABI allows (9 weight layers), a 1-unit output,
  a chunk shorter than one bunch (train is a no-op, like BP_GPU.cu:297-318), forward of 1 frame, CV of a ragged tail,
  weight cost with odd sizes, the same cases through the chained launches (bp_set_option "chain").
Prints one line per case; exit status 1 if any fails."""
import importlib
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
bp = importlib.import_module("dnn-for-speech-enhancement_b200")
import oracle_py as O  # noqa: E402  (the checker)

TOL = 1e-3   # whole path vs the tf32-conditioned oracle (tests/test_gpu_parity.py)


def rms(a):
    return float(np.sqrt(np.mean(np.square(a.astype(np.float64)))) + 1e-30)


def close(got, want, tol):
    return bool(np.isfinite(got).all()) and float(np.max(np.abs(got.astype(np.float64) - want))) <= tol * rms(want)


def case(name, sizes, bunch, n_frames, chain=0, **kw):
    w, b = O.glorot_init(sizes, seed=3)
    x, t = O.synth_data(max(n_frames, 1), sizes[0], sizes[-1], seed=9)
    x, t = x[:n_frames], t[:n_frames]
    o = O.Net(sizes, bunch, tf32=1, weights=w, bias=b, **kw)
    g = bp.BP_GPU(1, len(sizes), sizes, bunch, kw.get("lrate", 1.0), kw.get("momentum", 0.0), kw.get("weightcost", 0.0),
                  w, b, 0, 0.0, 0.0, activation=kw.get("activation", 0), device=0)
    g.set_option("chain", chain)
    ok = True
    if n_frames > 0:
        g.train(n_frames, x, t)
        o.train(n_frames, x, t)
    ws, bs = g.returnWeights()
    for l in range(1, len(sizes)):
        ok &= close(ws[l], o.w[l].astype(np.float64), TOL) and close(bs[l], o.b[l].astype(np.float64), 4 * TOL + 1e-6)
    ok &= g.counters()[1] == n_frames // bunch
    m = max(1, min(n_frames, bunch + 3))
    xf, tf = O.synth_data(m, sizes[0], sizes[-1], seed=10)
    ok &= close(g.forward(m, xf), o.forward(xf).astype(np.float64), 2 * TOL)
    cv, ref = g.CrossValid(m, xf, tf), o.crossvalid(xf, tf)
    ok &= abs(cv - ref) <= 4 * TOL * abs(ref) + 1e-6
    ok &= close(g.forward(1, xf[:1]), o.forward(xf[:1]).astype(np.float64), 2 * TOL)
    g.close()
    print(f"[{'OK ' if ok else 'BAD'}] {name}{' (chained launches)' if chain else ''}")
    return ok


def main():
    ok = True
    for chain in (0, 1):
        for args in (("bunch 37, ragged tail", [75, 96, 33], 37, 3 * 37 + 5, dict(lrate=0.7, momentum=0.9)),
                     ("bunch 1", [40, 24, 8], 1, 5, dict(lrate=0.1, momentum=0.5)),
                     ("one weight layer", [129, 65], 32, 96, dict(momentum=0.9)),
                     ("nine weight layers", [64, 72, 40, 96, 33, 80, 48, 56, 64, 20], 32, 64, dict(lrate=0.5)),
                     ("one output unit", [90, 70, 1], 32, 64, dict(momentum=0.9)),
                     ("chunk shorter than a bunch", [75, 96, 33], 64, 20, dict()),
                     ("weight cost, odd sizes", [257, 131, 67, 3], 40, 120, dict(lrate=0.5, momentum=0.9, weightcost=1e-3)),
                     ("sigmoid, odd sizes", [61, 45, 29], 24, 72, dict(activation=1, momentum=0.9))):
            try:
                ok &= case(args[0], args[1], args[2], args[3], chain, **args[4])
            except Exception:
                ok = False
                print(f"[EXC] {args[0]}{' (chained launches)' if chain else ''}")
                traceback.print_exc()
    print("EDGE CASES", "ALL OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
