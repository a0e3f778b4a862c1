"""Chained-products launch (bp_chain.cuh) against one launch per product: bit parity, then ms per bunch.
    python scripts/gpu_chain_check.py
Parity: the same net trained with chain=0 (and splitk=0, so that the output layer is summed in one pass as the chain
does) and chain=1 must end with identical weights, over edge shapes, dropout, sigmoid, 3xTF32 and the BASELINE shapes."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
bp = importlib.import_module("dnn-for-speech-enhancement_b200")
import oracle_py as O  # noqa: E402
import bench  # noqa: E402

CASES = [
    ("tiny ragged 75-96-33 B=37", [75, 96, 33], 37, 4, {}),
    ("one layer 129-65 B=32", [129, 65], 32, 3, {}),
    ("9 layers", [64, 72, 40, 96, 33, 80, 48, 56, 64, 20], 32, 2, {}),
    ("odd 300-261-131-33 B=96 wc", [300, 261, 131, 33], 96, 3, dict(weightcost=1e-3)),
    ("dropout 129-70-50-20 B=48", [129, 70, 50, 20], 48, 4, dict(dropoutflag=1, visible_omit=0.1, hid_omit=0.3)),
    ("sigmoid 61-45-29 B=24", [61, 45, 29], 24, 3, dict(activation=1)),
    ("3xTF32 300-512-129 B=256", [300, 512, 129], 256, 2, dict(math_mode=1)),
    ("pairs 2827-2048-2048-257 B=1024", [2827, 2048, 2048, 257], 1024, 2, {}),
    ("C2 x3 bunches", [2827, 2048, 2048, 2048, 257], 1024, 3, {}),
    ("C3 dropout B=2048", [3084, 2048, 2048, 2048, 257], 2048, 2, dict(dropoutflag=1, visible_omit=0.2, hid_omit=0.2)),
    ("C4 local B=512", [2827, 2048, 2048, 2048, 2048, 2048, 257], 512, 2, {}),
]


def train(sizes, bunch, nb, chain, kw):
    w, b = O.glorot_init(sizes, seed=3)
    x, t = O.synth_data(bunch * nb + 3, sizes[0], sizes[-1], seed=11)
    g = bp.BP_GPU(1, len(sizes), sizes, bunch, 1.0, 0.9, kw.get("weightcost", 0.0), w, b, kw.get("dropoutflag", 0),
                  kw.get("visible_omit", 0.0), kw.get("hid_omit", 0.0), device=0, seed=777,
                  activation=kw.get("activation", 0), math_mode=kw.get("math_mode", 0))
    g.set_option("chain", chain)
    g.set_option("splitk", 0)
    g.train(x.shape[0], x, t)
    ws, bs = g.returnWeights()
    out = g.forward(min(64, x.shape[0]), x[:64])
    g.close()
    return ws, bs, out, w


def main():
    ok = True
    for name, sizes, bunch, nb, kw in CASES:
        a = train(sizes, bunch, nb, 0, kw)
        c = train(sizes, bunch, nb, 1, kw)
        worst = 0.0
        same = True
        for l in range(1, len(sizes)):
            same &= np.array_equal(a[0][l], c[0][l]) and np.array_equal(a[1][l], c[1][l])
            worst = max(worst, float(np.abs(a[0][l] - c[0][l]).max()))
        moved = max(float(np.abs(c[0][l] - c[3][l]).max()) for l in range(1, len(sizes)))
        finite = all(np.isfinite(c[0][l]).all() for l in range(1, len(sizes)))
        print(f"[{'OK ' if same and finite and moved > 0 else 'BAD'}] {name}: weights moved by up to {moved:.3e}, "
              f"chain vs per-product max |diff| = {worst:.3e}", flush=True)
        ok &= same and finite and moved > 0
    print("PARITY", "ALL OK" if ok else "FAILED", flush=True)
    # ---- timing, alternating
    for wl in ("C2", "C3", "C4"):
        sizes, lb, dflag, vo, ho, _ = bench.WORKLOADS[wl]
        w, b = bench.glorot(sizes)
        g = bp.BP_GPU(1, len(sizes), sizes, lb, 1.0, 0.9, 0.0, w, b, dflag, vo, ho, seed=12345, device=0)
        cb = 32
        px, pt = bp.PinnedArray((cb * lb, sizes[0])), bp.PinnedArray((cb * lb, sizes[-1]))
        bench.synth(cb * lb, sizes[0], sizes[-1], seed=100, out_x=px.array, out_t=pt.array)
        g.upload_chunk(cb * lb, px.array, pt.array)
        fresh = dflag == 1 and vo > 0   # input dropout masks the resident rows in place: fresh upload per pass

        def passes(n):
            ms_ = 0.0
            for _ in range(n):
                if fresh:
                    g.upload_chunk(cb * lb, px.array, pt.array)
                g.sync()
                g.timer_start()
                g.train_resident(0, cb)
                ms_ += g.timer_stop()
            return ms_
        for rnd in range(2):
            for chain in (0, 1):
                g.set_option("chain", chain)
                passes(1)
                ms = passes(4) / (4 * cb)
                if fresh:
                    g.upload_chunk(cb * lb, px.array, pt.array)
                g.set_profiling(True)
                g.train_resident(0, 16)
                prof, n = g.profile()
                g.set_profiling(False)
                print(f"{wl} chain={chain}  {ms:.4f} ms/bunch  {lb / ms / 1e3:.3f} M frames/s  "
                      + " ".join(f"{k}={v / max(n, 1):.4f}" for k, v in prof.items()), flush=True)
        g.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
