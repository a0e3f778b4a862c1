#!/usr/bin/env python
"""A/B of scheduling / kernel-choice switches inside ONE process: one trainer, one resident chunk, the settings applied
through bp_set_option between timed runs (all of them leave every result bit unchanged), each setting measured twice in
alternation with the baseline so that drift of the box shows up.  One GPU-minute instead of one process per setting.
   python scripts/gpu_ab_inproc.py [C2|C3|C4] [steps]
Each line: setting, ms per bunch (both passes), M frames/s of the better pass, per-class ms of the second pass."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

bp = importlib.import_module("dnn-for-speech-enhancement_b200")

BASE = dict(fused_update=0, relu_mask=0, l2_prefetch=0, l2_persist=0, stages=0, dw_stream=0, small_pairs=0,
            sgd_early=6, pdl=1, tma_hint=1)
SETTINGS = [
    ("baseline", {}),
    ("fused_update", dict(fused_update=1)),
    ("relu_mask", dict(relu_mask=1)),
    ("fused_update + relu_mask", dict(fused_update=1, relu_mask=1)),
    ("l2_prefetch=4", dict(l2_prefetch=4)),
    ("l2_prefetch=8", dict(l2_prefetch=8)),
    ("l2_prefetch=16", dict(l2_prefetch=16)),
    ("l2_persist=64", dict(l2_persist=64)),
    ("l2_persist=96", dict(l2_persist=96)),
    ("stages=2", dict(stages=2)),
    ("stages=3", dict(stages=3)),
    ("stages=2 + l2_prefetch=8", dict(stages=2, l2_prefetch=8)),
    ("dw_stream", dict(dw_stream=1)),
    ("small_pairs", dict(small_pairs=1)),
    ("fused + mask + l2_persist=64", dict(fused_update=1, relu_mask=1, l2_persist=64)),
    ("fused + mask + l2_prefetch=8", dict(fused_update=1, relu_mask=1, l2_prefetch=8)),
    ("fused + mask + stages=2", dict(fused_update=1, relu_mask=1, stages=2)),
]


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    sizes, lb, dflag, vo, ho, train = bench.WORKLOADS[wl]
    assert train
    w, b = bench.glorot(sizes)
    g = bp.BP_GPU(1, len(sizes), sizes, lb, 1.0, 0.9, 0.0, w, b, dflag, vo, ho, seed=12345, device=0)
    cb = 32
    px, pt = bp.PinnedArray((cb * lb, sizes[0])), bp.PinnedArray((cb * lb, sizes[-1]))
    bench.synth(cb * lb, sizes[0], sizes[-1], seed=100, out_x=px.array, out_t=pt.array)
    g.upload_chunk(cb * lb, px.array, pt.array)

    def apply(opts):
        for k, v in dict(BASE, **opts).items():
            g.set_option(k, v)

    def measure(opts):
        apply(opts)
        g.train_resident(0, cb)          # warm-up with the setting in force
        g.sync()
        g.timer_start()
        done = 0
        while done < steps:
            g.train_resident(0, min(cb, steps - done))
            done += min(cb, steps - done)
        ms = g.timer_stop() / done
        g.set_profiling(True)
        g.train_resident(0, 16)
        prof, n = g.profile()
        g.set_profiling(False)
        return ms, {k: round(v / max(n, 1), 4) for k, v in prof.items()}

    for _ in range(3):
        g.train_resident(0, cb)
    results = {name: [] for name, _ in SETTINGS}
    for rnd in range(2):
        for name, opts in SETTINGS:
            results[name].append(measure(opts))
    apply({})
    base = min(m for m, _ in results["baseline"])
    print(f"{wl}, {lb} frames per bunch, {steps} timed bunches per pass; baseline {base:.4f} ms per bunch")
    for name, _ in SETTINGS:
        (m1, _p1), (m2, p2) = results[name]
        best = min(m1, m2)
        print(f"{name:34s} {m1:.4f} {m2:.4f} ms  {lb / best / 1e3:6.3f} M frames/s  {100 * (base / best - 1):+5.1f} %  {p2}")
    g.close()


if __name__ == "__main__":
    main()
