#!/usr/bin/env python
"""Launch timeline of a training bunch (bp_set_profiling(h, 2) / bp_get_timeline): completion time of every launch
since the start of the bunch, averaged over 16 bunches, caches as the pipeline leaves them.  Answers what the per-class
events and ncu's serialised, cache-flushed replays cannot: which launch of the dependency chain costs what in place
(the forward chain takes 0.119 ms inside a bunch against ~0.078 ms for its kernels alone).
   python scripts/gpu_timeline.py [C2|C3|C4]        (environment switches such as BP_L2_PREFETCH apply)"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload table, synthetic data, Glorot init)

bp = importlib.import_module("dnn-for-speech-enhancement_b200")


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    sizes, lb, dflag, vo, ho, train = bench.WORKLOADS[wl]
    assert train, "training workloads only"
    w, b = bench.glorot(sizes)
    g = bp.BP_GPU(1, len(sizes), sizes, lb, 1.0, 0.9, 0.0, w, b, dflag, vo, ho, seed=12345, device=0)
    cb = 32
    px, pt = bp.PinnedArray((cb * lb, sizes[0])), bp.PinnedArray((cb * lb, sizes[-1]))
    bench.synth(cb * lb, sizes[0], sizes[-1], seed=100, out_x=px.array, out_t=pt.array)
    g.upload_chunk(cb * lb, px.array, pt.array)
    for _ in range(4):
        g.train_resident(0, cb)
    g.sync()
    g.timer_start()
    g.train_resident(0, cb)
    ms_plain = g.timer_stop() / cb
    g.set_timeline(True)
    g.timer_start()
    g.train_resident(0, 16)
    ms_marked = g.timer_stop() / 16
    marks, nb = g.timeline()
    g.set_timeline(False)
    print(f"{wl}: {ms_plain:.4f} ms per bunch without marks, {ms_marked:.4f} ms with a mark behind every launch "
          f"({nb} bunches recorded)")
    print(f"{'launch':34s} {'done at (us)':>12s} {'since previous on its stream (us)':>34s}")
    last = {"compute": 0.0, "side": 0.0}
    for name, t in marks:
        st = "side" if "(side)" in name else "compute"
        prev = last[st]
        print(f"{name:34s} {t * 1e3:12.1f} {((t - prev) * 1e3):34.1f}")
        last[st] = t
    g.close()


if __name__ == "__main__":
    main()
