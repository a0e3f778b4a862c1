"""Per-tile timeline of the chained launches (bp_debug_chain_trace): where a bunch's forward / back-propagation
launch spends its time.   python scripts/gpu_chain_trace.py [C2|C3|C4] [detail]"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
bp = importlib.import_module("dnn-for-speech-enhancement_b200")
import bench  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    detail = "detail" in sys.argv[2:]
    opts = dict(a.split("=") for a in sys.argv[2:] if "=" in a)   # e.g. chain_prefetch=0 chain_fwd_only=1
    if wl.startswith("net:"):     # net:2048-2048:1024 = layer sizes : bunch
        _, ls, b_ = wl.split(":")
        sizes, lb, dflag, vo, ho = [int(v) for v in ls.split("-")], int(b_), 0, 0.0, 0.0
    else:
        sizes, lb, dflag, vo, ho, _ = bench.WORKLOADS[wl]
    w, b = bench.glorot(sizes)
    g = bp.BP_GPU(1, len(sizes), sizes, lb, 1.0, 0.9, 0.0, w, b, dflag, vo, ho, seed=12345, device=0)
    cb = int(opts.pop("cb", 32))
    px, pt = bp.PinnedArray((cb * lb, sizes[0])), bp.PinnedArray((cb * lb, sizes[-1]))
    bench.synth(cb * lb, sizes[0], sizes[-1], seed=100, out_x=px.array, out_t=pt.array)
    g.upload_chunk(cb * lb, px.array, pt.array)
    g.set_option("chain", 1)
    for k, v in opts.items():
        g.set_option(k, int(v))
    fresh = dflag == 1 and vo > 0   # input dropout masks the resident rows in place: fresh upload per pass
    ms = 0.0
    for i in range(5):
        if fresh:
            g.upload_chunk(cb * lb, px.array, pt.array)
        g.sync()
        g.timer_start()
        g.train_resident(0, cb)
        ms += g.timer_stop() if i else 0.0
    ms /= 4 * cb
    print(f"#### {wl} {opts}: {ms:.4f} ms per bunch")
    g.set_option("chain_trace", 1)
    if fresh:
        g.upload_chunk(cb * lb, px.array, pt.array)
    g.train_resident(0, cb)
    g.sync()
    lib = bp.load_library()
    lib.bp_debug_chain_trace.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    for which, name in ((0, "forward"), (1, "back-propagation"))[: 1 if opts.get("chain_fwd_only") == "1" else 2]:
        buf = C.create_string_buffer(1 << 20)
        rc = lib.bp_debug_chain_trace(g.handle, which, buf, len(buf))
        if rc:
            print(lib.bp_last_error().decode())
            return 1
        rows = np.array([[int(v) for v in l.split(",")] for l in buf.value.decode().strip().split("\n")])
        print(f"==== {wl} {name}: {len(rows)} tiles; times in us since the first stamp")
        print("prod tiles | deps satisfied (min..max) | accumulator ready (min..max) | stores issued (min..max) | "
              "published (min..max) | deps->acc us (median) | acc->stores issued | stores issued->published")
        for q in sorted(set(rows[:, 2])):
            r = rows[rows[:, 2] == q]
            us = r[:, 5:9] / 1e3
            print(f"{q:4d} {len(r):5d} | {us[:,0].min():7.1f} .. {us[:,0].max():7.1f} | {us[:,2].min():7.1f} .. {us[:,2].max():7.1f} | "
                  f"{us[:,1].min():7.1f} .. {us[:,1].max():7.1f} | {us[:,3].min():7.1f} .. {us[:,3].max():7.1f} | "
                  f"{np.median(us[:,2]-us[:,0]):6.2f} | {np.median(us[:,1]-us[:,2]):6.2f} | {np.median(us[:,3]-us[:,1]):6.2f}")
        if detail:
            for p in (0, 1, 36, 70, 73):
                r = rows[rows[:, 1] == p]
                print(f"  pair {p}: " + "  ".join(f"[q{x[2]} m{x[3]} n{x[4]}: {x[5]/1e3:.1f} {x[6]/1e3:.1f} {x[7]/1e3:.1f} {x[8]/1e3:.1f}]" for x in r))
    g.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
