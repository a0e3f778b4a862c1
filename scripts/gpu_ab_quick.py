"""Quick in-process A/B of a few option sets on one workload: python scripts/gpu_ab_quick.py C2 "chain=0" "chain=0 mc=2" ..."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
bp = importlib.import_module("dnn-for-speech-enhancement_b200")
wl = sys.argv[1]
sets = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in a.split()) for a in sys.argv[2:]]
sizes, lb, dflag, vo, ho, _ = bench.WORKLOADS[wl]
w, b = bench.glorot(sizes)
g = bp.BP_GPU(1, len(sizes), sizes, lb, 1.0, 0.9, 0.0, w, b, dflag, vo, ho, seed=12345, device=0)
cb = 32
px, pt = bp.PinnedArray((cb * lb, sizes[0])), bp.PinnedArray((cb * lb, sizes[-1]))
bench.synth(cb * lb, sizes[0], sizes[-1], seed=100, out_x=px.array, out_t=pt.array)
g.upload_chunk(cb * lb, px.array, pt.array)
fresh = dflag == 1 and vo > 0   # input dropout masks the resident rows in place: a fresh upload per pass (untimed)


def passes(n):
    ms = 0.0
    for _ in range(n):
        if fresh:
            g.upload_chunk(cb * lb, px.array, pt.array)
        g.sync()
        g.timer_start()
        g.train_resident(0, cb)
        ms += g.timer_stop()
    return ms


keys = sorted({k for s in sets for k in s})
base = {k: g.get_option(k) for k in keys}
for rnd in range(2):
    for s in sets:
        for k in keys:
            g.set_option(k, s.get(k, base[k]))
        passes(1)
        ms = passes(4) / (4 * cb)
        if fresh:
            g.upload_chunk(cb * lb, px.array, pt.array)
        g.set_profiling(True)
        g.train_resident(0, 16)
        prof, n = g.profile()
        g.set_profiling(False)
        print(f"{wl} {str(s):40s} {ms:.4f} ms/bunch  {lb / ms / 1e3:.3f} M frames/s  "
              + " ".join(f"{k}={v / max(n, 1):.4f}" for k, v in prof.items()), flush=True)
g.close()
