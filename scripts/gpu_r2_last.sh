# Last call of the round (1 GPU): GPU suite and smoke on the final library, a short bench line, and the host-side cost of
# queueing a bunch (is the per-product path anywhere near launch-bound?).
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.jsonl
echo "=== pytest -m gpu"; timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
echo "=== smoke"; timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (short)"; timeout 150 python bench.py --no-cpu-baseline --no-ref-gpu --no-3xtf32 --steady-seconds 0.5 > gpurun_out/r2l_bench_C2.json 2> gpurun_out/r2l_bench_C2.err; python -c "
import json; d = json.load(open('gpurun_out/r2l_bench_C2.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('raw_reader', {}).get('value'), d['roofline']['frac'], d['gpu_launches'])"
echo "=== host time to queue a bunch"; timeout 100 python - <<'PY' 2>&1 | grep -v "^this bunch"
import importlib, sys, time
sys.path.insert(0, ".")
import bench
bp = importlib.import_module("dnn-for-speech-enhancement_b200")
sizes, lb = [2827, 2048, 2048, 2048, 257], 1024
w, b = bench.glorot(sizes)
g = bp.BP_GPU(1, len(sizes), sizes, lb, 1.0, 0.9, 0.0, w, b, 0, 0.0, 0.0, seed=1, device=0)
cb = 32
px, pt = bp.PinnedArray((cb * lb, sizes[0])), bp.PinnedArray((cb * lb, sizes[-1]))
bench.synth(cb * lb, sizes[0], sizes[-1], seed=100, out_x=px.array, out_t=pt.array)
g.upload_chunk(cb * lb, px.array, pt.array)
for chain in (0, 1):
    g.set_option("chain", chain)
    g.train_resident(0, cb); g.sync()
    host, dev = [], []
    for _ in range(5):
        g.sync(); g.timer_start(); t0 = time.perf_counter()
        g.train_resident(0, cb)
        host.append((time.perf_counter() - t0) / cb * 1e3)
        dev.append(g.timer_stop() / cb)
    print(f"chain={chain}: host {min(host):.4f} ms per bunch to queue, device {min(dev):.4f} ms per bunch")
g.close()
PY
