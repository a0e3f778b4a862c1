# Round 2, call 4: branch-free epilogues (FWD_HID / DX / FWD_OUT) — whole GPU suite without -x, parity errors, bench.
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.jsonl
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15
echo "=== parity errors"; python - <<'PY'
import json
for l in open("gpurun_out/parity_errors.jsonl"):
    d = json.loads(l); print(f"{d['achieved']:.3e} (bound {d['bound']:.0e})  {d['test']} :: {d['what']}")
PY
echo "=== isolated"; timeout 60 python scripts/gpu_mc_probe.py quick 2>&1 | tail -9
echo "=== dxgap"; timeout 120 python scripts/gpu_mc_probe.py dxgap 2>&1 | tail -14
echo "=== bench C2"; timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2_bench4_C2.json 2> gpurun_out/r2_bench4_C2.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench4_C2.json"))
r = d["roofline"]
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "steady", d["steady"]["ms_per_step"], "3x", d.get("tf32x3"))
print(r["per_class_ms"]); print(r["launch_timeline"]["launch_us"]); print(r.get("isolated_dominant_kernel")); print(d.get("ref_gpu"))
PY
tail -3 gpurun_out/r2_bench4_C2.err
for w in C3 C5; do echo "=== bench $w"; timeout 300 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline --no-ref-gpu --no-3xtf32 --steady-seconds 0 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('per_class_ms'), {k: (v['value'] if isinstance(v, dict) and 'value' in v else v) for k, v in d['e2e'].items() if k in ('value','pipelined','raw_reader','raw_reader_pipelined')})"; done
