# Validation call (1 GPU) for the tree as committed: GPU suite, smoke, the driver's two bench arms, then two short probes
# (host -> device copy rates; in-process sweep of the two scheduling switches that take a value).
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.jsonl
echo "=== pytest -m gpu"; timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench (defaults)"; timeout 400 python bench.py > gpurun_out/r2i_bench_C2.json 2> gpurun_out/r2i_bench_C2.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2i_bench_C2.json"))
r = d["roofline"]
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "raw", d["e2e"].get("raw_reader", {}).get("value"), "steady", d["steady"]["ms_per_step"], d["clocks"])
print("roofline frac", r["frac"], "sustained", r["frac_of_sustained_peak"], r["per_class_ms"])
print(d.get("ref_gpu", {}).get("value"), d.get("tf32x3", {}).get("value"), d.get("cpu_baseline", {}).get("value"), d.get("gpu_launches"))
PY
tail -3 gpurun_out/r2i_bench_C2.err
echo "=== bench --impl reference"; timeout 300 python bench.py --impl reference --steps 4 --warmup 1 2>/dev/null | cut -c1-400
echo "=== h2d probe"; timeout 100 python scripts/gpu_h2d_probe.py 2>&1 | tail -5
echo "=== sweep sgd_early / splitk (C2)"
timeout 200 python scripts/gpu_ab_quick.py C2 "sgd_early=6" "sgd_early=2" "sgd_early=4" "sgd_early=8" "sgd_early=12" "sgd_early=16" "sgd_early=0" \
  "splitk=-1" "splitk=0" "splitk=2" "splitk=3" "splitk=4" "splitk=8" 2>&1 | grep -v "^this bunch" | tail -30
