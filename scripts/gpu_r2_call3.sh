# Round 2, call 3: the whole GPU suite (gates removed, config-shape parity tests added) + the new bench line.
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.jsonl
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
echo "=== parity errors"; cat gpurun_out/parity_errors.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(f\"{d['achieved']:.3e} (bound {d['bound']:.0e})  {d['test']} :: {d['what']}\")"
echo "=== bench C2 (defaults)"; timeout 400 python bench.py > gpurun_out/r2_bench3_C2.json 2> gpurun_out/r2_bench3_C2.err; cut -c1-6000 gpurun_out/r2_bench3_C2.json; tail -5 gpurun_out/r2_bench3_C2.err
echo "=== bench C2 --steps 20 --warmup 3 (as the driver)"; timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-3xtf32 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('raw_reader'), d['steady'])"
echo "=== bench reference arm"; OMP_NUM_THREADS=1 timeout 400 python bench.py --impl reference --steps 5 --warmup 3 | cut -c1-300
