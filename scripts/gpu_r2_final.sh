# Round 2 validation call (1 GPU): GPU suite, smoke, bench lines, ncu launch list + full captures, sanitizer.
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.jsonl
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench C2 (defaults)"; timeout 500 python bench.py > gpurun_out/r2_bench_final_C2.json 2> gpurun_out/r2_bench_final_C2.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_final_C2.json"))
r = d["roofline"]
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "raw", d["e2e"].get("raw_reader", {}).get("value"), "steady", d["steady"]["ms_per_step"], d["steady"]["clocks"])
print("roofline frac", r["frac"], "sustained", r["frac_of_sustained_peak"], r["per_class_ms"]); print(r["launch_timeline"]["launch_us"])
print(r.get("isolated_dominant_kernel")); print(d.get("ref_gpu")); print(d.get("tf32x3")); print(d.get("cpu_baseline")); print(d.get("roofline_sgd"))
PY
tail -3 gpurun_out/r2_bench_final_C2.err
for w in C3 C5 C4; do echo "=== bench $w"; timeout 300 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline --no-ref-gpu --no-3xtf32 --steady-seconds 0.5 > gpurun_out/r2_bench_final_$w.json 2> gpurun_out/r2_bench_final_$w.err; python -c "
import sys, json; d = json.load(open('gpurun_out/r2_bench_final_$w.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('per_class_ms'), {k: (v['value'] if isinstance(v, dict) and 'value' in v else v) for k, v in d['e2e'].items() if k in ('value','pipelined','raw_reader','raw_reader_pipelined')})"; tail -2 gpurun_out/r2_bench_final_$w.err; done
echo "=== ncu launch list (C2, one bunch = 14 launches; per-product path)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 42 --csv --log-file gpurun_out/r2_launches_C2.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-3xtf32 --no-timeline --no-e2e-raw --steady-seconds 0 > /dev/null 2>&1; tail -30 gpurun_out/r2_launches_C2.csv | cut -d, -f5,10-15 | cut -c1-200
echo "=== ncu --set full: C2 per-product bunch"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bp_gemm|bp_sgd|bp_out_finish" -s 42 -c 14 -o gpurun_out/r2_C2_full -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-3xtf32 --no-timeline --no-e2e-raw --steady-seconds 0 > /dev/null 2>&1; ls -la gpurun_out/r2_C2_full.ncu-rep
echo "=== ncu --set full: C2 chained bunch (2 chain launches + update)"
BP_CHAIN=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bp_chain|bp_sgd" -s 9 -c 3 -o gpurun_out/r2_C2_chain_full -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-3xtf32 --no-timeline --no-e2e-raw --steady-seconds 0 > /dev/null 2>&1; ls -la gpurun_out/r2_C2_chain_full.ncu-rep
echo "=== ncu --set full: C3 bunch / C5 batch (256-wide pair kernels)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bp_gemm" -s 36 -c 12 -o gpurun_out/r2_C3_full -f python bench.py --workload C3 --steps 4 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-3xtf32 --no-timeline --no-e2e-raw --steady-seconds 0 > /dev/null 2>&1; ls -la gpurun_out/r2_C3_full.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bp_gemm" -s 20 -c 5 -o gpurun_out/r2_C5_full -f python bench.py --workload C5 --steps 4 --warmup 3 --no-cpu-baseline --steady-seconds 0 > /dev/null 2>&1; ls -la gpurun_out/r2_C5_full.ncu-rep
echo "=== sanitizer"; bash scripts/gpu_sanitize.sh 2>&1 | tee gpurun_out/r2_sanitizer.log | tail -50
