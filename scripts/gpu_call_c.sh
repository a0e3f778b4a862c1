# One-box round-1 confirmation: multicast diagnosis, isolated GEMM timings, A/B of the L2 hints, GPU tests, the bench
# workloads and the ncu captures of the final configuration.
mkdir -p gpurun_out
echo "== multicast diag (MN-major A slices)"
MC_DIAG=1 BP_PAIRS=3 BP_MC=2 timeout 60 python scripts/gpu_mc_probe.py quick 2>&1 | head -70
echo "== isolated GEMMs, default selection"
timeout 90 python scripts/gpu_mc_probe.py 2>&1 | tail -14
bench() { env $1 timeout 150 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/bench_$2.err > gpurun_out/bench_$2.json; python - "$2" <<'PY'
import sys, json
try:
    d = json.loads(open(f"gpurun_out/bench_{sys.argv[1]}.json").read())
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'], 4), {k: round(v, 4) for k, v in d['roofline']['per_class_ms'].items()}, 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'], 3), d['clocks']['sm_mhz'])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
echo "== bench A/B"
bench "BP_TMA_HINT=1" hint1
bench "BP_TMA_HINT=0" hint0
bench "BP_TMA_HINT=1" hint1b
bench "BP_TMA_HINT=0" hint0b
echo "== pytest -m gpu"
timeout 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "== default bench (with cpu baseline)"
timeout 200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-2600 gpurun_out/bench_final.json
echo "== reference arm"
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-600
echo "== other workloads"
for w in C3 C5; do timeout 120 python bench.py --workload $w --steps 60 --warmup 10 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_$w.json; cut -c1-1700 gpurun_out/bench_$w.json; done
echo "== ncu launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 45 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu1.err; tail -3 gpurun_out/ncu1.err
echo "== ncu full"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"bp_gemm|bp_sgd|bp_out_finish" -s 28 -c 14 -f -o gpurun_out/prof_r1d python bench.py --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu2.err; tail -3 gpurun_out/ncu2.err
ls -la gpurun_out | head -30
