# One-box validation + A/B of the multicast-cluster GEMM variants, then tests / bench / ncu with the better setting.
mkdir -p gpurun_out
echo "== multicast probe"
MC_OK=1
for cfg in "3 1" "3 2" "3 4" "2 1" "2 2" "2 4"; do
  set -- $cfg
  BP_PAIRS=$1 BP_MC=$2 timeout 90 python scripts/gpu_mc_probe.py quick 2>&1 | tail -12
  [ ${PIPESTATUS[0]} -eq 0 ] || { [ "$2" != "1" ] && MC_OK=0; }
done
BP_VERBOSE=1 BP_MC=-1 timeout 120 python scripts/gpu_mc_probe.py 2>&1 | tail -15
[ ${PIPESTATUS[0]} -eq 0 ] || MC_OK=0
echo "MC_OK=$MC_OK"
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
bench "BP_MC=1" mc1
USE=""
if [ $MC_OK -eq 1 ]; then
  bench "BP_MC=-1" auto
  bench "BP_MC=1" mc1b
  bench "BP_MC=-1" autob
  USE=$(python - <<'PY'
import json
def v(n):
    try: return json.loads(open(f"gpurun_out/bench_{n}.json").read())['value']
    except Exception: return 0.0
a = max(v('auto'), v('autob')); b = max(v('mc1'), v('mc1b'))
print("BP_MC=-1" if a > b * 1.01 else "")
PY
)
fi
echo "== chosen env: '$USE'"
echo "== pytest -m gpu"
env $USE timeout 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "== default bench (with cpu baseline)"
env $USE timeout 200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-2500 gpurun_out/bench_final.json
echo "== other workloads"
for w in C3 C5; do env $USE timeout 120 python bench.py --workload $w --steps 60 --warmup 10 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_$w.json; cut -c1-900 gpurun_out/bench_$w.json; done
echo "== ncu launch list"
env $USE timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 45 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu1.err; tail -3 gpurun_out/ncu1.err
echo "== ncu full"
env $USE timeout 300 ncu --set full --clock-control none --import-source on -k regex:"bp_gemm|bp_sgd|bp_out_finish" -s 28 -c 14 -o gpurun_out/prof_r1d python bench.py --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu2.err; tail -3 gpurun_out/ncu2.err
ls -la gpurun_out | head -30
