# Round 2 call A (1 GPU): GPU suite, fused-update A/B, ncu launch list + full captures summarised on the box.
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.jsonl
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8
grep -h "C5 decode" gpurun_out/parity_errors.jsonl | cut -c1-200
echo "=== fused update A/B"; for w in C2 C3 C4; do timeout 200 python scripts/gpu_ab_quick.py $w "chain=0 fused_update=0" "chain=0 fused_update=1" 2>&1 | tail -4; done
echo "=== ncu launch list (C2, per-product path)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 28 --csv --log-file gpurun_out/r2_launches_C2.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-3xtf32 --no-timeline --no-e2e-raw --steady-seconds 0 > /dev/null 2>&1; wc -l gpurun_out/r2_launches_C2.csv
full() {  # name, extra env, -s, -c, regex, bench args...
  local name=$1 envs=$2 skip=$3 cnt=$4 rx=$5; shift 5
  env $envs timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o /tmp/$name -f python bench.py "$@" --steps 4 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-3xtf32 --no-timeline --no-e2e-raw --steady-seconds 0 > /dev/null 2>&1
  python scripts/ncu_summary.py /tmp/$name.ncu-rep > gpurun_out/$name.md 2>&1; tail -n +4 gpurun_out/$name.md | cut -c1-250
}
echo "=== ncu --set full: C2 per-product bunch"; full r2_C2_ncu_full BP_X=0 42 14 "bp_gemm|bp_sgd|bp_out_finish"
echo "=== ncu --set full: C2 chained bunch"; full r2_C2_chain_ncu_full BP_CHAIN=1 9 3 "bp_chain|bp_sgd"
echo "=== ncu --set full: C3 bunch"; full r2_C3_ncu_full BP_X=0 36 12 "bp_gemm" --workload C3
echo "=== ncu --set full: C5 batch"; full r2_C5_ncu_full BP_X=0 20 5 "bp_gemm" --workload C5
