# Final confirmation of the round: GPU tests, default bench line, ncu captures, then the optional probes.
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== default bench (with cpu baseline)"
timeout 120 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-3000 gpurun_out/bench_final.json
echo "== ncu launch list"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 45 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu1.err; tail -2 gpurun_out/ncu1.err
echo "== ncu full"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"bp_gemm|bp_sgd|bp_out_finish" -s 28 -c 14 -f -o gpurun_out/prof_r1d python bench.py --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu2.err; tail -2 gpurun_out/ncu2.err
echo "== isolated GEMMs, default selection"
timeout 60 python scripts/gpu_mc_probe.py quick 2>&1 | tail -10 | tee gpurun_out/isolated_default.log
echo "== multicast after the descriptor fix"
MC_DIAG=1 BP_PAIRS=3 BP_MC=2 timeout 40 python scripts/gpu_mc_probe.py quick 2>&1 | grep -v "^   n=" | tail -10 | tee gpurun_out/mc_fixed_128x2.log
BP_PAIRS=2 BP_MC=2 timeout 40 python scripts/gpu_mc_probe.py quick 2>&1 | tail -10 | tee gpurun_out/mc_fixed_256x2.log
