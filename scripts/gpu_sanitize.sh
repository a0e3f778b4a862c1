# compute-sanitizer over the hot path (the reference has no sanitizer coverage at all).
# memcheck on the edge-case script (every GEMM variant at odd shapes, per-product and chained launches, split-K and the
# finisher, SGD, fill) and on smoke() (dropout, ragged chunk, CV, decode + one C2-shaped bunch through the CTA-pair
# kernels and bp_chain_kernel); synccheck and racecheck on smoke().  racecheck does not model mbarrier / TMA completion,
# so its reports on the GEMM rings need reading, not counting.
# usage: gpurun --timeout 1200 -- 'bash scripts/gpu_sanitize.sh 2>&1 | tee gpurun_out/r2_sanitizer.log'
CS=/usr/local/cuda/bin/compute-sanitizer
mkdir -p gpurun_out
echo "== memcheck: edge cases (per-product and chained launches)"
timeout 420 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python scripts/gpu_edge_cases.py 2>&1 | tail -24
echo "== memcheck: smoke (dropout, ragged chunk, CV, decode, C2-shaped bunch: pair kernels + chain kernel)"
timeout 400 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
echo "== memcheck: device reader + raw decode"
timeout 300 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_raw_reader.py tests/test_pipeline_edges_chain.py -q -m gpu -k "decode or raw or pipelined" 2>&1 | tail -8
echo "== synccheck: smoke"
timeout 400 $CS --tool synccheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
echo "== racecheck: smoke (read the reports: mbarrier/TMA hand-overs are not modelled)"
timeout 500 $CS --tool racecheck --racecheck-report analysis --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -16
