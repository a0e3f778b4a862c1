# compute-sanitizer over the hot path at odd shapes (the reference has no sanitizer coverage at all).
# memcheck on everything the edge-case script and the smoke run launch (every GEMM variant incl. split-K and the
# finisher, SGD, input dropout, splice, fill); racecheck / synccheck on the smoke run only (they serialise heavily, and
# racecheck does not model mbarrier / TMA completion, so its reports on the GEMM rings need reading, not counting).
# usage: gpurun --timeout 900 -- 'bash scripts/gpu_sanitize.sh 2>&1 | tee gpurun_out/sanitize.log'
CS=/usr/local/cuda/bin/compute-sanitizer
mkdir -p gpurun_out
echo "== memcheck: edge cases"
timeout 420 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python scripts/gpu_edge_cases.py 2>&1 | tail -30
echo "== memcheck: smoke (dropout, ragged chunk, CV, decode)"
timeout 300 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12
echo "== memcheck: device reader + raw decode (tests, not under -x)"
timeout 300 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_raw_reader.py tests/test_zz_late_round1.py -q -m gpu -k "not bptrain" 2>&1 | tail -12
echo "== synccheck: smoke"
timeout 300 $CS --tool synccheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12
echo "== racecheck: smoke (read the reports: mbarrier/TMA hand-overs are not modelled)"
timeout 400 $CS --tool racecheck --racecheck-report analysis --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -25
