# First GPU call of round 2 (1 GPU, ~12 min of box time).  Round 1 ran out of GPU minutes before the last host-side and
# gated changes could be run on a B200; this call proves them in the order "what the round-end driver runs" first:
#   1. GPU test suite (includes tests/test_zz_late_round1.py, new since the last GPU call)
#   2. smoke()
#   3. default bench line (C2)                  -> gpurun_out/r2_bench_C2.json
#   4. fused-update and ReLU-mask bit parity + timing (edge shapes: gated tests of step 1)
#   5. reader throughput end to end from Pfiles (host reader on all cores, chunk prefetch, upload wait deferred)
#   6. C5 line with the raw-records e2e
# usage: gpurun --timeout 1200 -- 'bash scripts/gpu_round2_call1.sh 2>&1 | tee gpurun_out/r2_call1.log'
mkdir -p gpurun_out
echo "== 1. pytest -m gpu (as the driver runs it), then the gated tests of never-run code paths"
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
BP_TEST_UNPROVEN=1 timeout 400 python -m pytest tests/test_zz_late_round1.py tests/test_shim.py -q -m gpu 2>&1 | tail -8
echo "== 2. smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== 3. default bench"
timeout 200 python bench.py --e2e-raw --timeline > gpurun_out/r2_bench_C2.json 2> gpurun_out/r2_bench_C2.err; cut -c1-3500 gpurun_out/r2_bench_C2.json; tail -3 gpurun_out/r2_bench_C2.err
echo "== 3b. e2e with the old / new ordering of upload wait and bunch queueing in bp_train (twice each)"
bash scripts/gpu_ab.sh "BP_UPLOAD_WAIT_FIRST=1" "BP_UPLOAD_WAIT_FIRST=0" "BP_UPLOAD_WAIT_FIRST=1" "BP_UPLOAD_WAIT_FIRST=0"
echo "== 4. fused update: bit parity + timing (the edge shapes ran as gated tests in step 1)"
timeout 300 python scripts/gpu_fused_update_check.py 2>&1 | tail -20
echo "== 4b. ReLU bit mask: bit parity + timing"
timeout 300 python scripts/gpu_relu_mask_check.py 2>&1 | tail -24
echo "== 5. reader bench (host / gpu reader, prefetch on / off)"
timeout 400 python scripts/gpu_reader_bench.py 1500 2>&1 | tail -12
echo "== 5b. the reference's class BP_GPU, both implementations behind it (reference CUDA path vs ours through the shim)"
timeout 300 python scripts/gpu_ref_gpu_timing.py 16 2>&1 | tail -5
echo "== 6. C5"
timeout 200 python bench.py --workload C5 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_C5.json 2> gpurun_out/r2_bench_C5.err; cut -c1-2500 gpurun_out/r2_bench_C5.json; tail -3 gpurun_out/r2_bench_C5.err
