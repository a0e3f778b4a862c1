# Round 2, call 2: compute-sanitizer over the hot path + in-process A/B of every scheduling switch + traces.
mkdir -p gpurun_out
echo "=== A/B C2"; timeout 300 python scripts/gpu_ab_inproc.py C2 2>&1 | tail -22
echo "=== A/B C3"; timeout 300 python scripts/gpu_ab_inproc.py C3 2>&1 | tail -22
echo "=== A/B C4 (1 rank, 512 frames)"; timeout 300 python scripts/gpu_ab_inproc.py C4 2>&1 | tail -22
echo "=== pair trace"; timeout 120 python scripts/gpu_pair_trace.py 2>&1 | head -150
echo "=== isolated"; timeout 60 python scripts/gpu_mc_probe.py quick 2>&1 | tail -9
echo "=== dxgap"; timeout 120 python scripts/gpu_mc_probe.py dxgap 2>&1 | tail -14
echo "=== sanitizer"; bash scripts/gpu_sanitize.sh 2>&1 | tee gpurun_out/r2_sanitizer.log | tail -60
