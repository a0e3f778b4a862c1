# One-call measurement sweep for BASELINE.md (1 GPU).
mkdir -p gpurun_out
for w in C2 C3 C5; do
  echo "== $w"; timeout 300 python bench.py --workload $w --steps 100 --warmup 10 > gpurun_out/sum_$w.json 2> gpurun_out/sum_$w.err; cut -c1-1800 gpurun_out/sum_$w.json; tail -2 gpurun_out/sum_$w.err
done
echo "== C2 3xtf32"; timeout 300 python bench.py --math 3xtf32 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/sum_C2_3x.json 2>/dev/null; cut -c1-700 gpurun_out/sum_C2_3x.json
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/sum_ref.json 2>/dev/null; cut -c1-900 gpurun_out/sum_ref.json
echo "== REF-GPU"; timeout 600 python scripts/gpu_ref_gpu_timing.py 16 2>&1 | tail -3
