for cs in 1x1 2x1 1x2 2x2 4x1 1x4; do
  echo "== cluster $cs"
  BP_CLUSTER=$cs timeout 120 python tests/gpu_probe_gemm.py 2>&1 | grep -E "FAIL|ALL OK|SOME|2048 N= 1024 K= 2048|rc=" | cut -c1-160
  BP_CLUSTER=$cs timeout 200 python bench.py --steps 60 --warmup 10 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('value',round(d['value']), 'ms/step', round(d['ms_per_step'],4), d['roofline']['per_class_ms'], 'frac', round(d['roofline']['frac'],3))
    elif l: print(l[:200])
"
done
