# A/B of kernel options inside ONE box (numbers from different boxes differ by several %).
# usage: scripts/gpu_ab.sh "ENV1=a ENV2=b" "ENV1=c" ...   (each argument = one run's environment)
run() { echo "== $1"; env $1 timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['per_class_ms'].items()}, 'e2e', round(d['e2e']['value']), d['clocks'])"; }
for e in "$@"; do run "$e"; done
