# Multi-GPU check (one box, N GPUs): in-process group test, dp_check, and the bench line with dp_parity / exchange / the C4
# sub-line for the default schedule (and, with a second argument, the per-product launches beside it).
# usage: gpurun --gpus N -- bash scripts/gpu_r2_multi.sh N [ab]
N=${1:-2}
P=29700
mkdir -p gpurun_out
trun() { P=$((P+1)); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; }
echo "== in-process group (gpu_used=2)"; timeout 200 python -m pytest tests/test_multi_gpu_group.py -q -m gpu 2>&1 | tail -3
echo "== dp_check p2p N=$N"; ( export BP_DP=p2p; trun scripts/gpu_dp_check.py ) 2>&1 | grep -v "^W\|^\*\*\*\|^$\|OMP_NUM" | tail -5
VARIANTS=("BP_X=0"); [ -n "$2" ] && VARIANTS=("BP_X=0" "BP_CHAIN=0")
for e in "${VARIANTS[@]}"; do
  echo "== bench --gpus $N $e"
  tag=$(echo $e | tr ' =' '__')
  ( export $e; trun bench.py --gpus $N --steps 100 --warmup 10 --steady-seconds 0.5 ) 2> gpurun_out/r2_multi_n${N}_$tag.err | tee gpurun_out/r2_multi_n${N}_$tag.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in r.get('per_class_ms',{}).items()}, 'e2e', round(d['e2e']['value']), 'raw', round(d['e2e'].get('raw_reader',{}).get('value',0)))
print('   exchange', d.get('exchange'), d.get('schedule'), 'dp_parity', {k:v for k,v in (d.get('dp_parity') or {}).items() if k!='what'})
c=d.get('c4') or {}
print('   c4', {k:(round(v,4) if isinstance(v,float) else v) for k,v in c.items() if k in ('value','ms_per_step','bunch_per_gpu','exchange','error','tflops_per_gpu')}, {k:v for k,v in (c.get('dp_parity') or {}).items() if k!='what'})"
  grep -i "error\|timeout\|Traceback" gpurun_out/r2_multi_n${N}_$tag.err | head -5
done
