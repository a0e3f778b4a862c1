# Round 2 multi-GPU call: parity of the data-parallel step (both schedules, both exchange variants) and the bench line
# with dp_parity / exchange / C4 sub-line.   usage: gpurun --gpus N -- bash scripts/gpu_r2_multi.sh N
N=${1:-2}
P=29700
mkdir -p gpurun_out
trun() { P=$((P+1)); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; }
for e in "BP_CHAIN=1" "BP_CHAIN=0" "BP_CHAIN=0 BP_PEER_EARLY=1"; do
  echo "== dp_check p2p N=$N $e"
  ( export $e BP_DP=p2p; trun scripts/gpu_dp_check.py ) 2>&1 | grep -v "^W\|^\*\*\*\|^$\|OMP_NUM" | tail -5
done
for e in "BP_CHAIN=0" "BP_CHAIN=1" "BP_CHAIN=0 BP_PEER_EARLY=1" "BP_CHAIN=1"; do
  echo "== bench --gpus $N $e"
  tag=$(echo $e | tr ' =' '__')
  ( export $e; trun bench.py --gpus $N --steps 100 --warmup 10 --steady-seconds 0.5 ) 2> gpurun_out/r2_multi_n${N}_$tag.err | tee gpurun_out/r2_multi_n${N}_$tag.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in r.get('per_class_ms',{}).items()}, 'e2e', round(d['e2e']['value']), 'raw', round(d['e2e'].get('raw_reader',{}).get('value',0)))
print('   exchange', d.get('exchange'), 'peer_early', d.get('peer_early'), 'dp_parity', d.get('dp_parity'))
c=d.get('c4') or {}
print('   c4', {k:(round(v,4) if isinstance(v,float) else v) for k,v in c.items() if k in ('value','ms_per_step','bunch_per_gpu','exchange','error','tflops_per_gpu')}, c.get('per_class_ms'), c.get('dp_parity'))"
  grep -i "error\|timeout\|Traceback" gpurun_out/r2_multi_n${N}_$tag.err | head -5
done
