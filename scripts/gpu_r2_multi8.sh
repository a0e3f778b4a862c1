# Round 2, 8-GPU call: bench line (C2 weak scaling + C4 sub-line + dp_parity) for both schedules, NVLink counters around it.
N=${1:-8}
P=29600
mkdir -p gpurun_out
trun() { P=$((P+1)); timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; }
nvidia-smi topo -m 2>&1 | head -12
for e in "BP_CHAIN=1" "BP_CHAIN=0"; do
  echo "== bench --gpus $N $e"
  tag=$(echo $e | tr ' =' '__')
  nvidia-smi nvlink -gt d -i 0 2>&1 | grep -i "data\|link 0" | head -4 > gpurun_out/r2_nvlink_before_$tag.txt
  ( export $e; trun bench.py --gpus $N --steps 100 --warmup 10 --steady-seconds 1 ) 2> gpurun_out/r2_multi_n${N}_$tag.err | tee gpurun_out/r2_multi_n${N}_$tag.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in r.get('per_class_ms',{}).items()}, 'e2e', round(d['e2e']['value']), 'raw', round(d['e2e'].get('raw_reader',{}).get('value',0)), 'steady', d.get('steady',{}).get('ms_per_step'))
print('   exchange', d.get('exchange'), 'dp_parity', {k:v for k,v in (d.get('dp_parity') or {}).items() if k!='what'})
c=d.get('c4') or {}
print('   c4', {k:(round(v,4) if isinstance(v,float) else v) for k,v in c.items() if k in ('value','ms_per_step','bunch_per_gpu','exchange','error','tflops_per_gpu')}, {k:round(v,4) for k,v in (c.get('per_class_ms') or {}).items()}, {k:v for k,v in (c.get('dp_parity') or {}).items() if k!='what'})"
  nvidia-smi nvlink -gt d -i 0 2>&1 | grep -i "data\|link 0" | head -4 > gpurun_out/r2_nvlink_after_$tag.txt
  grep -i "error\|timeout\|Traceback" gpurun_out/r2_multi_n${N}_$tag.err | head -5
done
echo "== nvlink counters GPU 0 (before / after the BP_CHAIN=1 run)"; cat gpurun_out/r2_nvlink_before_BP_CHAIN_1.txt gpurun_out/r2_nvlink_after_BP_CHAIN_1.txt
