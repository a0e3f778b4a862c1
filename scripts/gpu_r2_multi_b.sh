# Overlapped per-layer exchange (BP_PEER_OVERLAP) against the serial one, chained path, N GPUs.
N=${1:-2}
P=29500
mkdir -p gpurun_out
trun() { P=$((P+1)); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; }
for e in ; do
  echo "== dp_check p2p N=$N $e"
  ( export $e BP_DP=p2p; trun scripts/gpu_dp_check.py ) 2>&1 | grep -v "^W\|^\*\*\*\|^$\|OMP_NUM" | tail -5
done
for e in "BP_PEER_OVERLAP=0" "BP_PEER_OVERLAP=1"; do
  echo "== bench --gpus $N $e"
  tag=$(echo $e | tr ' =' '__')
  ( export $e; trun bench.py --gpus $N --steps 100 --warmup 10 --steady-seconds 0.5 ) 2> gpurun_out/r2_multib_n${N}_$tag.err | tee gpurun_out/r2_multib_n${N}_$tag.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in r.get('per_class_ms',{}).items()}, 'e2e', round(d['e2e']['value']), 'raw', round(d['e2e'].get('raw_reader',{}).get('value',0)), 'steady', d.get('steady',{}).get('ms_per_step'))
print('   exchange', d.get('exchange'), 'dp_parity', {k:v for k,v in (d.get('dp_parity') or {}).items() if k!='what'})
c=d.get('c4') or {}
print('   c4', {k:(round(v,4) if isinstance(v,float) else v) for k,v in c.items() if k in ('value','ms_per_step','bunch_per_gpu','exchange','error','tflops_per_gpu')}, {k:round(v,4) for k,v in (c.get('per_class_ms') or {}).items()}, {k:v for k,v in (c.get('dp_parity') or {}).items() if k!='what'})"
  grep -i "error\|timeout\|Traceback" gpurun_out/r2_multib_n${N}_$tag.err | head -5
done
