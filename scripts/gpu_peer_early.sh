# Round-2 call at N GPUs (default 2): the two-part peer-memory exchange (BP_PEER_EARLY=1) against the default one.
#   1. parity: N ranks vs the 1-rank run, replicas bit-identical (scripts/gpu_dp_check.py), both settings
#   2. bench --gpus N, both settings, one box
N=${1:-2}
P=29900
mkdir -p gpurun_out
tr() { P=$((P+1)); timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; }
for e in 0 1; do
  echo "== dp_check p2p N=$N BP_PEER_EARLY=$e"
  BP_PEER_EARLY=$e BP_DP=p2p BP_VERBOSE=1 tr scripts/gpu_dp_check.py 2>&1 | grep -v "^W\|^\*\*\*\|^$\|OMP_NUM" | tail -6
done
for e in 0 1 0 1; do
  echo "== bench --gpus $N BP_PEER_EARLY=$e"
  BP_PEER_EARLY=$e BP_DP=p2p tr bench.py --gpus $N --steps 100 --warmup 10 2> gpurun_out/peer_early_$e.err | tee gpurun_out/peer_early_$e.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['per_class_ms'].items()}, 'e2e', round(d['e2e']['value']))"
  grep -i "error\|timeout" gpurun_out/peer_early_$e.err | head -5
done
