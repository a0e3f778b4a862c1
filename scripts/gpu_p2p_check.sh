# Peer-memory data parallelism on N GPUs (default 2): parity of both exchange modes, then an A/B of the bench.
N=${1:-2}
P=29600
tr() { P=$((P+1)); timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; }
echo "== dp_check p2p";  BP_DP=p2p BP_VERBOSE=1 tr scripts/gpu_dp_check.py 2>&1 | grep -v "^W\|^\*\*\*\|^$" | tail -8
echo "== dp_check nccl"; BP_DP=nccl tr scripts/gpu_dp_check.py 2>&1 | grep -v "^W\|^\*\*\*\|^$" | tail -5
echo "== group_check p2p"; BP_DP=p2p BP_VERBOSE=1 timeout 240 python scripts/gpu_group_check.py $N 2>&1 | tail -6
for m in nccl p2p nccl p2p; do
  echo "== bench --gpus $N BP_DP=$m"
  BP_DP=$m tr bench.py --gpus $N --steps 100 --warmup 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['per_class_ms'].items()}, 'e2e', round(d['e2e']['value']))"
done
