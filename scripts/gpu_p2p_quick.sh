# Quick 2-GPU validation of the peer-memory exchange: parity vs the 1-rank run, then bench p2p vs nccl (one box).
N=${1:-2}
P=29700
mkdir -p gpurun_out
tr() { P=$((P+1)); timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; }
echo "== dp_check p2p";  BP_DP=p2p BP_VERBOSE=1 tr scripts/gpu_dp_check.py 2>&1 | grep -v "^W\|^\*\*\*\|^$" | tail -8
for m in p2p nccl; do
  echo "== bench --gpus $N BP_DP=$m"
  BP_DP=$m BP_VERBOSE=1 tr bench.py --gpus $N --steps 100 --warmup 10 2> gpurun_out/p2p_$m.err | tee gpurun_out/p2p_bench_$m.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['per_class_ms'].items()}, 'e2e', round(d['e2e']['value']))"
  grep -i "libbpgpu\|error\|timeout" gpurun_out/p2p_$m.err | head -5
done
