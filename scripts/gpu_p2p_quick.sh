# 2-GPU bench in both exchange modes (one box).  Parity of both modes: scripts/gpu_dp_check.py / gpu_p2p_check.sh.
N=${1:-2}
P=29700
mkdir -p gpurun_out
tr() { P=$((P+1)); timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; }
for m in p2p nccl; do
  echo "== bench --gpus $N BP_DP=$m"
  BP_DP=$m BP_VERBOSE=1 tr bench.py --gpus $N --steps 100 --warmup 10 2> gpurun_out/p2p_$m.err | tee gpurun_out/p2p_bench_$m.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['per_class_ms'].items()}, 'e2e', round(d['e2e']['value']))"
  grep -i "libbpgpu\|error\|timeout" gpurun_out/p2p_$m.err | head -5
done
