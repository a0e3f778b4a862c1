# N-GPU check of the peer-memory exchange: parity against the 1-rank run, then the bench (one box).
N=${1:-4}
P=29800
mkdir -p gpurun_out
tr() { P=$((P+1)); timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; }
echo "== dp_check p2p N=$N";  BP_DP=p2p BP_VERBOSE=1 tr scripts/gpu_dp_check.py 2>&1 | grep -v "^W\|^\*\*\*\|^$\|OMP_NUM" | tail -8
echo "== bench --gpus $N (default exchange)"
BP_VERBOSE=1 tr bench.py --gpus $N --steps 100 --warmup 10 2> gpurun_out/p2p_n$N.err | tee gpurun_out/bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['per_class_ms'].items()}, 'e2e', round(d['e2e']['value']))"
grep -i "libbpgpu: grad\|error\|timeout" gpurun_out/p2p_n$N.err | head -5
