# BASELINE configs[4] on N GPUs of one box: C5 forward decode, N independent replicas (one process per GPU).
# usage: gpurun --gpus N -- bash scripts/gpu_r2_c5_multi.sh N
N=${1:-8}
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733 \
  bench.py --gpus $N --workload C5 --steps 40 --warmup 5 --chunk-bunches 8 --steady-seconds 0.5 --no-cpu-baseline \
  2> gpurun_out/r2k_c5_n$N.err | tee gpurun_out/r2k_c5_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('C5 N=%d' % d['n_gpus'], round(d['value']), 'frames/s', round(d['ms_per_step'], 4), 'ms per batch and GPU', d['config']['parallelism'])
print('  steady', round(d['steady']['value']), d['steady'].get('clocks'))
print('  e2e', {k: (round(v['value']) if isinstance(v, dict) and 'value' in v else v) for k, v in d['e2e'].items() if k in ('value', 'pipelined', 'raw_reader', 'raw_reader_pipelined')})
print('  clocks', d['clocks'])"
grep -i "error\|Traceback\|timeout" gpurun_out/r2k_c5_n$N.err | head -5
