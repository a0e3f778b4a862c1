timeout 200 python scripts/gpu_edge_cases.py 2>&1 | tail -24
timeout 300 python scripts/gpu_fused_update_check.py 2>&1 | tail -20
# First GPU call of the next round: A/B (one box) of the switches prepared but not yet measured at the end of round 1.
#   BP_L2_PREFETCH=k   producer issues L2-only TMA prefetches k k-blocks ahead of its shared-memory ring (+ under PDL)
#   BP_DW_STREAM=1     gradient tiles stored with st.global.cs (evict-first)
#   BP_STAGES=2|3      128-wide pair kernels with a 2- / 3-deep ring (2: two CTAs per SM, fill/drain overlap)
#   fused update       scripts/gpu_fused_update_check.py: bit-parity of EPI_DW_SGD vs bp_sgd_kernel, then ms/bunch
#   BP_L2_PERSIST=MB   weight arena pinned in L2 through an access-policy window (host-only switch, rank_create)
# Each line: frames/s, ms per bunch, per-class ms.  Keep what wins by > 1 % twice.  Then the timeline of fwd vs dX.
bash scripts/gpu_ab.sh "BP_L2_PREFETCH=0" "BP_L2_PREFETCH=4" "BP_L2_PREFETCH=8" "BP_L2_PREFETCH=16" \
                       "BP_DW_STREAM=1" "BP_DW_STREAM=1 BP_L2_PREFETCH=8" "BP_STAGES=2" "BP_STAGES=2 BP_L2_PREFETCH=8" \
                       "BP_STAGES=3" "BP_L2_PERSIST=64" "BP_L2_PERSIST=96" "BP_L2_PERSIST=64 BP_FUSED_UPDATE=1" "BP_RELU_MASK=1" "BP_RELU_MASK=1 BP_FUSED_UPDATE=1" "BP_L2_PREFETCH=0"
for s in 0 2; do echo "== isolated, BP_STAGES=$s"; BP_STAGES=$s timeout 60 python scripts/gpu_mc_probe.py quick 2>&1 | tail -7; done
timeout 120 python scripts/gpu_pair_trace.py 2>&1 | head -150
echo "== dX gap: operand layout vs epilogue"; timeout 120 python scripts/gpu_mc_probe.py dxgap 2>&1 | tail -14
echo "== launch timeline of a bunch (default, then with the L2 switches)"; for e in "BP_X=0" "BP_L2_PREFETCH=8" "BP_L2_PERSIST=64" "BP_FUSED_UPDATE=1"; do echo "-- $e"; env $e timeout 120 python scripts/gpu_timeline.py C2 2>&1 | tail -22; done
echo "== C4 (512 frames per GPU, 6 weight layers): lone CTAs vs 128-wide pairs for small products"; for e in "BP_X=0" "BP_SMALL_PAIRS=1" "BP_X=0" "BP_SMALL_PAIRS=1"; do echo "-- $e"; env $e timeout 120 python bench.py --workload C4 --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['per_class_ms'].items()})"; done
