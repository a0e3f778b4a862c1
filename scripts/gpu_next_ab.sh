# First GPU call of the next round: A/B (one box) of the switches prepared but not yet measured at the end of round 1.
#   BP_L2_PREFETCH=k   producer issues L2-only TMA prefetches k k-blocks ahead of its shared-memory ring (+ under PDL)
#   BP_DW_STREAM=1     gradient tiles stored with st.global.cs (evict-first)
# Each line: frames/s, ms per bunch, per-class ms.  Keep what wins by > 1 % twice.
bash scripts/gpu_ab.sh "BP_L2_PREFETCH=0" "BP_L2_PREFETCH=4" "BP_L2_PREFETCH=8" "BP_L2_PREFETCH=16" \
                       "BP_DW_STREAM=1" "BP_DW_STREAM=1 BP_L2_PREFETCH=8" "BP_L2_PREFETCH=0"
