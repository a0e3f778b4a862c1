# Side-car update / two-slice first-layer gradient: bit parity (tests), then in-process A/B on C2, C3 and C4's per-GPU share.
mkdir -p gpurun_out
echo "=== parity"; timeout 400 python -m pytest tests/test_pipeline_edges_chain.py -q -m gpu -k "sidecar" 2>&1 | tail -6
echo "=== A/B C2"
timeout 200 python scripts/gpu_ab_quick.py C2 "sgd_defer=0" "sgd_defer=2" "sgd_defer=3" "sgd_defer=4" "sgd_defer=2 sgd_defer_at=1" "sgd_defer=3 sgd_defer_at=1" \
  "sgd_defer=2 sgd_defer_tpcs=6" "sgd_defer=2 sgd_defer_tpcs=8" "sgd_defer=2 sgd_defer_tpcs=12" "sgd_defer=2 sgd_defer_tpcs=16" \
  "dw1_split=8" "dw1_split=9" "dw1_split=6" "sgd_defer=2 dw1_split=8" "sgd_defer=2 dw1_split=6" "sgd_defer=3 dw1_split=8" 2>&1 | grep -v "^this bunch" | tail -34
echo "=== A/B C3"
timeout 200 python scripts/gpu_ab_quick.py C3 "sgd_defer=0" "sgd_defer=2" "sgd_defer=3" "dw1_split=8" "sgd_defer=2 dw1_split=8" "sgd_defer=2 dw1_split=10" 2>&1 | grep -v "^this bunch" | tail -14
echo "=== A/B C4 (one GPU's share)"
timeout 200 python scripts/gpu_ab_quick.py C4 "sgd_defer=0" "sgd_defer=2" "sgd_defer=3" "sgd_defer=4" "sgd_defer=2 dw1_split=8" 2>&1 | grep -v "^this bunch" | tail -12
