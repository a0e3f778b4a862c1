#!/bin/bash
# Builds the REFERENCE's own CUDA implementation of the hot path (BP_GPU.cu + DevFunc.cu, cuBLAS FP32) for sm_100a,
# straight from the read-only sources under /root/reference, into oracle/_ref/ (git-ignored, travels with gpurun).
# TEST INFRASTRUCTURE ONLY: used by tests/test_reference_parity.py and bench.py's reference-GPU row as the checker /
# the thing compared against — never linked into or called by the product.
#
# Nothing from the reference is copied into the repo: objects are compiled from the sources where they lie; the one
# patched translation unit (P1, below) is produced by `sed` into a scratch directory outside the repo.
#   P1  DevFunc.cu:67,81 read `global void` instead of `__global__ void` (HEAD does not compile as shipped).
# Outputs:
#   oracle/_ref/BPtrain_ref    the reference CLI (HEAD semantics: ReLU, NAT block with the literal 129)
#   oracle/_ref/ref_reader_dump  tests/native/reader_dump.cc linked against the reference's Interface.o (host only)
#   oracle/_ref/ref_weights_dump  tests/native/weights_dump.cc linked against the reference's Interface.o (host only)
#   oracle/_ref/ref_harness_shim  oracle/ref_harness.cc + host/BP_GPU_shim.cc + OUR libbpgpu.so (same driver, our trainer)
#   oracle/_ref/BPtrain_shim  the reference's BPtrain.cc + Interface.cc + host/BP_GPU_shim.cc + OUR libbpgpu.so
#   oracle/_ref/ref_harness    oracle/ref_harness.cc (ours) linked against the reference's BP_GPU.o / DevFunc.o:
#                              drives class BP_GPU directly on binary blobs (no Pfile plumbing) and times train().
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
[ -d "$REF" ] || { echo "build_ref: $REF not present (GPU box?) - keeping prebuilt files"; exit 0; }
TMP=$(mktemp -d /tmp/bpref.XXXXXX)
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT"
sed 's/^global void/__global__ void/' "$REF/DevFunc.cu" > "$TMP/DevFunc_p1.cu"
$NVCC $ARCH -O2 -w -I"$REF" -c "$TMP/DevFunc_p1.cu" -o "$TMP/DevFunc.o"
$NVCC $ARCH -O2 -w -I"$REF" -c "$REF/BP_GPU.cu" -o "$TMP/BP_GPU.o"
g++ -O2 -w -I/usr/local/cuda/include -I"$REF" -c "$REF/Interface.cc" -o "$TMP/Interface.o"
$NVCC $ARCH -O2 -w -I"$REF" "$REF/BPtrain.cc" "$TMP/BP_GPU.o" "$TMP/DevFunc.o" "$TMP/Interface.o" \
      -o "$OUT/BPtrain_ref" -L/usr/local/cuda/lib64 -lcublas -lcurand -Xlinker -rpath=/usr/local/cuda/lib64
$NVCC $ARCH -O2 -w -I"$REF" "$HERE/ref_harness.cc" "$TMP/BP_GPU.o" "$TMP/DevFunc.o" \
      -o "$OUT/ref_harness" -L/usr/local/cuda/lib64 -lcublas -lcurand -Xlinker -rpath=/usr/local/cuda/lib64
# the reference's host-only reader (Interface.cc) behind tests/native/reader_dump.cc: runs without a GPU
g++ -O2 -w -I/usr/local/cuda/include -I"$REF" "$HERE/../tests/native/reader_dump.cc" "$TMP/Interface.o" \
      -o "$OUT/ref_reader_dump"
g++ -O2 -w -I/usr/local/cuda/include -I"$REF" "$HERE/../tests/native/weights_dump.cc" "$TMP/Interface.o" \
      -o "$OUT/ref_weights_dump"
# the reference's unmodified main() and Interface with OUR trainer behind its class BP_GPU: the
# binding dnn-for-speech-enhancement_b200/host/BP_GPU_shim.cc compiled against the reference's BP_GPU.h, linked with
# libbpgpu.so instead of BP_GPU.o / DevFunc.o / cuBLAS / cuRAND.  Proves the C-ABI covers the reference's call sites.
LIB="$HERE/../dnn-for-speech-enhancement_b200/lib"
if [ -f "$LIB/libbpgpu.so" ]; then
  g++ -O2 -w -I/usr/local/cuda/include -I"$REF" -c "$REF/BPtrain.cc" -o "$TMP/BPtrain_host.o"
  g++ -O2 -w -I/usr/local/cuda/include -I"$REF" -I"$HERE/../include" \
      -c "$HERE/../dnn-for-speech-enhancement_b200/host/BP_GPU_shim.cc" -o "$TMP/BP_GPU_shim.o"
  g++ -o "$OUT/BPtrain_shim" "$TMP/BPtrain_host.o" "$TMP/Interface.o" "$TMP/BP_GPU_shim.o" \
      -L"$LIB" -lbpgpu -Wl,-rpath,'$ORIGIN/../../dnn-for-speech-enhancement_b200/lib'
  # ... and the blob harness over the same binding: class BP_GPU driven exactly like ref_harness drives the reference's,
  # so that scripts/gpu_ref_gpu_timing.py times both implementations through one and the same C++ interface
  g++ -O2 -w -I/usr/local/cuda/include -I"$REF" -c "$HERE/ref_harness.cc" -o "$TMP/ref_harness_host.o"
  g++ -o "$OUT/ref_harness_shim" "$TMP/ref_harness_host.o" "$TMP/BP_GPU_shim.o" -L"$LIB" -lbpgpu \
      -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN/../../dnn-for-speech-enhancement_b200/lib' \
      -Wl,-rpath,/usr/local/cuda/lib64
fi
echo "built: $(ls "$OUT")"
