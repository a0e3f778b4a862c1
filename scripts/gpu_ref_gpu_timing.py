"""REF-GPU row of BASELINE.md: the reference's own CUDA path (class BP_GPU compiled unmodified from /root/reference into
oracle/_ref/ref_harness; cuBLAS FP32 + its element-wise kernels) timed on this B200 at the C2 shape.
The time is the wall time of BP_GPU::train() on one chunk (device-synchronised, includes the reference's own H2D of
the chunk from pageable memory), second repetition.   python scripts/gpu_ref_gpu_timing.py [bunches]"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    sizes, bunch = [2827, 2048, 2048, 2048, 257], 1024
    n = nb * bunch
    rng = np.random.default_rng(0)
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.blob"), os.path.join(d, "out.blob")
        with open(fin, "wb") as f:
            f.write(struct.pack(f"<i{len(sizes)}i5i", len(sizes), *sizes, bunch, n, 0, 0, 2))
            f.write(struct.pack("<5f", 1.0, 0.9, 0.0, 0.0, 0.0))
            for l in range(1, len(sizes)):
                r = 0.5 * np.sqrt(6.0) / np.sqrt(sizes[l - 1] + sizes[l])
                f.write(rng.uniform(-r, r, size=(sizes[l - 1], sizes[l])).astype("<f4").tobytes())
                f.write(np.zeros(sizes[l], "<f4").tobytes())
            f.write(rng.standard_normal((n, sizes[0]), dtype=np.float32).tobytes())
            f.write((0.5 * rng.standard_normal((n, sizes[-1]), dtype=np.float32)).tobytes())
        r = subprocess.run([H, fin, fout], capture_output=True, text=True, timeout=900, cwd=d)
        print(r.stdout.strip(), r.stderr.strip()[-300:])
        raw = np.fromfile(fout, dtype="<f4")
        ms = float(raw[-1])
        print(f"REF-GPU (reference CUDA path, cuBLAS FP32, B200): {n} frames in {ms:.2f} ms -> {n / (ms * 1e-3):.0f} "
              f"frames/s ({ms / nb * 1e3:.1f} us per bunch of {bunch}, H2D of the chunk included)")
        # the same driver, the same class interface, the same pageable host buffers — our trainer behind the binding
        # of INTEGRATION.md B (oracle/_ref/ref_harness_shim)
        hs = H + "_shim"
        if os.path.exists(hs):
            r = subprocess.run([hs, fin, fout], capture_output=True, text=True, timeout=900, cwd=d)
            print(r.stdout.strip(), r.stderr.strip()[-300:])
            ours = np.fromfile(fout, dtype="<f4")
            ms2 = float(ours[-1])
            rel = float(np.linalg.norm(ours[:-2].astype(np.float64) - raw[:-2]) / np.linalg.norm(raw[:-2]))
            print(f"OURS through class BP_GPU (BP_GPU_shim.cc, TF32): {n} frames in {ms2:.2f} ms -> "
                  f"{n / (ms2 * 1e-3):.0f} frames/s ({ms2 / nb * 1e3:.1f} us per bunch, pageable H2D included): "
                  f"{ms / ms2:.1f}x; weights after training differ by {rel:.2e} (relative Frobenius norm)")


if __name__ == "__main__":
    main()
