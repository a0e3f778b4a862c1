"""In-process multi-GPU group (bp_create with gpu_used = N, the path `BPtrain gpu_used=N` takes): N host threads + NCCL
ranks inside one process must reproduce the single-GPU result.    python scripts/gpu_group_check.py [N]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
bp = importlib.import_module("dnn-for-speech-enhancement_b200")
import oracle_py as O  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    sizes, B, nb = [300, 512, 384, 129], 256, 3
    w, b = O.glorot_init(sizes, seed=3)
    x, t = O.synth_data(B * nb + 11, sizes[0], sizes[-1], seed=7)
    g = bp.BP_GPU(n, len(sizes), sizes, B, 1.0, 0.9, 1e-4, w, b, 1, 0.1, 0.2, seed=77)     # group of n GPUs
    g.train(x.shape[0], x, t)
    gw, gb = g.returnWeights()
    cv = g.CrossValid(100, x[:100], t[:100])
    g.close()
    s = bp.BP_GPU(1, len(sizes), sizes, B, 1.0, 0.9, 1e-4, w, b, 1, 0.1, 0.2, seed=77, device=0)
    s.train(x.shape[0], x, t)
    sw, sb = s.returnWeights()
    cvs = s.CrossValid(100, x[:100], t[:100])
    s.close()
    ok = True
    for l in range(1, len(sizes)):
        d = float(np.linalg.norm((gw[l] - sw[l]).astype(np.float64)))
        nrm = float(np.linalg.norm((sw[l] - w[l]).astype(np.float64)))
        print(f"layer {l}: ||W_group - W_single|| / ||dW_single|| = {d / nrm:.3e}")
        ok &= d <= 2e-3 * nrm
    ok &= abs(cv - cvs) <= 1e-3 * abs(cvs)
    print("GROUP CHECK", "OK" if ok else "FAILED", cv, cvs)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
