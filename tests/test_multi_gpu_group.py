"""In-process multi-GPU group (the reference's `gpu_used = N` constructor argument, BP_GPU.cu:20-36): one host thread per
GPU inside one process, peer-memory exchange (NCCL only when the devices cannot map each other).  Needs >= 2 GPUs: the
single-GPU round-end run skips it; scripts/gpu_r2_multi.sh runs it on the multi-GPU boxes."""
import ctypes as C
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        rt = C.CDLL("libcudart.so")
        n = C.c_int(0)
        return n.value if rt.cudaGetDeviceCount(C.byref(n)) == 0 else 0
    except OSError:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("dropout", [0, 1])
def test_group_of_two_gpus_computes_the_single_gpu_step(oracle, dropout):
    bp = importlib.import_module("dnn-for-speech-enhancement_b200")
    sizes, B, nb = [300, 512, 384, 129], 256, 3
    w, b = oracle.glorot_init(sizes, seed=3)
    x, t = oracle.synth_data(B * nb + 5, sizes[0], sizes[-1], seed=7)
    kw = (1, 0.1, 0.2) if dropout else (0, 0.0, 0.0)
    one = bp.BP_GPU(1, len(sizes), sizes, B, 1.0, 0.9, 1e-4, w, b, *kw, seed=77, device=0)
    one.train(x.shape[0], x, t)
    w1, b1 = one.returnWeights()
    cv1 = one.CrossValid(100, x[:100], t[:100])
    one.close()
    two = bp.BP_GPU(2, len(sizes), sizes, B, 1.0, 0.9, 1e-4, w, b, *kw, seed=77)   # gpu_used = 2
    assert two.get_option("dp_exchange") in (1, 2)
    two.train(x.shape[0], x, t)
    w2, b2 = two.returnWeights()
    cv2 = two.CrossValid(100, x[:100], t[:100])
    two.close()
    for l in range(1, len(sizes)):
        d = np.linalg.norm((w2[l] - w1[l]).astype(np.float64)) / np.linalg.norm((w1[l] - w[l]).astype(np.float64))
        assert d <= 2e-3, f"layer {l}: ||W_2gpu - W_1gpu|| / ||dW|| = {d:.3e}"
    assert abs(cv2 - cv1) <= 1e-3 * abs(cv1)
