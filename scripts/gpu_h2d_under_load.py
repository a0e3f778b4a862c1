"""Host -> device copy rate of one `bp_train` call's input rows (8192 x 2827 floats into the padded layout) while the GPU
is idle and while C2 bunches train beside the copy (the situation of bench.py's e2e loop).  CUDA-event timed on the copy's
own stream.   usage: python scripts/gpu_h2d_under_load.py"""
import ctypes as C
import importlib
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

bp = importlib.import_module("dnn-for-speech-enhancement_b200")
rt = C.CDLL("libcudart.so")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
rt.cudaEventRecord.argtypes = [C.c_void_p, C.c_void_p]
rt.cudaEventSynchronize.argtypes = [C.c_void_p]
rt.cudaStreamCreateWithFlags.argtypes = [C.POINTER(C.c_void_p), C.c_uint]


def ck(e):
    assert e == 0, f"CUDA error {e}"


sizes, lb, cb = [2827, 2048, 2048, 2048, 257], 1024, 32
w, b = bench.glorot(sizes)
g = bp.BP_GPU(1, len(sizes), sizes, lb, 1.0, 0.9, 0.0, w, b, 0, 0.0, 0.0, seed=1, device=0)
px, pt = bp.PinnedArray((cb * lb, sizes[0])), bp.PinnedArray((cb * lb, sizes[-1]))
bench.synth(cb * lb, sizes[0], sizes[-1], seed=100, out_x=px.array, out_t=pt.array)
g.upload_chunk(cb * lb, px.array, pt.array)
g.train_resident(0, cb)
g.sync()

rows, k0, ld = 8192, 2827, 2848
dev, stream = C.c_void_p(), C.c_void_p()
ck(rt.cudaSetDevice(0))
ck(rt.cudaMalloc(C.byref(dev), C.c_size_t(rows * ld * 4)))
ck(rt.cudaStreamCreateWithFlags(C.byref(stream), 1))   # non-blocking
n = 16
ev = [C.c_void_p() for _ in range(n + 1)]
for e in ev:
    ck(rt.cudaEventCreate(C.byref(e)))
src = C.c_void_p(px.array.ctypes.data)


def copies(label):
    ck(rt.cudaEventRecord(ev[0], stream))
    for i in range(n):
        ck(rt.cudaMemcpy2DAsync(dev, ld * 4, src, k0 * 4, k0 * 4, rows, 1, stream))
        ck(rt.cudaEventRecord(ev[i + 1], stream))
    ck(rt.cudaEventSynchronize(ev[n]))
    gbs = []
    for i in range(n):
        t = C.c_float()
        ck(rt.cudaEventElapsedTime(C.byref(t), ev[i], ev[i + 1]))
        gbs.append(rows * k0 * 4 / t.value / 1e6)
    print(f"{label:44s} median {statistics.median(gbs):5.1f} GB/s   min {min(gbs):5.1f}   max {max(gbs):5.1f}")


copies("GPU idle")
for rep in range(2):
    for _ in range(8):          # ~60 ms of queued training; the 16 copies take ~30 ms
        g.train_resident(0, cb)
    copies("C2 bunches training beside the copy")
    g.sync()
copies("GPU idle again")
g.close()
