"""Which property of a host <-> device copy slows the C2 bunches training beside it (end to end)?
Device ms per bunch with: nothing beside it; the padded-row copy of bp_train (2-D, 93 MB); the same bytes as one flat
block; the flat block into a 2 MB destination over and over (stays in L2); a device -> host copy of 93 MB.
usage: python scripts/gpu_copy_interference_probe.py"""
import ctypes as C
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

bp = importlib.import_module("dnn-for-speech-enhancement_b200")
rt = C.CDLL("libcudart.so")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaStreamCreateWithFlags.argtypes = [C.POINTER(C.c_void_p), C.c_uint]
rt.cudaStreamSynchronize.argtypes = [C.c_void_p]


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
ck(rt.cudaStreamCreateWithFlags(C.byref(stream), 1))
src = C.c_void_p(px.array.ctypes.data)
flat = rows * k0 * 4
small = 2 << 20
H2D, D2H = 1, 2
kinds = {
    "nothing": lambda: None,
    "2-D rows into the padded layout (bp_train)": lambda: ck(rt.cudaMemcpy2DAsync(dev, ld * 4, src, k0 * 4, k0 * 4, rows, H2D, stream)),
    "one flat block, same bytes": lambda: ck(rt.cudaMemcpyAsync(dev, src, flat, H2D, stream)),
    "flat, 2 MB destination reused (46 copies each)": lambda: [ck(rt.cudaMemcpyAsync(dev, src, small, H2D, stream)) for _ in range(46)],
    "device -> host, flat, same bytes": lambda: ck(rt.cudaMemcpyAsync(src, dev, flat, D2H, stream)),
}
for rep in range(2):
    for name, issue in kinds.items():
        g.sync()
        for _ in range(40):
            issue()
        g.timer_start()
        for _ in range(4):
            g.train_resident(0, cb)
        ms = g.timer_stop() / (4 * cb)
        ck(rt.cudaStreamSynchronize(stream))
        print(f"{name:52s} {ms:.4f} ms per bunch")
g.close()
