"""Host -> device copy rates behind `bp_train`'s end-to-end figure: one C2 chunk of 8 bunches (8192 x 2827 floats, pinned) copied
(a) as one flat block, (b) row by row into the padded device layout (`cudaMemcpy2DAsync`, pitch 2848 floats — what
`rank_upload` issues), (c) as (b) plus the flat target block behind it, each alone on an idle GPU.  CUDA-event timed, best and
median of 20.   usage: python scripts/gpu_h2d_probe.py"""
import ctypes as C
import statistics

rt = C.CDLL("libcudart.so")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
rt.cudaEventRecord.argtypes = [C.c_void_p, C.c_void_p]
rt.cudaEventSynchronize.argtypes = [C.c_void_p]


def ck(e):
    assert e == 0, f"CUDA error {e}"


rows, k0, ld, nout = 8192, 2827, 2848, 257
host, htarg, dev, dtarg, stream, e0, e1 = (C.c_void_p() for _ in range(7))
ck(rt.cudaSetDevice(0))
ck(rt.cudaMallocHost(C.byref(host), C.c_size_t(rows * k0 * 4)))
ck(rt.cudaMallocHost(C.byref(htarg), C.c_size_t(rows * nout * 4)))
ck(rt.cudaMalloc(C.byref(dev), C.c_size_t(rows * ld * 4)))
ck(rt.cudaMalloc(C.byref(dtarg), C.c_size_t(rows * nout * 4)))
C.memset(host, 1, rows * k0 * 4)
C.memset(htarg, 1, rows * nout * 4)
ck(rt.cudaStreamCreate(C.byref(stream)))
ck(rt.cudaEventCreate(C.byref(e0)))
ck(rt.cudaEventCreate(C.byref(e1)))
H2D = 1


def timed(fn, nbytes, label):
    ms = []
    for i in range(23):
        ck(rt.cudaEventRecord(e0, stream))
        fn()
        ck(rt.cudaEventRecord(e1, stream))
        ck(rt.cudaEventSynchronize(e1))
        t = C.c_float()
        ck(rt.cudaEventElapsedTime(C.byref(t), e0, e1))
        if i >= 3:
            ms.append(t.value)
    print(f"{label:58s} {nbytes / 1e6:7.1f} MB  best {nbytes / min(ms) / 1e6:6.1f} GB/s  median {nbytes / statistics.median(ms) / 1e6:6.1f} GB/s")


timed(lambda: ck(rt.cudaMemcpyAsync(dev, host, rows * k0 * 4, H2D, stream)), rows * k0 * 4, "flat block (cudaMemcpyAsync)")
timed(lambda: ck(rt.cudaMemcpy2DAsync(dev, ld * 4, host, k0 * 4, k0 * 4, rows, H2D, stream)), rows * k0 * 4,
      "rows into the padded layout (cudaMemcpy2DAsync)")


def both():
    ck(rt.cudaMemcpy2DAsync(dev, ld * 4, host, k0 * 4, k0 * 4, rows, H2D, stream))
    ck(rt.cudaMemcpyAsync(dtarg, htarg, rows * nout * 4, H2D, stream))


timed(both, rows * (k0 + nout) * 4, "padded rows + flat targets (one bp_train call's copies)")
