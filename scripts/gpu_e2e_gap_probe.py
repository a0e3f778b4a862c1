"""Where bench.py's e2e loop loses its time: (1) device time per C2 bunch with and without host -> device copies streaming
beside it, (2) the host-side duration of every bp_train() call (returns when its copy has landed) and of the
bp_train_losses() call that follows, over 24 calls of 8 bunches.   usage: python scripts/gpu_e2e_gap_probe.py"""
import ctypes as C
import importlib
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

bp = importlib.import_module("dnn-for-speech-enhancement_b200")
rt = C.CDLL("libcudart.so")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
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


def bunches(with_copies):
    g.sync()
    if with_copies:
        for _ in range(40):      # ~73 ms of copies; 4 x 32 bunches take ~31 ms
            ck(rt.cudaMemcpy2DAsync(dev, ld * 4, src, k0 * 4, k0 * 4, rows, 1, stream))
    g.timer_start()
    for _ in range(4):
        g.train_resident(0, cb)
    ms = g.timer_stop() / (4 * cb)
    ck(rt.cudaStreamSynchronize(stream))
    return ms


e2e_cb = 8
n = e2e_cb * lb
for chain in (0, 1) if "chain" in sys.argv[1:] else (0,):
  g.set_option("chain", chain)
  g.upload_chunk(cb * lb, px.array, pt.array)   # the e2e loop below leaves an 8-bunch chunk resident
  g.train_resident(0, cb)
  print(f"--- chain={chain}")
  for rep in range(2):
    print(f"device ms per bunch: GPU to itself {bunches(False):.4f}   copies streaming beside it {bunches(True):.4f}")
  for rep in range(2):
      g.train(n, px.array[:n], pt.array[:n])
      g.sync()
      t_train, t_loss = [], []
      t_begin = time.perf_counter()
      for c in range(24):
          o = (c % (cb // e2e_cb)) * n
          t0 = time.perf_counter()
          g.train(n, px.array[o: o + n], pt.array[o: o + n])
          t1 = time.perf_counter()
          if c > 0:
              g.train_losses(age=1, max_n=e2e_cb)
          t2 = time.perf_counter()
          t_train.append((t1 - t0) * 1e3)
          t_loss.append((t2 - t1) * 1e3)
      g.sync()
      tot = (time.perf_counter() - t_begin) * 1e3
      print(f"24 calls of 8 bunches: {tot / 24:.3f} ms per call ({24 * n / tot / 1e3:.3f} M frames/s);  bp_train() median "
            f"{statistics.median(t_train):.3f} ms (min {min(t_train):.3f}, max {max(t_train):.3f});  bp_train_losses() median "
            f"{statistics.median(t_loss):.3f} ms (max {max(t_loss):.3f})")
      print("   bp_train ms:", " ".join(f"{v:.2f}" for v in t_train))
      print("   losses   ms:", " ".join(f"{v:.2f}" for v in t_loss))
g.close()
