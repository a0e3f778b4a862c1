"""Can small kernels run beside a resident bp_chain_kernel?  (bp_debug_coresidency)"""
import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
bp = importlib.import_module("dnn-for-speech-enhancement_b200")
lib = bp.load_library()
lib.bp_debug_coresidency.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
bps, thr = int(sys.argv[1]), int(sys.argv[2])
sizes, lb, dflag, vo, ho, _ = bench.WORKLOADS["C2"]
w, b = bench.glorot(sizes)
g = bp.BP_GPU(1, len(sizes), sizes, lb, 1.0, 0.9, 0.0, w, b, dflag, vo, ho, seed=12345, device=0)
px, pt = bp.PinnedArray((2 * lb, sizes[0])), bp.PinnedArray((2 * lb, sizes[-1]))
bench.synth(2 * lb, sizes[0], sizes[-1], seed=100, out_x=px.array, out_t=pt.array)
g.upload_chunk(2 * lb, px.array, pt.array)
g.train_resident(0, 2)
g.sync()
ms = C.c_float(0)
rc = lib.bp_debug_coresidency(g.handle, bps, thr, C.byref(ms))
print(f"blocks/SM {bps} threads {thr}: rc {rc} {'OK %.3f ms' % ms.value if rc == 0 else lib.bp_last_error().decode()}", flush=True)
