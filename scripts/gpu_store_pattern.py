"""Write throughput of the epilogue store pattern vs a tile-contiguous layout (bp_debug_store_pattern)."""
import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
bp = importlib.import_module("dnn-for-speech-enhancement_b200")
lib = bp.load_library()
lib.bp_debug_store_pattern.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
for mb in (32, 512):
    for ld in (2048, 2176, 8192):
        for ctas in (148, 592):
            r = []
            for layout in (0, 1):
                g = C.c_double(0)
                rc = lib.bp_debug_store_pattern(layout, ld, mb, 20, ctas, C.byref(g))
                assert rc == 0, lib.bp_last_error()
                r.append(g.value)
            print(f"buffer {mb:4d} MB  ld {ld:5d}  ctas {ctas:4d}:  row-major tiles {r[0]:7.0f} GB/s   contiguous tiles {r[1]:7.0f} GB/s")
