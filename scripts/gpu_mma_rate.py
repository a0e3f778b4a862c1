import ctypes as C, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "lib", "libbpgpu.so"))
lib.bp_last_error.restype = C.c_char_p
lib.bp_debug_mma_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
names = {0: "A=MN B=K ", 1: "A=K  B=K ", 2: "A=MN B=MN", 3: "A=K  B=MN"}
for bn in (64, 128, 256):
    for combo in range(4):
        row = []
        for mode in (0, 1, 2, 3):
            a, b = C.c_double(0), C.c_double(0)
            rc = lib.bp_debug_mma_rate(combo, bn, 2000, mode, C.byref(a), C.byref(b))
            if rc: print(lib.bp_last_error().decode()); raise SystemExit(1)
            row.append(f"mode{mode}: issue {a.value:6.1f} total {b.value:6.1f}")
        print(f"N={bn:3d} {names[combo]} cycles/MMA  " + " | ".join(row), flush=True)
