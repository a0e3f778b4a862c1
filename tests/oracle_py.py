"""ctypes binding of oracle/libbporacle.so — TEST INFRASTRUCTURE ONLY (see the header of oracle/bp_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libbporacle.so")
MAXLAYER = 10
_fp = C.POINTER(C.c_float)


class OrcCfg(C.Structure):
    _fields_ = [("numlayers", C.c_int), ("layersizes", C.c_int * MAXLAYER), ("bunchsize", C.c_int),
                ("lrate", C.c_float), ("momentum", C.c_float), ("weightcost", C.c_float), ("dropoutflag", C.c_int),
                ("visible_omit", C.c_float), ("hid_omit", C.c_float), ("activation", C.c_int), ("tf32", C.c_int),
                ("accum_double", C.c_int), ("seed", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
        L = C.CDLL(LIB_PATH)
        pp = C.POINTER(_fp)
        L.orc_train_bunch.argtypes = [C.POINTER(OrcCfg), pp, pp, pp, pp, C.c_int, _fp, _fp, C.c_uint32, C.c_int, _fp]
        L.orc_train_bunch.restype = None
        L.orc_train.argtypes = [C.POINTER(OrcCfg), pp, pp, pp, pp, C.c_int, _fp, _fp, C.POINTER(C.c_uint32)]
        L.orc_forward.argtypes = [C.POINTER(OrcCfg), pp, pp, C.c_int, _fp, _fp]
        L.orc_forward.restype = None
        L.orc_crossvalid.argtypes = [C.POINTER(OrcCfg), pp, pp, C.c_int, _fp, _fp]
        L.orc_crossvalid.restype = C.c_float
        L.orc_sgd.argtypes = [C.c_size_t, _fp, _fp, _fp, C.c_int, C.c_float, C.c_float, C.c_float]
        L.orc_sgd.restype = None
        L.orc_dropout_mask.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float]
        L.orc_philox.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_philox.restype = None
        for fn in (L.orc_gemm, L.orc_gemm_naive):
            fn.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, C.c_int]
            fn.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_fp)


class Net:
    """Weights + momentum state of one network, driven through the oracle's functions."""

    def __init__(self, layersizes, bunchsize, lrate=1.0, momentum=0.0, weightcost=0.0, dropoutflag=0,
                 visible_omit=0.0, hid_omit=0.0, activation=0, tf32=0, accum_double=0, seed=0x5EED5EED, weights=None,
                 bias=None):
        self.cfg = OrcCfg()
        self.cfg.numlayers = len(layersizes)
        for i, s in enumerate(layersizes):
            self.cfg.layersizes[i] = int(s)
        self.cfg.bunchsize = int(bunchsize)
        self.cfg.lrate, self.cfg.momentum, self.cfg.weightcost = lrate, momentum, weightcost
        self.cfg.dropoutflag, self.cfg.visible_omit, self.cfg.hid_omit = int(dropoutflag), visible_omit, hid_omit
        self.cfg.activation, self.cfg.tf32, self.cfg.accum_double = int(activation), int(tf32), int(accum_double)
        self.cfg.seed = int(seed)
        self.sizes = list(layersizes)
        L = len(layersizes)
        self.w = [None] + [np.array(weights[i], dtype=np.float32, order="C", copy=True).reshape(
            layersizes[i - 1], layersizes[i]) for i in range(1, L)]
        self.b = [None] + [np.array(bias[i], dtype=np.float32, copy=True).reshape(layersizes[i]) for i in range(1, L)]
        self.dw = [None] + [np.zeros_like(self.w[i]) for i in range(1, L)]
        self.db = [None] + [np.zeros_like(self.b[i]) for i in range(1, L)]
        self.step = C.c_uint32(0)

    def _arr(self, lst):
        a = (_fp * MAXLAYER)()
        for i in range(1, len(self.sizes)):
            a[i] = _p(lst[i])
        return a

    def train(self, n_frames, x, t):
        x = np.ascontiguousarray(x, dtype=np.float32)
        t = np.ascontiguousarray(t, dtype=np.float32)
        return lib().orc_train(C.byref(self.cfg), self._arr(self.w), self._arr(self.b), self._arr(self.dw),
                               self._arr(self.db), int(n_frames), _p(x), _p(t), C.byref(self.step))

    def train_bunch(self, x, t, frame0=0, want_out=False):
        x = np.ascontiguousarray(x, dtype=np.float32)
        t = np.ascontiguousarray(t, dtype=np.float32)
        B = x.shape[0]
        out = np.empty((B, self.sizes[-1]), dtype=np.float32) if want_out else None
        lib().orc_train_bunch(C.byref(self.cfg), self._arr(self.w), self._arr(self.b), self._arr(self.dw),
                              self._arr(self.db), B, _p(x), _p(t), self.step.value, frame0,
                              _p(out) if want_out else None)
        self.step.value += 1
        return out

    def forward(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[0]
        out = np.empty((n, self.sizes[-1]), dtype=np.float32)
        lib().orc_forward(C.byref(self.cfg), self._arr(self.w), self._arr(self.b), n, _p(x), _p(out))
        return out

    def crossvalid(self, x, t):
        x = np.ascontiguousarray(x, dtype=np.float32)
        t = np.ascontiguousarray(t, dtype=np.float32)
        return float(lib().orc_crossvalid(C.byref(self.cfg), self._arr(self.w), self._arr(self.b), x.shape[0], _p(x),
                                          _p(t)))


def set_threads(n):
    """OpenMP threads of the oracle's loops (n <= 0: query only); returns the count in force."""
    f = lib().orc_set_threads
    f.argtypes, f.restype = [C.c_int], C.c_int
    return int(f(int(n)))


def sgd(delta, w, grad, n, momentum, lr, wc):
    lib().orc_sgd(delta.size, _p(delta), _p(w), _p(grad), int(n), momentum, lr, wc)


def gemm(A, B, init=None, tf32=0, naive=False):
    """C = init + A @ B through the oracle's blocked GEMM core (or the naive triple loop it must equal bit for bit)."""
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    M, R = A.shape
    N = B.shape[1]
    out = np.empty((M, N), dtype=np.float32)
    ini = None if init is None else np.ascontiguousarray(init, dtype=np.float32)
    (lib().orc_gemm_naive if naive else lib().orc_gemm)(M, R, N, _p(A), _p(B), _p(ini) if ini is not None else None,
                                                        _p(out), int(tf32))
    return out


def dropout_mask(seed, step, tensor, frame, unit, p):
    return int(lib().orc_dropout_mask(seed, step, tensor, frame, unit, p))


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox(c, k, o)
    return [int(v) for v in o]


def glorot_init(layersizes, seed=3, beta=0.5):
    """Glorot-uniform beta*sqrt(6)/sqrt(n_in+n_out), zero biases — the reference's Gen_rand_net.cpp:84-101 scheme."""
    rng = np.random.default_rng(seed)
    L = len(layersizes)
    w, b = [None] * L, [None] * L
    for i in range(1, L):
        r = beta * np.sqrt(6.0) / np.sqrt(layersizes[i - 1] + layersizes[i])
        w[i] = rng.uniform(-r, r, size=(layersizes[i - 1], layersizes[i])).astype(np.float32)
        b[i] = np.zeros(layersizes[i], dtype=np.float32)
    return w, b


def synth_data(n, k0, nout, seed=1):
    """Deterministic synthetic (already normalised / spliced) frames and smooth targets."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, k0), dtype=np.float32)
    A = np.random.default_rng(seed + 1).standard_normal((min(k0, 64), nout)).astype(np.float32) / 8.0
    t = np.tanh(x[:, : A.shape[0]] @ A).astype(np.float32) + 0.1 * rng.standard_normal((n, nout), dtype=np.float32)
    return x, t.astype(np.float32)
