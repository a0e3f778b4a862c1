"""dnn-for-speech-enhancement_b200 — Python host-side mirror of the reference's trainer object.

The product is ``lib/libbpgpu.so`` (hand-written sm_100a CUDA behind the C-ABI of ``include/bp_gpu.h``); this module
is only the ctypes binding plus a class with the reference's own member names (``BP_GPU``: constructor, ``train``,
``CrossValid``, ``returnWeights`` — reference BP_GPU.h:40-88) so tests read like calls into the reference.

There is no CPU fallback: importing works anywhere (so the C-ABI surface can be checked), but constructing a
``BP_GPU`` without the built library or without a B200 raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbpgpu.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "bp_gpu.h")

BP_MAXLAYER = 10
BP_ACT_RELU, BP_ACT_SIGMOID = 0, 1
BP_MATH_TF32, BP_MATH_3XTF32 = 0, 1

_fp = C.POINTER(C.c_float)
_fpp = C.POINTER(_fp)


class BpConfig(C.Structure):
    """struct bp_config of include/bp_gpu.h."""
    _fields_ = [("device", C.c_int), ("world_size", C.c_int), ("rank", C.c_int), ("numlayers", C.c_int),
                ("layersizes", C.c_int * BP_MAXLAYER), ("bunchsize", C.c_int), ("lrate", C.c_float),
                ("momentum", C.c_float), ("weightcost", C.c_float), ("dropoutflag", C.c_int),
                ("visible_omit", C.c_float), ("hid_omit", C.c_float), ("activation", C.c_int),
                ("math_mode", C.c_int), ("seed", C.c_uint64)]


class BpRawChunk(C.Structure):
    """struct bp_raw_chunk of include/bp_gpu.h."""
    _fields_ = [("fea_dim", C.c_int), ("fea_context", C.c_int), ("targ_offset", C.c_int), ("nat", C.c_int),
                ("n_records", C.c_int), ("n_samples", C.c_int), ("fea_records", C.c_void_p),
                ("targ_records", C.c_void_p), ("mean", _fp), ("inv_std", _fp), ("sample_frame", C.POINTER(C.c_int)),
                ("sample_seg", C.POINTER(C.c_int)), ("sample_row", C.POINTER(C.c_int))]


class RawChunk:
    """One chunk for the device-side reader: raw big-endian Pfile records + the sample table (bp_raw_chunk)."""

    def __init__(self, fea_dim, fea_context, targ_offset, nat, fea_records, targ_records, mean, inv_std,
                 sample_frame, sample_seg=None, sample_row=None):
        self.fea_dim, self.fea_context, self.targ_offset, self.nat = int(fea_dim), int(fea_context), int(targ_offset), int(nat)
        self.fea_records = np.ascontiguousarray(fea_records).view(np.uint32).reshape(-1, self.fea_dim + 2)
        self.n_records = self.fea_records.shape[0]
        self.targ_records = None if targ_records is None else np.ascontiguousarray(targ_records).view(np.uint32)
        self.mean, self.inv_std = mean, inv_std
        self.sample_frame = np.ascontiguousarray(sample_frame, dtype=np.int32)
        self.sample_seg = self.sample_frame if sample_seg is None else np.ascontiguousarray(sample_seg, dtype=np.int32)
        self.sample_row = None if sample_row is None else np.ascontiguousarray(sample_row, dtype=np.int32)


class BpError(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """Load libbpgpu.so (built in-tree by ``__graft_entry__.build()`` / ``make -C csrc``). Fails loudly if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BpError(f"{LIB_PATH} is not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.bp_last_error.restype = C.c_char_p
    lib.bp_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_float,
                              C.c_float, C.c_float, _fpp, _fpp, C.c_int, C.c_float, C.c_float]
    lib.bp_create_ex.argtypes = [C.POINTER(C.c_void_p), C.POINTER(BpConfig), _fpp, _fpp]
    lib.bp_destroy.argtypes = [C.c_void_p]
    lib.bp_destroy.restype = None
    lib.bp_train.argtypes = [C.c_void_p, C.c_int, _fp, _fp]
    lib.bp_crossvalid.argtypes = [C.c_void_p, C.c_int, _fp, _fp, _fp]
    lib.bp_forward.argtypes = [C.c_void_p, C.c_int, _fp, _fp]
    lib.bp_return_weights.argtypes = [C.c_void_p, _fpp, _fpp]
    lib.bp_upload_chunk.argtypes = [C.c_void_p, C.c_int, _fp, _fp]
    lib.bp_train_resident.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.bp_forward_resident.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, C.POINTER(C.c_double)]
    lib.bp_sync.argtypes = [C.c_void_p]
    lib.bp_get_timeline.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_float), C.c_int,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.bp_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.bp_get_option.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
    lib.bp_begin_epoch.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int]
    lib.bp_set_dropout_seed.argtypes = [C.c_void_p, C.c_uint64]
    lib.bp_upload_raw_chunk.argtypes = [C.c_void_p, C.POINTER(BpRawChunk)]
    lib.bp_train_raw.argtypes = [C.c_void_p, C.POINTER(BpRawChunk)]
    lib.bp_decode_raw_submit.argtypes = [C.c_void_p, C.POINTER(BpRawChunk), _fp]
    lib.bp_decode_raw_wait.argtypes = [C.c_void_p]
    lib.bp_forward_submit.argtypes = [C.c_void_p, C.c_int, _fp, _fp]
    lib.bp_forward_wait.argtypes = [C.c_void_p]
    lib.bp_crossvalid_raw.argtypes = [C.c_void_p, C.POINTER(BpRawChunk), _fp, _fp]
    lib.bp_download_chunk.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, _fp]
    lib.bp_host_alloc.argtypes = [C.c_size_t]
    lib.bp_host_alloc.restype = C.c_void_p
    lib.bp_host_free.argtypes = [C.c_void_p]
    lib.bp_host_free.restype = None
    lib.bp_timer_start.argtypes = [C.c_void_p]
    lib.bp_timer_stop.argtypes = [C.c_void_p, _fp]
    lib.bp_get_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.bp_set_profiling.argtypes = [C.c_void_p, C.c_int]
    lib.bp_get_profile.argtypes = [C.c_void_p, C.c_float * 6, C.POINTER(C.c_uint64)]
    lib.bp_train_losses.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]
    lib.bp_comm_unique_id.argtypes = [C.c_char * 128]
    lib.bp_comm_init.argtypes = [C.c_void_p, C.c_char * 128]
    lib.bp_dropout_mask.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float]
    lib.bp_debug_gemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _fp, C.c_int, _fp, C.c_int, _fp, C.c_int, _fp,
                                  _fp, C.c_int, C.c_float, C.c_int, C.c_int, _fp]
    lib.bp_debug_sgd.argtypes = [C.c_int, _fp, _fp, _fp, C.c_int, C.c_float, C.c_float, C.c_float]
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise BpError(f"{what} failed (rc={rc}): {load_library().bp_last_error().decode(errors='replace')}")


def _as_f32(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None and a.shape != tuple(shape):
        a = a.reshape(shape)
    return a


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_fp)


class PinnedArray:
    """float32 array in page-locked host memory (bp_host_alloc) so chunk uploads are true asynchronous DMA."""

    def __init__(self, shape):
        lib = load_library()
        n = int(np.prod(shape))
        self._p = lib.bp_host_alloc(max(n, 1) * 4)
        if not self._p:
            raise BpError("bp_host_alloc failed")
        self.array = np.ctypeslib.as_array((C.c_float * n).from_address(self._p)).reshape(shape)

    def free(self):
        if self._p:
            load_library().bp_host_free(self._p)
            self._p = None
            self.array = None


class BP_GPU:
    """Mirror of the reference class ``BP_GPU`` (BP_GPU.h:40-88) over the C-ABI.

    ``weights`` / ``bias`` are sequences indexed like the reference's ``float **`` arguments: entry ``i`` belongs to
    weight layer ``i`` (1..numlayers-1) and entry 0 is unused (may be ``None``); ``weights[i]`` has shape
    ``(layersizes[i-1], layersizes[i])`` i.e. ``w[in*n_out + out]`` (BP_GPU.cu:188, Extend_rand_net.cpp:203).
    """

    def __init__(self, a_GPU_selected: int, a_numlayers: int, a_layersizes: Sequence[int], a_bunchsize: int,
                 a_lrate: float, a_momentum: float, a_weightcost: float, weights, bias, dropoutflag: int = 0,
                 visible_omit: float = 0.0, hid_omit: float = 0.0, *, activation: int = BP_ACT_RELU,
                 math_mode: int = BP_MATH_TF32, seed: int = 0x5EED5EED, device: Optional[int] = None,
                 world_size: int = 1, rank: int = 0):
        lib = load_library()
        self.numlayers = int(a_numlayers)
        self.layersizes = [int(x) for x in a_layersizes][: self.numlayers]
        self.bunchsize = int(a_bunchsize)
        self.world_size = int(world_size)
        self.rank = int(rank)
        self._h = C.c_void_p()
        w_arr, b_arr, self._keep = self._pack(weights, bias)
        if world_size == 1 and device is None and a_GPU_selected > 1:
            # reference-style in-process multi-GPU group (gpu_used = N)
            os.environ["BP_SEED"] = str(int(seed))
            os.environ["BP_ACTIVATION"] = "sigmoid" if activation == BP_ACT_SIGMOID else "relu"
            sizes = (C.c_int * self.numlayers)(*self.layersizes)
            _check(lib.bp_create(C.byref(self._h), int(a_GPU_selected), self.numlayers, sizes, self.bunchsize,
                                 a_lrate, a_momentum, a_weightcost, w_arr, b_arr, int(dropoutflag), visible_omit,
                                 hid_omit), "bp_create")
        else:
            cfg = BpConfig()
            cfg.device = int(device if device is not None else 0)
            cfg.world_size, cfg.rank = self.world_size, self.rank
            cfg.numlayers = self.numlayers
            for i, s in enumerate(self.layersizes):
                cfg.layersizes[i] = s
            cfg.bunchsize = self.bunchsize
            cfg.lrate, cfg.momentum, cfg.weightcost = a_lrate, a_momentum, a_weightcost
            cfg.dropoutflag, cfg.visible_omit, cfg.hid_omit = int(dropoutflag), visible_omit, hid_omit
            cfg.activation, cfg.math_mode, cfg.seed = int(activation), int(math_mode), int(seed)
            _check(lib.bp_create_ex(C.byref(self._h), C.byref(cfg), w_arr, b_arr), "bp_create_ex")
        self._keep = None

    # ------------------------------------------------------------------ helpers
    def _pack(self, weights, bias):
        L = self.numlayers
        keep: List[np.ndarray] = []
        w_arr, b_arr = (_fp * BP_MAXLAYER)(), (_fp * BP_MAXLAYER)()
        for i in range(1, L):
            w = _as_f32(weights[i], (self.layersizes[i - 1], self.layersizes[i]))
            b = _as_f32(bias[i], (self.layersizes[i],))
            keep += [w, b]
            w_arr[i], b_arr[i] = _ptr(w), _ptr(b)
        return w_arr, b_arr, keep

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            load_library().bp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ reference members
    def train(self, n_frames: int, indata, targ) -> None:
        """BP_GPU::train (BP_GPU.cu:241-331)."""
        x = _as_f32(indata).reshape(-1)
        t = _as_f32(targ).reshape(-1)
        assert x.size >= n_frames * self.layersizes[0] and t.size >= n_frames * self.layersizes[-1]
        _check(load_library().bp_train(self._h, int(n_frames), _ptr(x), _ptr(t)), "bp_train")

    def CrossValid(self, n_frames: int, indata, targ) -> float:
        """BP_GPU::CrossValid (BP_GPU.cu:408-479): returns the SUM of squared errors."""
        x = _as_f32(indata).reshape(-1)
        t = _as_f32(targ).reshape(-1)
        s = C.c_float(0)
        _check(load_library().bp_crossvalid(self._h, int(n_frames), _ptr(x), _ptr(t), C.byref(s)), "bp_crossvalid")
        return float(s.value)

    def returnWeights(self):
        """BP_GPU::returnWeights (BP_GPU.cu:910-923) -> (weights, bias) lists indexed 1..numlayers-1 (entry 0 None)."""
        L = self.numlayers
        ws: List[Optional[np.ndarray]] = [None] * L
        bs: List[Optional[np.ndarray]] = [None] * L
        w_arr, b_arr = (_fp * BP_MAXLAYER)(), (_fp * BP_MAXLAYER)()
        for i in range(1, L):
            ws[i] = np.empty((self.layersizes[i - 1], self.layersizes[i]), dtype=np.float32)
            bs[i] = np.empty((self.layersizes[i],), dtype=np.float32)
            w_arr[i], b_arr[i] = _ptr(ws[i]), _ptr(bs[i])
        _check(load_library().bp_return_weights(self._h, w_arr, b_arr), "bp_return_weights")
        return ws, bs

    # ------------------------------------------------------------------ extensions
    def forward(self, n_frames: int, indata) -> np.ndarray:
        """Decode: enhanced frames (the outputs cv_bunch_single computes and discards, BP_GPU.cu:445-473)."""
        x = _as_f32(indata).reshape(-1)
        out = np.empty((n_frames, self.layersizes[-1]), dtype=np.float32)
        _check(load_library().bp_forward(self._h, int(n_frames), _ptr(x), _ptr(out)), "bp_forward")
        return out

    def upload_chunk(self, n_frames: int, indata, targ=None) -> None:
        x = indata if isinstance(indata, np.ndarray) and indata.dtype == np.float32 else _as_f32(indata)
        t = None if targ is None else (targ if isinstance(targ, np.ndarray) and targ.dtype == np.float32
                                       else _as_f32(targ))
        _check(load_library().bp_upload_chunk(self._h, int(n_frames), _ptr(x), _ptr(t)), "bp_upload_chunk")

    def _raw_chunk(self, raw: "RawChunk"):
        rc = BpRawChunk()
        keep = []

        def ptr(a, dtype, ctype):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dtype)
            keep.append(a)
            return a.ctypes.data_as(ctype)

        rc.fea_dim, rc.fea_context, rc.targ_offset, rc.nat = raw.fea_dim, raw.fea_context, raw.targ_offset, raw.nat
        rc.n_records, rc.n_samples = raw.n_records, len(raw.sample_frame)
        rc.fea_records = ptr(raw.fea_records, np.uint32, C.c_void_p)
        rc.targ_records = ptr(raw.targ_records, np.uint32, C.c_void_p)
        rc.mean = ptr(raw.mean, np.float32, _fp)
        rc.inv_std = ptr(raw.inv_std, np.float32, _fp)
        rc.sample_frame = ptr(raw.sample_frame, np.int32, C.POINTER(C.c_int))
        rc.sample_seg = ptr(raw.sample_seg, np.int32, C.POINTER(C.c_int))
        rc.sample_row = ptr(raw.sample_row, np.int32, C.POINTER(C.c_int))
        return rc, keep

    def upload_raw_chunk(self, raw: "RawChunk") -> None:
        """Device-side reader: ship the chunk's Pfile records + sample table, assemble the samples on the GPU."""
        rc, _keep = self._raw_chunk(raw)
        _check(load_library().bp_upload_raw_chunk(self._h, C.byref(rc)), "bp_upload_raw_chunk")

    def train_raw(self, raw: "RawChunk") -> None:
        rc, _keep = self._raw_chunk(raw)
        _check(load_library().bp_train_raw(self._h, C.byref(rc)), "bp_train_raw")

    def crossvalid_raw(self, raw: "RawChunk", want_out: bool = False):
        rc, _keep = self._raw_chunk(raw)
        out = np.empty((len(raw.sample_frame), self.layersizes[-1]), dtype=np.float32) if want_out else None
        sq = C.c_float(0)
        _check(load_library().bp_crossvalid_raw(self._h, C.byref(rc), C.byref(sq), _ptr(out)), "bp_crossvalid_raw")
        return float(sq.value), out

    def decode_raw(self, raw: "RawChunk") -> np.ndarray:
        """Decode straight from raw Pfile records (no targets needed): bp_crossvalid_raw with a null score pointer —
        the records travel as they lie in the file (1/11 of the spliced rows' bytes), the rows are assembled on the
        device, the enhanced frames come back."""
        rc, _keep = self._raw_chunk(raw)
        out = np.empty((len(raw.sample_frame), self.layersizes[-1]), dtype=np.float32)
        _check(load_library().bp_crossvalid_raw(self._h, C.byref(rc), None, _ptr(out)), "bp_crossvalid_raw")
        return out

    def decode_raw_submit(self, raw: "RawChunk", out: np.ndarray) -> None:
        """Pipelined decode (bp_decode_raw_submit): queue the chunk and return; `out` (float32, n_samples x
        layersizes[-1], ideally a PinnedArray's array) holds the enhanced frames after the matching decode_raw_wait().
        At most two chunks in flight."""
        if out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"] or out.size < len(raw.sample_frame) * self.layersizes[-1]:
            raise ValueError("decode_raw_submit: out must be C-contiguous float32 with n_samples x layersizes[-1] elements")
        rc, _keep = self._raw_chunk(raw)
        _check(load_library().bp_decode_raw_submit(self._h, C.byref(rc), _ptr(out)), "bp_decode_raw_submit")

    def decode_raw_wait(self) -> None:
        _check(load_library().bp_decode_raw_wait(self._h), "bp_decode_raw_wait")

    def forward_submit(self, n_frames: int, indata: np.ndarray, out: np.ndarray) -> None:
        """Pipelined bp_forward (bp_forward_submit): spliced rows in, `out` complete after the matching forward_wait();
        at most two chunks in flight.  Both arrays float32 C-contiguous (PinnedArray for asynchronous copies)."""
        if indata.dtype != np.float32 or out.dtype != np.float32 or not indata.flags["C_CONTIGUOUS"] \
                or not out.flags["C_CONTIGUOUS"] or out.size < n_frames * self.layersizes[-1]:
            raise ValueError("forward_submit: float32 C-contiguous arrays, out >= n_frames x layersizes[-1]")
        _check(load_library().bp_forward_submit(self._h, int(n_frames), _ptr(indata), _ptr(out)), "bp_forward_submit")

    def forward_wait(self) -> None:
        _check(load_library().bp_forward_wait(self._h), "bp_forward_wait")

    def download_chunk(self, first_row: int, n_rows: int, want_targ: bool = True):
        x = np.empty((n_rows, self.layersizes[0]), dtype=np.float32)
        t = np.empty((n_rows, self.layersizes[-1]), dtype=np.float32) if want_targ else None
        _check(load_library().bp_download_chunk(self._h, int(first_row), int(n_rows), _ptr(x), _ptr(t)),
               "bp_download_chunk")
        return x, t

    def train_resident(self, first_bunch: int, n_bunches: int) -> None:
        _check(load_library().bp_train_resident(self._h, int(first_bunch), int(n_bunches)), "bp_train_resident")

    def forward_resident(self, first_frame: int, n_frames: int, want_out: bool = False, want_sqerr: bool = False):
        out = np.empty((n_frames, self.layersizes[-1]), dtype=np.float32) if want_out else None
        sq = C.c_double(0)
        _check(load_library().bp_forward_resident(self._h, int(first_frame), int(n_frames), _ptr(out),
                                                  C.byref(sq) if want_sqerr else None), "bp_forward_resident")
        return out, (float(sq.value) if want_sqerr else None)

    def set_option(self, name: str, value: int) -> None:
        """Run-time switch of an experimental code path (bp_set_option), e.g. ``set_option("fused_update", 1)``."""
        _check(load_library().bp_set_option(self._h, name.encode(), int(value)), "bp_set_option")

    def get_option(self, name: str) -> int:
        """Read-only state (bp_get_option): "dp_exchange" (0 single rank / 1 NCCL / 2 peer memory), "sm_clock_mhz"
        (SM clock measured on the device behind whatever is queued), or the value of a set_option switch."""
        v = C.c_int(0)
        _check(load_library().bp_get_option(self._h, name.encode(), C.byref(v)), "bp_get_option")
        return int(v.value)

    def sync(self) -> None:
        _check(load_library().bp_sync(self._h), "bp_sync")

    def begin_epoch(self, lrate: float, momentum: float, weightcost: float, reset_dropout_step: bool = True) -> None:
        """Epoch boundary inside one process: zero momentum deltas, new hyper-parameters, weights stay on the device —
        the state a fresh reference process starts from (BP_GPU.cu:10, 137-138) minus the .wts round-trip."""
        _check(load_library().bp_begin_epoch(self._h, lrate, momentum, weightcost, int(reset_dropout_step)),
               "bp_begin_epoch")

    def set_dropout_seed(self, seed: int) -> None:
        """Re-key the dropout masks from the next bunch on (bp_set_dropout_seed)."""
        _check(load_library().bp_set_dropout_seed(self._h, C.c_uint64(seed & (2 ** 64 - 1))), "bp_set_dropout_seed")

    def timer_start(self) -> None:
        _check(load_library().bp_timer_start(self._h), "bp_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        _check(load_library().bp_timer_stop(self._h, C.byref(ms)), "bp_timer_stop")
        return float(ms.value)

    def counters(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(load_library().bp_get_counters(self._h, C.byref(a), C.byref(b)), "bp_get_counters")
        return int(a.value), int(b.value)

    def set_profiling(self, on: bool) -> None:
        _check(load_library().bp_set_profiling(self._h, 1 if on else 0), "bp_set_profiling")

    def set_timeline(self, on: bool) -> None:
        """Launch timeline of the next <= 16 training bunches (bp_set_profiling(h, 2)); off = plain profiling off."""
        _check(load_library().bp_set_profiling(self._h, 2 if on else 0), "bp_set_profiling")

    def timeline(self):
        """[(label, ms since the bunch's start mark)] per launch, averaged over the recorded bunches, + their count."""
        lab = C.create_string_buffer(4096)
        ms = (C.c_float * 64)()
        n, nb = C.c_int(0), C.c_int(0)
        _check(load_library().bp_get_timeline(self._h, lab, 4096, ms, 64, C.byref(n), C.byref(nb)), "bp_get_timeline")
        names = lab.value.decode().split("\n")[: n.value]
        return [(names[i], float(ms[i])) for i in range(n.value)], nb.value

    def profile(self):
        ms = (C.c_float * 6)()
        n = C.c_uint64(0)
        _check(load_library().bp_get_profile(self._h, ms, C.byref(n)), "bp_get_profile")
        names = ["fwd", "dx", "dw", "sgd", "allreduce_wait", "sgd_upper"]
        return {k: float(v) for k, v in zip(names, ms)}, int(n.value)

    def train_losses(self, age: int = 0, max_n: int = 4096) -> np.ndarray:
        """Per-bunch sum of squared output error of the most recent (age 0) / previous (age 1) train call."""
        buf = (C.c_double * max_n)()
        n = C.c_int(0)
        _check(load_library().bp_train_losses(self._h, int(age), buf, max_n, C.byref(n)), "bp_train_losses")
        return np.array(buf[: n.value], dtype=np.float64)

    def comm_init(self, id128: bytes) -> None:
        buf = (C.c_char * 128).from_buffer_copy(id128)
        _check(load_library().bp_comm_init(self._h, buf), "bp_comm_init")


def comm_unique_id() -> bytes:
    buf = (C.c_char * 128)()
    _check(load_library().bp_comm_unique_id(buf), "bp_comm_unique_id")
    return bytes(buf.raw)


def dropout_mask(seed: int, step: int, layer: int, frame: int, unit: int, p: float) -> int:
    return int(load_library().bp_dropout_mask(seed, step, layer, frame, unit, p))


def declared_symbols(debug: bool = True) -> List[str]:
    """Entry points declared in include/bp_gpu.h (the drop-in boundary) and, with debug=True, include/bp_gpu_debug.h
    (test and bring-up aids) — parsed from the headers."""
    import re
    names = set()
    for path in [HEADER_PATH] + ([HEADER_PATH.replace("bp_gpu.h", "bp_gpu_debug.h")] if debug else []):
        txt = open(path).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(bp_[a-z_0-9]+)\s*\(", txt))
    return sorted(names)
