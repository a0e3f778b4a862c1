#!/usr/bin/env python
"""bench.py — frames/sec of the frame-wise DNN hot path (BASELINE.json metric) on N B200s.

A "step" is one bunch (minibatch) of spliced LPS frames through forward + back-prop + SGD update on the C2 workload of
BASELINE.json: 2827 -> 2048x3 (ReLU) -> 257, bunch 1024 per GPU (weak scaling: global bunch = 1024 * N), synthetic
frames, Glorot-initialised weights (Gen_rand_net scheme), lrate 1 (reference script), momentum 0.9.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C1|C2|C3|C4|C5]
  N > 1 is launched by `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...` (one rank/GPU).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs in HBM before the timed region, CUDA
events on the library's compute stream, max over ranks); `e2e` = the same work through the reference-facing C-ABI
call bp_train() with PINNED HOST buffers (H2D of every step's inputs inside the timed region, D2H of every step's loss).
`--impl reference` times the CPU restatement of the reference path (oracle/, the reference has no CPU implementation of
its own — SURVEY.md §8c) on the host cores.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (layersizes, local bunch, dropoutflag, visible_omit, hid_omit, train?)   [C1: sigmoid, see ACTIVATION]
    "C1": ([257, 512, 257], 128, 0, 0.0, 0.0, False),   # BASELINE configs[0]: the CPU plumbing case, forward only
    "C2": ([2827, 2048, 2048, 2048, 257], 1024, 0, 0.0, 0.0, True),
    "C3": ([3084, 2048, 2048, 2048, 257], 2048, 1, 0.2, 0.2, True),
    "C4": ([2827, 2048, 2048, 2048, 2048, 2048, 257], 512, 0, 0.0, 0.0, True),
    "C5": ([2827, 2048, 2048, 2048, 257], 8192, 0, 0.0, 0.0, False),
}


ACTIVATION = {"C1": 1}   # 1 = sigmoid (the reference's commented variant, DevFunc.cu:52,62); everything else ReLU


def flops_per_frame(sizes, train):
    P = sum(sizes[i] * sizes[i + 1] for i in range(len(sizes) - 1))
    return 2.0 * (3 * P - sizes[0] * sizes[1]) if train else 2.0 * P


def n_params(sizes):
    return sum((sizes[i] + 1) * sizes[i + 1] for i in range(len(sizes) - 1))


def glorot(sizes, seed=3, beta=0.5):
    rng = np.random.default_rng(seed)
    L = len(sizes)
    w, b = [None] * L, [None] * L
    for i in range(1, L):
        r = beta * np.sqrt(6.0) / np.sqrt(sizes[i - 1] + sizes[i])
        w[i] = rng.uniform(-r, r, size=(sizes[i - 1], sizes[i])).astype(np.float32)
        b[i] = np.zeros(sizes[i], dtype=np.float32)
    return w, b


def synth(n, k0, nout, seed, out_x=None, out_t=None):
    rng = np.random.default_rng(seed)
    x = out_x if out_x is not None else np.empty((n, k0), np.float32)
    t = out_t if out_t is not None else np.empty((n, nout), np.float32)
    rng.standard_normal((n, k0), dtype=np.float32, out=x)
    rng.standard_normal((n, nout), dtype=np.float32, out=t)
    t *= 0.5
    return x, t


class ClockSampler:
    """Samples SM clock / power / clock-event reasons through NVML (nvidia_ml_py) every ~2 ms from a thread while the
    warm-up and the timed region run, so that even a 60 ms timed region is covered (nvidia-smi -lms cannot)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.idx, self.rows, self.stop_flag, self.th, self.max_mhz = gpu_index, [], False, None, None
        self.in_region = False   # False, True (= the K timed steps) or a region name ("steady", "isolated")

    def _run(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            while not self.stop_flag:
                sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                pw = N.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    rs = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), pw, int(rs), self.in_region))
                time.sleep(0.002)
        except Exception:
            pass

    def start(self):
        import threading
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def region(self, name):
        """Clock summary of the samples taken while in_region == name (None if there were none)."""
        rows = [r for r in self.rows if r[3] == name]
        if not rows:
            return None
        reasons = sorted({n for r in rows for bit, n in self.REASONS.items() if r[2] & bit})
        return {"sm_mhz": float(np.median([r[0] for r in rows])), "sm_mhz_min": min(r[0] for r in rows),
                "power_w_max": max(r[1] for r in rows), "reasons": reasons, "samples": len(rows)}

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=5)
        timed = [r for r in self.rows if r[3] is True] or self.rows
        sm = [r[0] for r in timed]
        reasons = set()
        for r in timed:
            for bit, name in self.REASONS.items():
                if r[2] & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "power_w_max": max((r[1] for r in timed), default=None), "reasons": sorted(reasons),
                "samples": len(self.rows), "samples_in_timed_region": sum(1 for r in self.rows if r[3] is True)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class CpuPort:
    """The CPU restatement (oracle port, OpenMP over all host cores) stepping through the same workload: one step = one
    bunch of `frames` frames (train: forward + back-prop + update; decode: forward).  TEST INFRASTRUCTURE used as the
    reported baseline only (cpu_baseline / --impl reference) — never on the product path."""

    def __init__(self, sizes, frames, train, dropout=(0, 0.0, 0.0), pool=4, activation=0):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_py as O
        # torchrun exports OMP_NUM_THREADS=1 to its children: the baseline uses the host cores it reports
        self.threads = O.set_threads(host_cores())
        w, b = glorot(sizes)
        self.frames, self.train, self.pool, self.i = frames, train, pool, 0
        self.x, self.t = synth(frames * pool, sizes[0], sizes[-1], seed=1)
        self.net = O.Net(sizes, frames, lrate=1.0, momentum=0.9, dropoutflag=dropout[0], visible_omit=dropout[1],
                         hid_omit=dropout[2], activation=activation, weights=w, bias=b)

    def step(self):
        o = (self.i % self.pool) * self.frames
        self.i += 1
        if self.train:
            self.net.train(self.frames, self.x[o:o + self.frames], self.t[o:o + self.frames])
        else:
            self.net.forward(self.x[o:o + self.frames])

    def run(self, steps=None, min_seconds=None, max_steps=64):
        """Times `steps` steps, or (steps=None) as many as fit in ~min_seconds, at most max_steps."""
        n = 0
        t0 = time.perf_counter()
        while steps is None or steps > 0:
            self.step()
            n += 1
            dt = time.perf_counter() - t0
            if steps is not None:
                if n >= steps:
                    break
            elif dt >= min_seconds or n >= max_steps:
                break
        return n, time.perf_counter() - t0

    def describe(self, n):
        return (f"{n} step(s) of one {self.frames}-frame bunch each, "
                f"{'train step (fwd+bwd+SGD)' if self.train else 'forward'}, oracle port oracle/bp_oracle.c "
                f"(fp32, OpenMP, cache-blocked GEMM)")


def isolated_dominant_gemm(bp, sizes, lb, reps=2000, sampler=None):
    """The dominant product of the workload (a hidden layer's forward affine, units x frames x fan-in) launched alone,
    back to back `reps` times on its own stream and timed with CUDA events inside the library (bp_debug_gemm):
    operands L2-resident, no neighbours — the kernel's own ceiling, next to the in-place figure of the launch timeline.
    2000 launches (~35 ms) so that the host-side NVML sampler sees the loop (round 1 timed 20 launches after an idle
    gap and got 26 us where the same kernel runs 17-18 us once the clock is up — VERDICT r1)."""
    import ctypes as C
    lib = bp.load_library()
    M, N, K = sizes[2], lb, sizes[1]
    rng = np.random.default_rng(5)
    A = rng.standard_normal((K, M), dtype=np.float32)
    B = rng.standard_normal((N, K), dtype=np.float32)
    bias = np.zeros(M, np.float32)
    out = np.empty((N, M), np.float32)
    fp = C.POINTER(C.c_float)
    ms = C.c_float(0)
    old = os.environ.get("BP_DBG_REPS")
    rc = -1
    try:
        for n_launch, region in ((reps, False), (reps, "isolated")):   # first pass: clocks up, caches warm
            os.environ["BP_DBG_REPS"] = str(n_launch)
            if sampler is not None:
                sampler.in_region = region
            rc = lib.bp_debug_gemm(0, M, N, K, A.ctypes.data_as(fp), M, B.ctypes.data_as(fp), K,
                                   out.ctypes.data_as(fp), M, bias.ctypes.data_as(fp), None, 0, 1.0, 0, 0,
                                   C.byref(ms))
            if sampler is not None:
                sampler.in_region = False
            if rc != 0:
                break
    finally:
        if old is None:
            os.environ.pop("BP_DBG_REPS", None)
        else:
            os.environ["BP_DBG_REPS"] = old
    if rc != 0 or ms.value <= 0:
        return None
    return {"product": f"forward affine {M} units x {N} frames x {K} fan-in (bias + ReLU epilogue)", "launches": reps,
            "us_per_launch": ms.value * 1e3, "tflops": 2.0 * M * N * K / (ms.value * 1e-3) / 1e12,
            "clocks": sampler.region("isolated") if sampler is not None else None}


TRAFFIC_FILE = "r2_traffic.json"


class Ctx:
    """Rank plumbing shared by the measurements: trainer creation (+ communicator), barrier, max over ranks."""

    def __init__(self, bp, dist, rank, local_rank, world):
        self.bp, self.dist, self.rank, self.local_rank, self.world = bp, dist, rank, local_rank, world

    def create(self, sizes, gb, w, b, dropout, act, math_mode, world=None, seed=12345):
        world = self.world if world is None else world
        g = self.bp.BP_GPU(1, len(sizes), sizes, gb, 1.0, 0.9, 0.0, w, b, dropout[0], dropout[1], dropout[2],
                           seed=seed, device=self.local_rank, world_size=world, rank=self.rank if world > 1 else 0,
                           activation=act, math_mode=math_mode)
        if world > 1:
            import torch
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if self.rank == 0:
                idt = torch.tensor(list(self.bp.comm_unique_id()), dtype=torch.uint8, device="cuda")
            self.dist.broadcast(idt, 0)
            g.comm_init(bytes(idt.cpu().tolist()))
        return g

    def barrier(self, g=None):
        if g is not None:
            g.sync()
        if self.dist is not None:
            import torch
            torch.cuda.synchronize()
            self.dist.barrier()

    def max_over_ranks(self, v):
        if self.dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def timeline_summary(marks, nb_tl, sizes, lb):
    """Per-launch durations from bp_get_timeline (completion time of every launch of a bunch since its start mark,
    averaged over nb_tl bunches; an event sits behind every launch, so there is no PDL overlap across marks): each
    launch's duration = its completion minus the completion of the launch before it ON ITS STREAM."""
    done_at = dict(marks)
    out = {"bunches": nb_tl, "launch_done_at_us": {k: round(v * 1e3, 2) for k, v in marks}}
    prev = {"compute": 0.0, "side": 0.0}
    dur = {}
    for k, v in marks:
        st = "side" if "(side)" in k else "compute"
        if k == "start":
            continue
        # a side-stream launch cannot start before the compute-stream event it waits for; its duration is bounded
        # above by (done - previous side launch done): reported as such
        dur[k] = round((v - prev[st]) * 1e3, 2)
        prev[st] = v
    out["launch_us"] = dur
    if nb_tl > 0 and "fwd1" in done_at and "fwd2" in done_at:
        us = (done_at["fwd2"] - done_at["fwd1"]) * 1e3
        out["dominant"] = {"product": f"forward affine {sizes[2]} units x {lb} frames x {sizes[1]} fan-in, in place "
                                      "(completion of fwd2 minus completion of fwd1)",
                           "us_per_launch": us, "tflops": 2.0 * sizes[2] * lb * sizes[1] / (us * 1e-6) / 1e12}
    return out


def dp_parity(ctx, sizes, gb, dropout, act, math_mode, nb=4):
    """N ranks x (gb/N rows) with the gradient exchange must compute the 1-rank step on the full bunch (SURVEY.md §8e;
    the reference's dead multi-GPU code did not, BP_GPU.cu:841 vs :880).  Fresh trainers:
    (1) every replica's weights are bit-identical (checksum all-gathered) after nb global bunches,
    (2) rank 0 replays the same global bunches on a 1-rank trainer and reports the relative Frobenius error of the
        weight UPDATE — after ONE bunch (`rel_update_err`: the exchange itself; only the order in which the per-rank
        partial sums are added differs) and after all nb (`rel_update_err_after_nb`: the same difference fed back
        through nb - 1 further single-pass-TF32 steps, whose sensitivity to summation order is ~1e-2 per step on
        these inputs — tests/test_config_shapes.py)."""
    import torch
    import zlib
    bp, dist, rank, world = ctx.bp, ctx.dist, ctx.rank, ctx.world
    lb = gb // world
    w, b = glorot(sizes)
    # rows of rank r in global bunch i: rows [i*lb, (i+1)*lb) of synth(seed=500+r)
    x, t = synth(nb * lb, sizes[0], sizes[-1], seed=500 + rank)
    res = {"bunches": nb}
    snaps = {}
    for n_b in (1, nb):
        g = ctx.create(sizes, gb, w, b, dropout, act, math_mode, seed=777)
        ctx.barrier(g)
        g.train(n_b * lb, x[: n_b * lb], t[: n_b * lb])
        snaps[n_b] = g.returnWeights()
        ctx.barrier(g)
        g.close()
    ws, bs = snaps[nb]
    crc = 0
    for l in range(1, len(sizes)):
        crc = zlib.crc32(ws[l].tobytes(), crc)
        crc = zlib.crc32(bs[l].tobytes(), crc)
    allc = torch.zeros(world, dtype=torch.int64, device="cuda")
    allc[rank] = crc
    dist.all_reduce(allc)
    crcs = [int(v) for v in allc.cpu().tolist()]
    res["replicas_identical"] = len(set(crcs)) == 1
    res["weights_crc32"] = f"{crcs[0]:08x}"
    if rank == 0:
        xs = [synth(nb * lb, sizes[0], sizes[-1], seed=500 + r) for r in range(world)]
        xg = np.concatenate([xs[r][0][i * lb:(i + 1) * lb] for i in range(nb) for r in range(world)])
        tg = np.concatenate([xs[r][1][i * lb:(i + 1) * lb] for i in range(nb) for r in range(world)])

        def rel(n_b):
            s1 = ctx.create(sizes, gb, w, b, dropout, act, math_mode, world=1, seed=777)
            s1.train(n_b * gb, xg[: n_b * gb], tg[: n_b * gb])
            sw, _sb = s1.returnWeights()
            s1.close()
            num = den = 0.0
            for l in range(1, len(sizes)):
                d = float(np.linalg.norm((snaps[n_b][0][l].astype(np.float64) - sw[l]).ravel()))
                n = float(np.linalg.norm((sw[l].astype(np.float64) - w[l]).ravel()))
                num, den = num + d * d, den + n * n
            return float(np.sqrt(num / den))
        res["rel_update_err"] = rel(1)
        res["rel_update_err_after_nb"] = rel(nb)
        res["what"] = ("||W_dp - W_1rank||_F / ||W_1rank - W_0||_F over all layers: after one global bunch of "
                       f"{gb} frames, and after {nb}")
    ctx.barrier()
    return res


def c4_subline(ctx, bp, K, W, args, sampler):
    """BASELINE configs[3]: 2827 -> 2048 x 5 -> 257, GLOBAL bunch 4096 split over the N ranks (512 rows per GPU at
    N = 8), device-resident chunks, K timed steps after W warm-up steps; plus dp_parity of that net."""
    sizes, _lb, dflag, vo, ho, _train = WORKLOADS["C4"]
    world, rank = ctx.world, ctx.rank
    gb = 4096
    lb = gb // world
    w, b = glorot(sizes)
    g = ctx.create(sizes, gb, w, b, (dflag, vo, ho), 0, bp.BP_MATH_TF32)
    cbn = int(np.ceil(260e6 / (lb * sizes[0] * 4)))      # resident chunk larger than L2
    px, pt = bp.PinnedArray((cbn * lb, sizes[0])), bp.PinnedArray((cbn * lb, sizes[-1]))
    synth(cbn * lb, sizes[0], sizes[-1], seed=300 + rank, out_x=px.array, out_t=pt.array)
    g.upload_chunk(cbn * lb, px.array, pt.array)
    ctx.barrier(g)

    def run(n):
        done = 0
        while done < n:
            k = min(cbn, n - done)
            g.train_resident(0, k)
            done += k
    run(2 * W)
    ctx.barrier(g)
    if sampler is not None:
        sampler.in_region = "c4"
    g.timer_start()
    run(K)
    ms = g.timer_stop()
    if sampler is not None:
        sampler.in_region = False
    ctx.barrier(g)
    ms = ctx.max_over_ranks(ms)
    g.set_profiling(True)
    run(min(max(K, 32), 64))
    prof, nprof = g.profile()
    g.set_profiling(False)
    exch = {0: "none", 1: "nccl", 2: "p2p"}[g.get_option("dp_exchange")]
    ctx.barrier(g)
    g.close()
    px.free()
    pt.free()
    fl = flops_per_frame(sizes, True) * lb
    out = {"workload": "C4: " + "-".join(map(str, sizes)) + " train, global bunch 4096 (strong scaling over N)",
           "n_gpus": world, "bunch_per_gpu": lb, "steps": K, "value": K * gb / (ms * 1e-3), "unit": "frames/s",
           "ms_per_step": ms / K, "exchange": exch, "tflops_per_gpu": fl / (ms / K * 1e-3) / 1e12,
           "per_class_ms": {k: v / max(nprof, 1) for k, v in prof.items()},
           "clocks": sampler.region("c4") if sampler is not None else None}
    if not args.no_dp_parity:
        out["dp_parity"] = dp_parity(ctx, sizes, gb, (dflag, vo, ho), 0, bp.BP_MATH_TF32, nb=2)
    return out


def ref_gpu_arm(sizes, lb, dropout, nb=16):
    """The reference's OWN CUDA path on this GPU, beside ours: class BP_GPU of /root/reference compiled unmodified
    (+ the one-token compile fix DevFunc.cu:67,81) into oracle/_ref/ref_harness by oracle/build_ref.sh — cuBLAS FP32
    GEMMs + its element-wise kernels, BP_GPU::train (BP_GPU.cu:241-331) on one chunk of nb bunches from pageable host
    memory (its own H2D inside, as in its main loop), second repetition, wall time device-synchronised.  A checker
    binary run as a subprocess AFTER our timed regions; nothing of it is on the product path."""
    import struct
    import tempfile
    H = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(H):
        return {"unavailable": "oracle/_ref/ref_harness not built (needs /root/reference at build time)"}
    n = nb * lb
    rng = np.random.default_rng(0)
    try:
        with tempfile.TemporaryDirectory() as d:
            fin, fout = os.path.join(d, "in.blob"), os.path.join(d, "out.blob")
            w, b = glorot(sizes)
            with open(fin, "wb") as f:
                f.write(struct.pack(f"<i{len(sizes)}i5i", len(sizes), *sizes, lb, n, 0, dropout[0], 2))
                f.write(struct.pack("<5f", 1.0, 0.9, 0.0, dropout[1], dropout[2]))
                for l in range(1, len(sizes)):
                    f.write(w[l].astype("<f4").tobytes())
                    f.write(b[l].astype("<f4").tobytes())
                f.write(rng.standard_normal((n, sizes[0]), dtype=np.float32).tobytes())
                f.write((0.5 * rng.standard_normal((n, sizes[-1]), dtype=np.float32)).tobytes())
            r = subprocess.run([H, fin, fout], capture_output=True, text=True, timeout=600, cwd=d)
            if r.returncode != 0:
                return {"error": (r.stdout + r.stderr)[-300:]}
            ms = float(np.fromfile(fout, dtype="<f4")[-1])
    except Exception as e:
        return {"error": str(e)}
    return {"value": n / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms / nb, "steps": nb, "kind": "reference",
            "what": "the reference's unmodified BP_GPU::train (cuBLAS FP32 + its element-wise kernels, "
                    "oracle/_ref/ref_harness) on the same GPU, same net and bunch, one chunk of "
                    f"{nb} bunches from pageable host memory, H2D inside as in its own main loop"}


def decode_e2e(bp, g, args, sizes, lb, gb, cb, world, rank, px, barrier, max_over_ranks, K):
    """End-to-end decode figures (C5): bp_forward host in / host out, the pipelined variant, and the raw-record
    variants (device-side splice)."""
    n_calls = max(8, K // 4)
    out = None
    barrier(g)
    t0 = time.perf_counter()
    for c in range(n_calls):
        out = g.forward(lb, px.array[:lb])
    dt = time.perf_counter() - t0
    dt = max_over_ranks(dt)
    e2e = {"value": n_calls * gb / dt, "unit": "frames/s", "h2d_bytes_per_step": 4 * lb * sizes[0] * world,
           "d2h_bytes_per_step": 4 * lb * sizes[-1] * world, "api": "bp_forward() host in / host out"}
    # The same spliced rows, two batches in flight (bp_forward_submit / bp_forward_wait): H2D of batch k+1 beside
    # the forward pass of batch k and the read-back of batch k-1.
    try:
        pout = [bp.PinnedArray((lb, sizes[-1])) for _ in range(2)]
        g.forward_submit(lb, px.array[:lb], pout[0].array)
        g.forward_wait()
        barrier(g)
        t0 = time.perf_counter()
        g.forward_submit(lb, px.array[:lb], pout[0].array)
        for c in range(1, n_calls):
            g.forward_submit(lb, px.array[(c % cb) * lb: (c % cb + 1) * lb], pout[c & 1].array)
            g.forward_wait()
        g.forward_wait()
        dt_p = max_over_ranks(time.perf_counter() - t0)
        e2e["pipelined"] = {"value": n_calls * gb / dt_p, "unit": "frames/s",
                            "h2d_bytes_per_step": 4 * lb * sizes[0] * world,
                            "d2h_bytes_per_step": 4 * lb * sizes[-1] * world,
                            "api": "bp_forward_submit() / bp_forward_wait(), 2 batches in flight"}
    except Exception as e:   # an extra measurement must not hide the bench line
        e2e["pipelined"] = {"error": str(e)}
    # The same decode fed the way BPtrain reader=gpu feeds it: raw big-endian Pfile records from
    # pinned memory, 11-frame splice + normalisation on the device — 1/11 of the H2D bytes of the spliced rows.
    try:
        fea_dim, ctx_w = 257, 11
        if sizes[0] == fea_dim * ctx_w:
            n_rec = lb + ctx_w - 1
            prec = bp.PinnedArray((n_rec, fea_dim + 2))
            words = np.random.default_rng(7 + rank).standard_normal((n_rec, fea_dim + 2), dtype=np.float32)
            prec.array[:] = words.view(np.uint32).byteswap().view(np.float32)   # records are big-endian on disk
            raw = bp.RawChunk(fea_dim, ctx_w, ctx_w // 2, 0, prec.array, None, np.zeros(fea_dim, np.float32),
                              np.ones(fea_dim, np.float32), np.arange(lb, dtype=np.int32))
            g.decode_raw(raw)
            barrier(g)
            t0 = time.perf_counter()
            for c in range(n_calls):
                out = g.decode_raw(raw)
            dt_raw = max_over_ranks(time.perf_counter() - t0)
            e2e["raw_reader"] = {"value": n_calls * gb / dt_raw, "unit": "frames/s",
                                 "h2d_bytes_per_step": 4 * n_rec * (fea_dim + 2) * world,
                                 "d2h_bytes_per_step": 4 * lb * sizes[-1] * world,
                                 "api": "bp_crossvalid_raw() raw Pfile records in / enhanced frames out"}
            # ... and pipelined: two chunks in flight, upload of k+1 / forward of k+1 / read-back of k overlap
            pout = [bp.PinnedArray((lb, sizes[-1])) for _ in range(2)]
            g.decode_raw_submit(raw, pout[0].array)
            g.decode_raw_wait()
            barrier(g)
            n_pipe = 4 * n_calls
            t0 = time.perf_counter()
            g.decode_raw_submit(raw, pout[0].array)
            for c in range(1, n_pipe):
                g.decode_raw_submit(raw, pout[c & 1].array)
                g.decode_raw_wait()          # chunk c-1 is complete in pout[(c-1) & 1]
            g.decode_raw_wait()
            dt_pipe = max_over_ranks(time.perf_counter() - t0)
            e2e["raw_reader_pipelined"] = {"value": n_pipe * gb / dt_pipe, "unit": "frames/s", "calls": n_pipe,
                                           "h2d_bytes_per_step": 4 * n_rec * (fea_dim + 2) * world,
                                           "d2h_bytes_per_step": 4 * lb * sizes[-1] * world,
                                           "api": "bp_decode_raw_submit() / bp_decode_raw_wait(), 2 chunks in flight",
                                           "checksum": float(np.abs(pout[(n_pipe - 1) & 1].array).sum())}
    except Exception as e:   # an extra measurement must not hide the bench line
        e2e["raw_reader"] = {"error": str(e)}
    return e2e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--chunk-bunches", type=int, default=32, help="bunches resident per chunk (inputs > L2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timeline", action="store_true", help="(default on; kept for old command lines)")
    ap.add_argument("--e2e-raw", action="store_true", help="(default on; kept for old command lines)")
    ap.add_argument("--steady-seconds", type=float, default=2.0,
                    help="also run the same loop for this long after the K timed steps (0 = skip)")
    for flag in ("timeline", "e2e-raw", "dp-parity", "c4", "3xtf32", "ref-gpu"):
        ap.add_argument(f"--no-{flag}", action="store_true", help=f"skip the {flag} part of the line")
    ap.add_argument("--math", default="tf32", choices=["tf32", "3xtf32"],
                    help="tf32 = single-pass TF32 products (default); 3xtf32 = split precision, ~fp32 accuracy")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # Exactly ONE line on stdout: libraries (NCCL prints its version) write to fd 1 too, so fd 1 is pointed at stderr
    # for the duration of the run and the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    if world == 1 and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun (one rank per GPU)", file=sys.stderr)
        return 2
    sizes, lb, dflag, vo, ho, train = WORKLOADS[args.workload]
    act = ACTIVATION.get(args.workload, 0)
    K, W = args.steps, max(args.warmup, 3)
    gb = lb * world
    metric = "frames_per_sec"
    cfg = {"workload": f"{args.workload}: {'-'.join(map(str, sizes))} {'train (fwd+bwd+SGD)' if train else 'forward decode'}",
           "bunch_per_gpu": lb, "global_bunch": gb,
           "parallelism": f"dp{world}" if train or world == 1 else f"{world} independent replicas (no exchange)",
           "dropout": [dflag, vo, ho], "activation": "sigmoid" if act else "relu", "l2_policy": "inputs larger than L2 (resident chunk cycles)",
           "math": ("3xTF32 split-precision tensor-core products (fp32 storage, fp32 accumulate, ~fp32 accuracy)"
                    if args.math == "3xtf32" else "tf32 tensor-core products (fp32 storage, fp32 accumulate)")}

    # ------------------------------------------------------------------ reference arm: CPU restatement
    if args.impl == "reference":
        if rank != 0:
            return 0
        # One step = one bunch of the workload on the host cores.  The sample per step is the workload's own bunch
        # unless W + K of them would run for more than ~150 s on this host; then the bunch is cut (stated in `sample`).
        budget_s = 150.0
        frames = lb
        # `config` stays identical to our arm's (the driver compares them); this arm's arithmetic is stated beside it
        port = CpuPort(sizes, frames, train, (dflag, vo, ho), activation=act)
        _, t1 = port.run(steps=1)          # first call: thread start-up, page faults
        _, t1 = port.run(steps=1)
        if (K + W) * t1 > budget_s and frames > 64:
            frames = max(64, int(frames * budget_s / ((K + W) * t1)) // 64 * 64)
            port = CpuPort(sizes, frames, train, (dflag, vo, ho), activation=act)
            port.run(steps=1)
        port.run(steps=max(0, W - 2))
        n, dt = port.run(steps=K)
        fps = n * frames / dt
        cores = port.threads
        what = port.describe(n) + ("" if frames == lb else f"; bunch cut from {lb} to {frames} frames to bound the run")
        line = {"impl": "reference", "metric": metric, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": n, "warmup": W, "ms_per_step": 1e3 * dt / n,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": what},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "arithmetic": "literal fp32 on the host cores (one fused multiply-add per term, ascending k)",
                "note": "the reference has no CPU implementation of this path (CUDA+cuBLAS only); this is the literal "
                        "fp32 CPU restatement oracle/bp_oracle.c timed on the host cores"}
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm
    bp = importlib.import_module("dnn-for-speech-enhancement_b200")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        import datetime
        # a rank that dies must not leave the others (and the box) waiting for the default 10-minute watchdog
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=180))
    ctx = Ctx(bp, dist, rank, local_rank, world)
    w, b = glorot(sizes)
    math_mode = bp.BP_MATH_3XTF32 if args.math == "3xtf32" else bp.BP_MATH_TF32
    # Training shards a GLOBAL bunch over the ranks (gradient exchange).  Decode has no exchange step: every rank is an
    # independent replica decoding its own batches of `lb` frames ("replicas only").
    g = ctx.create(sizes, gb if train else lb, w, b, (dflag, vo, ho), act, math_mode, world=None if train else 1)
    barrier, max_over_ranks = ctx.barrier, ctx.max_over_ranks

    cb = args.chunk_bunches
    n_chunk = cb * lb
    px, pt = bp.PinnedArray((n_chunk, sizes[0])), bp.PinnedArray((n_chunk, sizes[-1]))
    synth(n_chunk, sizes[0], sizes[-1], seed=100 + rank, out_x=px.array, out_t=pt.array)
    g.upload_chunk(n_chunk, px.array, pt.array)
    barrier(g)  # ranks generate their chunks on shared host cores: line them up before the first collective bunch

    in_place_mask = train and dflag == 1 and vo > 0.0   # input dropout masks the resident rows in place (as the
    timed = {"ms": None}                                 # reference does): every pass needs a fresh upload, untimed

    def run_resident(h, n_steps, cbn=cb, rows=lb):
        done = 0
        while done < n_steps:
            k = min(cbn, n_steps - done)
            if in_place_mask:
                if timed["ms"] is not None:
                    timed["ms"] += h.timer_stop()
                h.upload_chunk(n_chunk, px.array, pt.array)
                h.sync()
                if timed["ms"] is not None:
                    h.timer_start()
            if train:
                h.train_resident(0, k)
            else:
                h.forward_resident(0, k * rows)
            done += k

    def timed_run(h, n_steps):
        """Device time of n_steps steps (CUDA events on the compute stream); re-uploads excluded."""
        timed["ms"] = 0.0 if in_place_mask else None
        h.timer_start()
        run_resident(h, n_steps)
        ms_ = h.timer_stop() + (timed["ms"] or 0.0)
        timed["ms"] = None
        return ms_

    # ---- value: device-resident, EXACTLY K timed steps
    run_resident(g, W)
    if rank == 0:   # NVML initialisation can take a second: make sure the sampler is live before the timed region
        t_wait = time.perf_counter()
        while not sampler.rows and time.perf_counter() - t_wait < 3.0:
            time.sleep(0.01)
    run_resident(g, W)  # EVERY rank: a bunch is a collective step (gradient exchange), ranks must run the same count
    launches0 = g.counters()[0]
    barrier(g)
    sampler.in_region = True
    ms = timed_run(g, K)
    sampler.in_region = False
    barrier(g)
    ms = max_over_ranks(ms)
    launches = g.counters()[0] - launches0
    value = K * gb / (ms * 1e-3)

    # ---- steady: the same loop for >= 2 s (the K-step region above lasts a few ms, i.e. it is a burst-clock figure;
    # this one shows what the workload does once power and clocks have settled), with its own clock samples
    steady = None
    if args.steady_seconds > 0:
        n_st = int(np.ceil(args.steady_seconds * 1e3 / (ms / K) / cb)) * cb
        barrier(g)
        sampler.in_region = "steady"
        ms_st = timed_run(g, n_st)
        sm_after = g.get_option("sm_clock_mhz")
        sampler.in_region = False
        barrier(g)
        ms_st = max_over_ranks(ms_st)
        steady = {"steps": n_st, "seconds": ms_st * 1e-3, "ms_per_step": ms_st / n_st,
                  "value": n_st * gb / (ms_st * 1e-3), "unit": "frames/s",
                  "sm_clock_mhz_on_device_after": sm_after}

    # ---- roofline split: same loop with per-class CUDA events (ring of the last 64 bunches, no host sync per bunch)
    prof = None
    if train:
        g.set_profiling(True)
        run_resident(g, min(max(K, 32), 64))
        prof, nprof = g.profile()
        g.set_profiling(False)

    # ---- launch timeline (an event behind every launch, 16 bunches): every product's in-place duration
    tl = None
    if train and not args.no_timeline and len(sizes) > 3:
        try:
            g.set_timeline(True)
            run_resident(g, 16)
            marks, nb_tl = g.timeline()
            g.set_timeline(False)
            tl = timeline_summary(marks, nb_tl, sizes, lb)
        except Exception as e:   # a measurement aid must not hide the bench line
            tl = {"error": str(e)}

    # ---- e2e: bp_train() from pinned host buffers, one chunk of `e2e_cb` bunches per call, loss read back per step.
    # At least 32 calls: with 2 calls (K = 20) the un-overlapped first upload was a third of the sample.
    e2e_cb = 8
    e2e = None
    if train:
        n_calls = max(32, K // e2e_cb)

        def e2e_loop(call, n):
            call(0)                         # warm-up call (buffers, pipeline fill)
            barrier(g)
            t0 = time.perf_counter()
            for c in range(n):
                call(c)
                if c > 0:
                    g.train_losses(age=1, max_n=e2e_cb)   # D2H of every step's loss (8 B/step), one call behind
            last = g.train_losses(age=0, max_n=e2e_cb)
            g.sync()
            return max_over_ranks(time.perf_counter() - t0), last

        def call_rows(c):
            o = (c % (cb // e2e_cb)) * e2e_cb * lb
            g.train(e2e_cb * lb, px.array[o: o + e2e_cb * lb], pt.array[o: o + e2e_cb * lb])

        dt, losses = e2e_loop(call_rows, n_calls)
        e2e = {"value": n_calls * e2e_cb * gb / dt, "unit": "frames/s", "calls": n_calls,
               "h2d_bytes_per_step": 4 * lb * (sizes[0] + sizes[-1]) * world, "d2h_bytes_per_step": 8 * world,
               "api": f"bp_train() on pinned host chunks of {e2e_cb} bunches + bp_train_losses()",
               "last_loss": float(losses[0]) / (lb * sizes[-1])}
        if not args.no_e2e_raw and sizes[0] in (257 * 11, 257 * 12):   # 12 = with the NAT block (C3)
            # The same chunks fed the way BPtrain reader=gpu feeds them: raw big-endian records (feature + target
            # Pfiles) from pinned memory, splice / normalise on the device: ~1/6 of the H2D bytes.
            try:
                fea_dim, ctx_w, n_s = 257, 11, e2e_cb * lb
                n_rec = n_s + ctx_w - 1
                rng = np.random.default_rng(11)            # the same records on every rank
                pf, ptg = bp.PinnedArray((n_rec, fea_dim + 2)), bp.PinnedArray((n_rec, sizes[-1] + 2))
                for pa, scale in ((pf, 1.0), (ptg, 0.5)):
                    words = rng.standard_normal(pa.array.shape, dtype=np.float32) * np.float32(scale)
                    pa.array[:] = words.view(np.uint32).byteswap().view(np.float32)
                nat = 1 if sizes[0] == fea_dim * (ctx_w + 1) else 0
                frames = np.arange(n_s, dtype=np.int32)
                raw = bp.RawChunk(fea_dim, ctx_w, ctx_w // 2, nat, pf.array, ptg.array, np.zeros(fea_dim, np.float32),
                                  np.ones(fea_dim, np.float32), frames,
                                  (frames // 300) * 300 if nat else None)   # "sentences" of 300 frames for the NAT mean
                dt_raw, _ = e2e_loop(lambda c: g.train_raw(raw), n_calls)
                # a raw chunk is the GLOBAL chunk: every rank is handed the same n_s samples and assembles its own
                # rows of each global bunch on its device, so one call trains n_s frames in total
                e2e["raw_reader"] = {"value": n_calls * n_s / dt_raw, "unit": "frames/s",
                                     "h2d_bytes_per_step": 4 * n_rec * (fea_dim + sizes[-1] + 4) * world * gb // n_s,
                                     "d2h_bytes_per_step": 8 * world,
                                     "api": f"bp_train_raw() on pinned raw-record chunks of {e2e_cb} bunches "
                                            "(device-side splice, BPtrain reader=gpu)"}
            except Exception as e:   # an extra measurement must not hide the bench line
                e2e["raw_reader"] = {"error": str(e)}
    else:
        e2e = decode_e2e(bp, g, args, sizes, lb, gb, cb, world, rank, px, barrier, max_over_ranks, K)

    extras = {}
    if world > 1:
        extras["exchange"] = {0: "none", 1: "nccl", 2: "p2p"}[g.get_option("dp_exchange")]
        extras["schedule"] = "chained launches" if g.get_option("chain") != 0 else "one launch per product"
    g.close()
    px.free()
    pt.free()

    # ---- N > 1: does the data-parallel step compute the 1-rank step?  (fresh trainers, 4 global bunches)
    if world > 1 and train and not args.no_dp_parity:
        try:
            extras["dp_parity"] = dp_parity(ctx, sizes, gb, (dflag, vo, ho), act, math_mode, nb=4)
        except Exception as e:
            extras["dp_parity"] = {"error": str(e)}
    # ---- N > 1: BASELINE configs[3] (C4: 5 x 2048, global bunch 4096 split over the N ranks) gets its own sub-line
    if world > 1 and args.workload == "C2" and not args.no_c4 and 4096 % world == 0:
        try:
            extras["c4"] = c4_subline(ctx, bp, K, W, args, sampler if rank == 0 else None)
        except Exception as e:
            extras["c4"] = {"error": str(e)}
    # ---- N = 1: the equal-precision mode (3xTF32 ~ fp32 accuracy) on the same workload
    if world == 1 and train and args.math == "tf32" and not args.no_3xtf32 and not in_place_mask:
        try:
            g3 = ctx.create(sizes, gb, w, b, (dflag, vo, ho), act, bp.BP_MATH_3XTF32)
            px3, pt3 = bp.PinnedArray((n_chunk, sizes[0])), bp.PinnedArray((n_chunk, sizes[-1]))
            synth(n_chunk, sizes[0], sizes[-1], seed=100, out_x=px3.array, out_t=pt3.array)
            g3.upload_chunk(n_chunk, px3.array, pt3.array)
            run_resident(g3, W)
            g3.sync()
            ms3 = timed_run(g3, K)
            extras["tf32x3"] = {"value": K * gb / (ms3 * 1e-3), "unit": "frames/s", "ms_per_step": ms3 / K, "steps": K,
                                "math": "3xTF32 split precision (A*B + A_lo*B + A*B_lo into one fp32 accumulator)"}
            g3.close()
            px3.free()
            pt3.free()
        except Exception as e:
            extras["tf32x3"] = {"error": str(e)}

    iso = None
    if rank == 0 and world == 1 and train and len(sizes) > 3:
        try:
            iso = isolated_dominant_gemm(bp, sizes, lb, sampler=sampler)
        except Exception as e:  # a measurement aid must not hide the bench line
            iso = {"error": str(e)}
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks, peak_src = load_peaks()
    tf32_burst, tf32_sust = peaks["bf16_tflops"] / 2.0, peaks["bf16_tflops_sustained"] / 2.0
    traffic = {}
    try:  # DRAM bytes of the committed ncu --set full capture (profiles/), per bunch / per launch
        traffic = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
    except Exception:
        pass
    line = {"metric": metric, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32x3" if args.math == "3xtf32" else "tf32",
            "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "impl": "ours"}
    line.update(extras)
    fl = flops_per_frame(sizes, train) * lb   # per rank per step
    if steady is not None:
        steady["clocks"] = sampler.region("steady")
        steady["tflops"] = fl / (steady["ms_per_step"] * 1e-3) / 1e12
        steady["frac_of_sustained_tf32_peak"] = steady["tflops"] / tf32_sust
        line["steady"] = steady
    ach_step = fl / (ms / K * 1e-3) / 1e12
    # `roofline`: the WHOLE step from ms_per_step (all products + update + gaps) against the burst TF32 peak — the
    # K-step region lasts milliseconds at the boost clock, so burst is its denominator; the sustained peak and the
    # >= 2 s figure are given beside it.  tcgen05 kind::tf32 issues at half the bf16 rate: peak = bf16 peak / 2.
    line["roofline"] = {"bound": "tensor",
                        "kernel": ("whole train step: bp_gemm2_kernel / bp_gemm_kernel products + bp_sgd_kernel"
                                   if train else "bp_gemm2_kernel / bp_gemm_kernel (forward chain)"),
                        "achieved": ach_step, "peak": tf32_burst, "unit": "TFLOP/s", "frac": ach_step / tf32_burst,
                        "frac_of_sustained_peak": ach_step / tf32_sust, "peak_sustained": tf32_sust,
                        "algorithmic_gflop_per_step": fl / 1e9,
                        "traffic": traffic.get("gemm_dram_bytes_per_bunch") if args.workload == "C2" else None,
                        "traffic_source": f"profiles/{TRAFFIC_FILE}: ncu --set full, dram__bytes_read+write summed over "
                                          "the GEMM launches of one bunch (cold caches, serialised replays)",
                        "peak_source": f"{peak_src}: bf16_tflops/2 (burst) and bf16_tflops_sustained/2"}
    if prof is not None and nprof > 0:
        line["roofline"]["per_class_ms"] = {k: v / nprof for k, v in prof.items()}
        if tl is not None:
            line["roofline"]["launch_timeline"] = tl
            dom = tl.get("dominant") if isinstance(tl, dict) else None
            if dom:
                dom["frac"] = dom["tflops"] / tf32_burst
                line["roofline"]["dominant_kernel_in_place"] = dom
        if iso:
            if "tflops" in iso:
                iso["frac"] = iso["tflops"] / tf32_burst
            line["roofline"]["isolated_dominant_kernel"] = iso
        if world == 1 and prof["sgd"] / nprof >= 0.003:
            sgd_ms = prof["sgd"] / nprof
            # algorithmic minimum 20 B/parameter over the parameters the timed launch updates: layer 1
            # when the upper layers were updated early, else all of them
            early = prof["sgd_upper"] > 0.0 and len(sizes) > 2
            sgd_bytes = 20.0 * ((sizes[0] + 1) * sizes[1] if early else n_params(sizes))
            line["roofline_sgd"] = {"bound": "hbm", "kernel": "bp_sgd_kernel (final launch: first layer)",
                                    "achieved": sgd_bytes / (sgd_ms * 1e-3) / 1e9,
                                    "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": sgd_bytes / (sgd_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                    "traffic": traffic.get("sgd_dram_bytes_per_launch") if args.workload == "C2" else None,
                                    "peak_source": peak_src}
        elif world > 1:
            line["exchange_ms"] = {"exposed_dw_tail": prof["dw"] / nprof, "owner_update_and_all_gather": prof["sgd"] / nprof,
                                   "note": "per-class CUDA events of rank 0; the gradient reduce-scatter runs inside "
                                           "the dW epilogues, the update + all-gather in bp_peer_sgd_kernel"}
            try:   # NVLink counters are not readable on the pool's boxes: the ALGORITHMIC bytes of the exchange instead
                line["nvlink"] = nvlink_algorithmic(sizes, world, prof["sgd"] / nprof)
            except Exception as e:
                line["nvlink"] = {"error": str(e)}
    if world == 1 and train and not args.no_ref_gpu:
        rg = ref_gpu_arm(sizes, lb, (dflag, vo, ho))
        if rg and "value" in rg:
            rg["ours_e2e_over_ref"] = e2e["value"] / rg["value"]
            rg["ours_device_over_ref"] = value / rg["value"]
        line["ref_gpu"] = rg
    if not args.no_cpu_baseline and world == 1:
        try:
            port = CpuPort(sizes, lb, train, (dflag, vo, ho), activation=act)
            port.run(steps=1)                                  # thread start-up, page faults
            n, dt = port.run(min_seconds=10.0, max_steps=64)   # ~10 s of CPU work on all host cores
            line["cpu_baseline"] = {"value": n * lb / dt, "unit": "frames/s", "cores": port.threads, "kind": "port",
                                    "sample": port.describe(n), "seconds": dt}
        except Exception as e:  # the oracle is test infrastructure; its absence must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "port",
                                    "sample": f"unavailable: {e}"}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def nvlink_algorithmic(sizes, world, all_gather_ms):
    """Bytes a rank stores into its peers' memory per step of the peer-memory exchange (csrc/bp_peer.cuh): the padded
    parameter arena, minus the 1/world of it the rank owns, once as partial gradients (reduce-scatter, from the dW
    epilogues) and once as new weights (all-gather, from bp_peer_sgd_kernel); every rank receives the same amounts."""
    arena = 4.0 * n_params_padded(sizes)
    leg = arena * (world - 1) / world
    return {"kind": "algorithmic (not a counter)", "arena_bytes": arena,
            "reduce_scatter_bytes_out_per_rank": leg, "all_gather_bytes_out_per_rank": leg,
            "bytes_in_per_rank_per_step": 2 * leg,
            "all_gather_gbs_out_per_rank": leg / (all_gather_ms * 1e-3) / 1e9 if all_gather_ms > 0 else None,
            "note": "reduce-scatter = remote stores of the dW epilogues (under the back-propagation launch); all-gather = "
                    "remote stores of bp_peer_sgd_kernel, timed together with the wait for every rank's partial "
                    "gradients and the owner's update"}


def n_params_padded(sizes):
    """Floats in the device parameter arena (row stride padded to 32, +1 bias row per layer) — what the SGD kernel
    actually streams; the algorithmic minimum is 20 B x n_params(sizes)."""
    tot = 0
    for i in range(len(sizes) - 1):
        ld = (sizes[i + 1] + 31) // 32 * 32
        tot += ((sizes[i] + 1) * ld + 63) // 64 * 64
    return tot


if __name__ == "__main__":
    sys.exit(main())
