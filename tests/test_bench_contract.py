"""bench.py contract on the CPU: the reference arm (`--impl reference`, the oracle port timed on the host cores) prints
exactly one JSON line with the keys the driver reads, non-zero ranks of a torchrun launch stay silent, and the FLOP /
parameter bookkeeping behind `roofline` equals SURVEY.md §8(d)'s table.  (The GPU arm needs a B200; its line is checked
by the driver at round end.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env_extra=None):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=600, env=env)


def test_reference_arm_prints_one_contract_line():
    p = _run(["--impl", "reference", "--steps", "2", "--warmup", "3", "--workload", "C2"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "frames_per_sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["steps"] == 2 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] - 1024) < 1e-6 * 1024   # one bunch per step
    assert d["config"]["workload"].startswith("C2: 2827-2048-2048-2048-257")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "1024-frame bunch" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3"],
             {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_multi_gpu_without_torchrun_is_refused():
    p = _run(["--gpus", "2", "--steps", "1"])
    assert p.returncode == 2 and "torchrun" in p.stderr and p.stdout.strip() == ""


def test_flop_and_parameter_bookkeeping_matches_survey_table():
    sys.path.insert(0, ROOT)
    import bench
    c2 = bench.WORKLOADS["C2"][0]
    c3 = bench.WORKLOADS["C3"][0]
    c4 = bench.WORKLOADS["C4"][0]
    # P, biases, forward and training MFLOP per frame
    assert sum(c2[i] * c2[i + 1] for i in range(4)) == 14704640 and bench.n_params(c2) == 14704640 + 6401
    assert abs(bench.flops_per_frame(c2, False) / 1e6 - 29.41) < 0.005
    assert abs(bench.flops_per_frame(c2, True) / 1e6 - 76.65) < 0.005
    assert abs(bench.flops_per_frame(c3, True) / 1e6 - 78.75) < 0.005
    assert bench.n_params(c4) == 23093248 + 10497 and abs(bench.flops_per_frame(c4, True) / 1e6 - 126.98) < 0.005
    assert bench.WORKLOADS["C4"][1] * 8 == 4096 and bench.WORKLOADS["C5"][1] == 8192 and not bench.WORKLOADS["C5"][5]


def test_exchange_bytes_of_the_peer_memory_step():
    """`nvlink` of the N > 1 line (algorithmic, the boxes expose no NVLink counters): each leg of the exchange moves the
    padded arena minus the rank's own 1/N share; the arena is >= the 58.84 MB of SURVEY.md §8(d)'s gradient size."""
    sys.path.insert(0, ROOT)
    import bench
    c2 = bench.WORKLOADS["C2"][0]
    for world in (2, 4, 8):
        nv = bench.nvlink_algorithmic(c2, world, 0.114)
        arena = nv["arena_bytes"]
        assert 4 * bench.n_params(c2) <= arena <= 1.01 * 4 * bench.n_params(c2)
        assert nv["all_gather_bytes_out_per_rank"] == arena * (world - 1) / world == nv["reduce_scatter_bytes_out_per_rank"]
        assert nv["bytes_in_per_rank_per_step"] == 2 * nv["all_gather_bytes_out_per_rank"]
        assert abs(nv["all_gather_gbs_out_per_rank"] - nv["all_gather_bytes_out_per_rank"] / 0.114e-3 / 1e9) < 1e-6
    assert bench.nvlink_algorithmic(c2, 8, 0.0)["all_gather_gbs_out_per_rank"] is None
    json.dumps(bench.nvlink_algorithmic(c2, 8, 0.1))
