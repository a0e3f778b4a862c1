"""Data-parallel parity on real GPUs (run under torchrun, one rank per GPU):
N ranks x (bunch/N rows) with NCCL all-reduce of the gradients must reproduce the 1-rank run on the full bunch.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/gpu_dp_check.py
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
bp = importlib.import_module("dnn-for-speech-enhancement_b200")
import oracle_py as O  # noqa: E402  (data / init helpers only)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    sizes, B, nb = [300, 512, 384, 129], 256, 3
    lb = B // world
    w, b = O.glorot_init(sizes, seed=3)
    x, t = O.synth_data(B * nb, sizes[0], sizes[-1], seed=7)
    kw = dict(dropoutflag=1, visible_omit=0.1, hid_omit=0.2)
    g = bp.BP_GPU(1, len(sizes), sizes, B, 1.0, 0.9, 1e-4, w, b, 1, 0.1, 0.2, seed=77, device=lr, world_size=world,
                  rank=rank)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(bp.comm_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    g.comm_init(bytes(idt.cpu().tolist()))
    # this rank's rows of every bunch
    xs = np.concatenate([x[i * B + rank * lb: i * B + (rank + 1) * lb] for i in range(nb)])
    ts = np.concatenate([t[i * B + rank * lb: i * B + (rank + 1) * lb] for i in range(nb)])
    g.train(nb * lb, xs, ts)
    ws, bs = g.returnWeights()
    g.close()
    ok = True
    # replicas must be bit-identical
    for l in range(1, len(sizes)):
        tw = torch.from_numpy(ws[l]).cuda()
        ref = tw.clone()
        dist.broadcast(ref, 0)
        same = bool((tw == ref).all().item())
        if not same:
            ok = False
            print(f"rank {rank}: layer {l} replica differs from rank 0", flush=True)
    if rank == 0:
        s = bp.BP_GPU(1, len(sizes), sizes, B, 1.0, 0.9, 1e-4, w, b, 1, 0.1, 0.2, seed=77, device=lr)
        s.train(B * nb, x, t)
        sw, sb = s.returnWeights()
        s.close()
        for l in range(1, len(sizes)):
            d = float(np.linalg.norm((ws[l] - sw[l]).astype(np.float64)))
            n = float(np.linalg.norm((sw[l] - w[l]).astype(np.float64)))
            print(f"layer {l}: ||W_dp - W_single|| / ||dW_single|| = {d / n:.3e}", flush=True)
            ok &= d <= 2e-3 * n
        print("DP CHECK", "OK" if ok else "FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
