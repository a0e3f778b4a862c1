"""world_size-2 data-parallel arithmetic on CPU (gloo): what each rank of the CUDA path computes, restated with the
oracle.  Rank r owns rows [r*B/W, (r+1)*B/W) of every global bunch, uses the GLOBAL bunch in both scale factors
(2/B and /B) and the global frame index in the dropout key; the sum of the per-rank gradients (an all-reduce) must
reproduce the single-rank step.  With momentum 0 the update is linear in the gradient, so
sum_r (w_r - w0) + w0 is that reduction."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    import oracle_py as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes, B = [24, 32, 20, 9], 16
    lb = B // world
    w, b = O.glorot_init(sizes, seed=3)
    x, t = O.synth_data(B, sizes[0], sizes[-1], seed=5)
    kw = dict(lrate=0.8, momentum=0.0, dropoutflag=1, visible_omit=0.1, hid_omit=0.25, seed=99)
    net = O.Net(sizes, B, weights=w, bias=b, **kw)            # cfg.bunchsize = GLOBAL bunch
    net.train_bunch(x[rank * lb:(rank + 1) * lb], t[rank * lb:(rank + 1) * lb], frame0=rank * lb)
    deltas = []
    for l in range(1, len(sizes)):
        d = torch.from_numpy(net.w[l] - w[l])
        dist.all_reduce(d)                                     # the gradient all-reduce
        deltas.append(d.numpy() + w[l])
    if rank == 0:
        ref = O.Net(sizes, B, weights=w, bias=b, **kw)
        ref.train_bunch(x, t)
        q.put(max(float(np.abs(deltas[l - 1] - ref.w[l]).max()) for l in range(1, len(sizes))))
    dist.destroy_process_group()


def test_two_rank_gradient_sum_equals_single_rank():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    err = q.get(timeout=10)
    assert err < 2e-6, err
