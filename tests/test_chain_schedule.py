"""Host logic of the chained launches (csrc/bp_chain_params.h: chain_shapes + chain_schedule), no GPU needed.

The device side relies on two properties of the schedule, whatever the real timing turns out to be:
  (1) every tile of every product is scheduled exactly once;
  (2) the union of "a pair processes its items in order" and "a tile waits for the tiles it depends on" is ACYCLIC —
      otherwise the persistent CTA pairs, which spin on each other's completion counters, deadlock.
Both are checked here for the BASELINE configs (C2, C3, C4 at 512 and 4096 / N rows, C5-shaped forward) and for edge
nets (one layer, nine layers, tiny bunches, one pair, 3xTF32 passes)."""
import ctypes as C
from collections import defaultdict

import numpy as np
import pytest

C2 = [2827, 2048, 2048, 2048, 257]
C3 = [3084, 2048, 2048, 2048, 257]
C4 = [2827, 2048, 2048, 2048, 2048, 2048, 257]
CASES = [("C2", C2, 1024, 74, 1), ("C3", C3, 2048, 74, 1), ("C4 per-GPU share", C4, 512, 74, 1),
         ("C4 on 2 GPUs", C4, 2048, 74, 1), ("C2 3xTF32", C2, 1024, 74, 3), ("one layer", [129, 65], 32, 74, 1),
         ("nine layers", [64, 72, 40, 96, 33, 80, 48, 56, 64, 20], 32, 74, 1), ("ragged", [75, 96, 33], 37, 74, 1),
         ("one pair", [300, 512, 129], 256, 1, 1), ("three pairs", C2, 1024, 3, 1), ("bunch 128", C2, 128, 74, 1)]


def plan(bp, sizes, rows, pairs, which, passes=1):
    lib = bp.load_library()
    f = lib.bp_debug_chain_plan
    ip = C.POINTER(C.c_int)
    f.argtypes = [C.c_int, ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, ip, ip, ip, C.c_int, ip, ip, ip, ip,
                  ip, ip, ip, C.POINTER(C.c_longlong)]
    mi, mp = 4096, 32
    arr = lambda n: np.zeros(n, np.int32)
    pair, prod, mt, nt = arr(mi), arr(mi), arr(mi), arr(mi)
    dep, dall, mts, nts, pn, nc = arr(mp), arr(mp), arr(mp), arr(mp), arr(mp), arr(mp)
    n_items, n_prods, span = C.c_int(0), C.c_int(0), C.c_longlong(0)
    ls = np.asarray(sizes, np.int32)
    p = lambda a: a.ctypes.data_as(ip)
    rc = f(len(sizes), p(ls), rows, pairs, which, passes, mi, C.byref(n_items), p(pair), p(prod), p(mt), p(nt), mp,
           C.byref(n_prods), p(dep), p(dall), p(mts), p(nts), p(pn), p(nc), C.byref(span))
    assert rc == 0, lib.bp_last_error().decode()
    n, q = n_items.value, n_prods.value
    return (dict(pair=pair[:n], prod=prod[:n], mt=mt[:n], nt=nt[:n]),
            dict(dep=dep[:q], dep_all=dall[:q], m_tiles=mts[:q], n_tiles=nts[:q], pair_n=pn[:q], n_cols=nc[:q]),
            span.value)


@pytest.mark.parametrize("which", [0, 1], ids=["forward", "back-propagation"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_schedule_is_complete_and_deadlock_free(bp, case, which):
    _name, sizes, rows, pairs, passes = case
    it, pr, span = plan(bp, sizes, rows, pairs, which, passes)
    n = len(it["prod"])
    # (1) every tile exactly once
    want = {(q, m, t) for q in range(len(pr["dep"])) for m in range(pr["m_tiles"][q]) for t in range(pr["n_tiles"][q])}
    got = list(zip(it["prod"].tolist(), it["mt"].tolist(), it["nt"].tolist()))
    assert len(got) == len(want) and set(got) == want
    assert span > 0 and it["pair"].max() < pairs
    # (2) acyclic: node = item; edges: previous item of the same pair -> item; every dependency tile -> item
    index = {g: i for i, g in enumerate(got)}
    succ = defaultdict(list)
    indeg = [0] * n
    last_of_pair = {}
    for i in range(n):        # items are listed pair by pair, in each pair's processing order
        p = int(it["pair"][i])
        if p in last_of_pair:
            succ[last_of_pair[p]].append(i)
            indeg[i] += 1
        last_of_pair[p] = i
    for i, (q, m, t) in enumerate(got):
        d = int(pr["dep"][q])
        if d < 0:
            continue
        if pr["dep_all"][q]:
            cols = range(pr["n_tiles"][d])
        else:
            f0, f1 = t * pr["pair_n"][q], min(pr["n_cols"][q], (t + 1) * pr["pair_n"][q])
            cols = range(f0 // pr["pair_n"][d], -(-f1 // pr["pair_n"][d]))
        for c in cols:
            for mm in range(pr["m_tiles"][d]):
                j = index[(d, mm, c)]
                succ[j].append(i)
                indeg[i] += 1
    ready = [i for i in range(n) if indeg[i] == 0]
    done = 0
    while ready:
        i = ready.pop()
        done += 1
        for j in succ[i]:
            indeg[j] -= 1
            if indeg[j] == 0:
                ready.append(j)
    assert done == n, "pair order + dependencies contain a cycle: the device would deadlock"


def test_c2_plan_shape(bp):
    """C2 at bunch 1024 on 74 pairs: 64 tiles of 256 x 128 per hidden-layer product, 16 for the 257-wide output layer;
    back-propagation: the dX chain first (priority), 256-wide dW tiles as filler."""
    it, pr, _ = plan(bp, C2, 1024, 74, 0)
    assert pr["pair_n"].tolist() == [128, 128, 128, 128] and (pr["m_tiles"] * pr["n_tiles"]).tolist() == [64, 64, 64, 16]
    assert pr["dep"].tolist() == [-1, 0, 1, 2]
    it, pr, _ = plan(bp, C2, 1024, 74, 1)
    assert pr["pair_n"].tolist() == [128, 128, 128, 256, 256, 256, 256]
    assert pr["dep"].tolist() == [-1, 0, 1, -1, 0, 1, 2] and pr["dep_all"].tolist() == [0, 0, 0, 1, 1, 1, 1]
    assert (pr["m_tiles"] * pr["n_tiles"]).tolist() == [64, 64, 64, 18, 72, 72, 96]
