"""Deterministic synthetic Pfile cases shared by tests/golden/make_reader_golden.py (reference reader) and
tests/test_reader.py (our reader)."""
import importlib
import os

import numpy as np

CASES = {
    # fea_dim 129 is the only width at which HEAD's NAT block (literal 129, Interface.cc:777-778) is self-consistent
    "129": dict(dim=129, out=129, ctx=11, off=5, nat=1, seed=11, lens=[30, 12, 5, 40, 25, 11, 10, 60, 33, 18],
                traincache=40, train="0-6", cv="7-9", rseed=7, hidden=4),
    "129b": dict(dim=129, out=129, ctx=11, off=5, nat=1, seed=5, lens=[50, 23, 75, 14, 90, 31], traincache=1000,
                 train="0-3", cv="4-5", rseed=123, hidden=3),
}


def _tools():
    return importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")


def make_inputs(d, case):
    T = _tools()
    rng = np.random.default_rng(case["seed"])
    dim, out = case["dim"], case["out"]
    mu = (-6 + 3 * np.sin(np.arange(dim) / 40.0)).astype(np.float32)
    ivar = np.full(dim, 1 / 1.5, dtype=np.float32)
    feas = [(rng.standard_normal((n, dim), dtype=np.float32) * 1.5 + mu).astype(np.float32) for n in case["lens"]]
    targs = [rng.standard_normal((n, out), dtype=np.float32) for n in case["lens"]]
    T.write_pfile(os.path.join(d, "fea.pfile"), feas)
    T.write_pfile(os.path.join(d, "targ.pfile"), targs)
    T.write_norm(os.path.join(d, "fea.norm"), mu, ivar)
    return feas, targs, mu, ivar


def layersizes(case):
    return [case["dim"] * case["ctx"] + (case["dim"] if case["nat"] else 0), case["hidden"], case["out"]]


def reader_args(d, case):
    ls = ",".join(str(s) for s in layersizes(case))
    return [f"fea_file={d}/fea.pfile", f"norm_file={d}/fea.norm", f"targ_file={d}/targ.pfile",
            f"outwts_file={d}/out.wts", f"log_file={d}/log.txt", "initwts_file=",
            f"train_sent_range={case['train']}", f"cv_sent_range={case['cv']}", f"fea_dim={case['dim']}",
            f"fea_context={case['ctx']}", f"targ_offset={case['off']}", f"traincache={case['traincache']}",
            "bunchsize=8", f"layersizes={ls}", "gpu_used=1", f"init_randem_seed={case['rseed']}", "momentum=0.5",
            "weightcost=0", "lrate=1", "dropoutflag=0", "visible_omit=0", "hid_omit=0"]


def parse_dump(path, case):
    ls = layersizes(case)
    raw = np.fromfile(path, dtype=np.uint8)
    pos = 0
    chunks = []
    for kind in (0, 1):
        n = int(raw[pos:pos + 4].view(np.int32)[0]); pos += 4
        for _ in range(n):
            cid, s = (int(v) for v in raw[pos:pos + 8].view(np.int32)); pos += 8
            x = raw[pos:pos + 4 * s * ls[0]].view(np.float32).reshape(s, ls[0]).copy(); pos += 4 * s * ls[0]
            t = raw[pos:pos + 4 * s * ls[-1]].view(np.float32).reshape(s, ls[-1]).copy(); pos += 4 * s * ls[-1]
            chunks.append((kind, cid, x, t))
    assert pos == raw.size
    return chunks
