"""Generates tests/golden/*.npz from the CPU oracle (literal-fp32 mode).  Re-run only when the oracle's arithmetic is
deliberately changed; tests/test_oracle.py::test_golden_fixtures re-derives and compares bit-for-bit, and the GPU
parity tests compare the CUDA path against the same files.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_py as O  # noqa: E402


def make(name, sizes, bunch, n, lrate, momentum, weightcost, dropoutflag=0, visible_omit=0.0, hid_omit=0.0, seed=77):
    w, b = O.glorot_init(sizes, seed=3)
    x, t = O.synth_data(n, sizes[0], sizes[-1], seed=1)
    net = O.Net(sizes, bunch, lrate=lrate, momentum=momentum, weightcost=weightcost, dropoutflag=dropoutflag,
                visible_omit=visible_omit, hid_omit=hid_omit, seed=seed, weights=w, bias=b)
    net.train(n, x, t)
    d = dict(sizes=np.array(sizes), bunch=bunch, lrate=lrate, momentum=momentum, weightcost=weightcost,
             dropoutflag=dropoutflag, visible_omit=visible_omit, hid_omit=hid_omit, seed=seed, x=x, t=t,
             fwd_out=net.forward(x), cv=net.crossvalid(x, t))
    for i in range(1, len(sizes)):
        d[f"w{i}"], d[f"b{i}"] = w[i], b[i]
        d[f"w{i}_out"], d[f"b{i}_out"] = net.w[i], net.b[i]
    np.savez_compressed(os.path.join(HERE, name), **d)
    print(name, "cv", d["cv"])


if __name__ == "__main__":
    make("tiny_train.npz", [33, 40, 24, 9], 16, 70, 0.5, 0.9, 1e-4)
    make("tiny_dropout.npz", [33, 40, 24, 9], 16, 64, 1.0, 0.5, 0.0, 1, 0.1, 0.2)
