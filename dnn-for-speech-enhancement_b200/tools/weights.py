"""Weight tooling: Python equivalents of the reference's offline helpers, same files in and out.

  gen     Gen_rand_net      (toolbox/weights/gen_rand_net/Gen_rand_net.cpp:84-101): uniform random weights in
          +-beta*sqrt(6)/sqrt(n_in+n_out) (flag 1, the ReLU recipe of Gen_rand_wts_for_ReLUs_forCudaTrain.pl:
          beta 0.5) or +-beta/sqrt(n_in) (flag 0), zero biases, written as a MAT-v4 `.wts` (:138-170).
  extend  Extend_rand_net   (Extend_rand_net.cpp:199-209, 268-284): widen the layers of a trained net; the old block
          keeps its values at w[m*new_n_out + n], new columns and new rows get fresh random values, new biases are 0.
  tomat   change_cudaSavedModels2matlabWeigths_4layers.m:10-26: `.wts` -> se_weights.mat with w_k = [W_k ; b_k]
          ((n_in+1) x n_out, the layout the external MATLAB decoder multiplies [x 1] with).

The reference draws from libc rand() seeded with time(NULL); these tools take an explicit seed (numpy PCG64), so files are
reproducible but not bit-equal to any particular reference run — only the distribution and the file format matter.

  python -m dnn-for-speech-enhancement_b200.tools.weights gen 2827,2048,2048,2048,257 out.wts [--flag 1 --beta 0.5 --seed 3]
  python -m dnn-for-speech-enhancement_b200.tools.weights extend in.wts 1548,2048,2048,129 0,512,512,0 out.wts
  python -m dnn-for-speech-enhancement_b200.tools.weights tomat mlp.25.wts 1548,2048,2048,129 se_weights25.mat
"""
import argparse

import numpy as np

from .pfile import read_wts, write_wts


def init_range(n_in, n_out, flag=1, beta=0.5):
    """Gen_rand_net.cpp:89-92."""
    if flag:
        return np.float32(beta * np.sqrt(np.float32(6.0)) / np.sqrt(np.float32(n_in + n_out)))
    return np.float32(beta * 1.0 / np.sqrt(np.float32(n_in)))


def gen_rand_net(layersizes, flag=1, beta=0.5, seed=0):
    rng = np.random.default_rng(seed)
    ws, bs = [None], [None]
    for i in range(1, len(layersizes)):
        n_in, n_out = layersizes[i - 1], layersizes[i]
        r = init_range(n_in, n_out, flag, beta)
        ws.append((r * rng.uniform(-1.0, 1.0, size=(n_in, n_out))).astype(np.float32))
        bs.append(np.zeros(n_out, np.float32))
    return ws, bs


def extend_rand_net(ws, bs, ori_layersizes, add_layersizes, beta=0.5, seed=0):
    """Widened copy: layersizes = ori + add (Extend_rand_net.cpp:121-128)."""
    rng = np.random.default_rng(seed)
    new = [o + a for o, a in zip(ori_layersizes, add_layersizes)]
    nws, nbs = [None], [None]
    for i in range(1, len(new)):
        oi, oo, ni, no = ori_layersizes[i - 1], ori_layersizes[i], new[i - 1], new[i]
        assert ws[i].shape == (oi, oo), (ws[i].shape, oi, oo)
        r = init_range(ni, no, 1, beta)
        w = np.zeros((ni, no), np.float32)
        w[:oi, :oo] = ws[i]
        w[:, oo:] = (r * rng.uniform(-1.0, 1.0, size=(ni, no - oo))).astype(np.float32)      # new columns, every row
        w[oi:, :oo] = (r * rng.uniform(-1.0, 1.0, size=(ni - oi, oo))).astype(np.float32)    # new rows, old columns
        b = np.zeros(no, np.float32)
        b[:oo] = bs[i]
        nws.append(w)
        nbs.append(b)
    return nws, nbs, new


def to_matlab_dict(ws, bs):
    """w_k = [weights_{k,k+1} bias_{k+1}']' : MATLAB sees weights_{k,k+1} as n_out x n_in, so w_k is (n_in+1) x n_out."""
    return {f"w{i}": np.vstack([ws[i], bs[i][None, :]]).astype(np.float32) for i in range(1, len(ws))}


def _sizes(s):
    return [int(v) for v in s.split(",")]


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    g = sub.add_parser("gen")
    g.add_argument("layersizes", type=_sizes)
    g.add_argument("out")
    g.add_argument("--flag", type=int, default=1)
    g.add_argument("--beta", type=float, default=0.5)
    g.add_argument("--seed", type=int, default=0)
    e = sub.add_parser("extend")
    e.add_argument("inp")
    e.add_argument("ori_layersizes", type=_sizes)
    e.add_argument("add_layersizes", type=_sizes)
    e.add_argument("out")
    e.add_argument("--beta", type=float, default=0.5)
    e.add_argument("--seed", type=int, default=0)
    m = sub.add_parser("tomat")
    m.add_argument("inp")
    m.add_argument("layersizes", type=_sizes)
    m.add_argument("out")
    a = ap.parse_args(argv)
    if a.cmd == "gen":
        ws, bs = gen_rand_net(a.layersizes, a.flag, a.beta, a.seed)
        write_wts(a.out, ws, bs)
        print(f"wrote {a.out}: layersizes {a.layersizes}")
    elif a.cmd == "extend":
        if len(a.ori_layersizes) != len(a.add_layersizes):
            ap.error("ori_layersizes and add_layersizes must have the same length")
        ws, bs = read_wts(a.inp, a.ori_layersizes)
        nws, nbs, new = extend_rand_net(ws, bs, a.ori_layersizes, a.add_layersizes, a.beta, a.seed)
        write_wts(a.out, nws, nbs)
        print(f"wrote {a.out}: layersizes {new}")
    else:
        from scipy.io import savemat
        ws, bs = read_wts(a.inp, a.layersizes)
        savemat(a.out, to_matlab_dict(ws, bs))
        print(f"wrote {a.out}: " + ", ".join(f"w{i} {ws[i].shape[0] + 1}x{ws[i].shape[1]}" for i in range(1, len(ws))))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
