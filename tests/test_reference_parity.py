"""GPU tests against the REFERENCE's own CUDA path: class BP_GPU compiled unmodified from /root/reference (plus the
one-token compile fix P1) into oracle/_ref/ref_harness by oracle/build_ref.sh, run here on the B200 with cuBLAS FP32.

Pins two things on identical inputs:
  (1) the CPU oracle's literal-fp32 mode IS the reference's arithmetic   (tight: 2e-5 * rms — summation order only),
  (2) our tcgen05 TF32 path matches the reference within the stated fp32 tolerance (1e-2 * rms, see test_gpu_parity).
"""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def run_reference(sizes, bunch, x, t, xcv, tcv, w, b, lrate, momentum, weightcost, dropoutflag=0, vis=0.0, hid=0.0,
                  reps=1):
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.blob"), os.path.join(d, "out.blob")
        with open(fin, "wb") as f:
            f.write(struct.pack(f"<i{len(sizes)}i5i", len(sizes), *sizes, bunch, x.shape[0], xcv.shape[0],
                                dropoutflag, reps))
            f.write(struct.pack("<5f", lrate, momentum, weightcost, vis, hid))
            for l in range(1, len(sizes)):
                f.write(np.ascontiguousarray(w[l], "<f4").tobytes())
                f.write(np.ascontiguousarray(b[l], "<f4").tobytes())
            for a in (x, t, xcv, tcv):
                f.write(np.ascontiguousarray(a, "<f4").tobytes())
        r = subprocess.run([HARNESS, fin, fout], capture_output=True, text=True, timeout=600, cwd=d)
        assert r.returncode == 0, r.stdout + r.stderr
        raw = np.fromfile(fout, dtype="<f4")
    pos = 0
    ws, bs = [None], [None]
    for l in range(1, len(sizes)):
        n = sizes[l - 1] * sizes[l]
        ws.append(raw[pos:pos + n].reshape(sizes[l - 1], sizes[l]).copy()); pos += n
        bs.append(raw[pos:pos + sizes[l]].copy()); pos += sizes[l]
    return ws, bs, float(raw[pos]), float(raw[pos + 1])


def rms(a):
    return float(np.sqrt(np.mean(np.square(a.astype(np.float64)))) + 1e-30)


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="oracle/_ref/ref_harness not built (needs /root/reference)")
@pytest.mark.parametrize("sizes,bunch", [([1548, 256, 192, 129], 128), ([75, 96, 130, 33], 32)])
def test_reference_cuda_vs_oracle_vs_ours(bp, oracle, sizes, bunch):
    x, t = oracle.synth_data(4 * bunch + 9, sizes[0], sizes[-1], seed=31)
    xcv, tcv = oracle.synth_data(bunch + 17, sizes[0], sizes[-1], seed=32)
    w, b = oracle.glorot_init(sizes, seed=3)
    hp = dict(lrate=1.0, momentum=0.9, weightcost=1e-4)
    rw, rb, rcv, _ms = run_reference(sizes, bunch, x, t, xcv, tcv, w, b, **hp)

    o = oracle.Net(sizes, bunch, weights=w, bias=b, **hp)          # literal fp32 restatement
    o.train(x.shape[0], x, t)
    g = bp.BP_GPU(1, len(sizes), sizes, bunch, hp["lrate"], hp["momentum"], hp["weightcost"], w, b, device=0)
    g.train(x.shape[0], x, t)
    gw, gb = g.returnWeights()
    gcv = g.CrossValid(xcv.shape[0], xcv, tcv)
    g.close()
    for l in range(1, len(sizes)):
        # (1) oracle == reference CUDA path up to fp32 summation order
        e = np.abs(o.w[l] - rw[l]).max()
        assert e <= 2e-5 * rms(rw[l]), f"oracle vs reference W{l}: {e:.3e} rms {rms(rw[l]):.3e}"
        e = np.abs(o.b[l] - rb[l]).max()
        assert e <= 1e-4 * rms(rb[l]) + 1e-7, f"oracle vs reference b{l}: {e:.3e}"
        # (2) ours == reference within the stated tolerance
        e = np.abs(gw[l] - rw[l]).max()
        assert e <= 1e-2 * rms(rw[l]), f"ours vs reference W{l}: {e:.3e} rms {rms(rw[l]):.3e}"
        # ... and the UPDATE itself (w - w0) agrees in relative Frobenius norm (element-wise max is dominated by the
        # few units whose ReLU' flips under rounding-level differences, see tests/test_gpu_parity.py::assert_fro)
        dref, dours, dorc = rw[l] - w[l], gw[l] - w[l], o.w[l] - w[l]
        fro = lambda a: float(np.linalg.norm(a.astype(np.float64).ravel()))
        print(f"layer {l}: ||dW_ours - dW_ref||/||dW_ref|| = {fro(dours - dref) / fro(dref):.3e}   "
              f"||dW_oracle - dW_ref||/||dW_ref|| = {fro(dorc - dref) / fro(dref):.3e}")
        assert fro(dours - dref) <= 5e-2 * fro(dref), f"ours vs reference dW{l}"
        assert fro(dorc - dref) <= 1e-3 * fro(dref), f"oracle vs reference dW{l}"
    ocv = o.crossvalid(xcv, tcv)
    assert abs(ocv - rcv) <= 1e-4 * abs(rcv)
    assert abs(gcv - rcv) <= 1e-2 * abs(rcv)
