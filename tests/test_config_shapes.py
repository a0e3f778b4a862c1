"""Parity at the SHAPES of BASELINE.json's configs (C2..C5) — the shapes that pick the 256-wide
CTA-pair kernels, the split-K output layer and the 7-layer net — against (a) the CPU oracle on the same seeded inputs
and (b) the reference's own CUDA trainer (class BP_GPU of /root/reference compiled unmodified into
oracle/_ref/ref_harness, cuBLAS FP32) at C2 size, in both math modes.  Every comparison records its ACHIEVED error
(parity_log) and is bounded at <= 2x what the B200 measured (profiles/r2_parity_errors.json).

What is compared: the weight UPDATE of one or two bunches (W_after - W_before, relative Frobenius norm: the part the
path computes; W itself is dominated by the unchanged initial weights) and the decode output (max |err| / rms)."""
import os

import numpy as np
import pytest

from test_reference_parity import HARNESS, run_reference

pytestmark = pytest.mark.gpu

C2 = [2827, 2048, 2048, 2048, 257]
C3 = [3084, 2048, 2048, 2048, 257]
C4 = [2827, 2048, 2048, 2048, 2048, 2048, 257]


def fro(a):
    return float(np.linalg.norm(np.asarray(a, np.float64).ravel()))


def rms(a):
    return float(np.sqrt(np.mean(np.square(np.asarray(a, np.float64)))) + 1e-30)


def rel_update_err(got, want, w0):
    return fro((got - w0) - (want - w0)) / (fro(want - w0) + 1e-30)


def test_c3_shape_dropout_bunch_vs_oracle(bp, oracle, parity_log):
    """C3: 3084-wide (NAT) input, dropout 0.2 / 0.2, bunch 2048 -> 256-wide pair kernels for forward / dX, Philox masks
    replayed by the oracle (BP_GPU.cu:484-673 with :534-551 masks)."""
    B = 2048
    x, t = oracle.synth_data(B, C3[0], C3[-1], seed=41)
    w, b = oracle.glorot_init(C3, seed=3)
    kw = dict(lrate=1.0, momentum=0.9, dropoutflag=1, visible_omit=0.2, hid_omit=0.2, seed=2024)
    o = oracle.Net(C3, B, tf32=1, weights=w, bias=b, **kw)
    g = bp.BP_GPU(1, len(C3), C3, B, 1.0, 0.9, 0.0, w, b, 1, 0.2, 0.2, seed=2024, device=0)
    g.train(B, x, t)
    o.train(B, x, t)
    ws, bs = g.returnWeights()
    g.close()
    for l in range(1, len(C3)):
        parity_log(f"C3 dW{l} vs tf32 oracle (rel. Frobenius)", rel_update_err(ws[l], o.w[l], w[l]), BOUND["c3_dw"])
        parity_log(f"C3 W{l} vs tf32 oracle (max/rms)", np.abs(ws[l] - o.w[l]).max() / rms(o.w[l]), BOUND["c3_w"])
        # masks really were applied: a dropped unit's fan-out... is covered by the bit-exact mask tests; here: it trained
        assert not np.array_equal(ws[l], w[l])


def test_c5_shape_forward_vs_oracle(bp, oracle, parity_log):
    """C5: forward decode of one 8192-frame batch through the C2 net (BP_GPU.cu:676-773), 256-wide pair kernels."""
    n = 8192
    x, _t = oracle.synth_data(n, C2[0], C2[-1], seed=43)
    w, b = oracle.glorot_init(C2, seed=3)
    o_tf = oracle.Net(C2, n, tf32=1, weights=w, bias=b)
    o_fp = oracle.Net(C2, n, tf32=0, weights=w, bias=b)
    g = bp.BP_GPU(1, len(C2), C2, n, 1.0, 0.9, 0.0, w, b, device=0)
    out = g.forward(n, x)
    g.close()
    ref_tf, ref_fp = o_tf.forward(x), o_fp.forward(x)
    assert np.isfinite(out).all()
    e_tf, e_fp = np.abs(out - ref_tf).max() / rms(ref_tf), np.abs(out - ref_fp).max() / rms(ref_fp)
    print(f"[parity] C5 decode: vs tf32 oracle {e_tf:.3e}, vs literal-fp32 oracle {e_fp:.3e} (max |err| / rms)")
    parity_log("C5 decode vs tf32 oracle (max/rms)", e_tf, BOUND["c5_tf32"])
    parity_log("C5 decode vs literal-fp32 oracle (max/rms)", e_fp, BOUND["c5_fp32"])
    parity_log("C5 decode vs literal-fp32 oracle (rel. Frobenius)", fro(out - ref_fp) / fro(ref_fp), BOUND["c5_fp32_fro"])


def test_c4_shape_local_bunch_vs_oracle(bp, oracle, parity_log):
    """C4's net (5 hidden layers) at one rank's share of the global bunch (512 of 4096 frames) — the 1-rank
    arithmetic every data-parallel rank runs; the exchange itself is covered by bench.py's dp_parity and
    tests/test_dp_gloo.py."""
    B = 512
    x, t = oracle.synth_data(2 * B, C4[0], C4[-1], seed=45)
    w, b = oracle.glorot_init(C4, seed=3)
    o = oracle.Net(C4, B, lrate=1.0, momentum=0.9, tf32=1, weights=w, bias=b)
    g = bp.BP_GPU(1, len(C4), C4, B, 1.0, 0.9, 0.0, w, b, device=0)
    g.train(2 * B, x, t)
    o.train(2 * B, x, t)
    ws, bs = g.returnWeights()
    g.close()
    for l in range(1, len(C4)):
        parity_log(f"C4 dW{l} vs tf32 oracle (rel. Frobenius)", rel_update_err(ws[l], o.w[l], w[l]), BOUND["c4_dw"])


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="oracle/_ref/ref_harness not built (needs /root/reference)")
@pytest.mark.parametrize("math", ["tf32", "3xtf32"])
def test_c2_size_vs_reference_cuda_binary(bp, oracle, parity_log, math):
    """Ours vs the reference's unmodified BP_GPU::train / CrossValid (cuBLAS FP32) at C2 size: two bunches of 1024 +
    a CV chunk with a ragged tail, identical blobs.  tf32 = the headline mode; 3xtf32 = the equal-precision mode."""
    B = 1024
    x, t = oracle.synth_data(2 * B, C2[0], C2[-1], seed=51)
    xcv, tcv = oracle.synth_data(B + 77, C2[0], C2[-1], seed=52)
    w, b = oracle.glorot_init(C2, seed=3)
    hp = dict(lrate=1.0, momentum=0.9, weightcost=0.0)
    rw, rb, rcv, _ms = run_reference(C2, B, x, t, xcv, tcv, w, b, **hp)
    g = bp.BP_GPU(1, len(C2), C2, B, 1.0, 0.9, 0.0, w, b, device=0,
                  math_mode=bp.BP_MATH_3XTF32 if math == "3xtf32" else bp.BP_MATH_TF32)
    g.train(2 * B, x, t)
    gw, gb = g.returnWeights()
    gcv = g.CrossValid(xcv.shape[0], xcv, tcv)
    g.close()
    for l in range(1, len(C2)):
        parity_log(f"C2 {math} dW{l} vs reference CUDA (rel. Frobenius)", rel_update_err(gw[l], rw[l], w[l]),
                   BOUND[f"c2_ref_dw_{math}"])
        parity_log(f"C2 {math} W{l} vs reference CUDA (max/rms)", np.abs(gw[l] - rw[l]).max() / rms(rw[l]),
                   BOUND[f"c2_ref_w_{math}"])
        parity_log(f"C2 {math} db{l} vs reference CUDA (rel. Frobenius)", rel_update_err(gb[l], rb[l], b[l]),
                   BOUND[f"c2_ref_dw_{math}"])
    parity_log(f"C2 {math} CV score vs reference CUDA (relative)", abs(gcv - rcv) / abs(rcv), BOUND[f"c2_ref_cv_{math}"])


# Bounds: first GPU run of this file uses the stated tolerances of tests/test_gpu_parity.py; they are then tightened to
# <= 2x the achieved values recorded in profiles/r2_parity_errors.json.
BOUND = {
    "c3_dw": 4e-2, "c3_w": 5e-6,          # achieved 2.0e-2 (layer 1) / 1.3e-6
    "c5_tf32": 3.5e-3, "c5_fp32": 3e-2, "c5_fp32_fro": 5e-3,   # achieved 1.7e-3 / (max over 2.1 M outputs) / -
    "c4_dw": 5e-2,                        # achieved 2.6e-2 (layer 1 of 6)
    "c2_ref_dw_tf32": 5e-2, "c2_ref_w_tf32": 2.5e-5, "c2_ref_cv_tf32": 1e-4,        # achieved 2.8e-2 / 1.1e-5 / 3.5e-5
    "c2_ref_dw_3xtf32": 6e-3, "c2_ref_w_3xtf32": 1e-5, "c2_ref_cv_3xtf32": 1e-4,    # achieved 2.9e-3 / 4.3e-6 / 4.9e-5
}
# Why the UPDATE of one bunch agrees only to a few percent in single-pass TF32: on these synthetic inputs (random-init
# net, noise-like targets) the gradient sums cancel heavily — the literal-fp32 oracle itself moves by 0.6 % (layer 1)
# when its accumulators are switched from float to double, the tf32-conditioned oracle by 1.5 %, and tf32 vs fp32
# differ by 3 % (measured with oracle/bp_oracle.c at the C3 shape).  3xTF32 is at the fp32 level (0.3 %).
