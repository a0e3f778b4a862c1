"""GPU parity tests: the CUDA path (through the C-ABI, via the Python mirror of BP_GPU) against the CPU oracle.

Tolerances (stated once, used everywhere):
  * integer / bit work (dropout masks, SGD update given identical gradients): bit-exact.
  * a single GEMM against float64 products of TF32-truncated operands (the tensor core truncates, measured by
    tests/gpu_probe_gemm.py): |err| <= 2e-4 * scale, scale = RMS magnitude of the compared tensor — what remains is
    summation order / the tensor core's internal accumulation.
  * the whole path against the oracle in tf32=1 mode: |err| <= 1e-3 * scale.  Looser than one GEMM because a 1-ulp
    difference in an fp32 activation can flip its TF32 truncation (a 2^-10 relative step) in the next layer.
  * against the literal-fp32 oracle (= the reference's cuBLAS-FP32 arithmetic): |err| <= 1e-2 * scale — the price of
    single-pass TF32 (10-bit mantissas, truncation); this is the "stated fp32 tolerance" of BASELINE.json.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TOL_GEMM = 2e-4
TOL_TF32 = 1e-3
TOL_FP32 = 1e-2


def rms(a):
    return float(np.sqrt(np.mean(np.square(a.astype(np.float64)))) + 1e-30)


def assert_fro(got, want, tol, what):
    """Relative Frobenius error.  Used for weight UPDATES of multi-layer nets: a handful of units whose pre-activation
    is within rounding of 0 get ReLU' = 0 on one side and 1 on the other (a discontinuity no tolerance on individual
    elements can absorb — the reference shows the same effect against itself under a different summation order), so
    individual columns may differ by whole terms while the update as a whole agrees."""
    d = np.linalg.norm((got.astype(np.float64) - want.astype(np.float64)).ravel())
    n = np.linalg.norm(want.astype(np.float64).ravel()) + 1e-30
    assert np.isfinite(got).all(), f"{what}: non-finite values"
    assert d <= tol * n, f"{what}: ||err||_F={d:.3e} > {tol:g} * ||want||_F {n:.3e}"


def assert_close(got, want, tol, what):
    err = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))))
    s = rms(want)
    assert np.isfinite(got).all(), f"{what}: non-finite values"
    assert err <= tol * s, f"{what}: max|err|={err:.3e} > {tol:g} * rms {s:.3e}"


def tf32_trunc(x):
    return (np.ascontiguousarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


# ---------------------------------------------------------------------------------------------- kernels in isolation
@pytest.mark.parametrize("kind", [3, 1, 2])
@pytest.mark.parametrize("shape", [(128, 128, 32), (257, 100, 96), (64, 36, 40), (300, 1030, 515), (2049, 257, 1024)])
def test_gemm_against_truncated_float64(bp, kind, shape):
    import ctypes as C
    lib = bp.load_library()
    M, N, K = shape
    rng = np.random.default_rng(M * 7 + N * 3 + K + kind)
    fp = C.POINTER(C.c_float)
    if kind == 3:
        A = rng.standard_normal((K, M), dtype=np.float32)
        B = rng.standard_normal((N, K), dtype=np.float32)
        ref = tf32_trunc(B).astype(np.float64) @ tf32_trunc(A).astype(np.float64)
    elif kind == 1:
        A = rng.standard_normal((M, K), dtype=np.float32)
        B = rng.standard_normal((N, K), dtype=np.float32)
        ref = tf32_trunc(B).astype(np.float64) @ tf32_trunc(A).astype(np.float64).T
    else:
        A = rng.standard_normal((K, M), dtype=np.float32)
        B = rng.standard_normal((K, N), dtype=np.float32)
        ref = tf32_trunc(B).astype(np.float64).T @ tf32_trunc(A).astype(np.float64)
    out = np.full((N, M), np.nan, dtype=np.float32)
    aux = np.ones((N, M), dtype=np.float32)
    rc = lib.bp_debug_gemm(kind, M, N, K, A.ctypes.data_as(fp), A.shape[1], B.ctypes.data_as(fp), B.shape[1],
                           out.ctypes.data_as(fp), M, None, aux.ctypes.data_as(fp) if kind == 1 else None, M, 1.0, 0,
                           0, None)
    assert rc == 0, lib.bp_last_error().decode()
    assert_close(out, ref, TOL_GEMM, f"gemm kind {kind} {shape}")


@pytest.mark.parametrize("kind", [3, 1, 2])
def test_gemm_3xtf32_is_fp32_accurate(bp, kind):
    """Split precision (A*B + A_lo*B + A*B_lo): error against the float64 product of the RAW fp32 operands."""
    import ctypes as C
    lib = bp.load_library()
    M, N, K = 300, 260, 1030
    rng = np.random.default_rng(kind)
    fp = C.POINTER(C.c_float)
    if kind == 3:
        A = rng.standard_normal((K, M), dtype=np.float32); B = rng.standard_normal((N, K), dtype=np.float32)
        ref = B.astype(np.float64) @ A.astype(np.float64)
    elif kind == 1:
        A = rng.standard_normal((M, K), dtype=np.float32); B = rng.standard_normal((N, K), dtype=np.float32)
        ref = B.astype(np.float64) @ A.astype(np.float64).T
    else:
        A = rng.standard_normal((K, M), dtype=np.float32); B = rng.standard_normal((K, N), dtype=np.float32)
        ref = B.astype(np.float64).T @ A.astype(np.float64)
    out = np.full((N, M), np.nan, dtype=np.float32)
    aux = np.ones((N, M), dtype=np.float32)
    rc = lib.bp_debug_gemm(kind, M, N, K, A.ctypes.data_as(fp), A.shape[1], B.ctypes.data_as(fp), B.shape[1],
                           out.ctypes.data_as(fp), M, None, aux.ctypes.data_as(fp) if kind == 1 else None, M, 1.0, 0,
                           bp.BP_MATH_3XTF32, None)
    assert rc == 0, lib.bp_last_error().decode()
    assert_close(out, ref, 2e-4, f"3xTF32 gemm kind {kind}")  # bounded by the tensor core's own accumulation (~3e-5 at K=2048)


def test_sgd_update_bit_exact(bp, oracle):
    import ctypes as C
    lib = bp.load_library()
    fp = C.POINTER(C.c_float)
    rng = np.random.default_rng(11)
    n = 100003
    for (mom, lr, wc, bunch) in [(0.9, 1.0, 0.0, 1024), (0.5, 0.3, 1e-4, 128), (0.0, 1.0, 0.0, 7)]:
        d0 = rng.standard_normal(n, dtype=np.float32) * 1e-2
        w0 = rng.standard_normal(n, dtype=np.float32)
        g = rng.standard_normal(n, dtype=np.float32) * 3
        d1, w1 = d0.copy(), w0.copy()
        oracle.sgd(d1, w1, g, bunch, mom, lr, wc)
        d2, w2 = d0.copy(), w0.copy()
        rc = lib.bp_debug_sgd(n, d2.ctypes.data_as(fp), w2.ctypes.data_as(fp), g.ctypes.data_as(fp), bunch, mom, lr, wc)
        assert rc == 0, lib.bp_last_error().decode()
        assert np.array_equal(d1, d2) and np.array_equal(w1, w2)


# ---------------------------------------------------------------------------------------------- whole path
def make_pair(bp, oracle, sizes, bunch, tf32, **kw):
    w, b = oracle.glorot_init(sizes, seed=kw.pop("init_seed", 3))
    onet = oracle.Net(sizes, bunch, tf32=tf32, weights=w, bias=b, **kw)
    g = bp.BP_GPU(1, len(sizes), sizes, bunch, kw.get("lrate", 1.0), kw.get("momentum", 0.0),
                  kw.get("weightcost", 0.0), w, b, kw.get("dropoutflag", 0), kw.get("visible_omit", 0.0),
                  kw.get("hid_omit", 0.0), activation=kw.get("activation", 0), seed=kw.get("seed", 0x5EED5EED),
                  device=0)
    return onet, g


@pytest.mark.parametrize("act", [0, 1])
def test_train_matches_oracle(bp, oracle, act):
    sizes = [75, 96, 130, 33]
    x, t = oracle.synth_data(5 * 32 + 7, sizes[0], sizes[-1], seed=21)
    kw = dict(lrate=0.7, momentum=0.9, weightcost=1e-4, activation=act)
    o_tf, g = make_pair(bp, oracle, sizes, 32, 1, **kw)
    o_fp, _g2 = make_pair(bp, oracle, sizes, 32, 0, **kw)
    _g2.close()
    g.train(x.shape[0], x, t)
    o_tf.train(x.shape[0], x, t)
    o_fp.train(x.shape[0], x, t)
    ws, bs = g.returnWeights()
    for l in range(1, len(sizes)):
        # compare the CHANGE of the weights (the part the path computes), scaled by its own rms
        dw_tf = o_tf.w[l] - o_tf.dw[l] * 0  # weights themselves
        assert_close(ws[l], dw_tf, TOL_TF32, f"W{l} vs tf32 oracle (act {act})")
        assert_close(bs[l], o_tf.b[l], 2 * TOL_TF32, f"b{l} vs tf32 oracle (act {act})")
        assert_close(ws[l], o_fp.w[l], TOL_FP32, f"W{l} vs fp32 oracle (act {act})")
    assert g.counters()[1] == 5  # trailing 7 frames skipped (BP_GPU.cu:315-318)
    g.close()


def test_split_k_output_layer_matches_oracle(bp, oracle):
    """A wide last hidden layer makes the runtime cut the output product's K range into slices
    (bp_runtime.cu out_layer_splits) and finish it in bp_out_finish_kernel: train, loss, CV and forward must still
    agree with the oracle (BP_GPU.cu:560-575 / kernSubClean DevFunc.cu:263)."""
    sizes = [75, 600, 33]   # 10 k-blocks of 64 -> 2 slices
    x, t = oracle.synth_data(4 * 64 + 9, sizes[0], sizes[-1], seed=33)
    kw = dict(lrate=0.5, momentum=0.9, weightcost=0.0)
    o_tf, g = make_pair(bp, oracle, sizes, 64, 1, **kw)
    g.train(x.shape[0], x, t)
    o_tf.train(x.shape[0], x, t)
    ws, bs = g.returnWeights()
    for l in range(1, len(sizes)):
        assert_close(ws[l], o_tf.w[l], TOL_TF32, f"split-K W{l}")
        assert_close(bs[l], o_tf.b[l], 2 * TOL_TF32, f"split-K b{l}")
    out = g.forward(x.shape[0], x)
    assert_close(out, o_tf.forward(x), TOL_TF32, "split-K forward")
    cv = g.CrossValid(x.shape[0], x, t)
    ref = o_tf.crossvalid(x, t)
    assert abs(cv - ref) <= 2e-3 * abs(ref), (cv, ref)
    g.close()


def test_3xtf32_training_matches_literal_fp32_oracle(bp, oracle):
    """BP_MATH_3XTF32 against the literal-fp32 oracle (= the reference's cuBLAS-FP32 arithmetic) at fp32-class
    tolerance: 1e-4 * rms on the weights, 5e-3 relative Frobenius on the update (ReLU' flips, see assert_fro)."""
    sizes = [75, 96, 130, 33]
    x, t = oracle.synth_data(5 * 32, sizes[0], sizes[-1], seed=21)
    w, b = oracle.glorot_init(sizes, seed=3)
    kw = dict(lrate=0.7, momentum=0.9, weightcost=1e-4)
    o = oracle.Net(sizes, 32, weights=w, bias=b, dropoutflag=1, visible_omit=0.1, hid_omit=0.2, seed=5, **kw)
    g = bp.BP_GPU(1, len(sizes), sizes, 32, kw["lrate"], kw["momentum"], kw["weightcost"], w, b, 1, 0.1, 0.2, seed=5,
                  device=0, math_mode=bp.BP_MATH_3XTF32)
    g.train(x.shape[0], x, t)
    o.train(x.shape[0], x, t)
    ws, bs = g.returnWeights()
    for l in range(1, len(sizes)):
        assert_close(ws[l], o.w[l], 1e-4, f"3xTF32 W{l}")
        assert_fro(ws[l] - w[l], o.w[l] - w[l], 5e-3, f"3xTF32 dW{l}")
    out = g.forward(100, x[:100])
    assert_close(out, o.forward(x[:100]), 2e-4, "3xTF32 forward (keep-scaled)")
    g.close()


def test_forward_crossvalid_partial_bunch_and_keep(bp, oracle):
    sizes = [60, 140, 257]
    x, t = oracle.synth_data(3 * 64 + 19, sizes[0], sizes[-1], seed=5)
    # no dropout: against the tf32-conditioned oracle.  With dropoutflag the reference scales the WEIGHTS by keep in
    # fp32 before the product (BP_GPU.cu:726-732) while we scale the accumulator, so the truncation points differ and
    # the comparison is against the literal fp32 oracle at the fp32 tolerance.
    for kw, mode, tol in ((dict(), 1, TOL_TF32), (dict(dropoutflag=1, visible_omit=0.1, hid_omit=0.2), 0, TOL_FP32)):
        o_tf, g = make_pair(bp, oracle, sizes, 64, mode, **kw)
        out = g.forward(x.shape[0], x)
        ref = o_tf.forward(x)
        assert_close(out, ref, tol, f"forward {kw}")
        cv = g.CrossValid(x.shape[0], x, t)
        assert abs(cv - o_tf.crossvalid(x, t)) <= tol * abs(cv)
        ws0, _ = g.returnWeights()
        g.CrossValid(x.shape[0], x, t)
        ws1, _ = g.returnWeights()
        assert all(np.array_equal(a, b) for a, b in zip(ws0[1:], ws1[1:]))  # CV never touches the weights
        g.close()


def test_dropout_training_matches_oracle_masks(bp, oracle):
    sizes = [40, 64, 48, 20]
    x, t = oracle.synth_data(4 * 32, sizes[0], sizes[-1], seed=8)
    kw = dict(lrate=1.0, momentum=0.5, dropoutflag=1, visible_omit=0.1, hid_omit=0.2, seed=987654321)
    o_tf, g = make_pair(bp, oracle, sizes, 32, 1, **kw)
    for (step, layer, f, u) in [(0, 0, 0, 0), (3, 1, 17, 5), (2, 2, 31, 47), (1, 0, 13, 39)]:
        assert bp.dropout_mask(kw["seed"], step, layer, f, u, 0.2) == oracle.dropout_mask(kw["seed"], step, layer, f,
                                                                                           u, 0.2)
    g.train(x.shape[0], x, t)
    o_tf.train(x.shape[0], x, t)
    ws, bs = g.returnWeights()
    for l in range(1, len(sizes)):
        assert_close(ws[l], o_tf.w[l], TOL_TF32, f"dropout W{l}")
    g.close()


@pytest.mark.parametrize("name", ["tiny_train.npz", "tiny_dropout.npz"])
def test_golden_fixture(bp, name):
    gf = np.load(os.path.join(GOLDEN, name))
    sizes = [int(s) for s in gf["sizes"]]
    L = len(sizes)
    w = [None] + [gf[f"w{i}"] for i in range(1, L)]
    b = [None] + [gf[f"b{i}"] for i in range(1, L)]
    g = bp.BP_GPU(1, L, sizes, int(gf["bunch"]), float(gf["lrate"]), float(gf["momentum"]), float(gf["weightcost"]),
                  w, b, int(gf["dropoutflag"]), float(gf["visible_omit"]), float(gf["hid_omit"]),
                  seed=int(gf["seed"]), device=0)
    g.train(gf["x"].shape[0], gf["x"], gf["t"])
    ws, bs = g.returnWeights()
    for i in range(1, L):
        assert_close(ws[i], gf[f"w{i}_out"], TOL_FP32, f"{name} W{i}")
    out = g.forward(gf["x"].shape[0], gf["x"])
    assert_close(out, gf["fwd_out"], 2 * TOL_FP32, f"{name} forward after training")
    g.close()


# ---------------------------------------------------------------------------------------------- full-size properties
C2 = [2827, 2048, 2048, 2048, 257]


def test_full_size_one_bunch_vs_oracle(bp, oracle):
    """One C2 bunch (B=1024) against the oracle (tf32 mode) — ~80 GFLOP on the host cores."""
    x, t = oracle.synth_data(1024, C2[0], C2[-1], seed=2)
    kw = dict(lrate=1.0, momentum=0.9)
    o_tf, g = make_pair(bp, oracle, C2, 1024, 1, **kw)
    g.train(1024, x, t)
    o_tf.train(1024, x, t)
    ws, bs = g.returnWeights()
    w0, _ = oracle.glorot_init(C2, seed=3)
    for l in range(1, len(C2)):
        assert_fro(ws[l] - w0[l], o_tf.w[l] - w0[l], 2e-2, f"C2 delta W{l}")  # the update itself, not w
        assert_close(ws[l], o_tf.w[l], TOL_TF32, f"C2 W{l}")
    g.close()


def test_full_size_properties(bp, oracle):
    """Size-independent properties at BASELINE.json's C2 shape: determinism, lr=0 idempotence, partial-bunch rule."""
    x, t = oracle.synth_data(2 * 1024 + 100, C2[0], C2[-1], seed=4)
    w, b = oracle.glorot_init(C2, seed=3)

    def run(lr, n):
        g = bp.BP_GPU(1, len(C2), C2, 1024, lr, 0.9, 0.0, w, b, device=0)
        g.train(n, x[:n], t[:n])
        ws, bs = g.returnWeights()
        out = g.forward(256, x[:256])
        g.close()
        return ws, bs, out

    a = run(1.0, 2 * 1024 + 100)
    bb = run(1.0, 2 * 1024)          # trailing 100 frames must be ignored
    c = run(0.0, 2 * 1024)           # lr = 0: weights unchanged bit-for-bit
    for l in range(1, len(C2)):
        assert np.array_equal(a[0][l], bb[0][l]) and np.array_equal(a[1][l], bb[1][l])
        assert np.array_equal(c[0][l], w[l]) and np.array_equal(c[1][l], b[l])
    assert np.array_equal(a[2], bb[2])
    assert not np.array_equal(a[0][1], w[1])
