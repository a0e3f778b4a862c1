/* bp_oracle.c — CPU restatement of the reference's frame-wise DNN hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this; the
 * product (libbpgpu.so, BPtrain) never links or calls it.
 *
 * The reference has NO CPU compute path (its only implementation is CUDA + cuBLAS), so this file
 * restates, in plain fp32 C, exactly the arithmetic of
 *   BP_GPU::train_bunch_single   BP_GPU.cu:484-673   (forward, back-prop, momentum-SGD update)
 *   BP_GPU::cv_bunch_single      BP_GPU.cu:676-773   (forward-only, keep-scaled weights)
 *   BP_GPU::train / CrossValid   BP_GPU.cu:241-331 / 408-479 (bunch loops, partial-bunch rules, squared error)
 * and the kernel bodies DevFunc.cu:20-45, 67-97, 166-182, 224-277, 313-318 with the GEMM conventions of
 * DevFunc.h:29-67.  Each function cites the lines it follows.
 *
 * PARITY STATUS: the reference ships no golden vectors or tests for this path => "parity unpinned"
 * by the reference's own fixtures.  It is pinned instead by (i) hand-derived known-answer tests in
 * tests/test_oracle.py, (ii) a float64 NumPy restatement, and (iii) runs of the reference's own CUDA binary
 * (oracle/_ref/BPtrain_ref, built by oracle/build_ref.sh) on the GPU box, compared in tests/test_reference_parity.py.
 *
 * Two deliberate, documented deviations from the reference (both unreproducible there):
 *   - dropout masks: the reference draws from cuRAND XORWOW seeded with time(NULL) (BP_GPU.cu:77-78); here masks come
 *     from Philox-4x32-10 keyed on (seed; frame/4, unit, tensor, step) — the same published algorithm the CUDA path
 *     uses, restated independently below.
 *   - `tf32` mode (0 = literal fp32, 1 = truncate, 2 = round-to-nearest-even) pre-conditions GEMM operands the way
 *     the tensor core does, so the CUDA path can be compared tightly; mode 0 is the reference's arithmetic.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXLAYER 10

typedef struct orc_cfg {
  int numlayers;               /* number of layer sizes */
  int layersizes[ORC_MAXLAYER];
  int bunchsize;               /* rows per bunch (global) */
  float lrate, momentum, weightcost;
  int dropoutflag;
  float visible_omit, hid_omit;
  int activation;              /* 0 ReLU (HEAD DevFunc.cu:67-97), 1 sigmoid (commented bodies :52,:62) */
  int tf32;                    /* 0 literal fp32, 1 truncate operands to TF32, 2 round operands to TF32 */
  int accum_double;            /* 1: accumulate GEMM sums in double ("truth" variant) */
  uint64_t seed;               /* dropout seed */
} orc_cfg;

/* ------------------------------------------------------------------------------------------------ Philox-4x32-10 */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  philox4x32_10(c, key[0], key[1]);
  memcpy(out, c, sizeof c);
}
/* cuRAND-style uniform in (0,1]: x*2^-32 + 2^-33 as one fused multiply-add. */
static float u32_uniform(uint32_t x) { return fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

/* 1 = dropped.  tensor 0 = network input, l = output of weight layer l.  kernDropout: DevFunc.cu:34-45 (r < p). */
int orc_dropout_mask(uint64_t seed, uint32_t step, uint32_t tensor, uint32_t frame, uint32_t unit, float p) {
  uint32_t c[4] = {frame >> 2, unit, tensor, step};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return u32_uniform(c[frame & 3]) < p ? 1 : 0;
}

/* ------------------------------------------------------------------------------------------------ helpers */
static float tf32_cond(float x, int mode) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if (mode == 1) u &= 0xFFFFE000u;
  else if (mode == 2) u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
/* kernSigmoid / kernDsigmoid bodies (ReLU at HEAD DevFunc.cu:74-77, 92-95; sigmoid variant :52, :62). */
static float act_f(float x, int act) { return act == 0 ? (x > 0 ? x : 0.0f) : 1.0f / (1.0f + expf(-x)); }
static float dact_f(float y, int act) { return act == 0 ? (y > 0 ? 1.0f : 0.0f) : (1.0f - y) * y; }

/* ------------------------------------------------------------------------------------------------ GEMM core
 * C[m][j] = init[j] (or 0), then C[m][j] = fmaf(A(m,r), Bm(r,j), C[m][j]) for r = 0, 1, ..., R-1 IN THAT ORDER: every
 * output element is the same chain of single-rounded fused multiply-adds a one-thread triple loop would produce (the
 * arithmetic of a textbook fp32 SGEMM).  Only the ORDER OF ELEMENTS is blocked for the cache hierarchy, never the order
 * of the terms of one element: R is cut into blocks of KC whose partial sums pass through C in fp32 (a store + load of
 * a float is exact), B is packed once into NR-wide column panels, A into MR-tall row panels per work item, and an
 * MR x NR register tile walks the panels.  A(m,r) = A[m*as_m + r*as_r], Bm(r,j) = B[r*bs_r + j*bs_j]; both operands are
 * conditioned (tf32 mode) while they are packed.  tests/test_oracle.py pins this routine against the naive loops. */
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define ORC_SIMD 1
#else
#define ORC_SIMD 0
#endif
#define ORC_MR 6
#define ORC_NR 16
#define ORC_KC 256
#define ORC_MC 192 /* rows per work item (multiple of MR) */
#define ORC_NC 256 /* columns per work item (multiple of NR) */

/* register tile: C[MR x NR] += Ap[kc x MR] (r-major) * Bp[kc x NR] (r-major), ascending r */
static inline void micro_tile(int kc, const float* Ap, const float* Bp, float* C, size_t ldc, int mr, int nr) {
#if ORC_SIMD
  if (mr == ORC_MR && nr == ORC_NR) {
    __m256 c[ORC_MR][2];
    for (int m = 0; m < ORC_MR; ++m) {
      c[m][0] = _mm256_loadu_ps(C + m * ldc);
      c[m][1] = _mm256_loadu_ps(C + m * ldc + 8);
    }
    for (int r = 0; r < kc; ++r) {
      const __m256 b0 = _mm256_loadu_ps(Bp + (size_t)r * ORC_NR);
      const __m256 b1 = _mm256_loadu_ps(Bp + (size_t)r * ORC_NR + 8);
#pragma GCC unroll 6
      for (int m = 0; m < ORC_MR; ++m) {
        const __m256 a = _mm256_broadcast_ss(Ap + (size_t)r * ORC_MR + m);
        c[m][0] = _mm256_fmadd_ps(a, b0, c[m][0]);
        c[m][1] = _mm256_fmadd_ps(a, b1, c[m][1]);
      }
    }
    for (int m = 0; m < ORC_MR; ++m) {
      _mm256_storeu_ps(C + m * ldc, c[m][0]);
      _mm256_storeu_ps(C + m * ldc + 8, c[m][1]);
    }
    return;
  }
#endif
  for (int m = 0; m < mr; ++m)
    for (int j = 0; j < nr; ++j) {
      float c = C[m * ldc + j];
      for (int r = 0; r < kc; ++r) c = fmaf(Ap[(size_t)r * ORC_MR + m], Bp[(size_t)r * ORC_NR + j], c);
      C[m * ldc + j] = c;
    }
}

static void gemm_core(int M, int R, int N, const float* A, size_t as_m, size_t as_r, const float* B, size_t bs_r,
                      size_t bs_j, const float* init, float* C, size_t ldc, int tf32) {
  if (M <= 0 || N <= 0) return;
  const int npan = (N + ORC_NR - 1) / ORC_NR;
  const size_t npad = (size_t)npan * ORC_NR;
  const int nkb = (R + ORC_KC - 1) / ORC_KC;
  /* Bp: k-block kb starts at kb*KC*npad; inside it panel p holds [kc][NR] (columns beyond N are zero, never stored) */
  float* Bp = (float*)malloc(((size_t)(R > 0 ? R : 1)) * npad * sizeof(float));
#pragma omp parallel for collapse(2) schedule(static)
  for (int kb = 0; kb < nkb; ++kb)
    for (int p = 0; p < npan; ++p) {
      const int r0 = kb * ORC_KC, kc = (R - r0 < ORC_KC) ? R - r0 : ORC_KC;
      float* dst = Bp + (size_t)r0 * npad + (size_t)p * kc * ORC_NR;
      for (int r = 0; r < kc; ++r)
        for (int j = 0; j < ORC_NR; ++j) {
          const int col = p * ORC_NR + j;
          dst[(size_t)r * ORC_NR + j] = col < N ? tf32_cond(B[(size_t)(r0 + r) * bs_r + (size_t)col * bs_j], tf32) : 0.0f;
        }
    }
  const int nic = (M + ORC_MC - 1) / ORC_MC, njc = (N + ORC_NC - 1) / ORC_NC;
#pragma omp parallel
  {
    float* Ap = (float*)malloc((size_t)ORC_MC * ORC_KC * sizeof(float));
#pragma omp for collapse(2) schedule(dynamic, 1)
    for (int ic = 0; ic < nic; ++ic)
      for (int jc = 0; jc < njc; ++jc) {
        const int m0 = ic * ORC_MC, mc = (M - m0 < ORC_MC) ? M - m0 : ORC_MC;
        const int j0 = jc * ORC_NC, nc = (N - j0 < ORC_NC) ? N - j0 : ORC_NC;
        for (int m = 0; m < mc; ++m) {
          float* c = C + (size_t)(m0 + m) * ldc + j0;
          for (int j = 0; j < nc; ++j) c[j] = init ? init[j0 + j] : 0.0f;
        }
        for (int kb = 0; kb < nkb; ++kb) {
          const int r0 = kb * ORC_KC, kc = (R - r0 < ORC_KC) ? R - r0 : ORC_KC;
          for (int mb = 0; mb < mc; mb += ORC_MR) { /* pack (and condition) this block of A: [mb/MR][r][MR] */
            const int mr = (mc - mb < ORC_MR) ? mc - mb : ORC_MR;
            float* dst = Ap + (size_t)(mb / ORC_MR) * kc * ORC_MR;
            for (int r = 0; r < kc; ++r)
              for (int m = 0; m < ORC_MR; ++m)
                dst[(size_t)r * ORC_MR + m] =
                    m < mr ? tf32_cond(A[(size_t)(m0 + mb + m) * as_m + (size_t)(r0 + r) * as_r], tf32) : 0.0f;
          }
          for (int jb = 0; jb < nc; jb += ORC_NR) {
            const int nr = (nc - jb < ORC_NR) ? nc - jb : ORC_NR;
            const float* bp = Bp + (size_t)r0 * npad + (size_t)((j0 + jb) / ORC_NR) * kc * ORC_NR;
            for (int mb = 0; mb < mc; mb += ORC_MR) {
              const int mr = (mc - mb < ORC_MR) ? mc - mb : ORC_MR;
              micro_tile(kc, Ap + (size_t)(mb / ORC_MR) * kc * ORC_MR, bp, C + (size_t)(m0 + mb) * ldc + j0 + jb, ldc,
                         mr, nr);
            }
          }
        }
      }
    free(Ap);
  }
  free(Bp);
}

/* Same contraction with every sum carried in double ("truth" variant for tolerance studies; small cases only). */
static void gemm_core_dbl(int M, int R, int N, const float* A, size_t as_m, size_t as_r, const float* B, size_t bs_r,
                          size_t bs_j, const float* init, float* C, size_t ldc, int tf32) {
#pragma omp parallel for schedule(static)
  for (int m = 0; m < M; ++m)
    for (int j = 0; j < N; ++j) {
      double acc = 0.0;
      for (int r = 0; r < R; ++r)
        acc += (double)tf32_cond(A[(size_t)m * as_m + (size_t)r * as_r], tf32) *
               (double)tf32_cond(B[(size_t)r * bs_r + (size_t)j * bs_j], tf32);
      C[(size_t)m * ldc + j] = (float)(acc + (init ? (double)init[j] : 0.0));
    }
}

/* Test hooks (tests/test_oracle.py): the blocked routine and the one-thread triple loop it must equal bit for bit.
 * A is M x R, B is R x N, both row-major and dense. */
void orc_gemm(int M, int R, int N, const float* A, const float* B, const float* init, float* C, int tf32) {
  gemm_core(M, R, N, A, (size_t)R, 1, B, (size_t)N, 1, init, C, (size_t)N, tf32);
}
void orc_gemm_naive(int M, int R, int N, const float* A, const float* B, const float* init, float* C, int tf32) {
  for (int m = 0; m < M; ++m)
    for (int j = 0; j < N; ++j) {
      float c = init ? init[j] : 0.0f;
      for (int r = 0; r < R; ++r)
        c = fmaf(tf32_cond(A[(size_t)m * R + r], tf32), tf32_cond(B[(size_t)r * N + j], tf32), c);
      C[(size_t)m * N + j] = c;
    }
}

/* X[B x N] = bias (kernMultiCopy, DevFunc.cu:166-182), then X += Y[B x K] * W[K x N] (SgemmNN, BP_GPU.cu:557-558).
 * W is the reference's w[in*N + out].  Sum over k in ascending order, one fma per term. */
static void gemm_fwd(int B, int K, int N, const float* Y, const float* W, const float* bias, float* X, int tf32,
                     int dbl) {
  (dbl ? gemm_core_dbl : gemm_core)(B, K, N, Y, (size_t)K, 1, W, (size_t)N, 1, bias, X, (size_t)N, tf32);
}

/* E[B x K] = D[B x N] * W^T (SgemmTN with beta 0, BP_GPU.cu:636): sum over n ascending. */
static void gemm_dx(int B, int K, int N, const float* D, const float* W, float* E, int tf32, int dbl) {
  (dbl ? gemm_core_dbl : gemm_core)(B, N, K, D, (size_t)N, 1, W, 1, (size_t)N, NULL, E, (size_t)K, tf32);
}

/* G[K x N] = Y^T[K x B] * D[B x N] (SgemmNT with beta 0, BP_GPU.cu:642): sum over frames ascending; gb[n] =
 * sum_f D[f][n] in frame order (kernAccSumrow with alpha 0, beta 1: BP_GPU.cu:647, DevFunc.cu:234-240). */
static void gemm_dw(int B, int K, int N, const float* Y, const float* D, float* G, float* gb, int tf32, int dbl) {
  (dbl ? gemm_core_dbl : gemm_core)(K, B, N, Y, 1, (size_t)K, D, (size_t)N, 1, NULL, G, (size_t)N, tf32);
  /* The CUDA path obtains the bias gradient from the same GEMM through an all-ones input column, so in tf32 mode the
   * addends are the conditioned D values; in literal mode they are D itself. */
  if (!dbl) {
    for (int j = 0; j < N; ++j) gb[j] = 0.0f * 0.0f + 1.0f * tf32_cond(D[j], tf32);
    for (int f = 1; f < B; ++f)
      for (int j = 0; j < N; ++j) gb[j] += 1.0f * tf32_cond(D[(size_t)f * N + j], tf32);
  } else {
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      for (int f = 0; f < B; ++f) s += (double)tf32_cond(D[(size_t)f * N + j], tf32);
      gb[j] = (float)s;
    }
  }
}

/* kernUpdatedelta (DevFunc.cu:313-318) then kernAccSum with beta 1 (DevFunc.cu:270-277):
 *   delta = momentum*delta - (1-momentum)*lr*(gradient/n + weightcost*w);  w = delta + 1.0*w
 * `n` is an int promoted to float.  Compiled with -ffp-contract=off: one rounding per operation. */
void orc_sgd(size_t size, float* delta, float* w, const float* grad, int n, float momentum, float lr, float wc) {
  const float c1 = (1 - momentum) * lr;
  const float nf = (float)n;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < size; ++i) {
    float t = grad[i] / nf;
    if (wc != 0.0f) t = t + wc * w[i];
    const float nd = momentum * delta[i] - c1 * t;
    delta[i] = nd;
    w[i] = nd + 1.0f * w[i];
  }
}

/* ------------------------------------------------------------------------------------------------ one train bunch
 * BP_GPU::train_bunch_single (BP_GPU.cu:484-673).  weights[l]/bias[l]/dweights[l]/dbias[l] for l = 1..L.
 * `in`, `targ`: B rows.  frame0 = global index of row 0 (mask key); B_scale = rows used in 2/B and /B (the global
 * bunch, identical to B on one GPU).  If out_opt != NULL it receives the linear outputs (B x Nout). */
void orc_train_bunch(const orc_cfg* c, float** weights, float** bias, float** dweights, float** dbias, int B,
                     const float* in, const float* targ, uint32_t step, int frame0, float* out_opt) {
  const int L = c->numlayers - 1;
  const int* ls = c->layersizes;
  float* y[ORC_MAXLAYER];   /* y[l] = output of layer l (post-dropout), y[0] = (masked copy of) input */
  float* dedx[ORC_MAXLAYER];
  memset(y, 0, sizeof y);
  memset(dedx, 0, sizeof dedx);

  /* Forward (BP_GPU.cu:518-585) */
  y[0] = (float*)malloc((size_t)B * ls[0] * sizeof(float));
  memcpy(y[0], in, (size_t)B * ls[0] * sizeof(float));
  for (int l = 1; l <= L; ++l) {
    const int K = ls[l - 1], N = ls[l];
    if (c->dropoutflag == 1) { /* :534-551 dropout on this layer's INPUT, in place, no rescale */
      const float p = (l == 1) ? c->visible_omit : c->hid_omit;
      float* yp = y[l - 1];
#pragma omp parallel for schedule(static)
      for (int f = 0; f < B; ++f)
        for (int u = 0; u < K; ++u)
          if (orc_dropout_mask(c->seed, step, (uint32_t)(l - 1), (uint32_t)(frame0 + f), (uint32_t)u, p))
            yp[(size_t)f * K + u] = 0.0f;
    }
    float* x = (float*)malloc((size_t)B * N * sizeof(float));
    gemm_fwd(B, K, N, y[l - 1], weights[l], bias[l], x, c->tf32, c->accum_double);
    if (l != L) { /* :560-562 */
#pragma omp parallel for schedule(static)
      for (size_t i = 0; i < (size_t)B * N; ++i) x[i] = act_f(x[i], c->activation);
    } /* else linear output (:564-571) */
    y[l] = x;
  }
  if (out_opt) memcpy(out_opt, y[L], (size_t)B * ls[L] * sizeof(float));

  /* Backward (BP_GPU.cu:588-671).  Updates are applied per layer exactly as the reference does; dedx of the layer
   * below is formed from the PRE-update weights (:636 precedes :643-652). */
  float* dedy = NULL;
  for (int l = L; l >= 1; --l) {
    const int K = ls[l - 1], N = ls[l];
    float* d = (float*)malloc((size_t)B * N * sizeof(float));
    if (l == L) { /* kernSubClean DevFunc.cu:253-268: (2.0f/rows)*(out - clean) */
      const float s = 2.0f / c->bunchsize;
#pragma omp parallel for schedule(static)
      for (size_t i = 0; i < (size_t)B * N; ++i) d[i] = s * (y[L][i] - targ[i]);
    } else { /* kernDsigmoid + kernVecMul :614-615 (on the post-dropout y) */
#pragma omp parallel for schedule(static)
      for (size_t i = 0; i < (size_t)B * N; ++i) d[i] = dact_f(y[l][i], c->activation) * dedy[i];
      free(dedy);
      dedy = NULL;
    }
    dedx[l] = d;
    if (l != 1) { /* :634-637 */
      dedy = (float*)malloc((size_t)B * K * sizeof(float));
      gemm_dx(B, K, N, d, weights[l], dedy, c->tf32, c->accum_double);
    }
    float* G = (float*)malloc((size_t)K * N * sizeof(float));
    float* gb = (float*)malloc((size_t)N * sizeof(float));
    gemm_dw(B, K, N, y[l - 1], d, G, gb, c->tf32, c->accum_double);
    orc_sgd((size_t)K * N, dweights[l], weights[l], G, c->bunchsize, c->momentum, c->lrate, c->weightcost); /* :643,651 */
    orc_sgd((size_t)N, dbias[l], bias[l], gb, c->bunchsize, c->momentum, c->lrate, 0.0f);                   /* :648,652 */
    free(G);
    free(gb);
  }
  for (int l = 0; l <= L; ++l) {
    free(y[l]);
    free(dedx[l]);
  }
}

/* BP_GPU::train (BP_GPU.cu:241-331): full bunches only; the trailing partial bunch is skipped (:297-318).
 * *step_io counts bunches (mask key) and is advanced.  Returns the number of bunches trained. */
int orc_train(const orc_cfg* c, float** weights, float** bias, float** dweights, float** dbias, int n_frames,
              const float* in, const float* targ, uint32_t* step_io) {
  const int B = c->bunchsize, K0 = c->layersizes[0], NO = c->layersizes[c->numlayers - 1];
  int nb = 0;
  for (int i = 0; i + B <= n_frames; i += B, ++nb) {
    orc_train_bunch(c, weights, bias, dweights, dbias, B, in + (size_t)i * K0, targ + (size_t)i * NO, *step_io, 0,
                    NULL);
    (*step_io)++;
  }
  return nb;
}

/* cv_bunch_single (BP_GPU.cu:676-773) over all frames, partial last bunch included (BP_GPU.cu:450).
 * With dropoutflag the weights are multiplied by keep in fp32 before the product (:726-732). out = n x Nout. */
void orc_forward(const orc_cfg* c, float** weights, float** bias, int n_frames, const float* in, float* out) {
  const int L = c->numlayers - 1;
  const int* ls = c->layersizes;
  const int B = c->bunchsize;
  for (int i = 0; i < n_frames; i += B) {
    const int n = (B > n_frames - i) ? (n_frames - i) : B;
    const float* prev = in + (size_t)i * ls[0];
    float* owned = NULL;
    for (int l = 1; l <= L; ++l) {
      const int K = ls[l - 1], N = ls[l];
      const float* W = weights[l];
      float* Ws = NULL;
      if (c->dropoutflag == 1) {
        const float keep = 1.0f - ((l == 1) ? c->visible_omit : c->hid_omit);
        Ws = (float*)malloc((size_t)K * N * sizeof(float));
#pragma omp parallel for schedule(static)
        for (size_t j = 0; j < (size_t)K * N; ++j) Ws[j] = W[j] * keep; /* kernWeightMultiP DevFunc.cu:20-33 */
        W = Ws;
      }
      float* x = (l == L) ? out + (size_t)i * N : (float*)malloc((size_t)n * N * sizeof(float));
      gemm_fwd(n, K, N, prev, W, bias[l], x, c->tf32, c->accum_double);
      if (l != L) {
#pragma omp parallel for schedule(static)
        for (size_t j = 0; j < (size_t)n * N; ++j) x[j] = act_f(x[j], c->activation);
      }
      free(Ws);
      free(owned);
      owned = (l == L) ? NULL : x;
      prev = x;
    }
  }
}

/* BP_GPU::CrossValid (BP_GPU.cu:408-479): sum of squared error accumulated in a host float, frame-major (:458-467). */
float orc_crossvalid(const orc_cfg* c, float** weights, float** bias, int n_frames, const float* in,
                     const float* targ) {
  const int NO = c->layersizes[c->numlayers - 1];
  float* out = (float*)malloc((size_t)n_frames * NO * sizeof(float));
  orc_forward(c, weights, bias, n_frames, in, out);
  float sq = 0.0f;
  for (int j = 0; j < n_frames; ++j)
    for (int d = 0; d < NO; ++d) {
      const float o = out[(size_t)j * NO + d], t = targ[(size_t)j * NO + d];
      sq = sq + (o - t) * (o - t);
    }
  free(out);
  return sq;
}

int orc_version(void) { return 100; }

/* Thread count of the OpenMP loops above.  Launchers such as torchrun export OMP_NUM_THREADS=1 to their children;
 * the CPU baseline must use the host cores it reports, so bench.py sets the count explicitly (n <= 0: leave as is). */
#include <omp.h>
int orc_set_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
}
