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
static float* cond_copy(const float* src, size_t n, int mode) {
  float* d = (float*)malloc(n * sizeof(float));
  if (mode == 0) memcpy(d, src, n * sizeof(float));
  else
    for (size_t i = 0; i < n; ++i) d[i] = tf32_cond(src[i], mode);
  return d;
}
/* kernSigmoid / kernDsigmoid bodies (ReLU at HEAD DevFunc.cu:74-77, 92-95; sigmoid variant :52, :62). */
static float act_f(float x, int act) { return act == 0 ? (x > 0 ? x : 0.0f) : 1.0f / (1.0f + expf(-x)); }
static float dact_f(float y, int act) { return act == 0 ? (y > 0 ? 1.0f : 0.0f) : (1.0f - y) * y; }

/* X[B x N] = bias (kernMultiCopy, DevFunc.cu:166-182), then X += Y[B x K] * W[K x N] (SgemmNN, BP_GPU.cu:557-558).
 * W is the reference's w[in*N + out].  Sum over k in ascending order, one fma per term. */
static void gemm_fwd(int B, int K, int N, const float* Y, const float* W, const float* bias, float* X, int tf32,
                     int dbl) {
  float* Yc = cond_copy(Y, (size_t)B * K, tf32);
  float* Wc = cond_copy(W, (size_t)K * N, tf32);
#pragma omp parallel for schedule(static)
  for (int f = 0; f < B; ++f) {
    float* x = X + (size_t)f * N;
    const float* y = Yc + (size_t)f * K;
    if (!dbl) {
      for (int j = 0; j < N; ++j) x[j] = bias[j];
      for (int k = 0; k < K; ++k) {
        const float a = y[k];
        const float* w = Wc + (size_t)k * N;
        for (int j = 0; j < N; ++j) x[j] = fmaf(a, w[j], x[j]);
      }
    } else {
      double* acc = (double*)calloc((size_t)N, sizeof(double));
      for (int k = 0; k < K; ++k) {
        const double a = y[k];
        const float* w = Wc + (size_t)k * N;
        for (int j = 0; j < N; ++j) acc[j] += a * (double)w[j];
      }
      for (int j = 0; j < N; ++j) x[j] = (float)(acc[j] + (double)bias[j]);
      free(acc);
    }
  }
  free(Yc);
  free(Wc);
}

/* E[B x K] = D[B x N] * W^T (SgemmTN with beta 0, BP_GPU.cu:636). */
static void gemm_dx(int B, int K, int N, const float* D, const float* W, float* E, int tf32, int dbl) {
  float* Dc = cond_copy(D, (size_t)B * N, tf32);
  float* Wt = (float*)malloc((size_t)K * N * sizeof(float)); /* Wt[n][k] = cond(W[k][n]) */
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) Wt[(size_t)n * K + k] = tf32_cond(W[(size_t)k * N + n], tf32);
#pragma omp parallel for schedule(static)
  for (int f = 0; f < B; ++f) {
    float* e = E + (size_t)f * K;
    const float* d = Dc + (size_t)f * N;
    if (!dbl) {
      for (int k = 0; k < K; ++k) e[k] = 0.0f;
      for (int n = 0; n < N; ++n) {
        const float a = d[n];
        const float* w = Wt + (size_t)n * K;
        for (int k = 0; k < K; ++k) e[k] = fmaf(a, w[k], e[k]);
      }
    } else {
      double* acc = (double*)calloc((size_t)K, sizeof(double));
      for (int n = 0; n < N; ++n) {
        const double a = d[n];
        const float* w = Wt + (size_t)n * K;
        for (int k = 0; k < K; ++k) acc[k] += a * (double)w[k];
      }
      for (int k = 0; k < K; ++k) e[k] = (float)acc[k];
      free(acc);
    }
  }
  free(Dc);
  free(Wt);
}

/* G[K x N] = Y^T[K x B] * D[B x N] (SgemmNT with beta 0, BP_GPU.cu:642); gb[n] = sum_f D[f][n] in frame order
 * (kernAccSumrow with alpha 0, beta 1: BP_GPU.cu:647, DevFunc.cu:234-240). */
static void gemm_dw(int B, int K, int N, const float* Y, const float* D, float* G, float* gb, int tf32, int dbl) {
  float* Yc = cond_copy(Y, (size_t)B * K, tf32);
  float* Dc = cond_copy(D, (size_t)B * N, tf32);
#pragma omp parallel for schedule(static)
  for (int k = 0; k < K; ++k) {
    float* g = G + (size_t)k * N;
    if (!dbl) {
      for (int j = 0; j < N; ++j) g[j] = 0.0f;
      for (int f = 0; f < B; ++f) {
        const float a = Yc[(size_t)f * K + k];
        const float* d = Dc + (size_t)f * N;
        for (int j = 0; j < N; ++j) g[j] = fmaf(a, d[j], g[j]);
      }
    } else {
      double* acc = (double*)calloc((size_t)N, sizeof(double));
      for (int f = 0; f < B; ++f) {
        const double a = Yc[(size_t)f * K + k];
        const float* d = Dc + (size_t)f * N;
        for (int j = 0; j < N; ++j) acc[j] += a * (double)d[j];
      }
      for (int j = 0; j < N; ++j) g[j] = (float)acc[j];
      free(acc);
    }
  }
  /* The CUDA path obtains the bias gradient from the same GEMM through an all-ones input column, so in tf32 mode the
   * addends are the conditioned D values; in literal mode they are D itself. */
  for (int j = 0; j < N; ++j) {
    if (!dbl) {
      float s = 0.0f * 0.0f + 1.0f * Dc[j];
      for (int f = 1; f < B; ++f) s += 1.0f * Dc[(size_t)f * N + j];
      gb[j] = s;
    } else {
      double s = 0.0;
      for (int f = 0; f < B; ++f) s += (double)Dc[(size_t)f * N + j];
      gb[j] = (float)s;
    }
  }
  free(Yc);
  free(Dc);
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
      for (size_t i = 0; i < (size_t)B * N; ++i) d[i] = s * (y[L][i] - targ[i]);
    } else { /* kernDsigmoid + kernVecMul :614-615 (on the post-dropout y) */
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
        for (size_t j = 0; j < (size_t)K * N; ++j) Ws[j] = W[j] * keep; /* kernWeightMultiP DevFunc.cu:20-33 */
        W = Ws;
      }
      float* x = (l == L) ? out + (size_t)i * N : (float*)malloc((size_t)n * N * sizeof(float));
      gemm_fwd(n, K, N, prev, W, bias[l], x, c->tf32, c->accum_double);
      if (l != L)
        for (size_t j = 0; j < (size_t)n * N; ++j) x[j] = act_f(x[j], c->activation);
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
