// ref_harness — drives the REFERENCE's own trainer class (BP_GPU, reference BP_GPU.h:40-88, compiled unmodified from
// /root/reference by oracle/build_ref.sh) on binary blobs, so the reference's CUDA path (cuBLAS FP32 + its element-wise
// kernels) can be compared with ours on identical inputs without the Pfile plumbing.  TEST INFRASTRUCTURE ONLY.
//
//   ref_harness in.blob out.blob
// in.blob  : int32 numlayers, sizes[numlayers], bunch, n_train, n_cv, dropoutflag, reps
//            float32 lrate, momentum, weightcost, visible_omit, hid_omit
//            per layer l=1..L: W[K*N], b[N]; train_in[n_train*K0], train_targ[n_train*NO]; cv_in, cv_targ
// out.blob : per layer W, b (after training); float32 cv_sum_sq_err; float32 train_ms (last rep of train(), wall,
//            device-synchronised, includes the reference's own H2D of the chunk)
// The calls mirror reference main(): constructor BPtrain.cc:31-32, train :53, returnWeights :57, CrossValid :77.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "BP_GPU.h"

static void rd(FILE* f, void* p, size_t n) {
  if (fread(p, 1, n, f) != n) {
    fprintf(stderr, "ref_harness: short read\n");
    exit(2);
  }
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int numlayers;
  rd(f, &numlayers, 4);
  int sizes[MAXLAYER] = {0};
  rd(f, sizes, 4 * numlayers);
  int bunch, n_train, n_cv, dropoutflag, reps;
  rd(f, &bunch, 4); rd(f, &n_train, 4); rd(f, &n_cv, 4); rd(f, &dropoutflag, 4); rd(f, &reps, 4);
  float lrate, momentum, weightcost, vis, hid;
  rd(f, &lrate, 4); rd(f, &momentum, 4); rd(f, &weightcost, 4); rd(f, &vis, 4); rd(f, &hid, 4);
  float* weights[MAXLAYER] = {0};
  float* bias[MAXLAYER] = {0};
  for (int l = 1; l < numlayers; ++l) {
    weights[l] = new float[(size_t)sizes[l - 1] * sizes[l]];
    bias[l] = new float[sizes[l]];
    rd(f, weights[l], 4 * (size_t)sizes[l - 1] * sizes[l]);
    rd(f, bias[l], 4 * (size_t)sizes[l]);
  }
  const int K0 = sizes[0], NO = sizes[numlayers - 1];
  std::vector<float> tin((size_t)n_train * K0), ttg((size_t)n_train * NO), cin((size_t)n_cv * K0),
      ctg((size_t)n_cv * NO);
  rd(f, tin.data(), 4 * tin.size()); rd(f, ttg.data(), 4 * ttg.size());
  rd(f, cin.data(), 4 * cin.size()); rd(f, ctg.data(), 4 * ctg.size());
  fclose(f);

  BP_GPU* t = new BP_GPU(1, numlayers, sizes, bunch, lrate, momentum, weightcost, weights, bias, dropoutflag, vis, hid);
  float train_ms = 0.f;
  for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    if (n_train > 0) t->train(n_train, tin.data(), ttg.data());
    cudaDeviceSynchronize();
    train_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  t->returnWeights(weights, bias);
  float cv = 0.f;
  if (n_cv > 0) cv = t->CrossValid(n_cv, cin.data(), ctg.data());
  FILE* o = fopen(argv[2], "wb");
  if (!o) return 2;
  for (int l = 1; l < numlayers; ++l) {
    fwrite(weights[l], 4, (size_t)sizes[l - 1] * sizes[l], o);
    fwrite(bias[l], 4, sizes[l], o);
  }
  fwrite(&cv, 4, 1, o);
  fwrite(&train_ms, 4, 1, o);
  fclose(o);
  printf("ref_harness: trained %d frames x %d reps, last train() %.3f ms, cv %.6f\n", n_train, reps, train_ms, cv);
  return 0;
}
