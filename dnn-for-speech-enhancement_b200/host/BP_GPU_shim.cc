// BP_GPU_shim.cc — the binding a maintainer of the reference would add: it implements the
// reference's class BP_GPU (declared in the reference's own, unmodified BP_GPU.h:40-88) on top of libbpgpu's C-ABI
// (include/bp_gpu.h).  Compiled against the reference's header and linked with the reference's unmodified BPtrain.cc
// and Interface.cc in place of BP_GPU.cu / DevFunc.cu, it yields the reference's binary with the B200 trainer behind it
// (the test infrastructure's reference build recipe produces exactly that binary; tests/test_shim.py).
//
// Nothing of the reference is copied here: this file only defines the member functions the reference declares.
#include <cstdio>
#include <cstdlib>

#include "BP_GPU.h"  // the reference header (-I<reference dir>); its CUDA / cuBLAS / cuRAND includes are headers only
#include "bp_gpu.h"  // -I<this repo>/include

// The reference keeps its device state behind the private pointer `dev` (BP_GPU.h:83); the shim parks the handle there.
#define BP_HANDLE reinterpret_cast<bp_handle*>(dev)

static void die(const char* what) {  // reference convention: print, exit(0) (BP_GPU.cu:20-24, 929-933)
  printf("%s: %s\n", what, bp_last_error());
  exit(0);
}

BP_GPU::BP_GPU(int a_GPU_selected, int a_numlayers, int* a_layersizes, int a_bunchsize, float a_lrate,
               float a_momentum, float a_weightcost, float** weights, float** bias, int a_dropoutflag,
               float a_visible_omit, float a_hid_omit) {  // BP_GPU.cu:10-197
  bp_handle* h = nullptr;
  if (bp_create(&h, a_GPU_selected, a_numlayers, a_layersizes, a_bunchsize, a_lrate, a_momentum, a_weightcost, weights,
                bias, a_dropoutflag, a_visible_omit, a_hid_omit) != BP_OK)
    die("BP_GPU");
  dev = reinterpret_cast<BP_WorkSpace*>(h);
  handles = nullptr;
  streams = nullptr;
  gen = nullptr;
  GPU_total = GPU_selected = a_GPU_selected;
  numlayers = a_numlayers;
  bunchsize = a_bunchsize;
  lrate = a_lrate;
  momentum = a_momentum;
  weightcost = a_weightcost;
  dropoutflag = a_dropoutflag;
  visible_omit = a_visible_omit;
  hid_omit = a_hid_omit;
  for (int i = 0; i < a_numlayers; ++i) layersizes[i] = a_layersizes[i];
}

BP_GPU::~BP_GPU() { bp_destroy(BP_HANDLE); }  // BP_GPU.cu:199-238

void BP_GPU::train(int n_frames, float* in, const float* targ) {  // BP_GPU.cu:241-331
  if (bp_train(BP_HANDLE, n_frames, in, targ) != BP_OK) die("train");
}

float BP_GPU::CrossValid(int n_frames, const float* in, const float* targ) {  // BP_GPU.cu:408-479
  float s = 0.0f;
  if (bp_crossvalid(BP_HANDLE, n_frames, in, targ, &s) != BP_OK) die("CrossValid");
  return s;
}

void BP_GPU::returnWeights(float** weights, float** bias) {  // BP_GPU.cu:910-923
  if (bp_return_weights(BP_HANDLE, weights, bias) != BP_OK) die("returnWeights");
}
