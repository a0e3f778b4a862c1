// weights_dump — the weight side of the data/IO surface, for a differential test against the reference: after
// Interface::Initial (random initialisation from srand48(init_randem_seed) — Interface.cc:338-352 — or a load from
// initwts_file) dump every layer's weights and biases as raw float32, then call Interface::Writeweights so that the
// .wts file named by outwts_file can be compared too.  Uses only members the reference's `Interface` and ours have in
// common, so the SAME source is compiled against both (oracle/build_ref.sh -> oracle/_ref/ref_weights_dump;
// host/Makefile -> bin/weights_dump).
//   weights_dump <out.bin> key=value ...        (same argv keys as BPtrain)
// out.bin: per weight layer i = 1..numlayers-1: weights[i] (layersizes[i-1]*layersizes[i] floats), bias[i].
#include <cstdio>
#include <cstdlib>

#include "Interface.h"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "wb");
  if (!f) return 2;
  Interface* io = new Interface;
  io->Initial(argc - 1, argv + 1);
  const WorkPara* p = io->para;
  for (int i = 1; i < io->numlayers; ++i) {
    fwrite(p->weights[i], 4, (size_t)p->layersizes[i - 1] * p->layersizes[i], f);
    fwrite(p->bias[i], 4, (size_t)p->layersizes[i], f);
  }
  fclose(f);
  io->Writeweights();
  delete io;   // closes (flushes) the .wts file
  return 0;
}
