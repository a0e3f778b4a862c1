// reader_dump — dumps every chunk the data/IO surface produces (shuffled train chunks, then CV chunks) to a binary
// file.  Uses only the public members that the reference's `Interface` (reference Interface.h:48-101) and our
// re-written one (dnn-for-speech-enhancement_b200/host/Interface.h) have in common, so the SAME source is compiled
// twice: against the reference (oracle/build_ref.sh -> oracle/_ref/ref_reader_dump, run in the build container to
// make tests/golden/reader_*.bin) and against ours (host/Makefile -> bin/reader_dump, run by tests/test_reader.py).
//   reader_dump <out.bin> key=value ...        (same argv keys as BPtrain)
// out.bin: int32 n_train_chunks; per chunk: int32 chunk_id, int32 samples, indata[samples*L0], targ[samples*NO];
//          int32 n_cv_chunks; per chunk the same.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "Interface.h"

static void dump(FILE* f, int id, int n, const WorkPara* p, int numlayers) {
  fwrite(&id, 4, 1, f);
  fwrite(&n, 4, 1, f);
  fwrite(p->indata, 4, (size_t)n * p->layersizes[0], f);
  fwrite(p->targ, 4, (size_t)n * p->layersizes[numlayers - 1], f);
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "wb");
  if (!f) return 2;
  Interface* io = new Interface;
  io->Initial(argc - 1, argv + 1);
  io->get_pfile_info();
  io->get_chunk_info(io->para->train_sent_range);
  int n = io->total_chunks;
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  io->GetRandIndex(order.data(), n);
  fwrite(&n, 4, 1, f);
  for (int i = 0; i < n; ++i) {
    int s = io->Readchunk(order[i]);
    dump(f, order[i], s, io->para, io->numlayers);
  }
  io->get_chunk_info_cv(io->para->cv_sent_range);
  n = io->cv_total_chunks;
  fwrite(&n, 4, 1, f);
  for (int i = 0; i < n; ++i) {
    int s = io->Readchunk_cv(i);
    dump(f, i, s, io->para, io->numlayers);
  }
  fclose(f);
  printf("reader_dump: %u train chunks (%u samples), %u cv chunks (%u samples)\n", io->total_chunks,
         io->total_samples, io->cv_total_chunks, io->cv_total_samples);
  return 0;
}
