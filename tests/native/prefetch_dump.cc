// prefetch_dump — reader_dump's output (same file format, see reader_dump.cc) produced the way BPtrain's training loop
// reads with prefetch=1: the host reader's chunks come through the ChunkPrefetcher thread, alternating between the two
// chunk-buffer pairs.  tests/test_reader.py requires the file to be byte-identical with reader_dump's (the serial loop
// over para->indata / para->targ, which in turn is bit-exact against the reference reader's golden chunks).
//   prefetch_dump <out.bin> key=value ...        (same argv keys as BPtrain)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ChunkPrefetch.h"
#include "Interface.h"

static void dump(FILE* f, int id, int n, const float* in, const float* targ, const WorkPara* p, int numlayers) {
  fwrite(&id, 4, 1, f);
  fwrite(&n, 4, 1, f);
  fwrite(in, 4, (size_t)n * p->layersizes[0], f);
  fwrite(targ, 4, (size_t)n * p->layersizes[numlayers - 1], f);
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
  {
    io->ensure_alt_buffers();
    ChunkPrefetcher ahead(order, [&](int chunk, int slot) { return io->Readchunk(chunk, slot); });
    int slot = 0;
    for (int i = 0; i < n; ++i) {
      const int s = ahead.next(&slot);
      if (s < 0) return 3;
      dump(f, order[i], s, io->chunk_in(slot), io->chunk_targ(slot), io->para, io->numlayers);
    }
    if (ahead.next(&slot) != -1) return 3;
  }
  io->get_chunk_info_cv(io->para->cv_sent_range);
  n = io->cv_total_chunks;
  fwrite(&n, 4, 1, f);
  for (int i = 0; i < n; ++i) {
    const int s = io->Readchunk_cv(i);
    dump(f, i, s, io->para->indata, io->para->targ, io->para, io->numlayers);
  }
  fclose(f);
  delete io;
  return 0;
}
