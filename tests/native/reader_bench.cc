// reader_bench — host-side chunk reader throughput without a GPU: times Interface::Readchunk (reader=host: byte-swap,
// normalise, splice, NAT, shuffle scatter on the host cores) or Interface::ReadchunkRaw (reader=gpu: positional reads
// of the record blocks + sample table) chunk by chunk over three passes of the training range.
//   reader_bench key=value ...        (same argv keys as BPtrain; reader=host|gpu)
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <vector>
#include "Interface.h"
static double now_s() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
int main(int argc, char** argv) {
  Interface* io = new Interface;
  io->Initial(argc, argv);
  io->get_pfile_info();
  RawChunk rc;
  for (int rep = 0; rep < 3; ++rep) {
    io->get_chunk_info(io->para->train_sent_range);
    int n = io->total_chunks;
    for (int i = 0; i < n; ++i) {
      double t0 = now_s();
      int s = io->para->reader_gpu ? io->ReadchunkRaw(i, &rc) : io->Readchunk(i);
      double dt = now_s() - t0;
      printf("rep %d chunk %d: %d samples %.1f ms -> %.2f M samples/s\n", rep, i, s, dt * 1e3, s / dt / 1e6);
    }
  }
  return 0;
}
