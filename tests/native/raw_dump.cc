// raw_dump — dumps what the device-side reader is fed: for every chunk (shuffled train chunks, then CV chunks, same
// order and same lrand48 stream as reader_dump) the raw Pfile records and the sample table that
// Interface::ReadchunkRaw / Readchunk_cvRaw plan.  tests/test_raw_reader.py replays the tables (numpy on CPU, the
// splice kernel on the GPU) and compares the rows bit for bit with the reference reader's golden chunks.
//   raw_dump <out.bin> key=value ...        (same argv keys as BPtrain; prefetch=1, the default, reads the training
//                                            chunks through the ChunkPrefetcher thread exactly as BPtrain does)
// out.bin: int32 fea_dim, ctx, targ_offset, nat, out_dim; float mean[fea_dim], inv_std[fea_dim];
//          int32 n_train_chunks; per chunk: int32 id, n_records, n_samples; fea words; targ words;
//          int32 sample_frame[n], sample_seg[n], sample_row[n];   int32 n_cv_chunks; the same.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "Interface.h"
#include "ChunkPrefetch.h"

static void dump(FILE* f, int id, const RawChunk& rc, int dim, int out) {
  fwrite(&id, 4, 1, f);
  fwrite(&rc.n_records, 4, 1, f);
  fwrite(&rc.n_samples, 4, 1, f);
  if (rc.n_samples == 0) return;
  fwrite(rc.fea_records, 4, (size_t)rc.n_records * (dim + 2), f);
  fwrite(rc.targ_records, 4, (size_t)rc.n_records * (out + 2), f);
  fwrite(rc.sample_frame.data(), 4, rc.n_samples, f);
  fwrite(rc.sample_seg.data(), 4, rc.n_samples, f);
  fwrite(rc.sample_row.data(), 4, rc.n_samples, f);
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "wb");
  if (!f) return 2;
  Interface* io = new Interface;
  io->Initial(argc - 1, argv + 1);
  io->get_pfile_info();
  const WorkPara* p = io->para;
  const int out = p->layersizes[io->numlayers - 1];
  const int hdr[5] = {p->fea_dim, p->fea_context, p->targ_offset, io->nat_block() ? 1 : 0, out};
  fwrite(hdr, 4, 5, f);
  fwrite(io->norm_mean(), 4, p->fea_dim, f);
  fwrite(io->norm_inv_std(), 4, p->fea_dim, f);
  RawChunk rc;
  io->get_chunk_info(io->para->train_sent_range);
  int n = io->total_chunks;
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  io->GetRandIndex(order.data(), n);
  fwrite(&n, 4, 1, f);
  if (p->prefetch && n > 1) {
    RawChunk slots[2];
    {
      ChunkPrefetcher ahead(order, [&](int chunk, int slot) { return io->ReadchunkRaw(chunk, &slots[slot]); });
      int slot = 0;
      for (int i = 0; i < n; ++i) {
        if (ahead.next(&slot) < 0) return 3;
        dump(f, order[i], slots[slot], p->fea_dim, out);
      }
      if (ahead.next(&slot) != -1) return 3;
    }
    io->free_raw(&slots[0]);
    io->free_raw(&slots[1]);
  } else {
    for (int i = 0; i < n; ++i) {
      io->ReadchunkRaw(order[i], &rc);
      dump(f, order[i], rc, p->fea_dim, out);
    }
  }
  io->get_chunk_info_cv(io->para->cv_sent_range);
  n = io->cv_total_chunks;
  fwrite(&n, 4, 1, f);
  for (int i = 0; i < n; ++i) {
    io->Readchunk_cvRaw(i, &rc);
    dump(f, i, rc, p->fea_dim, out);
  }
  io->free_raw(&rc);
  fclose(f);
  return 0;
}
