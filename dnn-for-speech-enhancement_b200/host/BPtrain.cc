// BPtrain — one training epoch + cross-validation per invocation, same command line, files and log lines as the
// reference binary (reference BPtrain.cc:16-101), driving the B200 trainer through the C-ABI of include/bp_gpu.h.
//
//   BPtrain fea_file=... norm_file=... targ_file=... outwts_file=... log_file=... initwts_file=...
//           train_sent_range=a-b cv_sent_range=c-d fea_dim=.. fea_context=.. targ_offset=.. traincache=.. bunchsize=..
//           layersizes=a,b,.. lrate=.. momentum=.. weightcost=.. dropoutflag=.. visible_omit=.. hid_omit=..
//           gpu_used=N init_randem_seed=..   [nat=0|1 activation=relu|sigmoid seed=.. decode_file=.. reader=host|gpu]
//           [prefetch=0|1  chunk i+1 is read by a second thread while chunk i is uploaded and queued (default 1)]
//           [epochs=N epoch_first=1 momentum_step=0.04 momentum_max=0.9 seed_step=345  with %d in outwts_file/log_file]
//
// epochs=N (> 1) runs the Perl driver's loop (finetune_DNN_speech_enhancement_dropout_NAT.pl:131-249) inside this
// process: weights stay on the device between epochs, every epoch still writes its own .wts and log, and the files are
// byte-identical to N chained single-epoch invocations (tests/test_cli_gpu.py).
//
// Success exit status is 1 and every error path is "log + exit(0)", as in the reference (BPtrain.cc:100; App. D Q10).
// The Perl epoch driver (finetune_DNN_speech_enhancement_dropout_NAT.pl) works unmodified against this binary.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <vector>

#include "../../include/bp_gpu.h"
#include "DecodeWriter.h"
#include "Interface.h"
#include "ChunkPrefetch.h"

static double now_s() {  // the reference uses time(NULL) (1 s steps); the log lines keep their format
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void die(Interface* io, const char* what) {
  fprintf(io->fp_log, "%s: %s\n", what, bp_last_error());
  fflush(io->fp_log);
  printf("%s: %s\n", what, bp_last_error());
  exit(0);
}

// RawChunk (host planner) -> bp_raw_chunk (C-ABI)
static bp_raw_chunk as_abi(const Interface* io, const RawChunk& rc) {
  bp_raw_chunk c{};
  c.fea_dim = io->para->fea_dim;
  c.fea_context = io->para->fea_context;
  c.targ_offset = io->para->targ_offset;
  c.nat = io->nat_block() ? 1 : 0;
  c.n_records = rc.n_records;
  c.n_samples = rc.n_samples;
  c.fea_records = rc.fea_records;
  c.targ_records = rc.targ_records;
  c.mean = io->norm_mean();
  c.inv_std = io->norm_inv_std();
  c.sample_frame = rc.sample_frame.data();
  c.sample_seg = rc.sample_seg.data();
  c.sample_row = rc.sample_row.data();
  return c;
}

// One epoch: training pass over the shuffled chunks, weight dump, CV pass (BPtrain.cc:36-92).
static void run_epoch(Interface* io, bp_handle* trainer, double t_epoch0) {
  WorkPara* para = io->para;
  // ---- train (BPtrain.cc:36-54)
  io->get_chunk_info(para->train_sent_range);
  std::vector<int> chunk_index(io->total_chunks);
  for (unsigned int i = 0; i < io->total_chunks; ++i) chunk_index[i] = i;
  io->GetRandIndex(chunk_index.data(), io->total_chunks);
  unsigned long long trained_samples = 0;
  RawChunk raw[2];  // reader=gpu: records + sample table of the current chunk (and of the one being prefetched)
  const double t_train0 = now_s();
  {
    // prefetch=1: the same Readchunk / ReadchunkRaw calls in the same order, one chunk ahead on a reader thread
    std::unique_ptr<ChunkPrefetcher> ahead;
    if (para->prefetch && io->total_chunks > 1) {
      if (!para->reader_gpu) io->ensure_alt_buffers();
      ahead.reset(new ChunkPrefetcher(chunk_index, [&](int chunk, int slot) {
        return para->reader_gpu ? io->ReadchunkRaw(chunk, &raw[slot]) : io->Readchunk(chunk, slot);
      }));
    }
    for (unsigned int i = 0; i < io->total_chunks; ++i) {
      int slot = 0;
      const int n = ahead              ? ahead->next(&slot)
                    : para->reader_gpu ? io->ReadchunkRaw(chunk_index[i], &raw[0])
                                       : io->Readchunk(chunk_index[i]);
      fprintf(io->fp_log, "Starting chunk %d of %d containing %d samples.\n", i + 1, io->total_chunks, n);
      fflush(io->fp_log);
      if (n > 0 && para->reader_gpu) {
        const bp_raw_chunk c = as_abi(io, raw[slot]);
        if (bp_train_raw(trainer, &c) != BP_OK) die(io, "train failed");
      } else if (n > 0 && bp_train(trainer, n, io->chunk_in(slot), io->chunk_targ(slot)) != BP_OK) {
        die(io, "train failed");
      }
      trained_samples += n;
    }
  }  // joins the reader thread before the Interface is used for the weight dump and the CV pass
  io->free_raw(&raw[1]);

  printf("begin to write weights\n");
  if (bp_return_weights(trainer, para->weights, para->bias) != BP_OK) die(io, "returnWeights failed");
  const double t_train = now_s() - t_train0;  // returnWeights has drained the device
  io->Writeweights();
  printf("finish to write weights\n\n");

  // ---- CV (BPtrain.cc:61-86)
  printf("begin to CV\n");
  fprintf(io->fp_log, "Starting CV.\n");
  io->get_chunk_info_cv(para->cv_sent_range);
  float squared_err = 0.0f;
  DecodeWriter decw;  // enhanced frames of the CV / decode pass (the reference drops them, BP_GPU.cu:445-473)
  if (para->decode_FN[0]) {
    std::string err;
    if (!decw.open(para->decode_FN, para->decode_format, para->decode_normFN,
                   para->layersizes[io->numlayers - 1], &err)) {
      fprintf(io->fp_log, "%s\n", err.c_str());
      printf("%s\n", err.c_str());
      exit(0);
    }
  }
  const bool fdec = decw.active();
  std::vector<float> dec;
  std::vector<int> rel_sent;
  auto write_decoded = [&](int n) {
    rel_sent.resize(n);
    const int s0 = io->sample_sent.empty() ? 0 : io->cv_first_sentence();
    for (int k = 0; k < n; ++k) rel_sent[k] = io->sample_sent[k] - s0;
    decw.append(dec.data(), n, rel_sent.data(), io->sample_frame_in_sent.data());
  };
  for (unsigned int i = 0; i < io->cv_total_chunks; ++i) {
    const int n = para->reader_gpu ? io->Readchunk_cvRaw(i, &raw[0]) : io->Readchunk_cv(i);
    printf("cur_chunk_samples=%d\n", n);
    if (n <= 0) continue;
    if (para->reader_gpu) {  // one upload serves the score and the decode output
      const bp_raw_chunk c = as_abi(io, raw[0]);
      float sq = 0.0f;
      if (fdec) dec.resize(static_cast<size_t>(n) * para->layersizes[io->numlayers - 1]);
      if (bp_crossvalid_raw(trainer, &c, &sq, fdec ? dec.data() : nullptr) != BP_OK) die(io, "CrossValid failed");
      squared_err += sq;
      if (fdec) write_decoded(n);
      continue;
    }
    float s = 0.0f;
    if (bp_crossvalid(trainer, n, para->indata, para->targ, &s) != BP_OK) die(io, "CrossValid failed");
    squared_err += s;
    if (fdec) {  // decode extension: enhanced (normalised-domain) LPS frames, raw little-endian float32 rows
      dec.resize(static_cast<size_t>(n) * para->layersizes[io->numlayers - 1]);
      if (bp_forward(trainer, n, para->indata, dec.data()) != BP_OK) die(io, "forward failed");
      write_decoded(n);
    }
  }
  decw.close();
  io->free_raw(&raw[0]);
  const float cvacc = squared_err / io->cv_total_samples;
  fprintf(io->fp_log, "CV over. squared error: %f\n", cvacc);
  fflush(io->fp_log);

  const double t_total = now_s() - t_epoch0;
  fprintf(io->fp_log, "Total cost time: %.1f s.\n", t_total);
  if (t_train > 0)  // added line: throughput of the training pass, reader included
    fprintf(io->fp_log, "Training throughput: %.0f frames/sec.\n", trained_samples / t_train);
  fflush(io->fp_log);
}

// Dropout key of the epoch this process (or this pass of the in-process loop) trains: seed= mixed with
// init_randem_seed, which the Perl driver advances by 345 per epoch (.pl:136) — so every epoch draws different masks,
// as the reference's time(NULL)-seeded cuRAND does (BP_GPU.cu:77-78), yet a run is reproducible from its command line.
static unsigned long long epoch_dropout_seed(const WorkPara* para) {
  unsigned long long z = para->seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(unsigned)para->init_randem_seed;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;   // splitmix64 finaliser
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

int main(int argc, char* argv[]) {
  const double t_start = now_s();

  Interface* io = new Interface;
  io->host_alloc = bp_host_alloc;  // page-locked chunk buffers -> asynchronous H2D overlapping the next Readchunk
  io->host_free = bp_host_free;
  io->Initial(argc, argv);
  WorkPara* para = io->para;

  if (para->activation == 1) setenv("BP_ACTIVATION", "sigmoid", 1);
  {
    char buf[64];
    snprintf(buf, sizeof buf, "%llu", epoch_dropout_seed(para));
    setenv("BP_SEED", buf, 1);
  }
  bp_handle* trainer = nullptr;
  if (bp_create(&trainer, para->gpu_used, io->numlayers, para->layersizes, para->bunchsize, para->lrate,
                para->momentum, para->weightcost, para->weights, para->bias, para->dropoutflag, para->visible_omit,
                para->hid_omit) != BP_OK)
    die(io, "GPU trainer creation failed");

  io->get_pfile_info();

  run_epoch(io, trainer, t_start);
  for (int epoch = 1; epoch < para->epochs; ++epoch) {
    // what a fresh process started by the Perl driver would begin with (bp_gpu.h: bp_begin_epoch), minus the weight
    // round-trip through the .wts file
    const double t0 = now_s();
    const float m = io->begin_epoch(epoch);
    if (bp_begin_epoch(trainer, para->lrate, m, para->weightcost, 1) != BP_OK) die(io, "begin_epoch failed");
    if (bp_set_dropout_seed(trainer, epoch_dropout_seed(para)) != BP_OK) die(io, "set_dropout_seed failed");
    run_epoch(io, trainer, t0);
  }

  printf("all finish!\n");
  bp_destroy(trainer);
  delete io;
  return 1;
}
