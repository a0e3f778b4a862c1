// Host data/IO surface of BPtrain (see Interface.h).  Behavioural spec = reference Interface.cc (cited per function);
// the code is new: table-driven argv parsing, one chunk planner and one sample assembler shared by train and CV.
#include "Interface.h"

#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace {

constexpr long kPfileHeaderBytes = 32768;  // PFILE_HEADER_SIZE, reference Interface.cc:13

inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }

inline float be_float(const float* p) {  // big-endian word -> host float
  uint32_t u;
  std::memcpy(&u, p, 4);
  u = bswap32(u);
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline int be_int(const float* p) {
  uint32_t u;
  std::memcpy(&u, p, 4);
  return static_cast<int>(bswap32(u));
}

// Static split of [0, n) over the host threads (at most 32, at least `grain` items each); fn(begin, end).
template <class Fn>
void parallel_for(int n, int grain, Fn fn) {
  const int hw = static_cast<int>(std::thread::hardware_concurrency());
  const int t = std::max(1, std::min({hw > 0 ? hw : 1, 32, n / std::max(1, grain)}));
  if (t == 1) {
    fn(0, n);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(t - 1);
  const int per = (n + t - 1) / t;
  for (int k = 1; k < t; ++k) th.emplace_back([=, &fn] { fn(std::min(n, k * per), std::min(n, (k + 1) * per)); });
  fn(0, std::min(n, per));
  for (std::thread& x : th) x.join();
}

enum ArgKind { kStr, kInt, kFloat };
struct ArgSpec {
  const char* key;
  ArgKind kind;
  size_t offset;
};
#define ARG(key, kind, field) {key, kind, offsetof(WorkPara, field)}
// Keys of reference Interface.cc:100-226 (same spelling) + our extensions.
const ArgSpec kArgs[] = {
    ARG("fea_file", kStr, fea_FN),
    ARG("norm_file", kStr, fea_normFN),
    ARG("targ_file", kStr, targ_FN),
    ARG("outwts_file", kStr, out_weightFN),
    ARG("log_file", kStr, log_FN),
    ARG("initwts_file", kStr, init_weightFN),
    ARG("train_sent_range", kStr, train_sent_range),
    ARG("cv_sent_range", kStr, cv_sent_range),
    ARG("fea_dim", kInt, fea_dim),
    ARG("fea_context", kInt, fea_context),
    ARG("targ_offset", kInt, targ_offset),
    ARG("dropoutflag", kInt, dropoutflag),
    ARG("traincache", kInt, traincache),
    ARG("bunchsize", kInt, bunchsize),
    ARG("gpu_used", kInt, gpu_used),
    ARG("init_randem_seed", kInt, init_randem_seed),
    ARG("momentum", kFloat, momentum),
    ARG("weightcost", kFloat, weightcost),
    ARG("lrate", kFloat, lrate),
    ARG("visible_omit", kFloat, visible_omit),
    ARG("hid_omit", kFloat, hid_omit),
    ARG("init_randem_weight_min", kFloat, init_randem_weight_min),
    ARG("init_randem_weight_max", kFloat, init_randem_weight_max),
    ARG("init_randem_bias_max", kFloat, init_randem_bias_max),
    ARG("init_randem_bias_min", kFloat, init_randem_bias_min),
    ARG("nat", kInt, nat),
    ARG("decode_file", kStr, decode_FN),
    ARG("decode_format", kStr, decode_format),
    ARG("decode_norm_file", kStr, decode_normFN),
    ARG("epochs", kInt, epochs),
    ARG("epoch_first", kInt, epoch_first),
    ARG("momentum_step", kFloat, momentum_step),
    ARG("momentum_max", kFloat, momentum_max),
    ARG("seed_step", kInt, seed_step),
};
#undef ARG

}  // namespace

// "%d" in a file-name pattern -> the epoch number (the Perl driver's mlp.$i.wts / mlp.$i.log, .pl:134-135).
std::string epoch_name(const char* pattern, int number) {
  std::string s(pattern);
  const size_t pos = s.find("%d");
  if (pos != std::string::npos) s.replace(pos, 2, std::to_string(number));
  return s;
}

Interface::Interface() { para = new WorkPara(); }

Interface::~Interface() {
  if (fp_data) fclose(fp_data);
  if (fp_targ) fclose(fp_targ);
  if (fp_out) fclose(fp_out);
  if (fp_log) fclose(fp_log);
  free_floats(para->indata, pinned_in[0]);
  free_floats(para->targ, pinned_targ[0]);
  free_floats(alt_indata, pinned_in[1]);
  free_floats(alt_targ, pinned_targ[1]);
  for (int i = 1; i < numlayers; ++i) {
    delete[] para->weights[i];
    delete[] para->bias[i];
  }
  delete para;
  delete[] chunk_frame_st;
  delete[] cv_chunk_frame_st;
  delete[] framesBeforeSent;
}

// Reference error convention: message into the log, then exit(0) (e.g. Interface.cc:246-265).
void Interface::fatal(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  FILE* out = fp_log ? fp_log : stderr;
  vfprintf(out, fmt, ap);
  va_end(ap);
  fflush(out);
  exit(0);
}

// Parameter echo, byte-compatible with reference Interface.cc:267-298 (logs were machine-read, .pl:113-121).
void Interface::echo_parameters() {
  fprintf(fp_log, "parameters input:\n");
  fprintf(fp_log, "fea_file:             %s\n", para->fea_FN);
  fprintf(fp_log, "norm_file:            %s\n", para->fea_normFN);
  fprintf(fp_log, "targ_file:            %s\n", para->targ_FN);
  fprintf(fp_log, "outwts_file:          %s\n", para->out_weightFN);
  fprintf(fp_log, "log_file:		          %s\n", para->log_FN);
  fprintf(fp_log, "initwts_file:         %s\n", para->init_weightFN);
  fprintf(fp_log, "train_sent_range:     %s\n", para->train_sent_range);
  fprintf(fp_log, "cv_sent_range:        %s\n", para->cv_sent_range);
  fprintf(fp_log, "fea_dim:		          %d\n", para->fea_dim);
  fprintf(fp_log, "fea_context:		      %d\n", para->fea_context);
  fprintf(fp_log, "bunchsize:		        %d\n", para->bunchsize);
  fprintf(fp_log, "gpu_used:		          %d\n", para->gpu_used);
  fprintf(fp_log, "train_cache:		      %d\n", para->traincache);
  fprintf(fp_log, "init_randem_seed:		  %d\n", para->init_randem_seed);
  fprintf(fp_log, "targ_offset:		      %d\n", para->targ_offset);
  fprintf(fp_log, "dropoutflag:		      %d\n", para->dropoutflag);
  fprintf(fp_log, "init_randem_weight_max:		  %f\n", para->init_randem_weight_max);
  fprintf(fp_log, "init_randem_weight_min:		  %f\n", para->init_randem_weight_min);
  fprintf(fp_log, "init_randem_bias_max:		    %f\n", para->init_randem_bias_max);
  fprintf(fp_log, "init_randem_bias_min:		    %f\n", para->init_randem_bias_min);
  fprintf(fp_log, "momentum:		                %f\n", para->momentum);
  fprintf(fp_log, "weightcost:		              %f\n", para->weightcost);
  fprintf(fp_log, "learnrate:		              %f\n", para->lrate);
  fprintf(fp_log, "visible_omit:		      %f\n", para->visible_omit);
  fprintf(fp_log, "hid_omit:		      %f\n", para->hid_omit);
  fprintf(fp_log, "layersizes:		              ");
  for (int j = 0; j < numlayers; ++j) fprintf(fp_log, "%d,", para->layersizes[j]);
  fprintf(fp_log, "\n");
  fprintf(fp_log, "Please check...\n");
}

// Momentum of epoch offset e as the Perl driver would hand it over: accumulated in double, printed with %.15g (Perl's
// number -> string rule), parsed with atof and narrowed to float (.pl:137,219 -> Interface.cc:176).
float epoch_momentum_value(double base, double step, double max, int e) {
  double m = base;
  for (int i = 0; i < e; ++i) m = m + step;
  if (e > 0 && m > max) m = max;
  char buf[64];
  snprintf(buf, sizeof buf, "%.15g", m);
  return static_cast<float>(atof(buf));
}
float Interface::epoch_momentum(int e) const {
  return epoch_momentum_value(base_momentum_d, momentum_step_d, momentum_max_d, e);
}

// Epoch boundary of the in-process loop: what a fresh process started by the Perl driver would see.
float Interface::begin_epoch(int e) {
  const int number = para->epoch_first + e;
  para->momentum = epoch_momentum(e);
  para->init_randem_seed = base_seed + e * para->seed_step;
  if (e > 0) {
    fclose(fp_log);
    fclose(fp_out);
    const std::string prev = epoch_name(para->out_weight_pattern, number - 1);
    snprintf(para->init_weightFN, MAXLINE, "%s", prev.c_str());  // for the echo only: the weights never left the device
    snprintf(para->out_weightFN, MAXLINE, "%s", epoch_name(para->out_weight_pattern, number).c_str());
    snprintf(para->log_FN, MAXLINE, "%s", epoch_name(para->log_pattern, number).c_str());
    if (!(fp_log = fopen(para->log_FN, "wt"))) {
      printf("can not open output log file: %s\n", para->log_FN);
      exit(0);
    }
    if (!(fp_out = fopen(para->out_weightFN, "wb"))) fatal("can not open output weights file: %s\n", para->out_weightFN);
    echo_parameters();
    fprintf(fp_log, "Loading Norm file...\nNorm file loaded.\nLoading Init weight file...\nInit weight file loaded.\n");
    srand48(para->init_randem_seed);  // a new process seeds once, then shuffles (Interface.cc:338)
    fflush(fp_log);
  }
  return para->momentum;
}

// Interface::Initial — reference Interface.cc:69-404: parse key=value args, open files, echo parameters, load norm
// file, allocate + initialise weights (random via drand48 or MAT-v4 file), check the input width, allocate chunk buffers.
void Interface::Initial(int argc, char** argv) {
  for (int i = 1; i < argc; ++i) {
    char* eq = strchr(argv[i], '=');
    if (!eq) fatal("Arg: %s  Format Error\n", argv[i]);
    std::string key(argv[i], eq - argv[i]);
    const char* val = eq + 1;
    if (key == "layersizes") {  // comma list -> numlayers (:227-243)
      int count = 0;
      const char* p = val;
      while (count < MAXLAYER) {
        para->layersizes[count++] = atoi(p);
        const char* c = strchr(p, ',');
        if (!c) break;
        p = c + 1;
      }
      numlayers = count;
      continue;
    }
    if (key == "activation") {
      para->activation = (strcmp(val, "sigmoid") == 0) ? 1 : 0;
      continue;
    }
    if (key == "math") {  // tf32 (default) | 3xtf32 (split precision, ~fp32 accuracy)
      setenv("BP_MATH", val, 1);
      continue;
    }
    if (key == "reader") {  // host (default, the reference's Readchunk) | gpu (device-side splice, §8f-1)
      para->reader_gpu = (strcmp(val, "gpu") == 0) ? 1 : 0;
      continue;
    }
    if (key == "prefetch") {  // reader=gpu: read chunk i+1 on a second thread while chunk i is uploaded (default 1)
      para->prefetch = atoi(val) != 0 ? 1 : 0;
      continue;
    }
    if (key == "seed") {
      para->seed = strtoull(val, nullptr, 0);
      continue;
    }
    // the momentum schedule is evaluated in double, as the Perl driver does (no `continue`: the table stores the floats)
    if (key == "momentum") base_momentum_d = atof(val);
    if (key == "momentum_step") momentum_step_d = atof(val);
    if (key == "momentum_max") momentum_max_d = atof(val);
    for (const ArgSpec& a : kArgs) {
      if (key != a.key) continue;
      char* base = reinterpret_cast<char*>(para) + a.offset;
      if (a.kind == kStr) {
        strncpy(base, val, MAXLINE - 1);
        base[MAXLINE - 1] = '\0';
      } else if (a.kind == kInt) {
        *reinterpret_cast<int*>(base) = atoi(val);
      } else {
        *reinterpret_cast<float*>(base) = static_cast<float>(atof(val));
      }
      break;
    }
    // unknown keys (e.g. numlayers= sent by the Perl driver) are accepted and ignored, like the reference
  }

  // in-process epochs: remember the command-line values, expand "%d" for the first epoch
  snprintf(para->out_weight_pattern, MAXLINE, "%s", para->out_weightFN);
  snprintf(para->log_pattern, MAXLINE, "%s", para->log_FN);
  base_seed = para->init_randem_seed;
  if (para->epochs < 1) para->epochs = 1;
  if (para->epochs > 1 || strstr(para->out_weightFN, "%d") || strstr(para->log_FN, "%d")) {
    snprintf(para->out_weightFN, MAXLINE, "%s", epoch_name(para->out_weight_pattern, para->epoch_first).c_str());
    snprintf(para->log_FN, MAXLINE, "%s", epoch_name(para->log_pattern, para->epoch_first).c_str());
  }
  if (!(fp_log = fopen(para->log_FN, "wt"))) {
    printf("can not open output log file: %s\n", para->log_FN);
    exit(0);
  }
  if (para->epochs > 1 && (!strstr(para->out_weight_pattern, "%d") || !strstr(para->log_pattern, "%d")))
    fatal("epochs > 1 needs %%d (the epoch number) in outwts_file and log_file\n");
  if (!(fp_data = fopen(para->fea_FN, "rb"))) fatal("can not open feature file: %s\n", para->fea_FN);
  if (!(fp_targ = fopen(para->targ_FN, "rb"))) fatal("can not open target file: %s\n", para->targ_FN);
  if (!(fp_out = fopen(para->out_weightFN, "wb"))) fatal("can not open output weights file: %s\n", para->out_weightFN);

  echo_parameters();

  if (numlayers < 2 || para->fea_dim <= 0 || para->fea_context <= 0 || para->traincache <= 0 || para->bunchsize <= 0)
    fatal("layersizes / fea_dim / fea_context / traincache / bunchsize missing or invalid\n");

  // Norm file (Interface.cc:301-326): header line, fea_dim means, header line, fea_dim inverse std-devs.
  FILE* fp_norm = fopen(para->fea_normFN, "rt");
  if (!fp_norm) fatal("can not open normalization file: %s\n", para->fea_normFN);
  fprintf(fp_log, "Loading Norm file...\n");
  mean.assign(para->fea_dim, 0.0f);
  dVar.assign(para->fea_dim, 1.0f);
  char line[MAXLINE];
  for (std::vector<float>* v : {&mean, &dVar}) {
    if (!fgets(line, MAXLINE, fp_norm)) fatal("norm file too short\n");
    for (int j = 0; j < para->fea_dim; ++j) {
      if (!fgets(line, MAXLINE, fp_norm)) fatal("norm file too short\n");
      (*v)[j] = static_cast<float>(atof(line));
    }
  }
  fclose(fp_norm);
  fprintf(fp_log, "Norm file loaded.\n");

  for (int i = 1; i < numlayers; ++i) {
    const size_t n = static_cast<size_t>(para->layersizes[i]) * para->layersizes[i - 1];
    para->weights[i] = new float[n]();
    para->bias[i] = new float[para->layersizes[i]]();
  }
  srand48(para->init_randem_seed);  // once, for weights AND shuffles (Interface.cc:338)

  if (para->init_weightFN[0] == '\0') {
    fprintf(fp_log, "Getting Randemed initial weights...\n");
    for (int i = 1; i < numlayers; ++i) {
      GetRandWeight(para->weights[i], para->init_randem_weight_min, para->init_randem_weight_max,
                    para->layersizes[i] * para->layersizes[i - 1]);
      GetRandWeight(para->bias[i], para->init_randem_bias_min, para->init_randem_bias_max, para->layersizes[i]);
    }
    fprintf(fp_log, "Randemed initial weights getted.\n");
  } else {
    // MAT-v4 weight file (Interface.cc:351-391): per layer {10, n_out, n_in, 0, namelen} name, data; then bias.
    FILE* fw = fopen(para->init_weightFN, "rb");
    if (!fw) fatal("can not open initial weights file: %s\n", para->init_weightFN);
    fprintf(fp_log, "Loading Init weight file...\n");
    for (int i = 1; i < numlayers; ++i) {
      int32_t hdr[5];
      char name[256];
      const int n_out = para->layersizes[i], n_in = para->layersizes[i - 1];
      if (fread(hdr, 4, 5, fw) != 5 || hdr[4] <= 0 || hdr[4] > 255 || fread(name, 1, hdr[4], fw) != (size_t)hdr[4])
        fatal("init weights file truncated\n");
      if (hdr[1] != n_out || hdr[2] != n_in) {
        fprintf(fp_log, "%d,%d,%d,%d\n", hdr[1], hdr[2], n_out, n_in);
        fatal("init weights node nums do not match\n");
      }
      if (fread(para->weights[i], 4, static_cast<size_t>(n_in) * n_out, fw) != static_cast<size_t>(n_in) * n_out)
        fatal("init weights file truncated\n");
      if (fread(hdr, 4, 5, fw) != 5 || hdr[4] <= 0 || hdr[4] > 255 || fread(name, 1, hdr[4], fw) != (size_t)hdr[4])
        fatal("init weights file truncated\n");
      if (hdr[2] != n_out || hdr[1] != 1) fatal("init bias node nums do not match\n");
      if (fread(para->bias[i], 4, n_out, fw) != static_cast<size_t>(n_out)) fatal("init weights file truncated\n");
    }
    fclose(fw);
    fprintf(fp_log, "Init weight file loaded.\n");
  }

  // Input width check.  HEAD: fea_dim*fea_context + fea_dim == layersizes[0] (NAT always on, Interface.cc:395);
  // the commented line :394 is the no-NAT check.  Both are accepted here; `nat=` forces one.
  const int spliced = para->fea_dim * para->fea_context;
  if (para->nat < 0) para->nat = (spliced + para->fea_dim == para->layersizes[0]) ? 1 : 0;
  use_nat = para->nat == 1;
  if (spliced + (use_nat ? para->fea_dim : 0) != para->layersizes[0])
    fatal("feadim times (+ 129 noise) context must be equal to layersizes[0]\n");

  para->indata = alloc_floats(static_cast<size_t>(para->layersizes[0]) * para->traincache, &pinned_in[0]);
  para->targ = alloc_floats(static_cast<size_t>(para->layersizes[numlayers - 1]) * para->traincache, &pinned_targ[0]);
  chunk_frame_st = new int[MAXCHUNK];
  cv_chunk_frame_st = new int[MAXCHUNK];
  fflush(fp_log);
}

// Chunk buffers: page-locked through the trainer's allocator when main() installed one (and it succeeds), else heap.
float* Interface::alloc_floats(size_t n, bool* pinned) {
  *pinned = false;
  if (host_alloc && host_free) {
    if (float* p = static_cast<float*>(host_alloc(n * sizeof(float)))) {
      *pinned = true;
      return p;
    }
  }
  return new float[n];
}

void Interface::free_floats(float* p, bool pinned) {
  if (!p) return;
  if (pinned) host_free(p);
  else delete[] p;
}

// Second chunk-buffer pair for the prefetching training loop (ChunkPrefetch.h); allocated on first use.
void Interface::ensure_alt_buffers() {
  if (alt_indata) return;
  alt_indata = alloc_floats(static_cast<size_t>(para->layersizes[0]) * para->traincache, &pinned_in[1]);
  alt_targ = alloc_floats(static_cast<size_t>(para->layersizes[numlayers - 1]) * para->traincache, &pinned_targ[1]);
}

// Interface::Writeweights — reference Interface.cc:411-465: MAT-v4, little-endian, weights then bias per layer.
// (The reference's debug dump weights.txt is written only when BP_DUMP_WEIGHTS_TXT is set.)
void Interface::Writeweights() {
  fprintf(fp_log, "Saving weights to file...\n");
  FILE* txt = getenv("BP_DUMP_WEIGHTS_TXT") ? fopen("weights.txt", "w") : nullptr;
  for (int i = 1; i < numlayers; ++i) {
    const int n_out = para->layersizes[i], n_in = para->layersizes[i - 1];
    char name[64];
    snprintf(name, sizeof name, "weights%d%d", i, i + 1);
    int32_t hdr[5] = {10, n_out, n_in, 0, static_cast<int32_t>(strlen(name) + 1)};
    fwrite(hdr, 4, 5, fp_out);
    fwrite(name, 1, hdr[4], fp_out);
    fwrite(para->weights[i], 4, static_cast<size_t>(n_in) * n_out, fp_out);
    if (txt)
      for (size_t j = 0; j < static_cast<size_t>(n_in) * n_out; ++j) fprintf(txt, "%f    ", para->weights[i][j]);
    snprintf(name, sizeof name, "bias%d", i + 1);
    int32_t hb[5] = {10, 1, n_out, 0, static_cast<int32_t>(strlen(name) + 1)};
    fwrite(hb, 4, 5, fp_out);
    fwrite(name, 1, hb[4], fp_out);
    fwrite(para->bias[i], 4, n_out, fp_out);
    if (txt)
      for (int j = 0; j < n_out; ++j) fprintf(txt, "%f    ", para->bias[i][j]);
  }
  if (txt) fclose(txt);
  fflush(fp_out);
  fprintf(fp_log, "Saving over.\n");
}

// get_uint / read_tail — reference Interface.cc:1057-1093.
void Interface::get_uint(const char* hdr, const char* argname, unsigned int* val) {
  const char* p = strstr(hdr, argname);
  if (!p) fatal("pfile header format is Not correct.\n");
  int count = 0;
  sscanf(p + strlen(argname), " %u%n", val, &count);
  if (count <= 1) fatal("%s num in pfile header is Not correct.\n", argname);
}

void Interface::read_tail(FILE* fp, long file_offset, unsigned int sentnum, int* out) {
  // sentence index = (num_sentences+1) big-endian int32; the leading 0 is skipped (Interface.cc:1083)
  fseek(fp, file_offset + 4, SEEK_SET);
  std::vector<uint32_t> raw(sentnum);
  if (fread(raw.data(), 4, sentnum, fp) != sentnum) fatal("pfile tail is Not correct.\n");
  for (unsigned int i = 0; i < sentnum; ++i) out[i] = static_cast<int>(bswap32(raw[i]));
}

// Interface::get_pfile_info — reference Interface.cc:468-555.
void Interface::get_pfile_info() {
  std::vector<char> header(kPfileHeaderBytes + 1, 0);
  fprintf(fp_log, "begin to read in_pfile\n");
  fseek(fp_data, 0, SEEK_SET);
  if (fread(header.data(), kPfileHeaderBytes, 1, fp_data) != 1) fatal("Failed to read data pfile header.\n");
  get_uint(header.data(), "-num_sentences", &total_sents);
  get_uint(header.data(), "-num_frames", &total_frames);
  framesBeforeSent = new int[total_sents];
  read_tail(fp_data, kPfileHeaderBytes + static_cast<long>(total_frames) * 4L * (2 + para->fea_dim), total_sents,
            framesBeforeSent);

  fprintf(fp_log, "begin to read target_pfile\n");
  unsigned int tsents = 0, tframes = 0;
  fseek(fp_targ, 0, SEEK_SET);
  if (fread(header.data(), kPfileHeaderBytes, 1, fp_targ) != 1) fatal("Failed to read target pfile header.\n");
  get_uint(header.data(), "-num_sentences", &tsents);
  get_uint(header.data(), "-num_frames", &tframes);
  std::vector<int> ttail(tsents);
  read_tail(fp_targ, kPfileHeaderBytes + static_cast<long>(tframes) * 4L * (2 + para->layersizes[numlayers - 1]),
            tsents, ttail.data());
  fprintf(fp_log, "tmpsentnum=%d,tmpframenum=%d,total_frames=%d\n", tsents, tframes, total_frames);
  if (tsents != total_sents || tframes != total_frames)
    fatal("frames or sentence num in target pfile and data pfile is not consistent.\n");
  fprintf(fp_log, "frames or sentence num in target pfile and data pfile is consistent.\n");
  for (unsigned int i = 0; i < total_sents; ++i)
    if (ttail[i] != framesBeforeSent[i]) fatal("tails in target pfile and data pfile is not consistent---%d.\n", i);
  fprintf(fp_log, "Get pfile info over: Training data has %u frames, %u sentences.\n", total_frames, total_sents);
}

Interface::Range Interface::parse_range(const char* range, const char* what) {
  const char* dash = strchr(range, '-');
  if (!dash) fatal("%ssent range: %s format error.\n", what, range);
  Range r;
  r.st = atoi(std::string(range, dash - range).c_str());
  r.en = atoi(dash + 1);
  if (r.en < r.st || r.st < 0 || r.en >= static_cast<int>(total_sents))
    fatal("%ssent range: %d to %d number error.\n", what, r.st, r.en);
  return r;
}

// Chunk planner — reference Interface.cc:558-620 (train) / 622-686 (cv), one implementation.
// A sentence of T frames yields T-(ctx-1) samples (0 if T < ctx); when the running count reaches traincache a new
// chunk starts mid-sentence and the ctx-1 context frames after the cut are lost again.
void Interface::plan_chunks(const Range& r, int* starts, unsigned int* n_chunks, unsigned int* n_samples) {
  const int ctx = para->fea_context, cache = para->traincache;
  int frame = r.st == 0 ? 0 : framesBeforeSent[r.st - 1];
  unsigned int count = 1;
  int in_chunk = 0;
  starts[0] = frame;
  for (int s = r.st; s <= r.en; ++s) {
    const int len = framesBeforeSent[s] - frame;
    frame = framesBeforeSent[s];
    in_chunk += len - (len >= ctx ? ctx - 1 : len);
    while (in_chunk >= cache) {
      const int next = frame - (in_chunk - cache);
      if (next < static_cast<int>(total_frames)) {
        if (count >= MAXCHUNK) fatal("too many chunks (> %d)\n", MAXCHUNK);
        starts[count++] = next;
        in_chunk = (frame - next > ctx - 1) ? (frame - next - ctx + 1) : 0;
      } else {
        // next == total_frames: the chunk is exactly full at the end of the file.  The reference spins forever in this
        // corner (Interface.cc:607-614 has no else); here the last chunk simply keeps its `cache` samples.
        break;
      }
    }
  }
  *n_chunks = count;
  *n_samples = (count - 1) * cache + in_chunk;
}

void Interface::get_chunk_info(char* range) {
  train_r = parse_range(range, "");
  plan_chunks(train_r, chunk_frame_st, &total_chunks, &total_samples);
  fprintf(fp_log, "Get chunk info over: Training sentences have %d chunks, %d samples.\n", total_chunks, total_samples);
}

void Interface::get_chunk_info_cv(char* range) {
  cv_r = parse_range(range, "cv ");
  plan_chunks(cv_r, cv_chunk_frame_st, &cv_total_chunks, &cv_total_samples);
  fprintf(fp_log, "Get cv chunk info over: CV sentences have %d chunks, %d samples.\n", cv_total_chunks,
          cv_total_samples);
}

// Raw records [first_frame, first_frame+n): (2+dim) big-endian words each; returns the sentence id of record 0.
void Interface::read_records(FILE* fp, int dim, long first_frame, int n_frames, std::vector<float>* rec,
                             int* first_sent) {
  const long rec_bytes = 4L * (dim + 2);
  if (fseek(fp, kPfileHeaderBytes + first_frame * rec_bytes, SEEK_SET) != 0) fatal("pfile cannot fseek to chunk.\n");
  rec->resize(static_cast<size_t>(n_frames) * (dim + 2));
  if (fread(rec->data(), rec_bytes, n_frames, fp) != static_cast<size_t>(n_frames)) fatal("pfile read failed.\n");
  *first_sent = be_int(rec->data());
}

// The sentence segments that intersect the record block [first, first+n_frames); `sent` = sentence of record `first`.
std::vector<Interface::Seg> Interface::segments(int first, int n_frames, int sent) const {
  std::vector<Seg> segs;
  int done = 0, frame = first, s = sent;
  while (done != n_frames) {
    const int len = (framesBeforeSent[s] > first + n_frames) ? (n_frames - done) : (framesBeforeSent[s] - frame);
    segs.push_back({done, len, s});
    frame = framesBeforeSent[s];
    ++s;
    done += len;
  }
  return segs;
}

void Interface::note_sample(int cur, const Seg& sg, int first, int j) {
  const int f = first + sg.begin + j + para->targ_offset;  // global index of the target frame
  sample_sent[cur] = sg.sent;
  sample_frame_in_sent[cur] = f - (sg.sent == 0 ? 0 : framesBeforeSent[sg.sent - 1]);
}

// Sample assembly — reference Readchunk Interface.cc:689-861 / Readchunk_cv :864-1034, one implementation:
// byte-swap + (x-mean)*dVar, 11-frame splice, NAT block (mean of the segment's first six frames, /6.0f, summed left to
// right), target frame j+targ_offset, rows scattered through a Fisher-Yates permutation (train) or in order (CV).
int Interface::assemble(int chunk_index, const int* starts, unsigned int n_chunks, unsigned int n_samples,
                        int sent_end, bool shuffle, float* in_dst, float* targ_dst) {
  const int dim = para->fea_dim, ctx = para->fea_context, in_w = para->layersizes[0];
  const int out_w = para->layersizes[numlayers - 1];
  const int first = starts[chunk_index];
  const bool last = static_cast<unsigned int>(chunk_index) == n_chunks - 1;
  const int n_frames = (last ? framesBeforeSent[sent_end] : starts[chunk_index + 1]) - first;
  const int samples = last ? static_cast<int>(n_samples) - para->traincache * chunk_index : para->traincache;

  if (n_frames <= 0 || samples <= 0) return 0;  // empty trailing chunk (range ends exactly on a chunk boundary)
  std::vector<int> order(samples);
  for (int i = 0; i < samples; ++i) order[i] = i;
  if (shuffle) GetRandIndex(order.data(), samples);

  // Both record blocks are fetched up front: the target block by a second thread while this one reads the features.
  std::vector<float>&rec = scratch_rec, &trec = scratch_trec, &fea = scratch_fea;  // kept across chunks: no page faults
  int sent = 0, tsent = 0;
  std::thread targ_reader([&] { read_records(fp_targ, out_w, first, n_frames, &trec, &tsent); });
  read_records(fp_data, dim, first, n_frames, &rec, &sent);
  // normalise: two fp32 operations, subtract then multiply (Interface.cc:745-746).  Every element is computed exactly
  // as in the serial loop; only the frames are dealt out to the host threads.
  if (fea.size() < static_cast<size_t>(n_frames) * dim) fea.resize(static_cast<size_t>(n_frames) * dim);
  parallel_for(n_frames, 256, [&](int f0, int f1) {
    for (int f = f0; f < f1; ++f) {
      const float* src = rec.data() + static_cast<size_t>(f) * (dim + 2) + 2;
      float* dst = fea.data() + static_cast<size_t>(f) * dim;
      for (int j = 0; j < dim; ++j) {
        float v = be_float(src + j);
        v -= mean[j];
        v *= dVar[j];
        dst[j] = v;
      }
    }
  });

  const std::vector<Seg> segs = segments(first, n_frames, sent);

  // Plan: sample `cur` = window j of segment sample_segi[cur], in file order (the order the reference's loops visit
  // them, Interface.cc:755-787); NAT vector per segment (mean of its first six frames, /6.0f, summed left to right;
  // the reference reads six frames even when the segment is shorter, frames past the block count as 0).
  sample_sent.assign(samples, 0);
  sample_frame_in_sent.assign(samples, 0);
  std::vector<int> sample_segi(samples), sample_j(samples);
  int cur = 0;
  for (size_t si = 0; si < segs.size(); ++si) {
    const Seg& sg = segs[si];
    for (int j = 0; j + ctx <= sg.len && cur < samples; ++j, ++cur) {
      note_sample(cur, sg, first, j);
      sample_segi[cur] = static_cast<int>(si);
      sample_j[cur] = j;
    }
  }
  const int produced = cur;
  std::vector<float> nat;
  if (use_nat) {
    nat.assign(segs.size() * static_cast<size_t>(dim), 0.0f);
    parallel_for(static_cast<int>(segs.size()), 4, [&](int s0, int s1) {
      for (int si = s0; si < s1; ++si) {
        const Seg& sg = segs[si];
        if (sg.len < ctx) continue;
        float* natp = nat.data() + static_cast<size_t>(si) * dim;
        for (int k = 0; k < dim; ++k) {
          float s6 = 0.0f;
          for (int q = 0; q < 6; ++q) {
            const int fr = sg.begin + q;
            const float v = fr < n_frames ? fea[static_cast<size_t>(fr) * dim + k] : 0.0f;
            s6 = q == 0 ? v : s6 + v;
          }
          natp[k] = s6 / 6.0f;
        }
      }
    });
  }
  // Scatter the input rows through the permutation: ~12 KB written per sample, the bulk of the host reader's work.
  parallel_for(produced, 64, [&](int c0, int c1) {
    for (int c = c0; c < c1; ++c) {
      const Seg& sg = segs[sample_segi[c]];
      float* row = in_dst + static_cast<size_t>(order[c]) * in_w;
      std::memcpy(row, fea.data() + static_cast<size_t>(sg.begin + sample_j[c]) * dim, sizeof(float) * dim * ctx);
      if (use_nat)
        std::memcpy(row + dim * ctx, nat.data() + static_cast<size_t>(sample_segi[c]) * dim, sizeof(float) * dim);
    }
  });

  // targets: frame j + targ_offset of each segment, not normalised (Interface.cc:815-816, 844-846)
  targ_reader.join();
  parallel_for(produced, 256, [&](int c0, int c1) {
    for (int c = c0; c < c1; ++c) {
      const Seg& sg = segs[sample_segi[c]];
      const float* src =
          trec.data() + static_cast<size_t>(sg.begin + sample_j[c] + para->targ_offset) * (out_w + 2) + 2;
      float* row = targ_dst + static_cast<size_t>(order[c]) * out_w;
      for (int k = 0; k < out_w; ++k) row[k] = be_float(src + k);
    }
  });
  (void)produced;
  return samples;
}

int Interface::Readchunk(int index) { return Readchunk(index, 0); }

int Interface::Readchunk(int index, int slot) {
  if (slot != 0) ensure_alt_buffers();
  return assemble(index, chunk_frame_st, total_chunks, total_samples, train_r.en, true, chunk_in(slot),
                  chunk_targ(slot));
}

int Interface::Readchunk_cv(int index) {
  return assemble(index, cv_chunk_frame_st, cv_total_chunks, cv_total_samples, cv_r.en, false, para->indata,
                  para->targ);
}

// Device-reader variant of `assemble`: same chunk geometry, same shuffle (one GetRandIndex call per
// chunk, so the lrand48 stream — hence every later chunk — is unchanged), but the records stay raw: the function only
// reads them into (page-locked) buffers and writes the sample table that bp_upload_raw_chunk consumes.
int Interface::assemble_raw(int chunk_index, const int* starts, unsigned int n_chunks, unsigned int n_samples,
                            int sent_end, bool shuffle, RawChunk* rc) {
  const int dim = para->fea_dim, ctx = para->fea_context;
  const int out_w = para->layersizes[numlayers - 1];
  const int first = starts[chunk_index];
  const bool last = static_cast<unsigned int>(chunk_index) == n_chunks - 1;
  const int n_frames = (last ? framesBeforeSent[sent_end] : starts[chunk_index + 1]) - first;
  const int samples = last ? static_cast<int>(n_samples) - para->traincache * chunk_index : para->traincache;
  rc->n_records = rc->n_samples = 0;
  if (n_frames <= 0 || samples <= 0) return 0;
  std::vector<int> order(samples);
  for (int i = 0; i < samples; ++i) order[i] = i;
  if (shuffle) GetRandIndex(order.data(), samples);

  auto grow = [&](void** buf, size_t* cap, size_t need) {
    if (need <= *cap) return;
    if (*buf) (host_free ? host_free : free)(*buf);
    *buf = host_alloc ? host_alloc(need) : malloc(need);
    if (!*buf) fatal("cannot allocate %zu bytes for the raw chunk.\n", need);
    *cap = need;
  };
  // The two record blocks are the only bulk host work left per chunk (~1 KB per frame): read them with a few
  // positional reads in parallel (page-cache copies run at 2-4 GB/s per thread), straight into the DMA buffers.
  std::atomic<bool> ok{true};
  std::vector<std::thread> readers;
  auto read_block = [&](FILE* fp, int width, void* dst) {
    const size_t rec_bytes = 4u * static_cast<size_t>(width + 2);
    const size_t total = rec_bytes * static_cast<size_t>(n_frames);
    const off_t base = static_cast<off_t>(kPfileHeaderBytes) + static_cast<off_t>(first) * rec_bytes;
    // 4 readers per block, 8 on hosts with >= 32 hardware threads (a page-cache copy runs at 2-4 GB/s per thread)
    const int parts = total > (8u << 20) ? (std::thread::hardware_concurrency() >= 32 ? 8 : 4) : 1;
    const size_t per = (total + parts - 1) / parts;
    const int fd = fileno(fp);
    for (int k = 0; k < parts; ++k) {
      const size_t b = std::min(total, per * k), e = std::min(total, per * (k + 1));
      readers.emplace_back([=, &ok] {
        size_t done = b;
        while (done < e) {
          const ssize_t got = pread(fd, static_cast<char*>(dst) + done, e - done, base + static_cast<off_t>(done));
          if (got <= 0) { ok = false; return; }
          done += static_cast<size_t>(got);
        }
      });
    }
  };
  grow(&rc->fea_records, &rc->fea_cap, static_cast<size_t>(n_frames) * (dim + 2) * 4);
  grow(&rc->targ_records, &rc->targ_cap, static_cast<size_t>(n_frames) * (out_w + 2) * 4);
  read_block(fp_data, dim, rc->fea_records);
  read_block(fp_targ, out_w, rc->targ_records);
  for (std::thread& t : readers) t.join();
  if (!ok) fatal("pfile read failed.\n");
  const int sent = be_int(static_cast<const float*>(rc->fea_records));

  rc->sample_frame.resize(samples);
  rc->sample_seg.resize(samples);
  rc->sample_row.resize(samples);
  sample_sent.assign(samples, 0);
  sample_frame_in_sent.assign(samples, 0);
  int cur = 0;
  for (const Seg& sg : segments(first, n_frames, sent))
    for (int j = 0; j + ctx <= sg.len && cur < samples; ++j, ++cur) {
      note_sample(cur, sg, first, j);
      rc->sample_frame[cur] = sg.begin + j;
      rc->sample_seg[cur] = sg.begin;
      rc->sample_row[cur] = order[cur];
    }
  if (cur != samples) fatal("chunk %d: planned %d samples but found %d.\n", chunk_index, samples, cur);
  rc->n_records = n_frames;
  rc->n_samples = samples;
  return samples;
}

int Interface::ReadchunkRaw(int index, RawChunk* rc) {
  return assemble_raw(index, chunk_frame_st, total_chunks, total_samples, train_r.en, true, rc);
}

int Interface::Readchunk_cvRaw(int index, RawChunk* rc) {
  return assemble_raw(index, cv_chunk_frame_st, cv_total_chunks, cv_total_samples, cv_r.en, false, rc);
}

void Interface::free_raw(RawChunk* rc) {
  for (void** b : {&rc->fea_records, &rc->targ_records}) {
    if (*b) (host_free ? host_free : free)(*b);
    *b = nullptr;
  }
  rc->fea_cap = rc->targ_cap = 0;
}

// Uniform weights from drand48 (Interface.cc:1036-1042): vec[i] = drand48()*(max-min)+min, evaluated in double.
void Interface::GetRandWeight(float* vec, float lo, float hi, int len) {
  for (int i = 0; i < len; ++i) vec[i] = drand48() * (hi - lo) + lo;
}

// Fisher-Yates from the back with lrand48 (Interface.cc:1044-1055).
void Interface::GetRandIndex(int* vec, int len) {
  for (int i = 0; i < len - 1; ++i) {
    const int idx = lrand48() % (len - i);
    const int tmp = vec[idx];
    vec[idx] = vec[len - 1 - i];
    vec[len - 1 - i] = tmp;
  }
}
