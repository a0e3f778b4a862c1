// Host-side data / IO surface of BPtrain, re-written from scratch for the B200 build.
//
// Mirrors the reference's class `Interface` + struct `WorkPara` (reference Interface.h:8-101) member for member so
// that main() reads like the reference's (BPtrain.cc:16-101): Initial, get_pfile_info, get_chunk_info(_cv),
// Readchunk(_cv), GetRandIndex, Writeweights, and the public fields para / fp_log / total_chunks / ...
// File formats (Pfile, norm file, MAT-v4 .wts), argv keys, RNG call order (srand48 -> weight init -> chunk shuffle ->
// per-chunk sample shuffle) and log lines follow the reference; the implementation does not.
//
// Extensions (all optional, defaults = reference behaviour):
//   nat=0|1        noise-aware-training block on/off (default: inferred from layersizes[0], HEAD hard-wires it on and
//                  hard-codes fea_dim 129 — Interface.cc:395,777-778; here the block is fea_dim wide, identical at 129)
//   activation=relu|sigmoid   (HEAD = relu, a source edit in the reference: DevFunc.cu:67-97)
//   seed=<u64>     dropout seed (reference: time(NULL), BP_GPU.cu:77-78);  decode_file=<path> enhanced CV frames
#pragma once
#include <cstdio>
#include <string>
#include <vector>

#define MAXLAYER 10
#define MAXLINE 1024
#define MAXCHUNK 102400

struct WorkPara {
  char fea_FN[MAXLINE] = "";
  char fea_normFN[MAXLINE] = "";
  int fea_dim = 0;
  int fea_context = 0;

  char targ_FN[MAXLINE] = "";
  int targ_offset = 0;
  int dropoutflag = 0;
  int traincache = 0;  // samples per chunk held in host memory
  int bunchsize = 0;
  int layersizes[MAXLAYER] = {0};
  float momentum = 0.0f;
  float weightcost = 0.0f;
  float lrate = 0.0f;
  float visible_omit = 0.0f;
  float hid_omit = 0.0f;

  char init_weightFN[MAXLINE] = "";
  char out_weightFN[MAXLINE] = "";
  char log_FN[MAXLINE] = "";

  char train_sent_range[MAXLINE] = "";
  char cv_sent_range[MAXLINE] = "";

  int gpu_used = 1;
  int init_randem_seed = 0;
  float init_randem_weight_min = -0.1f;  // reference defaults, Interface.cc:79-82
  float init_randem_weight_max = 0.1f;
  float init_randem_bias_max = 0.1f;
  float init_randem_bias_min = -0.1f;

  float* indata = nullptr;               // traincache x layersizes[0]   (pinned when a GPU is present)
  float* targ = nullptr;                 // traincache x layersizes[last]
  float* weights[MAXLAYER - 1] = {nullptr};  // index 1..numlayers-1, w[in*n_out + out]
  float* bias[MAXLAYER - 1] = {nullptr};

  // extensions
  int nat = -1;            // -1 = infer from layersizes[0]
  int reader_gpu = 0;      // reader=gpu: splice / normalise / shuffle on the device (default host, as the reference)
  int prefetch = 1;        // the training chunks are read one ahead on a second thread (ChunkPrefetch.h)
  int activation = 0;      // 0 relu, 1 sigmoid
  unsigned long long seed = 0x5eed5eedULL;
  char decode_FN[MAXLINE] = "";
  char decode_format[MAXLINE] = "raw";   // raw | pfile (DecodeWriter.h)
  char decode_normFN[MAXLINE] = "";      // optional de-normalisation of the decode output
  // epochs > 1: the Perl driver's loop (finetune_DNN_speech_enhancement_dropout_NAT.pl:131-249) inside one process.
  // Epoch e (1-based, numbered from epoch_first) uses momentum min(momentum_max, momentum + (e-1)*momentum_step),
  // init_randem_seed + (e-1)*seed_step, and the outwts_file / log_file names with "%d" replaced by its number.
  int epochs = 1;
  int epoch_first = 1;
  float momentum_step = 0.04f;   // .pl:137
  float momentum_max = 0.9f;     // .pl:219
  int seed_step = 345;           // .pl:136
  char out_weight_pattern[MAXLINE] = "";  // the names as given on the command line
  char log_pattern[MAXLINE] = "";
};

// One chunk for the device-side reader: the Pfile records as they lie in the file + the sample table
// (field for field the bp_raw_chunk of include/bp_gpu.h; buffers are owned and re-used by the Interface).
struct RawChunk {
  int n_records = 0, n_samples = 0;
  void* fea_records = nullptr;   // n_records x (2+fea_dim) big-endian words (page-locked when host_alloc is set)
  void* targ_records = nullptr;  // n_records x (2+layersizes[last]) words
  size_t fea_cap = 0, targ_cap = 0;
  std::vector<int> sample_frame, sample_seg, sample_row;
};

// Pure helpers of the in-process epoch loop (tested on the CPU against the Perl driver's own arithmetic).
float epoch_momentum_value(double base, double step, double max, int e);
std::string epoch_name(const char* pattern, int number);

class Interface {
 public:
  Interface();
  ~Interface();

  void Initial(int argc, char** argv);
  // epochs > 1: switch log / weight files, RNG seed and momentum to epoch `e` (0-based offset from epoch_first);
  // e = 0 is what Initial() already set up.  Returns the epoch's momentum.
  float begin_epoch(int e);
  void Writeweights();
  void get_pfile_info();
  void get_chunk_info(char* range);
  void get_chunk_info_cv(char* range);
  int Readchunk(int index);  // into para->indata / para->targ, as the reference
  // Prefetching training loop (ChunkPrefetch.h): the same chunk into buffer pair `slot` (0 = para->indata/targ,
  // 1 = a second pair allocated on first use).
  int Readchunk(int index, int slot);
  float* chunk_in(int slot) const { return slot == 0 ? para->indata : alt_indata; }
  float* chunk_targ(int slot) const { return slot == 0 ? para->targ : alt_targ; }
  void ensure_alt_buffers();
  int Readchunk_cv(int index);
  // reader=gpu: same chunks, same shuffle, but samples are assembled on the device (bp_upload_raw_chunk)
  int ReadchunkRaw(int index, RawChunk* rc);
  int Readchunk_cvRaw(int index, RawChunk* rc);
  void free_raw(RawChunk* rc);
  // (sentence, frame-in-sentence) of the target frame of every sample of the chunk read last, in file order (= row
  // order for CV chunks, which are not shuffled): what a decode writer needs to label its rows
  std::vector<int> sample_sent, sample_frame_in_sent;
  const float* norm_mean() const { return mean.data(); }
  const float* norm_inv_std() const { return dVar.data(); }
  bool nat_block() const { return use_nat; }
  int cv_first_sentence() const { return cv_r.st; }
  void GetRandIndex(int* vec, int len);

  WorkPara* para;

  unsigned int total_frames = 0;
  unsigned int total_sents = 0;
  unsigned int total_chunks = 0;
  unsigned int total_samples = 0;
  unsigned int cv_total_chunks = 0;
  unsigned int cv_total_samples = 0;

  int* framesBeforeSent = nullptr;  // end-exclusive frame index of each sentence
  int* chunk_frame_st = nullptr;
  int* cv_chunk_frame_st = nullptr;

  FILE* fp_log = nullptr;
  int numlayers = 0;
  int realbunchsize = 0;

  // Set by main() when the trainer hands out page-locked chunk buffers (bp_host_alloc); freed with `host_free`.
  void* (*host_alloc)(size_t) = nullptr;
  void (*host_free)(void*) = nullptr;

 private:
  struct Range {
    int st = 0, en = 0;
  };
  struct Seg {
    int begin, len;  // record index inside the chunk's block, frames
    int sent;        // sentence the segment belongs to
  };
  void note_sample(int cur, const Seg& sg, int first, int j);
  std::vector<Seg> segments(int first, int n_frames, int sent) const;
  int assemble_raw(int chunk_index, const int* starts, unsigned int n_chunks, unsigned int n_samples, int sent_end,
                   bool shuffle, RawChunk* rc);
  void fatal(const char* fmt, ...);
  void echo_parameters();
  float epoch_momentum(int e) const;
  double base_momentum_d = 0.0, momentum_step_d = 0.04, momentum_max_d = 0.9;  // schedule in double (Perl arithmetic)
  int base_seed = 0;
  Range parse_range(const char* range, const char* what);
  void plan_chunks(const Range& r, int* starts, unsigned int* n_chunks, unsigned int* n_samples);
  int assemble(int chunk_index, const int* starts, unsigned int n_chunks, unsigned int n_samples, int sent_end,
               bool shuffle, float* in_dst, float* targ_dst);
  float* alloc_floats(size_t n, bool* pinned);
  void free_floats(float* p, bool pinned);
  float* alt_indata = nullptr;
  float* alt_targ = nullptr;
  bool pinned_in[2] = {false, false}, pinned_targ[2] = {false, false};
  void read_records(FILE* fp, int dim, long first_frame, int n_frames, std::vector<float>* rec, int* first_sent);
  void get_uint(const char* hdr, const char* argname, unsigned int* val);
  void read_tail(FILE* fp, long file_offset, unsigned int sentnum, int* out);
  void GetRandWeight(float* vec, float lo, float hi, int len);

  FILE* fp_data = nullptr;
  FILE* fp_targ = nullptr;
  FILE* fp_out = nullptr;
  std::vector<float> mean, dVar;
  std::vector<float> scratch_rec, scratch_trec, scratch_fea;  // host assembler's record / normalised-frame buffers
  Range train_r, cv_r;
  bool use_nat = true;
};
