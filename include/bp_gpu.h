/* bp_gpu.h — C-ABI of libbpgpu.so: the B200-native (sm_100a) replacement for the reference's trainer object.
 *
 * The reference's drop-in seam is the C++ class BP_GPU (reference BP_GPU.h:40-88) used by main() at exactly four
 * sites: constructor BPtrain.cc:31-32, train :53, returnWeights :57, CrossValid :77, destructor :96.
 * Each entry point below names the reference member it replaces.  Plain pointers and sizes only; all `float*`
 * arguments are HOST memory owned by the caller unless stated otherwise.
 *
 * Error convention: the reference prints and exit(0)s (BP_GPU.cu:20-24, 929-933).  Here every call returns
 * BP_OK (0) or a negative BP_E* code and bp_last_error() gives the message; the CLI maps that to "log + exit(0)".
 * There is NO CPU fallback: without a usable sm_100 device bp_create fails with BP_ENODEV.
 */
#ifndef BP_GPU_H_
#define BP_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BP_MAXLAYER 10 /* reference BP_GPU.h:13 */

#define BP_OK 0
#define BP_EINVAL (-1)  /* bad argument */
#define BP_ENODEV (-2)  /* no sm_100 CUDA device / gpu_used out of range (reference BP_GPU.cu:17-25) */
#define BP_ECUDA (-3)   /* CUDA runtime / driver error */
#define BP_ENOMEM (-4)
#define BP_ECOMM (-5)   /* NCCL error */

#define BP_ACT_RELU 0    /* HEAD: DevFunc.cu:67-97 */
#define BP_ACT_SIGMOID 1 /* commented variant: DevFunc.cu:48-64 */

#define BP_MATH_TF32 0   /* single-pass TF32 tensor-core products, fp32 accumulate (throughput mode) */
#define BP_MATH_3XTF32 1 /* hi/lo split, three TF32 products per term: ~fp32 accuracy (parity mode) */

typedef struct bp_handle bp_handle;

/* Extended construction parameters (everything the reference fixes at compile time or leaves to the wall clock). */
typedef struct bp_config {
  int device;       /* CUDA device ordinal for this rank */
  int world_size;   /* data-parallel ranks (1 = single GPU) */
  int rank;         /* this rank, 0..world_size-1 */
  int numlayers;    /* number of layer SIZES (weight layers = numlayers-1), <= BP_MAXLAYER */
  int layersizes[BP_MAXLAYER];
  int bunchsize;    /* GLOBAL minibatch rows per step; each rank processes bunchsize/world_size rows */
  float lrate, momentum, weightcost;
  int dropoutflag;
  float visible_omit, hid_omit;
  int activation;   /* BP_ACT_* */
  int math_mode;    /* BP_MATH_* */
  uint64_t seed;    /* dropout RNG seed (reference: time(NULL), BP_GPU.cu:77-78) */
} bp_config;

/* BP_GPU::BP_GPU (BP_GPU.cu:10-197).  weights[i], bias[i] valid for i = 1..numlayers-1 (index 0 unused, as in the
 * reference); weights[i] = float[layersizes[i-1]*layersizes[i]] laid out w[in*n_out + out].  Copied, not retained.
 * gpu_used > 1 runs gpu_used data-parallel ranks (one host thread + NCCL rank per GPU) inside this process. */
int bp_create(bp_handle** h, int gpu_used, int numlayers, const int* layersizes, int bunchsize, float lrate,
              float momentum, float weightcost, float* const* weights, float* const* bias, int dropoutflag,
              float visible_omit, float hid_omit);

/* Same, with the extended configuration; one handle == one rank (one process per GPU under torchrun). */
int bp_create_ex(bp_handle** h, const bp_config* cfg, float* const* weights, float* const* bias);

/* BP_GPU::~BP_GPU (BP_GPU.cu:199-238). */
void bp_destroy(bp_handle* h);

/* BP_GPU::train (BP_GPU.cu:241-331): in = n_frames x layersizes[0], targ = n_frames x layersizes[last], row-major.
 * Uploads the chunk and runs floor(n_frames/bunch) full bunches of forward + back-prop + SGD update; a trailing
 * partial bunch is skipped exactly like the reference (:297-318).  The host buffers are not read after return.
 * For a rank handle (world_size > 1) n_frames/in/targ are this rank's shard: local bunch = bunchsize/world_size. */
int bp_train(bp_handle* h, int n_frames, const float* in, const float* targ);

/* BP_GPU::CrossValid (BP_GPU.cu:408-479): forward only (partial last bunch included, :450); *sum_sq_err receives
 * sum over frames and output dims of (out - targ)^2.  With dropoutflag the keep probability scales the products
 * (:705-746).  Weights are not modified (the reference's in-place rescale drift, :726-746, is not replicated). */
int bp_crossvalid(bp_handle* h, int n_frames, const float* in, const float* targ, float* sum_sq_err);

/* Decode extension of cv_bunch_single (BP_GPU.cu:676-773): writes the enhanced frames, out = n_frames x
 * layersizes[last] (the reference computes them and throws them away, :445-473). */
int bp_forward(bp_handle* h, int n_frames, const float* in, float* out);

/* BP_GPU::returnWeights (BP_GPU.cu:910-923): fills caller-owned arrays, same indexing as bp_create. */
int bp_return_weights(bp_handle* h, float* const* weights, float* const* bias);

/* Epoch boundary inside one process.  The reference runs one epoch per process and its Perl
 * driver re-invokes the binary with a new momentum (finetune_DNN_speech_enhancement_dropout_NAT.pl:131-165, 214-249):
 * every new process starts with zero momentum deltas (BP_GPU.cu:137-138 + zero-fill :938-940), the constructor's
 * lrate / momentum / weightcost (BP_GPU.cu:10) and — here — dropout step 0.  bp_begin_epoch puts a live trainer into
 * exactly that state without the weight round-trip through a .wts file: weights stay on the device, deltas are zeroed,
 * the three hyper-parameters are replaced, and the dropout step counter restarts iff reset_dropout_step != 0. */
int bp_begin_epoch(bp_handle* h, float lrate, float momentum, float weightcost, int reset_dropout_step);

/* Re-keys the dropout masks (Philox key) from the next bunch on.  The reference seeds cuRAND from time(NULL) in every
 * process (BP_GPU.cu:77-78), so its masks differ from epoch to epoch; BPtrain derives the key from seed= and
 * init_randem_seed (which the Perl driver advances by 345 per epoch, .pl:136) and calls this at every in-process epoch
 * boundary, so that an epoch draws the masks a fresh process started for it would draw. */
int bp_set_dropout_seed(bp_handle* h, uint64_t seed);

/* Run-time scheduling switches (defaults come from the environment variable of the same meaning at bp_create time;
 * scripts/SWITCHES.md lists them with their measured effect).  Results are unaffected — bit for bit, except "splitk",
 * which changes the summation order of the output layer.  Known names:
 *   "chain"          -1 automatic / 0 one launch per matrix product / 1 chained launches: all forward products of a bunch
 *                    in one persistent launch, the dX chain + every dW product in a second one (csrc/bp_chain.cuh).
 *                    Automatic = chained for data-parallel ranks on peer memory, per product on a single GPU (BP_CHAIN)
 *   "chain_trace"    1: bring-up aid, per-tile timestamps of the chained launches (bp_debug_chain_trace)
 *   process-wide switches: "pdl", "tma_hint", "pairs", "sgd_stream", "sgd_early", "splitk" — changeable between calls so
 *   that alternatives are A/B-timed inside one process (scripts/gpu_ab_quick.py)
 * Returns BP_EINVAL for an unknown name. */
int bp_set_option(bp_handle* h, const char* name, int value);

/* Read-only state of a handle, for the benchmark line and the tests:
 *   "dp_exchange"  0 = single rank, 1 = NCCL all-reduce, 2 = peer-memory exchange fused into the kernels
 *   "sm_clock_mhz" SM clock measured ON the device right now (a one-thread kernel on the compute stream compares
 *                  clock64 with globaltimer over ~50 us, after whatever is queued on that stream) — says at which
 *                  clock a timing loop really ran, which NVML sampling from the host cannot for loops of a few ms
 *   any bp_set_option name: its current value. */
int bp_get_option(bp_handle* h, const char* name, int* value);

/* Last error message of the calling thread ("" if none). */
const char* bp_last_error(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Device-resident control (measurement and pipelining): the same work as bp_train / bp_crossvalid split into
 * "upload a chunk" and "run bunches on the resident chunk" so that a benchmark can time the kernels with inputs
 * already in HBM, and a reader thread can overlap the next upload.
 * ------------------------------------------------------------------------------------------------------------- */
int bp_upload_chunk(bp_handle* h, int n_frames, const float* in, const float* targ /* may be NULL */);
/* NB with dropoutflag = 1 and visible_omit > 0 the input mask is applied IN PLACE to the resident rows, as the
 * reference does to its device copy (BP_GPU.cu:536-540): a row can be trained once per upload; training it again is
 * refused (BP_EINVAL) instead of silently stacking masks. */
int bp_train_resident(bp_handle* h, int first_bunch, int n_bunches);
int bp_forward_resident(bp_handle* h, int first_frame, int n_frames, float* out_host /* may be NULL */,
                        double* sum_sq_err /* may be NULL; needs resident targets */);
int bp_sync(bp_handle* h);

/* ---------------------------------------------------------------------------------------------------------------
 * Device-side chunk reader.  Replaces the sample assembly of Interface::Readchunk /
 * Readchunk_cv (Interface.cc:689-861, 864-1034): the caller passes the chunk's Pfile records exactly as they lie in
 * the file plus a table that says which record window each sample is; byte-swap, (x-mean)*inv_std (:745-746), the
 * fea_context-frame splice (:769-773), the NAT block = mean of the segment's first six frames (:776-779), the
 * target pick at +targ_offset (:844-846) and the scatter through the shuffle permutation (:731-735) run on the
 * device.  Rows are bit-identical with the host reader's.  Host traffic drops from 4*(layersizes[0]+layersizes[L])
 * bytes per sample to ~4*(fea_dim+layersizes[L]+4) bytes per frame.
 *
 * Data-parallel handles (in-process group or one process per rank): every rank is given the same bp_raw_chunk and keeps
 * rows [rank*B/G, (rank+1)*B/G) of every global bunch of B rows; the trailing partial bunch is dropped.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct bp_raw_chunk {
  int fea_dim;              /* floats per feature record */
  int fea_context;          /* frames spliced per sample */
  int targ_offset;          /* target record = first context record + targ_offset */
  int nat;                  /* 1 = append the NAT block (layersizes[0] = fea_dim*(fea_context+1)) */
  int n_records;            /* records in fea_records (and targ_records) */
  int n_samples;
  const void* fea_records;  /* n_records x (2+fea_dim) big-endian 32-bit words: sentence id, frame id, features */
  const void* targ_records; /* n_records x (2+layersizes[last]) words, or NULL (forward only) */
  const float* mean;        /* fea_dim, norm file part 1 */
  const float* inv_std;     /* fea_dim, norm file part 2 */
  const int* sample_frame;  /* n_samples: index of the sample's first context record */
  const int* sample_seg;    /* n_samples: index of the first record of its sentence segment (nat only) */
  const int* sample_row;    /* n_samples: destination row (the shuffle permutation), or NULL for file order */
} bp_raw_chunk;
int bp_upload_raw_chunk(bp_handle* h, const bp_raw_chunk* chunk);  /* then bp_train_resident / bp_forward_resident */
int bp_train_raw(bp_handle* h, const bp_raw_chunk* chunk);         /* = upload + train on floor(n_samples/B) bunches */
/* CrossValid (+ optional decode output, n_samples x layersizes[last]) over a raw chunk; all rows on device 0. */
int bp_crossvalid_raw(bp_handle* h, const bp_raw_chunk* chunk, float* sum_sq_err /* may be NULL */,
                      float* out /* may be NULL */);
/* Pipelined decode of a stream of raw chunks (BASELINE config C5; the reference computes the enhanced frames in
 * cv_bunch_single and drops them, BP_GPU.cu:445-473).  bp_decode_raw_submit queues records -> rows -> forward pass ->
 * device-to-host copy into `out` (n_samples x layersizes[last]; page-locked memory, bp_host_alloc, makes the copy
 * asynchronous) and returns as soon as the chunk's records and tables have been consumed (the caller may refill them).
 * bp_decode_raw_wait blocks until the OLDEST submitted chunk's frames are complete in its `out`.  At most two chunks
 * are in flight (a third submit returns BP_EINVAL), so the steady-state loop is  submit(k+1); wait()  [-> chunk k] :
 * upload of k+1, forward of k+1 and the read-back of k overlap.  Device 0 only; no targets needed. */
int bp_decode_raw_submit(bp_handle* h, const bp_raw_chunk* chunk, float* out);
int bp_decode_raw_wait(bp_handle* h);
/* The same pipeline for callers that keep the reference's host reader: `in` = n_frames spliced, normalised rows (what
 * bp_forward / BP_GPU::CrossValid take, BP_GPU.cu:408-479).  Shares the two in-flight slots with the raw variant;
 * bp_forward_submit returns when `in` may be refilled, bp_forward_wait when the oldest chunk's `out` is complete. */
int bp_forward_submit(bp_handle* h, int n_frames, const float* in, float* out);
int bp_forward_wait(bp_handle* h);
/* Reads rows of the resident chunk back (tests: bit-exactness of the device reader).  Single-rank handles. */
int bp_download_chunk(bp_handle* h, int first_row, int n_rows, float* in /* may be NULL */,
                      float* targ /* may be NULL */);

/* Pinned host memory for true asynchronous DMA of chunks (the reference uses pageable new[] buffers). */
void* bp_host_alloc(size_t bytes);
void bp_host_free(void* p);

/* CUDA-event stopwatch on the handle's compute stream. */
int bp_timer_start(bp_handle* h);
int bp_timer_stop(bp_handle* h, float* elapsed_ms);

/* Counters since creation: kernels launched by this library on this handle; train bunches executed. */
int bp_get_counters(bp_handle* h, uint64_t* kernel_launches, uint64_t* train_bunches);

/* Per-kernel device time of the most recent bunch when profiling is on (CUDA events around each launch class):
 * ms[0] = forward GEMMs, ms[1] = dX GEMMs (dW of the upper layers runs beside them), ms[2] = exposed tail of the dW
 * GEMMs, ms[3] = final SGD update launch (layer 1, or everything when BP_SGD_EARLY=0), ms[4] = all-reduce wait,
 * ms[5] = early SGD update of layers >= 2, which overlaps the first layer's dW GEMM.  Sums over the last <= 64 profiled bunches; records events only, no host sync per bunch. */
int bp_set_profiling(bp_handle* h, int on);
int bp_get_profile(bp_handle* h, float ms[6], uint64_t* bunches_profiled);
/* Launch timeline (bp_set_profiling(h, 2); measurement aid): an event behind EVERY launch of a training bunch, on the
 * stream it was issued on, for the next <= 16 bunches.  bp_get_timeline returns, per launch in issue order, its
 * label (newline-separated in `labels`) and its completion time in ms since the bunch's start mark, averaged over the
 * recorded bunches — the latency of the dependency chain launch by launch, caches as the pipeline leaves them.  The
 * records sit between the kernels, so programmatic dependent launch cannot overlap across them (~3 % pessimistic). */
int bp_get_timeline(bp_handle* h, char* labels, int labels_len, float* ms, int max_marks, int* n_marks,
                    int* bunches /* may be NULL */);

/* Training-loss monitor (our extension): sum over this handle's rows and output dims of (out - targ)^2 for each bunch
 * of the most recent (age 0) or the previous (age 1) bp_train / bp_train_resident call, accumulated in the
 * output-layer epilogue.  Waits only for that call to finish, so reading age 1 overlaps the current call's compute. */
int bp_train_losses(bp_handle* h, int age, double* out, int max_n, int* n_out);

/* Data-parallel communicator (NCCL over NVLink): rank 0 obtains an id, every rank passes the same 128 bytes. */
int bp_comm_unique_id(char id128[128]);
int bp_comm_init(bp_handle* h, const char id128[128]);

/* Mask that the fused dropout draws for (activation tensor `layer`, global frame, unit) at bunch `step`;
 * exposed so tests can replay masks against the oracle.  Returns 1 = dropped, 0 = kept. */
int bp_dropout_mask(uint64_t seed, uint32_t step, uint32_t layer, uint32_t frame, uint32_t unit, float p);

int bp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BP_GPU_H_ */
