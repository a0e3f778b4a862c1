// Decode output writer.  The reference computes the enhanced LPS frames in CrossValid and throws them
// away (BP_GPU.cu:445-473; the commented fwrite at :464-466); its users re-implement the forward pass in MATLAB to get
// them.  This writer stores the network output of the CV/decode pass:
//   decode_format=raw    little-endian float32 rows, one per sample (default)
//   decode_format=pfile  a QuickNet Pfile with the same record / sentence-table layout the reader consumes
//                        (Interface.cc:468-555): (sentence id, frame id, out_dim big-endian floats) per sample, sentence
//                        ids relative to the first decoded sentence, frame id = index of the target frame in its
//                        sentence, so the file lines up with the clean-target Pfile frame by frame
//   decode_norm_file=F   optional de-normalisation y = out / inv_std + mean with a norm file of out_dim entries in the
//                        reader's format (Interface.cc:301-326), i.e. the inverse of (x - mean) * inv_std
#pragma once
#include <cstdio>
#include <string>
#include <vector>

class DecodeWriter {
 public:
  // returns false (with `err` set) when a file cannot be opened or the norm file is malformed
  bool open(const char* path, const char* format, const char* norm_file, int out_dim, std::string* err);
  bool active() const { return fp_ != nullptr; }
  // n rows of out_dim floats; sent[i] / frame[i] = sentence (relative) and frame-in-sentence of row i
  void append(const float* rows, int n, const int* sent, const int* frame);
  void close();

 private:
  FILE* fp_ = nullptr;
  bool pfile_ = false;
  int dim_ = 0;
  std::vector<float> mean_, inv_std_;
  std::vector<float> scratch_;
  std::vector<unsigned int> frames_per_sent_;
  unsigned int n_frames_ = 0;
};
