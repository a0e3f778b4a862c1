#include "DecodeWriter.h"

#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace {
constexpr long kHeaderBytes = 32768;  // PFILE_HEADER_SIZE, reference Interface.cc:13
inline uint32_t be32(uint32_t v) { return __builtin_bswap32(v); }
}  // namespace

bool DecodeWriter::open(const char* path, const char* format, const char* norm_file, int out_dim, std::string* err) {
  dim_ = out_dim;
  pfile_ = format && strcmp(format, "pfile") == 0;
  if (format && format[0] && !pfile_ && strcmp(format, "raw") != 0) {
    *err = std::string("decode_format must be raw or pfile, not ") + format;
    return false;
  }
  if (norm_file && norm_file[0]) {
    FILE* fn = fopen(norm_file, "rt");
    if (!fn) {
      *err = std::string("can not open decode norm file: ") + norm_file;
      return false;
    }
    char line[1024];
    mean_.assign(dim_, 0.0f);
    inv_std_.assign(dim_, 1.0f);
    for (std::vector<float>* v : {&mean_, &inv_std_}) {
      bool ok = fgets(line, sizeof line, fn) != nullptr;
      for (int j = 0; ok && j < dim_; ++j) {
        ok = fgets(line, sizeof line, fn) != nullptr;
        if (ok) (*v)[j] = static_cast<float>(atof(line));
      }
      if (!ok) {
        fclose(fn);
        *err = "decode norm file too short";
        return false;
      }
    }
    fclose(fn);
  }
  fp_ = fopen(path, "wb");
  if (!fp_) {
    *err = std::string("can not open decode file: ") + path;
    return false;
  }
  if (pfile_) {  // header is written on close, when the counts are known
    std::vector<char> zero(kHeaderBytes, 0);
    fwrite(zero.data(), 1, kHeaderBytes, fp_);
  }
  return true;
}

void DecodeWriter::append(const float* rows, int n, const int* sent, const int* frame) {
  if (!fp_ || n <= 0) return;
  const float* src = rows;
  if (!mean_.empty()) {
    scratch_.resize(static_cast<size_t>(n) * dim_);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < dim_; ++j)
        scratch_[static_cast<size_t>(i) * dim_ + j] = rows[static_cast<size_t>(i) * dim_ + j] / inv_std_[j] + mean_[j];
    src = scratch_.data();
  }
  if (!pfile_) {
    fwrite(src, sizeof(float), static_cast<size_t>(n) * dim_, fp_);
    n_frames_ += n;
    return;
  }
  std::vector<uint32_t> rec(static_cast<size_t>(dim_) + 2);
  for (int i = 0; i < n; ++i) {
    rec[0] = be32(static_cast<uint32_t>(sent[i]));
    rec[1] = be32(static_cast<uint32_t>(frame[i]));
    const uint32_t* w = reinterpret_cast<const uint32_t*>(src + static_cast<size_t>(i) * dim_);
    for (int j = 0; j < dim_; ++j) rec[2 + j] = be32(w[j]);
    fwrite(rec.data(), 4, rec.size(), fp_);
    if (frames_per_sent_.size() <= static_cast<size_t>(sent[i])) frames_per_sent_.resize(sent[i] + 1, 0);
    frames_per_sent_[sent[i]]++;
    n_frames_++;
  }
}

void DecodeWriter::close() {
  if (!fp_) return;
  if (pfile_) {
    // sentence table: (num_sentences + 1) cumulative start frames, big-endian (Interface.cc:493, read_tail :1083)
    const unsigned int ns = static_cast<unsigned int>(frames_per_sent_.size());
    uint32_t acc = 0;
    uint32_t w = be32(acc);
    fwrite(&w, 4, 1, fp_);
    for (unsigned int s = 0; s < ns; ++s) {
      acc += frames_per_sent_[s];
      w = be32(acc);
      fwrite(&w, 4, 1, fp_);
    }
    std::string fmt = "dd" + std::string(dim_, 'f');
    char hdr[kHeaderBytes];
    memset(hdr, 0, sizeof hdr);
    const unsigned long long cells = static_cast<unsigned long long>(n_frames_) * (2 + dim_);
    snprintf(hdr, sizeof hdr,
             "-pfile_header version 0 size %ld\n-num_sentences %u\n-num_frames %u\n-first_feature_column 2\n"
             "-num_features %d\n-first_label_column %d\n-num_labels 0\n-format %s\n"
             "-data size %llu offset 0 ndim 2 nrow %u ncol %d\n-sent_table_data size %u offset %llu ndim 1\n-end\n",
             kHeaderBytes, ns, n_frames_, dim_, 2 + dim_, fmt.c_str(), cells, n_frames_, 2 + dim_, ns + 1, cells);
    fseek(fp_, 0, SEEK_SET);
    fwrite(hdr, 1, kHeaderBytes, fp_);
  }
  fclose(fp_);
  fp_ = nullptr;
}
