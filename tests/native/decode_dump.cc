// decode_dump — drives host/DecodeWriter.cc without a GPU: writes `n_sent` sentences of 3+s frames each with
// row[j] = 100*s + f + 0.25*j, in two append() calls, so tests/test_decode_writer.py can read the file back.
//   decode_dump <out> <raw|pfile> <norm_file|-> <dim> <n_sent>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "DecodeWriter.h"

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  const int dim = atoi(argv[4]), n_sent = atoi(argv[5]);
  DecodeWriter w;
  std::string err;
  if (!w.open(argv[1], argv[2], strcmp(argv[3], "-") ? argv[3] : "", dim, &err)) {
    printf("%s\n", err.c_str());
    return 1;
  }
  std::vector<float> rows;
  std::vector<int> sent, frame;
  for (int s = 0; s < n_sent; ++s) {
    if (s == 1) continue;  // a sentence too short to yield samples: must appear as an empty sentence
    for (int f = 0; f < 3 + s; ++f) {
      sent.push_back(s);
      frame.push_back(f + 5);
      for (int j = 0; j < dim; ++j) rows.push_back(100.0f * s + f + 0.25f * j);
    }
  }
  const int n = (int)sent.size(), half = n / 2;
  w.append(rows.data(), half, sent.data(), frame.data());
  w.append(rows.data() + (size_t)half * dim, n - half, sent.data() + half, frame.data() + half);
  w.close();
  return 0;
}
