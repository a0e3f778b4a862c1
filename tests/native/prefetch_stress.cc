// prefetch_stress — hand-over protocol of ChunkPrefetcher (host/ChunkPrefetch.h) under random timing: a reader that
// fills a slot with its chunk id after a random delay, a consumer that checks the slot holds exactly the expected id
// when it is handed over AND still when it is about to be released (i.e. the reader never writes a slot the consumer
// holds), the order of chunks, the -1 at the end, and early destruction with the reader blocked or mid-read.
//   prefetch_stress [n_chunks] [seed]      exit status 0 = all checks passed
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "ChunkPrefetch.h"

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 300;
  const unsigned seed = argc > 2 ? (unsigned)atoi(argv[2]) : 1u;
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = (i * 7919 + 13) % 100003;  // arbitrary, distinct-ish chunk ids
  std::vector<int> slot[2] = {std::vector<int>(4096, -1), std::vector<int>(4096, -1)};
  std::mt19937 rr(seed), rc(seed + 1);
  auto nap = [](std::mt19937& g) {
    const int us = (int)(g() % 4 == 0 ? g() % 300 : g() % 20);
    if (us) std::this_thread::sleep_for(std::chrono::microseconds(us));
  };
  int bad = 0;
  {
    ChunkPrefetcher ahead(order, [&](int chunk, int s) {
      nap(rr);
      for (int& v : slot[s]) v = chunk;  // a torn slot would show up as mixed ids on the consumer side
      return chunk % 1000;
    });
    for (int i = 0; i < n; ++i) {
      int s = -1;
      const int got = ahead.next(&s);
      if (got != order[i] % 1000 || s != (i & 1)) { printf("chunk %d: got %d in slot %d\n", i, got, s); ++bad; }
      for (int v : slot[s]) if (v != order[i]) { printf("chunk %d: slot holds %d at hand-over\n", i, v); ++bad; break; }
      nap(rc);
      for (int v : slot[s]) if (v != order[i]) { printf("chunk %d: slot overwritten while held (%d)\n", i, v); ++bad; break; }
    }
    int s = -1;
    if (ahead.next(&s) != -1) { printf("no end marker\n"); ++bad; }
    if (ahead.next(&s) != -1) { printf("end marker not sticky\n"); ++bad; }
  }
  for (int early = 0; early < 20; ++early) {  // destruction after `early` chunks: reader blocked or mid-read
    ChunkPrefetcher ahead(order, [&](int chunk, int s) {
      nap(rr);
      for (int& v : slot[s]) v = chunk;
      return 1;
    });
    int s = -1;
    for (int i = 0; i < early; ++i) ahead.next(&s);
    nap(rc);
  }
  ChunkPrefetcher empty(std::vector<int>{}, [](int, int) { return 0; });
  int s = -1;
  if (empty.next(&s) != -1) { printf("empty order: no end marker\n"); ++bad; }
  printf("prefetch_stress: %d chunks, %d failed checks\n", n, bad);
  return bad ? 1 : 0;
}
