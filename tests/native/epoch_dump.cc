// Test helper: prints the in-process epoch schedule (host/Interface.cc: epoch_momentum_value, epoch_name) so that
// tests/test_epoch_schedule.py can compare it with the arithmetic of the reference's Perl driver, run by perl itself.
//   epoch_dump <momentum> <step> <max> <epochs> <pattern>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "Interface.h"

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  const double base = atof(argv[1]), step = atof(argv[2]), mx = atof(argv[3]);
  const int n = atoi(argv[4]);
  for (int e = 0; e < n; ++e) {
    const float m = epoch_momentum_value(base, step, mx, e);
    unsigned int bits;
    memcpy(&bits, &m, 4);
    printf("%d %08x %s\n", e + 1, bits, epoch_name(argv[5], e + 1).c_str());
  }
  return 0;
}
