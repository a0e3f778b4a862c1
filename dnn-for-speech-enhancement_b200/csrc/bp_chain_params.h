// Host/device interface of the chained-products kernel (bp_chain.cuh): product table, schedule items, launch arguments,
// and the host-side list scheduler that turns a set of dependent products into per-pair item sequences.
#pragma once
#include <cuda.h>

#include <algorithm>
#include <cstdint>
#include <vector>

#include "bp_gemm_params.h"

namespace bp {

constexpr int CHAIN_MAX_STAGES = 4;
constexpr uint32_t CHAIN_A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 4;   // 32 KB: my 128 rows of A
constexpr uint32_t CHAIN_RING_BYTES = 6 * CHAIN_A_BYTES;              // 192 KB: 3 slots of 64 KB (some product has
                                                                      // 256-wide tiles: 32 KB of B per CTA) or 4 of 48 KB
constexpr int CHAIN_MAX_PROD = 20;
constexpr int CHAIN_MAX_ITEMS = 1400;   // tiles of one launch (C4's back-propagation: 562); larger nets use one launch
                                        // per product
constexpr int CHAIN_MAX_PAIRS = 80;
constexpr int CHAIN_ARRIVALS_PER_TILE = 8;                            // epilogue warps of a pair

struct ChainProd {
  GemmParams p;          // extents and epilogue parameters (bp_gemm_params.h); p.n_begin must be 0
  int epi;               // Epi: EPI_PLAIN | EPI_FWD_HID | EPI_FWD_OUT | EPI_DX
  int amn, bmn;          // 1 = MN-major operand (see bp_gemm.cuh "Operand fetch")
  int pair_n;            // 128 | 256
  int map_a, map_b;      // index into ChainArgs::maps; negative: ChainArgs::dyn[-1 - index]
  int map_a_lo, map_b_lo;
  int m_tiles, n_tiles;
  int cnt_base;          // my counters: counters[cnt_base + n_tile]
  int dep_prod;          // product that writes an operand of mine inside this launch, or -1
  int dep_all;           // 1: wait for all its n-tiles; 0: for those covering my tile's columns (frames)
  int per_bunch;         // bit 0: the host sets p.aux / p.sqerr to this bunch's targets / loss slot (output layer)
  int has_consumer;      // 1: some product of this launch waits for my tiles -> publish them (fence + counter)
};

struct ChainItem {
  short prod, mt, nt, pad;
};

// The whole launch description travels BY VALUE in the kernel parameters (constant bank, ~20 KB): the kernel indexes
// the item list and the product table with warp-uniform indices, so every per-tile quantity (tensor-map addresses,
// coordinates, descriptors, extents) is a UNIFORM value to the compiler and reaches the TMA / tcgen05 instructions
// without per-instruction vector-to-uniform register moves.  (A first version kept the tables in global / shared
// memory: ~20 R2UR per k-block in the producer, 40 in the MMA issuer, and a main loop 1.5x slower than bp_gemm2_kernel's.)
struct ChainArgs {
  ChainProd prods[CHAIN_MAX_PROD];
  ChainItem items[CHAIN_MAX_ITEMS];
  int pair_off[CHAIN_MAX_PAIRS + 1];   // items of pair i: [pair_off[i], pair_off[i+1])
  int n_prods;
  const CUtensorMap* maps;   // static tensor maps, global memory
  CUtensorMap dyn[4];        // per-bunch maps (this bunch's input rows as B operand of layer 1: fwd, dW; + low parts)
  uint32_t stage_bytes;      // ring slot: 32 KB of A + the widest product's B half (48 or 64 KB)
  int n_stages;              // CHAIN_RING_BYTES / stage_bytes (4 or 3)
  uint32_t* counters;        // this launch's counter set (all zero at launch)
  uint32_t* counters_next;   // the other set: zeroed by this launch
  int n_counters;
  unsigned long long* trace; // bring-up aid (null in production): per item 4 globaltimer stamps written by the leader
                             // CTA: dependencies satisfied, stores issued, accumulator complete, published
};
static_assert(sizeof(ChainArgs) < 32000, "kernel parameter space");

constexpr size_t chain_smem_bytes() {
  return size_t(CHAIN_RING_BYTES) + 1024 /*align*/ + 256 /*barriers*/;
}


// ---------------------------------------------------------------------------------------------- host: list scheduler
// Shape of one product as the scheduler sees it.
struct ChainShape {
  int m_tiles, n_tiles;  // pair tiles (256 rows x pair_n columns)
  int n_cols;            // N (columns = frames for forward / dX, fan-in + 1 for dW)
  int kb;                // 64-deep k-blocks per tile (x 3 in split-precision mode)
  int pair_n;
  int dep_prod, dep_all; // as ChainProd
  int prio;              // smaller = scheduled first among ready items (the dX chain before the dW filler)
};

// Tile width of a product inside a chain: 256-wide pair tiles (shared-memory roof = tensor roof) when they still cover
// >= 60 % of the pairs, else 128-wide ones (twice the tiles, shorter dependency chain) — the rule of pick_kernel.
inline int chain_pair_n(int M, int N, int pairs) {
  const int mt = (M + 255) / 256;
  return mt * ((N + 255) / 256) * 10 >= pairs * 6 ? 256 : 128;
}

// The products of one train bunch of `rows` frames as the scheduler sees them (sizes[0..L] = layer sizes).
//   which = 0, forward: product l-1 = layer l (M = units of l, N = frames, K = fan-in), waits for the tiles of layer
//           l-1 that cover its frames.
//   which = 1, back-propagation: dX_L .. dX_2 (product that turns dE/dX_l into dE/dX_{l-1}: M = units of l-1,
//           N = frames, K = units of l; waits for the dX product before it, same frames; priority 0 = the chain), then
//           dW_L .. dW_1 (M = units of l, N = fan-in + 1, K = frames, 256-wide; waits for ALL of dE/dX_l; the filler).
// bp_runtime.cu builds its product tables in exactly this order and checks them against this function.
inline std::vector<ChainShape> chain_shapes(const int* sizes, int L, int rows, int pairs, int passes, int which) {
  std::vector<ChainShape> v;
  const int mul = passes == 3 ? 3 : 1;
  auto mk = [&](int M, int N, int K, int pair_n, int dep, int dep_all, int prio) {
    ChainShape sh{};
    sh.m_tiles = (M + 255) / 256;
    sh.n_tiles = (N + pair_n - 1) / pair_n;
    sh.n_cols = N;
    sh.kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K * mul;
    sh.pair_n = pair_n;
    sh.dep_prod = dep;
    sh.dep_all = dep_all;
    sh.prio = prio;
    v.push_back(sh);
  };
  if (which == 0) {
    for (int l = 1; l <= L; ++l) mk(sizes[l], rows, sizes[l - 1], chain_pair_n(sizes[l], rows, pairs), l >= 2 ? l - 2 : -1, 0, l);
  } else {
    std::vector<int> dx_index(L + 2, -1);
    for (int l = L; l >= 2; --l) {
      dx_index[l] = (int)v.size();
      mk(sizes[l - 1], rows, sizes[l], chain_pair_n(sizes[l - 1], rows, pairs), l < L ? dx_index[l + 1] : -1, 0, 0);
    }
    for (int l = L; l >= 1; --l) mk(sizes[l], sizes[l - 1] + 1, rows, 256, l < L ? dx_index[l + 1] : -1, 1, 1 + (L - l));
  }
  return v;
}

// Cost model in SM cycles, from the clock64 traces of the pair kernels (profiles/r2b): a 64-deep k-block of a
// 128-wide pair tile takes ~600 cycles in steady state, of a 256-wide one ~1100; an epilogue ~2600 / ~4000.
inline long long chain_mainloop_cycles(const ChainShape& q) { return (long long)q.kb * (q.pair_n == 256 ? 1100 : 600) + 1200; }
inline long long chain_epilogue_cycles(const ChainShape& q) { return q.pair_n == 256 ? 4000 : 2600; }

// Event-driven list scheduling on `pairs` CTA pairs.  Repeatedly: take the pair that becomes free first; among the
// items whose dependencies have all been scheduled, give it the one that can start earliest on it (ties: priority,
// then product / tile order).  A pair is free for its next item's main loop when the current main loop ends (the
// epilogue overlaps: double-buffered TMEM); an item's results exist when its epilogue ends.  Every item starts, in
// simulated time, after its dependencies finished, and a pair's sequence has increasing start times — so the union of
// the pair sequences and the dependency edges is acyclic: no deadlock on the device, whatever the real timing is.
// Returns the simulated makespan (cycles).
inline long long chain_schedule(const std::vector<ChainShape>& prods, int pairs, std::vector<ChainItem>& items,
                                std::vector<int>& pair_off) {
  struct It {
    int prod, mt, nt;
    long long ready;   // all dependencies' results exist (-1: some dependency not scheduled yet)
    bool done;
  };
  std::vector<It> all;
  std::vector<int> first(prods.size() + 1, 0);
  for (size_t q = 0; q < prods.size(); ++q) {
    first[q] = (int)all.size();
    for (int nt = 0; nt < prods[q].n_tiles; ++nt)
      for (int mt = 0; mt < prods[q].m_tiles; ++mt) all.push_back({(int)q, mt, nt, -1, false});
  }
  first[prods.size()] = (int)all.size();
  // finish time of (product, n-tile) = max over its m-tiles; -1 while some m-tile is unscheduled
  std::vector<std::vector<long long>> col_done(prods.size());
  std::vector<std::vector<int>> col_left(prods.size());
  for (size_t q = 0; q < prods.size(); ++q) {
    col_done[q].assign(prods[q].n_tiles, 0);
    col_left[q].assign(prods[q].n_tiles, prods[q].m_tiles);
  }
  auto ready_time = [&](const It& it) -> long long {
    const ChainShape& q = prods[it.prod];
    if (q.dep_prod < 0) return 0;
    const ChainShape& d = prods[q.dep_prod];
    int j0 = 0, j1 = d.n_tiles;
    if (!q.dep_all) {
      const int f0 = it.nt * q.pair_n, f1 = std::min(q.n_cols, f0 + q.pair_n);
      j0 = f0 / d.pair_n;
      j1 = (f1 + d.pair_n - 1) / d.pair_n;
    }
    long long t = 0;
    for (int j = j0; j < j1; ++j) {
      if (col_left[q.dep_prod][j] > 0) return -1;
      t = std::max(t, col_done[q.dep_prod][j]);
    }
    return t;
  };
  std::vector<long long> avail(pairs, 0);
  std::vector<std::vector<ChainItem>> seq(pairs);
  size_t left = all.size();
  long long makespan = 0;
  while (left > 0) {
    int p = 0;
    for (int i = 1; i < pairs; ++i)
      if (avail[i] < avail[p]) p = i;
    // earliest possible start on this pair over all ready items ...
    long long min_start = -1;
    for (size_t i = 0; i < all.size(); ++i) {
      It& it = all[i];
      if (it.done) continue;
      it.ready = ready_time(it);
      if (it.ready < 0) continue;
      const long long start = std::max(it.ready, avail[p]);
      if (min_start < 0 || start < min_start) min_start = start;
    }
    if (min_start < 0) return -1;  // dependency cycle in the product table (a bug of the caller)
    // ... but a higher-priority item that becomes ready within `slack` of it goes first: a pair that finishes a
    // critical-path tile is free an epilogue earlier than the next critical-path tiles are ready, and must not grab a
    // long filler tile in between
    const long long slack = 8000;
    int best = -1;
    long long best_start = 0;
    for (size_t i = 0; i < all.size(); ++i) {
      const It& it = all[i];
      if (it.done || it.ready < 0) continue;
      const long long start = std::max(it.ready, avail[p]);
      if (start > min_start + slack) continue;
      const int pr = prods[it.prod].prio;
      if (best < 0 || pr < prods[all[best].prod].prio || (pr == prods[all[best].prod].prio && start < best_start)) {
        best = (int)i;
        best_start = start;
      }
    }
    It& it = all[best];
    const ChainShape& q = prods[it.prod];
    it.done = true;
    --left;
    avail[p] = best_start + chain_mainloop_cycles(q);
    const long long fin = avail[p] + chain_epilogue_cycles(q);
    col_done[it.prod][it.nt] = std::max(col_done[it.prod][it.nt], fin);
    --col_left[it.prod][it.nt];
    makespan = std::max(makespan, fin);
    seq[p].push_back(ChainItem{(short)it.prod, (short)it.mt, (short)it.nt, 0});
  }
  items.clear();
  pair_off.assign(pairs + 1, 0);
  for (int p = 0; p < pairs; ++p) {
    pair_off[p] = (int)items.size();
    items.insert(items.end(), seq[p].begin(), seq[p].end());
  }
  pair_off[pairs] = (int)items.size();
  return makespan;
}

}  // namespace bp
