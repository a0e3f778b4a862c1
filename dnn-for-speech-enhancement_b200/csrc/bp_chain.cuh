// Chained products: ONE persistent launch of CTA pairs walks a host-built schedule of tiles that belong to SEVERAL
// matrix products of a bunch (all forward layers; or the whole back-propagation: the dX chain and every dW product),
// with the dependencies between products expressed as counters in global memory instead of kernel boundaries.
//
// Why (measured, profiles/r2a, r2b): at bunch 1024 a hidden-layer product is 64 pair tiles on 74 SM pairs, ONE wave with
// one tile per CTA.  Its main loop runs at ~590 cycles per 64-deep k-block (87 % of the tensor pipe) = 10.4 us, but the
// launch is 17 us alone and 26-36 us inside a bunch: grid drain, launch, barrier/TMEM set-up, first-load latency and a
// non-overlapped epilogue are paid eleven times per bunch, 20 SMs idle throughout, and the dW products can only overlap
// the dX chain at kernel granularity.  Here a pair keeps its pipeline (shared-memory ring, barriers, TMEM) alive across
// tiles of different products: the epilogue of a tile overlaps the main loop of the next one (double-buffered TMEM),
// a tile of layer l starts as soon as the tiles of layer l-1 covering ITS frames are stored, and dW tiles fill the
// pairs the dX chain leaves idle.
//
// The reference runs the same products as ~11 cuBLAS calls + ~28 element-wise launches per bunch (BP_GPU.cu:484-673).
//
// Tile: 256 rows (2 CTAs x 128) x pair_n columns (128 or 256, per product), 64-deep k-blocks, tcgen05.mma
// cta_group::2 kind::tf32, operands of either major-ness chosen per product at run time (same 3-D tensor maps,
// descriptors and swizzles as bp_gemm2.cuh).  3 ring slots of 64 KB per CTA (A 32 KB + up to 32 KB of B), 2 x 256 TMEM
// columns.  Roles as in bp_gemm2.cuh: warp 0 TMA producer (both CTAs), warp 1 MMA issuer (leader CTA), warps 2-5
// epilogue (both CTAs); all three walk the same item list independently.
//
// Dependencies.  Product q owns one counter per n-tile; every epilogue warp of a pair (8) adds 1 to it after its part
// of the tile has been stored (stores -> fence.proxy.async -> __threadfence -> red.release.gpu).  A consumer item's
// PRODUCER warp polls (ld.acquire.gpu, bounded) the counters of the producer product that cover its frames — or all of
// them, for a dW product whose reduction runs over the frames — until each holds m_tiles * 8, then issues
// fence.proxy.async before its TMA loads (the data was written through the generic proxy, TMA reads through the async
// proxy).  Weights never change inside a chain launch.  Deadlock-freedom: the host's list scheduler starts an item only
// after its dependencies finished in simulated time, so "pair sequence + dependencies" is acyclic, and the grid is no
// larger than what is co-resident.  Two counter sets alternate between launches; a launch zeroes the other set.
#pragma once
#include "bp_chain_params.h"
#include "bp_gemm.cuh"

namespace bp {

__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add_u32(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Bounded poll of one dependency counter (one lane of the producer warp).
__device__ __forceinline__ void chain_wait_counter(const uint32_t* c, uint32_t target) {
  if (ld_acquire_gpu_u32(c) >= target) return;
  const long long t0 = clock64();
  while (ld_acquire_gpu_u32(c) < target) {
    __nanosleep(40);
    if (clock64() - t0 > kSpinLimit) {  // ~2 s: a scheduling bug becomes a CUDA error, never a hung box
      printf("bp_chain: dependency timeout block %d (counter %p = %u < %u)\n", (int)blockIdx.x, (const void*)c,
             ld_acquire_gpu_u32(c), target);
      __trap();
    }
  }
}

// MMA issue loop of ONE tile (leader CTA's warp 1), operand majors and tile width compiled in: see the call site.
template <bool kAMN, bool kBMN, int PAIR_N>
__device__ __forceinline__ void chain_mma_item(uint64_t* full, uint64_t* empty, uint64_t* tfull, uint32_t smem_base,
                                               uint32_t d_tmem, int num_it, int as, int& s, uint32_t& ph, int lane,
                                               uint32_t stage_bytes, int n_stages) {
  constexpr int BLOCK_M = GEMM_BLOCK_M, BLOCK_K = GEMM_BLOCK_K;
  constexpr uint32_t kDescHiK = (1024u >> 4) | (1u << 14) | (kLayoutSW128 << 29);
  constexpr uint32_t kDescHiMN = (512u >> 4) | (1u << 14) | (kLayoutSW128Base32 << 29);
  constexpr uint32_t kDescLoK = (16u >> 4) << 16;
  constexpr uint32_t kDescLoMN = ((uint32_t(BLOCK_K) * 128u) >> 4) << 16;
  constexpr uint32_t kAsub = BLOCK_M * 128, kBsub = (PAIR_N / 2) * 128;
  constexpr uint32_t idesc = make_idesc_tf32(2 * BLOCK_M, PAIR_N, kAMN ? 1u : 0u, kBMN ? 1u : 0u);
  constexpr int kPollLane = 1;
  for (int kb = 0; kb < num_it; ++kb) {
    if (lane == kPollLane) mbar_wait(&full[s], ph);
    __syncwarp();
    tc_fence_after();
    const uint32_t sa = smem_base + uint32_t(s) * stage_bytes;
    const uint32_t sb = sa + CHAIN_A_BYTES;
    if (elect_one()) {
      const uint32_t a_lo = (kAMN ? kDescLoMN : kDescLoK) + ((sa >> 4) & 0x3FFFu);
      const uint32_t b_lo = (kBMN ? kDescLoMN : kDescLoK) + ((sb >> 4) & 0x3FFFu);
#pragma unroll
      for (int k = 0; k < BLOCK_K / 8; ++k) {
        const uint32_t a_off = kAMN ? uint32_t(k) * 1024u : uint32_t(k / 4) * kAsub + uint32_t(k % 4) * 32u;
        const uint32_t b_off = kBMN ? uint32_t(k) * 1024u : uint32_t(k / 4) * kBsub + uint32_t(k % 4) * 32u;
        umma_tf32_2sm(d_tmem, a_lo + (a_off >> 4), kAMN ? kDescHiMN : kDescHiK, b_lo + (b_off >> 4),
                      kBMN ? kDescHiMN : kDescHiK, idesc, (k != 0 || kb != 0) ? 1u : 0u);
      }
      umma_commit_2sm(&empty[s], uint16_t(0x3u));
      if (kb == num_it - 1) umma_commit_2sm(&tfull[as], uint16_t(0x3u));
    }
    __syncwarp();
    if (++s == n_stages) { s = 0; ph ^= 1u; }
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) bp_chain_kernel(const __grid_constant__ ChainArgs a) {
  constexpr int BLOCK_M = GEMM_BLOCK_M, BLOCK_K = GEMM_BLOCK_K;
  constexpr uint32_t A_BYTES = CHAIN_A_BYTES;
  constexpr uint32_t TMEM_COLS = 512;
  const int kStages = a.n_stages;                 // kernel parameters: uniform
  const uint32_t STAGE_BYTES = a.stage_bytes;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw_addr);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + CHAIN_RING_BYTES);
  uint64_t* empty = full + CHAIN_MAX_STAGES;
  uint64_t* tfull = empty + CHAIN_MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kPollLane = 1;  // never the lane elect.sync picks, see bp_gemm.cuh
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1u;
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;

  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < a.n_counters; i += blockDim.x) a.counters_next[i] = 0u;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);  // 4 epilogue warps x 2 CTAs (used in the leader only)
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_ptr, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int it0 = a.pair_off[pair], it1 = a.pair_off[pair + 1];
  auto map_ptr = [&](int idx) -> const CUtensorMap* { return idx >= 0 ? a.maps + idx : &a.dyn[-1 - idx]; };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    int s = 0;
    uint32_t ph = 0;
    for (int i = it0; i < it1; ++i) {
      const ChainItem itm = a.items[i];
      const ChainProd& q = a.prods[itm.prod];
      const int pair_n = q.pair_n, half_n = pair_n >> 1;
      const bool amn = q.amn != 0, bmn = q.bmn != 0;
      const int m0 = itm.mt * 2 * BLOCK_M + int(rank) * BLOCK_M;
      const int n0 = itm.nt * pair_n + int(rank) * half_n;
      const int num_kb = (q.p.K + BLOCK_K - 1) / BLOCK_K;
      const int num_it = num_kb * (q.p.passes == 3 ? 3 : 1);
      const uint32_t b_bytes = uint32_t(half_n) * BLOCK_K * 4;
      const CUtensorMap* mA = map_ptr(q.map_a);
      const CUtensorMap* mB = map_ptr(q.map_b);
      const CUtensorMap* mAlo = map_ptr(q.map_a_lo);
      const CUtensorMap* mBlo = map_ptr(q.map_b_lo);
      const unsigned long long hint_a = q.p.hint_a, hint_b = q.p.hint_b;
      if (q.dep_prod >= 0) {  // operands written inside this launch: wait until the tiles that hold them are stored
        const ChainProd& d = a.prods[q.dep_prod];
        int j0 = 0, j1 = d.n_tiles;
        if (!q.dep_all) {
          const int f0 = itm.nt * pair_n, f1 = min(q.p.N, f0 + pair_n);
          j0 = f0 / d.pair_n;
          j1 = (f1 + d.pair_n - 1) / d.pair_n;
        }
        const uint32_t target = uint32_t(d.m_tiles) * CHAIN_ARRIVALS_PER_TILE;
        if (lane == kPollLane)
          for (int j = j0; j < j1; ++j) chain_wait_counter(a.counters + d.cnt_base + j, target);
        __syncwarp();
        fence_proxy_async_global();  // generic-proxy writes (acquired above) -> visible to my async-proxy (TMA) reads
      }
      if (a.trace != nullptr && leader && lane == 0) a.trace[4 * i + 0] = global_timer_ns();
      for (int it = 0, kb = 0, pass = 0; it < num_it; ++it, ++kb) {
        if (kb == num_kb) { kb = 0; ++pass; }
        const CUtensorMap* mapA = pass == 1 ? mAlo : mA;
        const CUtensorMap* mapB = pass == 2 ? mBlo : mB;
        if (lane == kPollLane) mbar_wait(&empty[s], ph ^ 1u);
        __syncwarp();
        if (elect_one()) {
          uint8_t* sa = smem + size_t(s) * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (leader) mbar_expect_tx(&full[s], 2 * (A_BYTES + b_bytes));  // both CTAs' bytes land on the leader's barrier
          const int a1 = amn ? kb * BLOCK_K : m0, a2 = amn ? m0 / 32 : kb * (BLOCK_K / 32);
          const int b1 = bmn ? kb * BLOCK_K : n0, b2 = bmn ? n0 / 32 : kb * (BLOCK_K / 32);
          if (hint_a) tma_load_3d_2sm_hint(sa, mapA, &full[s], 0, a1, a2, hint_a);
          else tma_load_3d_2sm(sa, mapA, &full[s], 0, a1, a2);
          if (hint_b) tma_load_3d_2sm_hint(sb, mapB, &full[s], 0, b1, b2, hint_b);
          else tma_load_3d_2sm(sb, mapB, &full[s], 0, b1, b2);
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int i = it0; i < it1; ++i) {
        const ChainItem itm = a.items[i];
        const ChainProd& q = a.prods[itm.prod];
        const int pair_n = q.pair_n;
        const bool amn = q.amn != 0, bmn = q.bmn != 0;
        const int num_it = ((q.p.K + BLOCK_K - 1) / BLOCK_K) * (q.p.passes == 3 ? 3 : 1);
        if (lane == kPollLane) mbar_wait(&tempty[as], aph ^ 1u);
        __syncwarp();
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * 256);
        // The instruction and shared-memory descriptors must reach tcgen05.mma in UNIFORM registers.  Per-product
        // values read from the table are ordinary (vector) registers to the compiler, and every one of them would be
        // moved across for every MMA (~5 moves x 8 MMAs per k-block: the first version of this loop ran at 1120 cycles
        // per k-block where bp_gemm2_kernel needs 590).  So the product's (majors, width) only SELECTS one of the
        // compile-time variants below, in which all of them are immediates.
        const int variant = (amn ? 2 : 0) + (bmn ? 1 : 0) + (pair_n == 256 ? 4 : 0);
        switch (variant) {
          case 2: chain_mma_item<true, false, 128>(full, empty, tfull, smem_base, d_tmem, num_it, as, s, ph, lane, STAGE_BYTES, kStages); break;
          case 0: chain_mma_item<false, false, 128>(full, empty, tfull, smem_base, d_tmem, num_it, as, s, ph, lane, STAGE_BYTES, kStages); break;
          case 3: chain_mma_item<true, true, 128>(full, empty, tfull, smem_base, d_tmem, num_it, as, s, ph, lane, STAGE_BYTES, kStages); break;
          case 6: chain_mma_item<true, false, 256>(full, empty, tfull, smem_base, d_tmem, num_it, as, s, ph, lane, STAGE_BYTES, kStages); break;
          case 4: chain_mma_item<false, false, 256>(full, empty, tfull, smem_base, d_tmem, num_it, as, s, ph, lane, STAGE_BYTES, kStages); break;
          case 7: chain_mma_item<true, true, 256>(full, empty, tfull, smem_base, d_tmem, num_it, as, s, ph, lane, STAGE_BYTES, kStages); break;
          default: __trap();  // K-major A with MN-major B: no product of the path has it
        }
        as ^= 1;
        if (as == 0) aph ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    const int qd = warp & 3;
    int as = 0;
    uint32_t aph = 0;
    for (int i = it0; i < it1; ++i) {
      const ChainItem itm = a.items[i];
      const ChainProd& q = a.prods[itm.prod];
      const GemmParams& p = q.p;   // constant bank: uniform, read-only to the compiler
      const int epi = q.epi, has_consumer = q.has_consumer, cnt_base = q.cnt_base;
      const int pair_n = q.pair_n;
      const int m0 = itm.mt * 2 * BLOCK_M + int(rank) * BLOCK_M;
      const int n0 = itm.nt * pair_n;
      const int m = m0 + qd * 32 + lane;
      const bool m_ok = m < p.M;
      DxPrefetch pre;
      // Y (EPI_DX) / the targets (training output layer) of the first chunks, fetched under the main loop
      const bool out_train = epi == EPI_FWD_OUT && p.aux != nullptr && p.out != nullptr && p.out2 == nullptr;
      if (epi == EPI_DX || out_train) pre.start(p, m, m_ok, n0);
      if (lane == 0) mbar_wait_backoff(&tfull[as], aph);
      __syncwarp();
      tc_fence_after();
      if (a.trace != nullptr && leader && warp == 2 && lane == 0) a.trace[4 * i + 2] = global_timer_ns();
      const uint32_t taddr = tmem_base + (uint32_t(qd * 32) << 16) + uint32_t(as * 256);
      float sq_local = 0.0f;
      if (epi == EPI_DX) {
        if (pair_n == 256) gemm_dx_epilogue<256>(p, pre, taddr, m, m_ok, n0);
        else gemm_dx_epilogue<128>(p, pre, taddr, m, m_ok, n0);
      } else if (out_train) {
        const float bias = m_ok ? __ldg(p.bias + m) : 0.0f;
        if (pair_n == 256) sq_local = gemm_fwdout_train_epilogue<256>(p, pre, taddr, m, m_ok, n0, bias, sq_local);
        else sq_local = gemm_fwdout_train_epilogue<128>(p, pre, taddr, m, m_ok, n0, bias, sq_local);
      } else {
        float bias = 0.0f;
        if (epi != EPI_PLAIN && m_ok) bias = __ldg(p.bias + m);
        const int chunks = pair_n >> 5;
#pragma unroll 1
        for (int c = 0; c < chunks; ++c) {
          const int nc = n0 + c * 32;
          if (nc >= p.N) break;
          uint32_t v[32];
          tmem_ld32(taddr + uint32_t(c * 32), v);
          tmem_ld_wait();
          if (epi == EPI_FWD_HID) gemm_epilogue_chunk<EPI_FWD_HID>(p, v, m, m_ok, nc, bias, sq_local);
          else if (epi == EPI_FWD_OUT) gemm_epilogue_chunk<EPI_FWD_OUT>(p, v, m, m_ok, nc, bias, sq_local);
          else gemm_epilogue_chunk<EPI_PLAIN>(p, v, m, m_ok, nc, bias, sq_local);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty[as]);
        else mbar_arrive_remote(&tempty[as], crank & ~1u);
      }
      if (a.trace != nullptr && leader && warp == 2 && lane == 0) a.trace[4 * i + 1] = global_timer_ns();
      // my part of the tile is stored: publish it to the consumers' producer warps (products nobody waits for inside
      // this launch — the dW gradients — are complete at kernel end like any kernel's output)
      if (has_consumer) {
        fence_proxy_async_global();
        __syncwarp();  // orders the other lanes' stores before lane 0's release (cumulative at gpu scope)
        if (lane == 0) red_release_gpu_add_u32(a.counters + cnt_base + itm.nt, 1u);
      }
      if (a.trace != nullptr && leader && warp == 2 && lane == 0) a.trace[4 * i + 3] = global_timer_ns();
      if (epi == EPI_FWD_OUT && p.sqerr != nullptr) {
        double sq = static_cast<double>(sq_local);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
        if (lane == 0 && sq != 0.0) atomicAdd(p.sqerr, sq);
      }
      as ^= 1;
      if (as == 0) aph ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves while the peer may still signal it or read its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

}  // namespace bp
