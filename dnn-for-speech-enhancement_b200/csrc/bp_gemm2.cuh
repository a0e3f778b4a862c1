// CTA-pair variant of bp_gemm.cuh: tcgen05.mma.cta_group::2 on a 256 x 256 pair tile (same products, same epilogues).
//
// Why: with fp32 operands a lone CTA's 128x128 tile needs 128 KB of shared-memory traffic (64 KB TMA fill + 64 KB
// UMMA operand reads) per 64-deep k-block = 1024 cycles at 128 B/clk for 512 tensor-pipe cycles: a 50 % roof (measured,
// DESIGN.md section 5).  In a pair, each CTA stages its own 128 rows of A and only HALF of the 256 B columns; one MMA
// instruction issued by the leader CTA drives both SMs' tensor cores (each computing 128 x 256 x 8 in 128 cycles) and
// reads the two B halves from both CTAs' shared memory.  Per CTA and k-block: 64 KB fill + 64 KB reads for 1024 pipe
// cycles -> the shared-memory roof coincides with the tensor roof, and the issue overhead per MMA is amortised over
// twice the work.
//
// Protocol (cluster of 2 CTAs; rank 0 = leader):
//   * TMEM: 512 columns allocated with cta_group::2 by warp 1 of both CTAs (2 tile buffers x 256 fp32 columns; each CTA
//     holds its own 128 accumulator rows).
//   * full[s]   (leader only, count 1): the leader's producer arms expect_tx for BOTH CTAs' bytes; all four TMA loads
//                of a stage (each CTA: its A rows, its B half, into its own smem) signal the LEADER's barrier
//                (cp.async.bulk.tensor ... .cta_group::2 with the leader's barrier address).
//   * empty[s]  (both CTAs, count 1): the leader's tcgen05.commit.cta_group::2 multicasts one arrival to both CTAs.
//   * tfull[b]  (both CTAs, count 1): same multicast commit after a tile's last k-block; each CTA's epilogue waits on
//                its own copy.
//   * tempty[b] (leader only, count 8): one arrival per epilogue warp of BOTH CTAs (the peer's arrive remotely).
//   The peer CTA's warp 1 only allocates / frees TMEM.  Everything else (3-D tensor maps, descriptors, warp-uniform
//   role loops, elect.sync, one polling lane, PDL, 3xTF32 passes) is as in bp_gemm.cuh.
//
// Tried and removed (measurements in DESIGN.md section 8): clusters of 2 / 4 pairs sharing their A rows by TMA multicast
// (no gain alone, -3 % forward / +15 % dW inside a bunch), a 2-deep ring with two CTAs per SM (-25 %), L2-only TMA
// prefetch ahead of the ring (-4..-9 %).
#pragma once
#include "bp_gemm.cuh"

namespace bp {

// Pair tile: 256 rows (2 x 128) x PAIR_N columns.  PAIR_N = 256: 64 KB stages x 3, shared-memory roof = tensor roof.
// PAIR_N = 128 (for products too small to fill the machine with 256-wide pair tiles, e.g. 2048 x 1024 outputs):
// 48 KB stages x 4, each CTA stages 64 B columns; 96 KB of shared-memory traffic per 512 pipe cycles -> 67 % roof
// instead of the lone CTA's 50 %.
template <int PAIR_N>
__host__ __device__ constexpr int gemm2_stages() { return PAIR_N == 256 ? 3 : 4; }

template <int PAIR_N>
constexpr size_t gemm2_smem_bytes() {
  return size_t(gemm2_stages<PAIR_N>()) * (GEMM_BLOCK_M + PAIR_N / 2) * GEMM_BLOCK_K * 4 + 1024 + 256;
}

template <bool kAMN, bool kBMN, int kEpi, int PAIR_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
bp_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo,
                const GemmParams p) {
  constexpr int BLOCK_M = GEMM_BLOCK_M, BLOCK_K = GEMM_BLOCK_K, BLOCK_N = PAIR_N;
  constexpr int kStages = gemm2_stages<PAIR_N>();
  static_assert(PAIR_N == 128 || PAIR_N == 256, "PAIR_N");
  constexpr int HALF_N = BLOCK_N / 2;                    // B columns staged by each CTA
  constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 4;    // my 128 rows of A
  constexpr uint32_t B_BYTES = HALF_N * BLOCK_K * 4;     // my half of B
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
  static_assert(TMEM_COLS <= 512, "TMEM");
  constexpr uint32_t kDescHiK = (1024u >> 4) | (1u << 14) | (kLayoutSW128 << 29);
  constexpr uint32_t kDescHiMN = (512u >> 4) | (1u << 14) | (kLayoutSW128Base32 << 29);
  constexpr uint32_t kDescLoK = (16u >> 4) << 16;
  constexpr uint32_t kDescLoMN = ((uint32_t(BLOCK_K) * 128u) >> 4) << 16;
  constexpr uint32_t kAsub = BLOCK_M * 128;
  constexpr uint32_t kBsub = HALF_N * 128;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw_addr);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(kStages) * STAGE_BYTES);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kPollLane = 1;
  const uint32_t rank = cluster_ctarank();     // position in my pair (0 = leader)
  const bool leader = rank == 0;

  const int num_m_tiles = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);            // pair tiles along M
  const int num_n_tiles = (p.N - p.n_begin + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int num_it = num_kb * (p.passes == 3 ? 3 : 1);
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);  // 4 epilogue warps x 2 CTAs (used in the pair leader only)
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
  cluster_sync_all();  // both CTAs' barriers exist before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();
  griddep_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    int s = 0;
    uint32_t ph = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int m0 = (t % num_m_tiles) * 2 * BLOCK_M + int(rank) * BLOCK_M;       // my 128 rows of A
      const int n0 = p.n_begin + (t / num_m_tiles) * BLOCK_N + int(rank) * HALF_N;  // my half of B
      for (int it = 0, kb = 0, pass = 0; it < num_it; ++it, ++kb) {
        if (kb == num_kb) { kb = 0; ++pass; }
        const CUtensorMap* mapA = pass == 1 ? &tmAlo : &tmA;
        const CUtensorMap* mapB = pass == 2 ? &tmBlo : &tmB;
        if (lane == kPollLane) mbar_wait(&empty[s], ph ^ 1u);
        __syncwarp();
        if (elect_one()) {
          uint8_t* sa = smem + size_t(s) * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (leader) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);  // both CTAs' bytes land on the leader's barrier
          const int b1 = kBMN ? kb * BLOCK_K : n0, b2 = kBMN ? n0 / 32 : kb * (BLOCK_K / 32);
          const int a1 = kAMN ? kb * BLOCK_K : m0, a2 = kAMN ? m0 / 32 : kb * (BLOCK_K / 32);
          if (p.hint_a) tma_load_3d_2sm_hint(sa, mapA, &full[s], 0, a1, a2, p.hint_a);
          else tma_load_3d_2sm(sa, mapA, &full[s], 0, a1, a2);
          if (p.hint_b) tma_load_3d_2sm_hint(sb, mapB, &full[s], 0, b1, b2, p.hint_b);
          else tma_load_3d_2sm(sb, mapB, &full[s], 0, b1, b2);
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_tf32(2 * BLOCK_M, BLOCK_N, kAMN ? 1u : 0u, kBMN ? 1u : 0u);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        if (lane == kPollLane) mbar_wait(&tempty[as], aph ^ 1u);
        __syncwarp();
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * BLOCK_N);
        for (int kb = 0; kb < num_it; ++kb) {
          if (lane == kPollLane) mbar_wait(&full[s], ph);
          __syncwarp();
          tc_fence_after();
          const uint32_t sa = smem_base + uint32_t(s) * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
          if (elect_one()) {
            // 14-bit start-address field (a shared-window address carries the CTA's cluster rank in its high bits)
            const uint32_t a_lo = (kAMN ? kDescLoMN : kDescLoK) + ((sa >> 4) & 0x3FFFu);
            const uint32_t b_lo = (kBMN ? kDescLoMN : kDescLoK) + ((sb >> 4) & 0x3FFFu);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 8; ++k) {
              const uint32_t a_off = kAMN ? uint32_t(k) * 1024u : uint32_t(k / 4) * kAsub + uint32_t(k % 4) * 32u;
              const uint32_t b_off = kBMN ? uint32_t(k) * 1024u : uint32_t(k / 4) * kBsub + uint32_t(k % 4) * 32u;
              umma_tf32_2sm(d_tmem, a_lo + (a_off >> 4), kAMN ? kDescHiMN : kDescHiK, b_lo + (b_off >> 4),
                            kBMN ? kDescHiMN : kDescHiK, idesc, (k != 0 || kb != 0) ? 1u : 0u);
            }
            umma_commit_2sm(&empty[s], uint16_t(0x3u));                           // slot s is free in both CTAs
            if (kb == num_it - 1) umma_commit_2sm(&tfull[as], uint16_t(0x3u));    // accumulators complete
          }
          __syncwarp();
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
        as ^= 1;
        if (as == 0) aph ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    const int q = warp & 3;
    int as = 0;
    uint32_t aph = 0;
    float sq_local = 0.0f;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int m0 = (t % num_m_tiles) * 2 * BLOCK_M + int(rank) * BLOCK_M;  // my 128 accumulator rows
      const int n0 = p.n_begin + (t / num_m_tiles) * BLOCK_N;  // all columns of my pair's tile
      const int m = m0 + q * 32 + lane;
      const bool m_ok = m < p.M;
      // EPI_DX: the Y values of the first two column chunks are fetched BEFORE waiting for the accumulator, i.e. under
      // the main loop, and the following chunks two steps ahead of their use (see gemm_dx_epilogue).
      // EPI_FWD_OUT in training (targets present, D_L stored, no raw output): the targets are pipelined like Y
      bool out_train = false;
      if constexpr (kEpi == EPI_FWD_OUT) out_train = p.aux != nullptr && p.out != nullptr && p.out2 == nullptr;
      DxPrefetch pre;
      if (kEpi == EPI_DX || out_train) pre.start(p, m, m_ok, n0);
      if (lane == 0) mbar_wait_backoff(&tfull[as], aph);
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BLOCK_N);
      float bias = 0.0f;
      if constexpr (kEpi == EPI_FWD_HID || kEpi == EPI_FWD_OUT) {
        if (m_ok) bias = __ldg(p.bias + m);
      }
      if constexpr (kEpi == EPI_DX) {
        gemm_dx_epilogue<BLOCK_N>(p, pre, taddr, m, m_ok, n0);
      } else if (out_train) {
        sq_local = gemm_fwdout_train_epilogue<BLOCK_N>(p, pre, taddr, m, m_ok, n0, bias, sq_local);
      } else {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          const int nc = n0 + c * 32;
          if (nc >= p.N) break;
          uint32_t v[32];
          tmem_ld32(taddr + uint32_t(c * 32), v);
          tmem_ld_wait();
          gemm_epilogue_chunk<kEpi>(p, v, m, m_ok, nc, bias, sq_local);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty[as]);
        else mbar_arrive_remote(&tempty[as], 0u);
      }
      as ^= 1;
      if (as == 0) aph ^= 1u;
    }
    if constexpr (kEpi == EPI_FWD_OUT) {
      if (p.sqerr != nullptr) {
        double sq = static_cast<double>(sq_local);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
        if (lane == 0 && sq != 0.0) atomicAdd(p.sqerr, sq);
      }
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
