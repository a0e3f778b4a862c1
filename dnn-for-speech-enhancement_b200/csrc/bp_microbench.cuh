// Bring-up microbenchmark: sustained tcgen05.mma kind::tf32 rate of ONE SM for a given operand-major combination and
// N, with operands resident in (uninitialised) shared memory — no TMA, no per-stage barriers.  Answers "how many cycles
// does a 128 x N x 8 TF32 MMA occupy the tensor pipe when both operands come from shared memory?".
#pragma once
#include "bp_ptx.cuh"

namespace bp {

template <bool kAMN, bool kBMN, int BN>
__global__ void __launch_bounds__(64, 1) bp_mma_rate_kernel(int iters, int mode, long long* out) {
  // mode bit 0: tcgen05.fence::after_thread_sync before every group of 8; bit 1: commit after every group of 8
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint64_t scratch;
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = 128 * 64 * 4, B_BYTES = BN * 64 * 4;
  constexpr uint32_t hiK = (1024u >> 4) | (1u << 14) | (kLayoutSW128 << 29);
  constexpr uint32_t hiMN = (512u >> 4) | (1u << 14) | (kLayoutSW128Base32 << 29);
  constexpr uint32_t loK = (16u >> 4) << 16, loMN = ((64u * 128u) >> 4) << 16;
  constexpr uint32_t idesc = make_idesc_tf32(128, BN, kAMN ? 1u : 0u, kBMN ? 1u : 0u);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&scratch, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t d = tmem_slot;
  if (warp == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode & 1) tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t ao = kAMN ? k * 1024u : (k / 4) * (128u * 128u) + (k % 4) * 32u;
          const uint32_t bo = kBMN ? k * 1024u : (k / 4) * (uint32_t(BN) * 128u) + (k % 4) * 32u;
          umma_tf32_lohi(d, (kAMN ? loMN : loK) | ((base + ao) >> 4), kAMN ? hiMN : hiK,
                         (kBMN ? loMN : loK) | ((base + A_BYTES + bo) >> 4), kBMN ? hiMN : hiK, idesc, 1u);
        }
        if (mode & 2) umma_commit(&scratch);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) {
      out[0] = t1 - t0;  // issue time
      out[1] = t2 - t0;  // issue + drain
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(d, BN);
  }
}

// Store-pattern microbenchmark: what the memory system does with an epilogue's writes.  Every CTA (128 threads) writes
// tiles of 128 rows x 128 floats: thread t writes float t of every row (one coalesced 512-byte row segment per CTA and
// row, 128 bytes per warp instruction — exactly the GEMM epilogues' pattern).  row_stride = 128: the tile is ONE
// contiguous 64 KB region; row_stride = ld of an activation matrix: 128 segments of 512 B, ld*4 bytes apart.
__global__ void __launch_bounds__(128) bp_store_pattern_kernel(float* buf, long long row_stride, long long tile_stride,
                                                               int tiles_per_panel, long long panel_stride,
                                                               long long n_tiles, float v) {
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    float* base = buf + (t / tiles_per_panel) * panel_stride + (t % tiles_per_panel) * tile_stride + threadIdx.x;
#pragma unroll 32
    for (int n = 0; n < 128; ++n) base[n * row_stride] = v;
  }
}

}  // namespace bp
