// tcgen05 / TMA / TMEM GEMM for the frame-wise DNN hot path (sm_100a only).
//
// One kernel template serves the three contractions the reference issues through cuBLAS
// (SgemmNN fwd DevFunc.h:45-55 / BP_GPU.cu:558,737; SgemmTN dX DevFunc.h:29-43 / BP_GPU.cu:636;
//  SgemmNT dW DevFunc.h:57-67 / BP_GPU.cu:642) with the reference's element-wise kernels fused into the
// epilogue (kernMultiCopy bias DevFunc.cu:166, kernSigmoid :67, kernDropout :34, kernSubClean :253,
// kernDsigmoid :81 + kernVecMul :244, kernAccSumrow :224 via an all-ones input column).
//
// Orientation ("units on lanes"): every product is computed as  Dt[M x N] = A[M x K] * B[N x K]^T  where the
// M (TMEM-lane) dimension is the one that is CONTIGUOUS in the output's memory, so that a warp's epilogue store of
// one accumulator column is a single coalesced 128-byte row segment:
//     fwd : M = units of layer l      N = frames          K = fan-in      A = W^T (MN-major)  B = Yprev (K-major)
//     dX  : M = units of layer l-1    N = frames          K = units of l  A = W   (K-major)   B = dEdX  (K-major)
//     dW  : M = units of layer l      N = fan-in (+1)     K = frames      A = dEdX^T (MN)     B = Yprev^T (MN)
// Output element (row m, col n) is stored at out[n * ldo + m].
//
// L2 traffic: with 128x128 tiles every CTA streams 32 KB per 32-deep k-block, i.e. 32 flop/B, and 128+ CTAs doing
// that exceed the L2->SM fabric (~10 TB/s measured) at a third of the TF32 tensor peak.  Thread-block clusters of
// CM x CN CTAs therefore share operand tiles by TMA multicast: the CN CTAs of a cluster row (same m-tile) split the A
// tile's 4-KB chunks between them and multicast each chunk to the whole row, likewise the CM CTAs of a column for B.
// A CTA's smem slot is written by its peers, so the slot's "empty" barrier counts one tcgen05.commit arrival from
// every CTA of its row and column (commit multicast), and a producer waits for all of them before re-filling.
//
// Pipeline: warp 0 = TMA producer (one elected lane), warp 1 = tcgen05.mma issuer (one elected lane) + TMEM owner,
// warps 2..5 = epilogue (tcgen05.ld -> registers -> fused math -> coalesced global stores).  kStages-deep smem ring
// (full/empty mbarriers), double-buffered TMEM accumulator (tfull/tempty mbarriers), persistent static tile loop.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "bp_ptx.cuh"
#include "bp_rng.cuh"

namespace bp {

enum Epi : int {
  EPI_PLAIN = 0,      // out = acc                               (dW gradient tile, debug)
  EPI_FWD_HID = 1,    // y = act(scale*acc + bias[m]) (+dropout)  (hidden layer forward)
  EPI_FWD_OUT = 2,    // o = scale*acc + bias[m]; optional out2 = o; optional out = gscale*(o - targ); optional sqerr
  EPI_DX = 3,         // out = act'(aux) * acc                    (back-prop through the non-linearity)
};

struct GemmParams {
  int M, N, K;            // logical extents (see header comment)
  float* out;             // primary output, element (m,n) at out[n*ldo + m]; may be null for EPI_FWD_OUT
  long long ldo;
  const float* bias;      // bias[m]                      (EPI_FWD_*)
  const float* aux;       // targ (EPI_FWD_OUT) or Y (EPI_DX), element (m,n) at aux[n*ldaux + m]
  long long ldaux;
  float* out2;            // raw linear output (EPI_FWD_OUT), element (m,n) at out2[n*ldo2 + m]
  long long ldo2;
  double* sqerr;          // if non-null (EPI_FWD_OUT): += sum (o - targ)^2
  float scale;            // multiplies the accumulator (inference-time keep probability, BP_GPU.cu:705-732)
  float gscale;           // 2/B of kernSubClean (DevFunc.cu:263)
  int act;                // 0 = ReLU (HEAD, DevFunc.cu:67-97), 1 = sigmoid (commented variant :52,:62)
  float drop_p;           // >0: zero y where u < drop_p (kernDropout DevFunc.cu:34-45), no rescale
  uint32_t seed_lo, seed_hi, step, layer;
  int frame0;             // global frame index of column n=0 (data-parallel shard offset), multiple of 4
  uint32_t dbg_mn_lbo, dbg_mn_sbo;  // bring-up overrides for the MN-major descriptor strides (0 = default)
  uint32_t dbg_flags;               // bit 0: skip the MMAs (TMA-only pipeline), bit 1: skip the TMA loads (MMA-only)
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 32;  // 32 fp32 = one 128-byte swizzle span
constexpr int GEMM_THREADS = 192;

template <int BLOCK_N, int kStages>
constexpr size_t gemm_smem_bytes() {
  return size_t(kStages) * (GEMM_BLOCK_M * GEMM_BLOCK_K * 4 + BLOCK_N * GEMM_BLOCK_K * 4) + 1024 /*align slack*/ +
         256 /*barriers*/;
}

__device__ __forceinline__ float act_fwd(float x, int act) {
  if (act == 0) return x > 0.0f ? x : 0.0f;
  return 1.0f / (1.0f + expf(-x));
}
__device__ __forceinline__ float act_bwd(float y, float e, int act) {
  if (act == 0) return y > 0.0f ? e : 0.0f;       // dydx = 1 -> dydx*dedy = dedy exactly
  return ((1.0f - y) * y) * e;
}

template <bool kAMN, bool kBMN, int kEpi, int BLOCK_N, int kStages, int CM, int CN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
bp_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmParams p) {
  constexpr int BLOCK_M = GEMM_BLOCK_M, BLOCK_K = GEMM_BLOCK_K;
  constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 4;
  constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 4;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t CHUNK_BYTES = 32 * BLOCK_K * 4;  // one MN-major 32-wide column chunk: BLOCK_K rows x 128 B
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
  constexpr int CSIZE = CM * CN;
  static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");
  static_assert(TMEM_COLS <= 512, "TMEM");
  static_assert(CSIZE >= 1 && CSIZE <= 8 && (BLOCK_M / 32) % CN == 0 && (BLOCK_N / 32) % CM == 0, "cluster shape");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(kStages) * STAGE_BYTES);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // Tiles are scheduled per cluster: cluster-tile ct covers m-tiles [mc*CM, mc*CM+CM) x n-tiles [nc*CN, nc*CN+CN);
  // CTA rank r of the cluster owns (rm, rn) = (r % CM, r / CM) of it.  Tiles past the edge are phantoms: they load
  // (TMA zero-fills), multiply and skip their stores, so every CTA of a cluster runs the same barrier protocol.
  const int num_mc = ((p.M + BLOCK_M - 1) / BLOCK_M + CM - 1) / CM;
  const int num_nc = ((p.N + BLOCK_N - 1) / BLOCK_N + CN - 1) / CN;
  const int num_ctiles = num_mc * num_nc;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int crank = CSIZE > 1 ? (int)cluster_ctarank() : 0;
  const int rm = crank % CM, rn = crank / CM;
  const int cid = blockIdx.x / CSIZE, num_clusters = gridDim.x / CSIZE;
  uint16_t mask_row = 0, mask_col = 0;  // CTAs sharing my A tile (same rm) / my B tile (same rn)
#pragma unroll
  for (int j = 0; j < CN; ++j) mask_row |= uint16_t(1u << (rm + CM * j));
#pragma unroll
  for (int i = 0; i < CM; ++i) mask_col |= uint16_t(1u << (rn * CM + i));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], CM + CN - 1);  // one commit arrival from every CTA of my row and column
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CSIZE > 1) cluster_sync_all();  // peers' barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int ct = cid; ct < num_ctiles; ct += num_clusters) {
        const int m0 = ((ct % num_mc) * CM + rm) * BLOCK_M;
        const int n0 = ((ct / num_mc) * CN + rn) * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[s], ph ^ 1u);
          if (p.dbg_flags & 2u) {  // measurement aid: no loads, the stage is "full" at once
            mbar_arrive(&full[s]);
            if (++s == kStages) { s = 0; ph ^= 1u; }
            continue;
          }
          mbar_expect_tx(&full[s], STAGE_BYTES);  // my A + my B, whoever delivers the chunks
          uint8_t* sa = smem + size_t(s) * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int k0 = kb * BLOCK_K;
          // Every operand tile is 32-wide chunks of 4 KB (K-major: 32 rows x 128 B; MN-major: 32 k-rows x 128 B).
          // Chunk c of my A tile is fetched by the row CTA with rn == c % CN and multicast to the row; same for B.
#pragma unroll
          for (int c = 0; c < BLOCK_M / 32; ++c) {
            if (c % CN != rn) continue;
            const int ci = kAMN ? m0 + 32 * c : k0, co = kAMN ? k0 : m0 + 32 * c;
            if constexpr (CN > 1) tma_load_2d_mc(sa + c * CHUNK_BYTES, &tmA, &full[s], ci, co, mask_row);
            else tma_load_2d(sa + c * CHUNK_BYTES, &tmA, &full[s], ci, co);
          }
#pragma unroll
          for (int c = 0; c < BLOCK_N / 32; ++c) {
            if (c % CM != rm) continue;
            const int ci = kBMN ? n0 + 32 * c : k0, co = kBMN ? k0 : n0 + 32 * c;
            if constexpr (CM > 1) tma_load_2d_mc(sb + c * CHUNK_BYTES, &tmB, &full[s], ci, co, mask_col);
            else tma_load_2d(sb + c * CHUNK_BYTES, &tmB, &full[s], ci, co);
          }
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BLOCK_M, BLOCK_N, kAMN ? 1u : 0u, kBMN ? 1u : 0u);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int ct = cid; ct < num_ctiles; ct += num_clusters) {
        mbar_wait(&tempty[as], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + size_t(s) * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const uint32_t mn_lbo = p.dbg_mn_lbo ? p.dbg_mn_lbo : CHUNK_BYTES;
          const uint32_t mn_sbo = p.dbg_mn_sbo ? p.dbg_mn_sbo : 512u;
          if ((p.dbg_flags & 1u) && CSIZE == 1) {  // measurement aid: consume the stage without multiplying
            mbar_arrive(&empty[s]);
            if (++s == kStages) { s = 0; ph ^= 1u; }
            continue;
          }
#pragma unroll
          for (int k = 0; k < BLOCK_K / 8; ++k) {
            // K-major (SWIZZLE_128B): rows of 128 B, 8-row groups 1024 B apart (SBO); one k-step = 8 fp32 = 32 B
            //   inside the swizzled row.
            // MN-major (SWIZZLE_128B_BASE32B): 32-wide column chunks CHUNK_BYTES apart (LBO); k-rows of 128 B,
            //   4-row swizzle groups 512 B apart (SBO); one k-step = 8 k-rows = 1024 B.
            const uint64_t adesc = kAMN ? make_smem_desc(sa + k * 1024, mn_lbo, mn_sbo, kLayoutSW128Base32)
                                        : make_smem_desc(sa + k * 32, 16, 1024, kLayoutSW128);
            const uint64_t bdesc = kBMN ? make_smem_desc(sb + k * 1024, mn_lbo, mn_sbo, kLayoutSW128Base32)
                                        : make_smem_desc(sb + k * 32, 16, 1024, kLayoutSW128);
            umma_tf32(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // release slot s in every CTA that fills it for me or that I fill (my row and column)
          if constexpr (CSIZE > 1) umma_commit_mc(&empty[s], uint16_t(mask_row | mask_col));
          else umma_commit(&empty[s]);
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
        umma_commit(&tfull[as]);
        as ^= 1;
        if (as == 0) aph ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    int as = 0;
    uint32_t aph = 0;
    float sq_local = 0.0f;
    for (int ct = cid; ct < num_ctiles; ct += num_clusters) {
      const int m0 = ((ct % num_mc) * CM + rm) * BLOCK_M;
      const int n0 = ((ct / num_mc) * CN + rn) * BLOCK_N;
      // One lane per warp polls (128 threads hammering try_wait on one mbarrier saturate the SM's barrier unit and
      // slow the producer/MMA hand-offs on the critical path — measured: ~450 ns per k-block regardless of work).
      if (lane == 0) mbar_wait_backoff(&tfull[as], aph);
      __syncwarp();
      tc_fence_after();
      const int m = m0 + q * 32 + lane;
      const bool m_ok = m < p.M;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BLOCK_N);
      float bias = 0.0f;
      if constexpr (kEpi == EPI_FWD_HID || kEpi == EPI_FWD_OUT) {
        if (m_ok) bias = __ldg(p.bias + m);
      }
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        const int nc = n0 + c * 32;
        if (nc >= p.N) break;
        uint32_t v[32];
        tmem_ld32(taddr + uint32_t(c * 32), v);
        tmem_ld_wait();
        if constexpr (kEpi == EPI_PLAIN) {
          if (m_ok) {
            float* o = p.out + size_t(nc) * p.ldo + m;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nc + j < p.N) o[size_t(j) * p.ldo] = __uint_as_float(v[j]);
          }
        } else if constexpr (kEpi == EPI_FWD_HID) {
          if (m_ok) {
            float* o = p.out + size_t(nc) * p.ldo + m;
            const bool drop = p.drop_p > 0.0f;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              float u4[4] = {1.0f, 1.0f, 1.0f, 1.0f};
              if (drop)
                philox_uniform4(p.seed_lo, p.seed_hi, uint32_t(p.frame0 + nc + j4 * 4) >> 2, uint32_t(m), p.layer,
                                p.step, u4);
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int j = j4 * 4 + jj;
                float y = act_fwd(fmaf(p.scale, __uint_as_float(v[j]), bias), p.act);
                if (drop && u4[jj] < p.drop_p) y = 0.0f;
                if (nc + j < p.N) o[size_t(j) * p.ldo] = y;
              }
            }
          }
        } else if constexpr (kEpi == EPI_FWD_OUT) {
          if (m_ok) {
            float tg[32];
            if (p.aux != nullptr) {
              const float* a = p.aux + size_t(nc) * p.ldaux + m;
#pragma unroll
              for (int j = 0; j < 32; ++j) tg[j] = (nc + j < p.N) ? __ldg(a + size_t(j) * p.ldaux) : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (nc + j < p.N) {
                const float o = fmaf(p.scale, __uint_as_float(v[j]), bias);
                if (p.out2 != nullptr) p.out2[size_t(nc + j) * p.ldo2 + m] = o;
                if (p.aux != nullptr) {
                  const float diff = o - tg[j];
                  if (p.out != nullptr) p.out[size_t(nc + j) * p.ldo + m] = p.gscale * diff;
                  sq_local = fmaf(diff, diff, sq_local);
                }
              }
            }
          }
        } else if constexpr (kEpi == EPI_DX) {
          if (m_ok) {
            float yv[32];
            const float* a = p.aux + size_t(nc) * p.ldaux + m;
#pragma unroll
            for (int j = 0; j < 32; ++j) yv[j] = (nc + j < p.N) ? __ldg(a + size_t(j) * p.ldaux) : 0.0f;
            float* o = p.out + size_t(nc) * p.ldo + m;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nc + j < p.N) o[size_t(j) * p.ldo] = act_bwd(yv[j], __uint_as_float(v[j]), p.act);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
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
  if constexpr (CSIZE > 1) cluster_sync_all();  // nobody leaves while peers may still signal / multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace bp
