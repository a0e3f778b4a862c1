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
// Operand fetch.  Every operand is a row-major fp32 matrix [R x ld] with ld % 32 == 0, viewed by ONE 3-D tensor map
// {32 floats, R rows (stride ld), ld/32 chunks (stride 128 B)}:
//   K-major  (reduction dim contiguous): rows = M/N index, chunks = k/32.   box {32, 128|BLOCK_N rows, 2 chunks}
//            -> smem [k-chunk][row][128 B], SWIZZLE_128B;            UMMA desc: SBO 1024, k-step 8 = +32 B in the row.
//   MN-major (M/N dim contiguous):       rows = k index,   chunks = mn/32.  box {32, 64 k-rows, 4|BLOCK_N/32 chunks}
//            -> smem [mn-chunk][k-row][128 B], SWIZZLE_128B_ATOM_32B; UMMA desc layout 1: LBO = 8192 (chunk), SBO 512
//               (4-row swizzle group), k-step 8 = +1024 B.  (tcgen05 accepts no other layout for MN-major 32-bit.)
// so a 64-deep k-block of a tile is ONE cp.async.bulk.tensor.3d per operand (32-64 KB).  Out-of-range reduction
// indices are zero-filled by TMA on at least one operand of every product (rows are exact extents; pad columns up
// to ld are kept zero); out-of-range M/N indices only reach accumulator rows/columns that are never stored.
//
// Pipeline (all role loops are WARP-UNIFORM, the asynchronous instructions are issued under elect.sync — issuing
// them from a divergent `if (lane == 0)` costs ~100 cycles per tcgen05.mma / TMA because every operand is first moved
// from vector to uniform registers — while every mbarrier wait is polled by ONE lane (not the issuing lane, see
// kPollLane) followed by __syncwarp; all measured with the clock64 trace below): warp 0 = TMA producer, warp 1 =
// tcgen05.mma issuer + TMEM owner, warps 2..5 = epilogue (tcgen05.ld -> registers -> fused math -> coalesced global
// stores).  kStages-deep smem ring of 64-deep k-blocks (full/empty mbarriers), double-buffered TMEM accumulator
// (tfull/tempty mbarriers), persistent static tile loop, every wait bounded (a protocol bug traps, never hangs).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "bp_gemm_params.h"
#include "bp_ptx.cuh"
#include "bp_rng.cuh"

namespace bp {

template <int BLOCK_N>
__host__ __device__ constexpr int gemm_stages() { return BLOCK_N == 128 ? 3 : 2; }

template <int BLOCK_N>
constexpr size_t gemm_smem_bytes() {
  return size_t(gemm_stages<BLOCK_N>()) * (GEMM_BLOCK_M + BLOCK_N) * GEMM_BLOCK_K * 4 + 1024 /*align slack*/ +
         256 /*barriers*/;
}

// x - trunc_tf32(x): exactly representable (<= 13 significant bits); the tensor core truncates the same way.
__device__ __forceinline__ float tf32_lo(float x) {
  return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}
// Output stores of the epilogues: explicitly to the GLOBAL space and invisible to the compiler's alias analysis.  When
// the parameter block does not come from the constant bank (bp_chain.cuh keeps a table of them in shared memory) its
// pointers are generic to the compiler: a plain `*o = y` becomes a generic ST that may alias the table itself, and every
// field (p.ldo, p.out, p.N ...) is re-read from shared memory after every store — 3x slower epilogues (measured).  No
// epilogue reads back what it stores, so no "memory" clobber is needed.
__device__ __forceinline__ void st_global_f32(float* p, float v) {
  asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v));
}
__device__ __forceinline__ float act_fwd(float x, int act) {
  if (act == 0) return x > 0.0f ? x : 0.0f;
  return 1.0f / (1.0f + expf(-x));
}
__device__ __forceinline__ float act_bwd(float y, float e, int act) {
  if (act == 0) return y > 0.0f ? e : 0.0f;       // dydx = 1 -> dydx*dedy = dedy exactly
  return ((1.0f - y) * y) * e;
}

// Hidden-layer forward epilogue of one 32-column chunk with every loop-invariant switch a template parameter:
// ACT 0 ReLU / 1 sigmoid; FL bit 0 dropout, bit 1 low part wanted (3xTF32), bit 2 ragged chunk (some columns >= N).
// All 32 values are formed first, then stored back to back (32 independent coalesced 128-byte warp stores in flight).
// Same operations in the same order as before the split: y = act(fma(scale, acc, bias)); dropped -> 0.
template <int ACT, int FL>
__device__ __forceinline__ void epi_fwd_hid_chunk(const GemmParams& p, const uint32_t (&v)[32], int m, int nc,
                                                      float bias) {
  constexpr bool kDrop = (FL & 1) != 0, kLo = (FL & 2) != 0, kRagged = (FL & 4) != 0;
  float y[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float x = fmaf(p.scale, __uint_as_float(v[j]), bias);
    if constexpr (ACT == 0) y[j] = x > 0.0f ? x : 0.0f;
    else y[j] = 1.0f / (1.0f + expf(-x));
  }
  if constexpr (kDrop) {
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      float u4[4];
      philox_uniform4(p.seed_lo, p.seed_hi, uint32_t(p.frame0 + nc + j4 * 4) >> 2, uint32_t(m), p.layer, p.step, u4);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        if (u4[jj] < p.drop_p) y[j4 * 4 + jj] = 0.0f;
    }
  }
  float* o = p.out + size_t(nc) * p.ldo + m;
  const int valid = kRagged ? p.N - nc : 32;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (!kRagged || j < valid) st_global_f32(o + size_t(j) * p.ldo, y[j]);
  if constexpr (kLo) {
    float* ol = p.out_lo + size_t(nc) * p.ldo + m;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (!kRagged || j < valid) st_global_f32(ol + size_t(j) * p.ldo, tf32_lo(y[j]));
  }
}

// Output-layer epilogue of one chunk (kernSubClean, DevFunc.cu:253-268, + the squared-error monitor), switches compiled
// in: kAux targets present (tg = targ[nc..nc+32)[m], loaded by the caller), kOut store (2/B)(o - targ), kOut2 store o, kLo low part of kOut, kRagged columns >= N exist.
// Returns sq + this chunk's sum of (o - targ)^2, accumulated in column order with one fma per term as before.
template <bool kAux, bool kOut, bool kOut2, bool kLo, bool kRagged>
__device__ __forceinline__ float epi_fwd_out_chunk(const GemmParams& p, const uint32_t (&v)[32], const float (&tg)[32],
                                                   int m, int nc, float bias, float sq) {
  const int valid = kRagged ? p.N - nc : 32;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (!kRagged || j < valid) {
      const float o = fmaf(p.scale, __uint_as_float(v[j]), bias);
      if constexpr (kOut2) st_global_f32(p.out2 + size_t(nc + j) * p.ldo2 + m, o);
      if constexpr (kAux) {
        const float diff = o - tg[j];
        if constexpr (kOut) {
          const float dv = p.gscale * diff;
          st_global_f32(p.out + size_t(nc + j) * p.ldo + m, dv);
          if constexpr (kLo) st_global_f32(p.out_lo + size_t(nc + j) * p.ldo + m, tf32_lo(dv));
        }
        sq = fmaf(diff, diff, sq);
      }
    }
  }
  return sq;
}

// dX epilogue of one chunk, same treatment: out = act'(y) * acc (kernDsigmoid + kernVecMul, DevFunc.cu:81-97, 244-250).
template <int ACT, bool kLo, bool kRagged>
__device__ __forceinline__ void epi_dx_chunk(const GemmParams& p, const uint32_t (&v)[32], const float (&yv)[32], int m,
                                             int nc) {
  float d[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float e = __uint_as_float(v[j]);
    if constexpr (ACT == 0) d[j] = yv[j] > 0.0f ? e : 0.0f;
    else d[j] = ((1.0f - yv[j]) * yv[j]) * e;
  }
  float* o = p.out + size_t(nc) * p.ldo + m;
  const int valid = kRagged ? p.N - nc : 32;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (!kRagged || j < valid) st_global_f32(o + size_t(j) * p.ldo, d[j]);
  if constexpr (kLo) {
    float* ol = p.out_lo + size_t(nc) * p.ldo + m;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (!kRagged || j < valid) st_global_f32(ol + size_t(j) * p.ldo, tf32_lo(d[j]));
  }
}

// Fused epilogue math for one chunk of 32 accumulator columns [nc, nc+32) of TMEM lane (= output row) m.
template <int kEpi>
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmParams& p, uint32_t (&v)[32], int m, bool m_ok, int nc,
                                                    float bias, float& sq_local, size_t plane_off = 0) {
  const bool whole = nc + 32 <= p.N;  // warp-uniform: all 32 columns of the chunk are real
  if constexpr (kEpi == EPI_PLAIN) {
    if (m_ok) {
      float* base = p.out + plane_off;
      if (p.scatter_n > 0) base = p.scatter[(p.chunk_base + (nc >> 5)) % p.scatter_n];
      float* o = base + size_t(nc) * p.ldo + m;
#pragma unroll
      for (int j = 0; j < 32; ++j, o += p.ldo)
        if (whole || nc + j < p.N) st_global_f32(o, __uint_as_float(v[j]));
    }
  } else if constexpr (kEpi == EPI_FWD_HID) {
    if (m_ok) {
      // warp-uniform switches, hoisted out of the 32-column loop: with them inside, every element carried ~6 branches
      // (activation kind, dropout, ragged tail, low part) and the hidden-layer forward product took 26 us where the
      // plain-epilogue product of the same shape takes 17 (profiles/r2b)
      const int flags = (p.drop_p > 0.0f ? 1 : 0) | (p.out_lo != nullptr ? 2 : 0) | (whole ? 0 : 4);
#define BP_HID(ACT, FL) epi_fwd_hid_chunk<ACT, FL>(p, v, m, nc, bias)
      if (p.act == 0) {
        switch (flags) {
          case 0: BP_HID(0, 0); break;
          case 1: BP_HID(0, 1); break;
          case 2: BP_HID(0, 2); break;
          case 3: BP_HID(0, 3); break;
          case 4: BP_HID(0, 4); break;
          case 5: BP_HID(0, 5); break;
          case 6: BP_HID(0, 6); break;
          default: BP_HID(0, 7); break;
        }
      } else {
        switch (flags) {
          case 0: BP_HID(1, 0); break;
          case 1: BP_HID(1, 1); break;
          case 2: BP_HID(1, 2); break;
          case 3: BP_HID(1, 3); break;
          case 4: BP_HID(1, 4); break;
          case 5: BP_HID(1, 5); break;
          case 6: BP_HID(1, 6); break;
          default: BP_HID(1, 7); break;
        }
      }
#undef BP_HID
    }
  } else if constexpr (kEpi == EPI_FWD_OUT) {
    if (m_ok) {
      // the combinations the runtime issues, each with its switches compiled in (see epi_fwd_out_chunk); anything else
      // takes the generic loop
      const bool aux = p.aux != nullptr, out = p.out != nullptr, out2 = p.out2 != nullptr, lo = p.out_lo != nullptr;
      float tg[32];
      if (aux) {
        const float* a = p.aux + size_t(nc) * p.ldaux + m;
#pragma unroll
        for (int j = 0; j < 32; ++j) tg[j] = (whole || nc + j < p.N) ? __ldg(a + size_t(j) * p.ldaux) : 0.0f;
      }
      if (aux && out && !out2) {                       // training: D_L = (2/B)(o - targ), loss monitor
        if (whole) { if (lo) sq_local = epi_fwd_out_chunk<true, true, false, true, false>(p, v, tg, m, nc, bias, sq_local);
                     else sq_local = epi_fwd_out_chunk<true, true, false, false, false>(p, v, tg, m, nc, bias, sq_local); }
        else       { if (lo) sq_local = epi_fwd_out_chunk<true, true, false, true, true>(p, v, tg, m, nc, bias, sq_local);
                     else sq_local = epi_fwd_out_chunk<true, true, false, false, true>(p, v, tg, m, nc, bias, sq_local); }
      } else if (!aux && !out && out2) {               // decode: raw linear output
        if (whole) epi_fwd_out_chunk<false, false, true, false, false>(p, v, tg, m, nc, bias, 0.0f);
        else epi_fwd_out_chunk<false, false, true, false, true>(p, v, tg, m, nc, bias, 0.0f);
      } else if (aux && !out) {                        // cross-validation: score, optionally the output too
        if (whole) { if (out2) sq_local = epi_fwd_out_chunk<true, false, true, false, false>(p, v, tg, m, nc, bias, sq_local);
                     else sq_local = epi_fwd_out_chunk<true, false, false, false, false>(p, v, tg, m, nc, bias, sq_local); }
        else       { if (out2) sq_local = epi_fwd_out_chunk<true, false, true, false, true>(p, v, tg, m, nc, bias, sq_local);
                     else sq_local = epi_fwd_out_chunk<true, false, false, false, true>(p, v, tg, m, nc, bias, sq_local); }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (whole || nc + j < p.N) {
            const float o = fmaf(p.scale, __uint_as_float(v[j]), bias);
            if (out2) p.out2[size_t(nc + j) * p.ldo2 + m] = o;
            if (aux) {
              const float diff = o - tg[j];
              if (out) {
                const float dv = p.gscale * diff;
                p.out[size_t(nc + j) * p.ldo + m] = dv;
                if (lo) p.out_lo[size_t(nc + j) * p.ldo + m] = tf32_lo(dv);
              }
              sq_local = fmaf(diff, diff, sq_local);
            }
          }
        }
      }
    }
  }
}

// EPI_DX with the Y operand (aux) software-pipelined: kernDsigmoid + kernVecMul (DevFunc.cu:81-97, 244-250) need
// Y[frame][unit] for every accumulator element; loading it only after the accumulator is complete put ~2.5 us of
// exposed global-load latency per 32-column chunk at the end of every dX GEMM (isolated 2048x1024x2048: dX 28.4 us vs
// 18.4 us for the forward product of the same size, profiles/r1d).  The epilogue warps are idle during the main loop, so
// they fetch the first two chunks then, and chunk c+2 while chunk c+1 is processed.  (Four chunks ahead — a whole
// 128-wide tile, 206 registers — is faster still with a warm L2, 24.7 us, but slower inside the bunch, where Y comes
// from HBM: dX class 0.101 vs 0.082 ms, ncu 43.5 vs 34.1 us per launch; two it is.)
struct DxPrefetch {
  float y[2][32];
  __device__ __forceinline__ void load(const GemmParams& p, int slot, int m, bool m_ok, int nc) {
    if (!m_ok || nc >= p.N) return;
    const float* a = p.aux + size_t(nc) * p.ldaux + m;
    const bool whole = nc + 32 <= p.N;
#pragma unroll
    for (int j = 0; j < 32; ++j, a += p.ldaux) y[slot][j] = (whole || nc + j < p.N) ? __ldg(a) : 0.0f;
  }
  __device__ __forceinline__ void start(const GemmParams& p, int m, bool m_ok, int n0) {
    load(p, 0, m, m_ok, n0);
    load(p, 1, m, m_ok, n0 + 32);
  }
};

__device__ __forceinline__ void gemm_dx_store(const GemmParams& p, const uint32_t (&v)[32], const float (&yv)[32], int m,
                                              bool m_ok, int nc) {
  if (!m_ok) return;
  const bool whole = nc + 32 <= p.N;      // warp-uniform, like act and out_lo: dispatched once per chunk
  const bool lo = p.out_lo != nullptr;
  if (p.act == 0) {
    if (whole) { if (lo) epi_dx_chunk<0, true, false>(p, v, yv, m, nc); else epi_dx_chunk<0, false, false>(p, v, yv, m, nc); }
    else       { if (lo) epi_dx_chunk<0, true, true>(p, v, yv, m, nc);  else epi_dx_chunk<0, false, true>(p, v, yv, m, nc); }
  } else {
    if (whole) { if (lo) epi_dx_chunk<1, true, false>(p, v, yv, m, nc); else epi_dx_chunk<1, false, false>(p, v, yv, m, nc); }
    else       { if (lo) epi_dx_chunk<1, true, true>(p, v, yv, m, nc);  else epi_dx_chunk<1, false, true>(p, v, yv, m, nc); }
  }
}

template <int BLOCK_N>
__device__ __forceinline__ void gemm_dx_epilogue(const GemmParams& p, DxPrefetch& pre, uint32_t taddr, int m, bool m_ok,
                                                 int n0) {
#pragma unroll 1
  for (int c = 0; c < BLOCK_N / 32; c += 2) {
    const int nc = n0 + c * 32;
    if (nc >= p.N) break;
    uint32_t v[32];
    tmem_ld32(taddr + uint32_t(c * 32), v);
    tmem_ld_wait();
    gemm_dx_store(p, v, pre.y[0], m, m_ok, nc);
    if (c + 2 < BLOCK_N / 32) pre.load(p, 0, m, m_ok, nc + 64);
    if (nc + 32 >= p.N) break;
    tmem_ld32(taddr + uint32_t((c + 1) * 32), v);
    tmem_ld_wait();
    gemm_dx_store(p, v, pre.y[1], m, m_ok, nc + 32);
    if (c + 3 < BLOCK_N / 32) pre.load(p, 1, m, m_ok, nc + 96);
  }
}

// Output-layer epilogue for training (targets present, D_L stored, no raw output) with the targets software-pipelined
// like EPI_DX's Y operand: the first two chunks are in registers before the accumulator is complete.
template <int BLOCK_N>
__device__ __forceinline__ float gemm_fwdout_train_epilogue(const GemmParams& p, DxPrefetch& pre, uint32_t taddr, int m,
                                                            bool m_ok, int n0, float bias, float sq) {
  auto one = [&](const uint32_t (&v)[32], const float (&tg)[32], int nc) {
    if (!m_ok) return;
    const bool whole = nc + 32 <= p.N, lo = p.out_lo != nullptr;
    if (whole) { if (lo) sq = epi_fwd_out_chunk<true, true, false, true, false>(p, v, tg, m, nc, bias, sq);
                 else sq = epi_fwd_out_chunk<true, true, false, false, false>(p, v, tg, m, nc, bias, sq); }
    else       { if (lo) sq = epi_fwd_out_chunk<true, true, false, true, true>(p, v, tg, m, nc, bias, sq);
                 else sq = epi_fwd_out_chunk<true, true, false, false, true>(p, v, tg, m, nc, bias, sq); }
  };
#pragma unroll 1
  for (int c = 0; c < BLOCK_N / 32; c += 2) {
    const int nc = n0 + c * 32;
    if (nc >= p.N) break;
    uint32_t v[32];
    tmem_ld32(taddr + uint32_t(c * 32), v);
    tmem_ld_wait();
    one(v, pre.y[0], nc);
    if (c + 2 < BLOCK_N / 32) pre.load(p, 0, m, m_ok, nc + 64);
    if (nc + 32 >= p.N) break;
    tmem_ld32(taddr + uint32_t((c + 1) * 32), v);
    tmem_ld_wait();
    one(v, pre.y[1], nc + 32);
    if (c + 3 < BLOCK_N / 32) pre.load(p, 1, m, m_ok, nc + 96);
  }
  return sq;
}

template <bool kAMN, bool kBMN, int kEpi, int BLOCK_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
bp_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo,
               const GemmParams p) {
  constexpr int BLOCK_M = GEMM_BLOCK_M, BLOCK_K = GEMM_BLOCK_K;
  constexpr int kStages = gemm_stages<BLOCK_N>();
  constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 4;
  constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 4;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
  static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");
  static_assert(TMEM_COLS <= 512, "TMEM");
  // UMMA shared-memory descriptors: constant high word, low word = start address >> 4 | LBO field.
  //   hi: [0,14) SBO>>4, [14,16) version 1, [29,32) layout;  lo: [0,14) addr>>4, [16,30) LBO>>4
  constexpr uint32_t kDescHiK = (1024u >> 4) | (1u << 14) | (kLayoutSW128 << 29);
  constexpr uint32_t kDescHiMN = (512u >> 4) | (1u << 14) | (kLayoutSW128Base32 << 29);
  constexpr uint32_t kDescLoK = (16u >> 4) << 16;
  constexpr uint32_t kDescLoMN = ((uint32_t(BLOCK_K) * 128u) >> 4) << 16;  // LBO = one mn-chunk = BLOCK_K rows x 128 B
  constexpr uint32_t kAsub = BLOCK_M * 128;  // K-major: bytes between the two 32-wide k-chunks of a stage
  constexpr uint32_t kBsub = BLOCK_N * 128;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;          // shared-space address of stage 0
  uint8_t* smem = smem_raw + (smem_base - raw_addr);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(kStages) * STAGE_BYTES);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // The lane that polls mbarriers must NOT be the lane elect.sync picks (lane 0) to issue tcgen05.commit: a thread's
  // barrier-unit operations are processed in order, and a try_wait queued behind its own pending commit-arrive only
  // returns when the committed MMAs have COMPLETED — which serialises issue with execution (measured: ~480 idle
  // tensor-pipe cycles per k-block).
  constexpr int kPollLane = 1;

  const int num_m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int num_n_tiles = (p.N - p.n_begin + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int num_passes = p.passes == 3 ? 3 : 1;
  // split-K: work item t = (tile, k-slice); slice s covers k-blocks [s*kb_per, min(num_kb, (s+1)*kb_per))
  const int k_splits = p.k_splits > 1 ? p.k_splits : 1;
  const int kb_per = (num_kb + k_splits - 1) / k_splits;
  const int num_items = num_tiles * k_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
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
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap
  // the tail of the previous kernel in the stream; nothing below may touch global memory before that kernel's
  // writes are visible.  Dependents may in turn start their own prologue as soon as our CTAs retire.
  griddep_wait();
  griddep_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform loop)
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < num_items; t += gridDim.x) {
      const int tile = t % num_tiles;
      const int m0 = (tile % num_m_tiles) * BLOCK_M;
      const int n0 = p.n_begin + (tile / num_m_tiles) * BLOCK_N;
      const int kb0 = (t / num_tiles) * kb_per;
      const int kb1 = min(num_kb, kb0 + kb_per);
      const int num_it = (kb1 - kb0) * num_passes;  // pipeline iterations of this work item
      for (int it = 0, kb = kb0, pass = 0; it < num_it; ++it, ++kb) {
        if (kb == kb1) { kb = kb0; ++pass; }
        const CUtensorMap* mapA = pass == 1 ? &tmAlo : &tmA;   // pass 0: A*B   pass 1: A_lo*B   pass 2: A*B_lo
        const CUtensorMap* mapB = pass == 2 ? &tmBlo : &tmB;
        if (lane == kPollLane) mbar_wait(&empty[s], ph ^ 1u);  // ONE lane polls (see header)
        __syncwarp();
        if (elect_one()) {
          uint8_t* sa = smem + size_t(s) * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full[s], STAGE_BYTES);
          const int a1 = kAMN ? kb * BLOCK_K : m0, a2 = kAMN ? m0 / 32 : kb * (BLOCK_K / 32);
          const int b1 = kBMN ? kb * BLOCK_K : n0, b2 = kBMN ? n0 / 32 : kb * (BLOCK_K / 32);
          if (p.hint_a) tma_load_3d_hint(sa, mapA, &full[s], 0, a1, a2, p.hint_a);
          else tma_load_3d(sa, mapA, &full[s], 0, a1, a2);
          if (p.hint_b) tma_load_3d_hint(sb, mapB, &full[s], 0, b1, b2, p.hint_b);
          else tma_load_3d(sb, mapB, &full[s], 0, b1, b2);
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop)
    constexpr uint32_t idesc = make_idesc_tf32(BLOCK_M, BLOCK_N, kAMN ? 1u : 0u, kBMN ? 1u : 0u);
    int s = 0;
    uint32_t ph = 0;
    int as = 0;
    uint32_t aph = 0;
    for (int t = blockIdx.x; t < num_items; t += gridDim.x) {
      const int kb0 = (t / num_tiles) * kb_per;
      const int num_it = (min(num_kb, kb0 + kb_per) - kb0) * num_passes;
      if (lane == kPollLane) mbar_wait(&tempty[as], aph ^ 1u);
      __syncwarp();
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + uint32_t(as * BLOCK_N);
      for (int kb = 0; kb < num_it; ++kb) {  // kb counts pipeline iterations (k-blocks x passes)
        if (lane == kPollLane) mbar_wait(&full[s], ph);
        __syncwarp();
        tc_fence_after();
        const uint32_t sa = smem_base + uint32_t(s) * STAGE_BYTES;
        const uint32_t sb = sa + A_BYTES;
        if (elect_one()) {
          {
            // descriptor low words: one base per operand per stage, then base + constant per k-step (the tensor pipe's
            // instruction queue is shallow, so every cycle of issue overhead beyond ~250 per k-block idles the pipe)
            const uint32_t a_lo = (kAMN ? kDescLoMN : kDescLoK) + (sa >> 4);
            const uint32_t b_lo = (kBMN ? kDescLoMN : kDescLoK) + (sb >> 4);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 8; ++k) {
              const uint32_t a_off = kAMN ? uint32_t(k) * 1024u : uint32_t(k / 4) * kAsub + uint32_t(k % 4) * 32u;
              const uint32_t b_off = kBMN ? uint32_t(k) * 1024u : uint32_t(k / 4) * kBsub + uint32_t(k % 4) * 32u;
              umma_tf32_lohi(d_tmem, a_lo + (a_off >> 4), kAMN ? kDescHiMN : kDescHiK, b_lo + (b_off >> 4),
                             kBMN ? kDescHiMN : kDescHiK, idesc, (k != 0 || kb != 0) ? 1u : 0u);
            }
            umma_commit(&empty[s]);  // slot s is free again once these MMAs have read it
          }
          if (kb == num_it - 1) umma_commit(&tfull[as]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
      as ^= 1;
      if (as == 0) aph ^= 1u;
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    int as = 0;
    uint32_t aph = 0;
    float sq_local = 0.0f;
    for (int t = blockIdx.x; t < num_items; t += gridDim.x) {
      const int tile = t % num_tiles;
      const int m0 = (tile % num_m_tiles) * BLOCK_M;
      const int n0 = p.n_begin + (tile / num_m_tiles) * BLOCK_N;
      const size_t plane_off = size_t(t / num_tiles) * size_t(p.split_stride);
      const int m = m0 + q * 32 + lane;
      const bool m_ok = m < p.M;
      DxPrefetch pre;
      if constexpr (kEpi == EPI_DX) pre.start(p, m, m_ok, n0);  // under the main loop (see gemm_dx_epilogue)
      if (lane == 0) mbar_wait_backoff(&tfull[as], aph);  // one poller per warp, long suspend hint
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BLOCK_N);
      float bias = 0.0f;
      if constexpr (kEpi == EPI_FWD_HID || kEpi == EPI_FWD_OUT) {
        if (m_ok) bias = __ldg(p.bias + m);
      }
      if constexpr (kEpi == EPI_DX) {
        gemm_dx_epilogue<BLOCK_N>(p, pre, taddr, m, m_ok, n0);
      } else {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          const int nc = n0 + c * 32;
          if (nc >= p.N) break;
          uint32_t v[32];
          tmem_ld32(taddr + uint32_t(c * 32), v);
          tmem_ld_wait();
          gemm_epilogue_chunk<kEpi>(p, v, m, m_ok, nc, bias, sq_local, plane_off);
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// Finisher of a split-K output-layer product: adds the k_splits partial planes (in slice order, so the result does not
// depend on scheduling) and applies the EPI_FWD_OUT epilogue.  ws plane s holds element (m,n) at ws[s*stride+n*ldw+m].
// One thread per (n, m) with m contiguous: every access is coalesced; ~1.2 MB per plane at the C2 shape.
__global__ void __launch_bounds__(256)
bp_out_finish_kernel(const float* __restrict__ ws, long long ldw, const GemmParams p) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int n = static_cast<int>(idx / ldw);
  const int m = static_cast<int>(idx - static_cast<long long>(n) * ldw);
  float sq_local = 0.0f;
  if (n < p.N && m < p.M) {
    const float* w = ws + static_cast<size_t>(n) * ldw + m;
    float acc = w[0];
    for (int s = 1; s < p.k_splits; ++s) acc += w[static_cast<size_t>(s) * p.split_stride];
    const float o = fmaf(p.scale, acc, __ldg(p.bias + m));
    if (p.out2 != nullptr) p.out2[static_cast<size_t>(n) * p.ldo2 + m] = o;
    if (p.aux != nullptr) {
      const float diff = o - __ldg(p.aux + static_cast<size_t>(n) * p.ldaux + m);
      if (p.out != nullptr) {
        const float dv = p.gscale * diff;
        p.out[static_cast<size_t>(n) * p.ldo + m] = dv;
        if (p.out_lo != nullptr) p.out_lo[static_cast<size_t>(n) * p.ldo + m] = tf32_lo(dv);
      }
      sq_local = diff * diff;
    }
  }
  if (p.sqerr != nullptr) {
    double sq = static_cast<double>(sq_local);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
    __shared__ double warp_sum[8];
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int i = 0; i < 8; ++i) tot += warp_sum[i];
      if (tot != 0.0) atomicAdd(p.sqerr, tot);
    }
  }
}

}  // namespace bp
