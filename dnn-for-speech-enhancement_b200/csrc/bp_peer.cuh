// Data-parallel exchange over NVLink peer memory (sm_100a), replacing "dW GEMM -> ncclAllReduce -> replicated SGD" by
//
//   dW GEMM epilogue  --(P2P stores, tile by tile)-->  owner's receive slab      reduce-scatter fused into the GEMM
//   owner:  G = sum over ranks (fixed order), momentum-SGD on its rows only       1/N of the update traffic per GPU
//   owner   --(P2P stores)-->  every replica's weight arena                       all-gather fused into the update
//
// Ownership: rows of the parameter arena in chunks of 32 (one epilogue chunk of the dW GEMM), dealt round-robin:
// chunk c of the concatenated layers belongs to rank c % N.  Every rank's partial gradient of a chunk is written by
// its GEMM epilogue straight into the owner's slab `recv[src]`; nothing is staged locally and nothing is reduced twice.
// Synchronisation is two monotone counters per (rank, peer) in the receiver's memory — "src's gradients of step s have
// landed" and "owner's weights of step s have landed" — written with a system-scope release after a system fence by a
// one-warp kernel that follows the producing kernels in stream order, and polled (bounded) with system-scope acquires.
// Replicas stay bit-identical: every replica receives the very same weight bits from the single owner of each row.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bp {

constexpr int kMaxPeers = 8;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

struct PeerFlags {
  unsigned long long* slot[kMaxPeers];  // slot[p] = address, in rank p's memory, of the counter this rank writes
};

// Launched after the kernels whose peer stores it publishes (stream order): one thread per peer.
__global__ void bp_peer_signal_kernel(const __grid_constant__ PeerFlags f, int world, unsigned long long value) {
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(f.slot[threadIdx.x], value);
  }
}

// Bounded spin until every peer's counter (in OUR memory) has reached `value`.  A protocol error traps (the context
// dies with an error) instead of hanging the GPU.
__device__ __forceinline__ void peer_wait_all(const unsigned long long* mine, int world, unsigned long long value) {
  if (threadIdx.x < world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(mine + threadIdx.x) < value) {
      __nanosleep(200);
      if (clock64() - t0 > 120000000000LL) {  // ~60 s: ranks read and assemble their own chunks, so a peer may be
                                               // seconds late (cold page cache); only a dead one should trap
        printf("bp_peer: timeout waiting for rank %d to reach step %llu\n", (int)threadIdx.x, value);
        __trap();
      }
    }
  }
}

__global__ void bp_peer_wait_kernel(const unsigned long long* mine, int world, unsigned long long value) {
  peer_wait_all(mine, world, value);
}
// Layer geometry of the arena for the owner-side update.
struct PeerLayers {
  int n;
  long long begin4[10];   // first float4 of the layer's block
  long long end4[10];
  int row4[10];           // float4 per row (ldN / 4)
  int bias_row[10];       // = K (weight cost does not apply to it)
  int chunk_base[10];     // global index of the layer's first 32-row chunk
};

struct PeerArenas {
  float4* w[kMaxPeers];     // every replica's weight arena, as mapped into this device
  float4* w_lo[kMaxPeers];  // 3xTF32 low parts, or null
};

// Owner-side reduce + momentum SGD + all-gather.  Same arithmetic, operation for operation, as bp_sgd_kernel
// (kernUpdatedelta + kernAccSum, DevFunc.cu:313-318, 270-277) on G = recv[0] + recv[1] + ... in rank order.
template <bool kHasWC>
__global__ void __launch_bounds__(256)
bp_peer_sgd_kernel(float4* __restrict__ delta, const float4* __restrict__ recv, long long arena4,
                   const __grid_constant__ PeerArenas pa, const __grid_constant__ PeerLayers pl, int world, int rank,
                   float nf, float momentum, float one_minus_m_lr, float weightcost,
                   const unsigned long long* grad_flags, unsigned long long step) {
  peer_wait_all(grad_flags, world, step);  // every rank's partial gradients of this step have landed in `recv`
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < arena4; i += stride) {
    int l = 0;
#pragma unroll 1
    while (l < pl.n - 1 && i >= pl.end4[l]) ++l;
    const int row = static_cast<int>((i - pl.begin4[l]) / pl.row4[l]);
    if ((pl.chunk_base[l] + (row >> 5)) % world != rank) continue;  // warp-uniform except at chunk edges
    float4 g = __ldcs(recv + i);
    for (int s = 1; s < world; ++s) {
      const float4 t = __ldcs(recv + s * arena4 + i);
      g.x = __fadd_rn(g.x, t.x);
      g.y = __fadd_rn(g.y, t.y);
      g.z = __fadd_rn(g.z, t.z);
      g.w = __fadd_rn(g.w, t.w);
    }
    const float4 d = delta[i];
    const float4 x = pa.w[rank][i];
    const float wc = (kHasWC && row != pl.bias_row[l]) ? weightcost : 0.0f;
    const float gv[4] = {g.x, g.y, g.z, g.w};
    float dv[4] = {d.x, d.y, d.z, d.w};
    float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float t = __fdiv_rn(gv[k], nf);
      if (kHasWC) t = __fadd_rn(t, __fmul_rn(wc, xv[k]));
      const float nd = __fsub_rn(__fmul_rn(momentum, dv[k]), __fmul_rn(one_minus_m_lr, t));
      dv[k] = nd;
      xv[k] = __fadd_rn(nd, xv[k]);
    }
    delta[i] = make_float4(dv[0], dv[1], dv[2], dv[3]);
    const float4 xn = make_float4(xv[0], xv[1], xv[2], xv[3]);
#pragma unroll 1
    for (int p = 0; p < world; ++p) pa.w[p][i] = xn;
    if (pa.w_lo[rank] != nullptr) {
      float4 lo;
      lo.x = xv[0] - __uint_as_float(__float_as_uint(xv[0]) & 0xFFFFE000u);
      lo.y = xv[1] - __uint_as_float(__float_as_uint(xv[1]) & 0xFFFFE000u);
      lo.z = xv[2] - __uint_as_float(__float_as_uint(xv[2]) & 0xFFFFE000u);
      lo.w = xv[3] - __uint_as_float(__float_as_uint(xv[3]) & 0xFFFFE000u);
#pragma unroll 1
      for (int p = 0; p < world; ++p) pa.w_lo[p][i] = lo;
    }
  }
}

}  // namespace bp
