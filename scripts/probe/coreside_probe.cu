// Which property of a resident kernel prevents a kernel of ANOTHER stream from starting beside it?  (B200 bring-up probe)
// spinner: `grid` CTAs x 192 threads wait for a flag; release: 1 thread on another stream sets it.  Variants: dynamic
// shared memory size, cluster launch, large kernel parameters, TMEM allocation.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
struct Big { char pad[20000]; unsigned long long* flag; int tmem; };
__device__ __forceinline__ unsigned long long ldacq(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
template <bool kBig>
__global__ void __launch_bounds__(192, 1) spinner(const __grid_constant__ Big b, unsigned long long* flag_small, int tmem, int* out) {
  extern __shared__ char sm[];
  __shared__ unsigned tslot;
  unsigned long long* flag = kBig ? b.flag : flag_small;
  const int use_tmem = kBig ? b.tmem : tmem;
  if (use_tmem && threadIdx.x < 32) {
    unsigned a = (unsigned)__cvta_generic_to_shared(&tslot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(a) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    sm[0] = 1;
    long long t0 = clock64();
    int ok = 1;
    while (ldacq(flag) == 0) { __nanosleep(200); if (clock64() - t0 > 3000000000LL) { ok = 0; break; } }
    if (blockIdx.x == 0) *out = ok;
  }
  __syncthreads();
  if (use_tmem && threadIdx.x < 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tslot) : "memory");
  }
}
__global__ void release(unsigned long long* flag) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(1ull) : "memory");
}
int main(int argc, char** argv) {
  const int grid = atoi(argv[1]), smem_kb = atoi(argv[2]), cluster = atoi(argv[3]), big = atoi(argv[4]), tmem = atoi(argv[5]);
  unsigned long long* flag; int* out; int h = -1;
  cudaMalloc(&flag, 64); cudaMemset(flag, 0, 64); cudaMalloc(&out, 4); cudaMemset(out, 0xff, 4);
  cudaStream_t a, c;
  cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking);
  // CUDA loads kernels lazily, and loading one may synchronise with running kernels: launch `release` once BEFORE the
  // spinner exists (argv[6] = 1), as a program that relies on concurrency must
  if (argc > 6 && atoi(argv[6])) { release<<<1, 1, 0, c>>>(flag + 4); cudaDeviceSynchronize(); }
  static Big b; b.flag = flag; b.tmem = tmem;
  cudaFuncSetAttribute(spinner<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
  cudaFuncSetAttribute(spinner<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem_kb * 1024; cfg.stream = a;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = cluster > 1 ? 1 : 0;
  cudaError_t e = big ? cudaLaunchKernelEx(&cfg, spinner<true>, b, flag, tmem, out) : cudaLaunchKernelEx(&cfg, spinner<false>, b, flag, tmem, out);
  release<<<1, 1, 0, c>>>(flag);
  cudaError_t e2 = cudaDeviceSynchronize();
  cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
  printf("grid %d smem %d KB cluster %d bigparams %d tmem %d: launch %s sync %s -> %s\n", grid, smem_kb, cluster, big, tmem,
         cudaGetErrorString(e), cudaGetErrorString(e2), h == 1 ? "RELEASED (concurrent)" : h == 0 ? "TIMED OUT (serialised)" : "?");
  return 0;
}
