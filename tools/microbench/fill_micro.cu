// Micro-benchmark (development): how long does a CTA take to bring a [32][K] bf16 activation block from L2 into shared
// memory — per-thread 16-byte loads vs one bulk copy per row — alone and with every SM doing it at once?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) fill_kernel(const uint4* __restrict__ src, int K, long long* out, int rounds) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  const int stride = K * 2 + 64, per_row = K >> 3, total = 32 * per_row;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned phase = 0;
  for (int it = 0; it < rounds; ++it) {
    __syncthreads();
    const unsigned long long t0 = gt();
    if (MODE == 0) {  // per-thread loads, 8 in flight
      for (int base = threadIdx.x; base < total; base += 512 * 8) {
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int i = base + j * 512; v[j] = make_uint4(0, 0, 0, 0); if (i < total) v[j] = __ldcg(src + i); }
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int i = base + j * 512; if (i < total) { const int r = i / per_row, c = i - r * per_row; *reinterpret_cast<uint4*>(smem + (size_t)r * stride + c * 16) = v[j]; } }
      }
    } else {  // one bulk copy per row, issued by 32 lanes of warp 0
      if (threadIdx.x < 32) {
        if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(32 * K * 2) : "memory");
        __syncwarp();
        const int r = threadIdx.x;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + (size_t)r * stride)),
                     "l"(reinterpret_cast<const unsigned char*>(src) + (size_t)r * K * 2), "r"(K * 2), "r"(s32(&bar)) : "memory");
      }
      unsigned ok = 0;
      while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(s32(&bar)), "r"(phase) : "memory");
      phase ^= 1;
    }
    __syncthreads();
    const unsigned long long t1 = gt();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[it] = (long long)(t1 - t0);
  }
}

int main() {
  const int Ks[2] = {768, 3072};
  uint4* src; long long* out;
  cudaMalloc(&src, 32 * 3072 * 2); cudaMemset(src, 1, 32 * 3072 * 2);
  cudaMalloc(&out, 64 * 8);
  cudaFuncSetAttribute(fill_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(fill_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  long long h[8];
  for (int ki = 0; ki < 2; ++ki) for (int grid : {1, 48, 148}) for (int mode = 0; mode < 2; ++mode) {
    const int K = Ks[ki];
    const size_t sm = 32 * (K * 2 + 64);
    if (mode == 0) fill_kernel<0><<<grid, 512, sm>>>(src, K, out, 8); else fill_kernel<1><<<grid, 512, sm>>>(src, K, out, 8);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
    printf("K=%4d grid=%3d %s: %s  ns per fill:", K, grid, mode ? "bulk/row " : "ldg+sts  ", cudaGetErrorString(e));
    for (int i = 0; i < 8; ++i) printf(" %lld", h[i]);
    printf("\n");
  }
  return 0;
}
