// Micro-benchmark (development): latency of ONE round of independent 16-byte loads per lane (8 in flight, each warp
// load 512 contiguous bytes at a pseudo-random offset) as a function of the footprint and of how many warps do it at once.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(512, 1) lat_kernel(const uint4* __restrict__ buf, size_t n_chunks /*512B chunks*/, int rounds, long long* out, unsigned seed, int active_warps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= active_warps) return;
  unsigned s = seed + (blockIdx.x * 16 + warp) * 2654435761u;
  uint4 acc = make_uint4(0, 0, 0, 0);
  long long total = 0, c0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    uint4 v[8];
    const unsigned long long t0 = gt();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s = s * 1664525u + 1013904223u;
      const size_t chunk = (size_t)(s >> 4) % n_chunks;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w) : "l"(buf + chunk * 32 + lane));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc.x ^= v[j].x; acc.y ^= v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
    if (acc.x == 0x12345678u) s ^= acc.y;  // dependency: the next round's addresses wait for this round's data
    total += (long long)(gt() - t0);
  }
  long long c1 = clock64();
  if (lane == 0 && warp == 0 && blockIdx.x == 0) { out[0] = total / rounds; out[1] = (c1 - c0) / rounds; out[2] = acc.z; }
}
int main() {
  long long* out; cudaMalloc(&out, 64);
  long long h[3];
  for (size_t mb : {8, 64, 200, 600}) {
    uint4* buf; cudaMalloc(&buf, mb << 20); cudaMemset(buf, 1, mb << 20);
    for (int grid : {1, 148}) for (int aw : {1, 16}) {
      lat_kernel<<<grid, 512>>>(buf, (mb << 20) / 512, 64, out, 12345u, aw);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
      printf("footprint %4zu MB grid %3d warps/CTA %2d: %s  round = %lld ns (%lld clk)\n", mb, grid, aw, cudaGetErrorString(e), h[0], h[1]);
    }
    cudaFree(buf);
  }
  return 0;
}
