// Shared host/device helpers for the vyom_b200 C-ABI library.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "../../include/vyom_b200.h"

namespace vy {

// ---- host: error plumbing ---------------------------------------------------------------------
void set_error(const char* fmt, ...);  // thread-local, defined in api.cu

#define VY_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::vy::set_error(__VA_ARGS__);        \
      return VY_ERR_INVALID_ARG;           \
    }                                      \
  } while (0)

#define VY_CUDA_OK(expr)                                                               \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::vy::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                      __LINE__);                                                       \
      return VY_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

#define VY_LAUNCH_OK()                                                                  \
  do {                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      ::vy::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                      __LINE__);                                                        \
      return VY_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

inline size_t dtype_size(int dt) { return dt == VY_BF16 ? 2 : 4; }
inline bool dtype_ok(int dt) { return dt == VY_F32 || dt == VY_BF16; }

int num_sms();  // cached, current device

extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// TMA descriptor creation (tensormap.cu). dims/strides innermost first; strides in BYTES for
// dims 1..rank-1 (dim 0 is contiguous). swizzle128: 0 none, 1 = SWIZZLE_128B (16-B atoms),
// 2 = SWIZZLE_128B_ATOM_32B (what MN-major tf32 UMMA operands need), 3 = SWIZZLE_64B (the epilogue's 64-byte staged
// rows). Returns 0 or VY_ERR_*.
int make_tensor_map(CUtensorMap* out, int dtype, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle128);

// Same, memoised on (pointer, geometry, box, swizzle).
int get_tensor_map_cached(CUtensorMap* out, int dtype, int rank, const void* base,
                          const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                          int swizzle);

// ---- launches with programmatic dependent launch (PDL) ------------------------------------------
// A training step is ~350 dependent kernels of 5-100 us. Each kernel is launched with the programmatic-stream-
// serialization attribute, calls pdl_trigger() first thing (so its successor may start launching as soon as every
// CTA of this grid is running or done) and pdl_wait() before its first access to global memory (which blocks until
// the predecessor grid has completed and its writes are visible). The successor's launch latency and prologue
// (barrier init, TMEM allocation, descriptor prefetch) then overlap the predecessor's tail instead of following it.
// Opt-in with VY_PDL=1 (measured on the captured training step: 13.04 ms with, 12.95 ms without — the persistent
// one-CTA-per-SM kernels leave no room for a successor's CTAs until they exit); off, the in-kernel instructions are no-ops.
bool pdl_enabled();  // api.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                         int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  int n = 1;
  if (cluster_x > 1) {  // thread-block cluster (grid.x must be a multiple of cluster_x)
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    n = 2;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_kernel_cluster(kern, grid, block, smem, st, 1, static_cast<Args&&>(args)...);
}

// ---- device: dtype-generic element access ---------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float ld_as_float(const void* p, int dt, int64_t i) {
  return dt == VY_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                       : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_from_float(void* p, int dt, int64_t i, float v) {
  if (dt == VY_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}
// 8 consecutive elements starting at element index i (i % 8 == 0, base 16B aligned)
__device__ __forceinline__ void ld8_as_float(const void* p, int dt, int64_t i, float (&v)[8]) {
  if (dt == VY_BF16) {
    uint4 raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = __bfloat1622float2(h[k]);
      v[2 * k] = f.x;
      v[2 * k + 1] = f.y;
    }
  } else {
    const float4* q = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
    float4 a = q[0], b = q[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
__device__ __forceinline__ void st8_from_float(void* p, int dt, int64_t i, const float (&v)[8]) {
  if (dt == VY_BF16) {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = raw;
  } else {
    float4* q = reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i);
    q[0] = make_float4(v[0], v[1], v[2], v[3]);
    q[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// ---- dropout: counter-based mask (Philox4x32-10), regenerated in backward instead of stored --------------------
// One call yields 128 random bits = eight 16-bit lanes = the keep decisions of one 8-element vector:
// element j of vector `vec` is kept iff lane j >= thresh (thresh = round(p * 65536)).
struct DropArgs {
  float p;                // 0 = off
  float scale;            // 1 / (1 - p)
  unsigned int thresh;    // round(p * 65536)
  unsigned int offset;    // distinguishes the call sites of one step
  unsigned long long seed;
  const int* step_ptr;    // device step counter or null (= 0)
};
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// keep bits (bit j = element j kept) of 8-element vector number `vec` of the call
__device__ __forceinline__ unsigned int dropout_keep8(const DropArgs& d, unsigned long long vec, unsigned int step) {
  const uint4 r = philox4x32_10(make_uint4(static_cast<unsigned int>(vec), static_cast<unsigned int>(vec >> 32), d.offset, step),
                                make_uint2(static_cast<unsigned int>(d.seed), static_cast<unsigned int>(d.seed >> 32)));
  const unsigned int w[4] = {r.x, r.y, r.z, r.w};
  unsigned int keep = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) keep |= (((w[j >> 1] >> (16 * (j & 1))) & 0xffffu) >= d.thresh ? 1u : 0u) << j;
  return keep;
}

// Exact-erf GELU (nn.GELU(), VyomAI/layers/ffn.py:8) without libdevice's branchy erff: with
// Q(x) = 1 - Phi(|x|) = 0.5 erfc(|x| / sqrt 2) from Abramowitz-Stegun 7.1.26 (|error| <= 0.75e-7, far
// below bf16 / tf32 resolution), gelu(x) = x Phi(x) and gelu'(x) = Phi(x) + x phi(x) share one
// exp2 and one reciprocal (two MUFU ops), ~14 FP32 instructions per element, no divergence.
__device__ __forceinline__ float vy_rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float vy_ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// returns Q(x) = 1 - Phi(|x|); e = exp(-x^2 / 2)
__device__ __forceinline__ float gauss_tail(float x, float& e) {
  const float t = vy_rcp_approx(fmaf(fabsf(x), 0.3275911f * 0.70710678118654752440f, 1.0f));
  e = vy_ex2_approx(x * x * (-0.5f * 1.44269504088896340736f));
  float p = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  p = fmaf(t, p, 0.5f * 1.421413741f);
  p = fmaf(t, p, 0.5f * -0.284496736f);
  p = fmaf(t, p, 0.5f * 0.254829592f);
  return p * (t * e);
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  const float q = gauss_tail(x, e);       // 1 - Phi(|x|)
  return fmaf(-fabsf(x), q, fmaxf(x, 0.f));  // x >= 0: x - x q;  x < 0: x q = -|x| q
}
__device__ __forceinline__ float dgelu_erf(float x) {
  float e;
  const float q = gauss_tail(x, e);
  const float cdf = x >= 0.f ? 1.0f - q : q;
  return fmaf(x * e, 0.39894228040143267794f, cdf);
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k = 0.7978845608028654f;
  return 0.5f * x * (1.0f + tanhf(k * (x + 0.044715f * x * x * x)));
}
__device__ __forceinline__ float dgelu_tanh(float x) {
  const float k = 0.7978845608028654f;
  const float u = k * (x + 0.044715f * x * x * x);
  const float t = tanhf(u);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * k * (1.0f + 3.0f * 0.044715f * x * x);
}
#endif

}  // namespace vy
