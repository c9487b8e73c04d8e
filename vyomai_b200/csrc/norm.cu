// Fused residual-add + LayerNorm, warp per row, rows held in registers between the statistics
// pass and the normalisation pass (one HBM read of x/residual, one write of y).
// HBM-bound: algorithmic bytes per row = (2 reads [+1 optional sum write] + 1 write) * H * sizeof.
#include "vy_common.cuh"
#include "vy_ptx.cuh"

namespace vy {

constexpr int NORM_WARPS = 8;

// NV = number of 8-element vectors each lane holds (H <= NV * 256)
template <int NV>
__global__ void __launch_bounds__(NORM_WARPS * 32)
add_layernorm_fwd_kernel(int rows, int H, const void* __restrict__ x, const void* __restrict__ res,
                         int io_dt, const void* __restrict__ gamma, const void* __restrict__ beta,
                         int p_dt, float eps, void* __restrict__ y, void* __restrict__ sum_out,
                         float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = H >> 3;
  for (int row = blockIdx.x * NORM_WARPS + warp; row < rows; row += gridDim.x * NORM_WARPS) {
    const long long base = static_cast<long long>(row) * H;
    float v[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        ld8_as_float(x, io_dt, base + vi * 8, v[i]);
        if (res) {
          float r[8];
          ld8_as_float(res, io_dt, base + vi * 8, r);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] += r[j];
        }
        if (sum_out) st8_from_float(sum_out, io_dt, base + vi * 8, v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[i][j];
      }
    }
    const float mean = warp_sum(sum) / H;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mean;
          sq += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / H + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float g[8], b[8], o[8];
        ld8_as_float(gamma, p_dt, vi * 8, g);
        ld8_as_float(beta, p_dt, vi * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
        st8_from_float(y, io_dt, base + vi * 8, o);
      }
    }
  }
}

// backward: persistent blocks stride over rows; per-lane dgamma/dbeta partials stay in registers,
// are combined across the block's warps in smem and written to partials[block][H].
template <int NV>
__global__ void __launch_bounds__(NORM_WARPS * 32)
add_layernorm_bwd_kernel(int rows, int H, const void* __restrict__ dy, const void* __restrict__ s,
                         int io_dt, const void* __restrict__ gamma, int p_dt,
                         const float* __restrict__ mean, const float* __restrict__ rstd,
                         void* __restrict__ dx, float* __restrict__ partials) {
  extern __shared__ float red[];  // [NORM_WARPS][2][H]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = H >> 3;
  float dg[NV][8], db[NV][8], g[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) ld8_as_float(gamma, p_dt, vi * 8, g[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dg[i][j] = 0.f;
      db[i][j] = 0.f;
    }
  }
  for (int row = blockIdx.x * NORM_WARPS + warp; row < rows; row += gridDim.x * NORM_WARPS) {
    const long long base = static_cast<long long>(row) * H;
    const float mu = mean[row], rs = rstd[row];
    float xh[NV][8], dyv[NV][8];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        ld8_as_float(dy, io_dt, base + vi * 8, dyv[i]);
        ld8_as_float(s, io_dt, base + vi * 8, xh[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (xh[i][j] - mu) * rs;
          const float t = dyv[i][j] * g[i][j];
          c1 += t;
          c2 += t * xh[i][j];
          dg[i][j] += dyv[i][j] * xh[i][j];
          db[i][j] += dyv[i][j];
        }
      }
    }
    c1 = warp_sum(c1) / H;
    c2 = warp_sum(c2) / H;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rs * (dyv[i][j] * g[i][j] - c1 - xh[i][j] * c2);
        st8_from_float(dx, io_dt, base + vi * 8, o);
      }
    }
  }
  float* my = red + static_cast<size_t>(warp) * 2 * H;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        my[vi * 8 + j] = dg[i][j];
        my[H + vi * 8 + j] = db[i][j];
      }
    }
  }
  __syncthreads();
  float* outp = partials + static_cast<size_t>(blockIdx.x) * 2 * H;
  for (int c = threadIdx.x; c < 2 * H; c += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < NORM_WARPS; ++w) a += red[static_cast<size_t>(w) * 2 * H + c];
    outp[c] = a;
  }
}

// column reduce of the per-block partials: 32 columns per block, 8 warps split the partial rows
__global__ void __launch_bounds__(256)
norm_bwd_reduce_kernel(int nparts, int H, const float* __restrict__ partials, void* __restrict__ dgamma,
                       void* __restrict__ dbeta, int out_dt, int accumulate) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float a = 0.f;
  if (c < 2 * H)
    for (int p = w; p < nparts; p += 8) a += partials[static_cast<size_t>(p) * 2 * H + c];
  red[w][lane] = a;
  __syncthreads();
  if (w == 0 && c < 2 * H) {
#pragma unroll
    for (int k = 1; k < 8; ++k) a += red[k][lane];
    void* dst = c < H ? dgamma : dbeta;
    const int cc = c < H ? c : c - H;
    if (accumulate) a += ld_as_float(dst, out_dt, cc);
    st_from_float(dst, out_dt, cc, a);
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int norm_common_checks(const VyNorm* p, const char* who) {
  VY_CHECK_ARG(p != nullptr, "%s: null params", who);
  if (!vy_device_ok()) {
    set_error("%s: no sm_100 device (there is no CPU fallback)", who);
    return VY_ERR_NO_DEVICE;
  }
  VY_CHECK_ARG(p->rows > 0 && p->H > 0, "%s: bad shape rows=%d H=%d", who, p->rows, p->H);
  VY_CHECK_ARG(p->H % 8 == 0 && p->H <= 2048, "%s: H must be a multiple of 8 and <= 2048 (got %d)", who, p->H);
  VY_CHECK_ARG(dtype_ok(p->io_dtype) && dtype_ok(p->param_dtype), "%s: bad dtype", who);
  return VY_OK;
}

}  // namespace vy

extern "C" int vy_norm_bwd_partial_rows(void) { return 2 * vy::num_sms(); }

extern "C" int vy_add_layernorm_fwd(const VyNorm* p) {
  using namespace vy;
  int rc = norm_common_checks(p, "vy_add_layernorm_fwd");
  if (rc != VY_OK) return rc;
  VY_CHECK_ARG(p->x && p->y && p->gamma && p->beta, "vy_add_layernorm_fwd: null pointer");
  VY_CHECK_ARG(aligned16(p->x) && aligned16(p->y) && aligned16(p->gamma) && aligned16(p->beta) &&
                   aligned16(p->residual) && aligned16(p->sum_out),
               "vy_add_layernorm_fwd: pointers must be 16-byte aligned");
  const int nv = (p->H + 255) / 256;
  int grid = (p->rows + NORM_WARPS - 1) / NORM_WARPS;
  const int maxgrid = num_sms() * 8;
  if (grid > maxgrid) grid = maxgrid;
  cudaStream_t st = static_cast<cudaStream_t>(p->stream);
#define VY_LN_FWD(NV)                                                                              \
  add_layernorm_fwd_kernel<NV><<<grid, NORM_WARPS * 32, 0, st>>>(                                  \
      p->rows, p->H, p->x, p->residual, p->io_dtype, p->gamma, p->beta, p->param_dtype, p->eps, p->y, \
      p->sum_out, p->mean, p->rstd)
  switch (nv) {
    case 1: VY_LN_FWD(1); break;
    case 2: VY_LN_FWD(2); break;
    case 3: VY_LN_FWD(3); break;
    case 4: VY_LN_FWD(4); break;
    case 5: VY_LN_FWD(5); break;
    case 6: VY_LN_FWD(6); break;
    case 7: VY_LN_FWD(7); break;
    default: VY_LN_FWD(8); break;
  }
#undef VY_LN_FWD
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_add_layernorm_bwd(const VyNorm* p) {
  using namespace vy;
  int rc = norm_common_checks(p, "vy_add_layernorm_bwd");
  if (rc != VY_OK) return rc;
  VY_CHECK_ARG(p->dy && p->s && p->dx && p->gamma && p->mean && p->rstd && p->dgamma && p->dbeta && p->partials,
               "vy_add_layernorm_bwd: null pointer");
  VY_CHECK_ARG(aligned16(p->dy) && aligned16(p->s) && aligned16(p->dx) && aligned16(p->gamma),
               "vy_add_layernorm_bwd: pointers must be 16-byte aligned");
  VY_CHECK_ARG(dtype_ok(p->dparam_dtype), "vy_add_layernorm_bwd: bad dparam_dtype");
  const int nv = (p->H + 255) / 256;
  int grid = (p->rows + NORM_WARPS - 1) / NORM_WARPS;
  const int nparts = vy_norm_bwd_partial_rows();
  if (grid > nparts) grid = nparts;
  const size_t smem = static_cast<size_t>(NORM_WARPS) * 2 * p->H * sizeof(float);
  cudaStream_t st = static_cast<cudaStream_t>(p->stream);
#define VY_LN_BWD(NV)                                                                               \
  do {                                                                                              \
    auto kern = add_layernorm_bwd_kernel<NV>;                                                       \
    if (smem > 48 * 1024)                                                                           \
      VY_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, NORM_WARPS * 32, smem, st>>>(p->rows, p->H, p->dy, p->s, p->io_dtype, p->gamma,    \
                                              p->param_dtype, p->mean, p->rstd, p->dx, p->partials); \
  } while (0)
  switch (nv) {
    case 1: VY_LN_BWD(1); break;
    case 2: VY_LN_BWD(2); break;
    case 3: VY_LN_BWD(3); break;
    case 4: VY_LN_BWD(4); break;
    case 5: VY_LN_BWD(5); break;
    case 6: VY_LN_BWD(6); break;
    case 7: VY_LN_BWD(7); break;
    default: VY_LN_BWD(8); break;
  }
#undef VY_LN_BWD
  VY_LAUNCH_OK();
  norm_bwd_reduce_kernel<<<(2 * p->H + 31) / 32, 256, 0, st>>>(grid, p->H, p->partials, p->dgamma, p->dbeta,
                                                               p->dparam_dtype, p->dparam_accumulate);
  VY_LAUNCH_OK();
  count_launch(2);
  return VY_OK;
}
