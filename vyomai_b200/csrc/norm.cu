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
                         float* __restrict__ mean_out, float* __restrict__ rstd_out, int kind, const DropArgs drop) {
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = H >> 3;
  {
    // gamma / beta are needed only after the row statistics: with few rows (a decode step: one warp, one row) that second,
    // dependent round trip to DRAM was half of the kernel's 11 us. Pull the lines into L2 now — before the predecessor is
    // even awaited (a prefetch of a line that is rewritten later is harmless: L2 is the point of coherence).
    const int pb = p_dt == VY_BF16 ? 2 : 4;
    for (int off = (warp * 32 + lane) * 128; off < H * pb; off += NORM_WARPS * 32 * 128) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const char*>(gamma) + off));
      if (beta) asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const char*>(beta) + off));
    }
  }
  pdl_wait();
  const unsigned int drop_step = (drop.p > 0.f && drop.step_ptr) ? static_cast<unsigned int>(*drop.step_ptr) : 0u;
  for (int row = blockIdx.x * NORM_WARPS + warp; row < rows; row += gridDim.x * NORM_WARPS) {
    const long long base = static_cast<long long>(row) * H;
    float v[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        ld8_as_float(x, io_dt, base + vi * 8, v[i]);
        if (drop.p > 0.f) {  // dropout(x) before the residual add (attention.py:70, ffn.py:38)
          const unsigned int keep = dropout_keep8(drop, static_cast<unsigned long long>(row) * nvec + vi, drop_step);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] = (keep >> j) & 1u ? v[i][j] * drop.scale : 0.f;
        }
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
    // kind 0: LayerNorm. kind 1 / 2: RMSNorm — no mean, y = w * xhat (2: Gemma's (1 + w) * xhat), optional shift beta
    const float mean = kind == VY_NORM_LAYER ? warp_sum(sum) / H : 0.f;
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
        if (beta) {
          ld8_as_float(beta, p_dt, vi * 8, b);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) b[j] = 0.f;
        }
        if (kind == VY_NORM_LAYER) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
        } else {
          // the reference rounds xhat to the activation dtype BEFORE the weight multiply
          // (models/custom_transformer.py:236-240: `self.weight * hidden_states.to(input_dtype)`)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float xh = v[i][j] * rstd;
            if (io_dt == VY_BF16) xh = __bfloat162float(__float2bfloat16_rn(xh));
            o[j] = xh * (kind == VY_NORM_RMS_GEMMA ? 1.f + g[j] : g[j]) + b[j];
          }
        }
        st8_from_float(y, io_dt, base + vi * 8, o);
      }
    }
  }
}

// Few rows (a decode step: 1-64 tokens): one CTA per row, 256 threads x 8 elements, so the row, gamma and beta are all
// requested at once and the statistics are two block reductions — the warp-per-row kernel above walks a 2048-wide row with ONE
// warp (8 dependent-looking vector loads per lane, then gamma after the statistics): 10-11 us per call in the PaliGemma-scale
// decode step, 37 calls per token.
__global__ void __launch_bounds__(256)
add_layernorm_fwd_small_kernel(int rows, int H, const void* __restrict__ x, const void* __restrict__ res, int io_dt,
                               const void* __restrict__ gamma, const void* __restrict__ beta, int p_dt, float eps,
                               void* __restrict__ y, void* __restrict__ sum_out, float* __restrict__ mean_out,
                               float* __restrict__ rstd_out, int kind) {
  __shared__ float s_red[2][8];
  pdl_trigger();
  const int row = blockIdx.x, t = threadIdx.x;
  const int nvec = H >> 3;
  const bool on = t < nvec;
  float g[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = b[j] = 0.f;
  if (on) {  // parameters do not depend on the predecessor: requested before it is awaited
    ld8_as_float(gamma, p_dt, t * 8, g);
    if (beta) ld8_as_float(beta, p_dt, t * 8, b);
  }
  pdl_wait();
  const long long base = static_cast<long long>(row) * H;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  if (on) {
    ld8_as_float(x, io_dt, base + t * 8, v);
    if (res) {
      float r[8];
      ld8_as_float(res, io_dt, base + t * 8, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
    if (sum_out) st8_from_float(sum_out, io_dt, base + t * 8, v);
  }
  auto block_sum = [&](float a, int slot) {
    a = warp_sum(a);
    if ((t & 31) == 0) s_red[slot][t >> 5] = a;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_red[slot][w];
    return tot;
  };
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) sum += v[j];
  const float mean = kind == VY_NORM_LAYER ? block_sum(sum, 0) / H : 0.f;
  float sq = 0.f;
  if (on) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[j] - mean;
      sq += d * d;
    }
  }
  const float rstd = rsqrtf(block_sum(sq, 1) / H + eps);
  if (t == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  if (on) {
    float o[8];
    if (kind == VY_NORM_LAYER) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd * g[j] + b[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float xh = v[j] * rstd;
        if (io_dt == VY_BF16) xh = __bfloat162float(__float2bfloat16_rn(xh));  // (same rounding point as the row kernel)
        o[j] = xh * (kind == VY_NORM_RMS_GEMMA ? 1.f + g[j] : g[j]) + b[j];
      }
    }
    st8_from_float(y, io_dt, base + t * 8, o);
  }
}

// backward: one pass over dy and s. Persistent CTAs (one per SM, 8 warps) stride over rows, warp per row; the raw
// 16-byte loads of the NEXT row are issued before the current row is processed, so every warp keeps two rows in
// flight (~7 MB over the chip: enough to cover HBM latency at full bandwidth with only 8 warps per SM). Per-lane
// column accumulators stay in registers: dgamma += dy*xhat, dbeta += dy and — when asked — dbias += dx (the bias
// gradient of the Linear whose output fed this LayerNorm: attention.py:69-71, ffn.py:37-39, saving a separate
// column-sum pass over dx). They are combined across the CTA's warps in smem and written to partials[cta][3][H].
// Row ring of the single-pass backward: every warp streams its rows of dy and s through LN_RING slots of shared memory
// with 1-D bulk copies (cp.async.bulk, mbarrier tx-count), issued LN_RING rows ahead by one lane. In-flight bytes per SM
// = 8 warps x LN_RING rows x 2 tensors x H x sizeof — what an HBM-bound kernel with one warp per row needs to cover the
// memory latency (registers only held ONE row ahead: 2.0 TB/s; see profiles/). The ring is dead when the column
// partials are staged, so `red` aliases it.
template <bool IO_BF16>
struct LnRing {
  static constexpr int SLOTS = IO_BF16 ? 4 : 2;
};
__device__ __forceinline__ void bulk_load_row(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int NV, bool IO_BF16, bool DROP>
__global__ void __launch_bounds__(NORM_WARPS * 32, 1)
add_layernorm_bwd_kernel(int rows, int H, const void* __restrict__ dy, const void* __restrict__ s,
                         const void* __restrict__ gamma, int p_dt, const float* __restrict__ mean,
                         const float* __restrict__ rstd, void* __restrict__ dx, int want_dbias, float* __restrict__ partials,
                         int kind, void* __restrict__ dx_drop, const DropArgs drop) {
  pdl_trigger();
  pdl_wait();
  const unsigned int drop_step = (DROP && drop.step_ptr) ? static_cast<unsigned int>(*drop.step_ptr) : 0u;
  extern __shared__ __align__(128) float red[];  // [NORM_WARPS][3][H] at the end; the row ring before that
  constexpr int IO_DT = IO_BF16 ? VY_BF16 : VY_F32;
  constexpr int RAW = IO_BF16 ? 1 : 2;  // uint4 per 8 elements
  constexpr int D = LnRing<IO_BF16>::SLOTS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = H >> 3;
  const uint32_t row_bytes = static_cast<uint32_t>(H) * (IO_BF16 ? 2u : 4u);
  // ring layout: [warp][slot][dy row | s row], then the mbarriers
  uint8_t* ring = reinterpret_cast<uint8_t*>(red) + static_cast<size_t>(warp) * D * 2 * row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(red) + static_cast<size_t>(NORM_WARPS) * D * 2 * row_bytes) + warp * D;
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < D; ++d) mbar_init(&bars[d], 1);
    fence_mbar_init();
  }
  __syncwarp();
  float g[NV][8], dg[NV][8], db[NV][8], dbi[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
#pragma unroll
    for (int j = 0; j < 8; ++j) g[i][j] = dg[i][j] = db[i][j] = dbi[i][j] = 0.f;
    if (vi < nvec) {
      ld8_as_float(gamma, p_dt, vi * 8, g[i]);
      if (kind == VY_NORM_RMS_GEMMA) {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[i][j] += 1.f;
      }
    }
  }
  const bool rms = kind != VY_NORM_LAYER;  // RMSNorm: xhat = x * rstd, dx = rstd * (dy g - xhat mean(dy g xhat))
  const int stride = gridDim.x * NORM_WARPS;
  float mu_r[D], rs_r[D];  // per-row statistics of the rows in flight (loaded when the row's copies are issued)
  auto issue = [&](int d, int row) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[d], 2 * row_bytes);
      uint8_t* slot = ring + static_cast<size_t>(d) * 2 * row_bytes;
      bulk_load_row(slot, reinterpret_cast<const uint8_t*>(dy) + static_cast<size_t>(row) * row_bytes, row_bytes, &bars[d]);
      bulk_load_row(slot + row_bytes, reinterpret_cast<const uint8_t*>(s) + static_cast<size_t>(row) * row_bytes, row_bytes, &bars[d]);
    }
  };
  const int first = blockIdx.x * NORM_WARPS + warp;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const int row = first + d * stride;
    mu_r[d] = rs_r[d] = 0.f;
    if (row < rows) {
      issue(d, row);
      mu_r[d] = rms ? 0.f : mean[row];
      rs_r[d] = rstd[row];
    }
  }
  // Two rows per step: their dependency chains (unpack -> row sums -> warp butterflies -> dx) are independent, and with
  // only two warps per SM sub-partition it is this instruction-level parallelism, not more loads in flight, that fills
  // the issue slots (one row per step: 18 us per call, latency-bound).
  uint32_t ph = 0;
  for (int base = first; base < rows; base += D * stride, ph ^= 1) {
#pragma unroll
    for (int d = 0; d < D; d += 2) {
      const int row0 = base + d * stride;
      if (row0 >= rows) break;
      const int row1 = row0 + stride;
      const bool two = row1 < rows;
      const float mu[2] = {mu_r[d], mu_r[d + 1]}, rs[2] = {rs_r[d], rs_r[d + 1]};
      float xh[2][NV][8], dyv[2][NV][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        mbar_wait(&bars[d + u], ph);
        const uint4* sd = reinterpret_cast<const uint4*>(ring + static_cast<size_t>(d + u) * 2 * row_bytes);
        const uint4* ss = reinterpret_cast<const uint4*>(ring + static_cast<size_t>(d + u) * 2 * row_bytes + row_bytes);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int vi = lane + i * 32;
          if (vi < nvec) {
            uint4 rd[RAW], rx[RAW];
#pragma unroll
            for (int k = 0; k < RAW; ++k) {
              rd[k] = sd[vi * RAW + k];
              rx[k] = ss[vi * RAW + k];
            }
            ld8_as_float(&rd[0], IO_DT, 0, dyv[u][i]);
            ld8_as_float(&rx[0], IO_DT, 0, xh[u][i]);
          }
        }
      }
      __syncwarp();  // every lane has read the slots: refill them with the rows LN_RING steps ahead
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int next = row0 + u * stride + D * stride;
        if (next < rows) {
          issue(d + u, next);
          mu_r[d + u] = rms ? 0.f : mean[next];
          rs_r[d + u] = rstd[next];
        }
      }
      float c1[2] = {0.f, 0.f}, c2[2] = {0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (lane + i * 32 < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              xh[u][i][j] = (xh[u][i][j] - mu[u]) * rs[u];
              const float t = dyv[u][i][j] * g[i][j];
              c1[u] += t;
              c2[u] += t * xh[u][i][j];
              dg[i][j] += dyv[u][i][j] * xh[u][i][j];
              db[i][j] += dyv[u][i][j];
            }
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {  // the four butterflies interleaved
        c1[0] += __shfl_xor_sync(0xffffffffu, c1[0], o);
        c2[0] += __shfl_xor_sync(0xffffffffu, c2[0], o);
        c1[1] += __shfl_xor_sync(0xffffffffu, c1[1], o);
        c2[1] += __shfl_xor_sync(0xffffffffu, c2[1], o);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        const float m1 = rms ? 0.f : c1[u] / H, m2 = c2[u] / H;
        const long long rbase = static_cast<long long>(row0 + u * stride) * H;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int vi = lane + i * 32;
          if (vi < nvec) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              o[j] = rs[u] * (dyv[u][i][j] * g[i][j] - m1 - xh[u][i][j] * m2);
              if (!DROP) dbi[i][j] += o[j];
            }
            st8_from_float(dx, IO_DT, rbase + vi * 8, o);
            if (DROP) {  // gradient of the dropped branch: the forward's mask, regenerated
              const unsigned int keep = dropout_keep8(drop, static_cast<unsigned long long>(row0 + u * stride) * nvec + vi, drop_step);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                o[j] = (keep >> j) & 1u ? o[j] * drop.scale : 0.f;
                dbi[i][j] += o[j];
              }
              st8_from_float(dx_drop, IO_DT, rbase + vi * 8, o);
            }
          }
        }
      }
    }
  }
  __syncthreads();  // all rings are drained (every issued row was consumed): the staging below reuses their memory
  float* my = red + static_cast<size_t>(warp) * 3 * H;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        my[vi * 8 + j] = dg[i][j];
        my[H + vi * 8 + j] = db[i][j];
        my[2 * H + vi * 8 + j] = dbi[i][j];
      }
    }
  }
  __syncthreads();
  float* outp = partials + static_cast<size_t>(blockIdx.x) * 3 * H;
  const int nout = want_dbias ? 3 * H : 2 * H;
  for (int c = threadIdx.x; c < nout; c += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < NORM_WARPS; ++w) a += red[static_cast<size_t>(w) * 3 * H + c];
    outp[c] = a;
  }
}

// Wide rows (H > 1024) do not fit the register accumulators above: dx by a light warp-per-row kernel, then the column
// sums over row strips (partials[strip][3][H]).
template <int NV>
__global__ void __launch_bounds__(NORM_WARPS * 32, 2)
add_layernorm_bwd_dx_kernel(int rows, int H, const void* __restrict__ dy, const void* __restrict__ s, int io_dt,
                            const void* __restrict__ gamma, int p_dt, const float* __restrict__ mean,
                            const float* __restrict__ rstd, void* __restrict__ dx) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = H >> 3;
  float g[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) ld8_as_float(gamma, p_dt, vi * 8, g[i]);
  }
  for (int row = blockIdx.x * NORM_WARPS + warp; row < rows; row += gridDim.x * NORM_WARPS) {
    const long long base = static_cast<long long>(row) * H;
    const float mu = mean[row], rs = rstd[row];
    float xh[NV][8], t[NV][8];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        ld8_as_float(dy, io_dt, base + vi * 8, t[i]);
        ld8_as_float(s, io_dt, base + vi * 8, xh[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (xh[i][j] - mu) * rs;
          t[i][j] *= g[i][j];
          c1 += t[i][j];
          c2 += t[i][j] * xh[i][j];
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
        for (int j = 0; j < 8; ++j) o[j] = rs * (t[i][j] - c1 - xh[i][j] * c2);
        st8_from_float(dx, io_dt, base + vi * 8, o);
      }
    }
  }
}

constexpr int NORM_MAX_STRIPS = 128;

// grid (ceil(H / 256), strips): a warp covers 256 consecutive columns (8 per lane), the 8 warps of the CTA take
// interleaved rows of the strip, 4 rows in flight each; partials[strip][3][H].
__global__ void __launch_bounds__(256)
norm_bwd_columns_kernel(int rows, int H, const void* __restrict__ dy, const void* __restrict__ s, const void* __restrict__ dx,
                        int io_dt, const float* __restrict__ mean, const float* __restrict__ rstd, int want_dbias,
                        float* __restrict__ partials) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][3][256 + 8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const int per = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float ag[8], ab[8], ax[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ag[j] = ab[j] = ax[j] = 0.f;
  if (c0 < H) {
    int r = r0 + w;
    for (; r + 24 < r1; r += 32) {
      float vd[4][8], vs[4][8], vx[4][8];
      float mu[4], rs[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long off = static_cast<long long>(r + u * 8) * H + c0;
        ld8_as_float(dy, io_dt, off, vd[u]);
        ld8_as_float(s, io_dt, off, vs[u]);
        if (want_dbias) ld8_as_float(dx, io_dt, off, vx[u]);
        mu[u] = mean[r + u * 8];
        rs[u] = rstd[r + u * 8];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          ag[j] += vd[u][j] * ((vs[u][j] - mu[u]) * rs[u]);
          ab[j] += vd[u][j];
          if (want_dbias) ax[j] += vx[u][j];
        }
    }
    for (; r < r1; r += 8) {
      float vd[8], vs[8], vx[8];
      const long long off = static_cast<long long>(r) * H + c0;
      ld8_as_float(dy, io_dt, off, vd);
      ld8_as_float(s, io_dt, off, vs);
      const float mu = mean[r], rs = rstd[r];
      if (want_dbias) ld8_as_float(dx, io_dt, off, vx);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ag[j] += vd[j] * ((vs[j] - mu) * rs);
        ab[j] += vd[j];
        if (want_dbias) ax[j] += vx[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[w][0][lane * 8 + j] = ag[j];
    red[w][1][lane * 8 + j] = ab[j];
    red[w][2][lane * 8 + j] = ax[j];
  }
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < H) {
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) a += red[k][v][threadIdx.x];
      partials[(static_cast<size_t>(blockIdx.y) * 3 + v) * H + c] = a;
    }
  }
}

// final reduce over the strips. CTA = 32 consecutive entries of the flattened [3][H] result; its 8 warps split the
// strips (coalesced 128-byte reads per warp and strip) and are combined in smem — ~70 CTAs of short loops instead of
// 9 CTAs each walking every strip.
__global__ void __launch_bounds__(256)
norm_bwd_reduce_kernel(int strips, int H, const float* __restrict__ partials, void* __restrict__ dgamma,
                       void* __restrict__ dbeta, void* __restrict__ dbias, int out_dt, int accumulate) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nv = dbias ? 3 : 2;
  const int idx = blockIdx.x * 32 + lane;  // entry of the flattened [nv][H] result
  float a = 0.f;
  if (idx < nv * H)
    for (int p = w; p < strips; p += 8) a += partials[static_cast<size_t>(p) * 3 * H + idx];
  red[w][lane] = a;
  __syncthreads();
  if (w == 0 && idx < nv * H) {
#pragma unroll
    for (int k = 1; k < 8; ++k) a += red[k][lane];
    const int v = idx / H, c = idx % H;
    void* dst = v == 0 ? dgamma : (v == 1 ? dbeta : dbias);
    if (accumulate) a += ld_as_float(dst, out_dt, c);
    st_from_float(dst, out_dt, c, a);
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

static int make_drop_args(const VyNorm* p, const char* who, DropArgs* d) {
  memset(d, 0, sizeof(*d));
  VY_CHECK_ARG(p->dropout_p >= 0.f && p->dropout_p < 1.f, "%s: dropout_p must lie in [0, 1) (got %g)", who, p->dropout_p);
  if (p->dropout_p == 0.f) return VY_OK;
  d->p = p->dropout_p;
  d->scale = 1.f / (1.f - p->dropout_p);
  d->thresh = static_cast<unsigned int>(p->dropout_p * 65536.f + 0.5f);
  d->offset = p->dropout_offset;
  d->seed = p->dropout_seed;
  d->step_ptr = p->dropout_step_ptr;
  return VY_OK;
}

}  // namespace vy

extern "C" int vy_norm_bwd_partial_rows(void) { return 2 * vy::num_sms(); }

extern "C" int vy_add_layernorm_fwd(const VyNorm* p) {
  using namespace vy;
  int rc = norm_common_checks(p, "vy_add_layernorm_fwd");
  if (rc != VY_OK) return rc;
  VY_CHECK_ARG(p->x && p->y && p->gamma && (p->beta || p->kind != VY_NORM_LAYER), "vy_add_layernorm_fwd: null pointer");
  VY_CHECK_ARG(p->kind >= VY_NORM_LAYER && p->kind <= VY_NORM_RMS_GEMMA, "vy_add_layernorm_fwd: bad kind %d", p->kind);
  VY_CHECK_ARG(aligned16(p->x) && aligned16(p->y) && aligned16(p->gamma) && aligned16(p->beta) &&
                   aligned16(p->residual) && aligned16(p->sum_out),
               "vy_add_layernorm_fwd: pointers must be 16-byte aligned");
  DropArgs drop;
  rc = make_drop_args(p, "vy_add_layernorm_fwd", &drop);
  if (rc != VY_OK) return rc;
  cudaStream_t st0 = static_cast<cudaStream_t>(p->stream);
  static const bool small_off = getenv("VY_NORM_SMALL") && atoi(getenv("VY_NORM_SMALL")) == 0;  // development: A/B
  if (p->rows <= 64 && drop.p <= 0.f && !small_off) {  // decode-sized calls: one CTA per row (see add_layernorm_fwd_small_kernel)
    VY_CUDA_OK(launch_kernel(add_layernorm_fwd_small_kernel, dim3(p->rows), dim3(256), 0, st0, p->rows, p->H, p->x, p->residual, p->io_dtype,
                             p->gamma, p->beta, p->param_dtype, p->eps, p->y, p->sum_out, p->mean, p->rstd, p->kind));
    VY_LAUNCH_OK();
    count_launch();
    return VY_OK;
  }
  const int nv = (p->H + 255) / 256;
  int grid = (p->rows + NORM_WARPS - 1) / NORM_WARPS;
  const int maxgrid = num_sms() * 8;
  if (grid > maxgrid) grid = maxgrid;
  cudaStream_t st = static_cast<cudaStream_t>(p->stream);
#define VY_LN_FWD(NV)                                                                              \
  VY_CUDA_OK(launch_kernel(add_layernorm_fwd_kernel<NV>, dim3(grid), dim3(NORM_WARPS * 32), 0, st,                                   \
      p->rows, p->H, p->x, p->residual, p->io_dtype, p->gamma, p->beta, p->param_dtype, p->eps, p->y, \
      p->sum_out, p->mean, p->rstd, p->kind, drop))
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
  VY_CHECK_ARG(p->dy && p->s && p->dx && p->gamma && (p->mean || p->kind != VY_NORM_LAYER) && p->rstd && p->dgamma && p->dbeta &&
                   p->partials,
               "vy_add_layernorm_bwd: null pointer");
  VY_CHECK_ARG(p->kind >= VY_NORM_LAYER && p->kind <= VY_NORM_RMS_GEMMA, "vy_add_layernorm_bwd: bad kind %d", p->kind);
  VY_CHECK_ARG(p->kind == VY_NORM_LAYER || p->H <= 1024, "vy_add_layernorm_bwd: RMSNorm backward supports H <= 1024 (got %d)", p->H);
  VY_CHECK_ARG(aligned16(p->dy) && aligned16(p->s) && aligned16(p->dx) && aligned16(p->gamma),
               "vy_add_layernorm_bwd: pointers must be 16-byte aligned");
  VY_CHECK_ARG(dtype_ok(p->dparam_dtype), "vy_add_layernorm_bwd: bad dparam_dtype");
  DropArgs drop;
  rc = make_drop_args(p, "vy_add_layernorm_bwd", &drop);
  if (rc != VY_OK) return rc;
  const bool dropping = drop.p > 0.f;
  if (dropping) {
    VY_CHECK_ARG(p->dx_drop != nullptr && aligned16(p->dx_drop), "vy_add_layernorm_bwd: dropout_p > 0 needs a 16-byte aligned dx_drop");
    if (p->H > 1024) {
      set_error("vy_add_layernorm_bwd: dropout backward supports H <= 1024 (got %d)", p->H);
      return VY_ERR_UNSUPPORTED;
    }
  }
  const int nv = (p->H + 255) / 256;
  int grid = (p->rows + NORM_WARPS - 1) / NORM_WARPS;
  if (grid > num_sms()) grid = num_sms();  // persistent: one CTA per SM (partials hold up to 2 * SMs * 2 * H floats)
  const size_t red_bytes = static_cast<size_t>(NORM_WARPS) * 3 * p->H * sizeof(float);
  const size_t ring_slots = p->io_dtype == VY_BF16 ? LnRing<true>::SLOTS : LnRing<false>::SLOTS;
  const size_t ring_bytes = static_cast<size_t>(NORM_WARPS) * ring_slots * 2 * p->H * dtype_size(p->io_dtype) + NORM_WARPS * ring_slots * 8;
  const size_t smem = red_bytes > ring_bytes ? red_bytes : ring_bytes;
  cudaStream_t st = static_cast<cudaStream_t>(p->stream);
#define VY_LN_BWD2(NV, BF, DR)                                                                        \
  do {                                                                                               \
    auto kern = add_layernorm_bwd_kernel<NV, BF, DR>;                                                \
    if (smem > 48 * 1024)                                                                            \
      VY_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    VY_CUDA_OK(launch_kernel(kern, dim3(grid), dim3(NORM_WARPS * 32), smem, st, p->rows, p->H, p->dy, p->s, p->gamma, p->param_dtype, p->mean, p->rstd, \
                                              p->dx, p->dbias != nullptr, p->partials, p->kind, p->dx_drop, drop));      \
  } while (0)
#define VY_LN_BWD(NV)                     \
  do {                                    \
    if (p->io_dtype == VY_BF16) { if (dropping) VY_LN_BWD2(NV, true, true); else VY_LN_BWD2(NV, true, false); } \
    else { if (dropping) VY_LN_BWD2(NV, false, true); else VY_LN_BWD2(NV, false, false); }           \
  } while (0)
  switch (nv) {
    case 1: VY_LN_BWD(1); break;
    case 2: VY_LN_BWD(2); break;
    case 3: VY_LN_BWD(3); break;
    case 4: VY_LN_BWD(4); break;
    default: break;
  }
#undef VY_LN_BWD
#undef VY_LN_BWD2
  int strips = grid;
  if (nv > 4) {
    int g2 = (p->rows + NORM_WARPS - 1) / NORM_WARPS;
    if (g2 > num_sms() * 8) g2 = num_sms() * 8;
#define VY_LN_DX(NV)                                                                                        \
  VY_CUDA_OK(launch_kernel(add_layernorm_bwd_dx_kernel<NV>, dim3(g2), dim3(NORM_WARPS * 32), 0, st, p->rows, p->H, p->dy, p->s, p->io_dtype, \
                                                                   p->gamma, p->param_dtype, p->mean, p->rstd, p->dx))
    switch (nv) {
      case 5: VY_LN_DX(5); break;
      case 6: VY_LN_DX(6); break;
      case 7: VY_LN_DX(7); break;
      default: VY_LN_DX(8); break;
    }
#undef VY_LN_DX
    VY_LAUNCH_OK();
    strips = (p->rows + 63) / 64;
    if (strips > NORM_MAX_STRIPS) strips = NORM_MAX_STRIPS;
    dim3 cgrid((p->H + 255) / 256, strips);
    VY_CUDA_OK(launch_kernel(norm_bwd_columns_kernel, dim3(cgrid), dim3(256), 0, st, p->rows, p->H, p->dy, p->s, p->dx, p->io_dtype, p->mean, p->rstd,
                                                   p->dbias != nullptr, p->partials));
  }
  VY_LAUNCH_OK();
  const int nvec = p->dbias ? 3 : 2;
  VY_CUDA_OK(launch_kernel(norm_bwd_reduce_kernel, dim3((nvec * p->H + 31) / 32), dim3(256), 0, st, strips, p->H, p->partials, p->dgamma, p->dbeta, p->dbias,
                                                                    p->dparam_dtype, p->dparam_accumulate));
  VY_LAUNCH_OK();
  count_launch(2);
  return VY_OK;
}
