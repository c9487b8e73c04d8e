// Bandwidth-bound helpers around the tensor-core kernels: embedding gather / scatter-add,
// patch extraction, greedy argmax, column sums (bias gradients), strided casts, fused
// softmax-cross-entropy, fused AdamW with global-norm clipping. All vectorised to 16-byte
// accesses where the layout allows; grid sizes are multiples of the SM count.
#include <cuda_fp16.h>

#include <type_traits>

#include "vy_common.cuh"
#include "vy_ptx.cuh"

namespace vy {

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

#define VY_NEED_DEVICE(who)                                             \
  do {                                                                  \
    if (!vy_device_ok()) {                                              \
      set_error("%s: no sm_100 device (there is no CPU fallback)", who); \
      return VY_ERR_NO_DEVICE;                                          \
    }                                                                   \
  } while (0)

// ------------------------------------------------------------------------------------------
// rows gather (+ positional add, + scale, + row remap). One warp per output row.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_fwd_kernel(int rows, int H, const long long* __restrict__ ids, const void* __restrict__ src,
                 long long ld_src, int dt, int vocab, int tokens_per_seq, int out_group_stride,
                 int out_row_off, const void* __restrict__ pos, int pos_row_off, float out_scale,
                 void* __restrict__ out, long long ld_out) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = H >> 3;
  for (int r = warp; r < rows; r += nwarps) {
    long long srow = r;
    if (ids) {
      srow = ids[r];
      if (srow < 0 || srow >= vocab) srow = 0;  // torch would raise; keep the kernel memory-safe
    }
    const int l = r % tokens_per_seq;
    const long long orow = static_cast<long long>(r / tokens_per_seq) * out_group_stride + l + out_row_off;
    for (int vi = lane; vi < nvec; vi += 32) {
      float v[8];
      ld8_as_float(src, dt, srow * ld_src + vi * 8, v);
      if (pos) {
        float p[8];
        ld8_as_float(pos, dt, static_cast<long long>(pos_row_off + l) * H + vi * 8, p);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += p[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= out_scale;
      st8_from_float(out, dt, orow * ld_out + vi * 8, v);
    }
  }
}

// scatter-add of output-row gradients into the embedding table / position table (fp32 atomics
// for fp32 tables, packed bf16x2 atomics for bf16 tables).
__device__ __forceinline__ void atomic_add_elem2(void* base, int dt, long long idx, float a, float b) {
  if (dt == VY_BF16) {
    atomicAdd(reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(base) + idx),
              __floats2bfloat162_rn(a, b));
  } else {
    float* p = reinterpret_cast<float*>(base) + idx;
    atomicAdd(p, a);
    atomicAdd(p + 1, b);
  }
}

__global__ void __launch_bounds__(256)
embed_bwd_kernel(int rows, int H, const long long* __restrict__ ids, int vocab, int tokens_per_seq,
                 int out_group_stride, int out_row_off, int pos_row_off, float scale,
                 const void* __restrict__ dout, long long ld_dout, int dt, void* __restrict__ dtable,
                 long long ld_table, void* __restrict__ dpos, long long pad_row, long long pos_pad_row) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = H >> 3;
  for (int r = warp; r < rows; r += nwarps) {
    long long srow = ids ? ids[r] : r;
    if (srow < 0 || srow >= vocab) continue;
    const int l = r % tokens_per_seq;
    const long long orow = static_cast<long long>(r / tokens_per_seq) * out_group_stride + l + out_row_off;
    // nn.Embedding(padding_idx=...) rows never receive a gradient (pad_row / pos_pad_row < 0: no such row)
    const bool to_table = dtable && srow != pad_row;
    const bool to_pos = dpos && (pos_row_off + l) != pos_pad_row;
    if (!to_table && !to_pos) continue;
    for (int vi = lane; vi < nvec; vi += 32) {
      float v[8];
      ld8_as_float(dout, dt, orow * ld_dout + vi * 8, v);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        if (to_table) atomic_add_elem2(dtable, dt, srow * ld_table + vi * 8 + j, v[j] * scale, v[j + 1] * scale);
        if (to_pos) atomic_add_elem2(dpos, dt, static_cast<long long>(pos_row_off + l) * H + vi * 8 + j, v[j] * scale, v[j + 1] * scale);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// patch extraction: NCHW pixels -> [B * nP, C * p * p] rows (c, i, j fastest = j), any dtype pair
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
patchify_kernel(int B, int C, int Hh, int Ww, int ph, int pw, const void* __restrict__ px, int in_dt,
                void* __restrict__ out, int out_dt, int vec, long long ld_out) {
  pdl_trigger();
  pdl_wait();
  const int gw = Ww / pw, gh = Hh / ph;
  const int K = C * ph * pw;
  if (ld_out <= 0) ld_out = K;
  const long long total = static_cast<long long>(B) * gh * gw * K;
  if (vec) {
    // 8 consecutive outputs = 8 consecutive pixels of one patch row (pw % 8 == 0): one 16/32-byte load and store
    const long long nvec = total >> 3;
    for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < nvec;
         v += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long i = v << 3;
      const int kk = static_cast<int>(i % K);
      const long long prow = i / K;
      const int s = static_cast<int>(prow % gw);
      const int r = static_cast<int>((prow / gw) % gh);
      const int b = static_cast<int>(prow / (static_cast<long long>(gw) * gh));
      const int j = kk % pw, ii = (kk / pw) % ph, c = kk / (pw * ph);
      const long long src = ((static_cast<long long>(b) * C + c) * Hh + (r * ph + ii)) * Ww + (s * pw + j);
      float x[8];
      ld8_as_float(px, in_dt, src, x);
      st8_from_float(out, out_dt, prow * ld_out + kk, x);
    }
    return;
  }
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kk = static_cast<int>(i % K);
    const long long prow = i / K;
    const int s = static_cast<int>(prow % gw);
    const int r = static_cast<int>((prow / gw) % gh);
    const int b = static_cast<int>(prow / (static_cast<long long>(gw) * gh));
    const int j = kk % pw, ii = (kk / pw) % ph, c = kk / (pw * ph);
    const long long src = ((static_cast<long long>(b) * C + c) * Hh + (r * ph + ii)) * Ww + (s * pw + j);
    st_from_float(out, out_dt, prow * ld_out + kk, ld_as_float(px, in_dt, src));
  }
}

// ------------------------------------------------------------------------------------------
// greedy next-token: argmax over the vocabulary, first index on ties (torch.topk(k=1) rule used
// by models/decoder.py:489-496 and generation_utils.py:179-189). One CTA per row.
// ------------------------------------------------------------------------------------------
// Rows are long (a vocabulary) and few (a decode batch), so a row is split over several CTAs: each finds the (max,
// first index) of its slice and folds it into out[r] with a 64-bit atomicMax on the packed key
// (order-preserving float bits << 32 | ~index): larger value wins, equal values keep the SMALLER index — the
// torch.topk(k=1) rule. out[] must be zero on entry (vy_argmax_rows clears it); a second tiny kernel unpacks.
__device__ __forceinline__ unsigned int float_order_bits(float v) {
  const unsigned int u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void __launch_bounds__(256)
argmax_rows_kernel(int rows, int V, int per, const void* __restrict__ x, long long ld, int dt,
                   unsigned long long* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  const int r = blockIdx.y;
  const int c0 = blockIdx.x * per, c1 = min(V, c0 + per);
  float best = -INFINITY;
  int bi = 0x7fffffff;
  const long long base = static_cast<long long>(r) * ld;
  const bool vec = dt == VY_BF16 && (ld & 7) == 0 && (c0 & 7) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  int c = c0 + threadIdx.x * 8;
  if (vec) {
    for (; c + 8 <= c1; c += blockDim.x * 8) {
      float v[8];
      ld8_as_float(x, dt, base + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v[j] > best) {  // ascending index within a thread: strict > keeps the first maximum
          best = v[j];
          bi = c + j;
        }
    }
    // ragged tail of the slice (at most 7 + misaligned leftovers): thread 0..7 scalar
    const int tail0 = c0 + ((c1 - c0) / 8) * 8;
    if (threadIdx.x < c1 - tail0) {
      const int ct = tail0 + threadIdx.x;
      const float v = ld_as_float(x, dt, base + ct);
      if (v > best || (v == best && ct < bi)) {
        best = v;
        bi = ct;
      }
    }
  } else {
    for (int cc = c0 + threadIdx.x; cc < c1; cc += blockDim.x) {
      const float v = ld_as_float(x, dt, base + cc);
      if (v > best || (v == best && cc < bi)) {
        best = v;
        bi = cc;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    if (v2 > best || (v2 == best && i2 < bi)) {
      best = v2;
      bi = i2;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    s_v[threadIdx.x >> 5] = best;
    s_i[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_v[w] > best || (s_v[w] == best && s_i[w] < bi)) {
        best = s_v[w];
        bi = s_i[w];
      }
    if (bi != 0x7fffffff) {  // a slice of only -inf / NaN contributes nothing (an all -inf / NaN row ends as index 0)
      const unsigned long long key = (static_cast<unsigned long long>(float_order_bits(best)) << 32) |
                                     static_cast<unsigned long long>(0xffffffffu - static_cast<unsigned int>(bi));
      atomicMax(out + r, key);
    }
  }
}
__global__ void __launch_bounds__(256)
argmax_unpack_kernel(int rows, unsigned long long* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const unsigned long long key = out[r];
  const long long idx = key == 0ull ? 0ll : static_cast<long long>(0xffffffffu - static_cast<unsigned int>(key & 0xffffffffull));
  reinterpret_cast<long long*>(out)[r] = idx;
}

// unpack + the greedy loop's bookkeeping in one launch (one CTA, rows <= 1024): tok[r] = argmax index, *pos += 1,
// tokens[r][*pos] = tok[r] — the tail of a captured decode step (decoder.py:489-500: append the token, advance cur_pos)
__global__ void __launch_bounds__(1024)
argmax_advance_kernel(int rows, unsigned long long* __restrict__ out, int* __restrict__ pos, long long* __restrict__ tokens, long long tokens_ld,
                      int tokens_cols) {
  pdl_trigger();
  pdl_wait();
  const int r = threadIdx.x;
  const int np = *pos + 1;
  if (r < rows) {
    const unsigned long long key = out[r];
    const long long idx = key == 0ull ? 0ll : static_cast<long long>(0xffffffffu - static_cast<unsigned int>(key & 0xffffffffull));
    reinterpret_cast<long long*>(out)[r] = idx;
    if (tokens && np >= 0 && np < tokens_cols) tokens[r * tokens_ld + np] = idx;
  }
  __syncthreads();  // every thread has read *pos
  if (r == 0) *pos = np;
}

// ------------------------------------------------------------------------------------------
// column sums (bias gradients): stage 1 partial sums over row chunks, stage 2 final reduce
// ------------------------------------------------------------------------------------------
constexpr int CS_CHUNKS = 64;
// One warp covers 256 consecutive columns (8 per lane, one 16-byte load per row); the 8 warps of a CTA
// take interleaved rows of the CTA's row chunk, 4 rows in flight each, and are combined in smem.
__global__ void __launch_bounds__(256)
colsum_partial_kernel(int R, int Cn, const void* __restrict__ x, long long ld, int dt, float* __restrict__ part, int vec_ok) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][256 + 8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const int per = (R + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(R, r0 + per);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (vec_ok && c0 + 8 <= Cn) {
    int r = r0 + w;
    for (; r + 24 < r1; r += 32) {  // 4 rows in flight per lane (8 measured slower: 22 vs 16 us per call in the step)
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) ld8_as_float(x, dt, static_cast<long long>(r + u * 8) * ld + c0, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[u][j];
    }
    for (; r < r1; r += 8) {
      float v[8];
      ld8_as_float(x, dt, static_cast<long long>(r) * ld + c0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  } else if (c0 < Cn) {
    const int lim = min(8, Cn - c0);
    for (int r = r0 + w; r < r1; r += 8)
      for (int j = 0; j < lim; ++j) acc[j] += ld_as_float(x, dt, static_cast<long long>(r) * ld + c0 + j);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[w][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < Cn) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += red[k][threadIdx.x];
    part[static_cast<long long>(blockIdx.y) * Cn + c] = a;
  }
}
__global__ void __launch_bounds__(128)
colsum_final_kernel(int Cn, int chunks, const float* __restrict__ part, long long pld, void* __restrict__ out, int out_dt, int accumulate,
                    float scale, const float* __restrict__ scale_ptr) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cn) return;
  float a = 0.f;
  for (int k = 0; k < chunks; ++k) a += part[static_cast<long long>(k) * pld + c];
  a *= scale_ptr ? scale * *scale_ptr : scale;
  if (accumulate) a += ld_as_float(out, out_dt, c);
  st_from_float(out, out_dt, c, a);
}

// ------------------------------------------------------------------------------------------
// strided 4-D cast/copy (inner dim contiguous), e.g. fp32 kv-cache prefix -> bf16 operands
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cast4d_kernel(int n0, int n1, int n2, int n3, const void* __restrict__ src, int sdt, long long s0, long long s1,
              long long s2, void* __restrict__ dst, int ddt, long long d0, long long d1, long long d2, int vec) {
  pdl_trigger();
  pdl_wait();
  const int w = vec ? 8 : 1;  // elements per thread along the contiguous inner dim
  const int n3v = n3 / w;
  const long long total = static_cast<long long>(n0) * n1 * n2 * n3v;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i3 = static_cast<int>(i % n3v) * w;
    long long t = i / n3v;
    const int i2 = static_cast<int>(t % n2);
    t /= n2;
    const int i1 = static_cast<int>(t % n1);
    const int i0 = static_cast<int>(t / n1);
    const long long so = i0 * s0 + i1 * s1 + i2 * s2 + i3, dof = i0 * d0 + i1 * d1 + i2 * d2 + i3;
    if (vec) {
      float x[8];
      ld8_as_float(src, sdt, so, x);
      st8_from_float(dst, ddt, dof, x);
    } else {
      st_from_float(dst, ddt, dof, ld_as_float(src, sdt, so));
    }
  }
}

// ------------------------------------------------------------------------------------------
// fused softmax cross-entropy: per row loss (fp32) and, in place, dlogits = (softmax - onehot) *
// grad_scale for rows whose label != ignore_index (0 elsewhere). One CTA per row, two passes.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
xent_kernel(int rows, int V, void* __restrict__ logits, long long ld, int dt, const long long* __restrict__ labels,
            long long ignore_index, const float* __restrict__ grad_scale_ptr, float grad_scale,
            float* __restrict__ loss_rows, int write_grad, int vec_ok) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_m[8], s_s[8];
  const int r = blockIdx.x;
  const long long base = static_cast<long long>(r) * ld;
  const long long lab = labels[r];
  const bool active = lab != ignore_index && lab >= 0 && lab < V;
  const int nvec = vec_ok ? (V >> 3) : 0;  // 8-element (16 B of bf16) chunks; the tail is scalar
  if (!active) {
    if (threadIdx.x == 0 && loss_rows) loss_rows[r] = 0.f;
    if (write_grad) {
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) st8_from_float(logits, dt, base + vi * 8, z);
      for (int c = nvec * 8 + threadIdx.x; c < V; c += blockDim.x) st_from_float(logits, dt, base + c, 0.f);
    }
    return;
  }
  // pass 1: online max / sum (one read of the row)
  float mx = -INFINITY, sum = 0.f;
  for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
    float x[8];
    ld8_as_float(logits, dt, base + vi * 8, x);
    float cm = x[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) cm = fmaxf(cm, x[j]);
    const float nm = fmaxf(mx, cm);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += __expf(x[j] - nm);
    sum = sum * __expf(mx - nm) + acc;
    mx = nm;
  }
  for (int c = nvec * 8 + threadIdx.x; c < V; c += blockDim.x) {
    const float x = ld_as_float(logits, dt, base + c);
    const float nm = fmaxf(mx, x);
    sum = sum * __expf(mx - nm) + __expf(x - nm);
    mx = nm;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, mx, o), s2 = __shfl_xor_sync(0xffffffffu, sum, o);
    const float nm = fmaxf(mx, m2);
    sum = (mx == -INFINITY ? 0.f : sum * __expf(mx - nm)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - nm));
    mx = nm;
  }
  if ((threadIdx.x & 31) == 0) {
    s_m[threadIdx.x >> 5] = mx;
    s_s[threadIdx.x >> 5] = sum;
  }
  __syncthreads();
  mx = s_m[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_m[w]);
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += (s_m[w] == -INFINITY ? 0.f : s_s[w] * __expf(s_m[w] - mx));
  const float lse = mx + __logf(sum);
  if (threadIdx.x == 0 && loss_rows) loss_rows[r] = lse - ld_as_float(logits, dt, base + lab);
  if (write_grad) {
    // pass 2: the row is re-read (L2-resident: a row is <= ~200 KB) and overwritten with the gradient
    const float gs = grad_scale_ptr ? *grad_scale_ptr * grad_scale : grad_scale;
    const float inv = 1.f / sum;
    for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
      float x[8];
      ld8_as_float(logits, dt, base + vi * 8, x);
      const int c0 = vi * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (__expf(x[j] - mx) * inv - (c0 + j == lab ? 1.f : 0.f)) * gs;
      st8_from_float(logits, dt, base + c0, x);
    }
    for (int c = nvec * 8 + threadIdx.x; c < V; c += blockDim.x) {
      const float pr = __expf(ld_as_float(logits, dt, base + c) - mx) * inv;
      st_from_float(logits, dt, base + c, (pr - (c == lab ? 1.f : 0.f)) * gs);
    }
  }
}

// ------------------------------------------------------------------------------------------
// The same loss with the column sums of the written gradient (= the vocabulary projection's bias gradient) taken in
// the same pass, so the [rows, V] gradient is not read again for them (824 MB at the captioner's shape). One persistent
// 1024-thread CTA per SM walks rows b, b + G, ...: thread t owns the 8-column vectors t, t + 1024, ... of every row
//   * the row lives in registers (<= 7 x 16 B per thread), read from HBM exactly once: raw bf16 for the maximum, then
//     overwritten by exp(x - max) as packed fp16 for the last pass (one exponential per element);
//   * per-column fp32 accumulators live in shared memory (V x 4 B <= 224 KB, two float4 planes so that consecutive
//     lanes touch consecutive 16-byte words) and are private to their owning thread: no atomics, fixed order;
//   * rows are asked into L2 two iterations ahead and the next row's registers are loaded in two batches inside the
//     last pass, once the vectors they replace have been consumed.
// The CTA's sums go to part[b][V rounded up to 8]; colsum_final_kernel adds the G partials in index order (deterministic).
// Measured (tools/xent_bench.py, 8256 x 50265): 415 us against 675 us for xent_kernel + colsum; ncu in profiles/.
// ------------------------------------------------------------------------------------------
constexpr int XC_THREADS = 1024, XC_MAXC = 7, XC_HALF = 4;
__device__ __forceinline__ void bf16x8_to_float(const uint4& raw, float (&x)[8]) {
  const unsigned int u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    x[2 * k] = __uint_as_float(u[k] << 16);
    x[2 * k + 1] = __uint_as_float(u[k] & 0xffff0000u);
  }
}
// NFULL: vectors k < NFULL of a thread exist for every thread (V >= NFULL * 8192) and are processed without bounds tests.
template <int NFULL>
__global__ void __launch_bounds__(XC_THREADS, 1)
xent_colsum_kernel(int rows, int V, __nv_bfloat16* logits, long long ld, const long long* __restrict__ labels,
                   long long ignore_index, const float* __restrict__ grad_scale_ptr, float grad_scale,
                   float* __restrict__ loss_rows, float* __restrict__ part) {
  pdl_trigger();
  extern __shared__ float4 xc_acc[];  // [2][nvec]
  __shared__ float s_mx[32], s_sum[32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int nvec = (V + 7) >> 3;  // the row stride covers whole vectors; a last partial vector's lanes >= V are padding
  const int tail = V & 7;
  constexpr float LOG2E = 1.4426950408889634f;
#define XC_HAS(k) ((k) < NFULL || t + (k) * XC_THREADS < nvec)
// The passes unpack the same registers; without this the compiler keeps the unpacked fp32 values of one pass alive for the
// next (56 more registers than there are) and spills them.
#define XC_OPAQUE()                                                                                        \
  _Pragma("unroll") for (int k_ = 0; k_ < XC_MAXC; ++k_)                                                  \
      asm volatile("" : "+r"(raw[k_].x), "+r"(raw[k_].y), "+r"(raw[k_].z), "+r"(raw[k_].w))
  for (int i = t; i < 2 * nvec; i += XC_THREADS) xc_acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  pdl_wait();
  __syncthreads();
  const float gs = grad_scale_ptr ? *grad_scale_ptr * grad_scale : grad_scale;
  uint4 raw[XC_MAXC];
  int r = blockIdx.x;
  // Rows are asked into L2 two iterations ahead (one 128-byte line per thread, no register and no scoreboard involved), so
  // the register loads of the next row, issued in two batches inside the last pass, find them there (~0.5 us instead of a
  // 2 us DRAM round, measured in tools/microbench) and anything that has to wait behind them waits that long at most.
  auto ask_l2 = [&](int rr) {
    if (rr < rows) {
      const char* b = reinterpret_cast<const char*>(logits + static_cast<long long>(rr) * ld);
      if (t * 8 < nvec) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + t * 128));  // nvec <= 7168: at most 896 lines
    }
  };
  if (r < rows) {
    const uint4* src = reinterpret_cast<const uint4*>(logits + static_cast<long long>(r) * ld);
#pragma unroll
    for (int k = 0; k < XC_MAXC; ++k)
      if (XC_HAS(k)) raw[k] = src[t + k * XC_THREADS];
    ask_l2(r + gridDim.x);
  }
  long long lab_nxt = r < rows ? labels[r] : 0;  // read one row ahead: the row's first branch depends on it
  for (; r < rows; r += gridDim.x) {
    __nv_bfloat16* row = logits + static_cast<long long>(r) * ld;
    const int rn = r + gridDim.x;
    const uint4* nxt = reinterpret_cast<const uint4*>(logits + static_cast<long long>(rn) * ld);
    const bool more = rn < rows;
    const long long lab = lab_nxt;
    if (more) lab_nxt = labels[rn];
    ask_l2(rn + gridDim.x);
    if (!(lab != ignore_index && lab >= 0 && lab < V)) {  // CTA-uniform
      if (t == 0 && loss_rows) loss_rows[r] = 0.f;
#pragma unroll
      for (int k = 0; k < XC_MAXC; ++k)
        if (XC_HAS(k)) {
          const int vi = t + k * XC_THREADS;
          reinterpret_cast<uint4*>(row)[vi] = make_uint4(0u, 0u, 0u, 0u);
          if (more) raw[k] = nxt[vi];
        }
      continue;
    }
    // the thread that holds the label's column reads its logit now and corrects that one element after the last pass
    // (softmax - onehot), so the passes themselves carry no per-element label test
    const int lab_vi = static_cast<int>(lab >> 3);
    const bool own = t == (lab_vi & (XC_THREADS - 1));
    float x_lab = 0.f;
    if (own) x_lab = __bfloat162float(row[lab]);
    if (tail) {  // padding lanes read as -inf: they add 0 to the sums and receive a zero gradient
#pragma unroll
      for (int k = 0; k < XC_MAXC; ++k)
        if (t + k * XC_THREADS == nvec - 1) {
          unsigned int* u = reinterpret_cast<unsigned int*>(&raw[k]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (2 * q >= tail) u[q] = (u[q] & 0xffff0000u) | 0xff80u;
            if (2 * q + 1 >= tail) u[q] = (u[q] & 0x0000ffffu) | 0xff800000u;
          }
        }
    }
    // pass 1: row maximum on the packed values
    unsigned int pm = 0xff80ff80u;
#pragma unroll
    for (int k = 0; k < XC_MAXC; ++k)
      if (XC_HAS(k)) {
        const unsigned int u[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) asm("max.bf16x2 %0, %0, %1;" : "+r"(pm) : "r"(u[q]));
      }
    float mx = fmaxf(__uint_as_float(pm << 16), __uint_as_float(pm & 0xffff0000u));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_mx[w] = mx;
    __syncthreads();
    mx = s_mx[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // pass 2: e = exp(x - max) * 2^14, summed in fp32 and left in the row's registers as packed fp16 (11 significant bits,
    // normal down to 2^-28 of the row maximum's weight; the bf16 result below keeps 8), so each element costs ONE exponential
    // — the kernel is bound by the MUFU / shared-memory instruction queue, not by HBM
    const float mb = fmaf(-mx, LOG2E, 14.f);
    XC_OPAQUE();
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < XC_MAXC; ++k)
      if (XC_HAS(k)) {
        float x[8];
        bf16x8_to_float(raw[k], x);
        float a = 0.f;
        unsigned int e2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float e0 = vy_ex2_approx(fmaf(x[2 * q], LOG2E, mb)), e1 = vy_ex2_approx(fmaf(x[2 * q + 1], LOG2E, mb));
          a += e0 + e1;
          const __half2 h = __floats2half2_rn(e0, e1);
          e2[q] = *reinterpret_cast<const unsigned int*>(&h);
        }
        sum += a;
        raw[k] = make_uint4(e2[0], e2[1], e2[2], e2[3]);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_sum[w] = sum;
    __syncthreads();
    sum = s_sum[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    // pass 3: softmax * scale written over the logits, column sums, next row requested in two batches
    const float sc = gs / sum;  // sum carries the 2^14 of e
#pragma unroll
    for (int k = 0; k < XC_MAXC; ++k) {
      if (XC_HAS(k)) {
        const int vi = t + k * XC_THREADS;
        const unsigned int u[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
        float4 a0 = xc_acc[vi], a1 = xc_acc[nvec + vi];  // both accumulator words requested before the arithmetic
        float x[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[q]));
          x[2 * q] = f.x * sc;
          x[2 * q + 1] = f.y * sc;
        }
        a0.x += x[0]; a0.y += x[1]; a0.z += x[2]; a0.w += x[3];
        a1.x += x[4]; a1.y += x[5]; a1.z += x[6]; a1.w += x[7];
        xc_acc[vi] = a0;
        xc_acc[nvec + vi] = a1;
        unsigned int o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __nv_bfloat162 pk = __floats2bfloat162_rn(x[2 * q], x[2 * q + 1]);
          o[q] = *reinterpret_cast<const unsigned int*>(&pk);
        }
        reinterpret_cast<uint4*>(row)[vi] = make_uint4(o[0], o[1], o[2], o[3]);
      }
      if (k == XC_HALF - 1 && more) {  // the first vectors are consumed: their registers take the next row's
#pragma unroll
        for (int kk = 0; kk < XC_HALF; ++kk)
          if (XC_HAS(kk)) raw[kk] = nxt[t + kk * XC_THREADS];
      }
    }
    if (more) {
#pragma unroll
      for (int kk = XC_HALF; kk < XC_MAXC; ++kk)
        if (XC_HAS(kk)) raw[kk] = nxt[t + kk * XC_THREADS];
    }
    if (own) {  // this thread wrote the label's vector above: same-thread order makes the corrections land after it
      const float g = vy_ex2_approx(fmaf(x_lab, LOG2E, mb)) * sc - gs;
      row[lab] = __float2bfloat16_rn(g);
      const int lab_j = static_cast<int>(lab & 7);
      reinterpret_cast<float*>(xc_acc)[(static_cast<long long>(lab_j >> 2) * nvec + lab_vi) * 4 + (lab_j & 3)] -= gs;
      if (loss_rows) loss_rows[r] = mx + (__logf(sum) - 14.f * 0.6931471805599453f) - x_lab;
    }
  }
  float4* dst = reinterpret_cast<float4*>(part + static_cast<long long>(blockIdx.x) * nvec * 8);  // rows of nvec * 8 floats
#pragma unroll
  for (int k = 0; k < XC_MAXC; ++k)
    if (XC_HAS(k)) {
      const int vi = t + k * XC_THREADS;
      dst[2 * vi] = xc_acc[vi];
      dst[2 * vi + 1] = xc_acc[nvec + vi];
    }
#undef XC_HAS
#undef XC_OPAQUE
}

// ------------------------------------------------------------------------------------------
// sum of squares of a flat gradient buffer -> one fp32 (atomic per CTA), then AdamW that reads the
// clip coefficient from device memory (no host sync): g' = g * min(1, max_norm / (norm + 1e-6)).
// ------------------------------------------------------------------------------------------
// Deterministic: every block stores its partial sum, the block that finishes last adds the partials up in index order
// and updates *out once — two data-parallel replicas that hold identical gradients therefore compute bit-identical clip
// coefficients and stay in lock step (a float atomicAdd per block would make the order, hence the rounding, vary).
// The partial / ticket buffers are shared by all launches: vy_sqnorm calls must not overlap on different streams.
constexpr int SQNORM_MAX_BLOCKS = 4096;
static __device__ float sqnorm_partials[SQNORM_MAX_BLOCKS];
static __device__ unsigned int sqnorm_ticket;

__global__ void __launch_bounds__(256)
sqnorm_kernel(long long n, const void* __restrict__ g, int dt, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_red[8];
  __shared__ bool s_last;
  float a = 0.f;
  const long long nvec = n >> 3;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long vi = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  {  // four 16-byte loads in flight per thread (ncu: one at a time left the kernel latency-bound at 3.9 of 6.5 TB/s)
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    for (; vi + 3 * stride < nvec; vi += 4 * stride) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) ld8_as_float(g, dt, (vi + u * stride) * 8, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) a4[u] += v[u][j] * v[u][j];
    }
    a = (a4[0] + a4[1]) + (a4[2] + a4[3]);
  }
  for (; vi < nvec; vi += stride) {
    float v[8];
    ld8_as_float(g, dt, vi * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) a += v[j] * v[j];
  }
  if (blockIdx.x == 0)
    for (long long i = (nvec << 3) + threadIdx.x; i < n; i += blockDim.x) {
      const float v = ld_as_float(g, dt, i);
      a += v * v;
    }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w];
    sqnorm_partials[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(&sqnorm_ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {  // fixed-order tree over the partials: lane-strided sums, then the warp butterfly
    __threadfence();
    float t = 0.f;
    if (threadIdx.x < 32) {
      for (int i = threadIdx.x; i < static_cast<int>(gridDim.x); i += 32) t += *const_cast<volatile float*>(&sqnorm_partials[i]);
      t = warp_sum(t);
      if (threadIdx.x == 0) {
        *out += t;
        sqnorm_ticket = 0;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
adamw_kernel(long long n, void* __restrict__ p, int p_dt, const void* __restrict__ g, int g_dt,
             float* __restrict__ m, float* __restrict__ v, float* __restrict__ master, float lr, float beta1,
             float beta2, float eps, float wd, float bc1, float bc2, const int* __restrict__ step_ptr,
             const float* __restrict__ sqnorm, float max_norm, float grad_div) {
  pdl_trigger();
  pdl_wait();
  float clip = 1.f / grad_div;
  if (sqnorm && max_norm > 0.f) {
    const float norm = sqrtf(*sqnorm) / grad_div;
    clip *= fminf(1.f, max_norm / (norm + 1e-6f));
  }
  if (step_ptr) {  // device-side step counter: the launch can sit in a replayed CUDA graph
    const float t = static_cast<float>(*step_ptr);
    bc1 = 1.f - powf(beta1, t);
    bc2 = 1.f - powf(beta2, t);
  }
  const float inv_bc1 = 1.f / bc1, inv_bc2 = 1.f / bc2, decay = 1.f - lr * wd;
  // n is padded to a multiple of 8 by the flat-buffer layout; 8 elements (16 B of bf16) per thread
  const long long nvec = n >> 3;
  for (long long vi8 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; vi8 < nvec;
       vi8 += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = vi8 * 8;
    float gi[8], w[8], mi[8], vv[8];
    ld8_as_float(g, g_dt, i, gi);
    if (master) ld8_as_float(master, VY_F32, i, w);
    else ld8_as_float(p, p_dt, i, w);
    ld8_as_float(m, VY_F32, i, mi);
    ld8_as_float(v, VY_F32, i, vv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gj = gi[j] * clip;
      mi[j] = beta1 * mi[j] + (1.f - beta1) * gj;
      vv[j] = beta2 * vv[j] + (1.f - beta2) * gj * gj;
      w[j] = w[j] * decay - lr * (mi[j] * inv_bc1) / (sqrtf(vv[j] * inv_bc2) + eps);
    }
    st8_from_float(m, VY_F32, i, mi);
    st8_from_float(v, VY_F32, i, vv);
    if (master) st8_from_float(master, VY_F32, i, w);
    st8_from_float(p, p_dt, i, w);
  }
  if (blockIdx.x == 0) {
    for (long long i = (nvec << 3) + threadIdx.x; i < n; i += blockDim.x) {
      const float gj = ld_as_float(g, g_dt, i) * clip;
      float w = master ? master[i] : ld_as_float(p, p_dt, i);
      const float mj = beta1 * m[i] + (1.f - beta1) * gj;
      const float vj = beta2 * v[i] + (1.f - beta2) * gj * gj;
      m[i] = mj;
      v[i] = vj;
      w = w * decay - lr * (mj * inv_bc1) / (sqrtf(vj * inv_bc2) + eps);
      if (master) master[i] = w;
      st_from_float(p, p_dt, i, w);
    }
  }
}

// ---- image-slot merge (the notebook-II captioner): rows whose slot index is >= 0 come from `b`, the others from `a` ----
// fwd: out[r] = slot[r] >= 0 ? b[slot[r]] : a[r].   bwd: da[r] = slot[r] >= 0 ? 0 : dout[r];  db[slot[r]] = dout[r].
// One 16-byte vector per thread; rows are H elements (H % 8 == 0 for bf16, % 4 for fp32), all buffers row-contiguous.
__global__ void __launch_bounds__(256)
slot_merge_kernel(int rows, int vec_per_row, const uint4* __restrict__ a, const uint4* __restrict__ b, const int* __restrict__ slot,
                  uint4* __restrict__ out, uint4* __restrict__ da, uint4* __restrict__ db, int backward, int n_b) {
  pdl_trigger();
  pdl_wait();
  const long long total = static_cast<long long>(rows) * vec_per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / vec_per_row), c = static_cast<int>(i - static_cast<long long>(r) * vec_per_row);
    int sl = slot[r];
    if (sl >= n_b) sl = -1;  // more <image> positions than source rows (the reference's masked_scatter raises; the host checks): never read past b
    if (!backward) {
      out[i] = sl >= 0 ? b[static_cast<long long>(sl) * vec_per_row + c] : a[i];
    } else {
      const uint4 g = a[i];  // (a = dout)
      if (da) da[i] = sl >= 0 ? make_uint4(0, 0, 0, 0) : g;
      if (db && sl >= 0) db[static_cast<long long>(sl) * vec_per_row + c] = g;
    }
  }
}

// stand-alone half-split RoPE: one warp per (b, h, l) row, lane j (+ 32, ...) <-> pair (j, j + D / 2)
__global__ void __launch_bounds__(256)
rope_kernel(int B, int H, int S, int D, const void* __restrict__ x, long long xsb, long long xsh, long long xsl, int dt,
            const float* __restrict__ cs, const float* __restrict__ sn, int pos0, int inverse, void* __restrict__ out,
            long long osb, long long osh, long long osl, const int* __restrict__ pos_ptr, int out_follows_pos, int copy_only) {
  pdl_trigger();
  pdl_wait();
  const int pos_dev = pos_ptr ? *pos_ptr : 0;  // device-side position: one captured graph serves every decode step
  pos0 += pos_dev;
  const long long out_shift = out_follows_pos ? static_cast<long long>(pos_dev) * osl : 0;
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long rows = static_cast<long long>(B) * H * S;
  for (long long r = warp; r < rows; r += nwarps) {
    const int l = static_cast<int>(r % S);
    const int h = static_cast<int>((r / S) % H);
    const int b = static_cast<int>(r / (static_cast<long long>(S) * H));
    const long long xi = b * xsb + h * xsh + l * xsl, oi = b * osb + h * osh + l * osl + out_shift;
    const int half = D >> 1;
    if (copy_only) {  // (values: appended to the cache unrotated)
      for (int j = lane; j < D; j += 32) st_from_float(out, dt, oi + j, ld_as_float(x, dt, xi + j));
      continue;
    }
    for (int j = lane; j < half; j += 32) {
      const float a = ld_as_float(x, dt, xi + j), bb = ld_as_float(x, dt, xi + j + half);
      const float c = cs[static_cast<long long>(pos0 + l) * half + j];
      float s = sn[static_cast<long long>(pos0 + l) * half + j];
      if (inverse) s = -s;
      st_from_float(out, dt, oi + j, a * c - bb * s);
      st_from_float(out, dt, oi + j + half, bb * c + a * s);
    }
  }
}

// RoPE + kv-cache append of a packed q|k|v projection in ONE launch (PaliGemma-scale decoder: three launches per layer before):
// head h of the [B, n_q + 2 n_kv, S, D] view is rotated in place (query), rotated into the key cache, or copied into the value
// cache at slot slot0 (+ *pos_ptr) + l. One warp per (b, h, l) row.
__global__ void __launch_bounds__(256)
rope_append_kernel(int B, int n_q, int n_kv, int S, int D, void* __restrict__ qkv, long long sb, long long sh, long long sl, int dt,
                   const float* __restrict__ cs, const float* __restrict__ sn, int pos0, int slot0, const int* __restrict__ pos_ptr,
                   void* __restrict__ kc, void* __restrict__ vc, long long c_sb, long long c_sh, long long c_sl) {
  pdl_trigger();
  pdl_wait();
  const int pos_dev = pos_ptr ? *pos_ptr : 0;
  pos0 += pos_dev;
  slot0 += pos_dev;
  const int Hh = n_q + 2 * n_kv;
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long rows = static_cast<long long>(B) * Hh * S;
  const int half = D >> 1;
  for (long long r = warp; r < rows; r += nwarps) {
    const int l = static_cast<int>(r % S);
    const int h = static_cast<int>((r / S) % Hh);
    const int b = static_cast<int>(r / (static_cast<long long>(S) * Hh));
    const long long xi = b * sb + h * sh + l * sl;
    if (h >= n_q + n_kv) {  // value: plain copy into the cache
      const long long oi = b * c_sb + (h - n_q - n_kv) * c_sh + static_cast<long long>(slot0 + l) * c_sl;
      for (int j = lane; j < D; j += 32) st_from_float(vc, dt, oi + j, ld_as_float(qkv, dt, xi + j));
      continue;
    }
    void* out = h < n_q ? qkv : kc;
    const long long oi = h < n_q ? xi : b * c_sb + (h - n_q) * c_sh + static_cast<long long>(slot0 + l) * c_sl;
    for (int j = lane; j < half; j += 32) {
      const float a = ld_as_float(qkv, dt, xi + j), bb = ld_as_float(qkv, dt, xi + j + half);
      const float c = cs[static_cast<long long>(pos0 + l) * half + j], s = sn[static_cast<long long>(pos0 + l) * half + j];
      st_from_float(out, dt, oi + j, a * c - bb * s);
      st_from_float(out, dt, oi + j + half, bb * c + a * s);
    }
  }
}

__global__ void __launch_bounds__(256)
act_bwd_kernel(long long n, const void* __restrict__ dy, const void* __restrict__ z, int dt, int act, void* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long nvec = n >> 3;
  for (long long vi = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; vi < nvec;
       vi += static_cast<long long>(gridDim.x) * blockDim.x) {
    float g[8], zz[8];
    ld8_as_float(dy, dt, vi * 8, g);
    ld8_as_float(z, dt, vi * 8, zz);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= (act == VY_ACT_GELU_TANH ? dgelu_tanh(zz[j]) : dgelu_erf(zz[j]));
    st8_from_float(out, dt, vi * 8, g);
  }
}

// dz of the gated MLP: 8 outputs of dh <-> 16 interleaved (gate, up) pre-activations
__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(long long nvec, const void* __restrict__ dh, const void* __restrict__ z, int dt, void* __restrict__ dz) {
  pdl_trigger();
  pdl_wait();
  for (long long vi = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; vi < nvec;
       vi += static_cast<long long>(gridDim.x) * blockDim.x) {
    float d[8], lo[8], hi[8];
    ld8_as_float(dh, dt, vi * 8, d);
    ld8_as_float(z, dt, vi * 16, lo);
    ld8_as_float(z, dt, vi * 16 + 8, hi);
    float x[16], o[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) { x[j] = lo[j]; x[8 + j] = hi[j]; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gte = x[2 * j], up = x[2 * j + 1];
      const float sg = 1.f / (1.f + __expf(-gte));
      o[2 * j] = d[j] * up * sg * (1.f + gte * (1.f - sg));  // silu'(g) = sigma(g) (1 + g (1 - sigma(g)))
      o[2 * j + 1] = d[j] * gte * sg;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { lo[j] = o[j]; hi[j] = o[8 + j]; }
    st8_from_float(dz, dt, vi * 16, lo);
    st8_from_float(dz, dt, vi * 16 + 8, hi);
  }
}

// x[i] *= *scale; a unit scale (what loss.backward() passes) costs one scalar load per thread and no traffic
__global__ void __launch_bounds__(256)
scale_by_ptr_kernel(long long n, void* __restrict__ x, int dt, const float* __restrict__ scale) {
  pdl_trigger();
  pdl_wait();
  const float sc = *scale;
  if (sc == 1.0f) return;
  const long long nvec = n >> 3;
  for (long long vi = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; vi < nvec;
       vi += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v[8];
    ld8_as_float(x, dt, vi * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= sc;
    st8_from_float(x, dt, vi * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const long long i = (nvec << 3) + threadIdx.x;
    st_from_float(x, dt, i, ld_as_float(x, dt, i) * sc);
  }
}

static int ew_grid(long long work_items, int per_block) {
  long long blocks = (work_items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace vy

using namespace vy;

extern "C" int vy_embed_fwd(const VyEmbed* p) {
  VY_CHECK_ARG(p != nullptr, "vy_embed_fwd: null params");
  VY_NEED_DEVICE("vy_embed_fwd");
  VY_CHECK_ARG(p->rows > 0 && p->H > 0 && p->H % 8 == 0, "vy_embed_fwd: bad shape rows=%d H=%d", p->rows, p->H);
  VY_CHECK_ARG(p->src && p->out && dtype_ok(p->dtype), "vy_embed_fwd: null pointer / bad dtype");
  VY_CHECK_ARG(aligned16(p->src) && aligned16(p->out) && aligned16(p->pos) && (p->ld_src * (long long)dtype_size(p->dtype)) % 16 == 0 &&
                   (p->ld_out * (long long)dtype_size(p->dtype)) % 16 == 0,
               "vy_embed_fwd: 16-byte alignment required");
  const int tps = p->tokens_per_seq > 0 ? p->tokens_per_seq : p->rows;
  const int ogs = p->out_group_stride > 0 ? p->out_group_stride : tps;
  VY_CUDA_OK(launch_kernel(embed_fwd_kernel, dim3(ew_grid(p->rows, 8)), dim3(256), 0, static_cast<cudaStream_t>(p->stream), 
      p->rows, p->H, reinterpret_cast<const long long*>(p->ids), p->src, p->ld_src, p->dtype, p->vocab, tps, ogs,
      p->out_row_off, p->pos, p->pos_row_off, p->out_scale == 0.f ? 1.f : p->out_scale, p->out, p->ld_out));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_embed_bwd(const VyEmbed* p) {
  VY_CHECK_ARG(p != nullptr, "vy_embed_bwd: null params");
  VY_NEED_DEVICE("vy_embed_bwd");
  VY_CHECK_ARG(p->rows > 0 && p->H > 0 && p->H % 8 == 0, "vy_embed_bwd: bad shape");
  VY_CHECK_ARG(p->dout && (p->dtable || p->dpos) && dtype_ok(p->dtype), "vy_embed_bwd: null pointer / bad dtype");
  const int tps = p->tokens_per_seq > 0 ? p->tokens_per_seq : p->rows;
  const int ogs = p->out_group_stride > 0 ? p->out_group_stride : tps;
  VY_CUDA_OK(launch_kernel(embed_bwd_kernel, dim3(ew_grid(p->rows, 8)), dim3(256), 0, static_cast<cudaStream_t>(p->stream), 
      p->rows, p->H, reinterpret_cast<const long long*>(p->ids), p->vocab, tps, ogs, p->out_row_off, p->pos_row_off,
      p->out_scale == 0.f ? 1.f : p->out_scale, p->dout, p->ld_out, p->dtype, p->dtable, p->ld_src, p->dpos,
      static_cast<long long>(p->padding_idx_plus1) - 1, static_cast<long long>(p->pos_padding_idx_plus1) - 1));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_patchify(const VyPatchify* p) {
  VY_CHECK_ARG(p != nullptr, "vy_patchify: null params");
  VY_NEED_DEVICE("vy_patchify");
  VY_CHECK_ARG(p->B > 0 && p->C > 0 && p->patch_h > 0 && p->patch_w > 0 && p->H % p->patch_h == 0 && p->W % p->patch_w == 0,
               "vy_patchify: image dimensions must be divisible by the patch size");
  VY_CHECK_ARG(p->pixels && p->out && dtype_ok(p->in_dtype) && dtype_ok(p->out_dtype), "vy_patchify: null pointer / bad dtype");
  const long long total = static_cast<long long>(p->B) * p->C * p->H * p->W;
  const long long Kp = static_cast<long long>(p->C) * p->patch_h * p->patch_w;
  VY_CHECK_ARG(p->ld_out == 0 || p->ld_out >= Kp, "vy_patchify: ld_out %lld is smaller than a patch row (%lld)", (long long)p->ld_out, Kp);
  const int vec = (p->patch_w % 8 == 0 && p->W % 8 == 0 && aligned16(p->pixels) && aligned16(p->out) &&
                   (p->ld_out == 0 || (p->ld_out * static_cast<long long>(dtype_size(p->out_dtype))) % 16 == 0)) ? 1 : 0;
  VY_CUDA_OK(launch_kernel(patchify_kernel, dim3(ew_grid(vec ? total / 8 : total, 1024)), dim3(256), 0,
                           static_cast<cudaStream_t>(p->stream), p->B, p->C, p->H, p->W, p->patch_h, p->patch_w, p->pixels,
                           p->in_dtype, p->out, p->out_dtype, vec, static_cast<long long>(p->ld_out)));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_argmax_rows(int rows, int V, const void* x, int64_t ld, int dtype, int64_t* out, void* stream) {
  VY_NEED_DEVICE("vy_argmax_rows");
  VY_CHECK_ARG(rows > 0 && V > 0 && x && out && dtype_ok(dtype), "vy_argmax_rows: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int slices = (2 * num_sms() + rows - 1) / rows;  // ~2 CTAs per SM over all rows
  const int max_slices = (V + 2047) / 2048;         // at least 2048 columns per slice
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  int per = (V + slices - 1) / slices;
  per = (per + 7) / 8 * 8;  // slices start on 16-byte boundaries of a bf16 row
  slices = (V + per - 1) / per;
  VY_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(int64_t) * rows, st));
  VY_CUDA_OK(launch_kernel(argmax_rows_kernel, dim3(slices, rows), dim3(256), 0, st, rows, V, per, x, ld, dtype,
                           reinterpret_cast<unsigned long long*>(out)));
  VY_LAUNCH_OK();
  VY_CUDA_OK(launch_kernel(argmax_unpack_kernel, dim3((rows + 255) / 256), dim3(256), 0, st, rows,
                           reinterpret_cast<unsigned long long*>(out)));
  VY_LAUNCH_OK();
  count_launch(2);
  return VY_OK;
}

extern "C" int vy_argmax_advance(int rows, int V, const void* x, int64_t ld, int dtype, int64_t* out, int32_t* pos, int64_t* tokens, int64_t tokens_ld,
                                 int tokens_cols, void* stream) {
  VY_NEED_DEVICE("vy_argmax_advance");
  VY_CHECK_ARG(rows > 0 && rows <= 1024 && V > 0 && x && out && pos && dtype_ok(dtype) && (!tokens || (tokens_ld >= tokens_cols && tokens_cols > 0)),
               "vy_argmax_advance: bad arguments (rows <= 1024)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int slices = (2 * num_sms() + rows - 1) / rows;  // ~2 CTAs per SM over all rows
  const int max_slices = (V + 2047) / 2048;         // at least 2048 columns per slice
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  int per = (V + slices - 1) / slices;
  per = (per + 7) / 8 * 8;  // slices start on 16-byte boundaries of a bf16 row
  slices = (V + per - 1) / per;
  VY_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(int64_t) * rows, st));
  VY_CUDA_OK(launch_kernel(argmax_rows_kernel, dim3(slices, rows), dim3(256), 0, st, rows, V, per, x, ld, dtype,
                           reinterpret_cast<unsigned long long*>(out)));
  VY_LAUNCH_OK();
  VY_CUDA_OK(launch_kernel(argmax_advance_kernel, dim3(1), dim3(1024), 0, st, rows, reinterpret_cast<unsigned long long*>(out), static_cast<int*>(pos),
                           reinterpret_cast<long long*>(tokens), static_cast<long long>(tokens_ld), tokens_cols));
  VY_LAUNCH_OK();
  count_launch(2);
  return VY_OK;
}

extern "C" int vy_colsum(int rows, int cols, const void* x, int64_t ld, int dtype, void* out, int out_dtype,
                         int accumulate, float scale, float* workspace, void* stream) {
  VY_NEED_DEVICE("vy_colsum");
  VY_CHECK_ARG(rows > 0 && cols > 0 && x && out && workspace && dtype_ok(dtype) && dtype_ok(out_dtype), "vy_colsum: bad arguments");
  const int colgroups = (cols + 255) / 256;
  int chunks = (4 * num_sms() + colgroups - 1) / colgroups;  // ~4 CTAs per SM in total
  if (chunks > CS_CHUNKS) chunks = CS_CHUNKS;
  if (chunks > (rows + 31) / 32) chunks = (rows + 31) / 32;
  if (chunks < 1) chunks = 1;
  dim3 grid(colgroups, chunks);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vec_ok = (aligned16(x) && (ld * static_cast<long long>(dtype_size(dtype))) % 16 == 0) ? 1 : 0;
  VY_CUDA_OK(launch_kernel(colsum_partial_kernel, dim3(grid), dim3(256), 0, st, rows, cols, x, ld, dtype, workspace, vec_ok));
  VY_LAUNCH_OK();
  VY_CUDA_OK(launch_kernel(colsum_final_kernel, dim3((cols + 127) / 128), dim3(128), 0, st, cols, chunks, workspace, static_cast<long long>(cols), out, out_dtype, accumulate,
                                                           scale == 0.f ? 1.f : scale, static_cast<const float*>(nullptr)));
  VY_LAUNCH_OK();
  count_launch(2);
  return VY_OK;
}

extern "C" int vy_colsum_workspace_floats(int cols) { return CS_CHUNKS * cols; }

extern "C" int vy_xent_colsum_chunks(int rows, int V, int dtype) {
  if (rows <= 0 || V <= 0 || dtype != VY_BF16 || (V + 7) / 8 > XC_MAXC * XC_THREADS) return 0;
  const int sms = num_sms();
  return rows < sms ? rows : sms;
}

extern "C" int vy_colsum_finish(int cols, int chunks, const float* part, int64_t part_ld, void* out, int out_dtype, int accumulate,
                                float scale, const float* scale_ptr, void* stream) {
  VY_NEED_DEVICE("vy_colsum_finish");
  VY_CHECK_ARG(cols > 0 && chunks > 0 && part && part_ld >= cols && out && dtype_ok(out_dtype), "vy_colsum_finish: bad arguments");
  VY_CUDA_OK(launch_kernel(colsum_final_kernel, dim3((cols + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), cols, chunks, part,
                           static_cast<long long>(part_ld), out, out_dtype, accumulate, scale == 0.f ? 1.f : scale, scale_ptr));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_cast4d(const VyCast4d* p) {
  VY_CHECK_ARG(p != nullptr, "vy_cast4d: null params");
  VY_NEED_DEVICE("vy_cast4d");
  VY_CHECK_ARG(p->n0 > 0 && p->n1 > 0 && p->n2 > 0 && p->n3 > 0 && p->src && p->dst && dtype_ok(p->src_dtype) && dtype_ok(p->dst_dtype),
               "vy_cast4d: bad arguments");
  const long long total = static_cast<long long>(p->n0) * p->n1 * p->n2 * p->n3;
  auto ok8 = [](const void* ptr, long long a, long long b_, long long c, int dt) {
    const long long es = static_cast<long long>(dtype_size(dt));
    return aligned16(ptr) && (a * es) % 16 == 0 && (b_ * es) % 16 == 0 && (c * es) % 16 == 0;
  };
  const int vec = (p->n3 % 8 == 0 && ok8(p->src, p->s0, p->s1, p->s2, p->src_dtype) && ok8(p->dst, p->d0, p->d1, p->d2, p->dst_dtype)) ? 1 : 0;
  VY_CUDA_OK(launch_kernel(cast4d_kernel, dim3(ew_grid(vec ? total / 8 : total, 1024)), dim3(256), 0,
                           static_cast<cudaStream_t>(p->stream), p->n0, p->n1, p->n2, p->n3, p->src, p->src_dtype, p->s0, p->s1,
                           p->s2, p->dst, p->dst_dtype, p->d0, p->d1, p->d2, vec));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_softmax_xent(const VyXent* p) {
  VY_CHECK_ARG(p != nullptr, "vy_softmax_xent: null params");
  VY_NEED_DEVICE("vy_softmax_xent");
  VY_CHECK_ARG(p->rows > 0 && p->V > 0 && p->logits && p->labels && dtype_ok(p->dtype), "vy_softmax_xent: bad arguments");
  if (p->colsum_part) {  // fused column sums: one persistent CTA per SM (see xent_colsum_kernel)
    const int chunks = vy_xent_colsum_chunks(p->rows, p->V, p->dtype);
    const int nvec = (p->V + 7) / 8;
    VY_CHECK_ARG(chunks > 0 && p->write_grad && aligned16(p->logits) && p->ld % 8 == 0 && p->ld >= nvec * 8 && aligned16(p->colsum_part),
                 "vy_softmax_xent: colsum_part needs write_grad, bf16 logits whose row stride is a multiple of 8 elements covering "
                 "V rounded up to 8, and vy_xent_colsum_chunks() > 0");
    const size_t smem = static_cast<size_t>(nvec) * 32;
    auto go = [&](auto kern) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, XC_MAXC * XC_THREADS * 32);
      if (e != cudaSuccess) return e;
      return launch_kernel(kern, dim3(chunks), dim3(XC_THREADS), smem, static_cast<cudaStream_t>(p->stream), p->rows, p->V,
                           static_cast<__nv_bfloat16*>(p->logits), static_cast<long long>(p->ld), reinterpret_cast<const long long*>(p->labels),
                           static_cast<long long>(p->ignore_index), p->grad_scale_ptr, p->grad_scale, p->loss_rows, p->colsum_part);
    };
    const int nfull = nvec / XC_THREADS;  // vectors every thread owns
    VY_CUDA_OK(nfull >= 6 ? go(xent_colsum_kernel<6>) : nfull >= 3 ? go(xent_colsum_kernel<3>) : go(xent_colsum_kernel<0>));
    VY_LAUNCH_OK();
    count_launch();
    return VY_OK;
  }
  VY_CUDA_OK(launch_kernel(xent_kernel, dim3(p->rows), dim3(256), 0, static_cast<cudaStream_t>(p->stream), 
      p->rows, p->V, p->logits, p->ld, p->dtype, reinterpret_cast<const long long*>(p->labels), p->ignore_index,
      p->grad_scale_ptr, p->grad_scale, p->loss_rows, p->write_grad,
      (aligned16(p->logits) && (p->ld * static_cast<long long>(dtype_size(p->dtype))) % 16 == 0) ? 1 : 0));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_swiglu_bwd(int64_t rows, int32_t inter, const void* dh, const void* z, int dtype, void* dz, void* stream) {
  VY_NEED_DEVICE("vy_swiglu_bwd");
  VY_CHECK_ARG(rows > 0 && inter > 0 && inter % 8 == 0 && dh && z && dz && dtype_ok(dtype) && aligned16(dh) && aligned16(z) && aligned16(dz),
               "vy_swiglu_bwd: bad arguments (intermediate size must be a multiple of 8, pointers 16-byte aligned)");
  const long long nvec = rows * (inter / 8);
  VY_CUDA_OK(launch_kernel(swiglu_bwd_kernel, dim3(ew_grid(nvec, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), nvec, dh, z, dtype, dz));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_sqnorm(int64_t n, const void* g, int dtype, float* out, void* stream) {
  VY_NEED_DEVICE("vy_sqnorm");
  VY_CHECK_ARG(n > 0 && g && out && dtype_ok(dtype) && aligned16(g), "vy_sqnorm: bad arguments");
  VY_CUDA_OK(launch_kernel(sqnorm_kernel, dim3(min(ew_grid(n, 8 * 256 * 4), SQNORM_MAX_BLOCKS)), dim3(256), 0, static_cast<cudaStream_t>(stream), n, g, dtype, out));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_adamw(const VyAdamW* p) {
  VY_CHECK_ARG(p != nullptr, "vy_adamw: null params");
  VY_NEED_DEVICE("vy_adamw");
  VY_CHECK_ARG(p->n > 0 && p->param && p->grad && p->exp_avg && p->exp_avg_sq && dtype_ok(p->param_dtype) && dtype_ok(p->grad_dtype),
               "vy_adamw: bad arguments");
  VY_CHECK_ARG(p->step >= 1 || p->step_ptr, "vy_adamw: step must be >= 1 (or pass step_ptr)");
  const float bc1 = 1.f - powf(p->beta1, static_cast<float>(p->step));
  const float bc2 = 1.f - powf(p->beta2, static_cast<float>(p->step));
  VY_CUDA_OK(launch_kernel(adamw_kernel, dim3(ew_grid(p->n, 1024)), dim3(256), 0, static_cast<cudaStream_t>(p->stream), 
      p->n, p->param, p->param_dtype, p->grad, p->grad_dtype, p->exp_avg, p->exp_avg_sq, p->master, p->lr, p->beta1,
      p->beta2, p->eps, p->weight_decay, bc1, bc2, p->step_ptr, p->grad_sqnorm, p->max_grad_norm,
      p->grad_div == 0.f ? 1.f : p->grad_div));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_rope_apply(const VyRope* p) {
  VY_CHECK_ARG(p != nullptr, "vy_rope_apply: null params");
  VY_NEED_DEVICE("vy_rope_apply");
  VY_CHECK_ARG(p->head_dim >= 2 && (p->head_dim & 1) == 0, "vy_rope_apply: head_dim must be even (got %d)", p->head_dim);
  VY_CHECK_ARG(p->B > 0 && p->H > 0 && p->S > 0 && p->x && p->out && (p->copy_only || (p->cos && p->sin)) && dtype_ok(p->dtype), "vy_rope_apply: bad arguments");
  const long long rows = static_cast<long long>(p->B) * p->H * p->S;
  VY_CUDA_OK(launch_kernel(rope_kernel, dim3(ew_grid(rows, 8)), dim3(256), 0, static_cast<cudaStream_t>(p->stream), 
      p->B, p->H, p->S, p->head_dim, p->x, p->x_sb, p->x_sh, p->x_sl, p->dtype, p->cos, p->sin, p->pos0, p->inverse, p->out, p->o_sb, p->o_sh, p->o_sl, p->pos_ptr, p->out_follows_pos, p->copy_only));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_rope_append(const VyRopeAppend* p) {
  VY_CHECK_ARG(p != nullptr, "vy_rope_append: null params");
  VY_NEED_DEVICE("vy_rope_append");
  VY_CHECK_ARG(p->B > 0 && p->S > 0 && p->n_q_heads >= 0 && p->n_kv_heads > 0 && p->head_dim >= 2 && (p->head_dim & 1) == 0,
               "vy_rope_append: bad shape");
  VY_CHECK_ARG(p->qkv && p->cos && p->sin && p->k_cache && p->v_cache && dtype_ok(p->dtype), "vy_rope_append: null pointer / bad dtype");
  VY_CHECK_ARG(p->pos0 >= 0 && p->slot0 >= 0 && (p->cache_slots <= 0 || p->pos_ptr || p->slot0 + p->S <= p->cache_slots),
               "vy_rope_append: slots [%d, %d) do not fit the %d slots of the cache", p->slot0, p->slot0 + p->S, p->cache_slots);
  VY_CHECK_ARG(p->rope_rows <= 0 || p->pos_ptr || p->pos0 + p->S <= p->rope_rows, "vy_rope_append: positions [%d, %d) exceed the %d rows of the tables",
               p->pos0, p->pos0 + p->S, p->rope_rows);
  const long long rows = static_cast<long long>(p->B) * (p->n_q_heads + 2 * p->n_kv_heads) * p->S;
  VY_CUDA_OK(launch_kernel(rope_append_kernel, dim3(ew_grid(rows, 8)), dim3(256), 0, static_cast<cudaStream_t>(p->stream), p->B, p->n_q_heads,
                           p->n_kv_heads, p->S, p->head_dim, p->qkv, p->sb, p->sh, p->sl, p->dtype, p->cos, p->sin, p->pos0, p->slot0, p->pos_ptr,
                           p->k_cache, p->v_cache, p->c_sb, p->c_sh, p->c_sl));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_scale_by_ptr(int64_t n, void* x, int dtype, const float* scale, void* stream) {
  VY_NEED_DEVICE("vy_scale_by_ptr");
  VY_CHECK_ARG(n > 0 && x && scale && dtype_ok(dtype) && aligned16(x), "vy_scale_by_ptr: bad arguments (x 16-byte aligned)");
  VY_CUDA_OK(launch_kernel(scale_by_ptr_kernel, dim3(ew_grid(n, 8 * 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                           static_cast<long long>(n), x, dtype, scale));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_slot_merge_fwd(int rows, int H, int dtype, const void* a, const void* b, int n_b, const int32_t* slot, void* out, void* stream) {
  VY_NEED_DEVICE("vy_slot_merge_fwd");
  VY_CHECK_ARG(rows > 0 && H > 0 && n_b >= 0 && dtype_ok(dtype) && a && b && slot && out && aligned16(a) && aligned16(b) && aligned16(out) &&
                   (static_cast<long long>(H) * dtype_size(dtype)) % 16 == 0,
               "vy_slot_merge_fwd: bad arguments (row bytes and pointers must be multiples of 16)");
  const int vpr = static_cast<int>(static_cast<long long>(H) * dtype_size(dtype) / 16);
  VY_CUDA_OK(launch_kernel(slot_merge_kernel, dim3(ew_grid(static_cast<long long>(rows) * vpr, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), rows, vpr,
                           static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<const int*>(slot), static_cast<uint4*>(out),
                           static_cast<uint4*>(nullptr), static_cast<uint4*>(nullptr), 0, n_b));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_slot_merge_bwd(int rows, int H, int dtype, const void* dout, const int32_t* slot, void* da, void* db, int n_b, void* stream) {
  VY_NEED_DEVICE("vy_slot_merge_bwd");
  VY_CHECK_ARG(rows > 0 && H > 0 && n_b >= 0 && dtype_ok(dtype) && dout && slot && (da || db) && aligned16(dout) && aligned16(da) && aligned16(db) &&
                   (static_cast<long long>(H) * dtype_size(dtype)) % 16 == 0,
               "vy_slot_merge_bwd: bad arguments (row bytes and pointers must be multiples of 16)");
  const int vpr = static_cast<int>(static_cast<long long>(H) * dtype_size(dtype) / 16);
  VY_CUDA_OK(launch_kernel(slot_merge_kernel, dim3(ew_grid(static_cast<long long>(rows) * vpr, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), rows, vpr,
                           static_cast<const uint4*>(dout), static_cast<const uint4*>(nullptr), static_cast<const int*>(slot),
                           static_cast<uint4*>(nullptr), static_cast<uint4*>(da), static_cast<uint4*>(db), 1, n_b));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_act_bwd(int64_t n, const void* dy, const void* z, int dtype, int act, void* out, void* stream) {
  VY_NEED_DEVICE("vy_act_bwd");
  VY_CHECK_ARG(n > 0 && n % 8 == 0 && dy && z && out && dtype_ok(dtype) && aligned16(dy) && aligned16(z) && aligned16(out),
               "vy_act_bwd: bad arguments (n must be a multiple of 8, pointers 16-byte aligned)");
  VY_CHECK_ARG(act == VY_ACT_GELU_ERF || act == VY_ACT_GELU_TANH, "vy_act_bwd: unsupported activation %d", act);
  VY_CUDA_OK(launch_kernel(act_bwd_kernel, dim3(ew_grid(n, 8 * 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), n, dy, z, dtype, act, out));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}
