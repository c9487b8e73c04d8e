// vy_attn_decode: single-token (Sq == 1) attention over a contiguous or paged kv-cache, HBM-bound.
//
// One CTA per (kv-split, kv head, batch row). Fuses, for the new token, the half-split RoPE of q
// and k, the kv-cache append at `start_pos`, and the attention over slots [0, start_pos] — with NO
// mask, exactly like the reference's decode step (models/decoder.py:355-362: mask is None when
// seqlen == 1; quirk Q3). GQA is a head-index map (q head i reads kv head i / n_rep), so the
// cache is streamed once per kv head instead of being re-materialised by repeat_kv
// (layers/attention.py:8-19, models/decoder.py:190-193).
//
// Data movement: a key/value row of 64 elements is read by 8 lanes x 8 elements (16-byte loads
// for bf16, 2 x 16 B for fp32), 4 rows per warp per step, UNROLL steps in flight; dot products
// are reduced with 3 xor-shuffles; softmax is online per lane-group and merged across lane
// groups (shuffles), warps (smem) and kv-splits (fp32 workspace + last-CTA ticket).
// Algorithmic bytes = 2 * B * Hkv * ctx * 64 * sizeof(cache dtype) per layer.
#include "vy_common.cuh"
#include "vy_ptx.cuh"

namespace vy {

constexpr int DEC_WARPS = 4;       // warps per CTA when the context is split over several CTAs
constexpr int DEC_WARPS_WIDE = 8;  // ... when one CTA owns a whole (batch row, kv head): no split, no workspace, no ticket
constexpr int DEC_UNROLL = 4;
constexpr int DEC_MAX_REP = 8;
constexpr int HD = 64;

struct DecodeDev {
  int B, Hq, Hkv, n_rep, start_pos, splits;
  const int* start_pos_ptr;  // device copy of start_pos (wins when set): lets a captured CUDA graph be replayed per token
  const void* qkv;  // [B, (Hq + 2 Hkv) * 64]
  long long ld_qkv;
  int qkv_dtype;
  const float* rope_cos;  // table [rows][32] (row = position + rope_pos_off), or null
  const float* rope_sin;
  int rope_pos_off;
  void* kcache;
  void* vcache;
  long long c_sb, c_sh, c_sl;  // element strides: batch (paged: block), head, slot
  const int* seqlens;      // optional [B]: per-row position of the new token (wins over start_pos); negative = skip the row
  const int* block_table;  // optional [B][table_stride]: paged cache, slot p of row b lives in block block_table[b][p / block_size]
  int table_stride, block_size, block_shift;  // block_shift = log2(block_size) when it is a power of two, else -1
  int cache_dtype;
  void* out;  // [B, Hq * 64]
  long long ld_out;
  int out_dtype;
  float scale_log2;
  float* ws;          // [B, Hkv, splits, n_rep, 66] partial (m, l, o[64])
  unsigned int* tickets;  // [B * Hkv], zero on entry, self-resetting
};

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned int bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <typename TC>
__device__ __forceinline__ void load_row8(const TC* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load_row8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 f = __bfloat1622float2(h[k]);
    v[2 * k] = f.x;
    v[2 * k + 1] = f.y;
  }
}
template <>
__device__ __forceinline__ void load_row8<float>(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// Merges the per-lane-group online-softmax states (m, l, o[8]) of one CTA — lane groups (shuffles), warps (smem), kv-splits
// (fp32 workspace + last-CTA ticket) — and writes the attention output of the NREP query heads of (b, kvh).
template <int NREP, int NWARPS>
__device__ __forceinline__ void decode_finish(const DecodeDev& g, float (&m)[NREP], float (&l)[NREP], float (&o)[NREP][8],
                                              float (*s_red)[DEC_MAX_REP][HD + 2], int split, int kvh, int b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = lane & 7, lk = lane >> 3;
  // ---- merge the 4 lane groups of each warp (xor 8, 16) ----
#pragma unroll
  for (int r = 0; r < NREP; ++r) {
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m[r], off);
      const float l2 = __shfl_xor_sync(0xffffffffu, l[r], off);
      const float mn = fmaxf(m[r], m2);
      const float a1 = (m[r] == -INFINITY) ? 0.f : exp2f(m[r] - mn);
      const float a2 = (m2 == -INFINITY) ? 0.f : exp2f(m2 - mn);
      l[r] = l[r] * a1 + l2 * a2;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float o2 = __shfl_xor_sync(0xffffffffu, o[r][j], off);
        o[r][j] = o[r][j] * a1 + o2 * a2;
      }
      m[r] = mn;
    }
    if (lk == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s_red[warp][r][ld * 8 + j] = o[r][j];
      if (ld == 0) {
        s_red[warp][r][HD] = m[r];
        s_red[warp][r][HD + 1] = l[r];
      }
    }
  }
  __syncthreads();

  // ---- merge warps; thread t < NREP*64 owns (head r, dim j) ----
  const int t = threadIdx.x;
  float M = -INFINITY, L = 0.f, O = 0.f;
  const int r_own = t / HD, j_own = t % HD;
  for (int rr = r_own; rr < NREP; rr += (NWARPS * 32) / HD) {
    M = -INFINITY; L = 0.f; O = 0.f;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) {
      const float mw = s_red[w][rr][HD], lw = s_red[w][rr][HD + 1], ow = s_red[w][rr][j_own];
      if (mw == -INFINITY) continue;
      const float mn = fmaxf(M, mw);
      const float a1 = (M == -INFINITY) ? 0.f : exp2f(M - mn);
      const float a2 = exp2f(mw - mn);
      L = L * a1 + lw * a2;
      O = O * a1 + ow * a2;
      M = mn;
    }
    if (g.splits == 1) {
      st_from_float(g.out, g.out_dtype, static_cast<long long>(b) * g.ld_out + (kvh * NREP + rr) * HD + j_own, O / L);
    } else {
      float* w = g.ws + ((((static_cast<long long>(b) * g.Hkv + kvh) * g.splits + split) * NREP + rr) * (HD + 2));
      w[j_own] = O;
      if (j_own == 0) {
        w[HD] = M;
        w[HD + 1] = L;
      }
    }
  }
  if (g.splits == 1) return;

  // ---- last CTA of this (b, kv head) combines the splits ----
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (t == 0) {
    const unsigned int prev = atomicAdd(&g.tickets[b * g.Hkv + kvh], 1u);
    s_last = (prev == static_cast<unsigned int>(g.splits - 1)) ? 1u : 0u;
    if (s_last) g.tickets[b * g.Hkv + kvh] = 0u;  // self-reset for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int rr = r_own; rr < NREP; rr += (NWARPS * 32) / HD) {
    M = -INFINITY; L = 0.f; O = 0.f;
    for (int si = 0; si < g.splits; ++si) {
      const float* w = g.ws + ((((static_cast<long long>(b) * g.Hkv + kvh) * g.splits + si) * NREP + rr) * (HD + 2));
      const float mw = __ldcg(w + HD), lw = __ldcg(w + HD + 1), ow = __ldcg(w + j_own);
      if (mw == -INFINITY) continue;
      const float mn = fmaxf(M, mw);
      const float a1 = (M == -INFINITY) ? 0.f : exp2f(M - mn);
      const float a2 = exp2f(mw - mn);
      L = L * a1 + lw * a2;
      O = O * a1 + ow * a2;
      M = mn;
    }
    st_from_float(g.out, g.out_dtype, static_cast<long long>(b) * g.ld_out + (kvh * NREP + rr) * HD + j_own, O / L);
  }
}

template <typename TC, int NREP, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32)
attn_decode_kernel(const DecodeDev g) {
  pdl_trigger();
  pdl_wait();
  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = lane & 7;   // which 8-element slice of the head dim
  const int lk = lane >> 3;  // key sub-index within the warp step (0..3)
  // slot / position of the new token: per row (continuous batching), from device memory (graph replay), or the host value
  const int sp = g.seqlens ? g.seqlens[b] : (g.start_pos_ptr ? *g.start_pos_ptr : g.start_pos);
  if (sp < 0) return;  // an idle batch slot
  // element offset of cache slot `pos` of this (row, kv head): contiguous [B, Hkv, len, 64] through strides, or paged
  // [blocks, block_size, Hkv, 64] through the row's block table (one int32 lookup per row, L1-resident)
  const int* tbl = g.block_table ? g.block_table + static_cast<long long>(b) * g.table_stride : nullptr;
  auto slot_off = [&](int pos) -> long long {
    if (tbl) {
      const int blk = g.block_shift >= 0 ? pos >> g.block_shift : pos / g.block_size;
      const int off = g.block_shift >= 0 ? pos & (g.block_size - 1) : pos - blk * g.block_size;
      return static_cast<long long>(__ldg(tbl + blk)) * g.c_sb + off * g.c_sl;
    }
    return pos * g.c_sl;
  };

  __shared__ float s_newk[HD];
  __shared__ float s_newv[HD];
  __shared__ float s_q[DEC_MAX_REP][HD];
  __shared__ float s_red[NWARPS][DEC_MAX_REP][HD + 2];

  // ---- new token: bias-added projections -> RoPE(q, k) -> smem; append k, v to the cache ----
  {
    const long long row = static_cast<long long>(b) * g.ld_qkv;
    const int t = threadIdx.x;
    for (int idx = t; idx < (NREP + 2) * HD; idx += blockDim.x) {
      const int which = idx / HD;  // 0..NREP-1: q heads, NREP: k, NREP+1: v
      const int j = idx % HD;
      int col;
      if (which < NREP) col = (kvh * NREP + which) * HD;
      else if (which == NREP) col = (g.Hq + kvh) * HD;
      else col = (g.Hq + g.Hkv + kvh) * HD;
      float x = ld_as_float(g.qkv, g.qkv_dtype, row + col + j);
      if (which <= NREP && g.rope_cos) {
        const int jj = j & 31;
        const float other = ld_as_float(g.qkv, g.qkv_dtype, row + col + (j < 32 ? j + 32 : j - 32));
        const int rr = sp + g.rope_pos_off;
        const float c = g.rope_cos[rr * 32 + jj], s = g.rope_sin[rr * 32 + jj];
        x = j < 32 ? x * c - other * s : x * c + other * s;
      }
      if (which < NREP) s_q[which][j] = x;
      else if (which == NREP) s_newk[j] = x;
      else s_newv[j] = x;
    }
  }
  __syncthreads();
  TC* kc = reinterpret_cast<TC*>(g.kcache) + (tbl ? 0 : b * g.c_sb) + kvh * g.c_sh;
  TC* vc = reinterpret_cast<TC*>(g.vcache) + (tbl ? 0 : b * g.c_sb) + kvh * g.c_sh;
  if (split == g.splits - 1 && threadIdx.x < HD) {
    // the cache stores what the reference stores: rotated k and raw v, rounded to the cache dtype
    st_from_float(kc, sizeof(TC) == 2 ? VY_BF16 : VY_F32, slot_off(sp) + threadIdx.x, s_newk[threadIdx.x]);
    st_from_float(vc, sizeof(TC) == 2 ? VY_BF16 : VY_F32, slot_off(sp) + threadIdx.x, s_newv[threadIdx.x]);
  }

  float q[NREP][8];
#pragma unroll
  for (int r = 0; r < NREP; ++r)
#pragma unroll
    for (int j = 0; j < 8; ++j) q[r][j] = s_q[r][ld * 8 + j] * g.scale_log2;

  float m[NREP], l[NREP], o[NREP][8];
#pragma unroll
  for (int r = 0; r < NREP; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[r][j] = 0.f;
  }

  // cached slots [0, start_pos) are split evenly; the new token (slot start_pos) is taken from
  // smem by the last split so no CTA has to wait for the append to become visible.
  const int per = (sp + g.splits - 1) / g.splits;
  const int k_begin = split * per;
  const int k_end = min(sp, k_begin + per);
  constexpr int KEYS_PER_ITER = NWARPS * 4 * DEC_UNROLL;

  for (int k0 = k_begin; k0 < k_end; k0 += KEYS_PER_ITER) {
    float kv[DEC_UNROLL][8], vv[DEC_UNROLL][8];
    int kidx[DEC_UNROLL];
#pragma unroll
    for (int u = 0; u < DEC_UNROLL; ++u) {
      kidx[u] = k0 + (u * NWARPS + warp) * 4 + lk;
      if (kidx[u] < k_end) {
        const long long off = slot_off(kidx[u]) + ld * 8;
        load_row8<TC>(kc + off, kv[u]);
        load_row8<TC>(vc + off, vv[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < DEC_UNROLL; ++u) {
      const bool valid = kidx[u] < k_end;
#pragma unroll
      for (int r = 0; r < NREP; ++r) {
        float s = 0.f;
        if (valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j) s += q[r][j] * kv[u][j];
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (valid) {
          const float mn = fmaxf(m[r], s);
          const float a = exp2f(m[r] - mn);
          const float p = exp2f(s - mn);
          l[r] = l[r] * a + p;
#pragma unroll
          for (int j = 0; j < 8; ++j) o[r][j] = o[r][j] * a + p * vv[u][j];
          m[r] = mn;
        }
      }
    }
  }
  // the new token
  if (split == g.splits - 1 && warp == 0 && lk == 0) {
#pragma unroll
    for (int r = 0; r < NREP; ++r) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += q[r][j] * s_newk[ld * 8 + j];
      s += __shfl_xor_sync(0x000000ffu, s, 1);
      s += __shfl_xor_sync(0x000000ffu, s, 2);
      s += __shfl_xor_sync(0x000000ffu, s, 4);
      const float mn = fmaxf(m[r], s);
      const float a = exp2f(m[r] - mn);
      const float p = exp2f(s - mn);
      l[r] = l[r] * a + p;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[r][j] = o[r][j] * a + p * s_newv[ld * 8 + j];
      m[r] = mn;
    }
  }

  decode_finish<NREP, NWARPS>(g, m, l, o, s_red, split, kvh, b);
}

// ---- copy-engine variant: bf16 cache with contiguous slots (slot stride = 64 elements, no block table) ----
// The key / value rows of one (batch row, kv head) are then ONE contiguous range, so the whole slice of a CTA is requested
// from HBM up front with bulk copies (cp.async.bulk -> shared memory, mbarrier tx-count) instead of 16-byte loads issued
// loop iteration by loop iteration: a 16 KB chunk = 128 keys, `stages` chunks of K and of V in flight per CTA (all of a
// C3-sized context at once), consumed in order and refilled while later chunks are still arriving. The LDG version keeps
// 128 B per thread in flight and pays one HBM round trip per 128 keys (ncu: 722 GB/s, long_scoreboard-bound).
constexpr int TD_CHUNK = 128;                       // keys per chunk
constexpr int TD_CHUNK_BYTES = TD_CHUNK * HD * 2;   // 16 KB of bf16
constexpr int TD_WARPS = 8;
constexpr int TD_MAX_STAGES = 6;

template <int NREP>
__global__ void __launch_bounds__(TD_WARPS * 32)
attn_decode_tma_kernel(const DecodeDev g, const int stages) {
  extern __shared__ __align__(128) unsigned char td_smem[];  // [stages][K chunk | V chunk]
  __shared__ __align__(8) uint64_t s_full[TD_MAX_STAGES];
  __shared__ __align__(8) uint64_t s_empty[TD_MAX_STAGES];
  __shared__ float s_newk[HD];
  __shared__ float s_newv[HD];
  __shared__ float s_q[DEC_MAX_REP][HD];
  pdl_trigger();
  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = lane & 7, lk = lane >> 3;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], TD_WARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  const int sp = g.seqlens ? g.seqlens[b] : (g.start_pos_ptr ? *g.start_pos_ptr : g.start_pos);
  if (sp < 0) return;  // an idle batch slot
  const __nv_bfloat16* kc = reinterpret_cast<const __nv_bfloat16*>(g.kcache) + b * g.c_sb + kvh * g.c_sh;
  const __nv_bfloat16* vc = reinterpret_cast<const __nv_bfloat16*>(g.vcache) + b * g.c_sb + kvh * g.c_sh;
  // cached slots [0, sp) are split evenly; the new token comes from smem (last split), as in the LDG kernel
  const int per = (sp + g.splits - 1) / g.splits;
  const int k_begin = split * per;
  const int k_end = min(sp, k_begin + per);
  const int nchunks = k_end > k_begin ? (k_end - k_begin + TD_CHUNK - 1) / TD_CHUNK : 0;
  auto issue = [&](int c) {  // one thread: request chunk c of K and V
    const int st = c % stages;
    const int key0 = k_begin + c * TD_CHUNK;
    const unsigned int bytes = static_cast<unsigned int>(min(TD_CHUNK, k_end - key0)) * (HD * 2);
    unsigned char* dst = td_smem + static_cast<size_t>(st) * (2 * TD_CHUNK_BYTES);
    mbar_arrive_expect_tx(&s_full[st], 2 * bytes);
    bulk_g2s(dst, kc + static_cast<long long>(key0) * HD, bytes, &s_full[st]);
    bulk_g2s(dst + TD_CHUNK_BYTES, vc + static_cast<long long>(key0) * HD, bytes, &s_full[st]);
  };
  if (threadIdx.x == 0) {
    const int n0 = min(nchunks, stages);
    for (int c = 0; c < n0; ++c) issue(c);
  }

  // ---- new token (while the copies fly): bias-added projections -> RoPE(q, k) -> smem; append k, v to the cache ----
  {
    const long long row = static_cast<long long>(b) * g.ld_qkv;
    for (int idx = threadIdx.x; idx < (NREP + 2) * HD; idx += blockDim.x) {
      const int which = idx / HD;  // 0..NREP-1: q heads, NREP: k, NREP+1: v
      const int j = idx % HD;
      int col;
      if (which < NREP) col = (kvh * NREP + which) * HD;
      else if (which == NREP) col = (g.Hq + kvh) * HD;
      else col = (g.Hq + g.Hkv + kvh) * HD;
      float x = ld_as_float(g.qkv, g.qkv_dtype, row + col + j);
      if (which <= NREP && g.rope_cos) {
        const int jj = j & 31;
        const float other = ld_as_float(g.qkv, g.qkv_dtype, row + col + (j < 32 ? j + 32 : j - 32));
        const int rr = sp + g.rope_pos_off;
        const float c = g.rope_cos[rr * 32 + jj], sn = g.rope_sin[rr * 32 + jj];
        x = j < 32 ? x * c - other * sn : x * c + other * sn;
      }
      if (which < NREP) s_q[which][j] = x;
      else if (which == NREP) s_newk[j] = x;
      else s_newv[j] = x;
    }
  }
  __syncthreads();
  if (split == g.splits - 1 && threadIdx.x < HD) {
    __nv_bfloat16* kw = reinterpret_cast<__nv_bfloat16*>(g.kcache) + b * g.c_sb + kvh * g.c_sh + static_cast<long long>(sp) * HD;
    __nv_bfloat16* vw = reinterpret_cast<__nv_bfloat16*>(g.vcache) + b * g.c_sb + kvh * g.c_sh + static_cast<long long>(sp) * HD;
    kw[threadIdx.x] = __float2bfloat16(s_newk[threadIdx.x]);
    vw[threadIdx.x] = __float2bfloat16(s_newv[threadIdx.x]);
  }

  float q[NREP][8];
#pragma unroll
  for (int r = 0; r < NREP; ++r)
#pragma unroll
    for (int j = 0; j < 8; ++j) q[r][j] = s_q[r][ld * 8 + j] * g.scale_log2;
  float m[NREP], l[NREP], o[NREP][8];
#pragma unroll
  for (int r = 0; r < NREP; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[r][j] = 0.f;
  }

  for (int c = 0; c < nchunks; ++c) {
    const int st = c % stages;
    const unsigned int parity = static_cast<unsigned int>(c / stages) & 1u;
    mbar_wait(&s_full[st], parity);
    const int keys = min(TD_CHUNK, k_end - (k_begin + c * TD_CHUNK));
    const unsigned char* kt = td_smem + static_cast<size_t>(st) * (2 * TD_CHUNK_BYTES);
    const unsigned char* vt = kt + TD_CHUNK_BYTES;
    // key (u * 8 + warp) * 4 + lk: a partly filled chunk is still spread over all warps. The 4 keys a lane group sees per
    // chunk are scored first (independent dot products and shuffles), then folded into the running softmax with ONE rescale
    // per head — a per-key online update is a chain of dependent ex2 / FMA that two warps per scheduler cannot hide.
    constexpr int KU = TD_CHUNK / (TD_WARPS * 4);
    float kvk[KU][8], vvk[KU][8];
    bool valid[KU];
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const int kidx = (u * TD_WARPS + warp) * 4 + lk;
      valid[u] = kidx < keys;
      if (valid[u]) {
        const uint4 kr = *reinterpret_cast<const uint4*>(kt + kidx * (HD * 2) + ld * 16);
        const uint4 vr = *reinterpret_cast<const uint4*>(vt + kidx * (HD * 2) + ld * 16);
        const __nv_bfloat162* kh = reinterpret_cast<const __nv_bfloat162*>(&kr);
        const __nv_bfloat162* vh = reinterpret_cast<const __nv_bfloat162*>(&vr);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 a = __bfloat1622float2(kh[k]), bb = __bfloat1622float2(vh[k]);
          kvk[u][2 * k] = a.x; kvk[u][2 * k + 1] = a.y;
          vvk[u][2 * k] = bb.x; vvk[u][2 * k + 1] = bb.y;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) kvk[u][k] = vvk[u][k] = 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < NREP; ++r) {
      float sc[KU];
#pragma unroll
      for (int u = 0; u < KU; ++u) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += q[r][j] * kvk[u][j];
        sc[u] = t;
      }
#pragma unroll
      for (int u = 0; u < KU; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 1);
#pragma unroll
      for (int u = 0; u < KU; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 2);
#pragma unroll
      for (int u = 0; u < KU; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 4);
      float mn = m[r];
#pragma unroll
      for (int u = 0; u < KU; ++u)
        if (valid[u]) mn = fmaxf(mn, sc[u]);
      if (mn != -INFINITY) {  // (-inf: this lane group has not seen a key yet)
        const float a = exp2f(m[r] - mn);
        float pp[KU], psum = 0.f;
#pragma unroll
        for (int u = 0; u < KU; ++u) {
          pp[u] = valid[u] ? exp2f(sc[u] - mn) : 0.f;
          psum += pp[u];
        }
        l[r] = l[r] * a + psum;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float acc = o[r][j] * a;
#pragma unroll
          for (int u = 0; u < KU; ++u) acc += pp[u] * vvk[u][j];
          o[r][j] = acc;
        }
        m[r] = mn;
      }
    }
    if (c + stages < nchunks) {  // the stage is needed again: hand it back, thread 0 refills it once every warp has
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
      if (threadIdx.x == 0) {
        mbar_wait(&s_empty[st], parity);
        issue(c + stages);
      }
    }
  }
  // the new token
  if (split == g.splits - 1 && warp == 0 && lk == 0) {
#pragma unroll
    for (int r = 0; r < NREP; ++r) {
      float sc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) sc += q[r][j] * s_newk[ld * 8 + j];
      sc += __shfl_xor_sync(0x000000ffu, sc, 1);
      sc += __shfl_xor_sync(0x000000ffu, sc, 2);
      sc += __shfl_xor_sync(0x000000ffu, sc, 4);
      const float mn = fmaxf(m[r], sc);
      const float a = exp2f(m[r] - mn);
      const float pp = exp2f(sc - mn);
      l[r] = l[r] * a + pp;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[r][j] = o[r][j] * a + pp * s_newv[ld * 8 + j];
      m[r] = mn;
    }
  }
  // every requested chunk has landed and been consumed: the ring now holds the warps' partial results (16.9 KB)
  __syncthreads();
  float (*s_red)[DEC_MAX_REP][HD + 2] = reinterpret_cast<float (*)[DEC_MAX_REP][HD + 2]>(td_smem);
  decode_finish<NREP, TD_WARPS>(g, m, l, o, s_red, split, kvh, b);
}

template <int NREP>
static int launch_decode_tma_r(const DecodeDev& g, int stages, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(stages) * 2 * TD_CHUNK_BYTES;
  static bool attr_done[64] = {};  // per device: the limit is an attribute of the function in that device's context
  int dev = 0;
  VY_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && !attr_done[dev]) {
    VY_CUDA_OK(cudaFuncSetAttribute(attn_decode_tma_kernel<NREP>, cudaFuncAttributeMaxDynamicSharedMemorySize, TD_MAX_STAGES * 2 * TD_CHUNK_BYTES));
    attr_done[dev] = true;
  }
  VY_CUDA_OK(launch_kernel(attn_decode_tma_kernel<NREP>, dim3(g.splits, g.Hkv, g.B), dim3(TD_WARPS * 32), smem, st, g, stages));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}
static int launch_decode_tma(const DecodeDev& g, int stages, cudaStream_t st) {
  switch (g.n_rep) {
    case 1: return launch_decode_tma_r<1>(g, stages, st);
    case 2: return launch_decode_tma_r<2>(g, stages, st);
    case 3: return launch_decode_tma_r<3>(g, stages, st);
    case 4: return launch_decode_tma_r<4>(g, stages, st);
    case 6: return launch_decode_tma_r<6>(g, stages, st);
    case 8: return launch_decode_tma_r<8>(g, stages, st);
    default:
      set_error("vy_attn_decode: unsupported q-heads per kv-head %d (1,2,3,4,6,8)", g.n_rep);
      return VY_ERR_UNSUPPORTED;
  }
}

template <typename TC, int NWARPS>
static int launch_decode_w(const DecodeDev& g, cudaStream_t st) {
  dim3 grid(g.splits, g.Hkv, g.B);
  dim3 block(NWARPS * 32);
  switch (g.n_rep) {
    case 1: VY_CUDA_OK(launch_kernel(attn_decode_kernel<TC, 1, NWARPS>, dim3(grid), dim3(block), 0, st, g)); break;
    case 2: VY_CUDA_OK(launch_kernel(attn_decode_kernel<TC, 2, NWARPS>, dim3(grid), dim3(block), 0, st, g)); break;
    case 3: VY_CUDA_OK(launch_kernel(attn_decode_kernel<TC, 3, NWARPS>, dim3(grid), dim3(block), 0, st, g)); break;
    case 4: VY_CUDA_OK(launch_kernel(attn_decode_kernel<TC, 4, NWARPS>, dim3(grid), dim3(block), 0, st, g)); break;
    case 6: VY_CUDA_OK(launch_kernel(attn_decode_kernel<TC, 6, NWARPS>, dim3(grid), dim3(block), 0, st, g)); break;
    case 8: VY_CUDA_OK(launch_kernel(attn_decode_kernel<TC, 8, NWARPS>, dim3(grid), dim3(block), 0, st, g)); break;
    default:
      set_error("vy_attn_decode: unsupported q-heads per kv-head %d (1,2,3,4,6,8)", g.n_rep);
      return VY_ERR_UNSUPPORTED;
  }
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}
template <typename TC>
static int launch_decode(const DecodeDev& g, cudaStream_t st) {
  // one CTA per (batch row, kv head) gets 8 warps (its context is not split); split contexts use 4-warp CTAs
  if (g.splits == 1) return launch_decode_w<TC, DEC_WARPS_WIDE>(g, st);
  return launch_decode_w<TC, DEC_WARPS>(g, st);
}

}  // namespace vy

extern "C" int vy_attn_decode_splits(int B, int Hkv, int start_pos) {
  // enough CTAs to cover the GPU ~4x, but keep >= 64 cached slots per split
  const int sms = vy::num_sms();
  // enough (batch row, kv head) pairs to occupy most SMs on their own: no split — one 8-warp CTA per pair streams the
  // whole context (no workspace round trip, no ticket, a single wave)
  if (3 * B * Hkv >= 2 * sms) return 1;
  int splits = (4 * sms + B * Hkv - 1) / (B * Hkv);
  const int max_by_len = (start_pos + 63) / 64;
  if (splits > max_by_len) splits = max_by_len;
  if (splits < 1) splits = 1;
  if (splits > 32) splits = 32;
  return splits;
}

// The copy-engine kernel serves bf16 caches whose slots are contiguous per (row, kv head) and unpaged.
static bool decode_tma_ok(const VyDecode* p) {
  static const bool off = getenv("VY_DECODE_ATTN_LDG") && atoi(getenv("VY_DECODE_ATTN_LDG")) != 0;  // development: A/B against the LDG kernel
  return !off && p->cache_dtype == VY_BF16 && p->block_table == nullptr && p->cache_sl == 64 && p->head_dim == 64;
}
// splits / stages of the copy-engine kernel for a context of at most `ctx` cached slots: <= 3 stages (96 KB) lets two CTAs share an
// SM, so aim for ~2 CTAs per SM; a CTA whose slice is <= 3 chunks then has its whole slice in flight from the start.
static void decode_tma_plan(int B, int Hkv, int ctx, int* splits, int* stages) {
  static const int f_splits = getenv("VY_DECODE_TMA_SPLITS") ? atoi(getenv("VY_DECODE_TMA_SPLITS")) : 0;
  static const int f_stages = getenv("VY_DECODE_TMA_STAGES") ? atoi(getenv("VY_DECODE_TMA_STAGES")) : 0;
  const int items = B * Hkv, sms = vy::num_sms();
  const int chunks = ctx > 0 ? (ctx + vy::TD_CHUNK - 1) / vy::TD_CHUNK : 1;
  // measured on config 3 (profiles/r02_decode_attn_ab.txt): once the (row, kv head) pairs alone cover 3/4 of the SMs a split only
  // adds the workspace round trip (GQA 128 pairs: 355 us/step unsplit or 2-way, 377 3-way; MHA 384 pairs: 381 unsplit, 399 2-way)
  int sp = 4 * items >= 3 * sms ? 1 : (2 * sms + items - 1) / items;
  if (sp > chunks) sp = chunks;
  if (sp > 32) sp = 32;
  if (sp < 1) sp = 1;
  if (f_splits > 0) sp = f_splits > chunks ? chunks : f_splits;
  const int per_chunks = ((ctx + sp - 1) / sp + vy::TD_CHUNK - 1) / vy::TD_CHUNK;
  int st = per_chunks < 3 ? (per_chunks < 1 ? 1 : per_chunks) : 3;
  if (static_cast<long long>(items) * sp <= sms && per_chunks > 3) st = per_chunks < vy::TD_MAX_STAGES ? per_chunks : vy::TD_MAX_STAGES;  // one CTA per SM anyway
  // 3-stage rings (96 KB) put two CTAs on an SM, 2-stage rings (64 KB) three: when the CTAs need more than two per SM but fit
  // three, one wave of 2-stage CTAs beats a full wave plus a thin second one (MHA config 3: 384 CTAs, 252.6 -> 236.8 us/step)
  if (st == 3 && static_cast<long long>(items) * sp > 2LL * sms && static_cast<long long>(items) * sp <= 3LL * sms) st = 2;
  if (f_stages > 0) st = f_stages > vy::TD_MAX_STAGES ? vy::TD_MAX_STAGES : f_stages;
  *splits = sp;
  *stages = st;
}

extern "C" int vy_attn_decode_plan(const VyDecode* p) {
  if (!p || p->B <= 0 || p->n_kv_heads <= 0) return 1;
  if (decode_tma_ok(p)) {
    int sp, st;
    decode_tma_plan(p->B, p->n_kv_heads, p->start_pos, &sp, &st);
    return sp;
  }
  return vy_attn_decode_splits(p->B, p->n_kv_heads, p->start_pos);
}

extern "C" int vy_attn_decode(const VyDecode* p) {
  using namespace vy;
  VY_CHECK_ARG(p != nullptr, "vy_attn_decode: null params");
  if (!vy_device_ok()) {
    set_error("vy_attn_decode: no sm_100 device (there is no CPU fallback)");
    return VY_ERR_NO_DEVICE;
  }
  VY_CHECK_ARG(p->head_dim == 64, "vy_attn_decode: head_dim must be 64 (got %d)", p->head_dim);
  VY_CHECK_ARG(p->B > 0 && p->n_q_heads > 0 && p->n_kv_heads > 0 && p->n_q_heads % p->n_kv_heads == 0,
               "vy_attn_decode: bad head counts");
  VY_CHECK_ARG(p->start_pos >= 0 && p->start_pos < p->cache_len, "vy_attn_decode: start_pos %d outside the cache (%d)",
               p->start_pos, p->cache_len);
  if (p->block_table)
    VY_CHECK_ARG(p->block_size > 0 && p->max_blocks_per_seq > 0 &&
                     static_cast<long long>(p->max_blocks_per_seq) * p->block_size >= p->start_pos + 1,
                 "vy_attn_decode: paged cache needs block_size > 0 and max_blocks_per_seq * block_size > start_pos");
  VY_CHECK_ARG(p->qkv && p->k_cache && p->v_cache && p->out, "vy_attn_decode: null pointer");
  VY_CHECK_ARG(dtype_ok(p->qkv_dtype) && dtype_ok(p->cache_dtype) && dtype_ok(p->out_dtype), "vy_attn_decode: bad dtype");
  VY_CHECK_ARG((p->rope_cos == nullptr) == (p->rope_sin == nullptr), "vy_attn_decode: rope tables must both be set or NULL");
  // start_pos is the position itself, or (start_pos_ptr / seqlens) its upper bound: either way the table must cover it
  if (p->rope_cos && p->rope_rows > 0)
    VY_CHECK_ARG(p->start_pos + p->rope_pos_off < p->rope_rows && (p->start_pos_ptr || p->seqlens || p->start_pos + p->rope_pos_off >= 0),
                 "vy_attn_decode: position %d (+ offset %d) lies outside the %d rows of the RoPE tables", p->start_pos,
                 p->rope_pos_off, p->rope_rows);
  const long long es = dtype_size(p->cache_dtype);
  VY_CHECK_ARG((reinterpret_cast<uintptr_t>(p->k_cache) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->v_cache) & 15) == 0 &&
                   (p->cache_sb * es) % 16 == 0 && (p->cache_sh * es) % 16 == 0 && (p->cache_sl * es) % 16 == 0,
               "vy_attn_decode: cache pointers/strides must keep 16-byte alignment");
  const bool tma = decode_tma_ok(p);
  int splits = p->splits > 0 ? p->splits : vy_attn_decode_plan(p);
  int stages = 0;
  if (tma) {
    int sp0;
    decode_tma_plan(p->B, p->n_kv_heads, p->start_pos, &sp0, &stages);
    if (splits != sp0) {  // the caller pinned another split: size the ring for its slice
      const int per_chunks = ((p->start_pos + splits - 1) / splits + TD_CHUNK - 1) / TD_CHUNK;
      stages = per_chunks < 1 ? 1 : (per_chunks > 3 ? 3 : per_chunks);
    }
  }
  if (splits > 1) VY_CHECK_ARG(p->workspace && p->tickets, "vy_attn_decode: split-kv needs workspace and tickets");

  DecodeDev g;
  g.B = p->B; g.Hq = p->n_q_heads; g.Hkv = p->n_kv_heads; g.n_rep = p->n_q_heads / p->n_kv_heads;
  g.start_pos = p->start_pos; g.splits = splits;
  g.start_pos_ptr = p->start_pos_ptr;
  g.qkv = p->qkv; g.ld_qkv = p->ld_qkv; g.qkv_dtype = p->qkv_dtype;
  g.rope_cos = p->rope_cos;
  g.rope_sin = p->rope_sin;
  g.rope_pos_off = p->rope_pos_off;
  g.kcache = p->k_cache; g.vcache = p->v_cache;
  g.c_sb = p->cache_sb; g.c_sh = p->cache_sh; g.c_sl = p->cache_sl; g.cache_dtype = p->cache_dtype;
  g.seqlens = p->seqlens;
  g.block_table = p->block_table; g.table_stride = p->max_blocks_per_seq; g.block_size = p->block_size;
  g.block_shift = -1;
  if (p->block_size > 0 && (p->block_size & (p->block_size - 1)) == 0)
    for (int sft = 0; sft < 31; ++sft)
      if ((1 << sft) == p->block_size) g.block_shift = sft;
  g.out = p->out; g.ld_out = p->ld_out; g.out_dtype = p->out_dtype;
  g.scale_log2 = 1.4426950408889634f / 8.0f;  // log2(e) / sqrt(64)
  g.ws = p->workspace; g.tickets = reinterpret_cast<unsigned int*>(p->tickets);
  cudaStream_t st = static_cast<cudaStream_t>(p->stream);
  if (tma) return launch_decode_tma(g, stages, st);
  if (p->cache_dtype == VY_BF16) return launch_decode<__nv_bfloat16>(g, st);
  return launch_decode<float>(g, st);
}
