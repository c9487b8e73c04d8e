// vy_decode_step: ONE persistent kernel per generated token for the GPT-style decoder (DecoderModel, bf16).
//
// What the reference runs per token (models/decoder.py:430-514 -> forward :324-374 -> DecoderLayer :222-250): embedding,
// then per layer q/k/v Linear -> RoPE -> cache append -> repeat_kv -> SDPA(mask=None) -> out Linear + residual +
// LayerNorm -> FFN (Linear, GELU, Linear) + residual (the LAYER INPUT, quirk Q2) + LayerNorm, then the LM head
// (Linear, GELU, LayerNorm, Linear to the vocabulary) and topk(1). With batch 32 that is ~213 MB of weights + kv-cache
// streamed once per token: an HBM-bound step whose floor is ~33 us, which the one-kernel-per-op form (22 single-wave
// GEMM launches of 8-20 us + 4 attention launches) misses by 10x because every launch pays its own latency chain.
//
// Here the whole step is one launch of one CTA per SM. Stages are separated by grid-wide barriers (an arrive counter per
// barrier in global memory, acquire polling by one thread per CTA):
//
//   per layer   QKV   [x = embedding row or LN2(previous layer's sum)] -> q|k|v = x Wqkv^T + b         (fp32 scratch)
//               ATTN  RoPE(q, k) at the current position, append k / v to the cache, online-softmax attention over the
//                     cache, split over CTAs, last-arriver combine                                       (bf16 scratch)
//               OUT   s1 = attn Wo^T + bo + x                                                            (fp32 scratch)
//               FFN1  a = gelu(LN1(s1) W1^T + b1)                                                        (bf16 scratch)
//               FFN2  s2 = a W2^T + b2 + x                                                               (fp32 scratch)
//   head        LMD   g = gelu(LN2(s2) Wd^T + bd)
//               LMV   logits = LN(g) Wv^T + bv  -> per-row argmax (packed 64-bit atomicMax, first index on ties)
//               FIN   token write-back, position += 1
//
// GEMM stages: tokens are the N side of mma.sync.m16n8k16 (bf16 in, fp32 accumulate), 16 output features x K is a unit.
// The activations of the stage ([B <= 32][K] bf16) live in shared memory (built by the stage prologue: LayerNorm of the
// fp32 sums, a copy, or the embedding gather); weights stream from HBM straight into the A fragments with 16-byte loads
// — thread (g, t) of a warp loads W[row g | g + 8][32 j + 8 t .. + 7], and the same k-permutation is applied to the
// activation fragments, so no shuffle or shared-memory staging of weights is needed. Small stages split K over the 8
// warps of a CTA (every load of a unit is in flight at once: the stage costs about one DRAM round trip) and reduce through
// shared memory in a fixed order (deterministic); the vocabulary projection gives every warp its own units and
// double-buffers the weight fragments in registers.
//
// Algorithmic bytes per step = all weights except the embedding table + 2 * B * L * h_kv * ctx * 64 * 2 (kv-cache).
#include <mutex>

#include "vy_common.cuh"
#include "vy_ptx.cuh"

namespace vy {

constexpr int DS_THREADS = 256;
constexpr int DS_WARPS = 8;
constexpr int DS_MAXB = 32;
constexpr int DS_HD = 64;
constexpr int DS_MAX_BARRIERS = 5 * VY_DECODE_MAX_LAYERS + 8;
constexpr int DS_RED = 16 * 33;  // one warp's partial tile: [16 features][32 tokens], rows padded against bank conflicts

typedef __nv_bfloat16 bf16;

struct DsLayer {
  const bf16* w_qkv; const bf16* b_qkv;
  const bf16* w_o;   const bf16* b_o;
  const bf16* ln1_g; const bf16* ln1_b;
  const bf16* w_1;   const bf16* b_1;
  const bf16* w_2;   const bf16* b_2;
  const bf16* ln2_g; const bf16* ln2_b;
  bf16* k_cache; bf16* v_cache;
};

struct DsParams {
  int B, H, Hq, Hkv, FF, V, L, NQKV;
  int ng;          // token groups of 8
  int cache_len, splits;
  long long c_sb, c_sh, c_sl;
  float eps_layer, eps_head;
  const bf16* emb;       // [vocab][H]
  const bf16* pos_table; // [max_pos][H] or null
  const float* rope_cos; // [rows][32] or null
  const float* rope_sin;
  const bf16* w_d; const bf16* b_d; const bf16* lnh_g; const bf16* lnh_b; const bf16* w_v; const bf16* b_v;
  int* pos;                   // device: cache slot / position of the token being fed
  long long* tok;             // device [B]: token fed to this step, overwritten with the next token
  long long* tokens_out;      // [B][ld_tokens] or null
  long long ld_tokens;
  void* logits;               // optional [B][ld_logits] bf16
  long long ld_logits;
  // scratch (global)
  bf16* xbuf;    // [B][H]   layer input x
  float* qkv;    // [B][NQKV]
  bf16* attn;    // [B][Hq*64]
  float* s1;     // [B][H]
  bf16* abuf;    // [B][FF]
  float* s2;     // [B][H]
  float* gbuf;   // [B][H]
  float* part;   // [B][Hkv][splits][n_rep][66]
  unsigned int* tickets;      // [B*Hkv]
  unsigned long long* amax;   // [B]
  unsigned int* bar;          // [DS_MAX_BARRIERS]
  int* abort_flag;
  long long* trace;           // optional: globaltimer stamps of CTA 0
  DsLayer layer[VY_DECODE_MAX_LAYERS];
};

// ---- small helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream(const void* p) {  // weights / cache: read once per step
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned int ds_order_bits(float v) {
  const unsigned int u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Grid-wide barrier number k of this launch. One arrive counter per barrier; CTA 0 clears counter k - 1 once it has
// passed barrier k (every CTA has stopped polling it by then) and the last counter at the start of the next launch.
// A wait that does not complete (a CTA that is not resident: the launch did not get every SM) raises the abort flag;
// every later barrier then falls through, so the kernel ends with garbage and a raised flag instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier(const DsParams& p, int k) {
  __syncthreads();
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0 && p.trace) p.trace[1 + 2 * k] = static_cast<long long>(globaltimer_ns());  // CTA 0 done with the stage
    __threadfence();
    atomicAdd(&p.bar[k], 1u);
    unsigned int spins = 0;
    while (ld_acquire_u32(&p.bar[k]) < gridDim.x) {
      if ((++spins & 255u) == 0) {
        if (*reinterpret_cast<volatile int*>(p.abort_flag) != 0) break;
        if (spins > (1u << 24)) {
          atomicExch(p.abort_flag, 1);
          break;
        }
      }
    }
    if (blockIdx.x == 0 && k > 0) p.bar[k - 1] = 0u;
    if (blockIdx.x == 0 && p.trace) p.trace[2 + 2 * k] = static_cast<long long>(globaltimer_ns());             // every CTA done
    __threadfence();
  }
  __syncthreads();
}

// ---- stage prologues: the stage's activations -> shared memory, bf16 [32][K] with a padded row stride ----
// row stride in bytes = K * 2 + 64: 16-byte fragment loads of 8 consecutive rows then fall into distinct banks
__device__ __forceinline__ int xs_stride(int K) { return K * 2 + 64; }

// Xs[r][:] = LayerNorm(src[r][:]) (fp32 [B][H], written by an earlier stage -> read through L2), optionally also written
// to xbuf (bf16 global: the residual operand of later stages) by the CTA whose index equals the row.
__device__ void fill_layernorm(const DsParams& p, unsigned char* Xs, const float* src, const bf16* gamma, const bf16* beta,
                               float eps, bf16* xout) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, stride = xs_stride(H);
  for (int r = warp; r < p.ng * 8; r += DS_WARPS) {
    bf16* dst = reinterpret_cast<bf16*>(Xs + static_cast<size_t>(r) * stride);
    if (r >= p.B) {
      for (int c = lane * 8; c < H; c += 256) *reinterpret_cast<uint4*>(dst + c) = make_uint4(0, 0, 0, 0);
      continue;
    }
    float v[8][4];  // H <= 1024: 8 float4 per lane
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = i * 128 + lane * 4;
      if (c < H) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(src + static_cast<size_t>(r) * H + c));
        v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
        sum += t.x + t.y + t.z + t.w;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / H;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i * 128 + lane * 4 < H) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d = v[i][j] - mean;
          sq += d * d;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / H + eps);
    const bool wr = xout != nullptr && static_cast<int>(blockIdx.x) == r;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = i * 128 + lane * 4;
      if (c < H) {
        const uint2 gr = *reinterpret_cast<const uint2*>(gamma + c);
        const uint2 br = *reinterpret_cast<const uint2*>(beta + c);
        const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&gr);
        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&br);
        const float2 g01 = __bfloat1622float2(g2[0]), g23 = __bfloat1622float2(g2[1]);
        const float2 b01 = __bfloat1622float2(b2[0]), b23 = __bfloat1622float2(b2[1]);
        uint2 o;
        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
        o2[0] = __floats2bfloat162_rn((v[i][0] - mean) * rstd * g01.x + b01.x, (v[i][1] - mean) * rstd * g01.y + b01.y);
        o2[1] = __floats2bfloat162_rn((v[i][2] - mean) * rstd * g23.x + b23.x, (v[i][3] - mean) * rstd * g23.y + b23.y);
        *reinterpret_cast<uint2*>(dst + c) = o;
        if (wr) *reinterpret_cast<uint2*>(xout + static_cast<size_t>(r) * H + c) = o;
      }
    }
  }
}

// Xs[r][:] = src[r][:] (bf16 [B][K] global, written by an earlier stage)
__device__ void fill_copy(const DsParams& p, unsigned char* Xs, const bf16* src, int K) {
  const int stride = xs_stride(K), per_row = K >> 3;
  const int total = p.ng * 8 * per_row;
  for (int i = threadIdx.x; i < total; i += DS_THREADS) {
    const int r = i / per_row, c = (i - r * per_row) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < p.B) v = __ldcg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(r) * K + c));
    *reinterpret_cast<uint4*>(Xs + static_cast<size_t>(r) * stride + c * 2) = v;
  }
}

// Xs[r][:] = emb[tok[r]][:] (+ pos_table[pos][:]); also written to xbuf by CTA r
__device__ void fill_embedding(const DsParams& p, unsigned char* Xs, int pos) {
  const int H = p.H, stride = xs_stride(H), per_row = H >> 3;
  const int total = p.ng * 8 * per_row;
  for (int i = threadIdx.x; i < total; i += DS_THREADS) {
    const int r = i / per_row, c = (i - r * per_row) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < p.B) {
      const long long t = p.tok[r];
      v = *reinterpret_cast<const uint4*>(p.emb + static_cast<size_t>(t) * H + c);
      if (p.pos_table) {
        const uint4 q = *reinterpret_cast<const uint4*>(p.pos_table + static_cast<size_t>(pos) * H + c);
        __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&v);
        const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = __hadd2(a[j], b[j]);  // bf16 add: what `hidden_state + pos_info` does in a bf16 model
      }
      if (static_cast<int>(blockIdx.x) == r) *reinterpret_cast<uint4*>(p.xbuf + static_cast<size_t>(r) * H + c) = v;
    }
    *reinterpret_cast<uint4*>(Xs + static_cast<size_t>(r) * stride + c * 2) = v;
  }
}

// ---- GEMM stage, K split over the warps of a CTA ---------------------------------------------------
// out[n][f] = epi(sum_k Xs[n][k] W[f][k] + bias[f]) for f in [0, N), n in [0, B). Unit = 16 features.
enum { DS_EPI_F32 = 0, DS_EPI_F32_RESID = 1, DS_EPI_GELU_BF16 = 2, DS_EPI_GELU_F32 = 3 };

template <int KB_PER_WARP>  // 32-wide k-blocks per warp: K = 256 * KB_PER_WARP
__device__ void gemm_ksplit(const DsParams& p, const unsigned char* Xs, float* red, const bf16* W, const bf16* bias, int N, int epi,
                            void* out, long long ldo, const bf16* resid) {
  constexpr int K = 256 * KB_PER_WARP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int stride = K * 2 + 64;
  const int units = (N + 15) >> 4;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int f0 = u * 16;
    const bool ok0 = f0 + g < N, ok1 = f0 + g + 8 < N;
    const bf16* w0 = W + static_cast<size_t>(f0 + g) * K + warp * (32 * KB_PER_WARP) + t * 8;
    const bf16* w1 = w0 + static_cast<size_t>(8) * K;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // k-blocks whose weight loads are in flight together (2 x 16-byte loads per thread and block)
    constexpr int CH = KB_PER_WARP % 6 == 0 ? 6 : (KB_PER_WARP % 4 == 0 ? 4 : KB_PER_WARP);
    static_assert(KB_PER_WARP % CH == 0, "chunking");
#pragma unroll 1
    for (int c0 = 0; c0 < KB_PER_WARP; c0 += CH) {
      uint4 ra[CH], rb[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        ra[j] = ok0 ? ldg_stream(w0 + (c0 + j) * 32) : make_uint4(0, 0, 0, 0);
        rb[j] = ok1 ? ldg_stream(w1 + (c0 + j) * 32) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int kbyte = (warp * (32 * KB_PER_WARP) + (c0 + j) * 32 + t * 8) * 2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < p.ng) {
            const uint4 xb = *reinterpret_cast<const uint4*>(Xs + static_cast<size_t>(i * 8 + g) * stride + kbyte);
            mma_bf16(acc[i], ra[j].x, rb[j].x, ra[j].y, rb[j].y, xb.x, xb.y);
            mma_bf16(acc[i], ra[j].z, rb[j].z, ra[j].w, rb[j].w, xb.z, xb.w);
          }
        }
      }
    }
    // partial tile of this warp -> red[warp][feature r][token n]
    float* my = red + warp * DS_RED;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = i * 8 + t * 2;
      my[g * 33 + n] = acc[i][0];
      my[g * 33 + n + 1] = acc[i][1];
      my[(g + 8) * 33 + n] = acc[i][2];
      my[(g + 8) * 33 + n + 1] = acc[i][3];
    }
    __syncthreads();
    // 512 results, 2 per thread: features r, r + 1 of token n (fixed summation order over the warps)
    {
      const int idx = threadIdx.x * 2;
      const int n = idx >> 4, r = idx & 15;
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int w = 0; w < DS_WARPS; ++w) {
        v0 += red[w * DS_RED + r * 33 + n];
        v1 += red[w * DS_RED + (r + 1) * 33 + n];
      }
      const int f = f0 + r;
      if (n < p.B && f < N) {
        const bool two = f + 1 < N;
        if (bias) {
          v0 += __bfloat162float(bias[f]);
          if (two) v1 += __bfloat162float(bias[f + 1]);
        }
        if (epi == DS_EPI_GELU_BF16 || epi == DS_EPI_GELU_F32) {
          v0 = gelu_erf(v0);
          v1 = gelu_erf(v1);
        }
        if (epi == DS_EPI_F32_RESID) {
          v0 += __bfloat162float(resid[static_cast<size_t>(n) * ldo + f]);
          if (two) v1 += __bfloat162float(resid[static_cast<size_t>(n) * ldo + f + 1]);
        }
        if (epi == DS_EPI_GELU_BF16) {
          bf16* o = reinterpret_cast<bf16*>(out) + static_cast<size_t>(n) * ldo + f;
          if (two) *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(v0, v1);
          else *o = __float2bfloat16_rn(v0);
        } else {
          float* o = reinterpret_cast<float*>(out) + static_cast<size_t>(n) * ldo + f;
          if (two) *reinterpret_cast<float2*>(o) = make_float2(v0, v1);
          else *o = v0;
        }
      }
    }
    __syncthreads();
  }
}

__device__ void gemm_ksplit_dispatch(const DsParams& p, int K, const unsigned char* Xs, float* red, const bf16* W, const bf16* bias,
                                     int N, int epi, void* out, long long ldo, const bf16* resid) {
  switch (K >> 8) {
    case 1: gemm_ksplit<1>(p, Xs, red, W, bias, N, epi, out, ldo, resid); break;
    case 2: gemm_ksplit<2>(p, Xs, red, W, bias, N, epi, out, ldo, resid); break;
    case 3: gemm_ksplit<3>(p, Xs, red, W, bias, N, epi, out, ldo, resid); break;
    case 4: gemm_ksplit<4>(p, Xs, red, W, bias, N, epi, out, ldo, resid); break;
    case 8: gemm_ksplit<8>(p, Xs, red, W, bias, N, epi, out, ldo, resid); break;
    case 12: gemm_ksplit<12>(p, Xs, red, W, bias, N, epi, out, ldo, resid); break;
    default: gemm_ksplit<16>(p, Xs, red, W, bias, N, epi, out, ldo, resid); break;
  }
}

// ---- vocabulary projection + greedy argmax: a warp owns whole units (16 features x K), register double buffering ----
template <int KB>  // k-blocks of 32: K = 32 * KB, KB % 4 == 0
__device__ void lm_head_argmax(const DsParams& p, const unsigned char* Xs, unsigned long long* s_keys) {
  constexpr int K = 32 * KB;
  constexpr int CH = 4;          // k-blocks per chunk
  constexpr int NCH = KB / CH;   // chunks per unit
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int stride = K * 2 + 64;
  const int N = p.V;
  const int units = (N + 15) >> 4;
  const int gw = blockIdx.x * DS_WARPS + warp, nw = gridDim.x * DS_WARPS;
  unsigned long long best[4][2];  // per token group: tokens 8 i + 2 t, + 1
#pragma unroll
  for (int i = 0; i < 4; ++i) best[i][0] = best[i][1] = 0ull;
  if (gw < units) {
    const int my_units = (units - gw + nw - 1) / nw;
    const int total = my_units * NCH;
    uint4 ra[2][CH], rb[2][CH];
    auto issue = [&](int buf, int ci) {
      const int u = gw + (ci / NCH) * nw, c = ci % NCH;
      const int f0 = u * 16;
      const bf16* w0 = p.w_v + static_cast<size_t>(f0 + g) * K + c * (CH * 32) + t * 8;
      const bool ok0 = f0 + g < N, ok1 = f0 + g + 8 < N;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        ra[buf][j] = ok0 ? ldg_stream(w0 + j * 32) : make_uint4(0, 0, 0, 0);
        rb[buf][j] = ok1 ? ldg_stream(w0 + static_cast<size_t>(8) * K + j * 32) : make_uint4(0, 0, 0, 0);
      }
    };
    issue(0, 0);
    float acc[4][4];
#pragma unroll 1
    for (int ci = 0; ci < total; ci += 2) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {  // static register buffer indices
        const int cc = ci + half;
        if (cc >= total) break;
        if (cc + 1 < total) issue(half ^ 1, cc + 1);
        const int c = cc % NCH;
        if (c == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const int kbyte = (c * (CH * 32) + j * 32 + t * 8) * 2;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i < p.ng) {
              const uint4 xb = *reinterpret_cast<const uint4*>(Xs + static_cast<size_t>(i * 8 + g) * stride + kbyte);
              mma_bf16(acc[i], ra[half][j].x, rb[half][j].x, ra[half][j].y, rb[half][j].y, xb.x, xb.y);
              mma_bf16(acc[i], ra[half][j].z, rb[half][j].z, ra[half][j].w, rb[half][j].w, xb.z, xb.w);
            }
          }
        }
        if (c == NCH - 1) {  // unit finished: bias, round to the model dtype (what the logits tensor holds), fold into the running best
          const int f0 = (gw + (cc / NCH) * nw) * 16;
          const int fa = f0 + g, fb = f0 + g + 8;
          const float ba = (p.b_v && fa < N) ? __bfloat162float(p.b_v[fa]) : 0.f;
          const float bb = (p.b_v && fb < N) ? __bfloat162float(p.b_v[fb]) : 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i < p.ng) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int n = i * 8 + t * 2 + e;
                const bf16 la = __float2bfloat16_rn(acc[i][e] + ba), lb = __float2bfloat16_rn(acc[i][2 + e] + bb);
                if (fa < N) {
                  const unsigned long long key = (static_cast<unsigned long long>(ds_order_bits(__bfloat162float(la))) << 32) | (0xffffffffu - static_cast<unsigned int>(fa));
                  best[i][e] = key > best[i][e] ? key : best[i][e];
                  if (p.logits && n < p.B) reinterpret_cast<bf16*>(p.logits)[static_cast<size_t>(n) * p.ld_logits + fa] = la;
                }
                if (fb < N) {
                  const unsigned long long key = (static_cast<unsigned long long>(ds_order_bits(__bfloat162float(lb))) << 32) | (0xffffffffu - static_cast<unsigned int>(fb));
                  best[i][e] = key > best[i][e] ? key : best[i][e];
                  if (p.logits && n < p.B) reinterpret_cast<bf16*>(p.logits)[static_cast<size_t>(n) * p.ld_logits + fb] = lb;
                }
              }
            }
          }
        }
      }
    }
  }
  // fold the 8 feature lanes (g) of every token, then one atomic per token and warp
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      unsigned long long k = best[i][e];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
        k = other > k ? other : k;
      }
      const int n = i * 8 + t * 2 + e;
      if (g == 0) s_keys[warp * 32 + n] = k;
    }
  __syncthreads();
  if (threadIdx.x < p.B) {  // one atomic per token and CTA
    unsigned long long k = 0ull;
#pragma unroll
    for (int w = 0; w < DS_WARPS; ++w) {
      const unsigned long long o = s_keys[w * 32 + threadIdx.x];
      k = o > k ? o : k;
    }
    if (k != 0ull) atomicMax(&p.amax[threadIdx.x], k);
  }
}

// ---- attention over the cache: one work item = (kv-split, kv head, batch row), all 8 warps of the CTA ----
template <int NREP>
__device__ void attention_items(const DsParams& p, const DsLayer& ly, int sp, unsigned char* smem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = lane & 7, lk = lane >> 3;
  float* s_newk = reinterpret_cast<float*>(smem);             // [64]
  float* s_newv = s_newk + DS_HD;                              // [64]
  float* s_q = s_newv + DS_HD;                                 // [NREP][64]
  float* s_red = s_q + 8 * DS_HD;                              // [8 warps][NREP][66]
  unsigned int* s_last = reinterpret_cast<unsigned int*>(s_red + DS_WARPS * 8 * (DS_HD + 2));
  const float scale_log2 = 1.4426950408889634f / 8.0f;
  const int items = p.B * p.Hkv * p.splits;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int split = it % p.splits, kvh = (it / p.splits) % p.Hkv, b = it / (p.splits * p.Hkv);
    // new token: projections (bias already added) -> RoPE(q, k) -> smem; the last split appends k, v to the cache
    const float* row = p.qkv + static_cast<size_t>(b) * p.NQKV;
    for (int idx = threadIdx.x; idx < (NREP + 2) * DS_HD; idx += DS_THREADS) {
      const int which = idx / DS_HD, j = idx % DS_HD;
      int col;
      if (which < NREP) col = (kvh * NREP + which) * DS_HD;
      else if (which == NREP) col = (p.Hq + kvh) * DS_HD;
      else col = (p.Hq + p.Hkv + kvh) * DS_HD;
      // the reference's bf16 model holds q / k / v as bf16 tensors before the rotation: same rounding point
      float x = __bfloat162float(__float2bfloat16_rn(__ldcg(row + col + j)));
      if (which <= NREP && p.rope_cos) {
        const float other = __bfloat162float(__float2bfloat16_rn(__ldcg(row + col + (j < 32 ? j + 32 : j - 32))));
        const float c = p.rope_cos[sp * 32 + (j & 31)], s = p.rope_sin[sp * 32 + (j & 31)];
        x = j < 32 ? x * c - other * s : x * c + other * s;
      }
      if (which < NREP) s_q[which * DS_HD + j] = x;
      else if (which == NREP) s_newk[j] = x;
      else s_newv[j] = x;
    }
    __syncthreads();
    bf16* kc = ly.k_cache + b * p.c_sb + kvh * p.c_sh;
    bf16* vc = ly.v_cache + b * p.c_sb + kvh * p.c_sh;
    if (split == p.splits - 1 && threadIdx.x < DS_HD) {
      kc[sp * p.c_sl + threadIdx.x] = __float2bfloat16_rn(s_newk[threadIdx.x]);
      vc[sp * p.c_sl + threadIdx.x] = __float2bfloat16_rn(s_newv[threadIdx.x]);
    }
    float q[NREP][8];
#pragma unroll
    for (int r = 0; r < NREP; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) q[r][j] = s_q[r * DS_HD + ld * 8 + j] * scale_log2;
    float m[NREP], l[NREP], o[NREP][8];
#pragma unroll
    for (int r = 0; r < NREP; ++r) {
      m[r] = -INFINITY;
      l[r] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[r][j] = 0.f;
    }
    const int per = (sp + p.splits - 1) / p.splits;
    const int k_begin = split * per;
    const int k_end = min(sp, k_begin + per);
    constexpr int UN = 4;
    constexpr int KEYS_PER_ITER = DS_WARPS * 4 * UN;
    for (int k0 = k_begin; k0 < k_end; k0 += KEYS_PER_ITER) {
      uint4 kraw[UN], vraw[UN];
      int kidx[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        kidx[u] = k0 + (u * DS_WARPS + warp) * 4 + lk;
        if (kidx[u] < k_end) {
          const long long off = kidx[u] * p.c_sl + ld * 8;
          kraw[u] = ldg_stream(kc + off);
          vraw[u] = ldg_stream(vc + off);
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const bool valid = kidx[u] < k_end;
        float kv[8], vv[8];
        {
          const __nv_bfloat162* hk = reinterpret_cast<const __nv_bfloat162*>(&kraw[u]);
          const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&vraw[u]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 a = __bfloat1622float2(hk[j]), c = __bfloat1622float2(hv[j]);
            kv[2 * j] = a.x; kv[2 * j + 1] = a.y;
            vv[2 * j] = c.x; vv[2 * j + 1] = c.y;
          }
        }
#pragma unroll
        for (int r = 0; r < NREP; ++r) {
          float s = 0.f;
          if (valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s += q[r][j] * kv[j];
          }
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          if (valid) {
            const float mn = fmaxf(m[r], s);
            const float a = exp2f(m[r] - mn);
            const float pr = exp2f(s - mn);
            l[r] = l[r] * a + pr;
#pragma unroll
            for (int j = 0; j < 8; ++j) o[r][j] = o[r][j] * a + pr * vv[j];
            m[r] = mn;
          }
        }
      }
    }
    if (split == p.splits - 1 && warp == 0 && lk == 0) {  // the new token, from smem (unrounded, like the eager kernel)
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
        const float pr = exp2f(s - mn);
        l[r] = l[r] * a + pr;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[r][j] = o[r][j] * a + pr * s_newv[ld * 8 + j];
        m[r] = mn;
      }
    }
    // merge lane groups (xor 8, 16), then warps through smem
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
        float* dst = s_red + (warp * 8 + r) * (DS_HD + 2);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[ld * 8 + j] = o[r][j];
        if (ld == 0) {
          dst[DS_HD] = m[r];
          dst[DS_HD + 1] = l[r];
        }
      }
    }
    __syncthreads();
    const int tdx = threadIdx.x;
    const int r_own = tdx / DS_HD, j_own = tdx % DS_HD;
    for (int rr = r_own; rr < NREP; rr += DS_THREADS / DS_HD) {
      float M = -INFINITY, Lsum = 0.f, O = 0.f;
#pragma unroll
      for (int w = 0; w < DS_WARPS; ++w) {
        const float* src = s_red + (w * 8 + rr) * (DS_HD + 2);
        const float mw = src[DS_HD], lw = src[DS_HD + 1], ow = src[j_own];
        if (mw == -INFINITY) continue;
        const float mn = fmaxf(M, mw);
        const float a1 = (M == -INFINITY) ? 0.f : exp2f(M - mn);
        const float a2 = exp2f(mw - mn);
        Lsum = Lsum * a1 + lw * a2;
        O = O * a1 + ow * a2;
        M = mn;
      }
      if (p.splits == 1) {
        p.attn[static_cast<size_t>(b) * (p.Hq * DS_HD) + (kvh * NREP + rr) * DS_HD + j_own] = __float2bfloat16_rn(O / Lsum);
      } else {
        float* w = p.part + ((((static_cast<size_t>(b) * p.Hkv + kvh) * p.splits + split) * NREP + rr) * (DS_HD + 2));
        w[j_own] = O;
        if (j_own == 0) {
          w[DS_HD] = M;
          w[DS_HD + 1] = Lsum;
        }
      }
    }
    if (p.splits > 1) {  // last CTA of this (row, kv head) combines the splits
      __threadfence();
      __syncthreads();
      if (tdx == 0) {
        const unsigned int prev = atomicAdd(&p.tickets[b * p.Hkv + kvh], 1u);
        *s_last = (prev == static_cast<unsigned int>(p.splits - 1)) ? 1u : 0u;
        if (*s_last) p.tickets[b * p.Hkv + kvh] = 0u;
      }
      __syncthreads();
      if (*s_last) {
        __threadfence();
        for (int rr = r_own; rr < NREP; rr += DS_THREADS / DS_HD) {
          float M = -INFINITY, Lsum = 0.f, O = 0.f;
          for (int si = 0; si < p.splits; ++si) {
            const float* w = p.part + ((((static_cast<size_t>(b) * p.Hkv + kvh) * p.splits + si) * NREP + rr) * (DS_HD + 2));
            const float mw = __ldcg(w + DS_HD), lw = __ldcg(w + DS_HD + 1), ow = __ldcg(w + j_own);
            if (mw == -INFINITY) continue;
            const float mn = fmaxf(M, mw);
            const float a1 = (M == -INFINITY) ? 0.f : exp2f(M - mn);
            const float a2 = exp2f(mw - mn);
            Lsum = Lsum * a1 + lw * a2;
            O = O * a1 + ow * a2;
            M = mn;
          }
          p.attn[static_cast<size_t>(b) * (p.Hq * DS_HD) + (kvh * NREP + rr) * DS_HD + j_own] = __float2bfloat16_rn(O / Lsum);
        }
      }
    }
    __syncthreads();  // smem is reused by the next item
  }
}

template <int NREP>
__global__ void __launch_bounds__(DS_THREADS, 1)
decode_step_kernel(const __grid_constant__ DsParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int Kmax = p.FF > p.H ? p.FF : p.H;
  unsigned char* Xs = smem;
  float* red = reinterpret_cast<float*>(smem + static_cast<size_t>(32) * xs_stride(Kmax));
  const int pos = *p.pos;  // read before anything can change it (FIN runs after the last barrier)
  int bar = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.trace) p.trace[0] = static_cast<long long>(globaltimer_ns());
  if (blockIdx.x == 0 && threadIdx.x == 0) p.bar[5 * p.L + 1] = 0u;  // the previous launch's last barrier (5 L + 2 per launch)
  const int qcols = p.Hq * DS_HD;

  for (int l = 0; l < p.L; ++l) {
    const DsLayer& ly = p.layer[l];
    // QKV
    const int cta = static_cast<int>(blockIdx.x);  // CTAs that own no unit of a stage skip its prologue
    if (cta < (p.NQKV + 15) / 16) {
      if (l == 0) fill_embedding(p, Xs, pos);
      else fill_layernorm(p, Xs, p.s2, p.layer[l - 1].ln2_g, p.layer[l - 1].ln2_b, p.eps_layer, p.xbuf);
    }
    __syncthreads();
    gemm_ksplit_dispatch(p, p.H, Xs, red, ly.w_qkv, ly.b_qkv, p.NQKV, DS_EPI_F32, p.qkv, p.NQKV, nullptr);
    grid_barrier(p, bar++);
    // ATTN
    attention_items<NREP>(p, ly, pos, smem);
    grid_barrier(p, bar++);
    // OUT: s1 = attn Wo^T + bo + x
    if (cta < (p.H + 15) / 16) fill_copy(p, Xs, p.attn, qcols);
    __syncthreads();
    gemm_ksplit_dispatch(p, qcols, Xs, red, ly.w_o, ly.b_o, p.H, DS_EPI_F32_RESID, p.s1, p.H, p.xbuf);
    grid_barrier(p, bar++);
    // FFN1: a = gelu(LN1(s1) W1^T + b1)
    if (cta < (p.FF + 15) / 16) fill_layernorm(p, Xs, p.s1, ly.ln1_g, ly.ln1_b, p.eps_layer, nullptr);
    __syncthreads();
    gemm_ksplit_dispatch(p, p.H, Xs, red, ly.w_1, ly.b_1, p.FF, DS_EPI_GELU_BF16, p.abuf, p.FF, nullptr);
    grid_barrier(p, bar++);
    // FFN2: s2 = a W2^T + b2 + x   (the residual is the layer INPUT: quirk Q2)
    if (cta < (p.H + 15) / 16) fill_copy(p, Xs, p.abuf, p.FF);
    __syncthreads();
    gemm_ksplit_dispatch(p, p.FF, Xs, red, ly.w_2, ly.b_2, p.H, DS_EPI_F32_RESID, p.s2, p.H, p.xbuf);
    grid_barrier(p, bar++);
  }
  // LM head: g = gelu(x Wd^T + bd), x = LN2(s2) of the last layer
  if (static_cast<int>(blockIdx.x) < (p.H + 15) / 16) fill_layernorm(p, Xs, p.s2, p.layer[p.L - 1].ln2_g, p.layer[p.L - 1].ln2_b, p.eps_layer, nullptr);
  __syncthreads();
  gemm_ksplit_dispatch(p, p.H, Xs, red, p.w_d, p.b_d, p.H, DS_EPI_GELU_F32, p.gbuf, p.H, nullptr);
  grid_barrier(p, bar++);
  // logits = LN(g) Wv^T + bv -> argmax
  fill_layernorm(p, Xs, p.gbuf, p.lnh_g, p.lnh_b, p.eps_head, nullptr);
  __syncthreads();
  switch (p.H >> 5) {
    case 8: lm_head_argmax<8>(p, Xs, reinterpret_cast<unsigned long long*>(red)); break;
    case 16: lm_head_argmax<16>(p, Xs, reinterpret_cast<unsigned long long*>(red)); break;
    case 24: lm_head_argmax<24>(p, Xs, reinterpret_cast<unsigned long long*>(red)); break;
    default: lm_head_argmax<32>(p, Xs, reinterpret_cast<unsigned long long*>(red)); break;
  }
  grid_barrier(p, bar++);
  // FIN: next token = unpacked argmax; it is the input of the next step and lands in tokens[:, pos + 1]
  if (blockIdx.x == 0) {
    if (threadIdx.x < p.B) {
      const unsigned long long key = __ldcg(&p.amax[threadIdx.x]);
      const long long idx = key == 0ull ? 0ll : static_cast<long long>(0xffffffffu - static_cast<unsigned int>(key & 0xffffffffull));
      p.tok[threadIdx.x] = idx;
      if (p.tokens_out) p.tokens_out[threadIdx.x * p.ld_tokens + pos + 1] = idx;
      p.amax[threadIdx.x] = 0ull;
    }
    if (threadIdx.x == 0) *p.pos = pos + 1;
  }
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct DsScratch {
  size_t xbuf, qkv, attn, s1, abuf, s2, gbuf, part, tickets, amax, bar, abort_flag, total;
};
static DsScratch scratch_layout(int B, int H, int Hq, int Hkv, int FF, int splits) {
  DsScratch s;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
  const int nrep = Hq / Hkv;
  s.bar = take(sizeof(unsigned int) * DS_MAX_BARRIERS);
  s.abort_flag = take(sizeof(int));
  s.tickets = take(sizeof(unsigned int) * B * Hkv);
  s.amax = take(sizeof(unsigned long long) * DS_MAXB);
  s.xbuf = take(sizeof(bf16) * B * H);
  s.qkv = take(sizeof(float) * B * (Hq + 2 * Hkv) * DS_HD);
  s.attn = take(sizeof(bf16) * B * Hq * DS_HD);
  s.s1 = take(sizeof(float) * B * H);
  s.abuf = take(sizeof(bf16) * B * FF);
  s.s2 = take(sizeof(float) * B * H);
  s.gbuf = take(sizeof(float) * B * H);
  s.part = take(sizeof(float) * B * Hkv * splits * nrep * (DS_HD + 2));
  s.total = o;
  return s;
}
constexpr int DS_MAX_SPLITS = 16;

}  // namespace vy

using namespace vy;

extern "C" int64_t vy_decode_step_workspace_bytes(int B, int H, int n_q_heads, int n_kv_heads, int ffn) {
  if (B <= 0 || H <= 0 || n_q_heads <= 0 || n_kv_heads <= 0 || ffn <= 0) return -1;
  return static_cast<int64_t>(scratch_layout(B, H, n_q_heads, n_kv_heads, ffn, DS_MAX_SPLITS).total);
}

extern "C" int vy_decode_step_status(const void* workspace) {
  if (!workspace) return -1;
  const DsScratch s = scratch_layout(1, 8, 1, 1, 8, 1);  // the flag's offset does not depend on the shape
  int v = -1;
  if (cudaMemcpy(&v, static_cast<const unsigned char*>(workspace) + s.abort_flag, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  return v;
}

extern "C" int vy_decode_step(const VyDecodeStep* q) {
  VY_CHECK_ARG(q != nullptr, "vy_decode_step: null params");
  if (!vy_device_ok()) {
    set_error("vy_decode_step: no sm_100 device (there is no CPU fallback)");
    return VY_ERR_NO_DEVICE;
  }
  const int B = q->B, H = q->H, Hq = q->n_q_heads, Hkv = q->n_kv_heads, FF = q->ffn, L = q->n_layers;
  VY_CHECK_ARG(B >= 1 && B <= DS_MAXB, "vy_decode_step: batch %d outside [1, %d]", B, DS_MAXB);
  VY_CHECK_ARG(q->head_dim == DS_HD && Hq > 0 && Hkv > 0 && Hq % Hkv == 0, "vy_decode_step: head_dim must be 64, n_q %% n_kv == 0");
  VY_CHECK_ARG(H == Hq * DS_HD, "vy_decode_step: hidden size %d != n_q_heads * 64", H);
  VY_CHECK_ARG(H % 256 == 0 && H <= 1024 && FF % 256 == 0 && FF <= 4096, "vy_decode_step: H (%d) and ffn (%d) must be multiples of 256, H <= 1024, ffn <= 4096", H, FF);
  {
    const int hb = H >> 8, fb = FF >> 8;
    auto okk = [](int kb) { return kb == 1 || kb == 2 || kb == 3 || kb == 4 || kb == 8 || kb == 12 || kb == 16; };
    VY_CHECK_ARG(okk(hb) && okk(fb), "vy_decode_step: unsupported H / ffn (%d / %d)", H, FF);
  }
  VY_CHECK_ARG(L >= 1 && L <= VY_DECODE_MAX_LAYERS, "vy_decode_step: n_layers %d outside [1, %d]", L, VY_DECODE_MAX_LAYERS);
  VY_CHECK_ARG(q->vocab > 0 && q->emb && q->w_d && q->ln_head_g && q->ln_head_b && q->w_v && q->pos && q->tok && q->workspace,
               "vy_decode_step: null pointer");
  VY_CHECK_ARG((q->rope_cos == nullptr) == (q->rope_sin == nullptr), "vy_decode_step: rope tables must both be set or NULL");
  VY_CHECK_ARG(q->cache_len > 0 && q->pos_bound >= 0 && q->pos_bound < q->cache_len, "vy_decode_step: pos_bound %d outside the cache (%d)",
               q->pos_bound, q->cache_len);
  VY_CHECK_ARG(!q->rope_cos || q->rope_rows <= 0 || q->pos_bound < q->rope_rows, "vy_decode_step: position bound %d exceeds the %d rows of the RoPE tables",
               q->pos_bound, q->rope_rows);
  VY_CHECK_ARG((q->cache_sb % 8) == 0 && (q->cache_sh % 8) == 0 && (q->cache_sl % 8) == 0, "vy_decode_step: cache strides must keep 16-byte alignment");
  const int nrep = Hq / Hkv;
  VY_CHECK_ARG(nrep == 1 || nrep == 2 || nrep == 3 || nrep == 4 || nrep == 6 || nrep == 8, "vy_decode_step: unsupported q-heads per kv-head %d", nrep);

  const int sms = num_sms();
  int splits = (2 * sms + B * Hkv - 1) / (B * Hkv);
  const int by_len = (q->pos_bound + 63) / 64;
  if (splits > by_len) splits = by_len;
  if (splits < 1) splits = 1;
  if (splits > DS_MAX_SPLITS) splits = DS_MAX_SPLITS;
  const DsScratch sc = scratch_layout(B, H, Hq, Hkv, FF, splits);
  VY_CHECK_ARG(q->workspace_bytes >= static_cast<int64_t>(scratch_layout(B, H, Hq, Hkv, FF, DS_MAX_SPLITS).total),
               "vy_decode_step: workspace too small (vy_decode_step_workspace_bytes)");
  VY_CHECK_ARG((reinterpret_cast<uintptr_t>(q->workspace) & 255) == 0, "vy_decode_step: workspace must be 256-byte aligned");

  DsParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.Hq = Hq; p.Hkv = Hkv; p.FF = FF; p.V = q->vocab; p.L = L; p.NQKV = (Hq + 2 * Hkv) * DS_HD;
  p.ng = (B + 7) / 8;
  p.cache_len = q->cache_len; p.splits = splits;
  p.c_sb = q->cache_sb; p.c_sh = q->cache_sh; p.c_sl = q->cache_sl;
  p.eps_layer = q->eps_layer; p.eps_head = q->eps_head;
  p.emb = static_cast<const bf16*>(q->emb);
  p.pos_table = static_cast<const bf16*>(q->pos_table);
  p.rope_cos = q->rope_cos; p.rope_sin = q->rope_sin;
  p.w_d = static_cast<const bf16*>(q->w_d); p.b_d = static_cast<const bf16*>(q->b_d);
  p.lnh_g = static_cast<const bf16*>(q->ln_head_g); p.lnh_b = static_cast<const bf16*>(q->ln_head_b);
  p.w_v = static_cast<const bf16*>(q->w_v); p.b_v = static_cast<const bf16*>(q->b_v);
  p.pos = q->pos; p.tok = reinterpret_cast<long long*>(q->tok);
  p.tokens_out = reinterpret_cast<long long*>(q->tokens_out); p.ld_tokens = q->ld_tokens;
  p.logits = q->logits; p.ld_logits = q->ld_logits;
  unsigned char* ws = static_cast<unsigned char*>(q->workspace);
  p.xbuf = reinterpret_cast<bf16*>(ws + sc.xbuf);
  p.qkv = reinterpret_cast<float*>(ws + sc.qkv);
  p.attn = reinterpret_cast<bf16*>(ws + sc.attn);
  p.s1 = reinterpret_cast<float*>(ws + sc.s1);
  p.abuf = reinterpret_cast<bf16*>(ws + sc.abuf);
  p.s2 = reinterpret_cast<float*>(ws + sc.s2);
  p.gbuf = reinterpret_cast<float*>(ws + sc.gbuf);
  p.part = reinterpret_cast<float*>(ws + sc.part);
  p.tickets = reinterpret_cast<unsigned int*>(ws + sc.tickets);
  p.amax = reinterpret_cast<unsigned long long*>(ws + sc.amax);
  p.bar = reinterpret_cast<unsigned int*>(ws + sc.bar);
  p.abort_flag = reinterpret_cast<int*>(ws + sc.abort_flag);
  p.trace = reinterpret_cast<long long*>(q->trace);
  for (int l = 0; l < L; ++l) {
    const VyDecodeLayer& s = q->layers[l];
    VY_CHECK_ARG(s.w_qkv && s.w_o && s.ln1_g && s.ln1_b && s.w_1 && s.w_2 && s.ln2_g && s.ln2_b && s.k_cache && s.v_cache,
                 "vy_decode_step: layer %d has a null pointer", l);
    const void* al[] = {s.w_qkv, s.w_o, s.w_1, s.w_2, s.k_cache, s.v_cache};
    for (const void* a : al) VY_CHECK_ARG((reinterpret_cast<uintptr_t>(a) & 15) == 0, "vy_decode_step: layer %d: weights / caches must be 16-byte aligned", l);
    DsLayer& d = p.layer[l];
    d.w_qkv = static_cast<const bf16*>(s.w_qkv); d.b_qkv = static_cast<const bf16*>(s.b_qkv);
    d.w_o = static_cast<const bf16*>(s.w_o); d.b_o = static_cast<const bf16*>(s.b_o);
    d.ln1_g = static_cast<const bf16*>(s.ln1_g); d.ln1_b = static_cast<const bf16*>(s.ln1_b);
    d.w_1 = static_cast<const bf16*>(s.w_1); d.b_1 = static_cast<const bf16*>(s.b_1);
    d.w_2 = static_cast<const bf16*>(s.w_2); d.b_2 = static_cast<const bf16*>(s.b_2);
    d.ln2_g = static_cast<const bf16*>(s.ln2_g); d.ln2_b = static_cast<const bf16*>(s.ln2_b);
    d.k_cache = static_cast<bf16*>(s.k_cache); d.v_cache = static_cast<bf16*>(s.v_cache);
  }
  const int Kmax = FF > H ? FF : H;
  const size_t smem = static_cast<size_t>(32) * (Kmax * 2 + 64) + DS_WARPS * DS_RED * sizeof(float);
  const size_t attn_smem = (2 * DS_HD + 8 * DS_HD + DS_WARPS * 8 * (DS_HD + 2)) * sizeof(float) + 16;
  VY_CHECK_ARG(attn_smem <= smem, "vy_decode_step: internal smem layout");
  cudaStream_t st = static_cast<cudaStream_t>(q->stream);
  int grid = sms;
  static const int margin = getenv("VY_DECODE_SM_MARGIN") ? atoi(getenv("VY_DECODE_SM_MARGIN")) : 0;
  if (margin > 0 && grid - margin >= 32) grid -= margin;
  VY_CHECK_ARG(grid >= DS_MAXB, "vy_decode_step: needs at least %d SMs", DS_MAXB);
#define DS_LAUNCH(NR)                                                                                        \
  do {                                                                                                       \
    auto kern = decode_step_kernel<NR>;                                                                      \
    static std::once_flag once;                                                                              \
    static cudaError_t attr_rc = cudaSuccess;                                                                \
    std::call_once(once, [&] { attr_rc = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }); \
    VY_CUDA_OK(attr_rc);                                                                                     \
    cudaLaunchConfig_t cfg;                                                                                  \
    memset(&cfg, 0, sizeof(cfg));                                                                            \
    cfg.gridDim = dim3(grid);                                                                                \
    cfg.blockDim = dim3(DS_THREADS);                                                                         \
    cfg.dynamicSmemBytes = smem;                                                                             \
    cfg.stream = st;                                                                                         \
    cudaLaunchAttribute attr[1];                                                                             \
    attr[0].id = cudaLaunchAttributeCooperative; /* every CTA resident, or the launch fails: the barriers cannot hang */ \
    attr[0].val.cooperative = 1;                                                                             \
    cfg.attrs = attr;                                                                                        \
    cfg.numAttrs = 1;                                                                                        \
    VY_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));                                                           \
  } while (0)
  switch (nrep) {
    case 1: DS_LAUNCH(1); break;
    case 2: DS_LAUNCH(2); break;
    case 3: DS_LAUNCH(3); break;
    case 4: DS_LAUNCH(4); break;
    case 6: DS_LAUNCH(6); break;
    default: DS_LAUNCH(8); break;
  }
#undef DS_LAUNCH
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}
