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

constexpr int DS_THREADS = 512;  // 16 warps: four per scheduler — the stages are latency chains, so issue-level parallelism is what they lack
constexpr int DS_WARPS = 16;
constexpr int DS_MAXB = 32;
constexpr int DS_HD = 64;
constexpr int DS_MAX_BARRIERS = 5 * VY_DECODE_MAX_LAYERS + 8;
constexpr int DS_RED = 16 * 32;  // one warp's partial tile: [16 features][32 tokens], token index XOR-swizzled by the feature

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
  bf16* xbuf;    // [B][H]   layer input x (= LN2 of the previous layer's sum)
  bf16* ybuf;    // [B][H]   LN1(s1) of the current layer / LN of the LM head
  unsigned int* ln_ticket;    // last-CTA election of the stages that end in a LayerNorm
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
// DRAM -> L2 prefetch. Measured on this part: one dependent round of loads costs ~0.5 us when it hits L2 and 2-2.5 us from
// DRAM at this working set (tools/microbench/lat_micro.cu) — five stages of two or three dependent rounds per layer is what
// made the per-op decode step ten times slower than its byte count. Weights and cached keys / values do not depend on the
// step's activations, so every CTA asks L2 for its share of the NEXT layer's bytes a whole layer ahead, one
// `prefetch.global.L2` per 128-byte line spread over all threads of the grid (through the load / store unit: bulk prefetches
// queue in the copy engine in front of the activation copies and delayed them).
__device__ __forceinline__ void prefetch_line_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// lines [0, bytes / 128) of a range, split evenly over (nparts * 512) threads; this CTA is part `part`
__device__ __forceinline__ void prefetch_range_l2(const void* base, long long bytes, int part, int nparts) {
  const long long lines = (bytes + 127) >> 7;
  const long long per = (lines + nparts - 1) / nparts;
  const long long lo = per * part, hi = lo + per < lines ? lo + per : lines;
  const char* c = static_cast<const char*>(base);
  for (long long i = lo + threadIdx.x; i < hi; i += DS_THREADS) prefetch_line_l2(c + (i << 7));
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
  __syncthreads();  // CTA-scope order of every thread's writes before thread 0's gpu-scope release (cumulativity)
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0 && p.trace) p.trace[1 + 2 * k] = static_cast<long long>(globaltimer_ns());  // CTA 0 done with the stage
    // one release-atomic instead of a full memory barrier + atomic: a gpu-scope fence costs ~1 us on this part
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&p.bar[k]) : "memory");
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
  }
  __syncthreads();  // thread 0's acquire + this barrier: the other CTAs' writes are visible to the whole CTA (read through L2)
}

// development: fine-grained %globaltimer marks of CTA 0 / thread 0 inside selected stages (trace[128 + slot])
__device__ __forceinline__ void ds_mark(const DsParams& p, int slot) {
  if (slot >= 0 && p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[128 + slot] = static_cast<long long>(globaltimer_ns());
}
__device__ __forceinline__ void ds_mark_any(const DsParams& p, int slot) {  // whichever CTA gets here (thread 0)
  if (slot >= 0 && p.trace && threadIdx.x == 0) p.trace[128 + slot] = static_cast<long long>(globaltimer_ns());
}

// ---- shared-memory plumbing ---------------------------------------------------------------------
// A stage's activations ([B <= 32][K] bf16) are brought into shared memory by the copy engine: ONE bulk copy per row
// (cp.async.bulk, mbarrier tx-count) issued by the lanes of warp 0 — measured 1.1 us for 196 KB per SM against 3.6 us for
// 16-byte loads + stores by 512 threads, and no instruction stream to fetch. Row stride = K * 2 + 64 bytes, so that the
// 16-byte fragment loads of 8 consecutive rows fall into distinct banks.
__device__ __forceinline__ int xs_stride(int K) { return K * 2 + 64; }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned int bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait_spin(unsigned long long* bar, unsigned int parity) {
  while (!mbar_try_wait(reinterpret_cast<uint64_t*>(bar), parity)) {
  }
}

enum { DS_IN_COPY = 0, DS_IN_EMBED = 2 };
enum { DS_EPI_F32 = 0, DS_EPI_F32_RESID = 1, DS_EPI_GELU_BF16 = 2, DS_EPI_GELU_F32 = 3 };

struct StageIn {
  int kind;            // DS_IN_COPY: bf16 [B][K] rows in global memory; DS_IN_EMBED: gather of the tokens' embedding rows
  const void* src;
  bf16* xout;          // EMBED: also publish the rows (bf16) here — the residual operand of later stages
};

// LayerNorm that FOLLOWS a stage: the stage's output is an fp32 pre-norm sum whose rows are spread over many CTAs. Every
// CTA of the stage takes a ticket when its units are done; the last one normalises all rows (two per warp) and writes them
// as bf16 — the next stage then only bulk-copies them. Doing the LayerNorm redundantly in every consuming CTA cost more
// instructions per step than all the GEMMs together.
struct PostLN {
  const float* src;    // [B][H] fp32, written by this stage (null: no LayerNorm after this stage)
  const bf16* gamma;
  const bf16* beta;
  float eps;
  bf16* dst;           // [B][H] bf16
  unsigned int* ticket;
};

__device__ __noinline__ void layernorm_rows(const DsParams& p, const PostLN& ln) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H;
  for (int r0 = warp; r0 < p.B; r0 += 2 * DS_WARPS) {
    float v[2][8][4];  // H <= 1024: 8 float4 per lane and row
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = r0 + u * DS_WARPS;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = i * 128 + lane * 4;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < H && r < p.B) t = __ldcg(reinterpret_cast<const float4*>(ln.src + static_cast<size_t>(r) * H + c));
        v[u][i][0] = t.x; v[u][i][1] = t.y; v[u][i][2] = t.z; v[u][i][3] = t.w;
      }
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i * 128 + lane * 4 < H) sum[u] += v[u][i][0] + v[u][i][1] + v[u][i][2] + v[u][i][3];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum[0] += __shfl_xor_sync(0xffffffffu, sum[0], o);
      sum[1] += __shfl_xor_sync(0xffffffffu, sum[1], o);
    }
    const float mean[2] = {sum[0] / H, sum[1] / H};
    float sq[2] = {0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i * 128 + lane * 4 < H) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float d = v[u][i][j] - mean[u];
            sq[u] += d * d;
          }
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sq[0] += __shfl_xor_sync(0xffffffffu, sq[0], o);
      sq[1] += __shfl_xor_sync(0xffffffffu, sq[1], o);
    }
    const float rstd[2] = {rsqrtf(sq[0] / H + ln.eps), rsqrtf(sq[1] / H + ln.eps)};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = i * 128 + lane * 4;
      if (c < H) {
        const uint2 gr = *reinterpret_cast<const uint2*>(ln.gamma + c);
        const uint2 br = *reinterpret_cast<const uint2*>(ln.beta + c);
        const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&gr);
        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&br);
        const float2 g01 = __bfloat1622float2(g2[0]), g23 = __bfloat1622float2(g2[1]);
        const float2 b01 = __bfloat1622float2(b2[0]), b23 = __bfloat1622float2(b2[1]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int r = r0 + u * DS_WARPS;
          if (r >= p.B) continue;
          uint2 o;
          __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
          o2[0] = __floats2bfloat162_rn((v[u][i][0] - mean[u]) * rstd[u] * g01.x + b01.x, (v[u][i][1] - mean[u]) * rstd[u] * g01.y + b01.y);
          o2[1] = __floats2bfloat162_rn((v[u][i][2] - mean[u]) * rstd[u] * g23.x + b23.x, (v[u][i][3] - mean[u]) * rstd[u] * g23.y + b23.y);
          *reinterpret_cast<uint2*>(ln.dst + static_cast<size_t>(r) * H + c) = o;
        }
      }
    }
  }
}

// ---- one GEMM stage: prologue (activations -> smem) + K-split GEMM + epilogue (+ LayerNorm by the last CTA) ------------
// out[n][f] = epi(sum_k X[n][k] W[f][k] + bias[f]) for f in [0, N), n in [0, B). Unit = 16 features x K.
// One copy of this code serves every K-split stage (__noinline__, run-time K): the step runs each stage once, so code
// that is instantiated per shape or per call site is fetched cold every time — a 24 k-instruction kernel spent more time
// on instruction fetch than on its data.
// The 16 warps work as one group (K a multiple of 512: every warp takes K / 16 of ONE unit) or as two groups of 8 warps on
// two units at once; a group synchronises on its own named barrier. The first chunk of a CTA's weight loads is issued
// BEFORE the prologue's data is awaited: weights do not depend on activations.
__device__ __noinline__ void stage_gemm(const DsParams& p, int K, const StageIn& in, unsigned char* Xs, float* red,
                                        unsigned long long* fill_bar, unsigned int& fill_phase, const bf16* W, const bf16* bias, int N,
                                        int epi, void* out, long long ldo, const bf16* resid, const PostLN& post, int mk = -1) {
  const int units = (N + 15) >> 4;
  if (static_cast<int>(blockIdx.x) >= units) return;  // this CTA owns no unit of the stage
  const int WG = (K & 511) == 0 ? 16 : 8;   // warps per group
  const int NG = DS_WARPS / WG;             // groups
  const int kbw = K / (32 * WG);            // 32-wide k-blocks per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = warp / WG, gwarp = warp - group * WG;
  const int tig = threadIdx.x - group * WG * 32;  // thread index inside the group
  const int g = lane >> 2, t = lane & 3;
  const int stride = K * 2 + 64;
  float* gred = red + group * (WG * DS_RED);
  const int npt = 512 / (WG * 32);  // results per thread in the reduction: 1 (16 warps) or 2 (8 warps)

  // ---- prologue, asynchronous part: one bulk copy per row ----
  const int rows = p.ng * 8;
  if (warp == 0) {
    fence_proxy_async_smem();  // earlier generic-proxy accesses of this smem are ordered before the copy engine's writes
    const unsigned int row_bytes = static_cast<unsigned int>(K) * 2u;
    if (lane == 0) mbar_arrive_expect_tx(reinterpret_cast<uint64_t*>(fill_bar), row_bytes * static_cast<unsigned int>(p.B));
    __syncwarp();
    if (lane < p.B) {
      const bf16* srow = in.kind == DS_IN_COPY ? static_cast<const bf16*>(in.src) + static_cast<size_t>(lane) * K
                                               : p.emb + static_cast<size_t>(p.tok[lane]) * p.H;  // embedding gather
      bulk_g2s(Xs + static_cast<size_t>(lane) * stride, srow, row_bytes, fill_bar);
    }
  } else if (warp == 1) {
    for (int r = p.B + (lane >> 3); r < rows; r += 4)  // padding rows of the last token group: zeros
      for (int c = (lane & 7) * 16; c < K * 2; c += 128) *reinterpret_cast<uint4*>(Xs + static_cast<size_t>(r) * stride + c) = make_uint4(0, 0, 0, 0);
  }

  constexpr int CH = 4;  // k-blocks whose weight loads are in flight together (2 x 16-byte loads per thread and block)
  uint4 ra[CH], rb[CH];
  int u = blockIdx.x + group * gridDim.x;
  auto load_chunk = [&](int unit, int c0) {
    const int f0 = unit * 16;
    const bool ok0 = f0 + g < N, ok1 = f0 + g + 8 < N;
    const bf16* w0 = W + static_cast<size_t>(f0 + g) * K + gwarp * (32 * kbw) + t * 8;
    const bf16* w1 = w0 + static_cast<size_t>(8) * K;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const bool inb = c0 + j < kbw;
      ra[j] = (ok0 && inb) ? ldg_stream(w0 + (c0 + j) * 32) : make_uint4(0, 0, 0, 0);
      rb[j] = (ok1 && inb) ? ldg_stream(w1 + (c0 + j) * 32) : make_uint4(0, 0, 0, 0);
    }
  };
  bool preloaded = false;
  if (u < units) {
    load_chunk(u, 0);
    preloaded = true;
  }

  // ---- prologue, synchronous part ----
  mbar_wait_spin(fill_bar, fill_phase);
  fill_phase ^= 1u;
  if (in.kind == DS_IN_EMBED && in.xout != nullptr && static_cast<int>(blockIdx.x) < p.B) {
    const int r = blockIdx.x;  // CTA r publishes row r for the residual adds of later stages
    for (int c = threadIdx.x * 8; c < p.H; c += DS_THREADS * 8)
      *reinterpret_cast<uint4*>(in.xout + static_cast<size_t>(r) * p.H + c) = *reinterpret_cast<const uint4*>(Xs + static_cast<size_t>(r) * stride + c * 2);
  }
  __syncthreads();  // (the zero rows written by warp 1)
  ds_mark(p, mk);   // prologue done

  for (; u < units; u += NG * gridDim.x) {
    const int f0 = u * 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // this thread's results of the unit (features er [, er + 1] of token en): their bias / residual loads go out now, so
    // their latency hides behind the weight stream instead of following the reduction
    const int en = (tig * npt) >> 4, er = (tig * npt) & 15;
    const int ef = f0 + er;
    const bool e_ok = en < p.B && ef < N, e_two = npt == 2 && ef + 1 < N;
    float eb0 = 0.f, eb1 = 0.f, ex0 = 0.f, ex1 = 0.f;
    if (e_ok) {
      if (bias) {
        eb0 = __bfloat162float(bias[ef]);
        if (e_two) eb1 = __bfloat162float(bias[ef + 1]);
      }
      if (epi == DS_EPI_F32_RESID) {  // (written by other CTAs in an earlier stage: L2 read)
        const unsigned short r0 = __ldcg(reinterpret_cast<const unsigned short*>(resid + static_cast<size_t>(en) * ldo + ef));
        ex0 = __bfloat162float(*reinterpret_cast<const bf16*>(&r0));
        if (e_two) {
          const unsigned short r1 = __ldcg(reinterpret_cast<const unsigned short*>(resid + static_cast<size_t>(en) * ldo + ef + 1));
          ex1 = __bfloat162float(*reinterpret_cast<const bf16*>(&r1));
        }
      }
    }
#pragma unroll 1
    for (int c0 = 0; c0 < kbw; c0 += CH) {
      if (!preloaded) load_chunk(u, c0);
      preloaded = false;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        if (c0 + j >= kbw) break;
        const int kbyte = (gwarp * (32 * kbw) + (c0 + j) * 32 + t * 8) * 2;
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
    ds_mark(p, mk < 0 ? -1 : mk + 1);  // weights consumed (thread 0's share)
    // partial tile of this warp -> gred[warp][feature r][token n ^ r]
    float* my = gred + gwarp * DS_RED;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = i * 8 + t * 2;
      my[g * 32 + (n ^ g)] = acc[i][0];
      my[g * 32 + ((n + 1) ^ g)] = acc[i][1];
      my[(g + 8) * 32 + (n ^ (g + 8))] = acc[i][2];
      my[(g + 8) * 32 + ((n + 1) ^ (g + 8))] = acc[i][3];
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(WG * 32) : "memory");
    ds_mark(p, mk < 0 ? -1 : mk + 2);  // every warp of the group has its partial in smem
    // 512 results over the group's threads (fixed summation order over the warps: deterministic)
    {
      const int n = en, r = er, f = ef;
      float v0 = 0.f, v1 = 0.f;
      for (int w = 0; w < WG; ++w) {
        v0 += gred[w * DS_RED + r * 32 + (n ^ r)];
        if (npt == 2) v1 += gred[w * DS_RED + (r + 1) * 32 + (n ^ (r + 1))];
      }
      if (e_ok) {
        const bool two = e_two;
        v0 += eb0;
        v1 += eb1;
        if (epi == DS_EPI_GELU_BF16 || epi == DS_EPI_GELU_F32) {
          v0 = gelu_erf(v0);
          v1 = gelu_erf(v1);
        }
        v0 += ex0;
        v1 += ex1;
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
    asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(WG * 32) : "memory");
  }

  // ---- LayerNorm of the finished sums by the last CTA of the stage ----
  if (post.src != nullptr) {
    unsigned int& s_last_cta = *reinterpret_cast<unsigned int*>(fill_bar + 32);  // (the dynamic region's tail: no static smem, the plan uses all 227 KB)
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int ctas = static_cast<unsigned int>(units < static_cast<int>(gridDim.x) ? units : static_cast<int>(gridDim.x));
      unsigned int prev;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(post.ticket) : "memory");
      s_last_cta = prev == ctas - 1u ? 1u : 0u;
      if (s_last_cta) *post.ticket = 0u;  // self-reset for the next stage / launch
    }
    __syncthreads();
    if (s_last_cta) {
      ds_mark_any(p, mk < 0 ? -1 : mk + 4);
      layernorm_rows(p, post);
      ds_mark_any(p, mk < 0 ? -1 : mk + 5);
    }
  }
}

// ---- vocabulary projection + greedy argmax ---------------------------------------------------------------
// The 77 MB of lm_head.decoder.weight are the largest single stream of the step. A unit (16 vocabulary rows x H) is ONE
// contiguous 24 KB block of the weight: it goes through a shared-memory ring filled by the copy engine with one bulk copy
// per unit, three slots per 8-warp group, so up to ~150 KB per SM are in flight without holding a register. Each group
// multiplies its unit (K split over its 8 warps; the unpadded weight rows cost a 2-way bank conflict on 6 loads per warp,
// irrelevant), reduces through smem, adds the bias, rounds to the model dtype and folds the 16 logits of every token into
// a running (value, index) maximum; one 64-bit atomicMax per token and CTA at the end (first index wins ties: the
// torch.topk(k=1) rule of models/decoder.py:489-496).
constexpr int DS_RING = 3;
__device__ __noinline__ void lm_head_argmax(const DsParams& p, const bf16* xin, unsigned char* Xs, float* red, unsigned long long* bars,
                                            unsigned int& fill_phase) {
  const int K = p.H;
  const int kbw = K / 256;  // k-blocks per warp (8 warps per group)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = warp >> 3, gwarp = warp & 7;
  const int tig = threadIdx.x & 255;
  const int g = lane >> 2, t = lane & 3;
  const int stride = K * 2 + 64;
  const int wstride = K * 2;  // weight rows in the ring: as in global memory
  const int N = p.V;
  const int units = (N + 15) >> 4;
  unsigned char* ring0 = Xs + ((static_cast<size_t>(32) * stride + 1023) & ~static_cast<size_t>(1023));
  unsigned char* ring = ring0 + static_cast<size_t>(group) * DS_RING * 16 * wstride;
  unsigned long long* fill_bar = bars;
  unsigned long long* full = bars + 1 + group * DS_RING;
  float* gred = red + group * (8 * DS_RED);
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(red + 2 * 8 * DS_RED);  // [2 groups][32]

  // this group's units: u = 2 * (blockIdx.x + i * gridDim.x) + group
  const int ustep = 2 * gridDim.x;
  const int u0 = 2 * blockIdx.x + group;
  auto issue = [&](int u, int slot) {  // one bulk copy per unit, by the first lane of the group
    if (gwarp != 0 || lane != 0 || u >= units) return;
    const unsigned int bytes = static_cast<unsigned int>(min(16, N - u * 16)) * static_cast<unsigned int>(wstride);
    fence_proxy_async_smem();
    mbar_arrive_expect_tx(reinterpret_cast<uint64_t*>(&full[slot]), bytes);
    bulk_g2s(ring + static_cast<size_t>(slot) * 16 * wstride, p.w_v + static_cast<size_t>(u) * 16 * K, bytes, &full[slot]);
  };
#pragma unroll
  for (int i = 0; i < DS_RING; ++i) issue(u0 + i * ustep, i);  // weights first: they do not wait for the activations

  // prologue: the normalised rows (bf16, written by the last CTA of the previous stage) -> Xs
  if (warp == 0) {
    fence_proxy_async_smem();
    if (lane == 0) mbar_arrive_expect_tx(reinterpret_cast<uint64_t*>(fill_bar), static_cast<unsigned int>(K) * 2u * static_cast<unsigned int>(p.B));
    __syncwarp();
    if (lane < p.B) bulk_g2s(Xs + static_cast<size_t>(lane) * stride, xin + static_cast<size_t>(lane) * K, static_cast<unsigned int>(K) * 2u, fill_bar);
  } else if (warp == 1) {
    for (int r = p.B + (lane >> 3); r < p.ng * 8; r += 4)
      for (int c = (lane & 7) * 16; c < K * 2; c += 128) *reinterpret_cast<uint4*>(Xs + static_cast<size_t>(r) * stride + c) = make_uint4(0, 0, 0, 0);
  }
  mbar_wait_spin(fill_bar, fill_phase);
  fill_phase ^= 1u;
  __syncthreads();

  unsigned long long best = 0ull;  // owner threads ((tig & 7) == 0): running maximum of token tig >> 3
  unsigned int phases = 0u;        // bit s: parity of ring slot s
  int slot = 0;
  for (int u = u0; u < units; u += ustep) {
    mbar_wait_spin(&full[slot], (phases >> slot) & 1u);
    phases ^= 1u << slot;
    const unsigned char* wrow = ring + static_cast<size_t>(slot) * 16 * wstride;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int kb = 0; kb < kbw; ++kb) {
      const int kbyte = ((gwarp * kbw + kb) * 32 + t * 8) * 2;
      const uint4 ra = *reinterpret_cast<const uint4*>(wrow + static_cast<size_t>(g) * wstride + kbyte);
      const uint4 rb = *reinterpret_cast<const uint4*>(wrow + static_cast<size_t>(g + 8) * wstride + kbyte);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < p.ng) {
          const uint4 xb = *reinterpret_cast<const uint4*>(Xs + static_cast<size_t>(i * 8 + g) * stride + kbyte);
          mma_bf16(acc[i], ra.x, rb.x, ra.y, rb.y, xb.x, xb.y);
          mma_bf16(acc[i], ra.z, rb.z, ra.w, rb.w, xb.z, xb.w);
        }
      }
    }
    float* my = gred + gwarp * DS_RED;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = i * 8 + t * 2;
      my[g * 32 + (n ^ g)] = acc[i][0];
      my[g * 32 + ((n + 1) ^ g)] = acc[i][1];
      my[(g + 8) * 32 + (n ^ (g + 8))] = acc[i][2];
      my[(g + 8) * 32 + ((n + 1) ^ (g + 8))] = acc[i][3];
    }
    asm volatile("bar.sync %0, 256;" ::"r"(1 + group) : "memory");  // slot consumed by all 8 warps, partials in smem
    issue(u + DS_RING * ustep, slot);                                 // refill it while the epilogue runs
    {
      const int n = tig >> 3, r = (tig & 7) * 2;  // token n, features r, r + 1
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        v0 += gred[w * DS_RED + r * 32 + (n ^ r)];
        v1 += gred[w * DS_RED + (r + 1) * 32 + (n ^ (r + 1))];
      }
      const int fa = u * 16 + r, fb = fa + 1;
      unsigned long long key = 0ull;
      if (n < p.B) {
        // bias, then the rounding of the model dtype (what the logits tensor of a bf16 model holds)
        if (fa < N) {
          const bf16 la = __float2bfloat16_rn(v0 + (p.b_v ? __bfloat162float(p.b_v[fa]) : 0.f));
          key = (static_cast<unsigned long long>(ds_order_bits(__bfloat162float(la))) << 32) | (0xffffffffu - static_cast<unsigned int>(fa));
          if (p.logits) reinterpret_cast<bf16*>(p.logits)[static_cast<size_t>(n) * p.ld_logits + fa] = la;
        }
        if (fb < N) {
          const bf16 lb = __float2bfloat16_rn(v1 + (p.b_v ? __bfloat162float(p.b_v[fb]) : 0.f));
          const unsigned long long kb2 = (static_cast<unsigned long long>(ds_order_bits(__bfloat162float(lb))) << 32) | (0xffffffffu - static_cast<unsigned int>(fb));
          key = kb2 > key ? kb2 : key;
          if (p.logits) reinterpret_cast<bf16*>(p.logits)[static_cast<size_t>(n) * p.ld_logits + fb] = lb;
        }
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {  // the 8 threads of a token
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
      }
      best = key > best ? key : best;
    }
    asm volatile("bar.sync %0, 256;" ::"r"(1 + group) : "memory");  // gred free for the next unit
    slot = slot + 1 == DS_RING ? 0 : slot + 1;
  }
  if ((tig & 7) == 0) s_keys[group * 32 + (tig >> 3)] = best;
  __syncthreads();
  if (threadIdx.x < p.B) {  // one atomic per token and CTA
    const unsigned long long a0 = s_keys[threadIdx.x], a1 = s_keys[32 + threadIdx.x];
    const unsigned long long k = a0 > a1 ? a0 : a1;
    if (k != 0ull) atomicMax(&p.amax[threadIdx.x], k);
  }
}

// ---- attention over the cache: one work item = (kv-split, kv head, batch row) per WARP ----
// A decode step has B * h_kv rows of work and 8 warps on each of 148 SMs: the context of every (row, kv head) is split so
// that there is about one item per warp of the grid, every warp runs its item without any CTA-wide synchronisation (the
// latency chain of an item — projections from L2, RoPE, cache rows, merge — is paid once per stage, not once per item
// and CTA), and the last warp to finish a (row, kv head) combines the splits (ticket in global memory).
__device__ __forceinline__ void attn_item_coords(const DsParams& p, int it, int sp, int& split, int& kvh, int& b, int& k_begin, int& k_end) {
  split = it % p.splits;
  kvh = (it / p.splits) % p.Hkv;
  b = it / (p.splits * p.Hkv);
  const int per = (sp + p.splits - 1) / p.splits;
  k_begin = split * per;
  k_end = min(sp, k_begin + per);
}

template <int NREP>
__device__ void attention_items(const DsParams& p, const DsLayer& ly, int sp, unsigned char* smem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = lane & 7, lk = lane >> 3;
  float* wq = reinterpret_cast<float*>(smem) + warp * (10 * DS_HD);  // [NREP <= 8][64] q, then new k [64], new v [64]
  float* s_newk = wq + 8 * DS_HD;
  float* s_newv = s_newk + DS_HD;
  const float scale_log2 = 1.4426950408889634f / 8.0f;
  const int items = p.B * p.Hkv * p.splits;
  const int nw = gridDim.x * DS_WARPS;
  for (int it = blockIdx.x * DS_WARPS + warp; it < items; it += nw) {
    int split, kvh, b, k_begin, k_end;
    attn_item_coords(p, it, sp, split, kvh, b, k_begin, k_end);
    // new token: projections (bias already added) -> RoPE(q, k) -> this warp's smem; lane j owns dims j and j + 32
    const float* row = p.qkv + static_cast<size_t>(b) * p.NQKV;
    {
      float lo[NREP + 2], hi[NREP + 2];
      float c = 1.f, sn = 0.f;
      if (p.rope_cos) {
        c = p.rope_cos[sp * 32 + lane];
        sn = p.rope_sin[sp * 32 + lane];
      }
#pragma unroll
      for (int h = 0; h < NREP + 2; ++h) {
        const int col = h < NREP ? (kvh * NREP + h) * DS_HD : (h == NREP ? (p.Hq + kvh) * DS_HD : (p.Hq + p.Hkv + kvh) * DS_HD);
        lo[h] = __ldcg(row + col + lane);
        hi[h] = __ldcg(row + col + lane + 32);
      }
#pragma unroll
      for (int h = 0; h < NREP + 2; ++h) {
        // the reference's bf16 model holds q / k / v as bf16 tensors before the rotation: same rounding point
        const float x0 = __bfloat162float(__float2bfloat16_rn(lo[h])), x1 = __bfloat162float(__float2bfloat16_rn(hi[h]));
        float* dst = h < NREP ? wq + h * DS_HD : (h == NREP ? s_newk : s_newv);
        if (h <= NREP && p.rope_cos) {
          dst[lane] = x0 * c - x1 * sn;
          dst[lane + 32] = x1 * c + x0 * sn;
        } else {
          dst[lane] = x0;
          dst[lane + 32] = x1;
        }
      }
    }
    __syncwarp();
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[128 + 16] = static_cast<long long>(globaltimer_ns());
    bf16* kc = ly.k_cache + b * p.c_sb + kvh * p.c_sh;
    bf16* vc = ly.v_cache + b * p.c_sb + kvh * p.c_sh;
    if (split == p.splits - 1) {  // the cache stores what the reference stores: rotated k, raw v, in the cache dtype
      kc[sp * p.c_sl + lane] = __float2bfloat16_rn(s_newk[lane]);
      kc[sp * p.c_sl + lane + 32] = __float2bfloat16_rn(s_newk[lane + 32]);
      vc[sp * p.c_sl + lane] = __float2bfloat16_rn(s_newv[lane]);
      vc[sp * p.c_sl + lane + 32] = __float2bfloat16_rn(s_newv[lane + 32]);
    }
    float q[NREP][8];
#pragma unroll
    for (int r = 0; r < NREP; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) q[r][j] = wq[r * DS_HD + ld * 8 + j] * scale_log2;
    float m[NREP], l[NREP], o[NREP][8];
#pragma unroll
    for (int r = 0; r < NREP; ++r) {
      m[r] = -INFINITY;
      l[r] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[r][j] = 0.f;
    }
    constexpr int UN = 4;  // 4 keys per warp and load step; UN steps in flight
    for (int k0 = k_begin; k0 < k_end; k0 += 4 * UN) {
      uint4 kraw[UN], vraw[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int ki = k0 + u * 4 + lk;
        kraw[u] = vraw[u] = make_uint4(0, 0, 0, 0);
        if (ki < k_end) {
          const long long off = ki * p.c_sl + ld * 8;
          kraw[u] = ldg_stream(kc + off);
          vraw[u] = ldg_stream(vc + off);
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const bool valid = k0 + u * 4 + lk < k_end;
        float kv[8], vv[8];
        const __nv_bfloat162* hk = reinterpret_cast<const __nv_bfloat162*>(&kraw[u]);
        const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&vraw[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a = __bfloat1622float2(hk[j]), cc = __bfloat1622float2(hv[j]);
          kv[2 * j] = a.x; kv[2 * j + 1] = a.y;
          vv[2 * j] = cc.x; vv[2 * j + 1] = cc.y;
        }
#pragma unroll
        for (int r = 0; r < NREP; ++r) {
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) s += q[r][j] * kv[j];
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
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[128 + 17] = static_cast<long long>(globaltimer_ns());
    if (split == p.splits - 1 && lk == 0) {  // the new token, from smem (unrounded, like the per-op kernel)
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
    // merge the 4 key lane groups (xor 8, 16): afterwards every lane holds the item's (m, l, o[ld * 8 ..])
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
    }
    bf16* orow = p.attn + static_cast<size_t>(b) * (p.Hq * DS_HD) + kvh * NREP * DS_HD;
    if (p.splits == 1) {
      if (lk == 0) {
#pragma unroll
        for (int r = 0; r < NREP; ++r) {
          uint4 w;
          __nv_bfloat162* w2 = reinterpret_cast<__nv_bfloat162*>(&w);
          const float inv = 1.f / l[r];
#pragma unroll
          for (int j = 0; j < 4; ++j) w2[j] = __floats2bfloat162_rn(o[r][2 * j] * inv, o[r][2 * j + 1] * inv);
          *reinterpret_cast<uint4*>(orow + r * DS_HD + ld * 8) = w;
        }
      }
      continue;
    }
    float* part = p.part + (((static_cast<size_t>(b) * p.Hkv + kvh) * p.splits + split) * NREP) * (DS_HD + 2);
    if (lk == 0) {
#pragma unroll
      for (int r = 0; r < NREP; ++r) {
        float* w = part + r * (DS_HD + 2);
#pragma unroll
        for (int j = 0; j < 8; j += 2) *reinterpret_cast<float2*>(w + ld * 8 + j) = make_float2(o[r][j], o[r][j + 1]);
        if (ld == 0) *reinterpret_cast<float2*>(w + DS_HD) = make_float2(m[r], l[r]);
      }
    }
    // the last warp of this (row, kv head) combines the splits
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[128 + 18] = static_cast<long long>(globaltimer_ns());
    __threadfence();
    __syncwarp();
    unsigned int last = 0;
    if (lane == 0) {
      const unsigned int prev = atomicAdd(&p.tickets[b * p.Hkv + kvh], 1u);
      last = prev == static_cast<unsigned int>(p.splits - 1) ? 1u : 0u;
      if (last) p.tickets[b * p.Hkv + kvh] = 0u;  // self-reset for the next stage / launch
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) continue;
    __threadfence();
    const float* base = p.part + ((static_cast<size_t>(b) * p.Hkv + kvh) * p.splits) * NREP * (DS_HD + 2);
    // lane s < splits fetches split s's (m, l) of every head at once; the weights exp2(m_s - M) / L then ride on shuffles
    // while every lane accumulates its two output dims over the splits, four splits' loads in flight at a time
#pragma unroll
    for (int r = 0; r < NREP; ++r) {
      float2 ml = make_float2(-INFINITY, 0.f);
      if (lane < p.splits) ml = __ldcg(reinterpret_cast<const float2*>(base + (static_cast<size_t>(lane) * NREP + r) * (DS_HD + 2) + DS_HD));
      float M = ml.x;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
      const float wgt = ml.x == -INFINITY ? 0.f : exp2f(ml.x - M);
      float Ls = ml.y * wgt;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) Ls += __shfl_xor_sync(0xffffffffu, Ls, o);
      float O0 = 0.f, O1 = 0.f;
      for (int s0 = 0; s0 < p.splits; s0 += 4) {
        float2 ov[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ov[j] = make_float2(0.f, 0.f);
          if (s0 + j < p.splits) ov[j] = __ldcg(reinterpret_cast<const float2*>(base + (static_cast<size_t>(s0 + j) * NREP + r) * (DS_HD + 2) + 2 * lane));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float ws = __shfl_sync(0xffffffffu, wgt, (s0 + j) & 31);
          if (s0 + j < p.splits) {
            O0 += ov[j].x * ws;
            O1 += ov[j].y * ws;
          }
        }
      }
      const float inv = 1.f / Ls;
      *reinterpret_cast<__nv_bfloat162*>(orow + r * DS_HD + 2 * lane) = __floats2bfloat162_rn(O0 * inv, O1 * inv);
    }
  }
}

// One layer's HBM bytes (projection / MLP weights and the cached keys / values up to `pos`) -> L2, shared out over the grid
__device__ __noinline__ void prefetch_layer(const DsParams& p, const DsLayer& ly, int pos) {
  const int part = blockIdx.x, nparts = gridDim.x;
  prefetch_range_l2(ly.w_qkv, static_cast<long long>(p.NQKV) * p.H * 2, part, nparts);
  prefetch_range_l2(ly.w_o, static_cast<long long>(p.H) * p.H * 2, part, nparts);
  prefetch_range_l2(ly.w_1, static_cast<long long>(p.FF) * p.H * 2, part, nparts);
  prefetch_range_l2(ly.w_2, static_cast<long long>(p.FF) * p.H * 2, part, nparts);
  // cache: (row, kv head) r holds slots [0, pos) contiguously when the slot stride is the head dim
  if (p.c_sl == DS_HD && pos > 0) {
    const int ranges = p.B * p.Hkv;
    for (int r = blockIdx.x; r < ranges; r += gridDim.x) {
      const long long off = static_cast<long long>(r / p.Hkv) * p.c_sb + static_cast<long long>(r % p.Hkv) * p.c_sh;
      prefetch_range_l2(ly.k_cache + off, static_cast<long long>(pos) * DS_HD * 2, 0, 1);
      prefetch_range_l2(ly.v_cache + off, static_cast<long long>(pos) * DS_HD * 2, 0, 1);
    }
  }
}
// part `which` of `of` of the vocabulary projection's weight
__device__ __forceinline__ void prefetch_vocab(const DsParams& p, int which, int of) {
  const long long bytes = static_cast<long long>(p.V) * p.H * 2;
  const long long piece = ((bytes / of) + 127) & ~127ll;
  const long long lo = piece * which;
  if (lo >= bytes) return;
  prefetch_range_l2(reinterpret_cast<const char*>(p.w_v) + lo, (lo + piece < bytes ? piece : bytes - lo), blockIdx.x, gridDim.x);
}

template <int NREP>
__global__ void __launch_bounds__(DS_THREADS, 1)
decode_step_kernel(const __grid_constant__ DsParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int Kmax = p.FF > p.H ? p.FF : p.H;
  unsigned char* Xs = smem;
  float* red = reinterpret_cast<float*>(smem + static_cast<size_t>(32) * xs_stride(Kmax));
  // behind the reduction tiles: 512 B of argmax keys, then the mbarriers — [0]: stage prologue fill; [1..6]: vocabulary ring (2 groups x 3 slots)
  unsigned long long* s_bars = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(red) + DS_WARPS * DS_RED * sizeof(float) + 512);
  const int pos = *p.pos;  // read before anything can change it (FIN runs after the last barrier)
  int bar = 0;
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0 && p.trace) p.trace[0] = static_cast<long long>(globaltimer_ns());
    if (blockIdx.x == 0) p.bar[5 * p.L + 1] = 0u;  // the previous launch's last barrier (5 L + 2 per launch)
#pragma unroll
    for (int i = 0; i < 1 + 2 * DS_RING; ++i) mbar_init(reinterpret_cast<uint64_t*>(&s_bars[i]), 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (blockIdx.x == gridDim.x - 1) {  // biases and LayerNorm vectors: a few KB each, first touched deep inside latency chains
    for (int l = 0; l < p.L; ++l) {
      const DsLayer& ly = p.layer[l];
      const bf16* v[8] = {ly.b_qkv, ly.b_o, ly.ln1_g, ly.ln1_b, ly.b_1, ly.b_2, ly.ln2_g, ly.ln2_b};
      const int n[8] = {p.NQKV, p.H, p.H, p.H, p.FF, p.H, p.H, p.H};
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (v[i]) prefetch_range_l2(v[i], static_cast<long long>(n[i]) * 2, 0, 1);
    }
    if (p.b_d) prefetch_range_l2(p.b_d, static_cast<long long>(p.H) * 2, 0, 1);
    prefetch_range_l2(p.lnh_g, static_cast<long long>(p.H) * 2, 0, 1);
    prefetch_range_l2(p.lnh_b, static_cast<long long>(p.H) * 2, 0, 1);
    if (p.b_v) prefetch_range_l2(p.b_v, static_cast<long long>(p.V) * 2, 0, 1);
  }
  unsigned int fill_phase = 0;
  const int qcols = p.Hq * DS_HD;
  StageIn in;
  const PostLN no_ln = {nullptr, nullptr, nullptr, 0.f, nullptr, nullptr};
  PostLN ln;

  for (int l = 0; l < p.L; ++l) {
    const DsLayer& ly = p.layer[l];
    // HBM -> L2 a layer ahead: the next layer's weights and cache rows; the vocabulary weight in two halves over the last two
    // layers (L2 holds 126 MB: the layer in use, the one being fetched and the head's 77 MB must not all be live at once)
    if (l + 1 < p.L) prefetch_layer(p, p.layer[l + 1], pos);
    if (l == p.L - 2 || p.L == 1) prefetch_vocab(p, 0, 2);
    if (l == p.L - 1) {
      prefetch_vocab(p, 1, 2);
      if (p.L == 1) prefetch_vocab(p, 0, 2);
      prefetch_range_l2(p.w_d, static_cast<long long>(p.H) * p.H * 2, blockIdx.x, gridDim.x);
    }
    // QKV: x = the tokens' embedding rows (layer 0) or LN2 of the previous layer's sum (written to xbuf by the last CTA of
    // its FFN2 stage). While it runs, L2 pulls in the cache rows of the attention stage and the output projection's weights.
    // The copy engine of an SM serves its requests in order: an L2 prefetch queued BEFORE a stage's activation copies makes
    // them wait for the prefetch's DRAM fetch — so prefetches are issued after the stage's own work, ahead of the barrier.
    in = l == 0 ? StageIn{DS_IN_EMBED, nullptr, p.xbuf} : StageIn{DS_IN_COPY, p.xbuf, nullptr};
    stage_gemm(p, p.H, in, Xs, red, &s_bars[0], fill_phase, ly.w_qkv, ly.b_qkv, p.NQKV, DS_EPI_F32, p.qkv, p.NQKV, nullptr, no_ln, l == 1 ? 0 : -1);
    ds_mark(p, l == 1 ? 3 : -1);
    grid_barrier(p, bar++);
    // ATTN
    attention_items<NREP>(p, ly, pos, smem);
    grid_barrier(p, bar++);
    // OUT: s1 = attn Wo^T + bo + x; its last CTA writes y = LN1(s1)
    in = StageIn{DS_IN_COPY, p.attn, nullptr};
    ln = PostLN{p.s1, ly.ln1_g, ly.ln1_b, p.eps_layer, p.ybuf, p.ln_ticket};
    stage_gemm(p, qcols, in, Xs, red, &s_bars[0], fill_phase, ly.w_o, ly.b_o, p.H, DS_EPI_F32_RESID, p.s1, p.H, p.xbuf, ln);
    grid_barrier(p, bar++);
    // FFN1: a = gelu(y W1^T + b1)
    in = StageIn{DS_IN_COPY, p.ybuf, nullptr};
    stage_gemm(p, p.H, in, Xs, red, &s_bars[0], fill_phase, ly.w_1, ly.b_1, p.FF, DS_EPI_GELU_BF16, p.abuf, p.FF, nullptr, no_ln);
    grid_barrier(p, bar++);
    // FFN2: s2 = a W2^T + b2 + x (the residual is the layer INPUT: quirk Q2); its last CTA writes the next x = LN2(s2)
    in = StageIn{DS_IN_COPY, p.abuf, nullptr};
    ln = PostLN{p.s2, ly.ln2_g, ly.ln2_b, p.eps_layer, p.xbuf, p.ln_ticket};
    stage_gemm(p, p.FF, in, Xs, red, &s_bars[0], fill_phase, ly.w_2, ly.b_2, p.H, DS_EPI_F32_RESID, p.s2, p.H, p.xbuf, ln, l == 1 ? 8 : -1);
    ds_mark(p, l == 1 ? 11 : -1);
    grid_barrier(p, bar++);
  }
  // LM head: g = gelu(x Wd^T + bd); the last CTA writes y = LN(g)
  in = StageIn{DS_IN_COPY, p.xbuf, nullptr};
  ln = PostLN{p.gbuf, p.lnh_g, p.lnh_b, p.eps_head, p.ybuf, p.ln_ticket};
  stage_gemm(p, p.H, in, Xs, red, &s_bars[0], fill_phase, p.w_d, p.b_d, p.H, DS_EPI_GELU_F32, p.gbuf, p.H, nullptr, ln);
  grid_barrier(p, bar++);
  // logits = y Wv^T + bv -> argmax; meanwhile L2 takes in layer 0's bytes for the NEXT step (its cache now reaches slot pos)
  prefetch_layer(p, p.layer[0], pos + 1);
  lm_head_argmax(p, p.ybuf, Xs, red, s_bars, fill_phase);
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
  size_t xbuf, ybuf, ln_ticket, qkv, attn, s1, abuf, s2, gbuf, part, tickets, amax, bar, abort_flag, total;
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
  s.ln_ticket = take(sizeof(unsigned int));
  s.xbuf = take(sizeof(bf16) * B * H);
  s.ybuf = take(sizeof(bf16) * B * H);
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
    VY_CHECK_ARG(hb >= 1 && hb <= 4 && fb >= 1 && fb <= 16, "vy_decode_step: unsupported H / ffn (%d / %d)", H, FF);
  }
  // the vocabulary stage keeps its weight ring (2 groups x 3 slots x 16 rows x H bf16) behind the bf16 activation rows
  VY_CHECK_ARG(static_cast<size_t>(32) * (FF * 2 + 64) >= ((static_cast<size_t>(32) * (H * 2 + 64) + 1023) / 1024) * 1024 + static_cast<size_t>(2 * DS_RING * 16) * H * 2,
               "vy_decode_step: ffn (%d) too small for the shared-memory plan of H = %d (needs ffn >= 4 H)", FF, H);
  VY_CHECK_ARG(q->pos_table == nullptr, "vy_decode_step: learned / sinusoidal position tables are not built into the one-launch step (RoPE models only)");
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
  int splits = (sms * DS_WARPS) / (B * Hkv);  // about one attention item per warp of the grid
  const int by_len = (q->pos_bound + 31) / 32;
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
  p.ybuf = reinterpret_cast<bf16*>(ws + sc.ybuf);
  p.ln_ticket = reinterpret_cast<unsigned int*>(ws + sc.ln_ticket);
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
  const size_t smem = static_cast<size_t>(32) * (Kmax * 2 + 64) + DS_WARPS * DS_RED * sizeof(float) + 1024;  // rows | reduction tiles | argmax keys + mbarriers
  VY_CHECK_ARG(smem <= 232448, "vy_decode_step: shared-memory plan (%zu bytes) exceeds the 227 KB of an SM", smem);
  const size_t attn_smem = static_cast<size_t>(DS_WARPS) * 10 * DS_HD * sizeof(float);
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
    std::call_once(once, [&] { attr_rc = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448); }); \
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
