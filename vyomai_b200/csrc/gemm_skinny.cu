// Small-batch ("skinny") linear layer for the decode step: out[n][f] = epi(sum_k X[n][k] W[f][k] + bias[f]) with at most 32
// rows of X — the shape of every projection of a single-token decode step (models/decoder.py:355-373 at seqlen 1:
// attention.py:87-95,69; ffn.py:33-37; decoder.py:267-275). HBM-bound: the only real traffic is W, read exactly once.
//
// vy_gemm routes here (from its swap-AB form: A = W [F][K], B = X [n][K], transposed_out) instead of the tcgen05 kernel,
// which at these shapes puts one 128-row weight tile per CTA on 6-24 SMs and walks K serially (ncu: 8.5-17 us for
// 1.2-4.7 MB). Here the weight matrix is cut into (16 features) x (K / ksplit) pieces, one 128-thread CTA each, so that
// 150-400 CTAs have ALL of W requested from HBM at once — every thread issues its 16-byte weight loads (<= 12 of them)
// before anything else, and under programmatic dependent launch even before the previous kernel has finished
// (weights_static: they do not depend on it). The math is mma.sync m16n8k16 (bf16, fp32 accumulate) with W as the
// 16-row operand and the tokens as the 8-column one; a warp's 8 consecutive k per thread serve two MMAs through a
// k-permutation that is applied to both operands alike, so all loads are 16 bytes wide. Partial sums meet in shared
// memory (4 warps), then — ksplit > 1 — across the thread-block CLUSTER formed by the ksplit CTAs of a feature block,
// through distributed shared memory in fixed rank order: no global scratch, no atomics, bit-identical run to run.
// Algorithmic bytes per call = F * K * 2 (+ the 32 x K activations, L2-resident).
#include "vy_common.cuh"
#include "vy_ptx.cuh"

namespace vy {

constexpr int SK_WARPS = 4;
constexpr int SK_THREADS = SK_WARPS * 32;
constexpr int SK_MAX_KB = 6;   // 32-wide k-blocks a warp has in flight at once (12 x 16-byte loads per thread)
constexpr int SK_TILE = 16 * 33;  // [16 features][32 tokens] fp32, padded rows

struct SkinnyDev {
  int F, K, n_tok, ksplit, kb_per_warp;  // features (output columns), reduction length, tokens (<= 32)
  int static_w;                          // weights may be read before griddepcontrol.wait
  const __nv_bfloat16* W;
  long long ldw;
  const __nv_bfloat16* X;
  long long ldx;
  const void* bias;
  int bias_dt;
  int act;
  const void* addend;
  long long ld_addend;
  int addend_dt;
  float out_scale;
  void* out;
  long long ld_out;
  int out_dt;
};

__device__ __forceinline__ void sk_mma(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 sk_ldg_stream(const void* p) {  // read-once weights: do not keep them in L1
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned sk_cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ float sk_ld_dsmem(unsigned local_addr, unsigned rank) {
  unsigned remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote));
  return v;
}
__device__ __forceinline__ void sk_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int NG>  // token groups of 8 (1..4)
__global__ void __launch_bounds__(SK_THREADS)
gemm_skinny_kernel(const SkinnyDev p) {
  __shared__ float s_part[SK_WARPS][SK_TILE];
  pdl_trigger();
  const int ks = blockIdx.x, unit = blockIdx.y;  // the cluster spans blockIdx.x
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int f0 = unit * 16;
  const int kbw = p.kb_per_warp;
  const int k0 = (ks * SK_WARPS + warp) * kbw * 32 + t * 8;  // this thread's first k
  const bool ok0 = f0 + g < p.F, ok1 = f0 + g + 8 < p.F;
  const __nv_bfloat16* w0 = p.W + static_cast<long long>(f0 + g) * p.ldw + k0;
  const __nv_bfloat16* w1 = w0 + 8 * p.ldw;
  // two register buffers of SK_MAX_KB k-blocks: the loads of chunk c + 1 are issued before the MMAs of chunk c
  uint4 ra[SK_MAX_KB], rb[SK_MAX_KB], ra2[SK_MAX_KB], rb2[SK_MAX_KB];
  auto load_w = [&](uint4 (&wa)[SK_MAX_KB], uint4 (&wb)[SK_MAX_KB], int c0) {  // k-blocks [c0, c0 + SK_MAX_KB) of this warp
#pragma unroll
    for (int j = 0; j < SK_MAX_KB; ++j) {
      wa[j] = (ok0 && c0 + j < kbw) ? sk_ldg_stream(w0 + (c0 + j) * 32) : make_uint4(0, 0, 0, 0);
      wb[j] = (ok1 && c0 + j < kbw) ? sk_ldg_stream(w1 + (c0 + j) * 32) : make_uint4(0, 0, 0, 0);
    }
  };
  if (p.static_w) load_w(ra, rb, 0);  // before the predecessor is awaited: (the first chunk of) the whole weight matrix is in flight
  pdl_wait();
  if (!p.static_w) load_w(ra, rb, 0);

  float acc[NG][4];
#pragma unroll
  for (int i = 0; i < NG; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  auto compute = [&](const uint4 (&wa)[SK_MAX_KB], const uint4 (&wb)[SK_MAX_KB], int c0) {
#pragma unroll
    for (int j = 0; j < SK_MAX_KB; ++j) {
      if (c0 + j < kbw) {
        uint4 xb[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) {
          const int n = i * 8 + g;
          xb[i] = n < p.n_tok ? __ldcg(reinterpret_cast<const uint4*>(p.X + static_cast<long long>(n) * p.ldx + k0 + (c0 + j) * 32)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < NG; ++i) {
          sk_mma(acc[i], wa[j].x, wb[j].x, wa[j].y, wb[j].y, xb[i].x, xb[i].y);
          sk_mma(acc[i], wa[j].z, wb[j].z, wa[j].w, wb[j].w, xb[i].z, xb[i].w);
        }
      }
    }
  };
#pragma unroll 1
  for (int c0 = 0; c0 < kbw; c0 += 2 * SK_MAX_KB) {
    if (c0 + SK_MAX_KB < kbw) load_w(ra2, rb2, c0 + SK_MAX_KB);
    compute(ra, rb, c0);
    if (c0 + 2 * SK_MAX_KB < kbw) load_w(ra, rb, c0 + 2 * SK_MAX_KB);
    if (c0 + SK_MAX_KB < kbw) compute(ra2, rb2, c0 + SK_MAX_KB);
  }
  // this warp's partial tile: [feature][token]
  float* my = s_part[warp];
#pragma unroll
  for (int i = 0; i < NG; ++i) {
    const int n = i * 8 + t * 2;
    my[g * 33 + n] = acc[i][0];
    my[g * 33 + n + 1] = acc[i][1];
    my[(g + 8) * 33 + n] = acc[i][2];
    my[(g + 8) * 33 + n + 1] = acc[i][3];
  }
  __syncthreads();
  // CTA partial -> s_part[0] (fixed warp order); thread -> token tid / 4, features (tid % 4) * 4 .. + 3
  const int tok = threadIdx.x >> 2, fq = (threadIdx.x & 3) * 4;
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int idx = (fq + j) * 33 + tok;
    v[j] = s_part[0][idx] + s_part[1][idx] + s_part[2][idx] + s_part[3][idx];
  }
  if (p.ksplit > 1) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) s_part[0][(fq + j) * 33 + tok] = v[j];
    sk_cluster_sync();  // every CTA's partial is in its own shared memory
    if (sk_cluster_rank() == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned a = smem_u32(&s_part[0][(fq + j) * 33 + tok]);
        float s = v[j];
        for (int r = 1; r < p.ksplit; ++r) s += sk_ld_dsmem(a, static_cast<unsigned>(r));
        v[j] = s;
      }
    }
    sk_cluster_sync();  // nobody leaves (and frees its shared memory) while rank 0 is still reading
    if (sk_cluster_rank() != 0) return;
  }
  if (tok >= p.n_tok || tok >= NG * 8) return;
  const int f = f0 + fq;
  if (f >= p.F) return;
  const float scale = p.out_scale == 0.f ? 1.f : p.out_scale;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (f + j < p.F) {
      float x = v[j];
      if (p.bias) x += ld_as_float(p.bias, p.bias_dt, f + j);
      if (p.act == VY_ACT_GELU_ERF) x = gelu_erf(x);
      else if (p.act == VY_ACT_GELU_TANH) x = gelu_tanh(x);
      if (p.addend) x += ld_as_float(p.addend, p.addend_dt, static_cast<long long>(tok) * p.ld_addend + f + j);
      v[j] = x * scale;
    }
  }
  const long long o = static_cast<long long>(tok) * p.ld_out + f;
  const bool full = f + 3 < p.F;
  if (p.out_dt == VY_BF16) {
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + o;
    if (full && (reinterpret_cast<uintptr_t>(op) & 7) == 0) {
      uint2 pk;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
      h[0] = __floats2bfloat162_rn(v[0], v[1]);
      h[1] = __floats2bfloat162_rn(v[2], v[3]);
      *reinterpret_cast<uint2*>(op) = pk;
    } else {
      for (int j = 0; j < 4 && f + j < p.F; ++j) op[j] = __float2bfloat16_rn(v[j]);
    }
  } else {
    float* op = reinterpret_cast<float*>(p.out) + o;
    if (full && (reinterpret_cast<uintptr_t>(op) & 15) == 0) *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
    else
      for (int j = 0; j < 4 && f + j < p.F; ++j) op[j] = v[j];
  }
}

// K split over the cluster: K / ksplit must be a multiple of 128 (4 warps x 32) and at most 768 per CTA; among the
// possible splits take the one that brings the CTA count closest to ~2 per SM without exceeding 8 CTAs per cluster.
static int skinny_ksplit(int F, int K) {
  const int units = (F + 15) / 16;
  static const int target_env = getenv("VY_SKINNY_CTAS") ? atoi(getenv("VY_SKINNY_CTAS")) : 0;  // development: sweep
  // one CTA per SM is enough while the whole matrix is a few MB (latency-bound; measured on the 768-wide decode step:
  // profiles/r02_decode_skinny_sweep_*.txt); tens of MB (the 2048 x 16384 Gemma projections) need several resident CTAs per
  // SM to keep enough loads in flight
  long long mb3 = static_cast<long long>(F) * K * 2 / (3 << 20);
  if (mb3 < 1) mb3 = 1;
  if (mb3 > 4) mb3 = 4;
  const int target = target_env > 0 ? target_env : num_sms() * static_cast<int>(mb3);
  int best = 0;
  long long best_score = -(1LL << 60);
  for (int s = 1; s <= 8; ++s) {
    if (K % s) continue;
    const int kc = K / s;
    if (kc % (SK_WARPS * 32)) continue;
    const long long ctas = static_cast<long long>(units) * s;
    long long score = ctas <= target ? ctas : target - (ctas - target) / 4;  // more CTAs up to the target, then mildly worse
    if (kc > SK_WARPS * SK_MAX_KB * 32) score -= num_sms() / 4;  // a second dependent round of weight loads per warp
    if (score > best_score) {
      best_score = score;
      best = s;
    }
  }
  return best;
}

bool skinny_applicable(const VyGemm* p) {
  static const bool off = getenv("VY_GEMM_SKINNY") && atoi(getenv("VY_GEMM_SKINNY")) == 0;  // development: A/B against the tcgen05 path
  if (off) return false;
  if (p->epi != VY_EPI_LINEAR || p->in_dtype != VY_BF16 || !p->transposed_out || p->a_mn_major || p->b_mn_major) return false;
  if (p->N > 32 || p->aux || p->addend2 || p->addend_row_mod || p->out_row_group) return false;
  if (p->act != VY_ACT_NONE && p->act != VY_ACT_GELU_ERF && p->act != VY_ACT_GELU_TANH) return false;
  // one wave of small CTAs is the whole point; a vocabulary-sized F (3142 feature blocks) would run ~4 latency-bound waves of
  // them and measured 2x slower than the 128-row tensor-memory tiles, which stream 77 MB at 3.5 TB/s
  if ((p->M + 15) / 16 > 4 * num_sms()) return false;
  return skinny_ksplit(p->M, p->K) > 0;
}

int launch_skinny(const VyGemm* p) {
  SkinnyDev d;
  memset(&d, 0, sizeof(d));
  d.F = p->M; d.K = p->K; d.n_tok = p->N;
  d.ksplit = skinny_ksplit(p->M, p->K);
  d.kb_per_warp = p->K / d.ksplit / (SK_WARPS * 32);
  d.static_w = p->weights_static != 0;
  d.W = static_cast<const __nv_bfloat16*>(p->A); d.ldw = p->lda;
  d.X = static_cast<const __nv_bfloat16*>(p->B); d.ldx = p->ldb;
  d.bias = p->bias; d.bias_dt = p->bias_dtype; d.act = p->act;
  d.addend = p->addend; d.ld_addend = p->ld_addend; d.addend_dt = p->addend_dtype;
  d.out_scale = p->out_scale; d.out = p->out; d.ld_out = p->ld_out; d.out_dt = p->out_dtype;
  const dim3 grid(d.ksplit, (p->M + 15) / 16), block(SK_THREADS);
  cudaStream_t st = static_cast<cudaStream_t>(p->stream);
  const int ng = (p->N + 7) / 8;
  cudaError_t e;
  switch (ng) {
    case 1: e = launch_kernel_cluster(gemm_skinny_kernel<1>, grid, block, 0, st, d.ksplit, d); break;
    case 2: e = launch_kernel_cluster(gemm_skinny_kernel<2>, grid, block, 0, st, d.ksplit, d); break;
    case 3: e = launch_kernel_cluster(gemm_skinny_kernel<3>, grid, block, 0, st, d.ksplit, d); break;
    default: e = launch_kernel_cluster(gemm_skinny_kernel<4>, grid, block, 0, st, d.ksplit, d); break;
  }
  VY_CUDA_OK(e);
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

}  // namespace vy

extern "C" int vy_gemm_is_small_batch(const VyGemm* p) { return p && p->M > 0 && p->N > 0 && p->K > 0 && vy::skinny_applicable(p) ? 1 : 0; }
