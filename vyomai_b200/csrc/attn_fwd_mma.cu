// vy_attn_fwd for head dims other than 64 (and for the prefix-LM mask): flash attention forward on mma.sync m16n8k16.
//
// The tensor-memory kernel (attn_fwd.cu) is specialised for head_dim 64, the only width the reference package's configs
// produce (hidden 768 / 12 heads). The PaliGemma-scale model of Examples/paligemma.ipynb needs two more: 72 (SigLIP:
// 1152 / 16 heads; cells 9 SiglipAttention) and 256 with ONE kv head (Gemma; cell 12 GemmaAttention), in inference only, where
// attention is a small share of the work (prefill: 264 tokens; decode: weights dominate). This kernel covers every head_dim
// that is a multiple of 8 up to 256 with one code path: 64 query rows per CTA (16 per warp), 64 keys per step through
// shared memory, Q K^T and P V on the legacy tensor-core path, online softmax in registers (log2 domain), fp32 accumulation.
//   * Masks: key padding, causal (k <= q_pos0 + l), prefix-LM (causal OR k < prefix_len[b]: cell 17 _update_causal_mask in
//     training form), all with the reference's additive finfo.min behaviour (a fully masked row averages v over all keys).
//   * GQA / MQA by head index. For single-token decode (Sq == 1) the n_rep query heads of one kv head become the ROWS of
//     one tile, so the cache is streamed once per kv head, not once per query head.
//   * q / k / v are addressed through (batch, head, token) element strides: packed projection outputs and kv-caches are
//     read in place.
#include "vy_common.cuh"
#include "vy_ptx.cuh"

namespace vy {

constexpr float AM_MASKED = -30000.0f;  // same finite "masked" score as attn_fwd.cu (log2 domain)
constexpr int AM_BM = 64, AM_BN = 64, AM_WARPS = 4;

struct AttnMmaDev {
  int B, Hq, Hkv, n_rep, D, Sq, Skv;
  const __nv_bfloat16* q;
  long long q_sb, q_sh, q_sl;
  const __nv_bfloat16* k;
  long long k_sb, k_sh, k_sl;
  const __nv_bfloat16* v;
  long long v_sb, v_sh, v_sl;
  int causal, q_pos0;
  const unsigned char* kpm;
  long long kpm_stride;
  const int* prefix_len;
  const int* pos_ptr;
  void* out;
  long long o_sb, o_sl;
  int out_dt;
  float* lse;
  float scale_log2;
  int pack_heads;  // Sq == 1: tile row r = query head kvh * n_rep + r
  int splits;      // pack_heads only: the key blocks are divided over gridDim.x CTAs, combined by the last one to finish
};

// split-key decode: partial (o[D], m, l) per (batch row, kv head, split, query head of the group) + a ticket per (row, kv head).
// Library-owned scratch (like the norm / squared-norm partials): calls that split must not run concurrently on two streams.
constexpr int AM_WS_FLOATS = 1 << 20;
constexpr int AM_MAX_TICKETS = 4096;
__device__ float am_ws[AM_WS_FLOATS];
__device__ unsigned int am_tickets[AM_MAX_TICKETS];

__device__ __forceinline__ void am_ldmatrix_x4(unsigned (&r)[4], unsigned addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void am_ldmatrix_x4_trans(unsigned (&r)[4], unsigned addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void am_mma(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned am_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<unsigned*>(&h);
}

// DP: head dim padded to a multiple of 16 (the k-step of Q K^T); rows in shared memory are DP * 2 + 16 bytes apart so that
// the 8 row addresses of an ldmatrix fall into distinct 16-byte bank groups.
template <int DP>
__global__ void __launch_bounds__(AM_WARPS * 32)
attn_fwd_mma_kernel(AttnMmaDev g) {
  extern __shared__ __align__(16) unsigned char am_smem[];
  constexpr int ROWB = DP * 2 + 16;
  unsigned char* sQ = am_smem;
  unsigned char* sK = sQ + AM_BM * ROWB;
  unsigned char* sV = sK + AM_BN * ROWB;
  pdl_trigger();
  pdl_wait();
  if (g.pos_ptr) {  // device-side position: keys [0, pos + Sq) of a cache with Skv slots
    const int pos = *g.pos_ptr;
    g.q_pos0 = pos;
    g.Skv = min(g.Skv, pos + g.Sq);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int b = blockIdx.z;
  const int m0 = blockIdx.x * AM_BM;
  // row r of the tile -> (query position l, query head h)
  const int h_blk = g.pack_heads ? blockIdx.y * g.n_rep : blockIdx.y;  // first / only query head of this CTA
  const int kvh = g.pack_heads ? blockIdx.y : blockIdx.y / g.n_rep;
  const int rows_valid = g.pack_heads ? g.n_rep : min(AM_BM, g.Sq - m0);
  const int vec_per_row = g.D >> 3;  // 16-byte vectors of real data per row

  // ---- Q tile -> smem (zero beyond D and beyond the valid rows) ----
  for (int i = threadIdx.x; i < AM_BM * (DP / 8); i += blockDim.x) {
    const int r = i / (DP / 8), c = i - r * (DP / 8);
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < rows_valid && c < vec_per_row) {
      const long long off = g.pack_heads ? static_cast<long long>(b) * g.q_sb + static_cast<long long>(h_blk + r) * g.q_sh
                                         : static_cast<long long>(b) * g.q_sb + static_cast<long long>(h_blk) * g.q_sh + static_cast<long long>(m0 + r) * g.q_sl;
      val = *reinterpret_cast<const uint4*>(g.q + off + c * 8);
    }
    *reinterpret_cast<uint4*>(sQ + r * ROWB + c * 16) = val;
  }

  float o_acc[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  // query position of this thread's two rows (gq, gq + 8 of the warp's 16)
  int qpos[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int r = warp * 16 + gq + e * 8;
    qpos[e] = g.q_pos0 + (g.pack_heads ? 0 : m0 + r);
  }
  const int prefix = g.prefix_len ? g.prefix_len[b] : 0;
  const unsigned char* kpm = g.kpm ? g.kpm + static_cast<long long>(b) * g.kpm_stride : nullptr;
  const __nv_bfloat16* kbase = g.k + static_cast<long long>(b) * g.k_sb + static_cast<long long>(kvh) * g.k_sh;
  const __nv_bfloat16* vbase = g.v + static_cast<long long>(b) * g.v_sb + static_cast<long long>(kvh) * g.v_sh;

  const int n_blocks = (g.Skv + AM_BN - 1) / AM_BN;
  int kb_lo = 0, kb_hi = n_blocks;
  if (g.splits > 1) {  // this CTA's share of the key blocks (possibly empty)
    const int per = (n_blocks + g.splits - 1) / g.splits;
    kb_lo = min(n_blocks, static_cast<int>(blockIdx.x) * per);
    kb_hi = min(n_blocks, kb_lo + per);
    __syncthreads();  // Q stored (the loop's first barrier may not run)
  }
  for (int kb = kb_lo; kb < kb_hi; ++kb) {
    const int k0 = kb * AM_BN;
    __syncthreads();  // previous block's K / V fully consumed (and, first time, Q stored)
    // 4 row-vectors of K and of V per thread in flight (8 independent 16-byte loads) before the first shared-memory store
    constexpr int KV_VECS = AM_BN * (DP / 8);
    for (int i0 = threadIdx.x; i0 < KV_VECS; i0 += 4 * AM_WARPS * 32) {
      uint4 kvv[4], vvv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * AM_WARPS * 32;
        const int r = i / (DP / 8), c = i - r * (DP / 8);
        kvv[u] = make_uint4(0, 0, 0, 0);
        vvv[u] = make_uint4(0, 0, 0, 0);
        if (i < KV_VECS && k0 + r < g.Skv && c < vec_per_row) {
          kvv[u] = *reinterpret_cast<const uint4*>(kbase + static_cast<long long>(k0 + r) * g.k_sl + c * 8);
          vvv[u] = *reinterpret_cast<const uint4*>(vbase + static_cast<long long>(k0 + r) * g.v_sl + c * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * AM_WARPS * 32;
        const int r = i / (DP / 8), c = i - r * (DP / 8);
        if (i < KV_VECS) {
          *reinterpret_cast<uint4*>(sK + r * ROWB + c * 16) = kvv[u];
          *reinterpret_cast<uint4*>(sV + r * ROWB + c * 16) = vvv[u];
        }
      }
    }
    __syncthreads();

    // ---- S = Q K^T for this warp's 16 rows x 64 keys ----
    float s_acc[AM_BN / 8][4];
#pragma unroll
    for (int i = 0; i < AM_BN / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s_acc[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < DP / 16; ++ks) {
      unsigned a[4];
      // A fragment: rows (lane % 16) of the warp's 16, 16-byte column block (lane / 16) of this k-step
      am_ldmatrix_x4(a, smem_u32(sQ + (warp * 16 + (lane & 15)) * ROWB + ks * 32 + (lane >> 4) * 16));
#pragma unroll
      for (int np = 0; np < AM_BN / 16; ++np) {
        // two key n-tiles (16 keys) x 16 d: matrices (keys 0-7, d 0-7), (keys 0-7, d 8-15), (keys 8-15, d 0-7), (keys 8-15, d 8-15)
        unsigned kf[4];
        am_ldmatrix_x4(kf, smem_u32(sK + (np * 16 + (lane & 7) + ((lane >> 4) << 3)) * ROWB + ks * 32 + ((lane >> 3) & 1) * 16));
        am_mma(s_acc[2 * np], a, kf[0], kf[1]);
        am_mma(s_acc[2 * np + 1], a, kf[2], kf[3]);
      }
    }

    // ---- scale, mask, online softmax ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < AM_BN / 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = j >> 1;
        const int kk = k0 + nt * 8 + tq * 2 + (j & 1);
        float sc;
        if (kk >= g.Skv) {
          sc = -INFINITY;  // tile padding: not a key at all
        } else {
          bool vis = kpm ? kpm[kk] != 0 : true;
          if (g.causal) vis = vis && (kk <= qpos[e] || kk < prefix);
          sc = vis ? s_acc[nt][j] * g.scale_log2 : AM_MASKED;
        }
        s_acc[nt][j] = sc;
        mx[e] = fmaxf(mx[e], sc);
      }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      mx[e] = fmaxf(mx[e], __shfl_xor_sync(0xffffffffu, mx[e], 1));
      mx[e] = fmaxf(mx[e], __shfl_xor_sync(0xffffffffu, mx[e], 2));
    }
    float alpha[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float mn = fmaxf(m_run[e], mx[e]);  // finite: every block holds at least one key (masked keys count, finite)
      alpha[e] = exp2f(m_run[e] - mn);
      m_run[e] = mn;
    }
#pragma unroll
    for (int nt = 0; nt < AM_BN / 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = exp2f(s_acc[nt][j] - m_run[j >> 1]);
        s_acc[nt][j] = p;
        rs[j >> 1] += p;
      }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      rs[e] += __shfl_xor_sync(0xffffffffu, rs[e], 1);
      rs[e] += __shfl_xor_sync(0xffffffffu, rs[e], 2);
      l_run[e] = l_run[e] * alpha[e] + rs[e];
    }
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      o_acc[i][0] *= alpha[0];
      o_acc[i][1] *= alpha[0];
      o_acc[i][2] *= alpha[1];
      o_acc[i][3] *= alpha[1];
    }

    // ---- O += P V: P comes straight from the score accumulators (their layout is the A-fragment layout) ----
#pragma unroll
    for (int kk2 = 0; kk2 < AM_BN / 16; ++kk2) {
      unsigned pa[4];
      pa[0] = am_pack(s_acc[2 * kk2][0], s_acc[2 * kk2][1]);
      pa[1] = am_pack(s_acc[2 * kk2][2], s_acc[2 * kk2][3]);
      pa[2] = am_pack(s_acc[2 * kk2 + 1][0], s_acc[2 * kk2 + 1][1]);
      pa[3] = am_pack(s_acc[2 * kk2 + 1][2], s_acc[2 * kk2 + 1][3]);
#pragma unroll
      for (int dp = 0; dp < DP / 16; ++dp) {
        // V rows = keys (the k index of this MMA), transposed on load: matrices (keys 0-7, d 0-7), (keys 8-15, d 0-7),
        // (keys 0-7, d 8-15), (keys 8-15, d 8-15)
        unsigned vf[4];
        am_ldmatrix_x4_trans(vf, smem_u32(sV + (kk2 * 16 + (lane & 7) + (((lane >> 3) & 1) << 3)) * ROWB + dp * 32 + (lane >> 4) * 16));
        am_mma(o_acc[2 * dp], pa, vf[0], vf[1]);
        am_mma(o_acc[2 * dp + 1], pa, vf[2], vf[3]);
      }
    }
  }

  if (g.splits > 1) {
    // ---- partial results -> scratch; the last CTA of this (batch row, kv head) combines them in split order ----
    const long long slot0 = (static_cast<long long>(b) * g.Hkv + kvh) * g.splits;
    const int stride = g.D + 2;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int r = warp * 16 + gq + e * 8;
      if (r >= rows_valid) continue;
      float* w = am_ws + ((slot0 + blockIdx.x) * g.n_rep + r) * stride;
#pragma unroll
      for (int i = 0; i < DP / 8; ++i) {
        const int d = i * 8 + tq * 2;
        if (d < g.D) {
          w[d] = o_acc[i][2 * e];
          w[d + 1] = o_acc[i][2 * e + 1];
        }
      }
      if (tq == 0) {
        w[g.D] = m_run[e];
        w[g.D + 1] = l_run[e];
      }
    }
    __shared__ unsigned int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(&am_tickets[b * g.Hkv + kvh], 1u);
      s_last = prev == static_cast<unsigned int>(g.splits - 1) ? 1u : 0u;
      if (s_last) am_tickets[b * g.Hkv + kvh] = 0u;  // self-reset for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // one (query head, 4 dims) item per thread and pass; all loads of an item are independent and issued together
    constexpr int MAXS = 16;
    for (int idx = threadIdx.x; idx < g.n_rep * (g.D >> 2); idx += blockDim.x) {
      const int r = idx / (g.D >> 2), d = (idx - r * (g.D >> 2)) << 2;
      float ms[MAXS], ls[MAXS];
      float4 os[MAXS];
#pragma unroll
      for (int sp = 0; sp < MAXS; ++sp) {
        if (sp < g.splits) {
          const float* w = am_ws + ((slot0 + sp) * g.n_rep + r) * stride;
          ms[sp] = __ldcg(w + g.D);
          ls[sp] = __ldcg(w + g.D + 1);
          os[sp] = make_float4(__ldcg(w + d), __ldcg(w + d + 1), __ldcg(w + d + 2), __ldcg(w + d + 3));
        } else {
          ms[sp] = -INFINITY;
          ls[sp] = 0.f;
          os[sp] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float M = -INFINITY;
#pragma unroll
      for (int sp = 0; sp < MAXS; ++sp) M = fmaxf(M, ms[sp]);
      float L = 0.f;
      float4 O = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int sp = 0; sp < MAXS; ++sp) {
        const float a = ms[sp] == -INFINITY ? 0.f : exp2f(ms[sp] - M);  // (an empty share has weight 0)
        L += ls[sp] * a;
        O.x += os[sp].x * a; O.y += os[sp].y * a; O.z += os[sp].z * a; O.w += os[sp].w * a;
      }
      const float inv = 1.f / L;
      const long long ob = static_cast<long long>(b) * g.o_sb + static_cast<long long>(h_blk + r) * g.D + d;
      st_from_float(g.out, g.out_dt, ob, O.x * inv);
      st_from_float(g.out, g.out_dt, ob + 1, O.y * inv);
      st_from_float(g.out, g.out_dt, ob + 2, O.z * inv);
      st_from_float(g.out, g.out_dt, ob + 3, O.w * inv);
    }
    return;
  }

  // ---- epilogue: O / l -> out[b, l, h * D + d] ----
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int r = warp * 16 + gq + e * 8;
    if (r >= rows_valid) continue;
    const int h = g.pack_heads ? h_blk + r : h_blk;
    const int l = g.pack_heads ? 0 : m0 + r;
    const float inv = 1.f / l_run[e];
    const long long obase = static_cast<long long>(b) * g.o_sb + static_cast<long long>(l) * g.o_sl + static_cast<long long>(h) * g.D;
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      const int d = i * 8 + tq * 2;
      if (d < g.D) {
        const float x0 = o_acc[i][2 * e] * inv, x1 = o_acc[i][2 * e + 1] * inv;
        if (g.out_dt == VY_BF16) {
          *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(g.out) + obase + d) = __floats2bfloat162_rn(x0, x1);
        } else {
          float* op = reinterpret_cast<float*>(g.out) + obase + d;
          op[0] = x0;
          op[1] = x1;
        }
      }
    }
    if (g.lse && tq == 0) g.lse[(static_cast<long long>(b) * g.Hq + h) * g.Sq + l] = m_run[e] + log2f(l_run[e]);
  }
}

template <int DP>
static int launch_attn_mma(const AttnMmaDev& g, cudaStream_t st) {
  constexpr int ROWB = DP * 2 + 16;
  const size_t smem = static_cast<size_t>(AM_BM + 2 * AM_BN) * ROWB;
  static bool attr_done[64] = {};
  int dev = 0;
  VY_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && !attr_done[dev]) {
    VY_CUDA_OK(cudaFuncSetAttribute(attn_fwd_mma_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_done[dev] = true;
  }
  const dim3 grid(g.pack_heads ? g.splits : (g.Sq + AM_BM - 1) / AM_BM, g.pack_heads ? g.Hkv : g.Hq, g.B);
  VY_CUDA_OK(launch_kernel(attn_fwd_mma_kernel<DP>, grid, dim3(AM_WARPS * 32), smem, st, g));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

// called by vy_attn_fwd (attn_fwd.cu) after its argument checks
int attn_fwd_mma(const VyAttn* p) {
  auto aligned16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
  VY_CHECK_ARG(p->head_dim >= 8 && p->head_dim <= 256 && p->head_dim % 8 == 0, "vy_attn_fwd: head_dim %d must be a multiple of 8 in [8, 256]", p->head_dim);
  VY_CHECK_ARG(p->qkv_dtype == VY_BF16, "vy_attn_fwd: q / k / v must be bf16");
  VY_CHECK_ARG(p->q_sl % 8 == 0 && p->q_sh % 8 == 0 && p->q_sb % 8 == 0 && p->k_sl % 8 == 0 && p->k_sh % 8 == 0 && p->k_sb % 8 == 0 &&
                   p->v_sl % 8 == 0 && p->v_sh % 8 == 0 && p->v_sb % 8 == 0 && aligned16(p->q) && aligned16(p->k) && aligned16(p->v),
               "vy_attn_fwd: q / k / v strides and pointers must keep 16-byte alignment");
  VY_CHECK_ARG((p->head_dim & 1) == 0 && (p->o_sl & 1) == 0 && (p->o_sb & 1) == 0, "vy_attn_fwd: output strides must be even");
  AttnMmaDev g;
  memset(&g, 0, sizeof(g));
  g.B = p->B; g.Hq = p->n_q_heads; g.Hkv = p->n_kv_heads; g.n_rep = p->n_q_heads / p->n_kv_heads; g.D = p->head_dim;
  g.Sq = p->Sq; g.Skv = p->Skv;
  g.q = static_cast<const __nv_bfloat16*>(p->q); g.q_sb = p->q_sb; g.q_sh = p->q_sh; g.q_sl = p->q_sl;
  g.k = static_cast<const __nv_bfloat16*>(p->k); g.k_sb = p->k_sb; g.k_sh = p->k_sh; g.k_sl = p->k_sl;
  g.v = static_cast<const __nv_bfloat16*>(p->v); g.v_sb = p->v_sb; g.v_sh = p->v_sh; g.v_sl = p->v_sl;
  g.causal = p->causal; g.q_pos0 = p->q_pos0; g.kpm = p->key_padding_mask; g.kpm_stride = p->kpm_stride;
  g.prefix_len = p->causal ? p->prefix_len : nullptr;
  g.pos_ptr = p->pos_ptr;
  g.out = p->out; g.o_sb = p->o_sb; g.o_sl = p->o_sl; g.out_dt = p->out_dtype; g.lse = p->lse;
  g.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(p->head_dim));
  // single-token decode of a grouped model: the group's query heads share one tile (a causal mask is vacuous at Sq == 1
  // only when the query sits at the end of the keys, which is what q_pos0 + 1 == Skv says)
  g.pack_heads = (p->Sq == 1 && g.n_rep > 1 && g.n_rep <= AM_BM && (!p->causal || p->pos_ptr || p->q_pos0 + 1 >= p->Skv)) ? 1 : 0;
  if (g.pack_heads) g.causal = 0;
  g.splits = 1;
  if (g.pack_heads && !p->lse) {
    // few (row, kv head) pairs: one CTA each would stream the whole context alone (65 us per Gemma layer at batch 1)
    const int pairs = p->B * p->n_kv_heads, blocks = (p->Skv + AM_BN - 1) / AM_BN;
    int sp = pairs >= num_sms() ? 1 : (num_sms() + pairs - 1) / pairs;
    if (sp > blocks) sp = blocks;
    if (sp > 16) sp = 16;
    const long long need = static_cast<long long>(pairs) * sp * g.n_rep * (p->head_dim + 2);
    if (sp > 1 && need <= AM_WS_FLOATS && pairs <= AM_MAX_TICKETS) g.splits = sp;
  }
  cudaStream_t st = static_cast<cudaStream_t>(p->stream);
  const int dp = (p->head_dim + 15) / 16 * 16;
  if (dp <= 64) return launch_attn_mma<64>(g, st);
  if (dp <= 80) return launch_attn_mma<80>(g, st);
  if (dp <= 128) return launch_attn_mma<128>(g, st);
  if (dp <= 192) return launch_attn_mma<192>(g, st);
  return launch_attn_mma<256>(g, st);
}

}  // namespace vy
