// vy_attn_bwd: flash-attention backward on tcgen05 / TMEM, operands fed by TMA. Three kernels:
//
//   1. dsum      D[b,h,l] = sum_j dO[b,l,h,j] * O[b,l,h,j]                  (HBM-bound, one pass)
//   2. dkdv      one CTA per (128-key tile, kv head, batch row); loops over the n_rep query heads that
//                share the kv head and over query tiles:
//                    S^T = K Q^T, dP^T = V dO^T          (M = keys, N = queries; accumulators in TMEM)
//                    P^T = exp2(S^T * c - lse), dS^T = P^T * (dP^T - D)     (thread <-> key row)
//                    dV += P^T dO, dK += dS^T Q           (accumulated in TMEM across the whole loop)
//                GQA needs no atomics: the loop over the query heads of the group is inside the CTA.
//   3. dq        one CTA per (128-query tile, q head, batch row); loops over key tiles:
//                    S = Q K^T, dP = dO V^T, dS = P * (dP - D), dQ += dS K
//                (recomputing S / dP here instead of atomically accumulating dQ from kernel 2 costs
//                 two extra small MMAs per tile pair and keeps the result deterministic)
// The epilogues multiply by 1/sqrt(d), undo the RoPE rotation of q / k (transpose rotation) and
// write straight into the packed [tokens, (Hq + 2 Hkv) * 64] gradient of the fused q|k|v projection.
// Masks follow vy_attn_fwd exactly (finite "finfo.min" scores, -inf beyond Skv).
#include <stdlib.h>

#include "vy_common.cuh"
#include "vy_ptx.cuh"

namespace vy {

constexpr int AB_T = 128;                    // tile edge (queries and keys)
constexpr int AB_D = 64;
constexpr int AB_TILE = AB_T * AB_D * 2;     // 16 KB
constexpr int AB_PBYTES = AB_T * AB_T * 2;   // 32 KB
constexpr float AB_MASKED = -30000.0f;  // must equal AT_MASKED of attn_fwd.cu
constexpr int AB_SMEM_DKDV = 2 * AB_TILE /*K,V*/ + 4 * AB_TILE /*Q,dO x2*/ + 2 * AB_PBYTES + 2 * 2 * AB_T * 4 + 256;
constexpr int AB_SMEM_DQ = 2 * AB_TILE /*Q,dO*/ + 4 * AB_TILE /*K,V x2*/ + AB_PBYTES + 256;

struct AttnBwdDev {
  int B, Hq, Hkv, Sq, Skv, n_rep, causal, q_pos0;
  const unsigned char* kpm;
  long long kpm_sb;
  const float* lse;
  const float* dsum;
  void* dq;
  long long ld_dq;
  void* dk;
  long long ld_dk;
  void* dv;
  long long ld_dv;
  int out_dtype;
  const float* rope_cos;
  const float* rope_sin;
  int rope_pos0;
  float scale_log2, scale;
};

// ---------------------------------------------------------------------------------------------
// 1. D = rowsum(dO * O) per (b, head, token). One warp per token row, 8 lanes per head.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_dsum_kernel(int B, int Hq, int Sq, const void* __restrict__ o, long long o_sb, long long o_sl, int o_dt,
                 const void* __restrict__ dout, long long do_sb, long long do_sl, int do_dt, float* __restrict__ dsum) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nchunks = Hq * 8;  // 8-element chunks per row
  for (long long row = warp; row < static_cast<long long>(B) * Sq; row += nwarps) {
    const int b = static_cast<int>(row / Sq), l = static_cast<int>(row % Sq);
    for (int c0 = 0; c0 < nchunks; c0 += 32) {
      const int c = c0 + lane;
      float acc = 0.f;
      if (c < nchunks) {
        float a[8], g[8];
        ld8_as_float(o, o_dt, b * o_sb + l * o_sl + c * 8, a);
        ld8_as_float(dout, do_dt, b * do_sb + l * do_sl + c * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += a[j] * g[j];
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if ((lane & 7) == 0 && c < nchunks) dsum[(static_cast<long long>(b) * Hq + (c >> 3)) * Sq + l] = acc;
    }
  }
}

// bf16 pair -> swizzled 16-byte chunk store helper: writes 32 packed values (4 chunks) of row `row`
// into a K-major [128 x 128] bf16 operand made of two [128 x 64] 128B-swizzle atoms.
__device__ __forceinline__ void store_p_chunk(uint8_t* base, int row, int c4, const uint32_t (&packed)[16]) {
  uint8_t* rowp = base + (c4 >> 1) * (AB_T * 128) + row * 128;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const int chunk = ((c4 & 1) * 4 + ch) ^ (row & 7);
    *reinterpret_cast<uint4*>(rowp + chunk * 16) =
        make_uint4(packed[ch * 4], packed[ch * 4 + 1], packed[ch * 4 + 2], packed[ch * 4 + 3]);
  }
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// write one 64-wide head row (thread-private fp32 values) with optional inverse RoPE and scaling
__device__ __forceinline__ void store_head_row(void* dst, int dt, long long off, float (&v)[AB_D], float scale,
                                               const float* cs, const float* sn) {
#pragma unroll
  for (int q8 = 0; q8 < 4; ++q8) {
    float o1[8], o2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = v[q8 * 8 + j] * scale, b = v[32 + q8 * 8 + j] * scale;
      if (cs) {
        const float c = cs[q8 * 8 + j], s = sn[q8 * 8 + j];
        o1[j] = a * c + b * s;   // transpose of the forward rotation
        o2[j] = b * c - a * s;
      } else {
        o1[j] = a;
        o2[j] = b;
      }
    }
    st8_from_float(dst, dt, off + q8 * 8, o1);
    st8_from_float(dst, dt, off + 32 + q8 * 8, o2);
  }
}

__device__ __forceinline__ bool bwd_noskip(const AttnBwdDev& g, int b) {
  return !g.causal || (g.kpm && g.kpm[b * g.kpm_sb] == 0);
}

// ---------------------------------------------------------------------------------------------
// 2. dK / dV
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                     const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                     const AttnBwdDev g) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + AB_TILE;
  uint8_t* sQ = smem + 2 * AB_TILE;    // [2]
  uint8_t* sdO = smem + 4 * AB_TILE;   // [2]
  uint8_t* sP = smem + 6 * AB_TILE;
  uint8_t* sdS = sP + AB_PBYTES;
  float* s_lse = reinterpret_cast<float*>(sdS + AB_PBYTES);  // [2][128]
  float* s_D = s_lse + 2 * AB_T;                              // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_D + 2 * AB_T);
  uint64_t* kv_full = bars;        // [1]
  uint64_t* qdo_full = bars + 1;   // [2]
  uint64_t* qdo_empty = bars + 3;  // [2]
  uint64_t* sdp_full = bars + 5;   // [1]
  uint64_t* pds_full = bars + 6;   // [1]
  uint64_t* pds_empty = bars + 7;  // [1]
  uint64_t* acc_full = bars + 8;   // [1]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * AB_T;
  const int kvh = blockIdx.y;
  const int b = blockIdx.z;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    tma_prefetch_desc(&tma_do);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qdo_full[s], 1);
      mbar_init(&qdo_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 128);
    mbar_init(pds_empty, 1);
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  pdl_wait();  // everything above is independent of the predecessor grid's output
  const uint32_t tm_S = tmem_base, tm_dP = tmem_base + 128, tm_dV = tmem_base + 256, tm_dK = tmem_base + 320;

  const int q_tiles = (g.Sq + AB_T - 1) / AB_T;
  int qt_begin = 0;
  if (!bwd_noskip(g, b)) {
    const int first_q = kv0 - g.q_pos0;  // first query row that can see key kv0
    if (first_q > 0) qt_begin = first_q / AB_T;
    if (qt_begin > q_tiles) qt_begin = q_tiles;
  }
  const int per_head = q_tiles - qt_begin;
  const int n_it = per_head * g.n_rep;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * AB_TILE);
      tma_load_4d(sK, &tma_k, kv_full, 0, kv0, kvh, b);
      tma_load_4d(sV, &tma_v, kv_full, 0, kv0, kvh, b);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        const int head = kvh * g.n_rep + it / per_head;
        const int q0 = (qt_begin + it % per_head) * AB_T;
        mbar_wait(&qdo_empty[s], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[s], 2 * AB_TILE);
        tma_load_4d(sQ + s * AB_TILE, &tma_q, &qdo_full[s], 0, q0, head, b);
        tma_load_4d(sdO + s * AB_TILE, &tma_do, &qdo_full[s], 0, q0, head, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_sc = make_idesc(1, AB_T, AB_T, 0, 0);   // K-major x K-major, N = 128
      constexpr uint32_t idesc_acc = make_idesc(1, AB_T, AB_D, 0, 1);  // A K-major, B MN-major, N = 64
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
      mbar_wait(kv_full, 0);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        const uint32_t q_addr = smem_u32(sQ + s * AB_TILE), do_addr = smem_u32(sdO + s * AB_TILE);
        mbar_wait(&qdo_full[s], (it >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AB_D / 16; ++k)
          umma_f16(tm_S, make_smem_desc_sw128(k_addr + k * 32, 16, 1024), make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                   idesc_sc, k != 0);
#pragma unroll
        for (int k = 0; k < AB_D / 16; ++k)
          umma_f16(tm_dP, make_smem_desc_sw128(v_addr + k * 32, 16, 1024), make_smem_desc_sw128(do_addr + k * 32, 16, 1024),
                   idesc_sc, k != 0);
        umma_commit(sdp_full);
        mbar_wait(pds_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < AB_T / 16; ++kk) {
          const uint32_t a_off = (kk >> 2) * (AB_T * 128) + (kk & 3) * 32;
          umma_f16(tm_dV, make_smem_desc_sw128(p_addr + a_off, 16, 1024),
                   make_smem_desc_sw128(do_addr + kk * 2048, 8192, 1024), idesc_acc, (it | kk) != 0);
        }
#pragma unroll
        for (int kk = 0; kk < AB_T / 16; ++kk) {
          const uint32_t a_off = (kk >> 2) * (AB_T * 128) + (kk & 3) * 32;
          umma_f16(tm_dK, make_smem_desc_sw128(ds_addr + a_off, 16, 1024),
                   make_smem_desc_sw128(q_addr + kk * 2048, 8192, 1024), idesc_acc, (it | kk) != 0);
        }
        umma_commit(&qdo_empty[s]);
        umma_commit(pds_empty);
      }
      umma_commit(acc_full);
    }
  } else if (warp >= 4) {
    const int qd = warp - 4;
    const int row = qd * 32 + lane;  // key row of this thread
    const int key = kv0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const int et = threadIdx.x - 128;
    const bool key_in = key < g.Skv;
    const bool key_vis = key_in && (!g.kpm || g.kpm[b * g.kpm_sb + key] != 0);

    for (int it = 0; it < n_it; ++it) {
      const int head = kvh * g.n_rep + it / per_head;
      const int q0 = (qt_begin + it % per_head) * AB_T;
      float* lse_s = s_lse + (it & 1) * AB_T;
      float* d_s = s_D + (it & 1) * AB_T;
      {
        const int qq = q0 + et;
        const long long idx = (static_cast<long long>(b) * g.Hq + head) * g.Sq + qq;
        lse_s[et] = qq < g.Sq ? g.lse[idx] : 0.f;
        d_s[et] = qq < g.Sq ? g.dsum[idx] : 0.f;
      }
      named_bar_sync(1, 128);
      mbar_wait(sdp_full, it & 1);
      tc_fence_after();
      mbar_wait(pds_empty, (it & 1) ^ 1);
#pragma unroll 1
      for (int c4 = 0; c4 < AB_T / 32; ++c4) {
        uint32_t sraw[32], praw[32];
        tmem_ld_x32(tm_S + lane_off + c4 * 32, sraw);
        tmem_ld_x32(tm_dP + lane_off + c4 * 32, praw);
        tmem_ld_wait();
        uint32_t pp[16], dd[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float pv[2], dv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = c4 * 32 + 2 * i + e;
            const int qq = q0 + c;
            const bool vis = key_vis && (!g.causal || key <= g.q_pos0 + qq);
            const float t = vis ? __uint_as_float(sraw[2 * i + e]) * g.scale_log2 : AB_MASKED;
            float p = exp2f(t - lse_s[c]);
            if (!key_in || qq >= g.Sq) p = 0.f;
            pv[e] = p;
            dv[e] = p * (__uint_as_float(praw[2 * i + e]) - d_s[c]);
          }
          pp[i] = pack_bf16(pv[0], pv[1]);
          dd[i] = pack_bf16(dv[0], dv[1]);
        }
        store_p_chunk(sP, row, c4, pp);
        store_p_chunk(sdS, row, c4, dd);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_full);
    }

    // epilogue: dV, dK of this key row
    float acc[AB_D];
    if (n_it > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    const long long tok = static_cast<long long>(b) * g.Skv + key;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      if (n_it > 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t raw[32];
          tmem_ld_x32((which == 0 ? tm_dV : tm_dK) + lane_off + h * 32, raw);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[h * 32 + i] = __uint_as_float(raw[i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < AB_D; ++i) acc[i] = 0.f;
      }
      if (key_in) {
        if (which == 0) {
          store_head_row(g.dv, g.out_dtype, tok * g.ld_dv + kvh * AB_D, acc, 1.f, nullptr, nullptr);
        } else {
          const float* cs = g.rope_cos ? g.rope_cos + static_cast<long long>(g.rope_pos0 + key) * 32 : nullptr;
          const float* sn = g.rope_sin ? g.rope_sin + static_cast<long long>(g.rope_pos0 + key) * 32 : nullptr;
          store_head_row(g.dk, g.out_dtype, tok * g.ld_dk + kvh * AB_D, acc, g.scale, cs, sn);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// 3. dQ
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                   const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                   const AttnBwdDev g) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + AB_TILE;
  uint8_t* sK = smem + 2 * AB_TILE;  // [2]
  uint8_t* sV = smem + 4 * AB_TILE;  // [2]
  uint8_t* sdS = smem + 6 * AB_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + AB_PBYTES);
  uint64_t* qdo_full = bars;       // [1]
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* sdp_full = bars + 5;   // [1]
  uint64_t* ds_full = bars + 6;    // [1]
  uint64_t* ds_empty = bars + 7;   // [1]
  uint64_t* acc_full = bars + 8;   // [1]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AB_T;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int kvh = head / g.n_rep;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    tma_prefetch_desc(&tma_do);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(qdo_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(ds_full, 128);
    mbar_init(ds_empty, 1);
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  pdl_wait();  // everything above is independent of the predecessor grid's output
  const uint32_t tm_S = tmem_base, tm_dP = tmem_base + 128, tm_dQ = tmem_base + 256;

  const int total = (g.Skv + AB_T - 1) / AB_T;
  int n_tiles = total;
  if (!bwd_noskip(g, b)) {
    const int last_q = min(q0 + AB_T, g.Sq) - 1;
    const int need = (g.q_pos0 + last_q) / AB_T + 1;
    if (need < n_tiles) n_tiles = need;
  }

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(qdo_full, 2 * AB_TILE);
      tma_load_4d(sQ, &tma_q, qdo_full, 0, q0, head, b);
      tma_load_4d(sdO, &tma_do, qdo_full, 0, q0, head, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * AB_TILE);
        tma_load_4d(sK + s * AB_TILE, &tma_k, &kv_full[s], 0, j * AB_T, kvh, b);
        tma_load_4d(sV + s * AB_TILE, &tma_v, &kv_full[s], 0, j * AB_T, kvh, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_sc = make_idesc(1, AB_T, AB_T, 0, 0);
      constexpr uint32_t idesc_acc = make_idesc(1, AB_T, AB_D, 0, 1);
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sdO), ds_addr = smem_u32(sdS);
      mbar_wait(qdo_full, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t k_addr = smem_u32(sK + s * AB_TILE), v_addr = smem_u32(sV + s * AB_TILE);
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AB_D / 16; ++k)
          umma_f16(tm_S, make_smem_desc_sw128(q_addr + k * 32, 16, 1024), make_smem_desc_sw128(k_addr + k * 32, 16, 1024),
                   idesc_sc, k != 0);
#pragma unroll
        for (int k = 0; k < AB_D / 16; ++k)
          umma_f16(tm_dP, make_smem_desc_sw128(do_addr + k * 32, 16, 1024), make_smem_desc_sw128(v_addr + k * 32, 16, 1024),
                   idesc_sc, k != 0);
        umma_commit(sdp_full);
        mbar_wait(ds_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < AB_T / 16; ++kk) {
          const uint32_t a_off = (kk >> 2) * (AB_T * 128) + (kk & 3) * 32;
          umma_f16(tm_dQ, make_smem_desc_sw128(ds_addr + a_off, 16, 1024),
                   make_smem_desc_sw128(k_addr + kk * 2048, 8192, 1024), idesc_acc, (j | kk) != 0);
        }
        umma_commit(&kv_empty[s]);
        umma_commit(ds_empty);
      }
      umma_commit(acc_full);
    }
  } else if (warp >= 4) {
    const int qd = warp - 4;
    const int row = qd * 32 + lane;
    const int qrow = q0 + row;
    const int qpos = g.q_pos0 + qrow;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const bool q_in = qrow < g.Sq;
    const long long sidx = (static_cast<long long>(b) * g.Hq + head) * g.Sq + qrow;
    const float lse = q_in ? g.lse[sidx] : 0.f;
    const float dsum = q_in ? g.dsum[sidx] : 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      const int kv0 = j * AB_T;
      uint32_t kb[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      if (g.kpm) {
        uint32_t w = 0;
        const int kbase = kv0 + lane * 4;
        const unsigned char* kp = g.kpm + b * g.kpm_sb;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (kbase + k < g.Skv && kp[kbase + k]) w |= 1u << k;
#pragma unroll
        for (int k = 0; k < 4; ++k) kb[k] = __ballot_sync(0xffffffffu, (w >> k) & 1u);
      }
      mbar_wait(sdp_full, j & 1);
      tc_fence_after();
      mbar_wait(ds_empty, (j & 1) ^ 1);
#pragma unroll 1
      for (int c4 = 0; c4 < AB_T / 32; ++c4) {
        uint32_t sraw[32], praw[32];
        tmem_ld_x32(tm_S + lane_off + c4 * 32, sraw);
        tmem_ld_x32(tm_dP + lane_off + c4 * 32, praw);
        tmem_ld_wait();
        uint32_t dd[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float dv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = c4 * 32 + 2 * i + e;
            const int key = kv0 + c;
            const bool vis = ((kb[c & 3] >> (c >> 2)) & 1u) && (!g.causal || key <= qpos);
            const float t = vis ? __uint_as_float(sraw[2 * i + e]) * g.scale_log2 : AB_MASKED;
            float p = exp2f(t - lse);
            if (key >= g.Skv || !q_in) p = 0.f;
            dv[e] = p * (__uint_as_float(praw[2 * i + e]) - dsum);
          }
          dd[i] = pack_bf16(dv[0], dv[1]);
        }
        store_p_chunk(sdS, row, c4, dd);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_full);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float acc[AB_D];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t raw[32];
      tmem_ld_x32(tm_dQ + lane_off + h * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[h * 32 + i] = __uint_as_float(raw[i]);
    }
    if (q_in) {
      // self-attention without a cache: query row l and key row l are the same token, rope row = rope_pos0 + l
      const float* cs = g.rope_cos ? g.rope_cos + static_cast<long long>(g.rope_pos0 + qrow) * 32 : nullptr;
      const float* sn = g.rope_sin ? g.rope_sin + static_cast<long long>(g.rope_pos0 + qrow) * 32 : nullptr;
      store_head_row(g.dq, g.out_dtype, (static_cast<long long>(b) * g.Sq + qrow) * g.ld_dq + head * AB_D, acc, g.scale, cs, sn);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int attn_bwd_fused_launch(const VyAttnBwd* p);  // attn_bwd_fused.cu: 1 = handled, 0 = shape not covered, < 0 = error

static int make_map4(CUtensorMap* out, const void* base, int S, int H, int B, long long sb, long long sh, long long sl) {
  uint64_t dims[4] = {static_cast<uint64_t>(AB_D), static_cast<uint64_t>(S), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
  uint64_t strides[4] = {0, static_cast<uint64_t>(sl) * 2, static_cast<uint64_t>(sh) * 2, static_cast<uint64_t>(sb) * 2};
  uint32_t box[4] = {AB_D, AB_T, 1, 1};
  return get_tensor_map_cached(out, VY_BF16, 4, base, dims, strides, box, 1);
}

}  // namespace vy

extern "C" int vy_attn_bwd(const VyAttnBwd* p) {
  using namespace vy;
  VY_CHECK_ARG(p != nullptr, "vy_attn_bwd: null params");
  if (!vy_device_ok()) {
    set_error("vy_attn_bwd: no sm_100 device (there is no CPU fallback)");
    return VY_ERR_NO_DEVICE;
  }
  VY_CHECK_ARG(p->head_dim == 64, "vy_attn_bwd: head_dim must be 64 (got %d)", p->head_dim);
  VY_CHECK_ARG(p->B > 0 && p->Sq > 0 && p->Skv > 0 && p->n_q_heads > 0 && p->n_kv_heads > 0 && p->n_q_heads % p->n_kv_heads == 0,
               "vy_attn_bwd: bad shape");
  VY_CHECK_ARG(p->q && p->k && p->v && p->o && p->dout && p->lse && p->dsum && p->dq && p->dk && p->dv, "vy_attn_bwd: null pointer");
  VY_CHECK_ARG(dtype_ok(p->o_dtype) && dtype_ok(p->out_dtype), "vy_attn_bwd: bad dtype");
  VY_CHECK_ARG((p->rope_cos == nullptr) == (p->rope_sin == nullptr), "vy_attn_bwd: rope tables must both be set or NULL");
  auto ok16 = [](const void* ptr, long long a, long long b_, long long c) {
    return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (a * 2) % 16 == 0 && (b_ * 2) % 16 == 0 && (c * 2) % 16 == 0;
  };
  VY_CHECK_ARG(ok16(p->q, p->q_sb, p->q_sh, p->q_sl) && ok16(p->k, p->k_sb, p->k_sh, p->k_sl) && ok16(p->v, p->v_sb, p->v_sh, p->v_sl) &&
                   ok16(p->dout, p->do_sb, p->do_sl, 64),
               "vy_attn_bwd: q/k/v/dout pointers and strides must keep 16-byte alignment");
  const long long eso = dtype_size(p->out_dtype);
  VY_CHECK_ARG((p->ld_dq * eso) % 16 == 0 && (p->ld_dk * eso) % 16 == 0 && (p->ld_dv * eso) % 16 == 0 &&
                   (reinterpret_cast<uintptr_t>(p->dq) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->dk) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(p->dv) & 15) == 0,
               "vy_attn_bwd: dq/dk/dv must keep 16-byte alignment");

  cudaStream_t st = static_cast<cudaStream_t>(p->stream);
  {
    const long long rows = static_cast<long long>(p->B) * p->Sq;
    long long blocks = (rows + 7) / 8;
    const long long cap = static_cast<long long>(num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    VY_CUDA_OK(launch_kernel(attn_dsum_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, st, p->B, p->n_q_heads, p->Sq, p->o, p->o_sb, p->o_sl, p->o_dtype,
                                                               p->dout, p->do_sb, p->do_sl, VY_BF16, p->dsum));
    VY_LAUNCH_OK();
  }

  {
    // short sequences: one fused kernel per (batch row, kv head); VY_ATTN_BWD_FUSED=0 forces the two-kernel path
    static const bool fused_on = !(getenv("VY_ATTN_BWD_FUSED") && atoi(getenv("VY_ATTN_BWD_FUSED")) == 0);
    if (fused_on) {
      const int f = attn_bwd_fused_launch(p);
      if (f < 0) return f;
      if (f == 1) {
        count_launch(2);
        return VY_OK;
      }
    }
  }

  CUtensorMap tq, tk, tv, tdo;
  int rc = make_map4(&tq, p->q, p->Sq, p->n_q_heads, p->B, p->q_sb, p->q_sh, p->q_sl);
  if (rc != VY_OK) return rc;
  rc = make_map4(&tk, p->k, p->Skv, p->n_kv_heads, p->B, p->k_sb, p->k_sh, p->k_sl);
  if (rc != VY_OK) return rc;
  rc = make_map4(&tv, p->v, p->Skv, p->n_kv_heads, p->B, p->v_sb, p->v_sh, p->v_sl);
  if (rc != VY_OK) return rc;
  rc = make_map4(&tdo, p->dout, p->Sq, p->n_q_heads, p->B, p->do_sb, 64, p->do_sl);
  if (rc != VY_OK) return rc;

  AttnBwdDev g;
  g.B = p->B; g.Hq = p->n_q_heads; g.Hkv = p->n_kv_heads; g.Sq = p->Sq; g.Skv = p->Skv;
  g.n_rep = p->n_q_heads / p->n_kv_heads; g.causal = p->causal; g.q_pos0 = p->q_pos0;
  g.kpm = p->key_padding_mask; g.kpm_sb = p->kpm_stride;
  g.lse = p->lse; g.dsum = p->dsum;
  g.dq = p->dq; g.ld_dq = p->ld_dq; g.dk = p->dk; g.ld_dk = p->ld_dk; g.dv = p->dv; g.ld_dv = p->ld_dv;
  g.out_dtype = p->out_dtype;
  g.rope_cos = p->rope_cos; g.rope_sin = p->rope_sin; g.rope_pos0 = p->rope_pos0;
  g.scale = 0.125f;
  g.scale_log2 = 1.4426950408889634f * 0.125f;

  static bool attr_set = false;
  if (!attr_set) {
    VY_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_DKDV));
    VY_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_DQ));
    attr_set = true;
  }
  dim3 grid_kv((p->Skv + AB_T - 1) / AB_T, p->n_kv_heads, p->B);
  VY_CUDA_OK(launch_kernel(attn_bwd_dkdv_kernel, dim3(grid_kv), dim3(256), AB_SMEM_DKDV, st, tq, tk, tv, tdo, g));
  VY_LAUNCH_OK();
  dim3 grid_q((p->Sq + AB_T - 1) / AB_T, p->n_q_heads, p->B);
  VY_CUDA_OK(launch_kernel(attn_bwd_dq_kernel, dim3(grid_q), dim3(256), AB_SMEM_DQ, st, tq, tk, tv, tdo, g));
  VY_LAUNCH_OK();
  count_launch(3);
  return VY_OK;
}
