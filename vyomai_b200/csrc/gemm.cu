// vy_gemm: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   warp 0      TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx-count)
//   warp 1      MMA issuer    (one thread, tcgen05.mma cta_group::1, fp32 accumulators in TMEM)
//   warp 2      TMEM allocator
//   warps 4..7  epilogue      (tcgen05.ld -> registers -> fused bias/act/residual/RoPE -> global)
//
// Two TMEM accumulator buffers let the epilogue of tile i overlap the mainloop of tile i+1.
// Tile = 128 x BN x (128 bytes of K): BK = 64 bf16 or 32 tf32 elements, so every smem stage has
// the same byte geometry for both input types. Operands may be K-major (nn.Linear layout) or
// MN-major (transposed storage, used by dgrad / wgrad) — only the TMA box and the UMMA
// descriptor change.
#include <mutex>
#include <unordered_map>
#include <vector>

#include "vy_common.cuh"
#include "vy_ptx.cuh"

namespace vy {

struct GemmDev {
  int M, N, K;
  int epi, act, transposed_out;
  const void* bias;
  int bias_dtype;
  const void* addend;
  long long ld_addend;
  int addend_dtype, addend_row_mod, addend_row_off;
  const void* addend2;
  long long ld_addend2;
  int addend2_dtype;
  void* aux;
  long long ld_aux;
  int aux_dtype;
  float out_scale;
  void* out;
  long long ld_out;
  int out_dtype, out_row_group, out_row_group_stride, out_row_off;
  int vec_ok;  // all row strides / bases allow 8-element vector access
  // qkv rope
  int tokens_per_seq, start_pos, kv_dst_pos0, n_q_heads, n_kv_heads;
  const float* rope_cos;
  const float* rope_sin;
  void* q_out;
  long long q_sb, q_sh, q_sl;
  void* k_out;
  long long k_sb, k_sh, k_sl;
  void* v_out;
  long long v_sb, v_sh, v_sl;
  int kv_out_dtype;
};

template <typename TIn, int BN_>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BN = BN_;
  static constexpr int EPB = 128 / sizeof(TIn);  // elements per 128-byte swizzle row
  static constexpr int BK = EPB;
  static constexpr int UMMA_K = 32 / sizeof(TIn);
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN >= 256 ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN <= 64) ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512));
  static constexpr int MN_BOX_BYTES = BK * 128;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 2 * BN * 4 /*bias*/ + 256;
  static constexpr int FMT = sizeof(TIn) == 2 ? 1 : 2;  // bf16 : tf32
  // MN-major operands: bf16 uses the plain 128B swizzle (8 k-rows per 1024-B group); tf32 must use
  // SWIZZLE_128B_BASE32B (4 k-rows per 512-B group) and the matching TMA mode.
  static constexpr int MN_SBO = sizeof(TIn) == 2 ? 1024 : 512;
  static constexpr int MN_LAYOUT = sizeof(TIn) == 2 ? 2 : 1;
  static constexpr int MN_TMA_SWIZZLE = sizeof(TIn) == 2 ? 1 : 2;
};

__device__ __forceinline__ float apply_act(int act, float x) {
  if (act == VY_ACT_GELU_ERF) return gelu_erf(x);
  if (act == VY_ACT_GELU_TANH) return gelu_tanh(x);
  return x;
}
__device__ __forceinline__ float apply_dact(int act, float z) {
  return act == VY_ACT_DGELU_ERF ? dgelu_erf(z) : dgelu_tanh(z);
}

__device__ __forceinline__ long long remap_out_row(const GemmDev& g, int r) {
  if (g.out_row_group > 0)
    return static_cast<long long>(r / g.out_row_group) * g.out_row_group_stride +
           (r % g.out_row_group) + g.out_row_off;
  return r;
}
__device__ __forceinline__ long long remap_add_row(const GemmDev& g, int r) {
  if (g.addend_row_mod > 0) return g.addend_row_off + (r % g.addend_row_mod);
  return r;
}

// --------------------------------------------------------------------------------------------
// epilogue bodies (executed by 128 threads; thread <-> accumulator row)
// --------------------------------------------------------------------------------------------
template <int BN>
__device__ __forceinline__ void epilogue_linear(const GemmDev& g, uint32_t tmem_acc, int m0, int n0,
                                                int row_in_tile, const float* bias_s,
                                                uint64_t* tmem_empty_bar) {
  const int grow = m0 + row_in_tile;
  const bool row_ok = grow < g.M;
  const float scale = g.out_scale == 0.f ? 1.f : g.out_scale;
  const bool fwd_act = g.act == VY_ACT_GELU_ERF || g.act == VY_ACT_GELU_TANH;
  const bool bwd_act = g.act == VY_ACT_DGELU_ERF || g.act == VY_ACT_DGELU_TANH;
  const long long orow = remap_out_row(g, grow);
  const long long arow = remap_add_row(g, grow);

#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    uint32_t raw[32];
    tmem_ld_x32(tmem_acc + c * 32, raw);
    tmem_ld_wait();
    if (c == BN / 32 - 1) {
      tc_fence_before();
      mbar_arrive(tmem_empty_bar);
    }
    const int gcol0 = n0 + c * 32;
    const int nvalid = g.N - gcol0;
    if (!row_ok || nvalid <= 0) continue;
    if (g.vec_ok && nvalid >= 32) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(raw[q * 8 + j]) + bias_s[c * 32 + q * 8 + j];
        const int col = gcol0 + q * 8;
        if (fwd_act) {
          if (g.aux) st8_from_float(g.aux, g.aux_dtype, static_cast<long long>(grow) * g.ld_aux + col, x);
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = apply_act(g.act, x[j]);
        } else if (bwd_act) {
          float z[8];
          ld8_as_float(g.aux, g.aux_dtype, static_cast<long long>(grow) * g.ld_aux + col, z);
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] *= apply_dact(g.act, z[j]);
        }
        if (g.addend) {
          float a[8];
          ld8_as_float(g.addend, g.addend_dtype, arow * g.ld_addend + col, a);
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] += a[j];
        }
        if (g.addend2) {
          float a[8];
          ld8_as_float(g.addend2, g.addend2_dtype, static_cast<long long>(grow) * g.ld_addend2 + col, a);
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] += a[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] *= scale;
        st8_from_float(g.out, g.out_dtype, orow * g.ld_out + col, x);
      }
    } else {
      const int lim = nvalid < 32 ? nvalid : 32;
      for (int j = 0; j < lim; ++j) {
        const int col = gcol0 + j;
        float x = __uint_as_float(raw[j]) + bias_s[c * 32 + j];
        if (fwd_act) {
          if (g.aux) st_from_float(g.aux, g.aux_dtype, static_cast<long long>(grow) * g.ld_aux + col, x);
          x = apply_act(g.act, x);
        } else if (bwd_act) {
          x *= apply_dact(g.act, ld_as_float(g.aux, g.aux_dtype, static_cast<long long>(grow) * g.ld_aux + col));
        }
        if (g.addend) x += ld_as_float(g.addend, g.addend_dtype, arow * g.ld_addend + col);
        if (g.addend2) x += ld_as_float(g.addend2, g.addend2_dtype, static_cast<long long>(grow) * g.ld_addend2 + col);
        st_from_float(g.out, g.out_dtype, orow * g.ld_out + col, x * scale);
      }
    }
  }
}

// swap-AB epilogue: accumulator row = logical output COLUMN (a weight row), accumulator column =
// logical output ROW (a token). Stores are scalar per thread but coalesced across the warp.
template <int BN>
__device__ __forceinline__ void epilogue_transposed(const GemmDev& g, uint32_t tmem_acc, int m0,
                                                    int n0, int row_in_tile,
                                                    uint64_t* tmem_empty_bar) {
  const int lc = m0 + row_in_tile;  // logical column
  const bool ok = lc < g.M;
  const float scale = g.out_scale == 0.f ? 1.f : g.out_scale;
  const bool fwd_act = g.act == VY_ACT_GELU_ERF || g.act == VY_ACT_GELU_TANH;
  const bool bwd_act = g.act == VY_ACT_DGELU_ERF || g.act == VY_ACT_DGELU_TANH;
  const float b = (ok && g.bias) ? ld_as_float(g.bias, g.bias_dtype, lc) : 0.f;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    uint32_t raw[32];
    tmem_ld_x32(tmem_acc + c * 32, raw);
    tmem_ld_wait();
    if (c == BN / 32 - 1) {
      tc_fence_before();
      mbar_arrive(tmem_empty_bar);
    }
    if (!ok) continue;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int lr = n0 + c * 32 + j;  // logical row
      if (lr < g.N) {
        float x = __uint_as_float(raw[j]) + b;
        if (fwd_act) {
          if (g.aux) st_from_float(g.aux, g.aux_dtype, static_cast<long long>(lr) * g.ld_aux + lc, x);
          x = apply_act(g.act, x);
        } else if (bwd_act) {
          x *= apply_dact(g.act, ld_as_float(g.aux, g.aux_dtype, static_cast<long long>(lr) * g.ld_aux + lc));
        }
        if (g.addend) x += ld_as_float(g.addend, g.addend_dtype, remap_add_row(g, lr) * g.ld_addend + lc);
        if (g.addend2) x += ld_as_float(g.addend2, g.addend2_dtype, static_cast<long long>(lr) * g.ld_addend2 + lc);
        st_from_float(g.out, g.out_dtype, remap_out_row(g, lr) * g.ld_out + lc, x * scale);
      }
    }
  }
}

// QKV projection epilogue: bias + in-register half-split RoPE + head-split scatter (+ kv-cache
// append through k_out/v_out strides). head_dim == 64: one 64-column group is one head.
template <int BN>
__device__ __forceinline__ void epilogue_qkv_rope(const GemmDev& g, uint32_t tmem_acc, int m0, int n0,
                                                  int row_in_tile, const float* bias_s,
                                                  uint64_t* tmem_empty_bar) {
  const int grow = m0 + row_in_tile;
  const bool row_ok = grow < g.M;
  const int b = row_ok ? grow / g.tokens_per_seq : 0;
  const int l = row_ok ? grow % g.tokens_per_seq : 0;
  const int pos = g.start_pos + l;
  const float* cs = g.rope_cos ? g.rope_cos + static_cast<long long>(pos) * 32 : nullptr;
  const float* sn = g.rope_sin ? g.rope_sin + static_cast<long long>(pos) * 32 : nullptr;
#pragma unroll 1
  for (int hgrp = 0; hgrp < BN / 64; ++hgrp) {
    uint32_t lo[32], hi[32];
    tmem_ld_x32(tmem_acc + hgrp * 64, lo);
    tmem_ld_x32(tmem_acc + hgrp * 64 + 32, hi);
    tmem_ld_wait();
    if (hgrp == BN / 64 - 1) {
      tc_fence_before();
      mbar_arrive(tmem_empty_bar);
    }
    const int gcol0 = n0 + hgrp * 64;
    if (!row_ok || gcol0 >= g.N) continue;
    const int head = gcol0 >> 6;
    void* dst;
    long long off;
    bool rotate;
    int dst_dt = g.kv_out_dtype;
    if (head < g.n_q_heads) {
      dst_dt = g.out_dtype;
      dst = g.q_out;
      off = b * g.q_sb + head * g.q_sh + l * g.q_sl;
      rotate = cs != nullptr;
    } else if (head < g.n_q_heads + g.n_kv_heads) {
      dst = g.k_out;
      off = b * g.k_sb + (head - g.n_q_heads) * g.k_sh + static_cast<long long>(g.kv_dst_pos0 + l) * g.k_sl;
      rotate = cs != nullptr;
    } else {
      dst = g.v_out;
      off = b * g.v_sb + (head - g.n_q_heads - g.n_kv_heads) * g.v_sh +
            static_cast<long long>(g.kv_dst_pos0 + l) * g.v_sl;
      rotate = false;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float x1[8], x2[8], o1[8], o2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        x1[j] = __uint_as_float(lo[q * 8 + j]) + bias_s[hgrp * 64 + q * 8 + j];
        x2[j] = __uint_as_float(hi[q * 8 + j]) + bias_s[hgrp * 64 + 32 + q * 8 + j];
      }
      if (rotate) {
        const float4 c0 = *reinterpret_cast<const float4*>(cs + q * 8);
        const float4 c1 = *reinterpret_cast<const float4*>(cs + q * 8 + 4);
        const float4 s0 = *reinterpret_cast<const float4*>(sn + q * 8);
        const float4 s1 = *reinterpret_cast<const float4*>(sn + q * 8 + 4);
        const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o1[j] = x1[j] * cc[j] - x2[j] * ss[j];
          o2[j] = x2[j] * cc[j] + x1[j] * ss[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o1[j] = x1[j];
          o2[j] = x2[j];
        }
      }
      st8_from_float(dst, dst_dt, off + q * 8, o1);
      st8_from_float(dst, dst_dt, off + 32 + q * 8, o2);
    }
  }
}

// --------------------------------------------------------------------------------------------
// kernel
// --------------------------------------------------------------------------------------------
template <typename TIn, int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(256, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const GemmDev g) {
  using Cfg = GemmCfg<TIn, BN>;
  constexpr int BM = Cfg::BM;
  constexpr int BK = Cfg::BK;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  float* bias_s = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 2 * BN);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (g.M + BM - 1) / BM;
  const int n_tiles = (g.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (g.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_s, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM;
        const int n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          uint8_t* a_dst = sA + s * Cfg::A_BYTES;
          uint8_t* b_dst = sB + s * Cfg::B_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d(a_dst, &tma_a, &full_bar[s], kb * BK, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BM / Cfg::EPB; ++i)
              tma_load_2d(a_dst + i * Cfg::MN_BOX_BYTES, &tma_a, &full_bar[s], m0 + i * Cfg::EPB, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(b_dst, &tma_b, &full_bar[s], kb * BK, n0);
          } else {
#pragma unroll
            for (int i = 0; i < BN / Cfg::EPB; ++i)
              tma_load_2d(b_dst + i * Cfg::MN_BOX_BYTES, &tma_b, &full_bar[s], n0 + i * Cfg::EPB, kb * BK);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(Cfg::FMT, BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      uint32_t it = 0;
      uint32_t local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const uint32_t acc = local & 1;
        const uint32_t acc_ph = (local >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + s * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / Cfg::UMMA_K; ++k) {
            const uint64_t ad = A_MN ? make_smem_desc_sw128(a_addr + k * Cfg::UMMA_K * 128, Cfg::MN_BOX_BYTES, Cfg::MN_SBO, Cfg::MN_LAYOUT)
                                     : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? make_smem_desc_sw128(b_addr + k * Cfg::UMMA_K * 128, Cfg::MN_BOX_BYTES, Cfg::MN_SBO, Cfg::MN_LAYOUT)
                                     : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            if constexpr (sizeof(TIn) == 2) umma_f16(d_tmem, ad, bd, idesc, (kb | k) != 0);
            else umma_tf32(d_tmem, ad, bd, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp - 4;  // == warp % 4: TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    const int et = threadIdx.x - 128;
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const uint32_t acc = local & 1;
      const uint32_t acc_ph = (local >> 1) & 1;
      const int m0 = (tile / n_tiles) * BM;
      const int n0 = (tile % n_tiles) * BN;
      float* bs = bias_s + acc * BN;
      if (!g.transposed_out) {
        for (int j = et; j < BN; j += 128) {
          const int col = n0 + j;
          bs[j] = (g.bias && col < g.N) ? ld_as_float(g.bias, g.bias_dtype, col) : 0.f;
        }
        named_bar_sync(1, 128);
      }
      mbar_wait(&tfull_bar[acc], acc_ph);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      if (g.epi == VY_EPI_QKV_ROPE) {
        if constexpr (BN >= 64) epilogue_qkv_rope<BN>(g, tmem_acc, m0, n0, row_in_tile, bs, &tempty_bar[acc]);
      } else if (g.transposed_out) {
        epilogue_transposed<BN>(g, tmem_acc, m0, n0, row_in_tile, &tempty_bar[acc]);
      } else {
        epilogue_linear<BN>(g, tmem_acc, m0, n0, row_in_tile, bs, &tempty_bar[acc]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
static int get_tmap_2d(CUtensorMap* out, int dtype, const void* base, uint64_t d0, uint64_t d1,
                       uint64_t stride1_bytes, uint32_t b0, uint32_t b1, int swz = 1) {
  uint64_t dims[2] = {d0, d1};
  uint64_t strides[2] = {0, stride1_bytes};
  uint32_t box[2] = {b0, b1};
  return get_tensor_map_cached(out, dtype, 2, base, dims, strides, box, swz);
}

template <typename TIn, int BN, bool A_MN, bool B_MN>
static int launch_gemm(const VyGemm* p, const GemmDev& g) {
  using Cfg = GemmCfg<TIn, BN>;
  const int dt = p->in_dtype;
  const size_t es = sizeof(TIn);
  CUtensorMap ta, tb;
  int rc;
  if (!A_MN)
    rc = get_tmap_2d(&ta, dt, p->A, p->K, p->M, p->lda * es, Cfg::BK, Cfg::BM);
  else
    rc = get_tmap_2d(&ta, dt, p->A, p->M, p->K, p->lda * es, Cfg::EPB, Cfg::BK, Cfg::MN_TMA_SWIZZLE);
  if (rc != VY_OK) return rc;
  if (!B_MN)
    rc = get_tmap_2d(&tb, dt, p->B, p->K, p->N, p->ldb * es, Cfg::BK, BN);
  else
    rc = get_tmap_2d(&tb, dt, p->B, p->N, p->K, p->ldb * es, Cfg::EPB, Cfg::BK, Cfg::MN_TMA_SWIZZLE);
  if (rc != VY_OK) return rc;

  auto kern = gemm_kernel<TIn, BN, A_MN, B_MN>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    VY_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int m_tiles = (p->M + Cfg::BM - 1) / Cfg::BM;
  const int n_tiles = (p->N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, 256, Cfg::SMEM_BYTES, static_cast<cudaStream_t>(p->stream)>>>(ta, tb, g);
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

template <typename TIn>
static int dispatch_gemm(const VyGemm* p, const GemmDev& g, int bn) {
  const bool amn = p->a_mn_major != 0, bmn = p->b_mn_major != 0;
  if (!amn && !bmn) {
    switch (bn) {
      case 32: return launch_gemm<TIn, 32, false, false>(p, g);
      case 64: return launch_gemm<TIn, 64, false, false>(p, g);
      case 128: return launch_gemm<TIn, 128, false, false>(p, g);
      default: return launch_gemm<TIn, 256, false, false>(p, g);
    }
  }
  if (bn < 128) bn = 128;
  if (!amn && bmn) return bn == 128 ? launch_gemm<TIn, 128, false, true>(p, g) : launch_gemm<TIn, 256, false, true>(p, g);
  if (amn && !bmn) return bn == 128 ? launch_gemm<TIn, 128, true, false>(p, g) : launch_gemm<TIn, 256, true, false>(p, g);
  return bn == 128 ? launch_gemm<TIn, 128, true, true>(p, g) : launch_gemm<TIn, 256, true, true>(p, g);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace vy

extern "C" int vy_gemm(const VyGemm* p) {
  using namespace vy;
  VY_CHECK_ARG(p != nullptr, "vy_gemm: null params");
  if (!vy_device_ok()) {
    set_error("vy_gemm: no sm_100 device (there is no CPU fallback)");
    return VY_ERR_NO_DEVICE;
  }
  VY_CHECK_ARG(p->M > 0 && p->N > 0 && p->K > 0, "vy_gemm: bad shape M=%d N=%d K=%d", p->M, p->N, p->K);
  VY_CHECK_ARG(dtype_ok(p->in_dtype), "vy_gemm: bad in_dtype %d", p->in_dtype);
  VY_CHECK_ARG(p->A && p->B, "vy_gemm: null operand");
  const size_t es = dtype_size(p->in_dtype);
  VY_CHECK_ARG(aligned16(p->A) && aligned16(p->B), "vy_gemm: A/B must be 16-byte aligned");
  VY_CHECK_ARG((p->lda * es) % 16 == 0 && (p->ldb * es) % 16 == 0,
               "vy_gemm: lda/ldb (%lld, %lld) must be multiples of 16 bytes", (long long)p->lda, (long long)p->ldb);
  VY_CHECK_ARG(p->lda >= (p->a_mn_major ? p->M : p->K) && p->ldb >= (p->b_mn_major ? p->N : p->K),
               "vy_gemm: leading dimension smaller than the contiguous extent");

  GemmDev g;
  memset(&g, 0, sizeof(g));
  g.M = p->M; g.N = p->N; g.K = p->K;
  g.epi = p->epi; g.act = p->act; g.transposed_out = p->transposed_out;
  g.bias = p->bias; g.bias_dtype = p->bias_dtype;
  g.addend = p->addend; g.ld_addend = p->ld_addend; g.addend_dtype = p->addend_dtype;
  g.addend_row_mod = p->addend_row_mod; g.addend_row_off = p->addend_row_off;
  g.addend2 = p->addend2; g.ld_addend2 = p->ld_addend2; g.addend2_dtype = p->addend2_dtype;
  g.aux = p->aux; g.ld_aux = p->ld_aux; g.aux_dtype = p->aux_dtype;
  g.out_scale = p->out_scale;
  g.out = p->out; g.ld_out = p->ld_out; g.out_dtype = p->out_dtype;
  g.out_row_group = p->out_row_group; g.out_row_group_stride = p->out_row_group_stride;
  g.out_row_off = p->out_row_off;

  if (p->epi == VY_EPI_LINEAR) {
    VY_CHECK_ARG(p->out != nullptr && dtype_ok(p->out_dtype), "vy_gemm: out / out_dtype invalid");
    VY_CHECK_ARG(p->act >= VY_ACT_NONE && p->act <= VY_ACT_DGELU_TANH, "vy_gemm: bad act %d", p->act);
    if (p->act == VY_ACT_DGELU_ERF || p->act == VY_ACT_DGELU_TANH)
      VY_CHECK_ARG(p->aux != nullptr, "vy_gemm: DGELU epilogue needs aux (saved pre-activation)");
    if (p->bias) VY_CHECK_ARG(dtype_ok(p->bias_dtype), "vy_gemm: bad bias_dtype");
    if (p->addend) VY_CHECK_ARG(dtype_ok(p->addend_dtype), "vy_gemm: bad addend_dtype");
    if (p->addend2) VY_CHECK_ARG(dtype_ok(p->addend2_dtype), "vy_gemm: bad addend2_dtype");
    if (p->aux) VY_CHECK_ARG(dtype_ok(p->aux_dtype), "vy_gemm: bad aux_dtype");
    auto vec = [](const void* ptr, long long ld, int dt) {
      if (!ptr) return true;
      return aligned16(ptr) && (ld * (long long)dtype_size(dt)) % 16 == 0;
    };
    // fp32 vector access moves 8 floats = 32 B as two 16-B halves: 16-B alignment suffices.
    g.vec_ok = vec(p->out, p->ld_out, p->out_dtype) && vec(p->aux, p->ld_aux, p->aux_dtype) &&
               vec(p->addend, p->ld_addend, p->addend_dtype) && vec(p->addend2, p->ld_addend2, p->addend2_dtype);
  } else if (p->epi == VY_EPI_QKV_ROPE) {
    VY_CHECK_ARG(!p->transposed_out, "vy_gemm: QKV_ROPE epilogue cannot be transposed");
    VY_CHECK_ARG(p->head_dim == 64, "vy_gemm: QKV_ROPE epilogue supports head_dim 64 (got %d)", p->head_dim);
    VY_CHECK_ARG(p->N == (p->n_q_heads + 2 * p->n_kv_heads) * 64, "vy_gemm: QKV_ROPE N mismatch");
    VY_CHECK_ARG(p->tokens_per_seq > 0 && p->M % p->tokens_per_seq == 0, "vy_gemm: M %% tokens_per_seq != 0");
    VY_CHECK_ARG((p->q_out || p->n_q_heads == 0) && ((p->k_out && p->v_out) || p->n_kv_heads == 0) && dtype_ok(p->out_dtype),
                 "vy_gemm: q/k/v outputs missing");
    VY_CHECK_ARG(p->start_pos >= 0 && p->kv_dst_pos0 >= 0, "vy_gemm: negative position");
    VY_CHECK_ARG((p->rope_cos == nullptr) == (p->rope_sin == nullptr), "vy_gemm: rope_cos/rope_sin must both be set or NULL");
    VY_CHECK_ARG(dtype_ok(p->kv_out_dtype), "vy_gemm: bad kv_out_dtype");
    auto okstr = [&](const void* ptr, long long sb, long long sh, long long sl, int dt) {
      const long long es_o = dtype_size(dt);
      return aligned16(ptr) && (sb * es_o) % 16 == 0 && (sh * es_o) % 16 == 0 && (sl * es_o) % 16 == 0;
    };
    VY_CHECK_ARG(okstr(p->q_out, p->q_sb, p->q_sh, p->q_sl, p->out_dtype) &&
                     okstr(p->k_out, p->k_sb, p->k_sh, p->k_sl, p->kv_out_dtype) &&
                     okstr(p->v_out, p->v_sb, p->v_sh, p->v_sl, p->kv_out_dtype),
                 "vy_gemm: q/k/v strides must keep 16-byte alignment");
    if (p->bias) VY_CHECK_ARG(dtype_ok(p->bias_dtype), "vy_gemm: bad bias_dtype");
    g.tokens_per_seq = p->tokens_per_seq; g.start_pos = p->start_pos; g.kv_dst_pos0 = p->kv_dst_pos0;
    g.n_q_heads = p->n_q_heads; g.n_kv_heads = p->n_kv_heads;
    g.rope_cos = p->rope_cos; g.rope_sin = p->rope_sin;
    g.q_out = p->q_out; g.q_sb = p->q_sb; g.q_sh = p->q_sh; g.q_sl = p->q_sl;
    g.k_out = p->k_out; g.k_sb = p->k_sb; g.k_sh = p->k_sh; g.k_sl = p->k_sl;
    g.v_out = p->v_out; g.v_sb = p->v_sb; g.v_sh = p->v_sh; g.v_sl = p->v_sl;
    g.kv_out_dtype = p->kv_out_dtype;
  } else {
    set_error("vy_gemm: unknown epilogue %d", p->epi);
    return VY_ERR_INVALID_ARG;
  }

  // tile-N choice: narrow tiles for swap-AB decode GEMMs, wide tiles once they still fill the GPU
  int bn;
  const int m_tiles = (p->M + 127) / 128;
  if (p->N <= 32) bn = 32;
  else if (p->N <= 64) bn = 64;
  else if (p->N <= 128) bn = 128;
  else bn = (m_tiles * ((p->N + 255) / 256) >= num_sms()) ? 256 : 128;
  if (p->epi == VY_EPI_QKV_ROPE && bn < 64) bn = 64;

  if (p->in_dtype == VY_BF16) return dispatch_gemm<__nv_bfloat16>(p, g, bn);
  return dispatch_gemm<float>(p, g, bn);
}
