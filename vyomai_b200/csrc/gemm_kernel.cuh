// vy_gemm: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   warp 0      TMEM allocator, then TMA producer (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx-count)
//   warp 1      MMA issuer    (one thread, tcgen05.mma, fp32 accumulators in TMEM; cta_group::2 in PAIR kernels: see gemm_kernel)
//   warps 2..9  epilogue      (tcgen05.ld -> registers -> fused bias/act/residual/RoPE/SwiGLU -> smem staging
//                              -> per-warp TMA bulk tensor stores (fast epilogues) or coalesced global stores;
//                              two warps per TMEM lane quarter)
//
// Two TMEM accumulator buffers let the epilogue of tile i overlap the mainloop of tile i+1; the row
// operand an epilogue reads (residual, saved pre-activation) is prefetched before the accumulator is
// awaited. Tile = 128 x BN x (128 bytes of K) on one SM, or 256 x BN on a CTA pair (two SMs, one M = 256 MMA,
// each CTA staging half of B: see gemm_kernel); BN in {32, 64, 128, 192, 256}. Flavour, BN and K split are chosen
// per call (gemm.cu: choose_tiling, VyGemm.hint_*): BK = 64 bf16 or 32 tf32 elements, so every smem stage has
// the same byte geometry for both input types. Operands may be K-major (nn.Linear layout) or
// MN-major (transposed storage, used by dgrad / wgrad) — only the TMA box and the UMMA
// descriptor change.
#pragma once
#include <stdlib.h>
#include "vy_common.cuh"
#include "vy_ptx.cuh"

namespace vy {

struct GemmDev {
  int M, N, K;
  int epi, act, transposed_out;
  int gate_act;  // VY_ACT_SWIGLU family: 0 = silu(gate) * up, 1 = gelu_tanh(gate) * up (GeGLU)
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
  int k_splits, kb_per_split;  // split-K: unit = (tile, split); raw fp32 partial tiles go to ws[split][M][N]
  float* ws;
  PoisonRef poison;  // raised by a wait that timed out (mbar_wait_soft); the host sees it in vy_gemm / vy_gemm_poison_peek / vy_gemm_poisoned
  int tma_store;  // fast epilogues write back with TMA stores (tma_out / tma_aux of the launch)
  int debug;  // development switches (VY_GEMM_DEBUG): 1 = epilogue drains TMEM only, 2 = producer skips TMA after the first ring fill
};

#ifdef VY_GEMM_TRACE
static __device__ long long vy_trace[10 * 16 * 4];  // [warp][tile][field] clock64 stamps of CTA 0
#define VY_TRACE(w, t, f) do { if (blockIdx.x == 0 && (t) < 16 && (threadIdx.x & 31) == 0) vy_trace[((w) * 16 + (t)) * 4 + (f)] = clock64(); } while (0)
#else
#define VY_TRACE(w, t, f) do { } while (0)
#endif

constexpr int GEMM_EPI_WARPS = 8;                        // two warps per TMEM lane quarter, each takes half of the tile's columns
constexpr int GEMM_FIRST_EPI_WARP = 2;                   // warp 0: TMEM allocator + TMA producer, warp 1: MMA issuer
constexpr int GEMM_THREADS = (GEMM_FIRST_EPI_WARP + GEMM_EPI_WARPS) * 32;  // 320 threads -> up to 200 registers each
constexpr int GEMM_STAGE_OUT = 2048;                     // per-warp staging of one [32 rows x 32 bf16] chunk (x2: out and aux)

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
  static constexpr int STAGES = BN >= 256 ? 4 : (BN >= 192 ? 4 : (BN >= 128 ? 6 : 8));
  static constexpr int TMEM_COLS = (2 * BN <= 64) ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512));
  static constexpr int MN_BOX_BYTES = BK * 128;
  static constexpr int EPI_STAGE_BYTES = GEMM_EPI_WARPS * 2 * GEMM_STAGE_OUT;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 2 * BN * 4 /*bias*/ + 256;
  // CTA pair (cta_group::2): each CTA stages its 128 rows of A and HALF of the B tile
  static constexpr int PAIR_B_BYTES = B_BYTES / 2;
  static constexpr int PAIR_STAGE_BYTES = A_BYTES + PAIR_B_BYTES;
  static constexpr int PAIR_STAGES = BN >= 256 ? 6 : (BN >= 192 ? 6 : 8);
  static constexpr int PAIR_SMEM_BYTES = PAIR_STAGES * PAIR_STAGE_BYTES + EPI_STAGE_BYTES + 2 * BN * 4 + 256;
  static constexpr int FMT = sizeof(TIn) == 2 ? 1 : 2;  // bf16 : tf32
  // MN-major operands: bf16 uses the plain 128B swizzle (8 k-rows per 1024-B group); tf32 must use
  // SWIZZLE_128B_BASE32B (4 k-rows per 512-B group) and the matching TMA mode.
  static constexpr int MN_SBO = sizeof(TIn) == 2 ? 1024 : 512;
  static constexpr int MN_LAYOUT = sizeof(TIn) == 2 ? 2 : 1;
  static constexpr int MN_TMA_SWIZZLE = sizeof(TIn) == 2 ? 1 : 2;
  static_assert(SMEM_BYTES <= 232448, "tile does not fit the 227 KB of shared memory");
  static_assert(PAIR_SMEM_BYTES <= 232448, "pair tile does not fit the 227 KB of shared memory");
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

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8_bf16(const uint4& raw, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __bfloat1622float2(h[k]);
    v[2 * k] = f.x;
    v[2 * k + 1] = f.y;
  }
}

// every epilogue warp releases its share of the accumulator buffer once its last tcgen05.ld retired
// (the barrier is named by a shared::cluster address: in a CTA pair it lives in the leader's shared memory)
__device__ __forceinline__ void release_acc(uint32_t tmem_empty_bar, int lane) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(tmem_empty_bar);
}

// --------------------------------------------------------------------------------------------
// Coalesced write-back of one [32 rows x 32 cols] bf16 chunk. Thread `lane` holds row `lane` as 16
// packed words; the chunk goes through this warp's 2 KB staging buffer (16-byte slots XOR-swizzled so
// that both the row-wise writes and the 8-rows-per-instruction read-back are bank-conflict free) and
// leaves as 64-byte row segments: 4 lanes per row, 8 rows per store instruction, every 32-byte sector
// written whole. wb[i] = destination of row (i * 8 + lane / 4) at this lane's 8-column slot (nullptr =
// skip the row).
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_rows_bf16(uint8_t* stage, int lane, const uint32_t (&pk)[16]) {
  uint8_t* rowp = stage + lane * 64;
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
}
__device__ __forceinline__ uint4 staged_slot(const uint8_t* stage, int lane, int i) {
  const int r = i * 8 + (lane >> 2);
  return *reinterpret_cast<const uint4*>(stage + r * 64 + (((lane & 3) ^ ((r >> 1) & 3)) << 4));
}

// --------------------------------------------------------------------------------------------
// Epilogue bodies. Warp (q = TMEM lane quarter, half) owns accumulator rows [32q, 32q + 32) and, of
// the tile's BN columns, the half [half * BN/2, (half + 1) * BN/2); thread <-> accumulator row.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ float gate_fn(float x, int gate_act) { return gate_act ? gelu_tanh(x) : x / (1.f + __expf(-x)); }

// general per-thread handling of one 32-column chunk: any dtype mix / alignment, ragged N edge
static __device__ __noinline__ void epilogue_chunk_general(const GemmDev& g, const uint32_t* raw, const float* bs, int grow,
                                                    long long orow, long long arow, int gcol0, int lim) {
  const float scale = g.out_scale == 0.f ? 1.f : g.out_scale;
  const bool fwd_act = g.act == VY_ACT_GELU_ERF || g.act == VY_ACT_GELU_TANH;
  const bool bwd_act = g.act == VY_ACT_DGELU_ERF || g.act == VY_ACT_DGELU_TANH;
  if (g.act == VY_ACT_SWIGLU) {
    // columns (2j, 2j+1) of the accumulator are (gate_j, up_j): 32 accumulator columns -> 16 outputs at column gcol0 / 2
    const long long obase = orow * g.ld_out + (gcol0 >> 1);
    if (g.vec_ok && lim == 32) {
#pragma unroll 1
      for (int j4 = 0; j4 < 2; ++j4) {
        float x[16], h[8];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(raw[j4 * 16 + j]) + bs[j4 * 16 + j];
        if (g.aux) {
          float lo[8], hi[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { lo[j] = x[j]; hi[j] = x[8 + j]; }
          st8_from_float(g.aux, g.aux_dtype, static_cast<long long>(grow) * g.ld_aux + gcol0 + j4 * 16, lo);
          st8_from_float(g.aux, g.aux_dtype, static_cast<long long>(grow) * g.ld_aux + gcol0 + j4 * 16 + 8, hi);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = gate_fn(x[2 * j], g.gate_act) * x[2 * j + 1] * scale;
        st8_from_float(g.out, g.out_dtype, obase + j4 * 8, h);
      }
    } else {
      for (int j = 0; j + 1 < lim; j += 2) {
        const float a = __uint_as_float(raw[j]) + bs[j], b = __uint_as_float(raw[j + 1]) + bs[j + 1];
        if (g.aux) {
          st_from_float(g.aux, g.aux_dtype, static_cast<long long>(grow) * g.ld_aux + gcol0 + j, a);
          st_from_float(g.aux, g.aux_dtype, static_cast<long long>(grow) * g.ld_aux + gcol0 + j + 1, b);
        }
        st_from_float(g.out, g.out_dtype, obase + (j >> 1), gate_fn(a, g.gate_act) * b * scale);
      }
    }
    return;
  }
  if (g.vec_ok && lim == 32) {
#pragma unroll 1
    for (int j4 = 0; j4 < 4; ++j4) {
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(raw[j4 * 8 + j]) + bs[j4 * 8 + j];
      const int col = gcol0 + j4 * 8;
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
    for (int j = 0; j < lim; ++j) {
      const int col = gcol0 + j;
      float x = __uint_as_float(raw[j]) + bs[j];
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

// general epilogue: any dtype mix / alignment / activation, one 32-column chunk at a time
template <int BN>
__device__ __forceinline__ void epilogue_linear_general(const GemmDev& g, uint32_t tmem_acc, int m0, int n0, int q, int half,
                                                        int lane, const float* bias_s, uint64_t* tfull_bar,
                                                        uint32_t tfull_phase, uint32_t tmem_empty_bar) {
  constexpr int WC = BN >= 64 ? BN / 2 : BN;  // columns per warp
  constexpr int NCH = WC / 32;
  if (BN < 64 && half) {  // narrow tiles: one warp per lane quarter does all the columns
    // an idle warp still arrives once per tile, and only once the tile exists: arriving unconditionally would let it run
    // ahead and complete a LATER phase of the barrier together with the working warps' arrivals for this one
    mbar_wait_soft(tfull_bar, tfull_phase, g.poison);
    release_acc(tmem_empty_bar, lane);
    return;
  }
  const int wcol0 = BN >= 64 ? half * WC : 0;
  const int grow = m0 + q * 32 + lane;
  const bool row_ok = grow < g.M;
  const long long orow = remap_out_row(g, grow);
  const long long arow = remap_add_row(g, grow);
  const int gc0 = n0 + wcol0;
  mbar_wait_soft(tfull_bar, tfull_phase, g.poison);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    uint32_t raw[32];
    tmem_ld_x32(tmem_acc + wcol0 + c * 32, raw);
    tmem_ld_wait();
    if (c == NCH - 1) release_acc(tmem_empty_bar, lane);
    const int nvalid = g.N - (gc0 + c * 32);
    if (row_ok && nvalid > 0)
      epilogue_chunk_general(g, raw, bias_s + wcol0 + c * 32, grow, orow, arow, gc0 + c * 32, nvalid < 32 ? nvalid : 32);
  }
}

// Fast epilogues (bf16 output / aux / addend, 16-byte aligned rows), one instantiation per fused
// operation so that the per-element code is branch-free and small enough to stay in the instruction cache:
enum { EPI_PLAIN = 0, EPI_ADD = 1, EPI_GELU = 2, EPI_DGELU = 3 };
//   EPI_PLAIN  out = scale * (acc + bias)
//   EPI_ADD    out = scale * (acc + bias + addend)                   (residual add / gradient accumulation)
//   EPI_GELU   aux = acc + bias (if aux);  out = scale * gelu_erf(acc + bias)
//   EPI_DGELU  out = scale * (acc + bias) * gelu_erf'(aux)
// Software pipeline per warp: the tcgen05.ld of chunk c+1 is in flight while chunk c is processed; the
// row operand (addend / aux) of the first two chunks is fetched before the accumulator is awaited and
// each later chunk's while its predecessor is processed; the two staging buffers alternate so one
// __syncwarp per chunk suffices. (With EPI_GELU + aux both buffers are used by every chunk.)
__device__ __forceinline__ void store_ragged8(__nv_bfloat16* dst, const uint4& v, int n) {
  const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < n) dst[k] = e[k];
}

// bias[col .. col + 8) as floats, straight from global memory: every lane reads the same 16 / 32 bytes (one L1
// broadcast), so the fast epilogues need neither a staged copy of the tile's bias nor the CTA-wide barrier behind it
__device__ __forceinline__ void load_bias8(const GemmDev& g, int col, float (&bb)[8]) {
  if (!g.bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) bb[j] = 0.f;
  } else if (col + 8 <= g.N) {
    if (g.bias_dtype == VY_BF16) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(g.bias) + col));
      unpack8_bf16(raw, bb);
    } else {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(g.bias) + col));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(g.bias) + col + 4));
      bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w;
      bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) bb[j] = col + j < g.N ? ld_as_float(g.bias, g.bias_dtype, col + j) : 0.f;
  }
}

template <int BN, int MODE, bool TMA>
__device__ __forceinline__ void epilogue_linear_fast(const GemmDev& g, uint32_t tmem_acc, int m0, int n0, int q, int half,
                                                     int lane, const float* bias_s, uint64_t* tfull_bar, uint32_t tfull_phase,
                                                     uint32_t tmem_empty_bar, uint8_t* stage, int trace_tile,
                                                     const CUtensorMap* tma_out, const CUtensorMap* tma_aux) {
  constexpr int WC = BN >= 64 ? BN / 2 : BN;
  constexpr int NCH = WC / 32;
  if (BN < 64 && half) {  // narrow tiles: one warp per lane quarter does all the columns
    // an idle warp still arrives once per tile, and only once the tile exists: arriving unconditionally would let it run
    // ahead and complete a LATER phase of the barrier together with the working warps' arrivals for this one
    mbar_wait_soft(tfull_bar, tfull_phase, g.poison);
    release_acc(tmem_empty_bar, lane);
    return;
  }
  const int wcol0 = BN >= 64 ? half * WC : 0;
  const int grow = m0 + q * 32 + lane;
  const bool row_ok = grow < g.M;
  const int gc0 = n0 + wcol0;  // first global column of this warp
  const float scale = g.out_scale == 0.f ? 1.f : g.out_scale;
  const bool save_aux = MODE == EPI_GELU && g.aux != nullptr;

  const __nv_bfloat16* pf_src = nullptr;
  if (row_ok) {
    if (MODE == EPI_DGELU) pf_src = reinterpret_cast<const __nv_bfloat16*>(g.aux) + static_cast<long long>(grow) * g.ld_aux + gc0;
    if (MODE == EPI_ADD) pf_src = reinterpret_cast<const __nv_bfloat16*>(g.addend) + remap_add_row(g, grow) * g.ld_addend + gc0;
  }
  uint4 pf[2][4];
  auto fetch = [&](uint4 (&dst)[4], int c) {
    if (MODE == EPI_DGELU || MODE == EPI_ADD) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dst[j] = make_uint4(0u, 0u, 0u, 0u);
        if (pf_src && c < NCH && gc0 + c * 32 + j * 8 + 8 <= g.N) dst[j] = *reinterpret_cast<const uint4*>(pf_src + c * 32 + j * 8);
      }
    }
  };
  fetch(pf[0], 0);
  fetch(pf[1], 1);

  // rows this lane writes in the coalesced phase (4 lanes per row, 8 rows per store instruction)
  __nv_bfloat16* wb_out[4];
  long long wb_aux_delta[4];  // aux row address relative to the out row address (elements)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (TMA) {  // the bulk tensor store needs no per-row pointers
      wb_out[i] = nullptr;
      wb_aux_delta[i] = 0;
      continue;
    }
    const int r = m0 + q * 32 + i * 8 + (lane >> 2);
    wb_out[i] = r < g.M ? reinterpret_cast<__nv_bfloat16*>(g.out) + remap_out_row(g, r) * g.ld_out + gc0 + (lane & 3) * 8 : nullptr;
    wb_aux_delta[i] = 0;
    if (save_aux && r < g.M)
      wb_aux_delta[i] = (reinterpret_cast<__nv_bfloat16*>(g.aux) + static_cast<long long>(r) * g.ld_aux + gc0 + (lane & 3) * 8) - wb_out[i];
  }
  const int sw = (lane >> 1) & 3;
  // bias of chunk c (32 columns = one or two cache lines) is pulled into L1 one chunk ahead: the broadcast loads in
  // process() then hit L1 instead of stalling every warp on L2 at the top of every chunk
  const int bias_es = g.bias_dtype == VY_BF16 ? 2 : 4;
  auto prefetch_bias = [&](int c) {
    if (g.bias && c < NCH && lane < 2) {
      const int col = gc0 + c * 32 + lane * 16;
      if (col < g.N) prefetch_l1(reinterpret_cast<const uint8_t*>(g.bias) + static_cast<long long>(col) * bias_es);
    }
  };
  prefetch_bias(0);

  mbar_wait_soft(tfull_bar, tfull_phase, g.poison);
  tc_fence_after();
  VY_TRACE(2 + half * 4 + ((q + 2) & 3), trace_tile, 1);

  auto process = [&](int c, const uint32_t (&raw)[32], uint4 (&pfc)[4], uint8_t* stg) {
    uint4 pc[4];
    if (MODE == EPI_DGELU || MODE == EPI_ADD) {
#pragma unroll
      for (int j = 0; j < 4; ++j) pc[j] = pfc[j];
      fetch(pfc, c + 2);
    }
    prefetch_bias(c + 1);
    const int nvalid = g.N - (gc0 + c * 32);
    if (nvalid <= 0) return;
    uint8_t* srow = stg + lane * 64;
    if (TMA) {
      // the buffer about to be overwritten must have been read by its bulk store: the out buffers alternate per chunk,
      // so only the store before the last has to be done — except with GELU + aux, where both buffers are used every chunk
      if (elect_one_sync()) {  // the lane that issued (and committed) the bulk stores: the same one is elected every time
        if (save_aux) tma_store_wait_read<0>();
        else tma_store_wait_read<1>();
      }
      __syncwarp();
    }
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      float x[8], bb[8];
      load_bias8(g, gc0 + c * 32 + j4 * 8, bb);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(raw[j4 * 8 + j]) + bb[j];
      if (MODE == EPI_GELU) {
        if (save_aux)
          *reinterpret_cast<uint4*>(stage + GEMM_STAGE_OUT + lane * 64 + ((j4 ^ sw) << 4)) =
              make_uint4(pack2_bf16(x[0], x[1]), pack2_bf16(x[2], x[3]), pack2_bf16(x[4], x[5]), pack2_bf16(x[6], x[7]));
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = gelu_erf(x[j]);
      } else if (MODE == EPI_DGELU) {
        float z[8];
        unpack8_bf16(pc[j4], z);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] *= dgelu_erf(z[j]);
      } else if (MODE == EPI_ADD) {
        float a[8];
        unpack8_bf16(pc[j4], a);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += a[j];
      }
      if (scale != 1.f) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] *= scale;
      }
      *reinterpret_cast<uint4*>(srow + ((j4 ^ sw) << 4)) =
          make_uint4(pack2_bf16(x[0], x[1]), pack2_bf16(x[2], x[3]), pack2_bf16(x[4], x[5]), pack2_bf16(x[6], x[7]));
    }
    if (TMA) {
      // [32 rows x 32 cols] bf16 box, SWIZZLE_64B == the staging swizzle; rows >= M / columns >= N are clipped by TMA
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one_sync()) {
        tma_store_2d(tma_out, stg, gc0 + c * 32, m0 + q * 32);
        if (save_aux) tma_store_2d(tma_aux, stage + GEMM_STAGE_OUT, gc0 + c * 32, m0 + q * 32);
        tma_store_commit();
      }
      return;
    }
    __syncwarp();
    if (nvalid >= 32) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (wb_out[i]) {
          *reinterpret_cast<uint4*>(wb_out[i] + c * 32) = staged_slot(stg, lane, i);
          if (save_aux) *reinterpret_cast<uint4*>(wb_out[i] + wb_aux_delta[i] + c * 32) = staged_slot(stage + GEMM_STAGE_OUT, lane, i);
        }
      }
    } else {  // the chunk crosses N: element-wise predicated stores
      const int n = nvalid - (lane & 3) * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (wb_out[i]) {
          store_ragged8(wb_out[i] + c * 32, staged_slot(stg, lane, i), n);
          if (save_aux) store_ragged8(wb_out[i] + wb_aux_delta[i] + c * 32, staged_slot(stage + GEMM_STAGE_OUT, lane, i), n);
        }
      }
    }
    if (save_aux) __syncwarp();  // the aux buffer is reused by the very next chunk
  };

  const uint32_t taddr = tmem_acc + wcol0;
  const int trace_w = 2 + half * 4 + ((q + 2) & 3);
  (void)trace_w;
  if (MODE == EPI_PLAIN || MODE == EPI_GELU) {
    // no row operand to rotate: one rolled loop body per chunk keeps the hot code small (instruction cache); the
    // tcgen05.ld latency is covered by the other epilogue warp of this SM sub-partition
    uint32_t raw[32];
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      tmem_ld_x32(taddr + c * 32, raw);
      tmem_ld_wait();
      if (c == NCH - 1) {
        release_acc(tmem_empty_bar, lane);
        VY_TRACE(trace_w, trace_tile, 2);
      }
      process(c, raw, pf[0], (save_aux || !(c & 1)) ? stage : stage + GEMM_STAGE_OUT);
    }
  } else {
    // the row operand rotates through two register sets (static indices => the chunk loop is unrolled by two); the
    // accumulator chunk itself needs no double buffering: tcgen05.ld + wait measured ~75 cycles
    uint32_t raw[32];
#pragma unroll 1
    for (int c = 0; c + 1 < NCH; c += 2) {
      tmem_ld_x32(taddr + c * 32, raw);
      tmem_ld_wait();
      process(c, raw, pf[0], stage);
      tmem_ld_x32(taddr + (c + 1) * 32, raw);
      tmem_ld_wait();
      if (c + 2 >= NCH) {
        release_acc(tmem_empty_bar, lane);
        VY_TRACE(trace_w, trace_tile, 2);
      }
      process(c + 1, raw, pf[1], stage + GEMM_STAGE_OUT);
    }
    if (NCH & 1) {  // odd chunk count: the last one is alone
      tmem_ld_x32(taddr + (NCH - 1) * 32, raw);
      tmem_ld_wait();
      release_acc(tmem_empty_bar, lane);
      VY_TRACE(trace_w, trace_tile, 2);
      process(NCH - 1, raw, pf[0], stage);
    }
  }
  __syncwarp();  // the next tile's first chunk reuses stage buffer 0
}


// SwiGLU, inference form (no pre-activation save): accumulator columns (2j, 2j+1) = (gate_j, up_j), so a PAIR of 32-column
// chunks yields one [32 rows x 32 cols] bf16 output chunk, staged and written back like the other fast epilogues
// (tma_out describes the N / 2 - column output). BN must be a multiple of 128 (an even number of chunks per warp).
template <int BN>
__device__ __forceinline__ void epilogue_swiglu_fast(const GemmDev& g, uint32_t tmem_acc, int m0, int n0, int q, int half, int lane,
                                                     uint64_t* tfull_bar, uint32_t tfull_phase, uint32_t tmem_empty_bar,
                                                     uint8_t* stage, const CUtensorMap* tma_out) {
  static_assert(BN % 128 == 0, "SwiGLU fast epilogue: an even number of 32-column chunks per warp");
  constexpr int WC = BN / 2;
  constexpr int NCH = WC / 32;
  const int wcol0 = half * WC;
  const int gc0 = n0 + wcol0;
  const float scale = g.out_scale == 0.f ? 1.f : g.out_scale;
  const int sw = (lane >> 1) & 3;
  mbar_wait_soft(tfull_bar, tfull_phase, g.poison);
  tc_fence_after();
  const uint32_t taddr = tmem_acc + wcol0;
#pragma unroll 1
  for (int c = 0; c < NCH; c += 2) {
    const bool live = gc0 + c * 32 < g.N;
    uint8_t* stg = (c & 2) ? stage + GEMM_STAGE_OUT : stage;
    uint8_t* srow = stg + lane * 64;
    if (live) {
      if (elect_one_sync()) tma_store_wait_read<1>();  // the bulk store that last read this buffer (two pairs ago) is done
      __syncwarp();
    }
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {  // one 32-column chunk at a time: 16 outputs = half of the staged row
      uint32_t raw[32];
      tmem_ld_x32(taddr + (c + cc) * 32, raw);
      tmem_ld_wait();
      if (c + 2 >= NCH && cc == 1) release_acc(tmem_empty_bar, lane);
      if (!live) continue;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {  // 16 accumulator columns -> 8 outputs
        float b0[8], b1[8], h[8];
        const int col = gc0 + (c + cc) * 32 + jj * 16;
        load_bias8(g, col, b0);
        load_bias8(g, col + 8, b1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gte = __uint_as_float(raw[jj * 16 + 2 * j]) + (j < 4 ? b0[2 * j] : b1[2 * j - 8]);
          const float up = __uint_as_float(raw[jj * 16 + 2 * j + 1]) + (j < 4 ? b0[2 * j + 1] : b1[2 * j - 7]);
          h[j] = gate_fn(gte, g.gate_act) * up * scale;
        }
        *reinterpret_cast<uint4*>(srow + (((cc * 2 + jj) ^ sw) << 4)) =
            make_uint4(pack2_bf16(h[0], h[1]), pack2_bf16(h[2], h[3]), pack2_bf16(h[4], h[5]), pack2_bf16(h[6], h[7]));
      }
    }
    if (!live) continue;
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one_sync()) {
      tma_store_2d(tma_out, stg, (gc0 + c * 32) >> 1, m0 + q * 32);
      tma_store_commit();
    }
  }
  __syncwarp();
}

// split-K epilogue: the raw fp32 accumulators of this (tile, split) unit go to the workspace slab of the split;
// vy_gemm's reduce kernel sums the slabs and applies bias / addend / scale.
template <int BN>
__device__ __forceinline__ void epilogue_splitk(const GemmDev& g, uint32_t tmem_acc, int m0, int n0, int q, int half, int lane,
                                                int split, uint64_t* tfull_bar, uint32_t tfull_phase, uint32_t tmem_empty_bar) {
  constexpr int WC = BN >= 64 ? BN / 2 : BN;
  constexpr int NCH = WC / 32;
  if (BN < 64 && half) {  // narrow tiles: one warp per lane quarter does all the columns
    // an idle warp still arrives once per tile, and only once the tile exists: arriving unconditionally would let it run
    // ahead and complete a LATER phase of the barrier together with the working warps' arrivals for this one
    mbar_wait_soft(tfull_bar, tfull_phase, g.poison);
    release_acc(tmem_empty_bar, lane);
    return;
  }
  const int wcol0 = BN >= 64 ? half * WC : 0;
  const int grow = m0 + q * 32 + lane;
  float* dst = g.ws + (static_cast<long long>(split) * g.M + grow) * g.N + n0 + wcol0;
  mbar_wait_soft(tfull_bar, tfull_phase, g.poison);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    uint32_t raw[32];
    tmem_ld_x32(tmem_acc + wcol0 + c * 32, raw);
    tmem_ld_wait();
    if (c == NCH - 1) release_acc(tmem_empty_bar, lane);
    if (grow < g.M) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (n0 + wcol0 + c * 32 + j * 4 + 4 <= g.N)  // N % 8 == 0 on this path
          *reinterpret_cast<uint4*>(dst + c * 32 + j * 4) = make_uint4(raw[j * 4], raw[j * 4 + 1], raw[j * 4 + 2], raw[j * 4 + 3]);
    }
  }
}

// swap-AB epilogue: accumulator row = logical output COLUMN (a weight row), accumulator column = logical output ROW
// (a token). Decode-shaped GEMMs run one tile per CTA, so every instruction of the epilogue is executed once and
// fetched cold: the code is kept SMALL instead of unrolled — each 32-column chunk is transposed through the
// staging buffer (fp32 [32 tokens][128 + 1 features], conflict-free) and finished by one rolled loop in which
// consecutive threads own consecutive features of a token, so bias / addend / aux / out accesses are coalesced.
template <int BN>
__device__ __forceinline__ void epilogue_transposed(const GemmDev& g, uint32_t tmem_base_acc, int m0, int n0, int q, int half,
                                                    int lane, int et, uint64_t* tfull_bar, uint32_t tfull_phase,
                                                    uint32_t tmem_empty_bar, float* stage_f) {
  constexpr int NCH = BN >= 32 ? BN / 32 : 1;
  constexpr int LDT = 128 + 1;
  const float scale = g.out_scale == 0.f ? 1.f : g.out_scale;
  const bool fwd_act = g.act == VY_ACT_GELU_ERF || g.act == VY_ACT_GELU_TANH;
  const bool bwd_act = g.act == VY_ACT_DGELU_ERF || g.act == VY_ACT_DGELU_TANH;
  mbar_wait_soft(tfull_bar, tfull_phase, g.poison);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    if (half == 0) {  // one warp per lane quarter moves the chunk; all eight finish it
      uint32_t raw[32];
      tmem_ld_x32(tmem_base_acc + (static_cast<uint32_t>(q * 32) << 16) + c * 32, raw);
      tmem_ld_wait();
      float* col0 = stage_f + q * 32 + lane;
#pragma unroll
      for (int j = 0; j < 32; ++j) col0[j * LDT] = __uint_as_float(raw[j]);
    }
    if (c == NCH - 1) release_acc(tmem_empty_bar, lane);
    named_bar_sync(2, GEMM_EPI_WARPS * 32);
#pragma unroll 1
    for (int idx = et; idx < 32 * 128; idx += GEMM_EPI_WARPS * 32) {
      const int j = idx >> 7, r = idx & 127;
      const int lc = m0 + r;             // logical column (feature)
      const int lr = n0 + c * 32 + j;    // logical row (token)
      if (lc >= g.M || lr >= g.N) continue;
      float x = stage_f[j * LDT + r] + (g.bias ? ld_as_float(g.bias, g.bias_dtype, lc) : 0.f);
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
    named_bar_sync(2, GEMM_EPI_WARPS * 32);  // the staging buffer is reused by the next chunk / tile
  }
}

// QKV projection epilogue: bias + in-register half-split RoPE + head-split scatter (+ kv-cache
// append through k_out/v_out strides). head_dim == 64: one 64-column group is one head. Of every
// head the warp with half = 0 rotates the pairs j in [0,16) (columns j and j + 32), the other warp
// the pairs j in [16,32) — so both warps stay busy for any number of heads per tile and a thread
// has both members of each pair. bf16 destinations leave through the staging buffer as full
// 32-byte sectors; fp32 destinations (an fp32 kv-cache) are written directly.
template <int BN>
__device__ __forceinline__ void epilogue_qkv_rope(const GemmDev& g, uint32_t tmem_acc, int m0, int n0, int q, int half,
                                                  int lane, const float* bias_s, uint64_t* tfull_bar, uint32_t tfull_phase,
                                                  uint32_t tmem_empty_bar, uint8_t* stage) {
  const int grow = m0 + q * 32 + lane;
  const bool row_ok = grow < g.M;
  const int b = row_ok ? grow / g.tokens_per_seq : 0;
  const int l = row_ok ? grow % g.tokens_per_seq : 0;
  const int j0 = half * 16;
  float cs[16], sn[16];
  if (g.rope_cos) {
    const float* cp = g.rope_cos + static_cast<long long>(g.start_pos + l) * 32 + j0;
    const float* sp = g.rope_sin + static_cast<long long>(g.start_pos + l) * 32 + j0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 c4 = *reinterpret_cast<const float4*>(cp + j * 4);
      const float4 s4 = *reinterpret_cast<const float4*>(sp + j * 4);
      cs[j * 4] = c4.x; cs[j * 4 + 1] = c4.y; cs[j * 4 + 2] = c4.z; cs[j * 4 + 3] = c4.w;
      sn[j * 4] = s4.x; sn[j * 4 + 1] = s4.y; sn[j * 4 + 2] = s4.z; sn[j * 4 + 3] = s4.w;
    }
  }
  // (batch, token) of the rows this lane writes in the coalesced phase
  int wb_b[4], wb_l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + q * 32 + i * 8 + (lane >> 2);
    wb_b[i] = r < g.M ? r / g.tokens_per_seq : -1;
    wb_l[i] = r < g.M ? r % g.tokens_per_seq : 0;
  }
  // this lane's 16-byte slot of a staged row: slots 0,1 = first-half columns j0.., slots 2,3 = columns 32 + j0..
  const int slot_col = ((lane & 2) ? 32 : 0) + j0 + (lane & 1) * 8;

  mbar_wait_soft(tfull_bar, tfull_phase, g.poison);
  tc_fence_after();

#pragma unroll 1
  for (int hgrp = 0; hgrp < BN / 64; ++hgrp) {
    uint32_t lo[16], hi[16];
    tmem_ld_x16(tmem_acc + hgrp * 64 + j0, lo);
    tmem_ld_x16(tmem_acc + hgrp * 64 + 32 + j0, hi);
    tmem_ld_wait();
    if (hgrp == BN / 64 - 1) release_acc(tmem_empty_bar, lane);
    const int gcol0 = n0 + hgrp * 64;
    if (gcol0 >= g.N) continue;
    const int head = gcol0 >> 6;
    uint8_t* dst;
    long long sb, sh, sl;
    int hh, tok0, dst_dt;
    bool rotate;
    if (head < g.n_q_heads) {
      dst = reinterpret_cast<uint8_t*>(g.q_out); sb = g.q_sb; sh = g.q_sh; sl = g.q_sl;
      hh = head; tok0 = 0; dst_dt = g.out_dtype; rotate = g.rope_cos != nullptr;
    } else if (head < g.n_q_heads + g.n_kv_heads) {
      dst = reinterpret_cast<uint8_t*>(g.k_out); sb = g.k_sb; sh = g.k_sh; sl = g.k_sl;
      hh = head - g.n_q_heads; tok0 = g.kv_dst_pos0; dst_dt = g.kv_out_dtype; rotate = g.rope_cos != nullptr;
    } else {
      dst = reinterpret_cast<uint8_t*>(g.v_out); sb = g.v_sb; sh = g.v_sh; sl = g.v_sl;
      hh = head - g.n_q_heads - g.n_kv_heads; tok0 = g.kv_dst_pos0; dst_dt = g.kv_out_dtype; rotate = false;
    }
    float o1[16], o2[16];
    const float* bs = bias_s + hgrp * 64 + j0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float x1 = __uint_as_float(lo[j]) + bs[j];
      const float x2 = __uint_as_float(hi[j]) + bs[32 + j];
      if (rotate) {
        o1[j] = x1 * cs[j] - x2 * sn[j];
        o2[j] = x2 * cs[j] + x1 * sn[j];
      } else {
        o1[j] = x1;
        o2[j] = x2;
      }
    }
    if (dst_dt == VY_BF16) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        pk[j] = pack2_bf16(o1[2 * j], o1[2 * j + 1]);
        pk[8 + j] = pack2_bf16(o2[2 * j], o2[2 * j + 1]);
      }
      stage_rows_bf16(stage, lane, pk);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (wb_b[i] >= 0) {
          __nv_bfloat16* rowp = reinterpret_cast<__nv_bfloat16*>(dst) + wb_b[i] * sb + hh * sh +
                                static_cast<long long>(tok0 + wb_l[i]) * sl;
          *reinterpret_cast<uint4*>(rowp + slot_col) = staged_slot(stage, lane, i);
        }
      }
      __syncwarp();
    } else if (row_ok) {
      const long long off = b * sb + hh * sh + static_cast<long long>(tok0 + l) * sl;
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        float a[8], c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a[j] = o1[h8 * 8 + j];
          c[j] = o2[h8 * 8 + j];
        }
        st8_from_float(dst, dst_dt, off + j0 + h8 * 8, a);
        st8_from_float(dst, dst_dt, off + 32 + j0 + h8 * 8, c);
      }
    }
  }
}

// --------------------------------------------------------------------------------------------
// kernel
// --------------------------------------------------------------------------------------------
// GATED: the kernel's only epilogue is the staged SwiGLU one (epilogue_swiglu_fast). It is a separate instantiation so that
// the register allocation of the other epilogues does not change with it (inlined next to them it cost the split-K /
// plain paths a few spills and ~5 % on the weight-gradient GEMMs).
template <typename TIn, int BN, bool A_MN, bool B_MN, bool PAIR, bool GATED = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_aux,
            const __grid_constant__ GemmDev g) {
  pdl_trigger();
  using Cfg = GemmCfg<TIn, BN>;
  constexpr int BM = Cfg::BM;
  constexpr int BK = Cfg::BK;
  constexpr int STAGES = PAIR ? Cfg::PAIR_STAGES : Cfg::STAGES;
  constexpr int B_BYTES = PAIR ? Cfg::PAIR_B_BYTES : Cfg::B_BYTES;      // this CTA's part of the B tile
  constexpr int STAGE_BYTES = Cfg::A_BYTES + B_BYTES;
  constexpr int BN_CTA = PAIR ? BN / 2 : BN;                            // B rows (output columns) staged by this CTA
  static_assert(!PAIR || !B_MN || BN_CTA % Cfg::EPB == 0, "an MN-major B half must be whole swizzle boxes");
  static_assert(!PAIR || BN % 16 == 0, "cta_group::2 MMAs take N in steps of 16");

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* epi_stage = smem + STAGES * STAGE_BYTES;
  float* bias_s = reinterpret_cast<float*>(epi_stage + Cfg::EPI_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 2 * BN);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  // broadcast through a shuffle so that the compiler knows the warp index (and everything derived from it: role, TMEM
  // lane quarter, staging buffer, store coordinates) is warp-uniform and keeps it in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // the 128B-swizzle atoms need a 1024-byte aligned base

  // Work distribution. Plain launch: CTA b walks units b, b + grid, ... (unit = (tile, k-split); k_splits == 1 unless
  // split-K). PAIR launch (2-CTA clusters, tcgen05 cta_group::2): the two CTAs of a cluster walk the same units, a unit
  // covers the m-tile pair (2p, 2p + 1) of one n-tile as ONE 256 x BN MMA tile. CTA rank r stages A rows of m-tile
  // 2p + r and B rows [r * BN/2, (r + 1) * BN/2) of the n-tile in its own shared memory, the leader (rank 0) issues the
  // M = 256 MMAs, which read both CTAs' shared memory and leave m-tile 2p + r's accumulator in CTA r's TMEM. Per SM
  // and MMA that is 128 + BN/2 operand rows out of shared memory instead of 128 + BN — the shared-memory read rate
  // (128 B/clk) is what holds a single-CTA 128 x BN MMA at ~75% of the tensor pipe's rate.
  //   full[s]    leader's: one arrive.expect_tx by the leader's producer, bytes of BOTH CTAs' loads complete on it
  //   empty[s]   per CTA:  the leader's tcgen05.commit arrives on both (multicast)
  //   tfull[a]   per CTA:  likewise
  //   tempty[a]  leader's: 2 x 8 epilogue warps, the peer's arrive through the cluster address
  const int rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const int worker = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int nworkers = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int m_tiles = (g.M + BM - 1) / BM;
  const int m_slots = PAIR ? (m_tiles + 1) / 2 : m_tiles;
  const int m_mul = PAIR ? 2 : 1;
  const int n_tiles = (g.N + BN - 1) / BN;
  const int num_tiles = m_slots * n_tiles;
  const int num_kb = (g.K + BK - 1) / BK;
  const int num_units = num_tiles * g.k_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (g.tma_store) {
      tma_prefetch_desc(&tma_out);
      tma_prefetch_desc(&tma_aux);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], PAIR ? 2 * GEMM_EPI_WARPS : GEMM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    if constexpr (PAIR) {
      tmem_alloc_pair(tmem_ptr_s, Cfg::TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr_s, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync();  // the peer's barriers must be initialised before anything can arrive on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  pdl_wait();  // everything above is independent of the predecessor grid's output

  if (warp == 0) {
    // ===================== TMA producer (whole warp, one elected lane issues) =====================
    {
      uint32_t it = 0, s = 0, ph = 0;  // k-blocks issued, ring position and its phase bit
      for (int unit = worker; unit < num_units; unit += nworkers) {
        const int tile = unit % num_tiles;
        const int m0 = ((tile / n_tiles) * m_mul + rank) * BM;  // may lie beyond M for the odd m-tile of the last pair: TMA zero-fills
        const int n0 = (tile % n_tiles) * BN + rank * BN_CTA;   // first B row this CTA stages
        const int kb0 = (unit / num_tiles) * g.kb_per_split;
        const int kb1 = min(num_kb, kb0 + g.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it, ph ^= (s + 1 == STAGES), s = (s + 1 == STAGES) ? 0 : s + 1) {
          mbar_wait_soft(&empty_bar[s], ph ^ 1, g.poison);
          if (elect_one_sync()) {
            uint8_t* a_dst = sA + s * Cfg::A_BYTES;
            uint8_t* b_dst = sB + s * B_BYTES;
            if ((g.debug & 2) && it >= STAGES) {
              if (rank == 0) mbar_arrive(&full_bar[s]);
            } else if constexpr (!PAIR) {
              mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
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
            } else {
              // both CTAs' bytes complete on the LEADER's full barrier, which gates the one MMA thread of the pair
              const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[s]), 0);
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
              if constexpr (!A_MN) {
                tma_load_2d_pair(a_dst, &tma_a, full_leader, kb * BK, m0);
              } else {
#pragma unroll
                for (int i = 0; i < BM / Cfg::EPB; ++i)
                  tma_load_2d_pair(a_dst + i * Cfg::MN_BOX_BYTES, &tma_a, full_leader, m0 + i * Cfg::EPB, kb * BK);
              }
              if constexpr (!B_MN) {  // tma_b was built with a box of BN / 2 rows
                tma_load_2d_pair(b_dst, &tma_b, full_leader, kb * BK, n0);
              } else {
#pragma unroll
                for (int i = 0; i < BN_CTA / Cfg::EPB; ++i)
                  tma_load_2d_pair(b_dst + i * Cfg::MN_BOX_BYTES, &tma_b, full_leader, n0 + i * Cfg::EPB, kb * BK);
              }
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, one elected lane issues; in a pair only the leader's) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(Cfg::FMT, PAIR ? 2 * BM : BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // The issue loop is kept to a handful of instructions per MMA: it shares its SM sub-partition's issue slots with
      // two epilogue warps, and at ~95 instructions per k-block (descriptors rebuilt from addresses, % and / for the
      // ring position) the loop itself, not the tensor pipe, set the pace of tiles narrower than 256. The descriptors of
      // stage 0 are built once; a stage adds its byte offset >> 4 to the 14-bit address field (the ring stays inside the
      // field's 256 KB, so nothing carries into the neighbouring fields) and a k-step adds a constant.
      const uint64_t a_desc0 = A_MN ? make_smem_desc_sw128(smem_u32(sA), Cfg::MN_BOX_BYTES, Cfg::MN_SBO, Cfg::MN_LAYOUT)
                                    : make_smem_desc_sw128(smem_u32(sA), 16, 1024);
      const uint64_t b_desc0 = B_MN ? make_smem_desc_sw128(smem_u32(sB), Cfg::MN_BOX_BYTES, Cfg::MN_SBO, Cfg::MN_LAYOUT)
                                    : make_smem_desc_sw128(smem_u32(sB), 16, 1024);
      const uint32_t a_hi = static_cast<uint32_t>(a_desc0 >> 32), b_hi = static_cast<uint32_t>(b_desc0 >> 32);
      const uint32_t a_lo0 = static_cast<uint32_t>(a_desc0), b_lo0 = static_cast<uint32_t>(b_desc0);
      constexpr uint32_t A_KSTEP = (A_MN ? Cfg::UMMA_K * 128 : 32) >> 4;  // one UMMA_K step inside a stage, in 16-byte units
      constexpr uint32_t B_KSTEP = (B_MN ? Cfg::UMMA_K * 128 : 32) >> 4;
      uint32_t s = 0, ph = 0;  // ring position and its phase bit
      uint32_t local = 0;
      for (int unit = worker; unit < num_units; unit += nworkers, ++local) {
        const int kb0 = (unit / num_tiles) * g.kb_per_split;
        const int nkb = min(num_kb, kb0 + g.kb_per_split) - kb0;
        const uint32_t acc = local & 1;
        const uint32_t acc_ph = (local >> 1) & 1;
        VY_TRACE(1, local, 0);
        mbar_wait_soft(&tempty_bar[acc], acc_ph ^ 1, g.poison);
        tc_fence_after();
        VY_TRACE(1, local, 1);
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int i = 0; i < nkb; ++i) {
          mbar_wait_soft(&full_bar[s], ph, g.poison);  // (cta-scope acquire: a cluster-scope one invalidates L1 on every k-block)
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t a_lo = a_lo0 + s * (Cfg::A_BYTES >> 4);
            const uint32_t b_lo = b_lo0 + s * (B_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BK / Cfg::UMMA_K; ++k) {
              const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + k * A_KSTEP);
              const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (b_lo + k * B_KSTEP);
              const uint32_t accum = (i | k) != 0;
              if constexpr (PAIR) {
                if constexpr (sizeof(TIn) == 2) umma_f16_pair(d_tmem, ad, bd, idesc, accum);
                else umma_tf32_pair(d_tmem, ad, bd, idesc, accum);
              } else {
                if constexpr (sizeof(TIn) == 2) umma_f16(d_tmem, ad, bd, idesc, accum);
                else umma_tf32(d_tmem, ad, bd, idesc, accum);
              }
            }
            if constexpr (PAIR) {
              umma_commit_pair(&empty_bar[s], 3);
              if (i == nkb - 1) umma_commit_pair(&tfull_bar[acc], 3);
            } else {
              umma_commit(&empty_bar[s]);
              if (i == nkb - 1) umma_commit(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        VY_TRACE(1, local, 2);
      }
    }
  } else if (warp >= GEMM_FIRST_EPI_WARP) {
    // ===================== epilogue =====================
    const int e = warp - GEMM_FIRST_EPI_WARP;
    const int q = warp & 3;  // warp % 4: the TMEM lane quarter this warp may access
    const int half = e >> 2;
    const int et = threadIdx.x - GEMM_FIRST_EPI_WARP * 32;
    uint8_t* stage = epi_stage + e * 2 * GEMM_STAGE_OUT;
    // which fused fast epilogue applies (-1: the general one)
    int fast_mode = -1;
    if (g.epi == VY_EPI_LINEAR && !g.transposed_out && g.vec_ok && g.out_dtype == VY_BF16 && !g.addend2 &&
        (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0) {
      const bool aux16 = !g.aux || g.aux_dtype == VY_BF16;
      const bool n8 = (g.N & 7) == 0;  // the prefetched row operand is read in 8-column vectors
      if (g.act == VY_ACT_NONE) fast_mode = !g.addend ? EPI_PLAIN : ((g.addend_dtype == VY_BF16 && n8) ? EPI_ADD : -1);
      else if (g.act == VY_ACT_GELU_ERF && !g.addend && aux16) fast_mode = EPI_GELU;
      else if (g.act == VY_ACT_DGELU_ERF && !g.addend && aux16 && n8) fast_mode = EPI_DGELU;
    }
    uint32_t local = 0;
    for (int unit = worker; unit < num_units; unit += nworkers, ++local) {
      const int tile = unit % num_tiles;
      const uint32_t acc = local & 1;
      const uint32_t acc_ph = (local >> 1) & 1;
      const int m0 = ((tile / n_tiles) * m_mul + rank) * BM;
      const int n0 = (tile % n_tiles) * BN;
      float* bs = bias_s + acc * BN;
      // where this warp announces that it has drained the accumulator: the (leader's) MMA thread waits there
      const uint32_t tempty_addr = PAIR ? mapa_u32(smem_u32(&tempty_bar[acc]), 0) : smem_u32(&tempty_bar[acc]);
      VY_TRACE(2 + e, local, 0);
      if (!GATED && !g.transposed_out && g.k_splits <= 1 && fast_mode < 0) {  // the fast epilogues read the bias straight from global
        for (int j = et; j < BN; j += GEMM_EPI_WARPS * 32) {
          const int col = n0 + j;
          bs[j] = (g.bias && col < g.N) ? ld_as_float(g.bias, g.bias_dtype, col) : 0.f;
        }
        named_bar_sync(1, GEMM_EPI_WARPS * 32);
      }
      const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      if constexpr (GATED) {
        epilogue_swiglu_fast<BN>(g, tmem_acc, m0, n0, q, half, lane, &tfull_bar[acc], acc_ph, tempty_addr, stage, &tma_out);
        continue;
      }
      if (g.debug & 1) {
        mbar_wait_soft(&tfull_bar[acc], acc_ph, g.poison);
        tc_fence_after();
        release_acc(tempty_addr, lane);
      } else if (g.k_splits > 1) {
        epilogue_splitk<BN>(g, tmem_acc, m0, n0, q, half, lane, unit / num_tiles, &tfull_bar[acc], acc_ph, tempty_addr);
      } else if (g.epi == VY_EPI_QKV_ROPE) {
        if constexpr (BN >= 64)
          epilogue_qkv_rope<BN>(g, tmem_acc, m0, n0, q, half, lane, bs, &tfull_bar[acc], acc_ph, tempty_addr, stage);
      } else if (g.transposed_out) {
        epilogue_transposed<BN>(g, tmem_base + acc * BN, m0, n0, q, half, lane, et, &tfull_bar[acc], acc_ph, tempty_addr,
                                reinterpret_cast<float*>(epi_stage));
      } else if (fast_mode == EPI_PLAIN) {
        if (g.tma_store)
          epilogue_linear_fast<BN, EPI_PLAIN, true>(g, tmem_acc, m0, n0, q, half, lane, bs, &tfull_bar[acc], acc_ph, tempty_addr, stage,
                                                    local, &tma_out, &tma_aux);
        else
          epilogue_linear_fast<BN, EPI_PLAIN, false>(g, tmem_acc, m0, n0, q, half, lane, bs, &tfull_bar[acc], acc_ph, tempty_addr, stage,
                                                     local, &tma_out, &tma_aux);
      } else if (fast_mode == EPI_ADD) {
        if (g.tma_store)
          epilogue_linear_fast<BN, EPI_ADD, true>(g, tmem_acc, m0, n0, q, half, lane, bs, &tfull_bar[acc], acc_ph, tempty_addr, stage,
                                                    local, &tma_out, &tma_aux);
        else
          epilogue_linear_fast<BN, EPI_ADD, false>(g, tmem_acc, m0, n0, q, half, lane, bs, &tfull_bar[acc], acc_ph, tempty_addr, stage,
                                                     local, &tma_out, &tma_aux);
      } else if (fast_mode == EPI_GELU) {
        if (g.tma_store)
          epilogue_linear_fast<BN, EPI_GELU, true>(g, tmem_acc, m0, n0, q, half, lane, bs, &tfull_bar[acc], acc_ph, tempty_addr, stage,
                                                    local, &tma_out, &tma_aux);
        else
          epilogue_linear_fast<BN, EPI_GELU, false>(g, tmem_acc, m0, n0, q, half, lane, bs, &tfull_bar[acc], acc_ph, tempty_addr, stage,
                                                     local, &tma_out, &tma_aux);
      } else if (fast_mode == EPI_DGELU) {
        if (g.tma_store)
          epilogue_linear_fast<BN, EPI_DGELU, true>(g, tmem_acc, m0, n0, q, half, lane, bs, &tfull_bar[acc], acc_ph, tempty_addr, stage,
                                                    local, &tma_out, &tma_aux);
        else
          epilogue_linear_fast<BN, EPI_DGELU, false>(g, tmem_acc, m0, n0, q, half, lane, bs, &tfull_bar[acc], acc_ph, tempty_addr, stage,
                                                     local, &tma_out, &tma_aux);
      } else {
        epilogue_linear_general<BN>(g, tmem_acc, m0, n0, q, half, lane, bs, &tfull_bar[acc], acc_ph, tempty_addr);
      }
      VY_TRACE(2 + e, local, 3);
    }
    if (g.tma_store && elect_one_sync()) tma_store_wait<0>();  // the staged tiles must have left smem (and landed) before exit
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync();  // the peer may still arrive on this CTA's barriers / read its smem until it is done too
  else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
#ifdef VY_GEMM_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const long long t0 = vy_trace[(1 * 16 + 0) * 4 + 0];
    for (int t = 0; t < 9; ++t) {
      printf("tile %d MMA: start %lld tempty_ok %lld issued %lld\n", t, vy_trace[(16 + t) * 4] - t0, vy_trace[(16 + t) * 4 + 1] - t0, vy_trace[(16 + t) * 4 + 2] - t0);
      for (int w = 2; w < 10; w += 3)
        printf("   epi w%d: top %lld tfull %lld released %lld end %lld\n", w, vy_trace[(w * 16 + t) * 4] - t0, vy_trace[(w * 16 + t) * 4 + 1] - t0,
               vy_trace[(w * 16 + t) * 4 + 2] - t0, vy_trace[(w * 16 + t) * 4 + 3] - t0);
    }
  }
#endif
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
static inline int get_tmap_2d(CUtensorMap* out, int dtype, const void* base, uint64_t d0, uint64_t d1,
                       uint64_t stride1_bytes, uint32_t b0, uint32_t b1, int swz = 1) {
  uint64_t dims[2] = {d0, d1};
  uint64_t strides[2] = {0, stride1_bytes};
  uint32_t box[2] = {b0, b1};
  return get_tensor_map_cached(out, dtype, 2, base, dims, strides, box, swz);
}

template <typename TIn, int BN, bool A_MN, bool B_MN, bool PAIR, bool GATED = false>
int launch_gemm(const VyGemm* p, const GemmDev& g) {
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
  if (!B_MN)  // a CTA of a pair stages half of the B tile
    rc = get_tmap_2d(&tb, dt, p->B, p->K, p->N, p->ldb * es, Cfg::BK, PAIR ? BN / 2 : BN);
  else
    rc = get_tmap_2d(&tb, dt, p->B, p->N, p->K, p->ldb * es, Cfg::EPB, Cfg::BK, Cfg::MN_TMA_SWIZZLE);
  if (rc != VY_OK) return rc;

  // write-back tensor maps of the fast epilogues: [32 cols x 32 rows] bf16 boxes, SWIZZLE_64B (the staging layout)
  CUtensorMap tout = ta, taux = ta;
  GemmDev gl = g;
  if (g.tma_store) {
    rc = get_tmap_2d(&tout, VY_BF16, p->out, p->act == VY_ACT_SWIGLU ? p->N / 2 : p->N, p->M, p->ld_out * 2, 32, 32, 3);
    if (rc == VY_OK && p->aux && (p->act == VY_ACT_GELU_ERF))
      rc = get_tmap_2d(&taux, VY_BF16, p->aux, p->N, p->M, p->ld_aux * 2, 32, 32, 3);
    if (rc != VY_OK) gl.tma_store = 0;  // fall back to the staged st.global write-back
  }
  auto kern = gemm_kernel<TIn, BN, A_MN, B_MN, PAIR, GATED>;
  constexpr int smem_bytes = PAIR ? Cfg::PAIR_SMEM_BYTES : Cfg::SMEM_BYTES;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    VY_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  const int m_tiles = (p->M + Cfg::BM - 1) / Cfg::BM;
  const int n_tiles = (p->N + BN - 1) / BN;
  const int m_slots = PAIR ? (m_tiles + 1) / 2 : m_tiles;
  const int units = m_slots * n_tiles * (g.k_splits > 1 ? g.k_splits : 1);
  // VY_GEMM_SM_MARGIN=n leaves n SMs to concurrently running kernels (the NCCL all-reduce of a data-parallel step): a
  // persistent grid that assumes every SM is its own waits for the stragglers that could not be co-scheduled.
  static const int sm_margin = getenv("VY_GEMM_SM_MARGIN") ? atoi(getenv("VY_GEMM_SM_MARGIN")) : 0;
  const int usable_sms = num_sms() - sm_margin > 16 ? num_sms() - sm_margin : num_sms();
  const int max_workers = PAIR ? usable_sms / 2 : usable_sms;
  const int workers = units < max_workers ? units : max_workers;
  const int grid = PAIR ? 2 * workers : workers;
  VY_CUDA_OK(launch_kernel_cluster(kern, dim3(grid), dim3(GEMM_THREADS), smem_bytes, static_cast<cudaStream_t>(p->stream),
                                   PAIR ? 2 : 1, ta, tb, tout, taux, gl));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

}  // namespace vy
