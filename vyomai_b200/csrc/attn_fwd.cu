// vy_attn_fwd: flash-style fused attention forward on tcgen05 / TMEM, operands fed by TMA.
//
// One CTA per (128-row q tile, q head, batch row); two CTAs co-reside per SM so the tensor pipe of
// one overlaps the softmax of the other. bf16 operands, head_dim 64, fp32 softmax / accumulation.
//   warp 0     TMA producer: Q once, then K_j / V_j (128 keys each) through a 2-stage ring
//   warp 1     MMA issuer:   S = Q K_j^T  (M128 N128 K64, accumulator in TMEM cols [0,128))
//                            O_j = P_j V_j (M128 N64  K128, TMEM cols [128,192); V is MN-major)
//   warp 2     TMEM allocator
//   warps 4-7  softmax: thread <-> query row. Two passes over S in TMEM (max, then exp2 + sum),
//              P written as bf16 into 128B-swizzled smem (the A operand of the second MMA), the
//              running output rescaled in registers (64 fp32 per thread).
// GQA is a head-index map in the TMA coordinate (q head h reads kv head h / n_rep) — nothing is
// repeated in memory. Masks reproduce the reference's additive finfo.min semantics: a masked key
// gets a large FINITE negative score, so a fully masked row degenerates to the uniform mean over
// all keys; keys beyond Skv (TMA zero fill) get -inf and never count.
#include "vy_common.cuh"
#include "vy_ptx.cuh"

namespace vy {

constexpr int AT_BM = 128;   // query rows per CTA
constexpr int AT_BN = 128;   // keys per tile
constexpr int AT_D = 64;
constexpr int AT_TILE = AT_BM * AT_D * 2;        // 16 KB: one [128 x 64] bf16 tile
constexpr int AT_P_BYTES = AT_BM * AT_BN * 2;    // 32 KB
constexpr int AT_SMEM = AT_TILE /*Q*/ + 2 * AT_TILE /*K*/ + 2 * AT_TILE /*V*/ + AT_P_BYTES + 256;
// finite, and small enough that lse = masked + log2(n) keeps ~1e-3 absolute precision in fp32, so a
// fully masked row is the uniform mean in forward AND backward (quirk Q4). exp2(masked - real) == 0.
constexpr float AT_MASKED = -30000.0f;

struct AttnDev {
  int B, Hq, Hkv, Sq, Skv, n_rep;
  int causal, q_pos0;
  const unsigned char* kpm;  // [B, Skv], 1 = key visible; or null
  long long kpm_sb;
  void* out;  // out[b, l, head * 64 + j]
  long long o_sb, o_sl;
  int out_dtype;
  float* lse;  // [B, Hq, Sq] log2-domain logsumexp, or null
  float scale_log2;
};

__device__ __forceinline__ int attn_num_kv_tiles(const AttnDev& g, int b, int q0) {
  const int total = (g.Skv + AT_BN - 1) / AT_BN;
  if (!g.causal) return total;
  // tiles strictly above the diagonal contribute exactly 0 unless a row is fully masked (then the
  // reference averages over ALL keys). With key 0 visible no row can be fully masked.
  if (g.kpm && g.kpm[b * g.kpm_sb] == 0) return total;
  const int last_q = min(q0 + AT_BM, g.Sq) - 1;
  const int need = (g.q_pos0 + last_q) / AT_BN + 1;
  return need < total ? need : total;
}

__global__ void __launch_bounds__(256, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                const __grid_constant__ CUtensorMap tma_v, const AttnDev g) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_TILE;
  uint8_t* sV = smem + 3 * AT_TILE;
  uint8_t* sP = smem + 5 * AT_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * AT_TILE + AT_P_BYTES);
  uint64_t* q_full = bars;          // [1]
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* v_full = bars + 3;      // [2]
  uint64_t* kv_empty = bars + 5;    // [2]
  uint64_t* s_full = bars + 7;      // [1]
  uint64_t* p_full = bars + 8;      // [1]
  uint64_t* o_full = bars + 9;      // [1]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BM;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int kvh = head / g.n_rep;

  if ((smem_u32(smem) & 1023u) != 0) __trap();  // swizzled tiles need a 1024-B aligned base

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_s, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  pdl_wait();  // everything above is independent of the predecessor grid's output
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  const int n_tiles = attn_num_kv_tiles(g, b, q0);

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, AT_TILE);
      tma_load_4d(sQ, &tma_q, q_full, 0, q0, head, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[s], AT_TILE);
        tma_load_4d(sK + s * AT_TILE, &tma_k, &k_full[s], 0, j * AT_BN, kvh, b);
        mbar_arrive_expect_tx(&v_full[s], AT_TILE);
        tma_load_4d(sV + s * AT_TILE, &tma_v, &v_full[s], 0, j * AT_BN, kvh, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(1, AT_BM, AT_BN, 0, 0);
      constexpr uint32_t idesc_o = make_idesc(1, AT_BM, AT_D, 0, 1);
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(&k_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + s * AT_TILE);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_f16(tmem_S, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                   make_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + s * AT_TILE);
#pragma unroll
        for (int kk = 0; kk < AT_BN / 16; ++kk)
          umma_f16(tmem_O,
                   make_smem_desc_sw128(p_addr + (kk >> 2) * (AT_BM * 128) + (kk & 3) * 32, 16, 1024),
                   make_smem_desc_sw128(v_addr + kk * 2048, 8192, 1024), idesc_o, kk != 0);
        umma_commit(o_full);
        umma_commit(&kv_empty[s]);
        if (j + 1 < n_tiles) issue_s(j + 1);
      }
    }
  } else if (warp >= 4) {
    const int qd = warp - 4;  // TMEM lane quarter
    const int row = qd * 32 + lane;
    const int qrow = q0 + row;
    const int qpos = g.q_pos0 + qrow;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    float m = -INFINITY, l = 0.f;
    float acc[AT_D];
#pragma unroll
    for (int i = 0; i < AT_D; ++i) acc[i] = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      const int kv0 = j * AT_BN;
      // key-padding bits of this tile: bit i of kb[k] <-> key kv0 + 4 i + k
      uint32_t kb[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      // Visibility of the tile's keys for this row in "prefix form": keys [0, vis_end) are visible, [vis_end, tile_valid)
      // are masked (finite score: quirk Q4), [tile_valid, 128) lie beyond Skv. Right padding, no padding and the causal
      // bound are all prefixes, so the per-element mask logic (bit extraction, two compares, two selects — it made the
      // softmax warps issue-bound at ~33 instructions per score) collapses to a comparison of the column index with a
      // per-row register; anything else (left padding) takes the general per-element path.
      const int tile_valid = min(AT_BN, g.Skv - kv0);
      int vis_end = tile_valid;
      bool prefix = true;
      if (g.kpm) {
        uint32_t w = 0;
        const int kbase = kv0 + lane * 4;
        const unsigned char* kp = g.kpm + b * g.kpm_sb;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (kbase + k < g.Skv && kp[kbase + k]) w |= 1u << k;
#pragma unroll
        for (int k = 0; k < 4; ++k) kb[k] = __ballot_sync(0xffffffffu, (w >> k) & 1u);
        const int cnt = __popc(kb[0]) + __popc(kb[1]) + __popc(kb[2]) + __popc(kb[3]);
        uint32_t expect = 0;  // this lane's 4 keys if the visible ones were exactly the first cnt of the tile
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (lane * 4 + k < cnt) expect |= 1u << k;
        prefix = __ballot_sync(0xffffffffu, expect != w) == 0u;
        vis_end = min(vis_end, cnt);
      }
      if (g.causal) vis_end = min(vis_end, max(qpos - kv0 + 1, 0));
      // tcgen05.ld is warp-collective: the chunk classification below must be warp-uniform, so it uses the smallest and
      // the largest vis_end of the warp's 32 rows (they differ only on the causal diagonal)
      const int vmin = __reduce_min_sync(0xffffffffu, vis_end), vmax = __reduce_max_sync(0xffffffffu, vis_end);

      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float mx = m;
      if (prefix) {
        // pass 1: row max of the visible raw scores (the scale is positive: applied once to the maximum)
        float rmax = -INFINITY;
#pragma unroll 1
        for (int c4 = 0; c4 < AT_BN / 32; ++c4) {
          if (c4 * 32 >= vmax) break;
          uint32_t raw[32];
          tmem_ld_x32(tmem_S + lane_off + c4 * 32, raw);
          tmem_ld_wait();
          if (c4 * 32 + 32 <= vmin) {
#pragma unroll
            for (int i = 0; i < 32; ++i) rmax = fmaxf(rmax, __uint_as_float(raw[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c4 * 32 + i < vis_end) rmax = fmaxf(rmax, __uint_as_float(raw[i]));
          }
        }
        mx = fmaxf(mx, rmax * g.scale_log2);
        if (vis_end < tile_valid) mx = fmaxf(mx, AT_MASKED);
      } else {
        auto score = [&](uint32_t raw, int c) -> float {
          const int key = kv0 + c;
          float t = __uint_as_float(raw) * g.scale_log2;
          const bool vis = ((kb[c & 3] >> (c >> 2)) & 1u) && (!g.causal || key <= qpos);
          t = vis ? t : AT_MASKED;
          return key < g.Skv ? t : -INFINITY;
        };
#pragma unroll 1
        for (int c4 = 0; c4 < AT_BN / 32; ++c4) {
          uint32_t raw[32];
          tmem_ld_x32(tmem_S + lane_off + c4 * 32, raw);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, score(raw[i], c4 * 32 + i));
        }
      }
      const float alpha = vy_ex2_approx(m - mx);  // m = -inf on the first tile -> 0
      m = mx;
      // fold in the previous tile's P V (already complete: the tensor pipe runs in issue order)
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t raw[32];
          tmem_ld_x32(tmem_O + lane_off + h * 32, raw);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[h * 32 + i] += __uint_as_float(raw[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < AT_D; ++i) acc[i] *= alpha;
      l *= alpha;
      // pass 2: p = exp2(t - m), row sum, bf16 P into swizzled smem (K-major, two 64-key atoms)
      const float pm = vy_ex2_approx(AT_MASKED - m);  // weight of a masked key: 0, or 1 when the row has seen no visible key
#pragma unroll 1
      for (int c4 = 0; c4 < AT_BN / 32; ++c4) {
        uint32_t packed[16];
        if (prefix && c4 * 32 >= vmax) {
          // whole chunk masked or beyond Skv for every row of the warp: constant weights, no TMEM read
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = c4 * 32 + 2 * i < tile_valid ? pm : 0.f;
            const float p1 = c4 * 32 + 2 * i + 1 < tile_valid ? pm : 0.f;
            l += p0 + p1;
            __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
            packed[i] = *reinterpret_cast<uint32_t*>(&h2);
          }
        } else {
          uint32_t raw[32];
          tmem_ld_x32(tmem_S + lane_off + c4 * 32, raw);
          tmem_ld_wait();
          if (prefix && c4 * 32 + 32 <= vmin) {
            // whole chunk visible for every row of the warp: one FFMA + one MUFU per score
            const float nm = -m;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float p0 = vy_ex2_approx(fmaf(__uint_as_float(raw[2 * i]), g.scale_log2, nm));
              const float p1 = vy_ex2_approx(fmaf(__uint_as_float(raw[2 * i + 1]), g.scale_log2, nm));
              l += p0 + p1;
              __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
              packed[i] = *reinterpret_cast<uint32_t*>(&h2);
            }
          } else if (prefix) {
            const float nm = -m;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c0 = c4 * 32 + 2 * i;
              float p0 = vy_ex2_approx(fmaf(__uint_as_float(raw[2 * i]), g.scale_log2, nm));
              float p1 = vy_ex2_approx(fmaf(__uint_as_float(raw[2 * i + 1]), g.scale_log2, nm));
              p0 = c0 < vis_end ? p0 : (c0 < tile_valid ? pm : 0.f);
              p1 = c0 + 1 < vis_end ? p1 : (c0 + 1 < tile_valid ? pm : 0.f);
              l += p0 + p1;
              __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
              packed[i] = *reinterpret_cast<uint32_t*>(&h2);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float p[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int c = c4 * 32 + 2 * i + e;
                const int key = kv0 + c;
                const bool vis = ((kb[c & 3] >> (c >> 2)) & 1u) && (!g.causal || key <= qpos);
                const float t = vis ? __uint_as_float(raw[2 * i + e]) * g.scale_log2 : AT_MASKED;
                p[e] = key < g.Skv ? vy_ex2_approx(t - m) : 0.f;
              }
              l += p[0] + p[1];
              __nv_bfloat162 h2 = __floats2bfloat162_rn(p[0], p[1]);
              packed[i] = *reinterpret_cast<uint32_t*>(&h2);
            }
          }
        }
        // 32 columns = 4 chunks of 16 B; atom = 64 columns
        const int atom = c4 >> 1;
        uint8_t* rowp = sP + atom * (AT_BM * 128) + row * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = ((c4 & 1) * 4 + ch) ^ (row & 7);
          *reinterpret_cast<uint4*>(rowp + chunk * 16) =
              make_uint4(packed[ch * 4], packed[ch * 4 + 1], packed[ch * 4 + 2], packed[ch * 4 + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // last tile's P V
    mbar_wait(o_full, (n_tiles - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t raw[32];
      tmem_ld_x32(tmem_O + lane_off + h * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[h * 32 + i] += __uint_as_float(raw[i]);
    }
    if (qrow < g.Sq) {
      const float inv = 1.f / l;
      const long long off = b * g.o_sb + static_cast<long long>(qrow) * g.o_sl + head * AT_D;
#pragma unroll
      for (int q8 = 0; q8 < AT_D / 8; ++q8) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = acc[q8 * 8 + i] * inv;
        st8_from_float(g.out, g.out_dtype, off + q8 * 8, o);
      }
      if (g.lse) g.lse[(static_cast<long long>(b) * g.Hq + head) * g.Sq + qrow] = m + log2f(l);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static int make_qkv_map(CUtensorMap* out, const void* base, int S, int H, int B, long long sb, long long sh,
                        long long sl) {
  uint64_t dims[4] = {static_cast<uint64_t>(AT_D), static_cast<uint64_t>(S), static_cast<uint64_t>(H),
                      static_cast<uint64_t>(B)};
  uint64_t strides[4] = {0, static_cast<uint64_t>(sl) * 2, static_cast<uint64_t>(sh) * 2,
                         static_cast<uint64_t>(sb) * 2};
  uint32_t box[4] = {AT_D, 128, 1, 1};
  return get_tensor_map_cached(out, VY_BF16, 4, base, dims, strides, box, 1);
}

}  // namespace vy

namespace vy {
int attn_fwd_mma(const VyAttn* p);  // attn_fwd_mma.cu
}

extern "C" int vy_attn_fwd(const VyAttn* p) {
  using namespace vy;
  VY_CHECK_ARG(p != nullptr, "vy_attn_fwd: null params");
  if (!vy_device_ok()) {
    set_error("vy_attn_fwd: no sm_100 device (there is no CPU fallback)");
    return VY_ERR_NO_DEVICE;
  }
  VY_CHECK_ARG(p->qkv_dtype == VY_BF16, "vy_attn_fwd: q/k/v must be bf16 (fp32 models hand bf16 operands to the tensor cores)");
  VY_CHECK_ARG(p->B > 0 && p->Sq > 0 && p->Skv > 0 && p->n_q_heads > 0 && p->n_kv_heads > 0 &&
                   p->n_q_heads % p->n_kv_heads == 0,
               "vy_attn_fwd: bad shape");
  VY_CHECK_ARG(p->q && p->k && p->v && p->out && dtype_ok(p->out_dtype), "vy_attn_fwd: null pointer / bad out dtype");
  if (p->causal) VY_CHECK_ARG(p->q_pos0 >= 0, "vy_attn_fwd: negative q_pos0");
  // head dims other than 64 and the prefix-LM mask: the mma.sync kernel (attn_fwd_mma.cu); VY_ATTN_MMA=1 sends everything there
  static const bool force_mma = getenv("VY_ATTN_MMA") && atoi(getenv("VY_ATTN_MMA")) != 0;
  if (p->head_dim != 64 || (p->causal && p->prefix_len) || p->pos_ptr || force_mma) return attn_fwd_mma(p);
  auto ok16 = [](const void* ptr, long long a, long long b_, long long c) {
    return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (a * 2) % 16 == 0 && (b_ * 2) % 16 == 0 && (c * 2) % 16 == 0;
  };
  VY_CHECK_ARG(ok16(p->q, p->q_sb, p->q_sh, p->q_sl) && ok16(p->k, p->k_sb, p->k_sh, p->k_sl) &&
                   ok16(p->v, p->v_sb, p->v_sh, p->v_sl),
               "vy_attn_fwd: q/k/v pointers and strides must keep 16-byte alignment");
  const long long eso = dtype_size(p->out_dtype);
  VY_CHECK_ARG((reinterpret_cast<uintptr_t>(p->out) & 15) == 0 && (p->o_sb * eso) % 16 == 0 && (p->o_sl * eso) % 16 == 0,
               "vy_attn_fwd: out pointer/strides must keep 16-byte alignment");
  if (p->causal) VY_CHECK_ARG(p->q_pos0 >= 0, "vy_attn_fwd: negative q_pos0");

  CUtensorMap tq, tk, tv;
  int rc = make_qkv_map(&tq, p->q, p->Sq, p->n_q_heads, p->B, p->q_sb, p->q_sh, p->q_sl);
  if (rc != VY_OK) return rc;
  rc = make_qkv_map(&tk, p->k, p->Skv, p->n_kv_heads, p->B, p->k_sb, p->k_sh, p->k_sl);
  if (rc != VY_OK) return rc;
  rc = make_qkv_map(&tv, p->v, p->Skv, p->n_kv_heads, p->B, p->v_sb, p->v_sh, p->v_sl);
  if (rc != VY_OK) return rc;

  AttnDev g;
  g.B = p->B; g.Hq = p->n_q_heads; g.Hkv = p->n_kv_heads; g.Sq = p->Sq; g.Skv = p->Skv;
  g.n_rep = p->n_q_heads / p->n_kv_heads;
  g.causal = p->causal; g.q_pos0 = p->q_pos0;
  g.kpm = p->key_padding_mask; g.kpm_sb = p->kpm_stride;
  g.out = p->out; g.o_sb = p->o_sb; g.o_sl = p->o_sl; g.out_dtype = p->out_dtype;
  g.lse = p->lse;
  g.scale_log2 = 1.4426950408889634f / 8.0f;

  static bool attr_set = false;
  if (!attr_set) {
    VY_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    attr_set = true;
  }
  dim3 grid((p->Sq + AT_BM - 1) / AT_BM, p->n_q_heads, p->B);
  VY_CUDA_OK(launch_kernel(attn_fwd_kernel, dim3(grid), dim3(256), AT_SMEM, static_cast<cudaStream_t>(p->stream), tq, tk, tv, g));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}
