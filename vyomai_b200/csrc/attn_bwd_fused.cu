// vy_attn_bwd, short-sequence path: dQ, dK and dV of one (batch row, kv head) in ONE CTA.
//
// The transformer blocks of the reference run attention over short sequences (128 decoder positions,
// 197 ViT tokens), where a key tile / query tile pair is most of the problem. Splitting the backward
// into a dK/dV kernel and a dQ kernel then recomputes S and dP and pays two prologues per head; here one
// CTA keeps K and V of its kv head resident and walks the (query head, key tile, query tile) pairs:
//
//     S^T = K Q^T,  dP^T = V dO^T                         (M = keys, N = queries; accumulators in TMEM)
//     P^T = exp2(S^T c - lse),  dS^T = P^T (dP^T - D)     (8 warps, thread <-> key row, bf16 into smem)
//     dV += P^T dO,  dK += dS^T Q,  dQ += dS K            (dS^T in smem is read K-major for dK and
//                                                          MN-major for dQ — same bytes, two descriptors)
//
// Two shapes are covered (everything else takes the two-kernel path of attn_bwd.cu):
//   mode A  Sq, Skv <= 128, any GQA group: iterations = the n_rep query heads of the group; dK/dV accumulate
//           over them (no atomics), dQ of a head is complete after its iteration and is flushed at once.
//   mode B  n_rep == 1, Sq, Skv <= 256: iterations = (key tile, query tile); dV/dK of a key tile accumulate
//           over query tiles, the dQ accumulators of both query tiles live in TMEM until the end.
// TMEM: S^T 128 + dP^T 128 + dV 64 + dK 64 + dQ 2 x 64 = 512 columns.
// Masks follow vy_attn_fwd exactly (finite "finfo.min" scores, zero weight beyond Skv / Sq); the epilogues
// multiply by 1/sqrt(d), undo the RoPE rotation of q / k and write into the packed dQKV gradient.
#include "vy_common.cuh"
#include "vy_ptx.cuh"

namespace vy {

constexpr int AF_T = 128;
constexpr int AF_D = 64;
constexpr int AF_TILE = AF_T * AF_D * 2;    // 16 KB
constexpr int AF_PBYTES = AF_T * AF_T * 2;  // 32 KB
constexpr float AF_MASKED = -30000.0f;      // must equal AT_MASKED of attn_fwd.cu
constexpr int AF_SOFTMAX_WARPS = 8;
constexpr int AF_THREADS = (2 + AF_SOFTMAX_WARPS) * 32;
constexpr int AF_SMEM = 8 * AF_TILE + 2 * AF_PBYTES + 2 * AF_T * 4 + 256;

struct AttnBwdFusedDev {
  int B, Hq, Hkv, Sq, Skv, n_rep, causal, q_pos0;
  int mode_b, QT, KT;
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

__device__ __forceinline__ uint32_t af_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 32 packed bf16 values (columns c4*32 .. +32 of row `row`) into a K-major [128 x 128] operand made of
// two [128 x 64] 128B-swizzle atoms
__device__ __forceinline__ void af_store_chunk(uint8_t* base, int row, int c4, const uint32_t (&packed)[16]) {
  uint8_t* rowp = base + (c4 >> 1) * (AF_T * 128) + row * 128;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const int chunk = ((c4 & 1) * 4 + ch) ^ (row & 7);
    *reinterpret_cast<uint4*>(rowp + chunk * 16) =
        make_uint4(packed[ch * 4], packed[ch * 4 + 1], packed[ch * 4 + 2], packed[ch * 4 + 3]);
  }
}

// rotation factors of one token for the pairs j0 .. j0+16 (four 16-byte loads per table)
struct AfRope {
  float cs[16], sn[16];
  int row;  // table row held, -1 = none
};
__device__ __forceinline__ void af_load_rope(AfRope& r, const float* cos_t, const float* sin_t, int row, int j0) {
  if (!cos_t || r.row == row) return;
  const float4* cp = reinterpret_cast<const float4*>(cos_t + static_cast<long long>(row) * 32 + j0);
  const float4* sp = reinterpret_cast<const float4*>(sin_t + static_cast<long long>(row) * 32 + j0);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 c4 = cp[j], s4 = sp[j];
    r.cs[j * 4] = c4.x; r.cs[j * 4 + 1] = c4.y; r.cs[j * 4 + 2] = c4.z; r.cs[j * 4 + 3] = c4.w;
    r.sn[j * 4] = s4.x; r.sn[j * 4 + 1] = s4.y; r.sn[j * 4 + 2] = s4.z; r.sn[j * 4 + 3] = s4.w;
  }
  r.row = row;
}

// One warp writes its 16 rotation pairs (columns j0..j0+16 and 32+j0..) of a 64-wide gradient row held in TMEM.
__device__ __forceinline__ void af_store_half_row(uint32_t taddr, int j0, void* dst, int dt, long long off, bool row_ok,
                                                  float scale, bool rotate, const AfRope& r) {
  uint32_t lo[16], hi[16];
  tmem_ld_x16(taddr + j0, lo);
  tmem_ld_x16(taddr + 32 + j0, hi);
  tmem_ld_wait();
  if (!row_ok) return;
#pragma unroll
  for (int h8 = 0; h8 < 2; ++h8) {
    float o1[8], o2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = __uint_as_float(lo[h8 * 8 + j]) * scale, b = __uint_as_float(hi[h8 * 8 + j]) * scale;
      if (rotate) {
        const float c = r.cs[h8 * 8 + j], s = r.sn[h8 * 8 + j];
        o1[j] = a * c + b * s;  // transpose of the forward rotation
        o2[j] = b * c - a * s;
      } else {
        o1[j] = a;
        o2[j] = b;
      }
    }
    st8_from_float(dst, dt, off + j0 + h8 * 8, o1);
    st8_from_float(dst, dt, off + 32 + j0 + h8 * 8, o2);
  }
}

__global__ void __launch_bounds__(AF_THREADS, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                      const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                      const __grid_constant__ AttnBwdFusedDev g) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;                   // [2] key tiles
  uint8_t* sV = smem + 2 * AF_TILE;     // [2]
  uint8_t* sQ = smem + 4 * AF_TILE;     // [2] mode A: ring over heads, mode B: query tiles
  uint8_t* sdO = smem + 6 * AF_TILE;    // [2]
  uint8_t* sP = smem + 8 * AF_TILE;
  uint8_t* sdS = sP + AF_PBYTES;
  float* s_lse = reinterpret_cast<float*>(sdS + AF_PBYTES);  // [128]
  float* s_D = s_lse + AF_T;                                  // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_D + AF_T);
  uint64_t* kv_full = bars;          // [2]
  uint64_t* qdo_full = bars + 2;     // [2]
  uint64_t* qdo_empty = bars + 4;    // [2]
  uint64_t* sdp_full = bars + 6;     // MMA -> softmax: S^T / dP^T ready
  uint64_t* pds_full = bars + 7;     // softmax -> MMA: P^T / dS^T in smem, S^T / dP^T consumed
  uint64_t* pds_empty = bars + 8;    // MMA -> softmax: P^T / dS^T consumed
  uint64_t* dq_full = bars + 9;      // MMA -> epilogue
  uint64_t* dq_empty = bars + 10;    // epilogue -> MMA (mode A re-uses the dQ accumulator every head)
  uint64_t* dkdv_full = bars + 11;
  uint64_t* dkdv_empty = bars + 12;  // mode B re-uses dK / dV for the second key tile
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kvh = blockIdx.x;
  const int b = blockIdx.y;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    tma_prefetch_desc(&tma_do);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&qdo_full[s], 1);
      mbar_init(&qdo_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, AF_SOFTMAX_WARPS);
    mbar_init(pds_empty, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, AF_SOFTMAX_WARPS);
    mbar_init(dkdv_full, 1);
    mbar_init(dkdv_empty, AF_SOFTMAX_WARPS);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  pdl_wait();  // everything above is independent of the predecessor grid's output
  const uint32_t tm_S = tmem_base, tm_dP = tmem_base + 128, tm_dV = tmem_base + 256, tm_dK = tmem_base + 320,
                 tm_dQ = tmem_base + 384;

  const bool mode_b = g.mode_b != 0;
  const int QT = g.QT, KT = g.KT;
  const int n_it = mode_b ? KT * QT : g.n_rep;
  // iteration -> (query head, key tile, query tile, Q/dO slot)
  auto it_head = [&](int it) { return mode_b ? kvh : kvh * g.n_rep + it; };
  auto it_kt = [&](int it) { return mode_b ? it / QT : 0; };
  auto it_qt = [&](int it) { return mode_b ? it % QT : 0; };
  auto it_slot = [&](int it) { return mode_b ? it % QT : (it & 1); };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(&kv_full[0], 2 * AF_TILE);
      tma_load_4d(sK, &tma_k, &kv_full[0], 0, 0, kvh, b);
      tma_load_4d(sV, &tma_v, &kv_full[0], 0, 0, kvh, b);
      if (mode_b) {
        for (int qt = 0; qt < QT; ++qt) {
          mbar_arrive_expect_tx(&qdo_full[qt], 2 * AF_TILE);
          tma_load_4d(sQ + qt * AF_TILE, &tma_q, &qdo_full[qt], 0, qt * AF_T, kvh, b);
          tma_load_4d(sdO + qt * AF_TILE, &tma_do, &qdo_full[qt], 0, qt * AF_T, kvh, b);
          if (qt == 0 && KT > 1) {
            mbar_arrive_expect_tx(&kv_full[1], 2 * AF_TILE);
            tma_load_4d(sK + AF_TILE, &tma_k, &kv_full[1], 0, AF_T, kvh, b);
            tma_load_4d(sV + AF_TILE, &tma_v, &kv_full[1], 0, AF_T, kvh, b);
          }
        }
      } else {
        for (int it = 0; it < n_it; ++it) {
          const int s = it & 1;
          mbar_wait(&qdo_empty[s], ((it >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&qdo_full[s], 2 * AF_TILE);
          tma_load_4d(sQ + s * AF_TILE, &tma_q, &qdo_full[s], 0, 0, it_head(it), b);
          tma_load_4d(sdO + s * AF_TILE, &tma_do, &qdo_full[s], 0, 0, it_head(it), b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_sc = make_idesc(1, AF_T, AF_T, 0, 0);   // K-major x K-major, N = 128
      constexpr uint32_t idesc_kv = make_idesc(1, AF_T, AF_D, 0, 1);   // A K-major, B MN-major, N = 64
      constexpr uint32_t idesc_dq = make_idesc(1, AF_T, AF_D, 1, 1);   // A MN-major (dS^T read transposed), B MN-major
      const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
      int dq_flushes = 0, dkdv_flushes = 0;
      for (int it = 0; it < n_it; ++it) {
        const int kt = it_kt(it), qt = it_qt(it), s = it_slot(it);
        const uint32_t k_addr = smem_u32(sK + kt * AF_TILE), v_addr = smem_u32(sV + kt * AF_TILE);
        const uint32_t q_addr = smem_u32(sQ + s * AF_TILE), do_addr = smem_u32(sdO + s * AF_TILE);
        if (mode_b) {
          if (qt == 0) mbar_wait(&kv_full[kt], 0);
          if (kt == 0) mbar_wait(&qdo_full[qt], 0);
        } else {
          if (it == 0) mbar_wait(&kv_full[0], 0);
          mbar_wait(&qdo_full[s], (it >> 1) & 1);
        }
        // S^T / dP^T are free: pds_full(it - 1) was awaited below before the previous iteration's second half
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AF_D / 16; ++k)
          umma_f16(tm_S, make_smem_desc_sw128(k_addr + k * 32, 16, 1024), make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                   idesc_sc, k != 0);
#pragma unroll
        for (int k = 0; k < AF_D / 16; ++k)
          umma_f16(tm_dP, make_smem_desc_sw128(v_addr + k * 32, 16, 1024), make_smem_desc_sw128(do_addr + k * 32, 16, 1024),
                   idesc_sc, k != 0);
        umma_commit(sdp_full);

        mbar_wait(pds_full, it & 1);
        // accumulator re-use: wait until the epilogue warps have drained what a previous flush handed them
        const bool dq_fresh = mode_b ? (kt == 0) : true;
        const bool kv_fresh = mode_b ? (qt == 0) : (it == 0);
        if (!mode_b && it > 0) mbar_wait(dq_empty, (dq_flushes - 1) & 1);
        if (mode_b && qt == 0 && kt > 0) mbar_wait(dkdv_empty, (dkdv_flushes - 1) & 1);
        tc_fence_after();
        const uint32_t dq_tm = tm_dQ + (mode_b ? qt * AF_D : 0);
#pragma unroll
        for (int kk = 0; kk < AF_T / 16; ++kk) {
          const uint32_t a_off = (kk >> 2) * (AF_T * 128) + (kk & 3) * 32;
          umma_f16(tm_dV, make_smem_desc_sw128(p_addr + a_off, 16, 1024),
                   make_smem_desc_sw128(do_addr + kk * 2048, 8192, 1024), idesc_kv, !(kv_fresh && kk == 0));
        }
#pragma unroll
        for (int kk = 0; kk < AF_T / 16; ++kk) {
          const uint32_t a_off = (kk >> 2) * (AF_T * 128) + (kk & 3) * 32;
          umma_f16(tm_dK, make_smem_desc_sw128(ds_addr + a_off, 16, 1024),
                   make_smem_desc_sw128(q_addr + kk * 2048, 8192, 1024), idesc_kv, !(kv_fresh && kk == 0));
        }
        // dQ[q, :] += sum_k dS[q, k] K[k, :]: A = dS^T bytes read MN-major (M = queries contiguous in a 128-B row,
        // 64 per atom, atoms 16 KB apart; K = key rows, 8 per 1024-B swizzle group)
#pragma unroll
        for (int kk = 0; kk < AF_T / 16; ++kk)
          umma_f16(dq_tm, make_smem_desc_sw128(ds_addr + kk * 2048, AF_T * 128, 1024),
                   make_smem_desc_sw128(k_addr + kk * 2048, 8192, 1024), idesc_dq, !(dq_fresh && kk == 0));
        umma_commit(pds_empty);
        if (!mode_b) umma_commit(&qdo_empty[s]);
        const bool dq_done = mode_b ? (kt == KT - 1) : true;
        const bool kv_done = mode_b ? (qt == QT - 1) : (it == n_it - 1);
        if (dq_done) {
          umma_commit(dq_full);
          ++dq_flushes;
        }
        if (kv_done) {
          umma_commit(dkdv_full);
          ++dkdv_flushes;
        }
      }
    }
  } else {
    // ===================== softmax / epilogue warps =====================
    const int e = warp - 2;
    const int qd = warp & 3;  // TMEM lane quarter of this warp
    const int half = e >> 2;  // which 64 of the tile's 128 columns
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const int et = threadIdx.x - 64;
    const int j0 = half * 16;
    int dq_flushes = 0, dkdv_flushes = 0;
    const bool rotate = g.rope_cos != nullptr;
    AfRope rope;
    rope.row = -1;
    const int rope_max = (g.Sq > g.Skv ? g.Sq : g.Skv) - 1;  // rows beyond the sequences are never stored: clamp their loads

    for (int it = 0; it < n_it; ++it) {
      const int head = it_head(it), kt = it_kt(it), qt = it_qt(it);
      const int key = kt * AF_T + row;
      {
        // rotation factors of the row this iteration's flush needs (query row first), fetched before the softmax so
        // the loads are long done when the accumulators arrive
        const bool dq_done_ = mode_b ? (kt == KT - 1) : true;
        const int need = dq_done_ ? qt * AF_T + row : key;
        af_load_rope(rope, g.rope_cos, g.rope_sin, g.rope_pos0 + (need < rope_max ? need : rope_max), j0);
      }
      const bool key_in = key < g.Skv;
      const bool key_vis = key_in && (!g.kpm || g.kpm[b * g.kpm_sb + key] != 0);
      const int q0 = qt * AF_T;
      // lse / D of this (head, query tile); the previous iteration's readers are past pds_full(it - 1), which every
      // warp has arrived on before any warp can be here (named barrier below)
      named_bar_sync(1, AF_SOFTMAX_WARPS * 32);
      if (et < AF_T) {
        const int qq = q0 + et;
        const long long idx = (static_cast<long long>(b) * g.Hq + head) * g.Sq + qq;
        s_lse[et] = qq < g.Sq ? g.lse[idx] : 0.f;
        s_D[et] = qq < g.Sq ? g.dsum[idx] : 0.f;
      }
      named_bar_sync(1, AF_SOFTMAX_WARPS * 32);
      mbar_wait(sdp_full, it & 1);
      tc_fence_after();
      mbar_wait(pds_empty, (it & 1) ^ 1);
      const int qlim = g.Sq - q0;               // columns >= qlim are beyond the sequence
      const int cmax = g.causal ? key - g.q_pos0 - q0 : -1;  // causal: column c visible iff c >= cmax
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c4 = half * 2 + cc;
        uint32_t sraw[32], praw[32];
        tmem_ld_x32(tm_S + lane_off + c4 * 32, sraw);
        tmem_ld_x32(tm_dP + lane_off + c4 * 32, praw);
        tmem_ld_wait();
        uint32_t pp[16], dd[16];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 l4 = *reinterpret_cast<const float4*>(s_lse + c4 * 32 + i4 * 4);
          const float4 d4 = *reinterpret_cast<const float4*>(s_D + c4 * 32 + i4 * 4);
          const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
          const float dsm[4] = {d4.x, d4.y, d4.z, d4.w};
          float pv[4], dv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = c4 * 32 + i4 * 4 + u;
            const bool vis = key_vis && c >= cmax;
            const float t = vis ? __uint_as_float(sraw[i4 * 4 + u]) * g.scale_log2 : AF_MASKED;
            float p = vy_ex2_approx(t - ls[u]);
            p = (key_in && c < qlim) ? p : 0.f;
            pv[u] = p;
            dv[u] = p * (__uint_as_float(praw[i4 * 4 + u]) - dsm[u]);
          }
          pp[i4 * 2] = af_pack(pv[0], pv[1]);
          pp[i4 * 2 + 1] = af_pack(pv[2], pv[3]);
          dd[i4 * 2] = af_pack(dv[0], dv[1]);
          dd[i4 * 2 + 1] = af_pack(dv[2], dv[3]);
        }
        af_store_chunk(sP, row, c4, pp);
        af_store_chunk(sdS, row, c4, dd);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);

      const bool dq_done = mode_b ? (kt == KT - 1) : true;
      const bool kv_done = mode_b ? (qt == QT - 1) : (it == n_it - 1);
      if (dq_done) {
        mbar_wait(dq_full, dq_flushes & 1);
        ++dq_flushes;
        tc_fence_after();
        const int qrow = q0 + row;
        const bool q_in = qrow < g.Sq;
        af_load_rope(rope, g.rope_cos, g.rope_sin, g.rope_pos0 + (qrow < rope_max ? qrow : rope_max), j0);
        af_store_half_row(tm_dQ + (mode_b ? qt * AF_D : 0) + lane_off, j0, g.dq, g.out_dtype,
                          (static_cast<long long>(b) * g.Sq + qrow) * g.ld_dq + head * AF_D, q_in, g.scale, rotate, rope);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_empty);
      }
      if (kv_done) {
        mbar_wait(dkdv_full, dkdv_flushes & 1);
        ++dkdv_flushes;
        tc_fence_after();
        const long long tok = static_cast<long long>(b) * g.Skv + key;
        af_store_half_row(tm_dV + lane_off, j0, g.dv, g.out_dtype, tok * g.ld_dv + kvh * AF_D, key_in, 1.f, false, rope);
        af_load_rope(rope, g.rope_cos, g.rope_sin, g.rope_pos0 + (key < rope_max ? key : rope_max), j0);
        af_store_half_row(tm_dK + lane_off, j0, g.dk, g.out_dtype, tok * g.ld_dk + kvh * AF_D, key_in, g.scale, rotate, rope);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dkdv_empty);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int af_make_map4(CUtensorMap* out, const void* base, int S, int H, int B, long long sb, long long sh, long long sl) {
  uint64_t dims[4] = {static_cast<uint64_t>(AF_D), static_cast<uint64_t>(S), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
  uint64_t strides[4] = {0, static_cast<uint64_t>(sl) * 2, static_cast<uint64_t>(sh) * 2, static_cast<uint64_t>(sb) * 2};
  uint32_t box[4] = {AF_D, AF_T, 1, 1};
  return get_tensor_map_cached(out, VY_BF16, 4, base, dims, strides, box, 1);
}

// Returns 1 if the shape is covered by the fused kernel (and launches it), 0 if not, negative on error.
int attn_bwd_fused_launch(const VyAttnBwd* p) {
  const int n_rep = p->n_q_heads / p->n_kv_heads;
  const bool mode_a = p->Sq <= AF_T && p->Skv <= AF_T;
  const bool mode_b = !mode_a && n_rep == 1 && p->Sq <= 2 * AF_T && p->Skv <= 2 * AF_T;
  if (!mode_a && !mode_b) return 0;
  CUtensorMap tq, tk, tv, tdo;
  int rc = af_make_map4(&tq, p->q, p->Sq, p->n_q_heads, p->B, p->q_sb, p->q_sh, p->q_sl);
  if (rc != VY_OK) return rc;
  rc = af_make_map4(&tk, p->k, p->Skv, p->n_kv_heads, p->B, p->k_sb, p->k_sh, p->k_sl);
  if (rc != VY_OK) return rc;
  rc = af_make_map4(&tv, p->v, p->Skv, p->n_kv_heads, p->B, p->v_sb, p->v_sh, p->v_sl);
  if (rc != VY_OK) return rc;
  rc = af_make_map4(&tdo, p->dout, p->Sq, p->n_q_heads, p->B, p->do_sb, 64, p->do_sl);
  if (rc != VY_OK) return rc;

  AttnBwdFusedDev g;
  memset(&g, 0, sizeof(g));
  g.B = p->B; g.Hq = p->n_q_heads; g.Hkv = p->n_kv_heads; g.Sq = p->Sq; g.Skv = p->Skv;
  g.n_rep = n_rep; g.causal = p->causal; g.q_pos0 = p->q_pos0;
  g.mode_b = mode_b ? 1 : 0;
  g.QT = (p->Sq + AF_T - 1) / AF_T;
  g.KT = (p->Skv + AF_T - 1) / AF_T;
  g.kpm = p->key_padding_mask; g.kpm_sb = p->kpm_stride;
  g.lse = p->lse; g.dsum = p->dsum;
  g.dq = p->dq; g.ld_dq = p->ld_dq; g.dk = p->dk; g.ld_dk = p->ld_dk; g.dv = p->dv; g.ld_dv = p->ld_dv;
  g.out_dtype = p->out_dtype;
  g.rope_cos = p->rope_cos; g.rope_sin = p->rope_sin; g.rope_pos0 = p->rope_pos0;
  g.scale = 0.125f;
  g.scale_log2 = 1.4426950408889634f * 0.125f;

  static bool attr_set = false;
  if (!attr_set) {
    VY_CUDA_OK(cudaFuncSetAttribute(attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM));
    attr_set = true;
  }
  dim3 grid(p->n_kv_heads, p->B);
  VY_CUDA_OK(launch_kernel(attn_bwd_fused_kernel, dim3(grid), dim3(AF_THREADS), AF_SMEM, static_cast<cudaStream_t>(p->stream), tq, tk, tv, tdo, g));
  VY_LAUNCH_OK();
  return 1;
}

}  // namespace vy
