// vy_gemm host side: argument checking, tile-width choice and dispatch to the kernel instantiations
// (gemm_kernel.cuh; instantiated in gemm_inst_*.cu so the translation units build in parallel).
#include <stdlib.h>

#include "gemm_kernel.cuh"

namespace vy {

#define VY_GEMM_EXTERN(T, BN, A, B) extern template int launch_gemm<T, BN, A, B>(const VyGemm*, const GemmDev&);
#define VY_GEMM_EXTERN_ALL(T)                                                                                   \
  VY_GEMM_EXTERN(T, 32, false, false) VY_GEMM_EXTERN(T, 64, false, false) VY_GEMM_EXTERN(T, 128, false, false) \
  VY_GEMM_EXTERN(T, 192, false, false) VY_GEMM_EXTERN(T, 256, false, false)                                    \
  VY_GEMM_EXTERN(T, 128, false, true) VY_GEMM_EXTERN(T, 192, false, true) VY_GEMM_EXTERN(T, 256, false, true)  \
  VY_GEMM_EXTERN(T, 128, true, false) VY_GEMM_EXTERN(T, 192, true, false) VY_GEMM_EXTERN(T, 256, true, false)  \
  VY_GEMM_EXTERN(T, 128, true, true) VY_GEMM_EXTERN(T, 192, true, true) VY_GEMM_EXTERN(T, 256, true, true)
VY_GEMM_EXTERN_ALL(__nv_bfloat16)
VY_GEMM_EXTERN_ALL(float)

template <typename TIn>
static int dispatch_gemm(const VyGemm* p, const GemmDev& g, int bn) {
  const bool amn = p->a_mn_major != 0, bmn = p->b_mn_major != 0;
#define VY_GEMM_CASE(BN_, A_, B_) \
  case BN_: return launch_gemm<TIn, BN_, A_, B_>(p, g)
  if (!amn && !bmn) {
    switch (bn) {
      VY_GEMM_CASE(32, false, false);
      VY_GEMM_CASE(64, false, false);
      VY_GEMM_CASE(128, false, false);
      VY_GEMM_CASE(192, false, false);
      default: return launch_gemm<TIn, 256, false, false>(p, g);
    }
  }
  if (bn < 128) bn = 128;  // MN-major operands arrive in 64-element (bf16) / 32-element (tf32) boxes
  if (!amn && bmn) {
    switch (bn) {
      VY_GEMM_CASE(128, false, true);
      VY_GEMM_CASE(192, false, true);
      default: return launch_gemm<TIn, 256, false, true>(p, g);
    }
  }
  if (amn && !bmn) {
    switch (bn) {
      VY_GEMM_CASE(128, true, false);
      VY_GEMM_CASE(192, true, false);
      default: return launch_gemm<TIn, 256, true, false>(p, g);
    }
  }
  switch (bn) {
    VY_GEMM_CASE(128, true, true);
    VY_GEMM_CASE(192, true, true);
    default: return launch_gemm<TIn, 256, true, true>(p, g);
  }
#undef VY_GEMM_CASE
}

// Tile width: the persistent grid walks ceil(tiles / SMs) waves of 128 x BN tiles, so the cost of a
// candidate is waves * BN, weighted by how well a tile of that width feeds the tensor pipe (at BN <= 128
// the A + B shared-memory reads per MMA reach the 128 B/clk of the SM; narrow tiles also amortise the
// fixed per-tile cost worse). Ties go to the wider tile.
static int choose_bn(int M, int N, bool mn_major, bool qkv) {
  static const int cand[5] = {256, 192, 128, 64, 32};
  static const double pen[5] = {1.0, 1.0, 1.12, 1.6, 2.6};
  const int m_tiles = (M + 127) / 128;
  const int sms = num_sms();
  int best = 256;
  double best_cost = 1e30;
  for (int i = 0; i < 5; ++i) {
    const int bn = cand[i];
    if (mn_major && bn < 128) continue;
    if (qkv && bn < 64) continue;
    if (bn > 32 && bn / 2 >= N) continue;  // more than half of the tile would be padding
    const long long tiles = static_cast<long long>(m_tiles) * ((N + bn - 1) / bn);
    const long long waves = (tiles + sms - 1) / sms;
    const double cost = static_cast<double>(waves) * bn * pen[i];
    if (cost < best_cost * 0.999) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
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
  static const int dbg = getenv("VY_GEMM_DEBUG") ? atoi(getenv("VY_GEMM_DEBUG")) : 0;
  g.debug = dbg;
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

  const int bn = choose_bn(p->M, p->N, p->a_mn_major || p->b_mn_major, p->epi == VY_EPI_QKV_ROPE);

  if (p->in_dtype == VY_BF16) return dispatch_gemm<__nv_bfloat16>(p, g, bn);
  return dispatch_gemm<float>(p, g, bn);
}
