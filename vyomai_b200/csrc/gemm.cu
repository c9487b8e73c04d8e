// vy_gemm host side: argument checking, tile-width choice and dispatch to the kernel instantiations
// (gemm_kernel.cuh; instantiated in gemm_inst_*.cu so the translation units build in parallel).
#include <stdlib.h>

#include <mutex>

#include "gemm_kernel.cuh"

namespace vy {

#define VY_GEMM_EXTERN(T, BN, A, B, P) extern template int launch_gemm<T, BN, A, B, P>(const VyGemm*, const GemmDev&);
#define VY_GEMM_EXTERN_ALL(T)                                                                                   \
  VY_GEMM_EXTERN(T, 32, false, false, false) VY_GEMM_EXTERN(T, 64, false, false, false) VY_GEMM_EXTERN(T, 128, false, false, false) \
  VY_GEMM_EXTERN(T, 192, false, false, false) VY_GEMM_EXTERN(T, 256, false, false, false)                                    \
  VY_GEMM_EXTERN(T, 128, false, true, false) VY_GEMM_EXTERN(T, 192, false, true, false) VY_GEMM_EXTERN(T, 256, false, true, false)  \
  VY_GEMM_EXTERN(T, 128, true, false, false) VY_GEMM_EXTERN(T, 192, true, false, false) VY_GEMM_EXTERN(T, 256, true, false, false)  \
  VY_GEMM_EXTERN(T, 128, true, true, false) VY_GEMM_EXTERN(T, 192, true, true, false) VY_GEMM_EXTERN(T, 256, true, true, false)
VY_GEMM_EXTERN_ALL(__nv_bfloat16)
VY_GEMM_EXTERN_ALL(float)
// CTA-pair (cta_group::2) kernels: bf16 only; an MN-major B half must be whole 64-column boxes, so no 192 there
#define VY_GEMM_EXTERN_PAIR(A, B, BN) VY_GEMM_EXTERN(__nv_bfloat16, BN, A, B, true)
VY_GEMM_EXTERN_PAIR(false, false, 128) VY_GEMM_EXTERN_PAIR(false, false, 192) VY_GEMM_EXTERN_PAIR(false, false, 256)
VY_GEMM_EXTERN_PAIR(true, false, 128) VY_GEMM_EXTERN_PAIR(true, false, 192) VY_GEMM_EXTERN_PAIR(true, false, 256)
VY_GEMM_EXTERN_PAIR(false, true, 128) VY_GEMM_EXTERN_PAIR(false, true, 256)
VY_GEMM_EXTERN_PAIR(true, true, 128) VY_GEMM_EXTERN_PAIR(true, true, 256)

// staged SwiGLU epilogue (inference form): bf16, both operands K-major, BN 128 | 256, either flavour
extern template int launch_gemm<__nv_bfloat16, 128, false, false, false, true>(const VyGemm*, const GemmDev&);
extern template int launch_gemm<__nv_bfloat16, 256, false, false, false, true>(const VyGemm*, const GemmDev&);
extern template int launch_gemm<__nv_bfloat16, 128, false, false, true, true>(const VyGemm*, const GemmDev&);
extern template int launch_gemm<__nv_bfloat16, 256, false, false, true, true>(const VyGemm*, const GemmDev&);
static int dispatch_gemm_gated(const VyGemm* p, const GemmDev& g, int bn, bool pair) {
  using T = __nv_bfloat16;
  if (pair) return bn == 128 ? launch_gemm<T, 128, false, false, true, true>(p, g) : launch_gemm<T, 256, false, false, true, true>(p, g);
  return bn == 128 ? launch_gemm<T, 128, false, false, false, true>(p, g) : launch_gemm<T, 256, false, false, false, true>(p, g);
}

static int dispatch_gemm_pair(const VyGemm* p, const GemmDev& g, int bn) {
  using T = __nv_bfloat16;
  const bool amn = p->a_mn_major != 0, bmn = p->b_mn_major != 0;
  if (!bmn) {
    if (bn == 128) return amn ? launch_gemm<T, 128, true, false, true>(p, g) : launch_gemm<T, 128, false, false, true>(p, g);
    if (bn == 192) return amn ? launch_gemm<T, 192, true, false, true>(p, g) : launch_gemm<T, 192, false, false, true>(p, g);
    return amn ? launch_gemm<T, 256, true, false, true>(p, g) : launch_gemm<T, 256, false, false, true>(p, g);
  }
  if (bn == 128) return amn ? launch_gemm<T, 128, true, true, true>(p, g) : launch_gemm<T, 128, false, true, true>(p, g);
  return amn ? launch_gemm<T, 256, true, true, true>(p, g) : launch_gemm<T, 256, false, true, true>(p, g);
}

template <typename TIn>
static int dispatch_gemm(const VyGemm* p, const GemmDev& g, int bn) {
  const bool amn = p->a_mn_major != 0, bmn = p->b_mn_major != 0;
#define VY_GEMM_CASE(BN_, A_, B_) \
  case BN_: return launch_gemm<TIn, BN_, A_, B_, false>(p, g)
  if (!amn && !bmn) {
    switch (bn) {
      VY_GEMM_CASE(32, false, false);
      VY_GEMM_CASE(64, false, false);
      VY_GEMM_CASE(128, false, false);
      VY_GEMM_CASE(192, false, false);
      default: return launch_gemm<TIn, 256, false, false, false>(p, g);
    }
  }
  if (bn < 128) bn = 128;  // MN-major operands arrive in 64-element (bf16) / 32-element (tf32) boxes
  if (!amn && bmn) {
    switch (bn) {
      VY_GEMM_CASE(128, false, true);
      VY_GEMM_CASE(192, false, true);
      default: return launch_gemm<TIn, 256, false, true, false>(p, g);
    }
  }
  if (amn && !bmn) {
    switch (bn) {
      VY_GEMM_CASE(128, true, false);
      VY_GEMM_CASE(192, true, false);
      default: return launch_gemm<TIn, 256, true, false, false>(p, g);
    }
  }
  switch (bn) {
    VY_GEMM_CASE(128, true, true);
    VY_GEMM_CASE(192, true, true);
    default: return launch_gemm<TIn, 256, true, true, false>(p, g);
  }
#undef VY_GEMM_CASE
}

// Kernel flavour, tile width and K split. The persistent grid walks ceil(units / workers) waves of units — 128 x BN x
// (K / splits) on one SM, or 256 x BN x (K / splits) on a CTA pair — and every candidate is scored with a small time
// model fitted to a sweep of all candidates over the GEMM shapes of the captioner step on B200 (tools/gemm_sweep.py,
// profiles/r01_gemm_sweep.txt; rms error 10%, total time of its choices within 4% of the per-shape best — callers that
// can afford to time the candidates themselves pass the winner as a hint instead, see VyGemm.hint_*):
//   t [us] = 14 + 3 * a + waves * (k-blocks per unit * a + e * BN / 256) (+ 3 + 0.058 * (4 * splits + 4) * M * N / 1e6)
// a = time of one 64-wide k-block of a unit (CTA pairs halve the B rows each SM reads from shared memory, which is what
// holds single-CTA tiles below the tensor pipe's rate), e = epilogue time per 256 columns of a unit, the last term the
// fp32 slab traffic of the split-K reduce pass.
struct Tiling {
  int pair, bn, splits;
  double cost;  // modelled time, us
};
static Tiling choose_tiling(int M, int N, int num_kb, bool mn_major, bool b_mn, bool qkv, double epi_cost, int max_splits,
                            int pair_lo, int pair_hi) {
  static const int cand[5] = {256, 192, 128, 64, 32};
  static const double kb_single[5] = {0.303, 0.261, 0.19, 0.15, 0.13};
  static const double kb_pair[3] = {0.27, 0.218, 0.162};
  static const int split_cand[6] = {1, 2, 3, 4, 6, 8};
  Tiling best = {0, 256, 1, 1e30};
  for (int pair = pair_lo; pair <= pair_hi; ++pair) {
    const int m_tiles = pair ? (M + 255) / 256 : (M + 127) / 128;
    const int workers = pair ? num_sms() / 2 : num_sms();
    for (int i = 0; i < (pair ? 3 : 5); ++i) {
      const int bn = cand[i];
      if (mn_major && bn < 128) continue;
      if (pair && b_mn && bn == 192) continue;  // an MN-major B half must be whole 64-column boxes
      if (qkv && bn < 64) continue;
      if (bn > 32 && bn / 2 >= N) continue;  // more than half of the tile would be padding
      const double a = pair ? kb_pair[i] : kb_single[i];
      const long long tiles = static_cast<long long>(m_tiles) * ((N + bn - 1) / bn);
      for (int si = 0; si < 6; ++si) {
        const int sp = split_cand[si];
        if (sp > max_splits) break;
        if (sp > 1 && num_kb / sp < 16) break;  // keep the mainloop of a unit long enough to amortise its epilogue
        const int kb_per = (num_kb + sp - 1) / sp;
        if (sp > 1 && static_cast<long long>(sp - 1) * kb_per >= num_kb) continue;  // an empty last split
        const long long waves = (tiles * sp + workers - 1) / workers;
        double cost = 14.0 + 3.0 * a + static_cast<double>(waves) * (kb_per * a + epi_cost * bn / 256.0);
        if (sp > 1) cost += 3.0 + 0.058 * (sp * 4.0 + 4.0) * M * N / 1.0e6;
        if (cost < best.cost * 0.999) best = {pair, bn, sp, cost};
      }
    }
  }
  return best;
}

// out[r, c] = scale * (sum_s ws[s][r][c] + bias[c] + addend[r, c]) — finishes a split-K GEMM. 8 columns per thread.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(int M, int N, int splits, const float* __restrict__ ws, const void* __restrict__ bias, int bias_dt,
                     const void* __restrict__ addend, long long ld_addend, int addend_dt, float scale, void* __restrict__ out,
                     long long ld_out, int out_dt) {
  pdl_trigger();
  pdl_wait();
  const long long nvec = static_cast<long long>(M) * (N >> 3);
  const long long slab = static_cast<long long>(M) * N;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < nvec;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(v / (N >> 3));
    const int c = static_cast<int>(v % (N >> 3)) * 8;
    float acc[8];
    ld8_as_float(ws, VY_F32, static_cast<long long>(r) * N + c, acc);
    for (int s = 1; s < splits; ++s) {
      float t[8];
      ld8_as_float(ws, VY_F32, s * slab + static_cast<long long>(r) * N + c, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += t[j];
    }
    if (bias) {
      float t[8];
      ld8_as_float(bias, bias_dt, c, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += t[j];
    }
    if (addend) {
      float t[8];
      ld8_as_float(addend, addend_dt, static_cast<long long>(r) * ld_addend + c, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += t[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= scale;
    st8_from_float(out, out_dt, static_cast<long long>(r) * ld_out + c, acc);
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Per-device pair of flags that a timed-out wait inside a GEMM kernel raises (mbar_wait_soft): one int of device memory
// (what the other waits of the same launch poll) and one int of mapped pinned host memory (what the host reads without
// synchronising). Allocated on first use on whichever device is current, so a process that drives several GPUs never
// hands one device's flag to another device's kernels.
struct PoisonSlot {
  int* dev = nullptr;
  int* host = nullptr;      // host address of the mirror
  int* host_dev = nullptr;  // device address of the mirror
};
static PoisonSlot* poison_slot() {
  static PoisonSlot slots[64];
  static std::mutex mu;
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  PoisonSlot& s = slots[d];
  if (!s.dev) {
    int* dv = nullptr;
    int* h = nullptr;
    int* hd = nullptr;
    if (cudaMalloc(&dv, sizeof(int)) != cudaSuccess) return nullptr;
    if (cudaMemset(dv, 0, sizeof(int)) != cudaSuccess) return nullptr;
    if (cudaHostAlloc(&h, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
    *h = 0;
    if (cudaHostGetDevicePointer(&hd, h, 0) != cudaSuccess) return nullptr;
    s.dev = dv; s.host = h; s.host_dev = hd;
  }
  return &s;
}
static std::atomic<int> g_gemm_launch_id{0};

}  // namespace vy

// development hook (tools/gemm_sweep.py): pin the kernel flavour, tile width and K split of subsequent vy_gemm calls
static int g_force_pair = -1, g_force_bn = 0, g_force_splits = 0;
extern "C" int vy_gemm_tune_override(int pair, int bn, int splits) {
  g_force_pair = pair;
  g_force_bn = bn;
  g_force_splits = splits;
  return VY_OK;
}

namespace vy {
bool skinny_applicable(const VyGemm* p);  // gemm_skinny.cu
int launch_skinny(const VyGemm* p);
}  // namespace vy

extern "C" int vy_gemm_poisoned(void) {
  using namespace vy;
  PoisonSlot* f = poison_slot();
  int v = -1;
  if (!f || cudaMemcpy(&v, f->dev, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (v != 0 || *reinterpret_cast<volatile int*>(f->host) != 0) {  // reading a raised flag lowers it again
    cudaMemset(f->dev, 0, sizeof(int));
    *reinterpret_cast<volatile int*>(f->host) = 0;
    return 1;
  }
  return 0;
}

extern "C" int vy_gemm_poison_peek(void) {
  using namespace vy;
  PoisonSlot* f = poison_slot();
  if (!f) return -1;
  return *reinterpret_cast<volatile int*>(f->host) != 0 ? 1 : 0;
}

extern "C" int vy_gemm(const VyGemm* p) {
  using namespace vy;
  VY_CHECK_ARG(p != nullptr, "vy_gemm: null params");
  if (!vy_device_ok()) {
    set_error("vy_gemm: no sm_100 device (there is no CPU fallback)");
    return VY_ERR_NO_DEVICE;
  }
  // GeGLU = the gated epilogue with gelu_tanh on the gate (Examples/paligemma.ipynb cell 11 GemmaMLP): same code path as SwiGLU
  VyGemm gated_copy;
  int gate_act = 0;
  if (p->act == VY_ACT_GEGLU_TANH) {
    gated_copy = *p;
    gated_copy.act = VY_ACT_SWIGLU;
    p = &gated_copy;
    gate_act = 1;
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
  PoisonSlot* ps = poison_slot();
  VY_CHECK_ARG(ps != nullptr, "vy_gemm: could not allocate the poison flags of this device");
  if (*reinterpret_cast<volatile int*>(ps->host) != 0) {
    set_error("vy_gemm: a barrier wait inside an earlier GEMM kernel on this device timed out (launch id %d) — its results "
              "and everything computed from them are invalid; vy_gemm_poisoned() acknowledges and lowers the flag",
              *reinterpret_cast<volatile int*>(ps->host));
    return VY_ERR_CUDA;
  }
  g.poison.flag = ps->dev;
  g.poison.host = ps->host_dev;
  int lid = g_gemm_launch_id.fetch_add(1, std::memory_order_relaxed) + 1;
  g.poison.id = lid & 0x7fffffff ? lid & 0x7fffffff : 1;
  g.M = p->M; g.N = p->N; g.K = p->K;
  g.epi = p->epi; g.act = p->act; g.transposed_out = p->transposed_out;
  g.gate_act = gate_act;
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
    VY_CHECK_ARG(p->act >= VY_ACT_NONE && p->act <= VY_ACT_SWIGLU, "vy_gemm: bad act %d", p->act);
    if (p->act == VY_ACT_SWIGLU)
      VY_CHECK_ARG((p->N & 1) == 0 && !p->addend && !p->addend2 && !p->transposed_out && p->out_row_group == 0,
                   "vy_gemm: SWIGLU needs an even N (interleaved gate/up rows) and takes no addend / transpose / row remap");
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
    VY_CHECK_ARG(p->kv_cap <= 0 || p->kv_dst_pos0 + p->tokens_per_seq <= p->kv_cap,
                 "vy_gemm: k/v rows [%d, %d) do not fit the %d token slots of k_out / v_out", p->kv_dst_pos0,
                 p->kv_dst_pos0 + p->tokens_per_seq, p->kv_cap);
    VY_CHECK_ARG(p->rope_rows <= 0 || !p->rope_cos || p->start_pos + p->tokens_per_seq <= p->rope_rows,
                 "vy_gemm: positions [%d, %d) exceed the %d rows of the RoPE tables", p->start_pos,
                 p->start_pos + p->tokens_per_seq, p->rope_rows);
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

  // at most 32 activation rows against a K-major weight (the projections of a decode step): weight-streaming kernel of
  // gemm_skinny.cu instead of 128-row tcgen05 tiles
  if (skinny_applicable(p)) return launch_skinny(p);

  // TMA-store write-back for the fast (all-bf16, 16-byte aligned, no row remap) epilogues; VY_GEMM_TMA_STORE=0 keeps st.global
  static const bool tma_store_on = !(getenv("VY_GEMM_TMA_STORE") && atoi(getenv("VY_GEMM_TMA_STORE")) == 0);
  g.tma_store = 0;
  // (SwiGLU: the staged write-back exists for the inference form only — no pre-activation save — and for bias pointers
  //  the fast bias loader can vector-load)
  const bool gated_fast = p->act == VY_ACT_SWIGLU && !p->aux && aligned16(p->bias) && p->in_dtype == VY_BF16 && !p->a_mn_major &&
                          !p->b_mn_major;
  if (tma_store_on && p->epi == VY_EPI_LINEAR && (p->act != VY_ACT_SWIGLU || gated_fast) && !p->transposed_out && g.vec_ok && p->out_dtype == VY_BF16 && !p->addend2 &&
      p->out_row_group == 0 && (!p->aux || p->aux_dtype == VY_BF16) && (!p->addend || p->addend_dtype == VY_BF16))
    g.tma_store = 1;
  const int bk = p->in_dtype == VY_BF16 ? 64 : 32;
  const int num_kb = (p->K + bk - 1) / bk;
  // split-K is offered for plain linear epilogues (what the weight gradients use) when the caller lent scratch space
  int max_splits = 1;
  if (p->workspace && p->epi == VY_EPI_LINEAR && p->act == VY_ACT_NONE && !p->transposed_out && !p->addend2 && !p->aux &&
      p->out_row_group == 0 && p->addend_row_mod == 0 && (p->N & 7) == 0 && g.vec_ok && aligned16(p->workspace) &&
      (!p->bias || aligned16(p->bias))) {
    const long long per = static_cast<long long>(p->M) * p->N * 4;
    max_splits = static_cast<int>(p->workspace_bytes / per < 8 ? p->workspace_bytes / per : 8);
    if (max_splits < 1) max_splits = 1;
  }
  // CTA-pair kernels (cta_group::2): bf16, at least two m-tiles, row-major result. Both flavours are scored by the time
  // model and the cheaper one runs; VY_GEMM_PAIR=0 / 1 pins the flavour.
  static const int pair_env = getenv("VY_GEMM_PAIR") ? atoi(getenv("VY_GEMM_PAIR")) : -1;
  const bool pair_ok = p->in_dtype == VY_BF16 && p->M > 128 && p->N >= 128 && !p->transposed_out;
  // precedence: development override (vy_gemm_tune_override) > environment > the caller's hint > the time model
  int pair_pin = g_force_pair >= 0 ? g_force_pair : pair_env;
  if (pair_pin < 0 && p->hint_flavour == 1) pair_pin = 0;
  if (pair_pin < 0 && p->hint_flavour == 2 && pair_ok) pair_pin = 1;
  const bool mn_major = p->a_mn_major || p->b_mn_major;
  const bool qkv = p->epi == VY_EPI_QKV_ROPE;
  const double epi_cost = p->act == VY_ACT_GELU_ERF || p->act == VY_ACT_GELU_TANH ? 3.7
                          : (p->act == VY_ACT_DGELU_ERF || p->act == VY_ACT_DGELU_TANH ? 6.2 : 2.5);
  Tiling tl = choose_tiling(p->M, p->N, num_kb, mn_major, p->b_mn_major != 0, qkv, epi_cost, max_splits,
                            pair_ok && pair_pin == 1 ? 1 : 0, pair_ok && pair_pin != 0 ? 1 : 0);
  const bool pair = tl.pair != 0;
  static const int force_bn_env = getenv("VY_GEMM_FORCE_BN") ? atoi(getenv("VY_GEMM_FORCE_BN")) : 0;  // development: pin the tile width
  int bn = tl.bn;
  if (g_force_bn || force_bn_env) {
    bn = g_force_bn ? g_force_bn : force_bn_env;
  } else if (p->hint_bn) {
    const int hb = p->hint_bn;
    const bool known = hb == 32 || hb == 64 || hb == 128 || hb == 192 || hb == 256;
    const bool fits = (!mn_major || hb >= 128) && (!qkv || hb >= 64) && (!pair || hb >= 128) &&
                      !(pair && p->b_mn_major && hb == 192) && (hb == 32 || hb / 2 < p->N);
    if (known && fits) {
      bn = hb;
      if (p->hint_splits == 0) tl.splits = 1;  // the model's split belonged to the model's width
    }
  }
  if (pair && bn < 128) bn = 128;
  if (pair && p->b_mn_major && bn == 192) bn = 256;
  if (p->act == VY_ACT_SWIGLU && g.tma_store && (bn == 192 || bn < 128)) bn = bn == 192 ? 256 : 128;  // whole chunk pairs per warp
  auto split_ok = [&](int sp) {
    return sp >= 1 && sp <= max_splits && (sp == 1 || static_cast<long long>(sp - 1) * ((num_kb + sp - 1) / sp) < num_kb);
  };
  if (g_force_splits > 0) {
    VY_CHECK_ARG(split_ok(g_force_splits), "vy_gemm: forced split %d not possible here (max %d, %d k-blocks)", g_force_splits,
                 max_splits, num_kb);
    tl.splits = g_force_splits;
  } else if (p->hint_splits > 0 && split_ok(p->hint_splits)) {
    tl.splits = p->hint_splits;
  }
  g.k_splits = tl.splits;
  g.kb_per_split = (num_kb + tl.splits - 1) / tl.splits;
  g.ws = static_cast<float*>(p->workspace);
  if (tl.splits > 1) {
    const int rc = pair ? dispatch_gemm_pair(p, g, bn)
                        : (p->in_dtype == VY_BF16 ? dispatch_gemm<__nv_bfloat16>(p, g, bn) : dispatch_gemm<float>(p, g, bn));
    if (rc != VY_OK) return rc;
    const long long nvec = static_cast<long long>(p->M) * (p->N >> 3);
    long long blocks = (nvec + 255) / 256;
    if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
    VY_CUDA_OK(launch_kernel(splitk_reduce_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(p->stream), 
        p->M, p->N, tl.splits, g.ws, p->bias, p->bias_dtype, p->addend, p->ld_addend, p->addend_dtype,
        p->out_scale == 0.f ? 1.f : p->out_scale, p->out, p->ld_out, p->out_dtype));
    VY_LAUNCH_OK();
    count_launch();
    return VY_OK;
  }

  if (p->act == VY_ACT_SWIGLU && g.tma_store) return dispatch_gemm_gated(p, g, bn, pair);
  if (pair) return dispatch_gemm_pair(p, g, bn);
  if (p->in_dtype == VY_BF16) return dispatch_gemm<__nv_bfloat16>(p, g, bn);
  return dispatch_gemm<float>(p, g, bn);
}
