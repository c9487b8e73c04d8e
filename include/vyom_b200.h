/* vyom_b200.h — C ABI of the B200-native transformer-block hot path (libvyom_b200.so).
 *
 * The reference (Ajax0564/VyomAI) has no FFI of its own: its hot path is a handful of PyTorch
 * calls inside VyomAI/layers/*.py. Each entry point below replaces one of those call sites; the
 * comment above every declaration cites the reference lines it stands in for (paths relative to
 * the reference checkout). INTEGRATION.md shows the ctypes binding a maintainer adds on the
 * reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer on the current device
 *     unless the name ends in _host; the caller owns all buffers (outputs and workspaces);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 on success, a negative VY_ERR_* otherwise; vy_last_error() returns a
 *     thread-local message for the last failure on the calling thread; nothing throws;
 *   - dtype arguments take VY_F32 or VY_BF16. All reductions / softmax / norm statistics are
 *     fp32. Tensor-core contractions use bf16 inputs (kind::f16) or tf32 (kind::tf32, for fp32
 *     tensors) with fp32 accumulation in TMEM.
 *   - there is no CPU path: on a machine without an sm_100 GPU every compute entry point
 *     returns VY_ERR_NO_DEVICE.
 */
#ifndef VYOM_B200_H_
#define VYOM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VY_ABI_VERSION 7

#if defined(__GNUC__)
#define VY_API __attribute__((visibility("default")))
#else
#define VY_API
#endif

enum { VY_F32 = 0, VY_BF16 = 1 };

enum {
  VY_OK = 0,
  VY_ERR_INVALID_ARG = -1,
  VY_ERR_CUDA = -2,
  VY_ERR_NO_DEVICE = -3,
  VY_ERR_UNSUPPORTED = -4
};

/* activation selector of the GEMM epilogue (VyomAI/layers/ffn.py:7-15 `_ACT_`, exact-erf GELU is
 * the default; the tanh form is what the PaliGemma notebook's SigLIP/Gemma MLPs use). The D*
 * variants multiply the accumulator by act'(aux) and are used by the backward pass. */
enum {
  VY_ACT_NONE = 0,
  VY_ACT_GELU_ERF = 1,
  VY_ACT_GELU_TANH = 2,
  VY_ACT_DGELU_ERF = 3,
  VY_ACT_DGELU_TANH = 4,
  VY_ACT_SWIGLU = 5, /* gated MLP: see VyGemm.act */
  VY_ACT_GEGLU_TANH = 6 /* same layout and rules as VY_ACT_SWIGLU with out = gelu_tanh(gate) * up (GemmaMLP, Examples/paligemma.ipynb cell 11) */
};

enum { VY_EPI_LINEAR = 0, VY_EPI_QKV_ROPE = 1 };

VY_API int vy_version(void);
VY_API const char* vy_last_error(void);
/* number of kernels this library has launched from the calling process (all threads) */
VY_API int64_t vy_launch_count(void);
/* Programmatic dependent launch of this library's kernels (each one triggers its dependents at entry and waits for its
 * predecessor before touching global memory): on != 0 enables, 0 disables, negative returns to the process default (the
 * VY_PDL environment variable, off). Returns the previous override (-1 = none). Affects launches made after the call, by
 * every thread of the process: meant to bracket a CUDA-graph capture (vyomai_b200/decode_graph.py), not to be toggled
 * around individual calls while other threads launch. */
VY_API int vy_set_pdl(int on);
/* 1 if the current device is sm_100 (B200), else 0 */
VY_API int vy_device_ok(void);
/* sizeof() of the parameter struct called `name` ("VyGemm", ...) as compiled into the library, or -1;
 * bindings compare it with their own layout so a stale .so / header pair fails at load time. */
VY_API int vy_abi_sizeof(const char* name);

/* ------------------------------------------------------------------------------------------
 * vy_gemm — tcgen05/TMEM GEMM fed by TMA with a fused epilogue.
 *   acc[m,n] = sum_k A[m,k] * B[n,k]                      (fp32 accumulation in TMEM)
 * replaces nn.Linear at VyomAI/layers/attention.py:87-95,114-116,161-173,190-192,587,607 (q/k/v
 * and fused qkv projections), attention.py:47-51,69 (output projection), ffn.py:21,30,33-37
 * (intermediate + GELU + out), models/decoder.py:258-275 (LM head) and the stride==kernel
 * nn.Conv2d patch embedding at models/vision_encoder.py:83-88,114-115 (as a GEMM over patch
 * rows), plus their autograd backward (dgrad uses B MN-major, wgrad uses both MN-major).
 *
 * Operand storage: K-major means the K index is contiguous (A is [M][K] with row stride lda,
 * B is [N][K] with row stride ldb — exactly an nn.Linear weight). MN-major means the operand is
 * stored transposed ([K][M] / [K][N]) and ld* is the stride between consecutive k rows.
 * Strides are in elements; base pointers 16-byte aligned; ld* * sizeof(dtype) % 16 == 0.
 *
 * VY_EPI_LINEAR epilogue, for logical output element (r,c) (= (m,n), or (n,m) if
 * transposed_out, which is how the small-batch decode GEMMs run "swap-AB"):
 *   x = acc + bias[c]
 *   if aux && act in {GELU_*}:  aux[r,c] = x        (pre-activation saved for backward)
 *   x = act(x)                  (D* variants: x = x * act'(aux[r,c]))
 *   if addend: x += addend[ar, c],  ar = addend_row_mod ? addend_row_off + r % addend_row_mod : r
 *   if addend2: x += addend2[r, c]       (backward: the two residual gradients of a layer)
 *   out[orow, c] = out_scale * x,
 *       orow = out_row_group ? (r / out_row_group) * out_row_group_stride + r % out_row_group
 *                              + out_row_off : r
 * (residual add of attention.py:71 / ffn.py:39 = addend; ViT "2*(patch + pos)" of
 *  vision_encoder.py:125-127 + positional_embeddings.py:222-226 = addend with row_mod, out_scale
 *  2 and the row-group remap that leaves room for the cls row.)
 *
 * VY_EPI_QKV_ROPE epilogue (head_dim must be 64): the N axis is [q heads | k heads | v heads];
 * row r is token (b = r / tokens_per_seq, l = r % tokens_per_seq) at position start_pos + l.
 * Adds bias, rotates q and k heads in registers with the half-split RoPE of
 * positional_embeddings.py:140-182 using rope_cos/rope_sin[pos][j] (fp32 tables of d/2 columns,
 * pass NULL for absolute/sinusoidal models), and scatters straight into q_out[b,h,l,:],
 * k_out[b,hk,kv_dst_pos0+l,:], v_out[...] through the given element strides — i.e. the
 * "b l (h d) -> b h l d" rearrange (attention.py:118-120) and the kv-cache append
 * (kv_cache.py:355-359) are fused into the projection. k/v rows are written at token index
 * kv_dst_pos0 + l (= start_pos for a cache, 0 for a fresh buffer); q rows at l.
 *
 * Small batches (transposed_out, N <= 32 activation rows, bf16, K-major operands, K / s a multiple of 128 for some s <= 8, at most ~9.4k output features, act NONE / GELU_*, no aux / addend2 / row remaps) do not use the tensor-memory kernel at
 * all: csrc/gemm_skinny.cu streams the weights once through 150-400 CTAs (16 features x K / s each, mma.sync, partial sums
 * combined through the shared memory of a thread-block cluster). VY_GEMM_SKINNY=0 in the environment disables it.
 * ------------------------------------------------------------------------------------------ */
typedef struct VyGemm {
  int32_t M, N, K;
  int32_t in_dtype; /* dtype of A and B */
  const void* A;
  int64_t lda;
  int32_t a_mn_major;
  const void* B;
  int64_t ldb;
  int32_t b_mn_major;

  int32_t epi; /* VY_EPI_* */
  int32_t act; /* VY_ACT_*. VY_ACT_SWIGLU: B holds the gate and up projections INTERLEAVED (row 2j = gate_j, row 2j+1 =
                * up_j; N = 2 * intermediate), out has N / 2 columns: out[r, j] = silu(z[r, 2j]) * z[r, 2j+1] with z = A B^T
                * + bias — `down_proj(act_fn(gate_proj(x)) * up_proj(x))`'s first half as ONE GEMM
                * (VyomAI/models/custom_transformer.py:76-89; Examples/simple_vllm.ipynb FeedForward). aux (optional)
                * receives z ([rows, N]) for vy_swiglu_bwd. No addend / transposed_out / split-K with it. */
  int32_t transposed_out;
  const void* bias; /* [cols] or NULL */
  int32_t bias_dtype;
  const void* addend; /* or NULL */
  int64_t ld_addend;
  int32_t addend_dtype;
  int32_t addend_row_mod, addend_row_off;
  const void* addend2; /* second residual stream (same row index as the output row), or NULL */
  int64_t ld_addend2;
  int32_t addend2_dtype;
  void* aux; /* pre-activation: written by GELU_*, read by DGELU_*; or NULL */
  int64_t ld_aux;
  int32_t aux_dtype;
  float out_scale; /* 0 is treated as 1 */
  void* out;
  int64_t ld_out;
  int32_t out_dtype;
  int32_t out_row_group, out_row_group_stride, out_row_off;

  /* VY_EPI_QKV_ROPE only */
  int32_t tokens_per_seq, start_pos, kv_dst_pos0, head_dim, n_q_heads, n_kv_heads;
  const float* rope_cos; /* [>= start_pos + tokens_per_seq][head_dim/2] */
  const float* rope_sin;
  void* q_out;
  int64_t q_sb, q_sh, q_sl; /* element strides of batch, head, token; head_dim contiguous */
  void* k_out;
  int64_t k_sb, k_sh, k_sl;
  void* v_out;
  int64_t v_sb, v_sh, v_sl;
  int32_t kv_out_dtype; /* dtype of k_out / v_out (a kv-cache may be fp32 while q_out is bf16) */
  int32_t kv_cap;       /* token slots of k_out / v_out (0 = unchecked): kv_dst_pos0 + tokens_per_seq must not exceed it —
                         * the reference fails on the slice assignment (kv_cache.py:355-359), this call returns
                         * VY_ERR_INVALID_ARG instead of writing past the cache */
  int32_t rope_rows;    /* rows of rope_cos / rope_sin (0 = unchecked): start_pos + tokens_per_seq must not exceed it */

  /* optional split-K scratch (fp32). When a GEMM has too few output tiles to fill the SMs and a long K (the
   * weight gradients dW = dY^T X, K = tokens), the library splits K over several CTAs per tile, writes fp32
   * partial tiles here and finishes with a reduce kernel that applies bias / addend / scale. It uses as many
   * splits as fit in workspace_bytes (>= splits * M * N * 4); NULL / 0 disables split-K. */
  void* workspace;
  int64_t workspace_bytes;

  /* optional tiling hints (0 = let the library's time model decide). A caller that has timed the candidates on its own
   * shapes (vyomai_b200/gemm_tune.py does, once per shape) passes the winner here; a hint that does not apply to the
   * call (pair kernels need bf16 and M > 128, widths below 128 need K-major operands, a split needs workspace and must
   * not leave an empty slab) is ignored. */
  int32_t hint_flavour; /* 1 = single-CTA 128 x BN tiles, 2 = CTA-pair (cta_group::2) 256 x BN tiles */
  int32_t hint_bn;      /* 32, 64, 128, 192 or 256 */
  int32_t hint_splits;  /* K splits, 1..8 */

  /* Non-zero: the K-major weight operand (A in the swap-AB / transposed_out form) is not written by any kernel still in
   * flight on this stream. The small-batch kernel (N <= 32 activation rows, bf16) then requests the weights BEFORE it
   * waits for its predecessor under programmatic dependent launch (vy_set_pdl). Leave 0 when an optimizer step may
   * immediately precede the call. */
  int32_t weights_static;

  void* stream;
} VyGemm;

VY_API int vy_gemm(const VyGemm* p);
/* 1 if vy_gemm would serve `p` with the small-batch weight-streaming kernel (tiling hints do not apply to it), else 0 */
VY_API int vy_gemm_is_small_batch(const VyGemm* p);
/* Self-check of the GEMM kernels' barrier protocol: a wait inside a kernel that times out raises a device flag instead
 * of faulting, and the launch finishes with undefined results. Returns the flag (0 = every vy_gemm so far ran its
 * protocol to completion, 1 = some launch did not, -1 = the flag could not be read); a raised flag is lowered by the
 * read. Synchronises the device. */
VY_API int vy_gemm_poisoned(void);
/* Same flag without synchronising (a pinned host mirror the kernels write on a timeout): 1 once a GEMM of the current
 * device has timed out and nobody has acknowledged it through vy_gemm_poisoned(), else 0. While it is raised every
 * vy_gemm call on that device fails with VY_ERR_CUDA, so a corrupted result cannot feed later work silently. */
VY_API int vy_gemm_poison_peek(void);
/* Development hook (tools/gemm_sweep.py): pin the kernel flavour (pair: -1 auto, 0 single CTA, 1 CTA pair), the tile
 * width (bn: 0 auto) and the K split (splits: 0 auto) of subsequent vy_gemm calls of this process. */
VY_API int vy_gemm_tune_override(int pair, int bn, int splits);

/* ------------------------------------------------------------------------------------------
 * vy_add_layernorm_fwd / _bwd — y = LayerNorm(x + residual) * gamma + beta, warp per row.
 * replaces `self.layernorm(hidden_states + input_tensor)` at VyomAI/layers/attention.py:52-54,71
 * and VyomAI/layers/ffn.py:25,39 (biased variance, eps = config.layer_norm_eps), and the plain
 * LayerNorm of the LM head (models/decoder.py:259-261,270; residual = NULL).
 * mean / rstd (fp32, [rows]) are optional outputs of fwd and required inputs of bwd.
 * bwd: dx[rows,H] (= d residual as well), dgamma/dbeta partials are reduced into dgamma[H], dbeta[H]
 * (fp32 or bf16; overwritten, or accumulated into when dparam_accumulate); dbias (optional) receives the column
 * sums of dx, i.e. the gradient of the bias of the Linear whose output x was (attention.py:69, ffn.py:37). xhat is recomputed from the saved pre-norm sum `s`
 * (s = x + residual; pass the tensor fwd wrote to sum_out, or x when residual was NULL).
 * ------------------------------------------------------------------------------------------ */
#define VY_NORM_LAYER 0
#define VY_NORM_RMS 1
#define VY_NORM_RMS_GEMMA 2

typedef struct VyNorm {
  int32_t rows, H;
  const void* x;
  const void* residual; /* or NULL */
  int32_t io_dtype;     /* dtype of x, residual, y, sum_out, dy, dx */
  const void* gamma;
  const void* beta;
  int32_t param_dtype;
  float eps;
  void* y;
  void* sum_out; /* optional: x + residual (needed by bwd when residual != NULL) */
  float* mean;   /* optional [rows] */
  float* rstd;   /* optional [rows] */
  /* bwd only */
  const void* dy;
  const void* s;  /* pre-norm sum saved by fwd */
  void* dx;
  void* dgamma; /* [H], dtype dparam_dtype */
  void* dbeta;  /* [H], dtype dparam_dtype */
  void* dbias;  /* optional [H]: column sums of dx = the bias gradient of the Linear that produced x */
  int32_t dparam_dtype;      /* VY_F32 (default 0) or VY_BF16 */
  int32_t dparam_accumulate; /* 1: dgamma/dbeta += result (accumulate straight into the parameters' .grad) */
  float* partials; /* workspace: 2 * vy_norm_bwd_partial_rows() * H floats (holds the per-strip column partials) */
  /* VY_NORM_LAYER (0, default): LayerNorm as above. VY_NORM_RMS: y = gamma * xhat, xhat = x * rsqrt(mean(x^2) + eps),
   * with xhat rounded to the activation dtype before the multiply (VyomAI/models/custom_transformer.py:227-241;
   * Examples/simple_vllm.ipynb cell 2 RMSNorm, whose optional `shift` is `beta`). VY_NORM_RMS_GEMMA: y = (1 + gamma) *
   * xhat (Examples/paligemma.ipynb GemmaRMSNorm). For both, beta may be NULL, mean is unused, and bwd's dbeta
   * receives the column sums of dy (meaningful only when a shift is trained). */
  int32_t kind;
  /* Dropout on x before the residual add — `self.dropout(hidden_states)` at VyomAI/layers/attention.py:70 and
   * VyomAI/layers/ffn.py:38, live in .train() with hidden_dropout_prob (0.1 in every reference config):
   *   y = Norm(keep * x / (1 - p) + residual),  keep ~ Bernoulli(1 - p) per element.
   * The mask is never stored: element e of the call is kept iff a 16-bit lane of Philox4x32-10(key = dropout_seed,
   * counter = (e / 8, dropout_offset, step)) is >= round(p * 65536), so bwd regenerates it from the same
   * (seed, offset, step). `step` is *dropout_step_ptr (device int32, e.g. the trainer's step counter, so a captured
   * CUDA graph draws a fresh mask on every replay) or 0. dropout_p == 0 disables all of it.
   * bwd: dx (= d residual) is the gradient of the pre-norm sum; dx_drop (required when dropout_p > 0) receives
   * keep * dx / (1 - p) = the gradient of x, and dbias is then its column sum. */
  float dropout_p;
  uint64_t dropout_seed;
  uint32_t dropout_offset;
  const int32_t* dropout_step_ptr;
  void* dx_drop;
  void* stream;
} VyNorm;

VY_API int vy_add_layernorm_fwd(const VyNorm* p);
VY_API int vy_add_layernorm_bwd(const VyNorm* p);
VY_API int vy_norm_bwd_partial_rows(void);

/* ------------------------------------------------------------------------------------------
 * vy_attn_fwd — flash-style fused attention forward (tcgen05 + TMEM + TMA), Sq >= 1.
 * replaces repeat_kv + F.scaled_dot_product_attention(q, k, v, attn_mask) + "b h l d -> b l (h d)"
 * at VyomAI/layers/attention.py:128-132 (EncoderAttention), :205-213 (EncoderAttentionGqa),
 * :283-287, :368-377 (DecoderAttention[Gqa]), :464-468, :563-571 (cross attention), :619-623
 * (VisionAttention) and VyomAI/models/decoder.py:107-111,190-199.
 *   out[b, l, h*64 + :] = softmax_k( q[b,h,l,:] . k[b,h/n_rep,k,:] / 8 + M[b,l,k] ) v[b,h/n_rep,k,:]
 * q/k/v are bf16 [B, heads, S, 64] through element strides (sb, sh, sl; head_dim contiguous) — so
 * k/v may point straight into a kv-cache. The additive float mask of the reference is passed in
 * factored form: key_padding_mask[b,k] (uint8, 1 = visible; the `attention_mask` of
 * models/encoder.py:161-164) and `causal` with q_pos0 = start_pos (key k visible to query l iff
 * k <= q_pos0 + l; models/decoder.py:376-419). Masked scores behave like "+ finfo.min": a row with
 * no visible key returns the mean of v over ALL Skv keys.
 * lse (optional, fp32 [B, n_q_heads, Sq]) receives log2-domain logsumexp for vy_attn_bwd.
 * head_dim 64 runs on the tensor-memory kernel described above. Any other head_dim that is a multiple of 8 up to 256
 * (SigLIP 72, Gemma 256: Examples/paligemma.ipynb cells 9, 12) and every call with a prefix-LM mask runs on a second,
 * mma.sync-based forward kernel (csrc/attn_fwd_mma.cu; "64" in the layout notes above reads head_dim, the scale is
 * 1/sqrt(head_dim)); at Sq == 1 it packs the query heads of a kv group into one tile (MQA / GQA decode over a cache).
 * vy_attn_bwd exists for head_dim 64 only. The packed-head decode form splits the keys over several CTAs when there are few
 * (row, kv head) pairs and combines them through a library-owned scratch buffer: such calls must not run concurrently on two
 * streams of one device (like vy_sqnorm / vy_add_layernorm_bwd's partials).
 * ------------------------------------------------------------------------------------------ */
typedef struct VyAttn {
  int32_t B, n_q_heads, n_kv_heads, head_dim;
  int32_t Sq, Skv;
  int32_t qkv_dtype; /* VY_BF16 */
  const void* q;
  int64_t q_sb, q_sh, q_sl;
  const void* k;
  int64_t k_sb, k_sh, k_sl;
  const void* v;
  int64_t v_sb, v_sh, v_sl;
  int32_t causal, q_pos0;
  const uint8_t* key_padding_mask; /* [B, Skv] or NULL */
  int64_t kpm_stride;
  void* out; /* [B, Sq, n_q_heads*64] through o_sb, o_sl */
  int64_t o_sb, o_sl;
  int32_t out_dtype;
  float* lse;
  /* PREFIX_LM mask (with causal != 0): int32 [B], key k is visible to query l iff k <= q_pos0 + l OR k < prefix_len[b] — the
   * training mask of Examples/paligemma.ipynb cell 17 `_update_causal_mask` (image + prompt tokens attend to each other in
   * both directions, the suffix is causal). NULL = plain causal. */
  const int32_t* prefix_len;
  /* Device-side position (mma.sync kernel only; NULL = off): the keys are [0, min(Skv, *pos_ptr + Sq)) and q_pos0 = *pos_ptr —
   * Skv is then the capacity of the cache the k / v pointers address, so one captured graph serves every decode step. */
  const int32_t* pos_ptr;
  void* stream;
} VyAttn;

VY_API int vy_attn_fwd(const VyAttn* p);

/* ------------------------------------------------------------------------------------------
 * vy_attn_bwd — flash-attention backward (tcgen05 + TMEM + TMA): the autograd of the SDPA call
 * sites listed under vy_attn_fwd, of repeat_kv (dk/dv summed over the query heads of a group) and of
 * apply_rotary_pos_emb (the rotation is undone on dq / dk before they are written).
 * Inputs are the forward's bf16 q/k/v (post-RoPE, [B, heads, S, 64] via strides), its output o
 * ([B, Sq, Hq*64] via o_sb/o_sl, any dtype), the bf16 output gradient dout (same layout through
 * do_sb/do_sl) and lse from vy_attn_fwd. dsum is a [B, Hq, Sq] fp32 workspace. Results go to
 * dq [B*Sq, >= Hq*64] (row stride ld_dq, head h at column h*64), dk / dv [B*Skv, >= Hkv*64] — pass
 * three column offsets of one packed [tokens, (Hq+2Hkv)*64] buffer to get the gradient of the
 * fused q|k|v projection directly. rope_cos/rope_sin are the forward's tables (NULL = no RoPE),
 * rope_pos0 the table row of position 0 (normally 0).
 * ------------------------------------------------------------------------------------------ */
typedef struct VyAttnBwd {
  int32_t B, n_q_heads, n_kv_heads, head_dim;
  int32_t Sq, Skv;
  const void* q;
  int64_t q_sb, q_sh, q_sl;
  const void* k;
  int64_t k_sb, k_sh, k_sl;
  const void* v;
  int64_t v_sb, v_sh, v_sl;
  const void* o;
  int64_t o_sb, o_sl;
  int32_t o_dtype;
  const void* dout; /* bf16 */
  int64_t do_sb, do_sl;
  const float* lse;
  float* dsum;
  int32_t causal, q_pos0;
  const uint8_t* key_padding_mask;
  int64_t kpm_stride;
  const float* rope_cos;
  const float* rope_sin;
  int32_t rope_pos0;
  void* dq;
  int64_t ld_dq;
  void* dk;
  int64_t ld_dk;
  void* dv;
  int64_t ld_dv;
  int32_t out_dtype;
  void* stream;
} VyAttnBwd;

VY_API int vy_attn_bwd(const VyAttnBwd* p);

/* vy_rope_apply — stand-alone half-split RoPE for the public apply_rotary_pos_emb(q, k, freqs)
 * helper (layers/positional_embeddings.py:155-182): out[b,h,l,:] = rotate(x[b,h,l,:], angle row
 * pos0 + l). cos/sin are fp32 [rows][head_dim / 2]; inverse = 1 applies the transpose rotation. Any even head_dim (64 in the
 * package's models, 256 in the Gemma decoder of Examples/paligemma.ipynb cell 11, whose `out` is the kv-cache slot itself). */
typedef struct VyRope {
  int32_t B, H, S, head_dim;
  const void* x;
  int64_t x_sb, x_sh, x_sl;
  int32_t dtype;
  const float* cos;
  const float* sin;
  int32_t pos0, inverse;
  void* out;
  int64_t o_sb, o_sh, o_sl;
  /* Device-side position (all optional, zero = off) so that ONE captured CUDA graph serves every decode step:
   * pos_ptr: int32 scalar added to pos0; out_follows_pos: token l is written at out token index l + *pos_ptr (out = base of a
   * kv-cache: the append); copy_only: no rotation, a plain strided copy (values go into the cache unrotated). */
  const int32_t* pos_ptr;
  int32_t out_follows_pos, copy_only;
  void* stream;
} VyRope;

VY_API int vy_rope_apply(const VyRope* p);

/* vy_rope_append — apply_rotary_pos_emb(q, k) + StaticCache.update(k, v) of a PACKED q|k|v projection in one launch
 * (Examples/paligemma.ipynb cell 12 GemmaAttention.forward + cell 28): `qkv` is addressed as [B, n_q + 2 n_kv, S, head_dim]
 * through (sb, sh, sl); query heads are rotated in place, key heads rotated into k_cache, value heads copied into v_cache
 * ([B, n_kv, slots, head_dim] through c_sb, c_sh, c_sl) at slot slot0 + l; token l uses table row pos0 + l (PaliGemma: slot + 1).
 * pos_ptr (device int32, optional) is added to both pos0 and slot0, so a captured graph serves every decode step; cache_slots /
 * rope_rows (0 = unchecked) bound the host-known part. cos / sin: fp32 [rows][head_dim / 2]. */
typedef struct VyRopeAppend {
  int32_t B, S, n_q_heads, n_kv_heads, head_dim;
  void* qkv;
  int64_t sb, sh, sl;
  int32_t dtype; /* of qkv and the caches */
  const float* cos;
  const float* sin;
  int32_t pos0, slot0;
  const int32_t* pos_ptr;
  void* k_cache;
  void* v_cache;
  int64_t c_sb, c_sh, c_sl;
  int32_t cache_slots, rope_rows;
  void* stream;
} VyRopeAppend;

VY_API int vy_rope_append(const VyRopeAppend* p);

/* vy_act_bwd — out[i] = dy[i] * act'(z[i]) (act = VY_ACT_GELU_ERF / VY_ACT_GELU_TANH): the GELU
 * backward of the LM head (models/decoder.py:269), whose upstream gradient comes out of a LayerNorm
 * backward rather than a GEMM (inside the FFN the same factor rides in the dgrad GEMM epilogue). */
VY_API int vy_act_bwd(int64_t n, const void* dy, const void* z, int dtype, int act, void* out, void* stream);

/* vy_scale_by_ptr — x[i] *= *scale (device scalar, fp32) over a contiguous buffer; returns immediately on the device when
 * *scale == 1. Applies an arbitrary upstream gradient to the d logits the fused LM-head + cross-entropy node has already
 * written (loss.backward() passes 1, a scaled or accumulated loss does not). */
VY_API int vy_scale_by_ptr(int64_t n, void* x, int dtype, const float* scale, void* stream);

/* vy_slot_merge_fwd / _bwd — `inputs_embeds.masked_scatter(input_ids == image_token_index, image_features)` of the notebook-II
 * captioner (Examples/vyom-ai-accelerate-multimodel-2t4.ipynb cell 1, VisionLanguageModel.forward) and its autograd:
 *   fwd: out[r, :] = slot[r] >= 0 ? b[slot[r], :] : a[r, :]         (a: word-embedding rows, b: image-feature rows)
 *   bwd: da[r, :] = slot[r] >= 0 ? 0 : dout[r, :];   db[slot[r], :] = dout[r, :]     (da or db may be NULL)
 * slot[r] is the running count of image tokens before row r (masked_scatter consumes source rows in row-major order) or
 * -1; every b row is used at most once, rows of db no slot points at are left untouched (zero them first). All buffers
 * are row-contiguous [rows or n_b, H]; H * sizeof(dtype) % 16 == 0. A slot index >= n_b (more image positions than source
 * rows: the reference raises) is treated as -1, so b / db are never addressed out of bounds. */
VY_API int vy_slot_merge_fwd(int rows, int H, int dtype, const void* a, const void* b, int n_b, const int32_t* slot, void* out, void* stream);
VY_API int vy_slot_merge_bwd(int rows, int H, int dtype, const void* dout, const int32_t* slot, void* da, void* db, int n_b, void* stream);

/* vy_swiglu_bwd — gradient of h[r, j] = silu(z[r, 2j]) * z[r, 2j+1] w.r.t. the interleaved pre-activations:
 * dz[r, 2j] = dh[r, j] * z[r, 2j+1] * silu'(z[r, 2j]), dz[r, 2j+1] = dh[r, j] * silu(z[r, 2j]); dz then feeds the
 * ordinary dgrad / wgrad GEMMs against the interleaved weight. dh [rows, I], z / dz [rows, 2 I], contiguous, I % 8 == 0. */
VY_API int vy_swiglu_bwd(int64_t rows, int32_t inter, const void* dh, const void* z, int dtype, void* dz, void* stream);

/* ------------------------------------------------------------------------------------------
 * vy_attn_decode — single-token attention over the contiguous kv-cache, fused with the new
 * token's RoPE and cache append. Replaces, for seqlen == 1, the chain
 *   apply_rotary_pos_emb (layers/positional_embeddings.py:155-182)
 *   -> kv_cache.update(layer, k, v, start_pos)  (layers/kv_cache.py:323-361, 198-236; per-layer
 *      variants :114-146, :31-59)
 *   -> repeat_kv (layers/attention.py:8-19) -> F.scaled_dot_product_attention(mask=None)
 *      (models/decoder.py:105-109,188-197; layers/attention.py:281-285,366-375)
 *   -> "b h l d -> b l (h d)" (attention.py:287,377).
 * qkv is the packed projection output [B, (n_q + 2 n_kv) * 64] (bias added, not rotated).
 * The cache is [B, n_kv, cache_len, 64] through element strides (cache_sb, cache_sh, cache_sl);
 * slots [0, start_pos) must already hold the context; slot start_pos is written by this call
 * (rotated k, raw v — what the reference stores) and attention covers [0, start_pos]. No mask is
 * applied (the reference passes mask=None at decode). head_dim must be 64.
 * splits = 0 lets the library choose (vy_attn_decode_splits); workspace must hold
 * B * n_kv * splits * (n_q/n_kv) * 66 floats, tickets B * n_kv zero-initialised uint32 (the kernel
 * leaves them zero again).
 * ------------------------------------------------------------------------------------------ */
typedef struct VyDecode {
  int32_t B, n_q_heads, n_kv_heads, head_dim;
  int32_t start_pos, cache_len;
  const int32_t* start_pos_ptr; /* optional device copy of start_pos (wins over the host value; start_pos is then
                                   only the upper bound used to size the kv-split): a captured CUDA graph of one
                                   decode step can be replayed for every token */
  const void* qkv;
  int64_t ld_qkv;
  int32_t qkv_dtype;
  const float* rope_cos; /* fp32 [rope_rows][32] or NULL; the new token's angles are row (position + rope_pos_off) */
  const float* rope_sin;
  int32_t rope_pos_off;  /* table row of cache slot 0 (0 when the tables start at position 0; negative offsets are how
                            a caller that only holds the angle rows of the CURRENT positions addresses them) */
  int32_t rope_rows;     /* rows of the tables (0 = unchecked); start_pos + rope_pos_off must lie inside */
  void* k_cache;
  void* v_cache;
  int64_t cache_sb, cache_sh, cache_sl;
  int32_t cache_dtype;
  /* continuous batching / paged cache (Examples/simple_vllm.ipynb cell 2: PagedKVManager, the slot_mapping write
   * `k_cache[slots // block_size, slots % block_size] = k` and flash_attn_with_kvcache(cache_seqlens, block_table)):
   *   seqlens     optional device int32 [B]: row b's context length = the position of its new token (wins over
   *               start_pos / start_pos_ptr, which then only bound the kv-split); a negative entry skips the row.
   *   block_table optional device int32 [B, max_blocks_per_seq]: the caches are then [num_blocks, block_size, n_kv, 64]
   *               pools — cache_sb is the BLOCK stride, cache_sl the slot stride, cache_sh the head stride — and slot p
   *               of row b lives in block block_table[b][p / block_size] at offset p % block_size. The new token is
   *               appended there too; its block must already be in the table. */
  const int32_t* seqlens;
  const int32_t* block_table;
  int32_t max_blocks_per_seq, block_size;
  void* out; /* [B, n_q * 64] */
  int64_t ld_out;
  int32_t out_dtype;
  int32_t splits;
  float* workspace;
  uint32_t* tickets;
  void* stream;
} VyDecode;

VY_API int vy_attn_decode(const VyDecode* p);
VY_API int vy_attn_decode_splits(int B, int n_kv_heads, int start_pos);
/* kv-splits vy_attn_decode will use for `p` when p->splits == 0 (so that the caller can size workspace: B * n_kv * splits *
 * n_rep * 66 floats, and tickets: B * n_kv). bf16 caches with contiguous slots (cache_sl == 64, no block table) take the
 * copy-engine kernel, whose split policy differs from vy_attn_decode_splits' (which the other layouts keep). */
VY_API int vy_attn_decode_plan(const VyDecode* p);

/* ------------------------------------------------------------------------------------------
 * vy_decode_step — ONE kernel launch per generated token for DecoderModel (bf16): the body of the reference's greedy
 * loop (VyomAI/models/decoder.py:470-513 -> forward :324-374 -> DecoderLayer.forward :222-250 -> LMHead :267-275 ->
 * topk(1) :489-496) with a single-token input and a whole-model static kv-cache (layers/kv_cache.py:255-361):
 *   x = word_embeddings[tok] (+ position_embeddings[pos]);  per layer: q|k|v projection, RoPE at `pos`, cache append at
 *   slot `pos`, attention over slots [0, pos] WITHOUT a mask (quirk Q3), LN(dense(.) + x), LN(W2 gelu(W1 .) + x) (the FFN
 *   residual is the layer input, quirk Q2);  logits = decoder(LN(gelu(dense(x)))); next = first index of the row maximum.
 * Persistent: one CTA per SM, stages separated by grid-wide barriers, weights and cache streamed from HBM once, every
 * intermediate kept in an L2-resident scratch (`workspace`). On return (stream order) tok[b] holds the next token,
 * tokens_out[b][pos + 1] too (if given), and *pos has been incremented — so a captured CUDA graph of this ONE launch
 * replays the whole greedy loop. Constraints: bf16 weights / caches, batch <= 32, head_dim 64, H = 64 * n_q_heads a
 * multiple of 256 (<= 1024), ffn a multiple of 256 (<= 4096). q|k|v weights / biases of a layer are one packed matrix
 * [ (n_q + 2 n_kv) * 64, H ] (rows q | k | v). `workspace`: vy_decode_step_workspace_bytes() bytes, 256-byte aligned,
 * ZEROED once by the caller before first use (it holds the barrier counters) and private to one stream.
 * vy_decode_step_status(workspace): 0, or 1 if a grid barrier of some earlier step timed out (results invalid).
 * ------------------------------------------------------------------------------------------ */
#define VY_DECODE_MAX_LAYERS 24

typedef struct VyDecodeLayer {
  const void* w_qkv; const void* b_qkv;   /* attention.{query,key,value}.{weight,bias} packed (bias may be NULL) */
  const void* w_o;   const void* b_o;     /* attention.out.dense */
  const void* ln1_g; const void* ln1_b;   /* attention.out.layernorm */
  const void* w_1;   const void* b_1;     /* feed_forward.intermediate */
  const void* w_2;   const void* b_2;     /* feed_forward.out */
  const void* ln2_g; const void* ln2_b;   /* feed_forward.layernorm */
  void* k_cache;     void* v_cache;       /* this layer's [B', n_kv, cache_len, 64] bf16 caches */
} VyDecodeLayer;

typedef struct VyDecodeStep {
  int32_t B, H, n_q_heads, n_kv_heads, head_dim, ffn, vocab, n_layers;
  const VyDecodeLayer* layers;  /* HOST array of n_layers entries */
  const void* emb;              /* word_embeddings.weight [vocab, H] */
  const void* pos_table;        /* learned / sinusoidal position rows [>= cache_len, H], or NULL (RoPE models) */
  const float* rope_cos;        /* fp32 [rope_rows][32] or NULL */
  const float* rope_sin;
  int32_t rope_rows;
  const void* w_d; const void* b_d;             /* lm_head.dense */
  const void* ln_head_g; const void* ln_head_b; /* lm_head.layer_norm */
  const void* w_v; const void* b_v;             /* lm_head.decoder [vocab, H], lm_head.bias */
  float eps_layer, eps_head;
  int32_t cache_len;
  int64_t cache_sb, cache_sh, cache_sl;  /* element strides of the caches: batch, head, slot */
  int32_t pos_bound;            /* upper bound of *pos over the life of this launch description (sizes the kv-split) */
  int32_t* pos;                 /* device: position / cache slot of the token being fed; incremented by the step */
  int64_t* tok;                 /* device [B]: in = token fed to this step, out = the greedy next token */
  int64_t* tokens_out;          /* optional device [B][ld_tokens]: next token also stored at column pos + 1 */
  int64_t ld_tokens;
  void* logits;                 /* optional device bf16 [B][ld_logits]: the step's logits */
  int64_t ld_logits;
  void* workspace;
  int64_t workspace_bytes;
  int64_t* trace;               /* optional device [2 * (5 * n_layers + 2) + 1] (development, tools/decode_bench.py --trace):
                                   %globaltimer (ns) of CTA 0 at launch [0], when it reaches grid barrier k [1 + 2k] and when
                                   that barrier opens [2 + 2k] — a step's time stage by stage, own work vs waiting */
  void* stream;
} VyDecodeStep;

VY_API int vy_decode_step(const VyDecodeStep* p);
VY_API int64_t vy_decode_step_workspace_bytes(int B, int H, int n_q_heads, int n_kv_heads, int ffn);
VY_API int vy_decode_step_status(const void* workspace);

/* ------------------------------------------------------------------------------------------
 * vy_embed_fwd / vy_embed_bwd — row gather with fused positional add, scale and row remap.
 * replaces nn.Embedding lookup + "hidden_state + pos_info" at VyomAI/models/encoder.py:146-152,
 * models/decoder.py:343-350, models/multimodel.py:162-180 (the captioner writes its text rows
 * after the prepended image row: out_group_stride = S+1, out_row_off = 1), and the cls-row
 * "2 * (cls + pos[0])" of models/vision_encoder.py:119-127 (ids = NULL, ld_src = 0, out_scale = 2).
 *   srow = ids ? ids[r] : r;  l = r % tokens_per_seq
 *   out[(r / tokens_per_seq) * out_group_stride + l + out_row_off, :] =
 *        out_scale * (src[srow * ld_src + :] + (pos ? pos[(pos_row_off + l) * H + :] : 0))
 * bwd scatter-adds dout rows (same remap, times out_scale) into dtable[srow] and/or dpos[...]
 * with atomics (the autograd of the lookups above). ids are int64. H % 8 == 0.
 * ------------------------------------------------------------------------------------------ */
typedef struct VyEmbed {
  int32_t rows, H;
  const int64_t* ids; /* [rows] or NULL */
  const void* src;    /* table [vocab, H] (ld_src) or direct rows */
  int64_t ld_src;
  int32_t dtype; /* of src, pos, out, dout, dtable, dpos */
  int32_t vocab;
  int32_t tokens_per_seq, out_group_stride, out_row_off;
  const void* pos; /* [>= pos_row_off + tokens_per_seq, H] or NULL */
  int32_t pos_row_off;
  float out_scale; /* 0 = 1 */
  void* out;
  int64_t ld_out; /* also the row stride of dout in bwd */
  /* bwd only */
  const void* dout;
  void* dtable; /* or NULL */
  void* dpos;   /* or NULL */
  /* nn.Embedding(padding_idx=...) never receives a gradient for that row (models/encoder.py:100-104 word_embeddings;
   * layers/positional_embeddings.py:20-24 builds the learned position table with padding_idx = pad_token_id too).
   * Stored as index + 1 so that zero-initialised structs mean "no padding row": bwd skips ids[r] == padding_idx_plus1 - 1
   * for dtable and position row == pos_padding_idx_plus1 - 1 for dpos. */
  int64_t padding_idx_plus1;
  int64_t pos_padding_idx_plus1;
  void* stream;
} VyEmbed;

VY_API int vy_embed_fwd(const VyEmbed* p);
VY_API int vy_embed_bwd(const VyEmbed* p);

/* vy_patchify — NCHW pixels -> [B * (H/ph) * (W/pw), C*ph*pw] patch rows, (c, i, j) order with j
 * fastest: the im2col of the stride==kernel nn.Conv2d at VyomAI/models/vision_encoder.py:83-88,114
 * so that the patch embedding is one vy_gemm against pixel_seq.weight viewed as [hidden, C*ph*pw]. */
typedef struct VyPatchify {
  int32_t B, C, H, W, patch_h, patch_w;
  const void* pixels;
  int32_t in_dtype;
  void* out;
  int32_t out_dtype;
  int64_t ld_out; /* row stride of out in elements; 0 = C * patch_h * patch_w. A wider stride pads each patch row (columns beyond the
                   * patch are left untouched) — SigLIP's 14 x 14 patches give 588 columns, which vy_gemm needs padded to 592 */
  void* stream;
} VyPatchify;

VY_API int vy_patchify(const VyPatchify* p);

/* vy_argmax_rows — greedy token selection: out[r] = first index of the row maximum, the
 * torch.topk(k=1) rule of VyomAI/models/decoder.py:489-496 and generation_utils.py:179-189
 * (argmax of softmax(logits) == argmax of logits). out is int64 [rows]. */
VY_API int vy_argmax_rows(int rows, int V, const void* x, int64_t ld, int dtype, int64_t* out, void* stream);
/* The same plus the greedy loop's bookkeeping (models/decoder.py:489-500) in the same launch sequence: out[r] = argmax, then
 * *pos += 1 and, when `tokens` is given, tokens[r * tokens_ld + *pos] = out[r] (skipped when *pos falls outside
 * [0, tokens_cols)). rows <= 1024. pos is a device int32 scalar: the tail of a captured single-token decode step. */
VY_API int vy_argmax_advance(int rows, int V, const void* x, int64_t ld, int dtype, int64_t* out, int32_t* pos, int64_t* tokens, int64_t tokens_ld,
                             int tokens_cols, void* stream);

/* vy_colsum — out[c] (+)= scale * sum_r x[r, c]: the bias gradient of every nn.Linear on the path.
 * workspace: vy_colsum_workspace_floats(cols) floats. */
VY_API int vy_colsum(int rows, int cols, const void* x, int64_t ld, int dtype, void* out, int out_dtype,
                     int accumulate, float scale, float* workspace, void* stream);
VY_API int vy_colsum_workspace_floats(int cols);
/* Second stage of a two-stage column sum on its own: out[c] (+)= scale * (*scale_ptr if given) * sum_k part[k * part_ld + c],
 * k < chunks, added in index order. Used for the partial sums vy_softmax_xent leaves in VyXent.colsum_part. */
VY_API int vy_colsum_finish(int cols, int chunks, const float* part, int64_t part_ld, void* out, int out_dtype, int accumulate,
                            float scale, const float* scale_ptr, void* stream);

/* vy_cast4d — strided 4-D copy with dtype conversion (inner dim contiguous on both sides), e.g.
 * an fp32 StaticCacheOne prefix (layers/kv_cache.py:295-312) to the bf16 attention operands. */
typedef struct VyCast4d {
  int32_t n0, n1, n2, n3;
  const void* src;
  int32_t src_dtype;
  int64_t s0, s1, s2;
  void* dst;
  int32_t dst_dtype;
  int64_t d0, d1, d2;
  void* stream;
} VyCast4d;

VY_API int vy_cast4d(const VyCast4d* p);

/* vy_softmax_xent — fused softmax cross-entropy over the vocabulary, the training loss of the
 * captioner / CLM notebooks (Examples/vyom-ai-accelerate-multimodel-2t4.ipynb cell 1 `loss_fn`:
 * F.cross_entropy(logits.view(-1, V), labels.view(-1), ignore_index)). loss_rows[r] = logsumexp -
 * logit[label] (0 for ignored rows). With write_grad the logits buffer is overwritten IN PLACE by
 * d loss / d logits = (softmax - onehot) * grad_scale * (*grad_scale_ptr if given). */
typedef struct VyXent {
  int32_t rows, V;
  void* logits;
  int64_t ld;
  int32_t dtype;
  const int64_t* labels;
  int64_t ignore_index;
  const float* grad_scale_ptr; /* device scalar or NULL */
  float grad_scale;
  float* loss_rows; /* [rows] or NULL */
  int32_t write_grad;
  /* Optional, with write_grad: [vy_xent_colsum_chunks(rows, V, dtype)][V rounded up to 8] fp32; needs ld % 8 == 0 and
   * ld >= V rounded up to 8 (columns V .. of the logits rows are then written with zeros). The column sums of the written gradient
   * (the bias gradient of the vocabulary projection, models/decoder.py:267-275 `self.decoder` of LMHead) are then taken in
   * the same pass as per-CTA partial sums, to be added up by vy_colsum_finish: the [rows, V] gradient is not re-read. */
  float* colsum_part;
  void* stream;
} VyXent;

VY_API int vy_softmax_xent(const VyXent* p);
/* Rows of VyXent.colsum_part for this shape; 0 = the fused column sums are not available (needs bf16, V <= 57344). */
VY_API int vy_xent_colsum_chunks(int rows, int V, int dtype);

/* vy_sqnorm — *out += sum(g^2) over a flat gradient buffer (out must be zeroed by the caller);
 * vy_adamw — fused AdamW over a flat parameter buffer (torch.optim.AdamW semantics: decoupled
 * weight decay, bias correction), with the global-norm clip coefficient computed on device from
 * *grad_sqnorm (clip_grad_norm_(…, max_grad_norm) of the notebooks' training loop) and grad_div
 * dividing the gradient first (data-parallel mean). master (optional fp32 copy) is updated and the
 * bf16/fp32 param rewritten from it. */
VY_API int vy_sqnorm(int64_t n, const void* g, int dtype, float* out, void* stream);

typedef struct VyAdamW {
  int64_t n;
  void* param;
  int32_t param_dtype;
  const void* grad;
  int32_t grad_dtype;
  float* exp_avg;
  float* exp_avg_sq;
  float* master; /* or NULL */
  float lr, beta1, beta2, eps, weight_decay;
  int32_t step; /* 1-based */
  const int32_t* step_ptr; /* optional device copy of the step counter (wins over `step`): lets a
                              captured CUDA graph of the whole training step be replayed */
  const float* grad_sqnorm; /* device scalar or NULL */
  float max_grad_norm;      /* <= 0: no clipping */
  float grad_div;           /* 0 = 1 */
  void* stream;
} VyAdamW;

VY_API int vy_adamw(const VyAdamW* p);

/* ------------------------------------------------------------------------------------------
 * Data-parallel optimizer step over NVLink peer memory (one node, <= 8 ranks) — replaces, for the notebooks' training loop
 * (Examples/vyom-ai-accelerate-multimodel-2t4.ipynb cell 1 main(): DDP gradient all-reduce + clip_grad_norm_(1.0) +
 * AdamW on every rank), the pair "all-reduce the whole gradient, run the whole optimizer everywhere" by a sharded step
 * whose kernels read and write the other ranks' buffers directly:
 *   vy_dp_barrier       device-side rendezvous of the ranks (flags in symmetric memory, generation counted; capturable);
 *   vy_dp_reduce_shard  gshard[i - lo] = sum_r grad_r[i] for i in [lo, hi) (peer loads, rank order, fp32) and the shard's
 *                       sum of squares written to slot [rank] of every rank's `scalars`;
 *   vy_dp_adamw_shard   after a barrier: clip coefficient from the N published partial norms (the norm is that of the MEAN
 *                       gradient, like clip_grad_norm_ on DDP-averaged gradients), torch.optim.AdamW semantics on the shard
 *                       with fp32 master weights / moments that exist for the shard only, new parameters stored into EVERY
 *                       rank's parameter buffer; after one more barrier all ranks hold bit-identical parameters.
 * All pointer tables are HOST arrays of `world` device pointers that are valid on the calling rank's device (peer-mapped:
 * cudaIpc / CUDA VMM, e.g. torch.distributed._symmetric_memory). flags: uint32 [world] per rank, zero-initialised; scalars:
 * float [world] per rank; epoch: local device uint32 (zero-initialised); error_flag: local device int32, raised when a
 * barrier wait times out (a rank that never arrives) instead of hanging the GPU. */
typedef struct VyDpGroup {
  int32_t world, rank;
  void* const* grads;      /* [world] flat gradient buffers (dtype of the model), or NULL for a barrier-only group */
  void* const* params;     /* [world] flat parameter buffers */
  uint32_t* const* flags;  /* [world] barrier flags */
  float* const* scalars;   /* [world] scalar slots */
  uint32_t* epoch;
  int32_t* error_flag;
} VyDpGroup;

typedef struct VyDpReduce {
  int64_t lo, hi;   /* this rank's shard of the flat buffers, in elements, multiples of 8 */
  int32_t dtype;    /* of the gradient buffers */
  float* gshard;    /* out: fp32 [hi - lo] summed gradient */
  float* sq_local;  /* out: device scalar, sum of squares of gshard */
  void* stream;
} VyDpReduce;

typedef struct VyDpAdamW {
  int64_t lo, hi;
  int32_t param_dtype;
  const float* gshard;
  float* exp_avg;     /* fp32 [hi - lo] */
  float* exp_avg_sq;  /* fp32 [hi - lo] */
  float* master;      /* fp32 [hi - lo] master copy of the shard, or NULL (fp32 parameters) */
  float lr, beta1, beta2, eps, weight_decay, max_grad_norm; /* max_grad_norm <= 0: no clipping */
  int32_t step;
  const int32_t* step_ptr; /* optional device step counter (captured CUDA graphs) */
  void* stream;
} VyDpAdamW;

VY_API int vy_dp_barrier(const VyDpGroup* g, void* stream);
VY_API int vy_dp_reduce_shard(const VyDpGroup* g, const VyDpReduce* p);
VY_API int vy_dp_adamw_shard(const VyDpGroup* g, const VyDpAdamW* p);

#ifdef __cplusplus
}
#endif
#endif /* VYOM_B200_H_ */
