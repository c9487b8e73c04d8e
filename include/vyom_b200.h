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

#define VY_ABI_VERSION 1

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
  VY_ACT_DGELU_TANH = 4
};

enum { VY_EPI_LINEAR = 0, VY_EPI_QKV_ROPE = 1 };

VY_API int vy_version(void);
VY_API const char* vy_last_error(void);
/* number of kernels this library has launched from the calling process (all threads) */
VY_API int64_t vy_launch_count(void);
/* 1 if the current device is sm_100 (B200), else 0 */
VY_API int vy_device_ok(void);

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
 * k_out[b,hk,start_pos+l,:], v_out[...] through the given element strides — i.e. the
 * "b l (h d) -> b h l d" rearrange (attention.py:118-120) and the kv-cache append
 * (kv_cache.py:355-359) are fused into the projection.
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
  int32_t act; /* VY_ACT_* */
  int32_t transposed_out;
  const void* bias; /* [cols] or NULL */
  int32_t bias_dtype;
  const void* addend; /* or NULL */
  int64_t ld_addend;
  int32_t addend_dtype;
  int32_t addend_row_mod, addend_row_off;
  void* aux; /* pre-activation: written by GELU_*, read by DGELU_*; or NULL */
  int64_t ld_aux;
  int32_t aux_dtype;
  float out_scale; /* 0 is treated as 1 */
  void* out;
  int64_t ld_out;
  int32_t out_dtype;
  int32_t out_row_group, out_row_group_stride, out_row_off;

  /* VY_EPI_QKV_ROPE only */
  int32_t tokens_per_seq, start_pos, head_dim, n_q_heads, n_kv_heads;
  const float* rope_cos; /* [>= start_pos + tokens_per_seq][head_dim/2] */
  const float* rope_sin;
  void* q_out;
  int64_t q_sb, q_sh, q_sl; /* element strides of batch, head, token; head_dim contiguous */
  void* k_out;
  int64_t k_sb, k_sh, k_sl;
  void* v_out;
  int64_t v_sb, v_sh, v_sl;

  void* stream;
} VyGemm;

VY_API int vy_gemm(const VyGemm* p);

/* ------------------------------------------------------------------------------------------
 * vy_add_layernorm_fwd / _bwd — y = LayerNorm(x + residual) * gamma + beta, warp per row.
 * replaces `self.layernorm(hidden_states + input_tensor)` at VyomAI/layers/attention.py:52-54,71
 * and VyomAI/layers/ffn.py:25,39 (biased variance, eps = config.layer_norm_eps), and the plain
 * LayerNorm of the LM head (models/decoder.py:259-261,270; residual = NULL).
 * mean / rstd (fp32, [rows]) are optional outputs of fwd and required inputs of bwd.
 * bwd: dx[rows,H] (= d residual as well), dgamma/dbeta partials are reduced into fp32
 * dgamma[H], dbeta[H] (overwritten). xhat is recomputed from the saved pre-norm sum `s`
 * (s = x + residual; pass the tensor fwd wrote to sum_out, or x when residual was NULL).
 * ------------------------------------------------------------------------------------------ */
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
  float* dgamma; /* [H] fp32 */
  float* dbeta;  /* [H] fp32 */
  float* partials; /* workspace: 2 * vy_norm_bwd_partial_rows() * H floats */
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
  void* stream;
} VyAttn;

VY_API int vy_attn_fwd(const VyAttn* p);

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
  const void* qkv;
  int64_t ld_qkv;
  int32_t qkv_dtype;
  const float* rope_cos; /* [cache_len][32] fp32 or NULL */
  const float* rope_sin;
  void* k_cache;
  void* v_cache;
  int64_t cache_sb, cache_sh, cache_sl;
  int32_t cache_dtype;
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

#ifdef __cplusplus
}
#endif
#endif /* VYOM_B200_H_ */
