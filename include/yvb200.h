/* yvb200 -- C ABI of the B200-native ViLBERT hot path (libyvb200.so).
 *
 * The reference (JeremyLinky/YouTube-VLN) is pure Python on top of torch ATen; it has no FFI of its own.
 * The boundary it would bind for this path is therefore the set of operators that vilbert/vilbert.py
 * composes; each entry point below names the reference lines it replaces.  All pointers are DEVICE
 * pointers owned by the caller (PyTorch's caching allocator on the host side); the library never
 * allocates or frees device memory, never synchronises, and enqueues everything on `stream`.
 * Return value: 0 on success, non-zero on error with a message available from yv_last_error().
 *
 * Number formats
 *   f32       : row-major float matrices [rows, ld]
 *   "planes"  : a GEMM operand is two bf16 matrices, plane 0 = hi = bf16(x), plane 1 = lo = bf16(x - hi).
 *               passes == 3 contracts hi*hi + hi*lo + lo*hi in fp32 (the "bf16x3" parity mode, ~2^-16
 *               relative operand error); passes == 1 uses the hi planes only (plain bf16).
 */
#ifndef YVB200_H
#define YVB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* yv_stream_t; /* cudaStream_t */

const char* yv_last_error(void);
int yv_version(void);
/* kernels launched by this library since load (evidence for bench.py's "gpu_launches") */
uint64_t yv_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * yv_gemm: D = epilogue(alpha * A.B^T)  on tcgen05 tensor cores (TMA-staged tiles, TMEM accumulators; 128x128 per CTA or
 * 256x{128,256} per CTA pair)
 * replaces every nn.Linear / torch.matmul on the path: vilbert/vilbert.py:285-287,294,306,322,352,365,
 * 414-416,423,435,450,479,492,555-573,577,591,597,613,641,644,831,846,864,883,906,968,1358 and their
 * autograd dgrad / wgrad twins.
 * An operand is a (possibly batched, possibly transposed) view of a bf16 plane pair:
 *   mn_major == 0: element (r, k) at ptr[r*ld + k]          (rows index M or N, `inner` is K)
 *   mn_major == 1: element (r, k) at ptr[k*ld + r]          (rows index K, `inner` is M or N)
 * batch z = b1*nb0 + b0 adds b0*sb0 + b1*sb1; the lo plane sits plane_stride elements after hi.
 * ld, sb0, sb1, plane_stride must be multiples of 8 elements and ptr 16-byte aligned (TMA).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const void* ptr;
    int64_t inner, rows, ld;
    int64_t nb0, sb0, nb1, sb1;
    int64_t plane_stride;
    int32_t mn_major, _pad;
} YvOperand;

enum { YV_ACT_NONE = 0, YV_ACT_GELU = 1, YV_ACT_RELU = 2, YV_ACT_MUL_GELU_GRAD = 3, YV_ACT_MUL_RELU_MASK = 4 };

typedef struct {
    int32_t M, N, K, passes;      /* passes: 1 (bf16) or 3 (bf16x3) */
    YvOperand a, b;
    float alpha;
    int32_t act;
    const float* bias;            /* [N] or NULL */
    float* aux_out;               /* optional: pre-activation (alpha*acc + bias) saved for backward */
    const float* aux_in;          /* MUL_GELU_GRAD: pre-activation; MUL_RELU_MASK: forward output */
    const float* residual;        /* optional f32 addend (may alias out32 -> accumulate) */
    float* out32;                 /* optional f32 output */
    int64_t ld_out, out_sb0, out_sb1; /* layout shared by out32 / aux_out / aux_in / residual */
    void* out_planes;             /* optional bf16 plane-pair output */
    int64_t ld_pl, pl_sb0, pl_sb1, pl_plane_stride;
    float drop_p;                 /* dropout after the activation, before the residual add */
    uint32_t drop_site;
    const uint64_t* rng;          /* device {seed, step}; NULL or drop_p == 0 disables dropout */
    int32_t out32_zeroed;         /* the caller already zero-filled out32 (split-K launches skip their memset node) */
    int32_t _pad2;
} YvGemm;
int yv_gemm(const YvGemm* g, yv_stream_t stream);
/* number of K splits yv_gemm would use for this problem (1 = none); lets a caller zero-fill the output early */
int yv_gemm_splits(const YvGemm* g);
/* Tuning / test knob (process-wide, not thread-safe): which kernel yv_gemm launches.  0 = automatic (default),
 * 32 / 64 = one CTA per 128x128 tile (half-SM ring / persistent), 2 = CTA pairs (tcgen05 cta_group::2, 256-row
 * pair tiles, width chosen per problem), 128 / 256 = CTA pairs with that tile width. */
int yv_gemm_set_variant(int variant);

/* fp32 -> bf16 plane pair (activations entering the path, e.g. the 2048-d region features) */
int yv_split_planes(const float* src, int64_t ld_src, void* planes, int64_t ld_dst, int64_t plane_stride,
                    int64_t rows, int64_t cols, yv_stream_t stream);
/* all weights in one launch: seg table lives on the device */
typedef struct {
    const float* src;
    int64_t dst_off;   /* element offset into the hi plane */
    int64_t numel;
    int64_t first_blk; /* prefix sum of ceil(numel / 2048) */
} YvSplitSeg;
int yv_split_multi(const YvSplitSeg* segs_dev, int32_t nseg, int64_t total_blocks, void* planes,
                   int64_t plane_stride, yv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused multi-tensor AdamW -- SURVEY.md 8(f) "next" #1; replaces the per-parameter Python loop of
 * vilbert/optimization.py:141-187 (bias-corrected Adam, eps added after sqrt(v), decoupled decay applied to the
 * updated weight).  hyper_dev = {lr, step_size, beta1, beta2, eps, 1-beta1, 1-beta2}.  Segments with plane pointers also get the
 * updated weight re-split into bf16 hi/lo planes (saves the per-step yv_split_multi pass).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float* p;
    const float* g;
    float* m;
    float* v;
    void* plane_hi;      /* NULL for parameters that are not GEMM operands */
    void* plane_lo;
    int64_t numel;
    int64_t first_blk;   /* prefix sum of ceil(numel / 2048) */
    float weight_decay;
    int32_t _pad;
} YvAdamSeg;
int yv_adamw_multi(const YvAdamSeg* segs_dev, int32_t nseg, int64_t total_blocks, const float* hyper_dev,
                   yv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Batch masking on the device -- SURVEY.md 8(f) "next" #3; replaces utils/dataset/common.py:213-270 (randomize_tokens)
 * and :272-300 (randomize_regions) of the reference's CPU data path.  The uniform draws are inputs (p in [0,1) per
 * token / region, replacement token ids, optional host-picked action-word positions), so the result is a pure
 * function of its arguments and bit-exact against the reference.  tokens / features are updated in place.
 * ---------------------------------------------------------------------------------------------- */
int yv_mask_tokens(int64_t* tokens, const uint8_t* mask, const float* p, const int64_t* random_tokens,
                   const uint8_t* forced /* may be NULL */, int64_t mask_id, int64_t* targets, int64_t n,
                   yv_stream_t stream);
int yv_mask_regions(float* features /* [rows, F] */, const float* probs /* [rows, C] */, const int64_t* mask,
                    const float* p, float* targets /* [rows, C] */, int64_t* targets_mask, int64_t rows, int32_t F,
                    int32_t C, yv_stream_t stream);

/* dropout RNG state {seed, step}: step += 1 (captured once per training step) */
int yv_rng_advance(uint64_t* rng, yv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm (vilbert/vilbert.py:204-217: biased variance, eps inside the sqrt), optionally followed by
 * dropout (embeddings: :254-255, :1367-1368).  y32 and/or y_planes may be NULL.  stats = {mean, rstd}[M].
 * bwd: dx = LN'(dy) (+ dx_add), dgamma/dbeta accumulated with atomics into zero-initialised buffers.
 *      dx_planes (optional) = dx * dropmask(pre_site): the operand of the preceding dense layer's
 *      dgrad/wgrad when that layer's output went through dropout before the residual add (:323-324);
 *      dbias (optional, zero-initialised) accumulates the column sums of that same masked gradient, i.e.
 *      the bias gradient of the preceding dense layer.
 * ---------------------------------------------------------------------------------------------- */
int yv_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, float* y32,
                     void* y_planes, int64_t plane_stride, float* stats, int64_t M, int32_t C, float drop_p,
                     uint32_t drop_site, const uint64_t* rng, yv_stream_t stream);
int yv_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* stats,
                     float post_drop_p, uint32_t post_drop_site, const float* dx_add, float* dx32,
                     void* dx_planes, int64_t plane_stride, float pre_drop_p, uint32_t pre_drop_site,
                     const uint64_t* rng, float* dgamma, float* dbeta, float* dbias, int64_t M, int32_t C,
                     yv_stream_t stream);

/* The same backward in two launches (autograd of vilbert/vilbert.py:204-217 again): _dx computes only what the rest of the
 * backward pass waits for (dx, dx_planes); _cols accumulates dgamma / dbeta (zero-initialised by the caller) and can trail
 * on another stream next to the weight gradients.  (The bias gradient of the preceding dense layer is then
 * yv_colsum_planes of dx_planes.) */
int yv_layernorm_bwd_dx(const float* dy, const float* x, const float* gamma, const float* stats, const float* dx_add,
                        float* dx32, void* dx_planes, int64_t plane_stride, float pre_drop_p, uint32_t pre_drop_site,
                        const uint64_t* rng, int64_t M, int32_t C, yv_stream_t stream);
int yv_layernorm_bwd_cols(const float* dy, const float* x, const float* stats, float* dgamma, float* dbeta, int64_t M,
                          int32_t C, yv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * attention softmax (vilbert/vilbert.py:295-304, 424-433, 578-589, 598-611):
 *   P = softmax(S * scale + mask[pair, key]); S is overwritten by P (kept for backward);
 *   planes = dropout(P) split to bf16 hi/lo.   rows = pairs*heads*Tq, row r belongs to pair r / rows_per_pair.
 * bwd: dS = scale * P * (dPd*m - sum_j P*dPd*m) with m the recomputed dropout multiplier -> planes.
 * ---------------------------------------------------------------------------------------------- */
int yv_softmax_fwd(float* s, int64_t ld_s, const float* mask, int64_t rows, int32_t cols, int64_t rows_per_pair,
                   float scale, void* p_planes, int64_t ld_p, int64_t plane_stride, float drop_p,
                   uint32_t drop_site, const uint64_t* rng, yv_stream_t stream);
int yv_softmax_bwd(const float* p, const float* dpd, int64_t ld_s, int64_t rows, int32_t cols, float scale,
                   void* ds_planes, int64_t ld_p, int64_t plane_stride, float drop_p, uint32_t drop_site,
                   const uint64_t* rng, yv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * yv_attn_fwd / yv_attn_bwd: fused scaled-dot-product attention on tcgen05 (one launch each):
 *   O = dropout(softmax(Q K^T * scale + mask[pair, key])) V
 * replaces torch.matmul -> +mask -> nn.Softmax -> nn.Dropout -> torch.matmul of vilbert/vilbert.py:294-306
 * (BertSelfAttention), :423-435 (BertImageSelfAttention) and :577-589 / :597-611 (the two directions of
 * BertBiAttention) plus their autograd backward.  Scores and probabilities stay in TMEM / shared memory; the forward
 * saves one log-sum-exp per query row and the backward recomputes the probabilities (same dropout mask: the counter
 * RNG is keyed by (pair, head, query, key) exactly like yv_softmax_fwd).
 * A YvHeadView is a [pairs, rows, heads*dh] view of a bf16 plane pair (e.g. the Q third of a fused Q|K|V projection
 * output): element (pair, row, head, d) at ptr[pair*pair_stride + row*ld + head*dh + d], lo plane plane_stride later.
 * dh must be 64 or 128 (yv_attn_supported); ptr 16-byte aligned; ld, pair_stride, plane_stride multiples of 8.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const void* ptr;
    int64_t ld, plane_stride, pair_stride;
    int32_t rows, _pad;
} YvHeadView;

typedef struct {
    int32_t pairs, heads, dh, passes;   /* passes: 1 (bf16) or 3 (bf16x3, also for P V) */
    YvHeadView q, k, v;                 /* q.rows = Tq, k.rows = v.rows = Tk */
    const float* mask;                  /* [pairs, Tk] additive (0 / -10000, vilbert/vilbert.py:1268-1287) or NULL */
    float scale;                        /* 1 / sqrt(dh) */
    float drop_p;
    uint32_t drop_site, _pad;
    const uint64_t* rng;
    void* out_planes;                   /* context [pairs*Tq, heads*dh] as a plane pair (heads merged, :307-309) */
    int64_t ld_out, out_plane_stride;
    float* out32;                       /* optional fp32 copy of the context */
    int64_t ld_out32;
    float* lse;                         /* [pairs*heads*Tq] row log-sum-exp, input of yv_attn_bwd (may be NULL) */
} YvAttnFwd;
int yv_attn_supported(int32_t dh, int32_t passes);
int yv_attn_fwd(const YvAttnFwd* a, yv_stream_t stream);

typedef struct {
    int32_t pairs, heads, dh, passes;
    YvHeadView q, k, v;
    YvHeadView dout, out;               /* gradient of the context and the forward's context, rows = Tq */
    const float* mask;
    float scale;
    float drop_p;
    uint32_t drop_site, _pad;
    const uint64_t* rng;
    const float* lse;
    YvHeadView dq, dk, dv;              /* outputs (plane pairs); dq.rows = Tq, dk.rows = dv.rows = Tk */
    void* workspace;                    /* scratch (contents irrelevant), yv_attn_bwd_workspace_bytes() bytes: per-query-
                                           tile partial dK / dV, added up by the last CTA of each (pair, head) */
    size_t workspace_bytes;
    uint32_t* tickets;                  /* [pairs*heads], zero before the first launch; every launch leaves it zero */
} YvAttnBwd;
size_t yv_attn_bwd_workspace_bytes(int32_t pairs, int32_t heads, int32_t dh, int32_t Tq, int32_t Tk);
int yv_attn_bwd(const YvAttnBwd* a, yv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * embeddings
 *   text  (vilbert/vilbert.py:240-253): out[m] = word[tok[m]] + pos[m % T] + type[seg[m]]
 *   image (vilbert/vilbert.py:1361-1365): out[m] = W5.loc[0:5]+b5 + W4.loc[5:9]+b4 + W2.loc[9:11]+b2 + seq[loc[11]]
 *          (the 2048->1024 projection is a yv_gemm with this tensor as its residual)
 * bwd scatters with atomics into zero-initialised (or accumulating) gradient buffers.
 * ---------------------------------------------------------------------------------------------- */
int yv_embed_text_fwd(const int64_t* tok, const int64_t* seg, const float* word, const float* pos,
                      const float* type, float* out, int64_t M, int32_t T, int32_t H, yv_stream_t stream);
int yv_embed_text_bwd(const int64_t* tok, const int64_t* seg, const float* dout, float* dword, float* dpos,
                      float* dtype_, int64_t M, int32_t T, int32_t H, int32_t padding_idx, yv_stream_t stream);
int yv_embed_loc_fwd(const float* loc, const float* w5, const float* b5, const float* w4, const float* b4,
                     const float* w2, const float* b2, const float* seq, float* out, int64_t M, int32_t H,
                     yv_stream_t stream);
int yv_embed_loc_bwd(const float* loc, const float* dout, float* dw5, float* db5, float* dw4, float* db4,
                     float* dw2, float* db2, float* dseq, int64_t M, int32_t H, yv_stream_t stream);

/* column sums of an f32 matrix (bias gradients): out[c] (+)= sum_r x[r, c] */
int yv_colsum(const float* x, int64_t ld, int64_t rows, int32_t cols, float* out, int32_t accumulate,
              yv_stream_t stream);
int yv_colsum_planes(const void* planes, int64_t ld, int64_t plane_stride, int64_t rows, int32_t cols,
                     float* out, int32_t accumulate, yv_stream_t stream);
/* Gradient exchange over peer-to-peer copies (the reference gets its averaging from DistributedDataParallel's NCCL
 * all-reduce, utils/distributed.py:97-99): mean over the ranks, summed in rank order, of one chunk of the flat gradient
 * buffer.  `own` = this rank's copy of the chunk (overwritten with the mean); row step-1 of `stage` (rows `stage_stride`
 * floats apart) = the copy pulled from rank (rank - step) mod world, step = 1..world-1.  n floats, a multiple of 4. */
int yv_mean_chunks(float* own, const float* stage, int64_t stage_stride, int32_t world, int32_t rank, int64_t n,
                   yv_stream_t stream);
/* backward of the GEMM-epilogue activations when the upstream gradient arrives as f32:
 * planes = dy * act'(aux)  (GELU: aux = saved pre-activation, vilbert/vilbert.py:113-119; ReLU: aux = output) */
int yv_act_bwd_split(const float* dy, int64_t ld_dy, const float* aux, int64_t ld_aux, int32_t act, void* planes,
                     int64_t ld_dst, int64_t plane_stride, int64_t rows, int64_t cols, float* dbias /* optional column
                     sums = bias gradient, zeroed by the call */, yv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * fused losses (utils/utils_init.py:117-135)
 *   ce : masked-language cross entropy, ignore_index = -1, mean over kept rows.
 *        loss_sum[0] += sum_rows(-log p[target]); count[0] += kept rows;  dlogits planes (optional) get
 *        (softmax - onehot) * grad_scale[0] / max(count,1) for kept rows, 0 otherwise (two-phase: call with
 *        dl_planes == NULL first to obtain count, or pass count_in).
 *   kl : masked-vision KL(target || softmax(logits)) summed over rows with mask, / max(1, sum(mask)).
 * ---------------------------------------------------------------------------------------------- */
int yv_ce_loss(const float* logits, int64_t ld, const int64_t* target, int64_t rows, int32_t cols,
               float* loss_sum, float* count, yv_stream_t stream);
int yv_ce_grad(const float* logits, int64_t ld, const int64_t* target, int64_t rows, int32_t cols,
               const float* count, const float* gscale, float* dl32, void* dl_planes, int64_t ld_p,
               int64_t plane_stride, yv_stream_t stream);
int yv_kl_loss(const float* logits, int64_t ld, const float* target, int64_t ld_t, const int64_t* mask,
               int64_t rows, int32_t cols, float* loss_sum, float* count, yv_stream_t stream);
int yv_kl_grad(const float* logits, int64_t ld, const float* target, int64_t ld_t, const int64_t* mask,
               int64_t rows, int32_t cols, const float* count, const float* gscale, float* dl32,
               void* dl_planes, int64_t ld_p, int64_t plane_stride, yv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* YVB200_H */
