/* rgbnm_b200.h -- C ABI of the B200-native DCT-domain ViT hot path.
 *
 * Drop-in boundary for the hot path of JeongsooP/RGB-no-more.
 * Every entry point is `extern "C"`, takes plain pointers and sizes (device pointers
 * are raw `void*`/typed pointers into HBM, `stream` is a `cudaStream_t` passed as
 * `void*`), and returns an `rgbnm_status` (0 = OK).  No torch types cross this line.
 * Reference citations are file:line relative to the reference checkout.
 *
 * The shared library is `rgb_no_more_b200/librgbnm_b200.so`, built by
 * `__graft_entry__.build()` with nvcc for sm_100a only.
 */
#ifndef RGBNM_B200_H
#define RGBNM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    RGBNM_OK = 0,
    RGBNM_ERR_NOT_JPEG = 1,
    RGBNM_ERR_CORRUPT = 2,
    RGBNM_ERR_UNSUPPORTED = 3,
    RGBNM_ERR_PROGRESSIVE = 4,
    RGBNM_ERR_BUFFER = 5,
    RGBNM_ERR_IO = 6,
    RGBNM_ERR_CUDA = 7,
    RGBNM_ERR_ARG = 8
} rgbnm_status;

const char* rgbnm_strerror(int code);
/* Text of the last CUDA error seen by this thread's calls ("" if none). */
const char* rgbnm_last_cuda_error(void);
/* ABI version; bumped whenever a struct below changes. */
int rgbnm_abi_version(void);

/* ------------------------------------------------------------------------------
 * (B1) JPEG front end -- host.  Replaces dct_manip.read_coefficients
 *      (dct_manip/dct_manip.cpp:152-178 -> read_coefficients_using :98-150 ->
 *      extract_channel :78-96).  Huffman decode stays on the host (north_star).
 * ---------------------------------------------------------------------------- */
typedef struct {
    int32_t width, height, ncomp, progressive;
    int32_t hb[3], wb[3];       /* height/width in 8x8 blocks per component      */
    int32_t dsh[3], dsw[3];     /* downsampled height/width (the `dimensions` tensor) */
    int32_t hsamp[3], vsamp[3];
} rgbnm_jpeg_info;

int rgbnm_jpeg_info_from_memory(const uint8_t* data, size_t size, rgbnm_jpeg_info* info);

/* One image.  y: hb[0]*wb[0]*64 int16; cbcr: 2*hb[1]*wb[1]*64 int16 (Cb plane then Cr
 * plane; ignored for grayscale); quant: ncomp*64 int16, natural order; dims: ncomp*2
 * int32 (may be NULL); clamp_flag: set to 1 iff some dequantised coefficient leaves
 * [-1024, 1016], i.e. the clamp of datasets.py:288-290 is live (may be NULL). */
int rgbnm_jpeg_read_coefficients(const uint8_t* data, size_t size, int16_t* y, size_t y_capacity,
                                 int16_t* cbcr, size_t c_capacity, int16_t* quant, int32_t* dims,
                                 int32_t* clamp_flag);

/* Batch decode on `nthreads` host threads (0 = all cores) straight into the batch
 * layout the fused kernel reads: y [n][hb][wb][64], cbcr [n][2][hb/2][wb/2][64],
 * quant [n][3][64].  Every image must be hb x wb blocks, 4:2:0 or grayscale.
 * Replaces DataLoader workers + default collate (datasets.py:542-556). */
int rgbnm_jpeg_decode_batch(const uint8_t* const* data, const size_t* sizes, int n, int hb, int wb,
                            int16_t* y, int16_t* cbcr, int16_t* quant, uint8_t* clamp_flags,
                            int32_t* status, int nthreads);

/* Plan-first variant: last_block_row[i] (may be NULL = whole image) is the last LUMA block row image i's crop window
 * needs (crop_i + crop_size - 1 of its plan, drawn before decoding).  The scan is decoded up to and including the MCU row
 * that holds it and then abandoned; block rows below keep whatever the buffers held (the fused kernel reads the crop
 * window only, custom_transforms.py:557-629), and the clamp flag covers the decoded part.  Same arguments otherwise. */
int rgbnm_jpeg_decode_batch_rows(const uint8_t* const* data, const size_t* sizes, int n, int hb, int wb,
                                 int16_t* y, int16_t* cbcr, int16_t* quant, uint8_t* clamp_flags,
                                 int32_t* status, int nthreads, const int32_t* last_block_row);

int rgbnm_jpeg_read_file(const char* path, uint8_t** out, size_t* size);
void rgbnm_free(void* p);

/* Coefficient writer for fixtures: mirror of write_coefficients (dct_manip.cpp:265-313).
 * Luma sampling chroma_h x chroma_v (1 or 2), chroma 1x1; standard Annex-K tables. */
int rgbnm_jpeg_write_coefficients(int width, int height, int ncomp, int chroma_h, int chroma_v,
                                  const int16_t* y, const int16_t* cbcr, const int16_t* quant,
                                  uint8_t** out, size_t* out_size);

/* ------------------------------------------------------------------------------
 * (K0) Fused DCT-domain data path -- device.  One launch replaces, per image:
 *   dequantise+clamp (datasets.py:288-293) -> crop (dct_ops.py:584-599) ->
 *   resize (dct_ops.py:529-580) -> RandomFlip_DCT (custom_transforms.py:913-942) ->
 *   RandAugment_dct ops (custom_transforms.py:944-1127) -> ToRange (:406-466) ->
 *   grouped-embedding rearrange + sub-block conversion + collapse
 *   (models/plainvit.py:200-216)
 * and writes the (B, 196, 384) input of the patch-projection Linear.
 * ---------------------------------------------------------------------------- */
enum rgbnm_op {
    RGBNM_OP_NOP = 0, RGBNM_OP_TRANSLATE_X = 1, RGBNM_OP_TRANSLATE_Y = 2, RGBNM_OP_ROT90 = 3,
    RGBNM_OP_CUTOUT = 4, RGBNM_OP_BRIGHTNESS = 5, RGBNM_OP_CONTRAST = 6, RGBNM_OP_COLOR = 7,
    RGBNM_OP_AUTOCONTRAST = 8, RGBNM_OP_AUTOSATURATION = 9, RGBNM_OP_POSTERIZE = 10,
    RGBNM_OP_SHARPNESS = 11, RGBNM_OP_MIDFREQ = 12, RGBNM_OP_GRAYSCALE = 13,
    RGBNM_OP_CHROMADROP = 14, RGBNM_OP_SOLARIZE_ADD = 15, RGBNM_OP_INVERT = 16,
    RGBNM_OP_FREQ_ENHANCE = 17,  /* every coefficient but the DC term * f, Y and CbCr (dct_ops.py:1015-1034) */
    RGBNM_OP_EQUALIZE = 18,      /* histogram equalisation of the luma DC plane (dct_ops.py:916-955) */
    RGBNM_OP_SOLARIZE = 19       /* blocks whose luma DC > f inverted; chroma block (r,c) follows luma block (2r,2c) (dct_ops.py:631-651) */
};
#define RGBNM_MAX_OPS 4
#define RGBNM_FILTER_SLOTS 48

typedef struct {
    int16_t code;      /* enum rgbnm_op */
    int16_t p[8];      /* integer parameters (see rgb_no_more_b200/plan.py:resolve_op) */
    int16_t pad;
    float f;           /* float parameter, fp32-exact */
} rgbnm_plan_op;       /* 24 bytes */

typedef struct {
    int16_t crop_i, crop_j, crop_size;  /* luma crop window in blocks (chroma = /2) */
    int16_t flip;                       /* RandomFlip_DCT applied */
    int16_t n_ops;
    int16_t clamp_in;                   /* 1: apply the dequant clamp (0 only if the decoder proved it idle) */
    int16_t needs_stats;                /* plan holds Brightness / AutoContrast / AutoSaturation */
    int16_t train;                      /* RandAugment stage present (entry clamp) */
    rgbnm_plan_op ops[RGBNM_MAX_OPS];
} rgbnm_plan;          /* 112 bytes */

/* Per-launch constant tables (device pointers). */
typedef struct {
    const float* filters;        /* [RGBNM_FILTER_SLOTS][64] multiplicative 8x8 filters */
    const int16_t* posterize_lut;/* [6][2048] */
    int16_t* equalize_lut;       /* [n][RGBNM_MAX_OPS][2048] scratch: rgbnm_k0_dcstats writes the per-image DC mapping of every
                                    Equalize op (and the per-block 0/1 mask of every Solarize op, index r * grid + c),
                                    rgbnm_k0_fused reads it; may be NULL when no plan holds such an op */
} rgbnm_k0_tables;

#define RGBNM_K0_OUT_F32 0
#define RGBNM_K0_OUT_BF16 1
#define RGBNM_K0_OUT_INT16_PLANES 2   /* debug/parity: int16 planes handed to ToRange */

/* stats: float [n][RGBNM_MAX_OPS][2] scratch written by rgbnm_k0_dcstats and read by
 * rgbnm_k0_fused (Brightness: mean|dc|*m; AutoContrast/AutoSaturation: min, max).
 * y/cbcr/quant: quantised int16 coefficients in the batch layout above (hb x wb luma
 * blocks).  out: [n][196][384] (f32 or bf16) or, in INT16_PLANES mode,
 * [n][(28*28 + 2*14*14)*64] int16. */
int rgbnm_k0_dcstats(const int16_t* y, const int16_t* cbcr, const int16_t* quant, const rgbnm_plan* plans,
                     const rgbnm_k0_tables* tables, float* stats, int n, int hb, int wb, void* stream);
int rgbnm_k0_fused(const int16_t* y, const int16_t* cbcr, const int16_t* quant, const rgbnm_plan* plans,
                   const rgbnm_k0_tables* tables, const float* stats, void* out, int out_mode, int n,
                   int hb, int wb, void* stream);
/* Number of kernels the two calls above launch for one batch (for gpu_launches accounting). */
int rgbnm_k0_launch_count(void);

/* The same two calls with an explicit output layout.
 *   RGBNM_K0_LAYOUT_VIT16 (the calls above): planes resized to 28 x 28 luma blocks (crop side 14 / 28 / 56),
 *     out [n][196][384] = per 16 x 16 patch [Y: A16 . X . A16^T (256) | Cb 64 | Cr 64]       models/plainvit.py:200-216
 *   RGBNM_K0_LAYOUT_SWIN4: the SwinV2 data path (datasets.py:370-382: RandomResizedCrop_DCT(32) / Resize_DCT(32)):
 *     planes resized to 32 x 32 luma blocks (crop side 16 / 32 / 64), out [n][64*64][24] = per 4 x 4 patch
 *     [Y 4x4 (16) | Cb 2x2 (4) | Cr 2x2 (4)]: every 8 x 8 block DECOMPOSED, D = A^T . X . A with A = A(4,2) / A(2,4),
 *     and read out through the reference's interleaved rearrange "(p1 pdh) (p2 pdw)" -- token (2h + i%2, 2w + j%2)
 *     takes D[i][j] at feature (i/2)*4 + j/2 (luma; chroma: 4h + i%4, 4w + j%4, (i/4)*2 + j/4)
 *     models/swinv2.py:505-576, models/plainvit.py:50-88.  INT16_PLANES: [n][(32*32 + 2*16*16)*64]. */
#define RGBNM_K0_LAYOUT_VIT16 0
#define RGBNM_K0_LAYOUT_SWIN4 1
/*   RGBNM_K0_LAYOUT_VIT16_NOSUB: `--no_subblock` (PatchEmbedding_DCT_Group with use_subblock = False, plainvit.py:173-216): the VIT16
 *     geometry without the A16 products, out [n][196][384] = per 16 x 16 patch [Y 256 | Cb 64 | Cr 64] where the luma part is the
 *     tile of the four un-converted 8 x 8 blocks, row-major ('b c (h pdh) (w pdw) p1 p2 -> b c h w (pdh p1) (pdw p2)') */
#define RGBNM_K0_LAYOUT_VIT16_NOSUB 2
int rgbnm_k0_dcstats_ex(const int16_t* y, const int16_t* cbcr, const int16_t* quant, const rgbnm_plan* plans,
                        const rgbnm_k0_tables* tables, float* stats, int n, int hb, int wb, int layout, void* stream);
int rgbnm_k0_fused_ex(const int16_t* y, const int16_t* cbcr, const int16_t* quant, const rgbnm_plan* plans,
                      const rgbnm_k0_tables* tables, const float* stats, void* out, int out_mode, int layout, int n,
                      int hb, int wb, void* stream);


/* ------------------------------------------------------------------------------
 * (K1,K3,K5,K6,K8) Dense bf16 contractions of models/plainvit.py on tcgen05 tensor cores.
 * Logical problem: C[M,N] = sum_k A[m,k] * B[n,k]  (fp32 accumulation in TMEM).
 *   nn.Linear forward  (plainvit.py:194,441,443,485-490): A = activations [M,K], B = weight [N,K]
 *   dgrad              : A = dY [M,N'], B = transposed bf16 weight copy
 *   wgrad (RGBNM_EPI_WGRAD_ATOMIC): A and B are stored [K][M] and [K][N] (reduction index = token
 *                        = slowest), out_f32[M][N] += alpha * C, split over `splits` CTAs
 * All device pointers; bf16 operands; leading dimensions in elements, multiples of 8.
 * ---------------------------------------------------------------------------- */
enum rgbnm_epilogue {
    RGBNM_EPI_STORE = 0,        /* C = acc (+ bias)                          bf16 */
    RGBNM_EPI_RESIDUAL = 1,     /* C = acc + bias + aux                      plainvit.py:475-479 */
    RGBNM_EPI_GELU = 2,         /* C = acc + bias, C2 = gelu_erf(C)          plainvit.py:485-487 */
    RGBNM_EPI_DGELU = 3,        /* C = acc * gelu_erf'(aux)                  backward of the above */
    RGBNM_EPI_POSEMB = 4,       /* C = acc + bias + posemb[row % period]     plainvit.py:194-198, 97-121 */
    RGBNM_EPI_WGRAD_ATOMIC = 5, /* out_f32 += alpha * acc                    fp32 red.add */
    RGBNM_EPI_F32 = 6,          /* out_f32 = acc + bias                      fp32 (logits) */
    RGBNM_EPI_GELU_ACT = 7,     /* C = gelu_erf(acc + bias)                  inference: the pre-activation is not kept (swinv2.py:30-31) */
    RGBNM_EPI_LNRES = 8,        /* C = aux + LayerNorm_N(acc + bias) * ln_gamma + ln_beta   post-norm residual, swinv2.py:302-306;
                                   N <= 384, N % 32 == 0 (the row statistics are taken inside one output tile) */
    RGBNM_EPI_LN = 9            /* C = LayerNorm_N(acc + bias) * ln_gamma + ln_beta         patch_embed.norm (swinv2.py:568), PatchMerging.norm (:361) */
};

typedef struct {
    const void* A; const void* B;
    void* C; void* C2; const void* aux;
    const float* bias; const float* posemb; float* out_f32;
    long long lda, ldb, ldc, ldaux, ldo;
    int M, N, K;
    int epilogue;               /* enum rgbnm_epilogue */
    int pos_period;
    int splits;                 /* WGRAD_ATOMIC: split count over the reduction; <= 0 = chosen by the library */
    float alpha;
    int trans_out;              /* WGRAD_ATOMIC: out_f32[n * ldo + m] += ... (lets the caller put the longer side on M) */
    int perm_heads;             /* WGRAD_ATOMIC: > 0: the M index is in the kernel's qkv order (q|k|v head-major, rgbnm_weight_prep); */
    int perm_head_dim;          /*   results are added at the reference row h*3D + d*3 + which ("(h d qkv)", plainvit.py:447) */
    const float* ln_gamma;      /* LNRES: fp32 [N], 16-byte aligned */
    const float* ln_beta;
    float ln_eps;
} rgbnm_gemm_args;

int rgbnm_gemm_bf16(const rgbnm_gemm_args* args, void* stream);


/* ------------------------------------------------------------------------------
 * (K2, K7, K8, a31) Memory-bound ViT kernels around the contractions.
 * ---------------------------------------------------------------------------- */
/* nn.LayerNorm(emb, eps) over the last dimension (plainvit.py:513,522,551); x, y bf16 [rows][emb],
 * gamma/beta fp32, statistics fp32 [rows] kept for backward.  emb in {192, 384, 768}. */
int rgbnm_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                        int rows, int emb, float eps, void* stream);
/* dx = LN'(dy) (+ dres if non-NULL: the residual branch gradient, plainvit.py:475-479); dgamma/dbeta +=; if dxsum is
 * non-NULL, dxsum[emb] += column sums of the stored dx (= the bias gradient of the Linear that wrote the residual stream,
 * saving a separate pass over dx) */
int rgbnm_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                        const void* dres, void* dx, float* dgamma, float* dbeta, float* dxsum, int rows, int emb, void* stream);
/* The same with broadcast upstream gradients: dy holds rows / rows_per_dy_row rows and dy row r / rows_per_dy_row serves row r
 * (the token mean of the classification head, plainvit.py:551: every token of an image receives d(pooled) / tokens). */
int rgbnm_layernorm_bwd_ex(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                           const void* dres, void* dx, float* dgamma, float* dbeta, float* dxsum, int rows, int emb,
                           int rows_per_dy_row, void* stream);
/* RandomMixup_DCT on the bf16 embed input (utils/cls_transforms.py:135-182): out[b] = lam[0] * x[b] + lam[1] * x[(b-1) mod batch];
 * lam = 2 device floats; per_image = elements per image (multiple of 8); out != x. */
int rgbnm_mixup_bf16(const void* x, void* out, const float* lam, int batch, long long per_image, void* stream);
/* out[cols] += column sums of a bf16 matrix (bias gradients of nn.Linear).  qkv_heads > 0: the columns are in the kernel's
 * qkv order and column c is added at its reference position (see perm_heads above). */
int rgbnm_colsum_bf16(const void* a, long long ld, int rows, int cols, float* out, int qkv_heads, int head_dim, void* stream);
/* fp32 master weight [n][k] -> bf16 working copy (rows regrouped q|k|v head-major when qkv_heads > 0,
 * undoing the "(h d qkv)" interleave of plainvit.py:447) and, if wt_bf16 != NULL, its transpose [k][n] */
int rgbnm_weight_prep(const float* w, int n, int k, int qkv_heads, int head_dim, void* w_bf16, void* wt_bf16, void* stream);
/* The same for every Linear of the model in one launch.  descs_dev: device array; first_tile = running sum of
 * ceil(n/32) * ceil(k/32) over the preceding descriptors (total_tiles = the sum over all); bias / bias_k (may be NULL):
 * qkv bias regrouped into kernel order alongside. */
typedef struct {
    const float* w; void* w_bf16; void* wt_bf16; const float* bias; float* bias_k;
    int32_t n, k, qkv_heads, head_dim, first_tile, pad;
} rgbnm_wprep_desc;
int rgbnm_weight_prep_batch(const rgbnm_wprep_desc* descs_dev, int n_desc, int total_tiles, void* stream);
int rgbnm_qkv_perm_vec(const float* src, float* dst, int n, int heads, int head_dim, int inverse, void* stream);
int rgbnm_qkv_unperm_rows_add(const float* src, float* dst, int n, int k, int heads, int head_dim, void* stream);
/* *out += sum(g^2) */
int rgbnm_sumsq_f32(const float* g, long long n, float* out, void* stream);
/* clip_grad_norm_(max_norm) + AdamW(weight_decay=0) + decoupled decay `p -= decay * p` on the first n_decay
 * elements (train.py:163-172, pipeline_utils.py:536, custom_optims.py:37-43).  gnorm_sq = device scalar with
 * sum(g^2) of the unscaled gradient.  hyper = 9 device floats (so that a captured CUDA graph can be replayed with
 * a new learning rate): lr, beta1, beta2, eps, 1-beta1^t, 1-beta2^t, decay (= lr/base_lr*wd), grad_scale
 * (multiplies g first: 1/world after a SUM allreduce), max_norm (<= 0: no clipping). */
int rgbnm_adamw_step(float* p, const float* g, float* m, float* v, long long n, long long n_decay, const float* gnorm_sq,
                     const float* hyper, void* stream);


/* (K4) attention core, forward: o = softmax(q k^T * scale) v per (image, head) (plainvit.py:450-461; the reference's
 * scale is 1/sqrt(emb_size)).  qkv bf16 [B][N][3*H*D] = [q | k | v] head-major, o bf16 [B][N][H*D], lse fp32 [B][H][N]
 * (log-sum-exp of the scaled scores, for backward).  N = 196, D = 64. */
int rgbnm_attention_fwd(const void* qkv, void* o, float* lse, int B, int N, int H, int D, float scale, void* stream);

/* (K8) attention core, backward: dqkv bf16 [B][N][3*H*D] = [dq | dk | dv] from dout bf16 [B][N][H*D]; P is recomputed from
 * qkv and lse; dvec fp32 [B][H][N] is scratch for rowsum(dout * o).  Three launches (prep, dQ kernel, dK/dV kernel). */
int rgbnm_attention_bwd(const void* dout, const void* qkv, const void* o, const float* lse, void* dqkv, float* dvec, int B,
                        int N, int H, int D, float scale, void* stream);


/* ------------------------------------------------------------------------------
 * (a33) SwinV2 DCT forward path (models/swinv2.py).  The dense contractions (qkv / proj / fc1+GELU / fc2 / patch-merging
 * reduction / head) go through rgbnm_gemm_bf16; these are the memory-bound kernels around them.
 * ---------------------------------------------------------------------------- */
/* Post-norm residual of SwinTransformerBlock.forward (swinv2.py:302-306) and plain nn.LayerNorm (patch_embed.norm :568,
 * PatchMerging.norm :361, final norm :696): y = (res ? res : 0) + LayerNorm(x) * gamma + beta; x, res, y bf16 [rows][emb],
 * gamma/beta fp32; emb even, <= 1536. */
int rgbnm_layernorm_res_fwd(const void* x, const float* gamma, const float* beta, const void* res, void* y, int rows,
                            int emb, float eps, void* stream);
/* WindowAttention.forward core (swinv2.py:152-177) on tokens kept in image order: window partition, cyclic shift by
 * `shift` and their inverses (swinv2.py:39-66, 283-300) are folded into the kernel's gather / scatter.
 *   qkv   bf16 [B*H*W][3*C], columns (which, head, d) as after the qkv Linear (swinv2.py:153-155)
 *   out   bf16 [B*H*W][C], columns (head, d) = input of the proj Linear
 *   bias  fp32 [heads][window^2][window^2] = 16 * sigmoid(cpb_mlp(relative_coords_table))[relative_position_index]
 *   scale fp32 [heads] = exp(min(logit_scale, log(100)))                            (swinv2.py:158-168)
 * scores = normalize(q) . normalize(k)^T * scale + bias (+ -100 between different shift regions, swinv2.py:227-242),
 * softmax, . v.  Round 1: window = 8, head dimension C / heads = 32 (SwinV2-T, utils/configs.py:123-137). */
int rgbnm_window_attention_fwd(const void* qkv, void* out, const float* bias, const float* scale, int B, int H, int W,
                               int C, int heads, int window, int shift, void* stream);
/* PatchMerging gather (swinv2.py:353-358): out bf16 [B][H/2][W/2][4*C] = [x(2h,2w) | x(2h+1,2w) | x(2h,2w+1) | x(2h+1,2w+1)] */
int rgbnm_patch_merge_gather(const void* x, void* out, int B, int H, int W, int C, void* stream);
/* AdaptiveAvgPool1d(1) over tokens (swinv2.py:697-699): out bf16 [B][C] = mean_l x[b][l][c] */
int rgbnm_token_mean_bf16(const void* x, void* out, int B, int L, int C, void* stream);

/* ---- SwinV2 training (first correct CUDA versions of the backward kernels; fp32 arithmetic on the CUDA cores) ---- */
/* Post-norm residual with stochastic depth (swinv2.py:302-306, DropPath): y = (res ? res : 0) + s * (LayerNorm(x) * gamma + beta),
 * s = row_scale ? row_scale[row / rows_per_scale] : 1 (row_scale: fp32 per image = mask / keep_prob).  emb even, <= 768. */
int rgbnm_layernorm_res_scaled_fwd(const void* x, const float* gamma, const float* beta, const void* res, const float* row_scale,
                                   int rows_per_scale, void* y, int rows, int emb, float eps, void* stream);
/* Backward of the LayerNorm branch of the above: dx bf16 = d loss / d x from dy (= d loss / d y; the residual's gradient is
 * dy itself), dgamma / dbeta fp32 [emb] +=.  Statistics are recomputed from x. */
int rgbnm_layernorm_res_bwd(const void* dy, const void* x, const float* gamma, const float* row_scale, int rows_per_scale,
                            void* dx, float* dgamma, float* dbeta, int rows, int emb, float eps, void* stream);
/* The same with `dxsum` (emb floats, accumulated; may be NULL): column sums of the dx it writes = the bias gradient of the Linear
 * whose output the LayerNorm takes (swinv2.py:302-306: fc2 / attn.proj), so no separate rgbnm_colsum_bf16 launch.  dxsum needs
 * emb in {96, 192, 384, 768}. */
int rgbnm_layernorm_res_bwd_ex(const void* dy, const void* x, const float* gamma, const float* row_scale, int rows_per_scale,
                               void* dx, float* dgamma, float* dbeta, float* dxsum, int rows, int emb, float eps, void* stream);
/* Backward of rgbnm_window_attention_fwd: dqkv bf16 [B*H*W][3*C] from dout bf16 [B*H*W][C] (P is recomputed from qkv);
 * dbias fp32 [heads][64][64] += d loss / d bias tile (the caller differentiates 16 * sigmoid(cpb_mlp(.))[index] through it),
 * dscale fp32 [heads] += d loss / d scale (scale = exp(min(logit_scale, log 100))). */
int rgbnm_window_attention_bwd(const void* qkv, const void* dout, const float* bias, const float* scale, void* dqkv, float* dbias,
                               float* dscale, int B, int H, int W, int C, int heads, int window, int shift, void* stream);
/* Inverse of rgbnm_patch_merge_gather: dx bf16 [B][H][W][C] from dy bf16 [B][H/2][W/2][4*C] */
int rgbnm_patch_merge_scatter(const void* dy, void* dx, int B, int H, int W, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RGBNM_B200_H */
