// K0: the fused DCT-domain data-path kernel (sm_100a).
//
// One launch replaces, per image (reference file:line):
//   dequantise + clamp                     datasets.py:288-293
//   crop                                   utils/dct_ops.py:584-599
//   resize (x2 up / identity / x2 down)    utils/dct_ops.py:436-580
//   RandomFlip_DCT                         utils/custom_transforms.py:913-942
//   RandAugment_dct ops + per-op clamp     utils/custom_transforms.py:944-1127
//   ToRange(-1,1,-1024,1016)               utils/custom_transforms.py:406-466
//   rearrange + sub-block conversion + collapse + concat   models/plainvit.py:200-216
// and writes the (B,196,384) operand of the patch-projection GEMM.
//
// Work decomposition: one warp = one pair of horizontally adjacent tokens = 8 luma + 4
// chroma post-resize blocks.  Everything between the global loads and the global stores
// stays in registers / warp-private shared memory; the only barriers are __syncwarp().
//   row pass    : lane = one 16-coefficient source row (two 16-byte loads), dequantise,
//                 1-D transform along the row  -> R (smem)
//   column pass : lane = one column of one output block: 1-D transform down the column,
//                 round, flip/RandAugment ops, ToRange  -> luma 16x16 tile S (smem) /
//                 chroma straight to HBM
//   sub-block   : A16 . S . A16^T on the two 16x16 luma tiles, rows then columns -> HBM
// Geometric ops (flip, translate, rot90, cutout, chroma drop) cost nothing: they are
// resolved per block by walking the plan backwards to the source block (trace_back) and
// by tracking one in-block "transposed" flag going forwards.
// The kernel is HBM-bound by design (DESIGN.md "K0 roofline"); tensor cores are not used.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/rgbnm_b200.h"
#include "common.cuh"
#include "k0_common.cuh"

namespace k0 {

constexpr int WARPS_PER_CTA = 4;
constexpr int R_STRIDE = 136;        // floats per block in R: 16 rows x 8 + 8 pad (bank spread)
constexpr int S_LD = 17;             // padded leading dimension of the 16x16 tiles
constexpr int PAIRS_PER_IMAGE = 98;  // 14 token rows x 7 token pairs

struct __align__(16) WarpSmem {
    float R[12 * R_STRIDE];   // row-pass output; re-used as tile T in the sub-block pass
    float S[2 * 16 * S_LD];   // luma tiles in logical block coordinates, ToRange'd
    float qf[3 * 64];         // fp32 quantisation tables of the current image
    rgbnm_plan plan;
    int info[12];
    int img;
};

__device__ __forceinline__ int pack_info(int sr, int sc, int child_r, int child_c, int zero) {
    return (sr & 0xff) | ((sc & 0xff) << 8) | (child_r << 16) | (child_c << 17) | ((zero + 1) << 18);
}

// Apply flip + RandAugment ops to the 8 values (physical column c, rows 0..7) of one block.
// `comp`: 0 Y, 1 Cb, 2 Cr.  Returns the final "transposed" flag.
__device__ __forceinline__ bool run_ops(float (&v)[8], const rgbnm_plan& pl, int comp, int c, int zero,
                                        const rgbnm_k0_tables& tb, const float* __restrict__ stats) {
    bool T = false;
    int start = 0;
    if (zero >= 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.0f;
        start = zero + 1;
    } else if (pl.flip) {
        if (c & 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = -v[i];
        }
    }
    if (pl.train) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = clampf(v[i]);
    }
    // the transpose flag depends on every rot90 in the list, also those before `start`
    for (int k = 0; k < start; ++k) T ^= (pl.ops[k].code == RGBNM_OP_ROT90);
    for (int k = start; k < pl.n_ops; ++k) {
        const rgbnm_plan_op& op = pl.ops[k];
        const int code = op.code;
        bool touched = true;
        if (code == RGBNM_OP_ROT90) {
            T = !T;
            // ccw: negate logical odd rows; cw: negate logical odd columns (dct_ops.py:116-128)
            const bool rows = op.p[0] > 0;
            const bool by_phys_col = (rows == T);   // logical row == physical column iff T
            if (by_phys_col) {
                if (c & 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = -v[i];
                }
            } else {
#pragma unroll
                for (int i = 1; i < 8; i += 2) v[i] = -v[i];
            }
        } else if (code == RGBNM_OP_BRIGHTNESS) {
            if (comp == 0 && c == 0) v[0] = rint_magic(v[0] + stats[2 * k]);
        } else if (code == RGBNM_OP_CONTRAST) {
            if (comp == 0 && c == 0) v[0] = rint_magic(v[0] * op.f);
        } else if (code == RGBNM_OP_COLOR) {
            if (comp != 0 && c == 0) v[0] = rint_magic(v[0] * op.f);
        } else if (code == RGBNM_OP_AUTOCONTRAST || code == RGBNM_OP_AUTOSATURATION) {
            const bool mine = (code == RGBNM_OP_AUTOCONTRAST) ? (comp == 0) : (comp != 0);
            const float lo = stats[2 * k], hi = stats[2 * k + 1];
            if (mine && c == 0 && lo != hi) {
                const float z = __fdiv_rn(v[0] - lo, hi - lo);
                v[0] = rint_magic(CLAMP_LO + z * (CLAMP_HI - CLAMP_LO));
            }
        } else if (code == RGBNM_OP_POSTERIZE) {
            if (c == 0) v[0] = float(tb.posterize_lut[op.p[0] * 2048 + int(v[0]) + 1024]);
        } else if (code == RGBNM_OP_SHARPNESS || code == RGBNM_OP_MIDFREQ) {
            if (comp == 0) {
                const float* F = tb.filters + op.p[0] * 64 + c;   // symmetric: F[i][c] == F[c][i]
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = rint_magic(clampf(v[i] * __ldg(F + 8 * i)));
            }
        } else if (code == RGBNM_OP_SOLARIZE_ADD) {
            if (comp == 0 && c == 0 && v[0] < 0.0f) v[0] += float(op.p[0]);
        } else if (code == RGBNM_OP_INVERT) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = -v[i];
        } else {
            touched = false;   // geometric / zeroing ops: handled by trace_back
        }
        (void)touched;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = clampf(v[i]);   // custom_transforms.py:1019-1020
    }
    return T;
}

__device__ __forceinline__ float to_range(float v) {
    // ToRange: ((x - (-1024)) / 2040) * 2 + (-1), same rounding sequence as the reference
    const float z = __fdiv_rn(v + 1024.0f, 2040.0f);
    return -1.0f + z * 2.0f;
}

template <int OUT_MODE>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k0_fused_kernel(const int16_t* __restrict__ y, const int16_t* __restrict__ cbcr, const int16_t* __restrict__ quant,
                const rgbnm_plan* __restrict__ plans, rgbnm_k0_tables tb, const float* __restrict__ stats_all,
                void* __restrict__ out_, int n_images, int hb, int wb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSmem& ws = reinterpret_cast<WarpSmem*>(smem_raw)[warp];
    if (lane == 0) ws.img = -1;
    __syncwarp();

    const int total = n_images * PAIRS_PER_IMAGE;
    const int hc = hb >> 1, wc = wb >> 1;
    for (int item = blockIdx.x * WARPS_PER_CTA + warp; item < total; item += gridDim.x * WARPS_PER_CTA) {
        const int img = item / PAIRS_PER_IMAGE;
        const int rem = item - img * PAIRS_PER_IMAGE;
        const int th = rem / 7, tp = rem - th * 7;

        // ---- per-image state -----------------------------------------------------------------
        if (ws.img != img) {
            __syncwarp();
            const int* psrc = reinterpret_cast<const int*>(plans + img);
            int* pdst = reinterpret_cast<int*>(&ws.plan);
            if (lane < int(sizeof(rgbnm_plan) / 4)) pdst[lane] = __ldg(psrc + lane);
            for (int k = lane; k < 192; k += 32) ws.qf[k] = float(__ldg(quant + size_t(img) * 192 + k));
            if (lane == 0) ws.img = img;
            __syncwarp();
        }
        const rgbnm_plan& pl = ws.plan;
        const int mode = mode_of(pl.crop_size);
        const float* stats = stats_all + size_t(img) * RGBNM_MAX_OPS * 2;

        // ---- block bookkeeping: lanes 0..11 trace one block each ------------------------------
        if (lane < 12) {
            const int b = lane;
            int comp, r, c;
            if (b < 8) {
                comp = 0;
                r = 2 * th + ((b & 3) >> 1);
                c = 2 * (2 * tp + (b >> 2)) + (b & 1);
            } else {
                comp = 1 + ((b - 8) & 1);
                r = th;
                c = 2 * tp + ((b - 8) >> 1);
            }
            const Trace t = trace_back(pl, comp, r, c);
            const int ci = comp == 0 ? pl.crop_i : (pl.crop_i >> 1);
            const int cj = comp == 0 ? pl.crop_j : (pl.crop_j >> 1);
            int sr, sc, chr = 0, chc = 0;
            if (mode == MODE_DOWN2) { sr = ci + 2 * t.r; sc = cj + 2 * t.c; }
            else if (mode == MODE_IDENT) { sr = ci + t.r; sc = cj + t.c; }
            else { sr = ci + (t.r >> 1); sc = cj + (t.c >> 1); chr = t.r & 1; chc = t.c & 1; }
            ws.info[b] = pack_info(sr, sc, chr, chc, t.zero);
        }
        __syncwarp();

        // ---- row pass ---------------------------------------------------------------------------
        const size_t y_img = size_t(img) * hb * wb * 64;
        const size_t c_img = size_t(img) * 2 * hc * wc * 64;
        if (mode == MODE_DOWN2) {
#pragma unroll 2
            for (int k = 0; k < 6; ++k) {
                const int b = 2 * k + (lane >> 4), i = lane & 15;
                const int inf = ws.info[b];
                if (((inf >> 18) & 7) != 0) continue;   // zeroed block: nothing to load
                const int comp = b < 8 ? 0 : 1 + ((b - 8) & 1);
                const int16_t* base = comp == 0 ? y + y_img : cbcr + c_img + size_t(comp - 1) * hc * wc * 64;
                const int W = comp == 0 ? wb : wc;
                const int srow = (inf & 0xff) + (i >> 3), scol = (inf >> 8) & 0xff;
                const int4* p = reinterpret_cast<const int4*>(base + (size_t(srow) * W + scol) * 64 + (i & 7) * 8);
                const int4 ra = __ldg(p), rb = __ldg(p + 8);
                const float* q = ws.qf + comp * 64 + (i & 7) * 8;
                float xl[8], xr[8], o[8];
                if (pl.clamp_in) { dequant8<true>(ra, q, xl); dequant8<true>(rb, q, xr); }
                else             { dequant8<false>(ra, q, xl); dequant8<false>(rb, q, xr); }
                down2_1d<1, 1>(xl, xr, o);
                float4* dst = reinterpret_cast<float4*>(ws.R + b * R_STRIDE + i * 8);
                dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                dst[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int b = 4 * k + (lane >> 3), i = lane & 7;
                const int inf = ws.info[b];
                if (((inf >> 18) & 7) != 0) continue;
                const int comp = b < 8 ? 0 : 1 + ((b - 8) & 1);
                const int16_t* base = comp == 0 ? y + y_img : cbcr + c_img + size_t(comp - 1) * hc * wc * 64;
                const int W = comp == 0 ? wb : wc;
                const int srow = inf & 0xff, scol = (inf >> 8) & 0xff;
                const int4 ra = __ldg(reinterpret_cast<const int4*>(base + (size_t(srow) * W + scol) * 64 + i * 8));
                const float* q = ws.qf + comp * 64 + i * 8;
                float x[8], o[8];
                if (pl.clamp_in) dequant8<true>(ra, q, x); else dequant8<false>(ra, q, x);
                if (mode == MODE_UP2) {
                    up2_1d<1>(x, (inf >> 17) & 1, o);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = x[j];
                }
                float4* dst = reinterpret_cast<float4*>(ws.R + b * R_STRIDE + i * 8);
                dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                dst[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
        }
        __syncwarp();

        // ---- column pass + ops + ToRange ---------------------------------------------------------
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            const int b = 4 * k + (lane >> 3), c = lane & 7;
            const int inf = ws.info[b];
            const int zero = ((inf >> 18) & 7) - 1;
            const int comp = b < 8 ? 0 : 1 + ((b - 8) & 1);
            float v[8];
            if (zero < 0) {
                const float* col = ws.R + b * R_STRIDE + c;
                if (mode == MODE_DOWN2) {
                    float xl[8], xr[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { xl[i] = col[8 * i]; xr[i] = col[8 * (i + 8)]; }
                    down2_1d<1, 2>(xl, xr, v);
                } else if (mode == MODE_UP2) {
                    float x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = col[8 * i];
                    up2_1d<2>(x, (inf >> 16) & 1, v);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = col[8 * i];
                }
                if (mode != MODE_IDENT) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = rint_magic(v[i]);   // torch.round -> int16 (dct_ops.py:577-578)
                }
            }
            const bool T = run_ops(v, pl, comp, c, zero, tb, stats);

            if (OUT_MODE == RGBNM_K0_OUT_INT16_PLANES) {
                int16_t* o16 = reinterpret_cast<int16_t*>(out_) + size_t(img) * PLANE_ELEMS;
                int blk;
                if (b < 8) blk = (2 * th + ((b & 3) >> 1)) * GRID_Y + 2 * (2 * tp + (b >> 2)) + (b & 1);
                else blk = GRID_Y * GRID_Y + ((comp - 1) * GRID_C + th) * GRID_C + 2 * tp + ((b - 8) >> 1);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int li = T ? c : i, lj = T ? i : c;
                    o16[size_t(blk) * 64 + li * 8 + lj] = int16_t(int(v[i]));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = to_range(v[i]);
                if (b < 8) {
                    float* tile = ws.S + (b >> 2) * 16 * S_LD;
                    const int r0 = ((b & 3) >> 1) * 8, c0 = (b & 1) * 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int li = T ? c : i, lj = T ? i : c;
                        tile[(r0 + li) * S_LD + c0 + lj] = v[i];
                    }
                } else {
                    const int token = th * 14 + 2 * tp + ((b - 8) >> 1);
                    const size_t o = (size_t(img) * TOKENS + token) * FEAT + 256 + (comp - 1) * 64;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int li = T ? c : i, lj = T ? i : c;
                        if (OUT_MODE == RGBNM_K0_OUT_F32) reinterpret_cast<float*>(out_)[o + li * 8 + lj] = v[i];
                        else reinterpret_cast<__nv_bfloat16*>(out_)[o + li * 8 + lj] = __float2bfloat16_rn(v[i]);
                    }
                }
            }
        }
        __syncwarp();

        // ---- sub-block conversion: A16 . S . A16^T per token (plainvit.py:50-69) -----------------
        if (OUT_MODE != RGBNM_K0_OUT_INT16_PLANES) {
            const int tok = lane >> 4, rc = lane & 15;
            float* Tt = ws.R + tok * 16 * S_LD;          // R is dead: reuse as T
            {
                const float* row = ws.S + tok * 16 * S_LD + rc * S_LD;
                float xl[8], xr[8], o[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) { xl[j] = row[j]; xr[j] = row[8 + j]; }
                a16_1d(xl, xr, o);                        // S . A16^T  (along the row)
#pragma unroll
                for (int j = 0; j < 16; ++j) Tt[rc * S_LD + j] = o[j];
            }
            __syncwarp();
            {
                float xl[8], xr[8], o[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) { xl[i] = Tt[i * S_LD + rc]; xr[i] = Tt[(i + 8) * S_LD + rc]; }
                a16_1d(xl, xr, o);                        // A16 . (S A16^T)  (down the column)
                const int token = th * 14 + 2 * tp + tok;
                const size_t ob = (size_t(img) * TOKENS + token) * FEAT;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (OUT_MODE == RGBNM_K0_OUT_F32) reinterpret_cast<float*>(out_)[ob + i * 16 + rc] = o[i];
                    else reinterpret_cast<__nv_bfloat16*>(out_)[ob + i * 16 + rc] = __float2bfloat16_rn(o[i]);
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace k0

extern "C" int rgbnm_k0_fused(const int16_t* y, const int16_t* cbcr, const int16_t* quant, const rgbnm_plan* plans,
                              const rgbnm_k0_tables* tables, const float* stats, void* out, int out_mode, int n,
                              int hb, int wb, void* stream) {
    using namespace k0;
    if (!y || !cbcr || !quant || !plans || !tables || !stats || !out || n < 0) return RGBNM_ERR_ARG;
    if (hb < 2 || wb < 2 || hb > 255 || wb > 255 || (hb & 1) || (wb & 1)) return RGBNM_ERR_ARG;
    if (n == 0) return RGBNM_OK;
    static int num_sms = 0;
    const size_t smem = sizeof(WarpSmem) * WARPS_PER_CTA;
    if (num_sms == 0) {
        int dev = 0;
        RGBNM_CUDA_CHECK(cudaGetDevice(&dev));
        RGBNM_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        RGBNM_CUDA_CHECK(cudaFuncSetAttribute(k0_fused_kernel<RGBNM_K0_OUT_F32>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        RGBNM_CUDA_CHECK(cudaFuncSetAttribute(k0_fused_kernel<RGBNM_K0_OUT_BF16>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        RGBNM_CUDA_CHECK(cudaFuncSetAttribute(k0_fused_kernel<RGBNM_K0_OUT_INT16_PLANES>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    }
    const int items = n * PAIRS_PER_IMAGE;
    const int ctas_needed = (items + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const int grid = ctas_needed < num_sms * 5 ? ctas_needed : num_sms * 5;   // 5 resident CTAs per SM
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 block(WARPS_PER_CTA * 32);
    if (out_mode == RGBNM_K0_OUT_F32)
        k0_fused_kernel<RGBNM_K0_OUT_F32><<<grid, block, smem, st>>>(y, cbcr, quant, plans, *tables, stats, out, n, hb, wb);
    else if (out_mode == RGBNM_K0_OUT_BF16)
        k0_fused_kernel<RGBNM_K0_OUT_BF16><<<grid, block, smem, st>>>(y, cbcr, quant, plans, *tables, stats, out, n, hb, wb);
    else if (out_mode == RGBNM_K0_OUT_INT16_PLANES)
        k0_fused_kernel<RGBNM_K0_OUT_INT16_PLANES><<<grid, block, smem, st>>>(y, cbcr, quant, plans, *tables, stats, out, n, hb, wb);
    else
        return RGBNM_ERR_ARG;
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

extern "C" int rgbnm_k0_launch_count(void) { return 2; }
