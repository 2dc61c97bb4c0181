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
// Work decomposition: one warp = one token row of one image (7 pairs of horizontally
// adjacent tokens, processed one after the other; per-image state -- plan, fp32 quantisation
// tables -- is loaded once per warp).  One pair = 8 luma + 4 chroma post-resize blocks.
// Everything between the global loads and the global stores stays in registers and
// warp-private shared memory; the only barriers are __syncwarp().
//   P1 row pass  : lane = one source coefficient row (one or two 16-byte loads), exact
//                  I2F-free dequantisation, 1-D resize transform along the row -> RY / RC
//   P2 col pass  : luma: lane = one of the 2 x 16 tile columns: 1-D resize transform down
//                  the column for both stacked blocks, round, flip/RandAugment ops, ToRange,
//                  then immediately the column half of the sub-block conversion (A16 . S)
//                  -> TT.  chroma: lane = one column of one block -> CT
//   P3 out pass  : luma: lane = one row of (A16 . S): the row half ((.) A16^T), 16 contiguous
//                  outputs -> two (bf16) / four (fp32) 16-byte stores.  chroma: lane = one
//                  block row -> one / two 16-byte stores.
// Geometric ops (flip, translate, rot90, cutout, chroma drop) cost nothing: they are resolved
// per block by walking the plan backwards to the source block (trace_back) and by tracking one
// in-block "transposed" flag going forwards.
// The kernel is HBM-bound by design (DESIGN.md "K0 roofline"); tensor cores are not used.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "../../include/rgbnm_b200.h"
#include "common.cuh"
#include "k0_common.cuh"

namespace k0 {

constexpr int WARPS_PER_CTA = 1;
// warp tasks per image = chroma block rows (ViT: 14 token rows; Swin: 16 rows of 4 x 64 tokens); one task walks
// GRID_C / 2 "pairs" (pair = 2 horizontally adjacent chroma positions = 8 luma + 4 chroma post-resize blocks)
constexpr int RY_TOK = 32 * 16 + 16;     // floats per token in RY (+16: the two tokens land on disjoint banks)
constexpr int RC_BLK = 136;              // floats per chroma block in RC (16 rows x 8, +8 bank spread)
constexpr int TT_LD = 20;                // leading dimension of the 16x16 tiles in TT (16-byte rows, conflict-free)
constexpr int TT_TOK = 16 * TT_LD + 16;
constexpr int CT_LD = 12;
constexpr int CT_BLK = 8 * CT_LD + 8;

constexpr int SW_RUN = 8 * 24;           // Swin: one output run = 8 consecutive tokens x 24 features
template <int LAYOUT>
struct __align__(16) WarpSmemT {
    float RY[2 * RY_TOK];     // luma row-pass output; re-used as TT (A16 . S) after the column pass
    float RC[4 * RC_BLK];     // chroma row-pass output; re-used as CT (finished chroma blocks)
    float qf[3 * 64];         // fp32 quantisation tables of the current image
    float cq[3 * 64];         // -(2^23 + 2^15) * qf (dequant8)
    rgbnm_plan plan;
    int info[12];
    int pad[4];
    // Swin only: the pair's 4 token rows x 8 tokens x 24 features, staged so that the global stores are contiguous runs
    float OUTS[LAYOUT == RGBNM_K0_LAYOUT_SWIN4 ? 4 * SW_RUN : 4];
};
static_assert(2 * TT_TOK <= 2 * RY_TOK && 4 * CT_BLK <= 4 * RC_BLK, "aliased tiles must fit");

__device__ __forceinline__ int pack_info(int sr, int sc, int child_r, int child_c, int zero) {
    return (sr & 0xff) | ((sc & 0xff) << 8) | (child_r << 16) | (child_c << 17) | ((zero + 1) << 18);
}

__device__ __forceinline__ float clamp_hi(float x) { return fminf(x, CLAMP_HI); }

// Apply flip + RandAugment ops to the 8 values (physical column c, rows 0..7) of one block.
// `comp`: 0 Y, 1 Cb, 2 Cr.  Returns the final "transposed" flag.
// Clamp bookkeeping (custom_transforms.py:1019-1020 clamps both planes after every op): after
// the entry clamp every value is inside [-1024, 1016], so the per-op clamp can only act on
// values the op has just changed -- negations (-(-1024) = 1024) and the DC term.
__device__ __forceinline__ bool run_ops(float (&v)[8], const rgbnm_plan& pl, int comp, int c, int zero,
                                        const rgbnm_k0_tables& tb, const float* __restrict__ stats, int img, int fr, int fc,
                                        int grid_y) {
    bool T = false;
    int start = 0;
    if (zero >= 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.0f;
        start = zero + 1;
    } else {
        if (pl.flip && (c & 1)) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = -v[i];
        }
        if (pl.train) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = clampf(v[i]);     // custom_transforms.py:1107-1108
        }
    }
    // the transpose flag depends on every rot90 in the list, also those before `start`
    for (int k = 0; k < start; ++k) T ^= (pl.ops[k].code == RGBNM_OP_ROT90);
    for (int k = start; k < pl.n_ops; ++k) {
        const rgbnm_plan_op& op = pl.ops[k];
        const int code = op.code;
        if (code == RGBNM_OP_ROT90) {
            T = !T;
            // ccw: negate logical odd rows; cw: negate logical odd columns (dct_ops.py:116-128)
            const bool rows = op.p[0] > 0;
            const bool by_phys_col = (rows == T);   // logical row == physical column iff T
            if (by_phys_col) {
                if (c & 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = clamp_hi(-v[i]);
                }
            } else {
#pragma unroll
                for (int i = 1; i < 8; i += 2) v[i] = clamp_hi(-v[i]);
            }
        } else if (code == RGBNM_OP_SHARPNESS || code == RGBNM_OP_MIDFREQ) {
            if (comp == 0) {
                const float* F = tb.filters + op.p[0] * 64 + c;   // symmetric: F[i][c] == F[c][i]
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = rint_magic(clampf(v[i] * __ldg(F + 8 * i)));
            }
        } else if (code == RGBNM_OP_INVERT) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = clamp_hi(-v[i]);
        } else if (code == RGBNM_OP_SOLARIZE) {
            // mask plane of this op (luma DC > threshold when the op ran) from the statistics pre-pass; a chroma block follows
            // the luma block at twice its coordinates AT THAT TIME: walk the later geometric ops back from (fr, fc)
            if (tb.equalize_lut != nullptr) {
                int r = fr, cc = fc;
                if (grid_y == 32) position_at_op<32>(pl, comp, k, r, cc); else position_at_op<28>(pl, comp, k, r, cc);
                if (comp != 0) { r *= 2; cc *= 2; }
                const bool inside = r >= 0 && cc >= 0 && r < grid_y && cc < grid_y;
                if (inside && tb.equalize_lut[(size_t(img) * RGBNM_MAX_OPS + k) * 2048 + r * grid_y + cc] != 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = clamp_hi(-v[i]);
                }
            }
        } else if (code == RGBNM_OP_FREQ_ENHANCE) {
            // every coefficient but DCT[0,0] (physical (row 0, column 0) whatever the transpose flag) * f, rounded, clamped
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i != 0 || c != 0) v[i] = clampf(rint_magic(v[i] * op.f));
        } else if (c == 0) {
            // DC-only ops: one value of one lane in eight
            float d = v[0];
            if (code == RGBNM_OP_BRIGHTNESS) {
                if (comp == 0) d = rint_magic(d + stats[2 * k]);
            } else if (code == RGBNM_OP_CONTRAST) {
                if (comp == 0) d = rint_magic(d * op.f);
            } else if (code == RGBNM_OP_COLOR) {
                if (comp != 0) d = rint_magic(d * op.f);
            } else if (code == RGBNM_OP_AUTOCONTRAST || code == RGBNM_OP_AUTOSATURATION) {
                const bool mine = (code == RGBNM_OP_AUTOCONTRAST) ? (comp == 0) : (comp != 0);
                const float lo = stats[2 * k], hi = stats[2 * k + 1];
                if (mine && lo != hi) {
                    const float z = __fdiv_rn(d - lo, hi - lo);
                    d = rint_magic(CLAMP_LO + z * (CLAMP_HI - CLAMP_LO));
                }
            } else if (code == RGBNM_OP_POSTERIZE) {
                d = float(tb.posterize_lut[op.p[0] * 2048 + int(d) + 1024]);
            } else if (code == RGBNM_OP_SOLARIZE_ADD) {
                if (comp == 0 && d < 0.0f) d += float(op.p[0]);
            } else if (code == RGBNM_OP_EQUALIZE) {
                // per-image DC mapping built by the statistics pre-pass (histogram equalisation, dct_ops.py:916-955)
                if (comp == 0 && tb.equalize_lut != nullptr)
                    d = float(tb.equalize_lut[(size_t(img) * RGBNM_MAX_OPS + k) * 2048 + int(d) + 1024]);
            }
            v[0] = clampf(d);
        }
        // geometric / zeroing ops: handled by trace_back
    }
    return T;
}

__device__ __forceinline__ float to_range(float v) {
    // ToRange: ((x - (-1024)) / 2040) * 2 + (-1) with the reference's rounding sequence.  The division
    // n / 2040 is one Newton step on n * fl(1/2040): correctly rounded for every integer n in [0, 2040]
    // (exhaustively checked: tests/test_k0_math.py on the CPU, test_to_range_bit_exact_all_values on the GPU).
    constexpr float R = 1.0f / 2040.0f;
    const float n = v + 1024.0f;
    const float q0 = n * R;
    const float e = fmaf(-q0, 2040.0f, n);
    const float z = fmaf(e, R, q0);
    return fmaf(z, 2.0f, -1.0f);      // z * 2 is exact, so the fused form rounds like mul-then-add
}

template <int OUT_MODE>
__device__ __forceinline__ void store_row16(void* __restrict__ out_, size_t off, const float (&o)[16]) {
    if (OUT_MODE == RGBNM_K0_OUT_F32) {
        float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_) + off);
#pragma unroll
        for (int k = 0; k < 4; ++k) p[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
    } else {
        uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out_) + off);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            __nv_bfloat162 a = __floats2bfloat162_rn(o[8 * k], o[8 * k + 1]), b = __floats2bfloat162_rn(o[8 * k + 2], o[8 * k + 3]);
            __nv_bfloat162 c = __floats2bfloat162_rn(o[8 * k + 4], o[8 * k + 5]), d = __floats2bfloat162_rn(o[8 * k + 6], o[8 * k + 7]);
            p[k] = make_uint4(*reinterpret_cast<unsigned*>(&a), *reinterpret_cast<unsigned*>(&b),
                              *reinterpret_cast<unsigned*>(&c), *reinterpret_cast<unsigned*>(&d));
        }
    }
}

template <int OUT_MODE>
__device__ __forceinline__ void store_row8(void* __restrict__ out_, size_t off, const float4& a, const float4& b) {
    if (OUT_MODE == RGBNM_K0_OUT_F32) {
        float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_) + off);
        p[0] = a;
        p[1] = b;
    } else {
        __nv_bfloat162 x = __floats2bfloat162_rn(a.x, a.y), y = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 z = __floats2bfloat162_rn(b.x, b.y), w = __floats2bfloat162_rn(b.z, b.w);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out_) + off) =
            make_uint4(*reinterpret_cast<unsigned*>(&x), *reinterpret_cast<unsigned*>(&y),
                       *reinterpret_cast<unsigned*>(&z), *reinterpret_cast<unsigned*>(&w));
    }
}

// ---- P1 helpers: one lane-item = one source coefficient row of output block `b` ---------------
// All global loads of a half-pair are issued before the first one is consumed (one DRAM-latency
// exposure per half-pair instead of one per item).
struct RowSrc {
    const int4* p;      // first 16-byte load (nullptr: block is zeroed by an op, nothing to load)
    const float* q;     // fp32 table row (8 entries) and, 192 floats further, its dequant bias row
    int inf;
};

template <int MODE, class WarpSmem>
__device__ __forceinline__ RowSrc row_src(const WarpSmem& ws, int b, int i, const int16_t* __restrict__ y_img,
                                          const int16_t* __restrict__ c_img, int wb, int hc, int wc) {
    RowSrc r;
    r.inf = ws.info[b];
    const int comp = b < 8 ? 0 : 1 + ((b - 8) & 1);
    const int16_t* base = comp == 0 ? y_img : c_img + size_t(comp - 1) * hc * wc * 64;
    const int W = comp == 0 ? wb : wc;
    const int scol = (r.inf >> 8) & 0xff;
    const int srow = (r.inf & 0xff) + (MODE == MODE_DOWN2 ? (i >> 3) : 0);
    r.p = reinterpret_cast<const int4*>(base + (size_t(srow) * W + scol) * 64 + (i & 7) * 8);
    if (((r.inf >> 18) & 7) != 0) r.p = nullptr;
    r.q = ws.qf + comp * 64 + (i & 7) * 8;
    return r;
}

template <int MODE>
__device__ __forceinline__ void row_compute(const RowSrc& r, const int4& ra, const int4& rb, bool clamp, float* dst) {
    if (r.p == nullptr) return;
    float o[8];
    if (MODE == MODE_DOWN2) {
        float xl[8], xr[8];
        dequant8x2(ra, rb, r.q, r.q + 192, clamp, xl, xr);
        down2_1d<1, 1>(xl, xr, o);
    } else {
        float x[8];
        dequant8(ra, r.q, r.q + 192, clamp, x);
        if (MODE == MODE_UP2) {
            up2_1d<1>(x, (r.inf >> 17) & 1, o);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = x[j];
        }
    }
    float4* d4 = reinterpret_cast<float4*>(dst);
    d4[0] = make_float4(o[0], o[1], o[2], o[3]);
    d4[1] = make_float4(o[4], o[5], o[6], o[7]);
}

// luma block b = tok*4 + bi*2 + bj -> RY[tok][bi*YROWS + i][bj*8 ..];  chroma block b -> RC[b-8][i][..]
template <int MODE, class WarpSmem>
__device__ __forceinline__ void row_pass(WarpSmem& ws, int lane, const int16_t* __restrict__ y_img,
                                         const int16_t* __restrict__ c_img, int wb, int hc, int wc, bool clamp) {
    constexpr int YROWS = MODE == MODE_DOWN2 ? 16 : 8;
    constexpr int HALVES = MODE == MODE_DOWN2 ? 2 : 1;
#pragma unroll 1
    for (int h = 0; h < HALVES; ++h) {
        RowSrc src[3];
        float* dst[3];
        int4 ra[3], rb[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            int b, i;
            if (MODE == MODE_DOWN2) {
                i = lane & 15;
                b = j < 2 ? 2 * (2 * h + j) + (lane >> 4) : 8 + 2 * h + (lane >> 4);
            } else {
                i = lane & 7;
                b = j < 2 ? 4 * j + (lane >> 3) : 8 + (lane >> 3);
            }
            src[j] = row_src<MODE>(ws, b, i, y_img, c_img, wb, hc, wc);
            dst[j] = j < 2 ? ws.RY + (b >> 2) * RY_TOK + (((b >> 1) & 1) * YROWS + i) * 16 + (b & 1) * 8
                           : ws.RC + (b - 8) * RC_BLK + i * 8;
            ra[j] = rb[j] = make_int4(0, 0, 0, 0);
            if (src[j].p != nullptr) {
                ra[j] = __ldg(src[j].p);
                if (MODE == MODE_DOWN2) rb[j] = __ldg(src[j].p + 8);
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) row_compute<MODE>(src[j], ra[j], rb[j], clamp, dst[j]);
    }
}

// P2 resize half for one block column: 8 values down physical column `col_ptr` (stride `ld` floats).
__device__ __forceinline__ void col_item(int mode, const float* __restrict__ col_ptr, int ld, int inf, float (&v)[8]) {
    if (mode == MODE_DOWN2) {
        float xl[8], xr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { xl[i] = col_ptr[ld * i]; xr[i] = col_ptr[ld * (i + 8)]; }
        down2_1d<1, 2>(xl, xr, v);
    } else if (mode == MODE_UP2) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = col_ptr[ld * i];
        up2_1d<2>(x, (inf >> 16) & 1, v);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = col_ptr[ld * i];
    }
    if (mode != MODE_IDENT) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = rint_magic(v[i]);   // torch.round -> int16 (dct_ops.py:577-578)
    }
}

// ---- Swin tail of one pair: per-block decomposition D = A^T . X . A (A(4,2) luma, A(2,4) chroma), the reference's
// interleaved "(p1 pdh) (p2 pdw)" read-out, staging of the pair's 4 x 8 tokens x 24 features and contiguous stores.
// On entry: luma tiles S (two units x 16 x 16, 2 x 2 blocks each) at RY + u * TT_TOK (ld TT_LD), finished chroma
// blocks (unit u, comp) at RC + (2u + comp-1) * CT_BLK (ld CT_LD).
template <int OUT_MODE, class WarpSmem>
__device__ __forceinline__ void swin_out(WarpSmem& ws, int lane, int img, int th, int tp, void* __restrict__ out_) {
    // column halves, in place (each lane reads and writes only its own column)
    {
        float* TT = ws.RY + (lane >> 4) * TT_TOK + (lane & 15);
#pragma unroll
        for (int bi = 0; bi < 2; ++bi) {
            float x[8], o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = TT[(8 * bi + i) * TT_LD];
            decomp_y_1d(x, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) TT[(8 * bi + i) * TT_LD] = o[i];
        }
        float* CT = ws.RC + (lane >> 3) * CT_BLK + (lane & 7);
        float x[8], o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = CT[i * CT_LD];
        decomp_c_1d(x, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) CT[i * CT_LD] = o[i];
    }
    __syncwarp();
    // row halves -> staging OUTS[a = token row 0..3][u][b = token col 0..3][24]
    {
        const int u = lane >> 4, r = lane & 15, bi = r >> 3, i = r & 7;
        const float4* src = reinterpret_cast<const float4*>(ws.RY + u * TT_TOK + r * TT_LD);
        const float4 q0 = src[0], q1 = src[1], q2 = src[2], q3 = src[3];
        const float xs[2][8] = {{q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w}, {q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w}};
        const int a = 2 * bi + (i & 1), p1 = i >> 1;
#pragma unroll
        for (int bj = 0; bj < 2; ++bj) {
            float o[8];
            decomp_y_1d(xs[bj], o);
#pragma unroll
            for (int pdw = 0; pdw < 2; ++pdw) {
                float* dst = ws.OUTS + a * SW_RUN + (u * 4 + 2 * bj + pdw) * 24 + p1 * 4;
                *reinterpret_cast<float4*>(dst) = make_float4(o[pdw], o[2 + pdw], o[4 + pdw], o[6 + pdw]);
            }
        }
    }
    {
        const int cb = lane >> 3, i = lane & 7, u = cb >> 1, comp1 = cb & 1;
        const float4* src = reinterpret_cast<const float4*>(ws.RC + cb * CT_BLK + i * CT_LD);
        const float4 q0 = src[0], q1 = src[1];
        const float x[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        float o[8];
        decomp_c_1d(x, o);
        const int a = i & 3, p1 = i >> 2;
#pragma unroll
        for (int pdw = 0; pdw < 4; ++pdw) {
            float* dst = ws.OUTS + a * SW_RUN + (u * 4 + pdw) * 24 + 16 + comp1 * 4 + p1 * 2;
            *reinterpret_cast<float2*>(dst) = make_float2(o[pdw], o[4 + pdw]);
        }
    }
    __syncwarp();
    // contiguous runs: token row 4*th + a, tokens 8*tp .. 8*tp + 7
    const size_t base = ((size_t(img) * 64 + 4 * th) * 64 + 8 * tp) * 24;
    if (OUT_MODE == RGBNM_K0_OUT_F32) {
        float* out = reinterpret_cast<float*>(out_);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int e = k * 32 + lane, a = e / 48, w = e - a * 48;           // 48 float4 per run
            const float4 v = *reinterpret_cast<const float4*>(ws.OUTS + a * SW_RUN + w * 4);
            *reinterpret_cast<float4*>(out + base + size_t(a) * 64 * 24 + w * 4) = v;
        }
    } else {
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(out_);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int e = k * 32 + lane, a = e / 24, w = e - a * 24;           // 24 chunks of 8 bf16 per run
            const float4 v0 = *reinterpret_cast<const float4*>(ws.OUTS + a * SW_RUN + w * 8);
            const float4 v1 = *reinterpret_cast<const float4*>(ws.OUTS + a * SW_RUN + w * 8 + 4);
            __nv_bfloat162 x = __floats2bfloat162_rn(v0.x, v0.y), y = __floats2bfloat162_rn(v0.z, v0.w);
            __nv_bfloat162 z = __floats2bfloat162_rn(v1.x, v1.y), t = __floats2bfloat162_rn(v1.z, v1.w);
            *reinterpret_cast<uint4*>(out + base + size_t(a) * 64 * 24 + w * 8) =
                make_uint4(*reinterpret_cast<unsigned*>(&x), *reinterpret_cast<unsigned*>(&y),
                           *reinterpret_cast<unsigned*>(&z), *reinterpret_cast<unsigned*>(&t));
        }
    }
}

template <int OUT_MODE, int LAYOUT>
__device__ __forceinline__ void process_row(WarpSmemT<LAYOUT>& ws, int lane, int img, int th, const int16_t* __restrict__ y_img,
                                            const int16_t* __restrict__ c_img, const rgbnm_k0_tables& tb,
                                            const float* __restrict__ stats, void* __restrict__ out_, int wb, int hc, int wc) {
    constexpr int GRID_Y = Geo<LAYOUT>::GRID_Y, GRID_C = Geo<LAYOUT>::GRID_C;
    constexpr int TOKENS = Geo<LAYOUT>::TOKENS, FEAT = Geo<LAYOUT>::FEAT, PLANE_ELEMS = Geo<LAYOUT>::PLANE_ELEMS;
    constexpr int PAIRS_PER_ROW = GRID_C / 2;
    constexpr bool SWIN = LAYOUT == RGBNM_K0_LAYOUT_SWIN4;
    const rgbnm_plan& pl = ws.plan;
    const int mode = mode_of(pl.crop_size, GRID_Y);
    const bool clamp = pl.clamp_in != 0;
    const int yrows = mode == MODE_DOWN2 ? 16 : 8;       // row-pass rows per output block
#pragma unroll 1
    for (int tp = 0; tp < PAIRS_PER_ROW; ++tp) {
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
            const Trace t = trace_back<GRID_Y>(pl, comp, r, c);
            const int ci = comp == 0 ? pl.crop_i : (pl.crop_i >> 1);
            const int cj = comp == 0 ? pl.crop_j : (pl.crop_j >> 1);
            int sr, sc, chr = 0, chc = 0;
            if (mode == MODE_DOWN2) { sr = ci + 2 * t.r; sc = cj + 2 * t.c; }
            else if (mode == MODE_IDENT) { sr = ci + t.r; sc = cj + t.c; }
            else { sr = ci + (t.r >> 1); sc = cj + (t.c >> 1); chr = t.r & 1; chc = t.c & 1; }
            ws.info[b] = pack_info(sr, sc, chr, chc, t.zero);
        }
        __syncwarp();

        // ---- P1: row pass -------------------------------------------------------------------------
        if (mode == MODE_DOWN2) row_pass<MODE_DOWN2>(ws, lane, y_img, c_img, wb, hc, wc, clamp);
        else if (mode == MODE_IDENT) row_pass<MODE_IDENT>(ws, lane, y_img, c_img, wb, hc, wc, clamp);
        else row_pass<MODE_UP2>(ws, lane, y_img, c_img, wb, hc, wc, clamp);
        __syncwarp();

        // ---- P2: column pass + ops + ToRange: three rounds of 4 blocks x 8 columns ---------------
        // round 0 / 1: the luma blocks of token 0 / 1 -> tile S (aliases RY), round 2: chroma -> CT (aliases RC)
        bool T = false;
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            const int b = 4 * k + (lane >> 3), c = lane & 7;
            const int inf = ws.info[b];
            const int zero = ((inf >> 18) & 7) - 1;
            const int comp = k < 2 ? 0 : 1 + (b & 1);
            float v[8];
            if (zero < 0) {
                const float* colp = k < 2 ? ws.RY + k * RY_TOK + (((b >> 1) & 1) * yrows) * 16 + (b & 1) * 8 + c
                                          : ws.RC + (b - 8) * RC_BLK + c;
                col_item(mode, colp, k < 2 ? 16 : 8, inf, v);
            }
            __syncwarp();                                         // this round's RY / RC reads are done: S / CT may overwrite
            int fr, fc;            // final-grid position of this lane's block
            if (k < 2) { fr = 2 * th + ((b >> 1) & 1); fc = 2 * (2 * tp + k) + (b & 1); }
            else { fr = th; fc = 2 * tp + ((b - 8) >> 1); }
            T = run_ops(v, pl, comp, c, zero, tb, stats, img, fr, fc, GRID_Y);

            if (OUT_MODE == RGBNM_K0_OUT_INT16_PLANES) {
                int16_t* o16 = reinterpret_cast<int16_t*>(out_) + size_t(img) * PLANE_ELEMS;
                int blk;
                if (k < 2) blk = (2 * th + ((b >> 1) & 1)) * GRID_Y + 2 * (2 * tp + k) + (b & 1);
                else blk = GRID_Y * GRID_Y + ((comp - 1) * GRID_C + th) * GRID_C + 2 * tp + ((b - 8) >> 1);
                (void)TOKENS; (void)FEAT;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int li = T ? c : i, lj = T ? i : c;
                    o16[size_t(blk) * 64 + li * 8 + lj] = int16_t(int(v[i]));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = to_range(v[i]);
                float* tile;
                int ld;
                if (k < 2) { tile = ws.RY + k * TT_TOK + (((b >> 1) & 1) * 8) * TT_LD + (b & 1) * 8; ld = TT_LD; }
                else { tile = ws.RC + (b - 8) * CT_BLK; ld = CT_LD; }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int li = T ? c : i, lj = T ? i : c;
                    tile[li * ld + lj] = v[i];
                }
            }
        }
        __syncwarp();
        if (OUT_MODE == RGBNM_K0_OUT_INT16_PLANES) continue;

        if (SWIN) {
            swin_out<OUT_MODE>(ws, lane, img, th, tp, out_);
            __syncwarp();
            continue;
        }
        // ---- P2b: column half of the sub-block conversion, in place: S -> A16 . S ----------------
        const int tok = lane >> 4;
        float* TT = ws.RY + tok * TT_TOK;
        {
            const int col = lane & 15;
            float xl[8], xr[8], t[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) { xl[i] = TT[i * TT_LD + col]; xr[i] = TT[(8 + i) * TT_LD + col]; }
            a16_1d(xl, xr, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) TT[i * TT_LD + col] = t[i];
        }
        __syncwarp();

        // ---- P3: row half of the sub-block conversion (plainvit.py:50-69) + stores ---------------
        {
            const int row = lane & 15;                            // lane = (tok, row)
            const float4* src = reinterpret_cast<const float4*>(TT + row * TT_LD);
            const float4 a = src[0], b = src[1], c = src[2], d = src[3];
            const float xl[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            const float xr[8] = {c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
            float o[16];
            a16_1d(xl, xr, o);                                    // (A16 . S) . A16^T  (along the row)
            const int token = th * 14 + 2 * tp + tok;
            store_row16<OUT_MODE>(out_, (size_t(img) * TOKENS + token) * FEAT + row * 16, o);
        }
        {
            const int cb = lane >> 3, row = lane & 7;             // lane = (chroma block tok*2 + comp-1, row)
            const float4* src = reinterpret_cast<const float4*>(ws.RC + cb * CT_BLK + row * CT_LD);
            const int token = th * 14 + 2 * tp + (cb >> 1);
            store_row8<OUT_MODE>(out_, (size_t(img) * TOKENS + token) * FEAT + 256 + (cb & 1) * 64 + row * 8, src[0], src[1]);
        }
        __syncwarp();
    }
}

template <int OUT_MODE, int LAYOUT>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, LAYOUT == RGBNM_K0_LAYOUT_SWIN4 ? 18 : 25)
k0_fused_kernel(const int16_t* __restrict__ y, const int16_t* __restrict__ cbcr, const int16_t* __restrict__ quant,
                const rgbnm_plan* __restrict__ plans, rgbnm_k0_tables tb, const float* __restrict__ stats_all,
                void* __restrict__ out_, int n_images, int hb, int wb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    using WarpSmem = WarpSmemT<LAYOUT>;
    constexpr int ROWS_PER_IMAGE = Geo<LAYOUT>::GRID_C;
    WarpSmem& ws = reinterpret_cast<WarpSmem*>(smem_raw)[warp];
    const int task = blockIdx.x * WARPS_PER_CTA + warp;
    if (task >= n_images * ROWS_PER_IMAGE) return;
    const int img = task / ROWS_PER_IMAGE, th = task - img * ROWS_PER_IMAGE;

    // ---- per-image state ------------------------------------------------------------------------
    {
        const int* psrc = reinterpret_cast<const int*>(plans + img);
        int* pdst = reinterpret_cast<int*>(&ws.plan);
        if (lane < int(sizeof(rgbnm_plan) / 4)) pdst[lane] = __ldg(psrc + lane);
        for (int k = lane; k < 192; k += 32) {
            const float q = float(__ldg(quant + size_t(img) * 192 + k));
            ws.qf[k] = q;
            ws.cq[k] = -DEQ_BIAS * q;
        }
    }
    __syncwarp();
    const int hc = hb >> 1, wc = wb >> 1;
    const int16_t* y_img = y + size_t(img) * hb * wb * 64;
    const int16_t* c_img = cbcr + size_t(img) * 2 * hc * wc * 64;
    const float* stats = stats_all + size_t(img) * RGBNM_MAX_OPS * 2;
    if (mode_of(ws.plan.crop_size, Geo<LAYOUT>::GRID_Y) == MODE_BAD) return;
    process_row<OUT_MODE, LAYOUT>(ws, lane, img, th, y_img, c_img, tb, stats, out_, wb, hc, wc);
}

}  // namespace k0

template <int OUT_MODE, int LAYOUT>
static int launch_k0(const int16_t* y, const int16_t* cbcr, const int16_t* quant, const rgbnm_plan* plans,
                     const rgbnm_k0_tables* tables, const float* stats, void* out, int n, int hb, int wb, cudaStream_t st) {
    using namespace k0;
    static bool configured = false;
    const size_t smem = sizeof(WarpSmemT<LAYOUT>) * WARPS_PER_CTA;
    if (!configured) {
        RGBNM_CUDA_CHECK(cudaFuncSetAttribute(k0_fused_kernel<OUT_MODE, LAYOUT>,
                                              cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured = true;
    }
    const int tasks = n * Geo<LAYOUT>::GRID_C;
    const int grid = (tasks + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    k0_fused_kernel<OUT_MODE, LAYOUT><<<grid, dim3(WARPS_PER_CTA * 32), smem, st>>>(y, cbcr, quant, plans, *tables, stats, out, n, hb, wb);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

// second-generation ViT-layout kernel (k0_vit2.cu): packed fp32x2 arithmetic, quads of tokens
int rgbnm_k0_vit2_launch(const int16_t* y, const int16_t* cbcr, const int16_t* quant, const rgbnm_plan* plans, const rgbnm_k0_tables* tables,
                         const float* stats, void* out, int out_mode, int nosub, int n, int hb, int wb, cudaStream_t st);

extern "C" int rgbnm_k0_fused_ex(const int16_t* y, const int16_t* cbcr, const int16_t* quant, const rgbnm_plan* plans,
                                 const rgbnm_k0_tables* tables, const float* stats, void* out, int out_mode, int layout,
                                 int n, int hb, int wb, void* stream) {
    if (!y || !cbcr || !quant || !plans || !tables || !stats || !out || n < 0) return RGBNM_ERR_ARG;
    if (hb < 2 || wb < 2 || hb > 255 || wb > 255 || (hb & 1) || (wb & 1)) return RGBNM_ERR_ARG;
    if (n == 0) return RGBNM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (layout == RGBNM_K0_LAYOUT_VIT16_NOSUB) return rgbnm_k0_vit2_launch(y, cbcr, quant, plans, tables, stats, out, out_mode, 1, n, hb, wb, st);
    if (layout == RGBNM_K0_LAYOUT_VIT16) {
        static const bool v1 = getenv("RGBNM_K0_V1") != nullptr;      // first-generation kernel, A/B measurements only
        if (!v1) return rgbnm_k0_vit2_launch(y, cbcr, quant, plans, tables, stats, out, out_mode, 0, n, hb, wb, st);
    }
#define RGBNM_K0_CASE(OM, LY) \
    if (out_mode == OM && layout == LY) return launch_k0<OM, LY>(y, cbcr, quant, plans, tables, stats, out, n, hb, wb, st)
    RGBNM_K0_CASE(RGBNM_K0_OUT_F32, RGBNM_K0_LAYOUT_VIT16);
    RGBNM_K0_CASE(RGBNM_K0_OUT_BF16, RGBNM_K0_LAYOUT_VIT16);
    RGBNM_K0_CASE(RGBNM_K0_OUT_INT16_PLANES, RGBNM_K0_LAYOUT_VIT16);
    RGBNM_K0_CASE(RGBNM_K0_OUT_F32, RGBNM_K0_LAYOUT_SWIN4);
    RGBNM_K0_CASE(RGBNM_K0_OUT_BF16, RGBNM_K0_LAYOUT_SWIN4);
    RGBNM_K0_CASE(RGBNM_K0_OUT_INT16_PLANES, RGBNM_K0_LAYOUT_SWIN4);
#undef RGBNM_K0_CASE
    return RGBNM_ERR_ARG;
}

extern "C" int rgbnm_k0_fused(const int16_t* y, const int16_t* cbcr, const int16_t* quant, const rgbnm_plan* plans,
                              const rgbnm_k0_tables* tables, const float* stats, void* out, int out_mode, int n,
                              int hb, int wb, void* stream) {
    return rgbnm_k0_fused_ex(y, cbcr, quant, plans, tables, stats, out, out_mode, RGBNM_K0_LAYOUT_VIT16, n, hb, wb, stream);
}

extern "C" int rgbnm_k0_launch_count(void) { return 2; }
