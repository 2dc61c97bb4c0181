// Device helpers shared by the fused DCT kernel (k0_fused.cu) and its DC-statistics
// pre-pass (k0_dcstats.cu).  Both translation units must evaluate the *same* IEEE
// operation sequences for a resized block's DC term, so the 1-D transforms live here and
// the library is compiled with --fmad=false (FMA only where written as fmaf()).
//
// Math (reference utils/dct_ops.py:150-208, 436-527):
//   A16 = D16 . blockdiag(D8, D8)^T is the orthonormal 16x16 "conversion matrix".
//   Structure exploited here (verified in tests/test_k0_math.py):
//     A16[k][8+j] = (-1)^(k+j) A16[k][j]                     (mirror symmetry)
//     A16[2m][j]  = delta(j, m) / sqrt(2)                    (even rows are 2-sparse)
//     A16[2m+1][j] = kB[m][j], a dense 8x8 block             (odd rows)
//   so a 16-point product costs 8 mul + 64 fma + 16 add instead of 256 fma, and the
//   8-output (downsample) form 4 mul + 32 fma + 12 add instead of 128.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rgbnm_b200.h"

namespace k0 {

constexpr float KE = 0.707106781f;
constexpr float CLAMP_LO = -1024.0f;
constexpr float CLAMP_HI = 1016.0f;
// Output geometry per layout (include/rgbnm_b200.h RGBNM_K0_LAYOUT_*):
//   VIT16: 28 x 28 luma blocks -> 196 tokens x [Y 16x16 | Cb 8x8 | Cr 8x8]   (models/plainvit.py:200-216, patch 16)
//   SWIN4: 32 x 32 luma blocks -> 4096 tokens x [Y 4x4 | Cb 2x2 | Cr 2x2]    (models/swinv2.py:505-576, patch 4)
template <int LAYOUT>
struct Geo {
    static constexpr int GRID_Y = LAYOUT == RGBNM_K0_LAYOUT_SWIN4 ? 32 : 28;   // luma blocks per side after resize
    static constexpr int GRID_C = GRID_Y / 2;
    static constexpr int TOKENS = LAYOUT == RGBNM_K0_LAYOUT_SWIN4 ? 4096 : 196;
    static constexpr int FEAT = LAYOUT == RGBNM_K0_LAYOUT_SWIN4 ? 24 : 384;
    static constexpr int PLANE_ELEMS = (GRID_Y * GRID_Y + 2 * GRID_C * GRID_C) * 64;
};

enum Mode { MODE_DOWN2 = 0, MODE_IDENT = 1, MODE_UP2 = 2, MODE_BAD = 3 };

// crop side -> resize case for a G x G output grid
__device__ __forceinline__ int mode_of(int crop_size, int G) {
    return crop_size == 2 * G ? MODE_DOWN2 : crop_size == G ? MODE_IDENT : 2 * crop_size == G ? MODE_UP2 : MODE_BAD;
}

#define K0_KB_TABLE                                                                                               \
    {                                                                                                             \
        {0.637643577f, 0.298637845f, -0.0584927049f, 0.0240878632f, -0.0124921762f, 0.00706026221f,               \
         -0.00392845722f, 0.00177477869f},                                                                        \
        {-0.215305887f, 0.544633646f, 0.381218413f, -0.0950727443f, 0.0436403016f, -0.0234808956f,                \
         0.0127634981f, -0.00570323591f},                                                                         \
        {0.132584711f, -0.221907938f, 0.508053606f, 0.400770851f, -0.106061464f, 0.0493435375f, -0.0252556743f,   \
         0.0109887194f},                                                                                          \
        {-0.0985193279f, 0.150923057f, -0.202355499f, 0.497064887f, 0.406474087f, -0.107836242f, 0.0475687588f,   \
         -0.0195524384f},                                                                                         \
        {0.0808527229f, -0.119774931f, 0.139934337f, -0.196652263f, 0.495290108f, 0.404699308f, -0.102133006f,    \
         0.0365800394f},                                                                                          \
        {-0.0708680044f, 0.103354298f, -0.114071695f, 0.138159559f, -0.198427042f, 0.500993344f, 0.393710589f,    \
         -0.0825805681f},                                                                                         \
        {0.0653123269f, -0.094519257f, 0.101579519f, -0.115846474f, 0.143862794f, -0.209415762f, 0.520545782f,    \
         0.357130549f},                                                                                           \
        {-0.0628024108f, 0.0905907998f, -0.0962940357f, 0.107282755f, -0.126835194f, 0.163415233f,                \
         -0.245995801f, 0.60312635f},                                                                             \
    }

// o[k] = SCALE * sum_j (A16[k][j] xl[j] + A16[k][8+j] xr[j]),  k = 0..7  (downsample-by-2 half)
template <int SCALE_NUM, int SCALE_DEN>
__device__ __forceinline__ void down2_1d(const float (&xl)[8], const float (&xr)[8], float (&o)[8]) {
    constexpr float B[8][8] = K0_KB_TABLE;
    constexpr float s = float(SCALE_NUM) / float(SCALE_DEN);
    o[0] = (KE * s) * (xl[0] + xr[0]);
    o[2] = (KE * s) * (xl[1] - xr[1]);
    o[4] = (KE * s) * (xl[2] + xr[2]);
    o[6] = (KE * s) * (xl[3] - xr[3]);
    float u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = (j & 1) ? (xl[j] + xr[j]) : (xl[j] - xr[j]);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        float acc = (B[m][0] * s) * u[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) acc = fmaf(B[m][j] * s, u[j], acc);
        o[2 * m + 1] = acc;
    }
}

// o[k] = sum_j (A16[k][j] xl[j] + A16[k][8+j] xr[j]),  k = 0..15  (sub-block conversion)
__device__ __forceinline__ void a16_1d(const float (&xl)[8], const float (&xr)[8], float (&o)[16]) {
    constexpr float B[8][8] = K0_KB_TABLE;
    float u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float sm = xl[j] + xr[j], df = xl[j] - xr[j];
        o[2 * j] = KE * ((j & 1) ? df : sm);
        u[j] = (j & 1) ? sm : df;
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        float acc = B[m][0] * u[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) acc = fmaf(B[m][j], u[j], acc);
        o[2 * m + 1] = acc;
    }
}

// o[a] = SCALE * sum_b x[b] A16[b][8*child + a],  a = 0..7  (upsample-by-2, one child half)
template <int SCALE>
__device__ __forceinline__ void up2_1d(const float (&x)[8], int child, float (&o)[8]) {
    constexpr float B[8][8] = K0_KB_TABLE;
    constexpr float s = float(SCALE);
    const float cs = child ? -1.0f : 1.0f;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        float w = (B[0][a] * s) * x[1];
        w = fmaf(B[1][a] * s, x[3], w);
        w = fmaf(B[2][a] * s, x[5], w);
        w = fmaf(B[3][a] * s, x[7], w);
        const float v = (a < 4) ? (KE * s) * x[2 * a] : 0.0f;
        float r = fmaf(cs, w, v);
        if (a & 1) r = child ? -r : r;
        o[a] = r;
    }
}

// ---------------------------------------------------------------------------------------
// Block DECOMPOSITION of the SwinV2 patch embedding (models/swinv2.py:505-576, plainvit.py:50-69 with
// combine=False): D = A^T . X . A with A = A(4,2) for luma (8x8 -> 2x2 sub-blocks of 4x4) and A = A(2,4)
// for chroma (8x8 -> 4x4 sub-blocks of 2x2).  One call = one 1-D pass: o[j] = sum_k A[k][j] x[k].
// A(4,2) has the structure of A16 at half the size: A[k][4+j] = (-1)^(k+j) A[k][j], even rows
// A[2m][j] = delta(j, m) / sqrt(2), odd rows A[2m+1][j] = kB4[m][j] -> 4 mul + 16 fma + 8 add.
// ---------------------------------------------------------------------------------------
#define K0_KB4_TABLE                                                                  \
    {                                                                                 \
        {0.64072883f, 0.2939689f, -0.052791018f, 0.016183784f},                       \
        {-0.22499399f, 0.5593675f, 0.36294368f, -0.068974815f},                       \
        {0.15033615f, -0.24921477f, 0.5431836f, 0.3467599f},                          \
        {-0.1274489f, 0.19642372f, -0.26539865f, 0.61215854f},                        \
    }
__device__ __forceinline__ void decomp_y_1d(const float (&x)[8], float (&o)[8]) {
    constexpr float B[4][4] = K0_KB4_TABLE;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float w = B[0][j] * x[1];
        w = fmaf(B[1][j], x[3], w);
        w = fmaf(B[2][j], x[5], w);
        w = fmaf(B[3][j], x[7], w);
        const float e = KE * x[2 * j];
        o[j] = e + w;
        o[4 + j] = (j & 1) ? (w - e) : (e - w);
    }
}
// A(2,4) = D8 . blockdiag(D2 x 4)^T, dense (values of dct_ops.generate_conversion_matrix(2, 4), fp32)
#define K0_A24_TABLE                                                                                                       \
    {                                                                                                                      \
        {0.5f, 0.0f, 0.5f, 0.0f, 0.5f, 0.0f, 0.5f, 0.0f},                                                                  \
        {0.6407289f, 0.052791085f, 0.26539853f, 0.12744892f, -0.26539862f, 0.12744893f, -0.6407289f, 0.05279105f},         \
        {0.46193975f, 0.1913417f, -0.4619398f, 0.19134171f, -0.46193963f, -0.19134182f, 0.46193984f, -0.19134165f},        \
        {0.22499405f, 0.36294374f, -0.5431836f, -0.15033633f, 0.54318374f, -0.15033615f, -0.2249942f, 0.3629437f},         \
        {0.0f, 0.5f, 0.0f, -0.5f, 0.0f, 0.5f, 0.0f, -0.5f},                                                                \
        {-0.15033624f, 0.5431837f, 0.36294368f, -0.22499393f, -0.36294374f, -0.22499414f, 0.15033603f, 0.5431839f},        \
        {-0.1913417f, 0.46193975f, 0.19134156f, 0.46193993f, 0.1913418f, -0.4619395f, -0.19134162f, -0.46194f},            \
        {-0.12744884f, 0.26539847f, -0.052791126f, 0.6407287f, 0.052791115f, 0.6407289f, 0.12744878f, 0.26539934f},        \
    }
__device__ __forceinline__ void decomp_c_1d(const float (&x)[8], float (&o)[8]) {
    constexpr float A[8][8] = K0_A24_TABLE;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        // rows 0 and 4 of A(2,4) are 4-sparse (+-1/2): the zero products are dropped at compile time
        float acc = 0.0f;
        bool first = true;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (A[k][j] == 0.0f) continue;
            acc = first ? A[k][j] * x[k] : fmaf(A[k][j], x[k], acc);
            first = false;
        }
        o[j] = acc;
    }
}

__device__ __forceinline__ float rint_magic(float x) {
    // round-half-even for |x| < 2^22 (torch.round semantics), two full-rate FADDs
    return (x + 12582912.0f) - 12582912.0f;
}
__device__ __forceinline__ float clampf(float x) { return fminf(fmaxf(x, CLAMP_LO), CLAMP_HI); }

// dequantise 8 int16 packed in an int4: x[j] = float(v[j]) * q[j], exactly, without I2F.
// The biased 16-bit pattern (v ^ 0x8000) is dropped into the mantissa of 2^23 with one PRMT,
// giving m = 2^23 + 32768 + v; then x = fma(m, q, cq) with cq = -(2^23 + 32768) * q, which is
// exact: cq = -2^15 * 257 * q is representable for q <= 255 and |v * q| < 2^24.
// q8 / cq8: 8 consecutive fp32 table entries (16-byte aligned, shared memory).
__device__ __forceinline__ void dequant8_regs(const int4& raw, const float (&q)[8], const float (&cq)[8], bool clamp,
                                              float (&x)[8]) {
    const unsigned w[4] = {unsigned(raw.x) ^ 0x80008000u, unsigned(raw.y) ^ 0x80008000u,
                           unsigned(raw.z) ^ 0x80008000u, unsigned(raw.w) ^ 0x80008000u};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float lo = __uint_as_float(__byte_perm(w[p], 0x4B000000u, 0x7610));
        const float hi = __uint_as_float(__byte_perm(w[p], 0x4B000000u, 0x7632));
        x[2 * p] = fmaf(lo, q[2 * p], cq[2 * p]);
        x[2 * p + 1] = fmaf(hi, q[2 * p + 1], cq[2 * p + 1]);
    }
    if (clamp) {     // datasets.py:288-290; the decoder proves it idle for ordinary JPEGs (plan.clamp_in = 0)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = clampf(x[j]);
    }
}
#define K0_LOAD_Q(q8, cq8)                                                                                          \
    const float4 qa = *reinterpret_cast<const float4*>(q8), qb = *reinterpret_cast<const float4*>((q8) + 4);        \
    const float4 ca = *reinterpret_cast<const float4*>(cq8), cb = *reinterpret_cast<const float4*>((cq8) + 4);      \
    const float q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};                                            \
    const float cq[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w}
__device__ __forceinline__ void dequant8(const int4& raw, const float* __restrict__ q8, const float* __restrict__ cq8,
                                         bool clamp, float (&x)[8]) {
    K0_LOAD_Q(q8, cq8);
    dequant8_regs(raw, q, cq, clamp, x);
}
// two horizontally adjacent blocks share the table row
__device__ __forceinline__ void dequant8x2(const int4& ra, const int4& rb, const float* __restrict__ q8,
                                           const float* __restrict__ cq8, bool clamp, float (&xl)[8], float (&xr)[8]) {
    K0_LOAD_Q(q8, cq8);
    dequant8_regs(ra, q, cq, clamp, xl);
    dequant8_regs(rb, q, cq, clamp, xr);
}
constexpr float DEQ_BIAS = 8421376.0f;   // 2^23 + 2^15

// ---------------------------------------------------------------------------------------
// Geometry: where does the block at final-grid position (r, c) of plane `comp` come from?
// Walks the op list backwards (RandAugment ops, then RandomFlip_DCT).  Returns the stage
// after which the block is identically zero (-1: never) and the post-resize position.
// ---------------------------------------------------------------------------------------
struct Trace {
    int r, c;     // position in the post-resize grid
    int zero;     // -1, or index of the op that zeroed the block
};

template <int GRID_Y>
__device__ __forceinline__ Trace trace_back(const rgbnm_plan& pl, int comp, int r, int c) {
    const int G = comp == 0 ? GRID_Y : GRID_Y / 2;
    Trace t{r, c, -1};
    for (int k = pl.n_ops - 1; k >= 0; --k) {
        const rgbnm_plan_op& op = pl.ops[k];
        const int code = op.code;
        if (code == RGBNM_OP_TRANSLATE_X) {
            const int nc = t.c - op.p[comp == 0 ? 0 : 1];
            if (nc < 0 || nc >= G) { t.zero = k; return t; }
            t.c = nc;
        } else if (code == RGBNM_OP_TRANSLATE_Y) {
            const int nr = t.r - op.p[comp == 0 ? 0 : 1];
            if (nr < 0 || nr >= G) { t.zero = k; return t; }
            t.r = nr;
        } else if (code == RGBNM_OP_ROT90) {
            const int rr = t.r, cc = t.c;
            if (op.p[0] > 0) { t.r = cc; t.c = G - 1 - rr; }   // torch.rot90(k=+1): out[i][j] = in[j][G-1-i]
            else             { t.r = G - 1 - cc; t.c = rr; }   // k=-1: out[i][j] = in[G-1-j][i]
        } else if (code == RGBNM_OP_CUTOUT) {
            const int o = comp == 0 ? 0 : 4;
            if (t.r >= op.p[o] && t.r < op.p[o + 1] && t.c >= op.p[o + 2] && t.c < op.p[o + 3]) { t.zero = k; return t; }
        } else if (code == RGBNM_OP_GRAYSCALE) {
            if (comp != 0) { t.zero = k; return t; }
        } else if (code == RGBNM_OP_CHROMADROP) {
            if (comp == 1 + op.p[0]) { t.zero = k; return t; }
        }
    }
    if (pl.flip) t.c = G - 1 - t.c;
    return t;
}

// Position of the block that ends at final-grid position (r, c) at the moment op `k` runs: the geometric ops after k
// walked backwards (zeroing ops do not move blocks).  Used by ops that look at another block (Solarize: chroma follows the
// luma block at twice its coordinates).
template <int GRID_Y>
__device__ __forceinline__ void position_at_op(const rgbnm_plan& pl, int comp, int k, int& r, int& c) {
    const int G = comp == 0 ? GRID_Y : GRID_Y / 2;
    for (int j = pl.n_ops - 1; j > k; --j) {
        const rgbnm_plan_op& op = pl.ops[j];
        const int code = op.code;
        if (code == RGBNM_OP_TRANSLATE_X) c -= op.p[comp == 0 ? 0 : 1];
        else if (code == RGBNM_OP_TRANSLATE_Y) r -= op.p[comp == 0 ? 0 : 1];
        else if (code == RGBNM_OP_ROT90) {
            const int rr = r, cc = c;
            if (op.p[0] > 0) { r = cc; c = G - 1 - rr; }
            else             { r = G - 1 - cc; c = rr; }
        }
    }
}

}  // namespace k0
