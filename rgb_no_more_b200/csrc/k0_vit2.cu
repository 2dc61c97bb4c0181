// K0, ViT layout, second generation: the fused DCT-domain data-path kernel on packed fp32x2 arithmetic (sm_100a).
//
// Same contract as k0_fused.cu (rows a2-a25 of SURVEY.md 8a; reference file:line cited there) and bit-identical results --
// every 1-D transform below evaluates the operation sequence of its scalar twin in k0_common.cuh, which the DC-statistics
// pre-pass (k0_dcstats.cu) shares -- but organised around what bounded the first kernel: instruction issue.
//
//   * Blackwell's FFMA2 / FADD2 / FMUL2 (PTX fma/add/sub/mul.rn.f32x2) do two IEEE fp32 operations per issue slot.  A lane
//     therefore always works on the SAME line (coefficient row or column) of TWO horizontally adjacent tokens at once: the
//     pair (token 0, token 1) is the .x / .y of every float2 from the global loads to the global stores; basis constants
//     are broadcast immediates, quantisation-table entries broadcast registers.  No lane ever needs a value of the other
//     half, so nothing is shuffled or transposed in registers.
//   * Work unit = QUAD of tokens (2 token rows x 2 token columns = 2 packed pairs): 16 luma + 8 chroma post-resize blocks.
//     Every pass then has exactly 32 independent lane items per round (no idle lanes, no divergent luma / chroma split):
//       R  row pass   : lane = one source coefficient row of one block pair: 16-byte loads, I2F-free dequantisation, 1-D resize
//                       along the row -> tile T (fp32 pairs, shared memory)
//       C  col pass   : lane = one tile column: 1-D resize down the column, round, flip / RandAugment ops, ToRange -> tile S
//       P2b / P3      : column / row half of the 16 x 16 sub-block conversion A16 . S . A16^T, bf16 pack, 16-byte stores
//     Chroma re-uses the luma tile space after the luma token rows have been stored.
//   * Tiles are 16 pairs wide with a row pitch of 144 bytes: 16-byte row-pass stores of 8 consecutive rows, 8-byte column reads of
//     16 consecutive columns and 16-byte row reads of 8 consecutive rows are all bank-conflict free with plain linear addressing
//     (compile-time offsets, no swizzle arithmetic).  Pairs travel as 64-bit registers (mov.b64 views, ld/st.shared.b64 on
//     32-bit shared-window addresses): no pack / unpack moves, no generic-address conversion.
//   * Quads are dealt to warps DYNAMICALLY (a device ticket counter; 49 quads per image, 20 warps per SM resident): the work of a
//     quad depends on its image's plan (x2-down reads 4x the coefficients of x2-up, RandAugment ops add column-pass work) and a
//     batch of 256 images holds only ~4 quads per warp, so fixed ranges left warps idle at the end of the training mix.
//
// HBM-bound by design; CUDA cores only (north_star reserves the tensor cores for the ViT contractions).
#include <atomic>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstddef>

#include "../../include/rgbnm_b200.h"
#include "common.cuh"
#include "k0_common.cuh"

namespace k0v2 {
using namespace k0;

typedef unsigned long long p2;            // packed pair: token 0 in the low word, token 1 in the high word
// 20 resident warps / SM (4 CTAs x 5 warps, <= 96 registers: 92 used, no spills).  Measured on B200 (profiles/r02_k0_experiments.md):
// 16, 18 and 20 warps / SM land within 4 % of each other; 20 is the best in both the eval geometry and the training mix.
#ifndef K0V2_WARPS
#define K0V2_WARPS 5
#endif
#ifndef K0V2_CTAS
#define K0V2_CTAS 4
#endif
constexpr int WARPS = K0V2_WARPS;         // warps per CTA (independent: no block-level barrier)
constexpr int CTAS_PER_SM = K0V2_CTAS;    // shared memory: CTAS x (WARPS x 11,440 + 1 KB) <= 228 KB
constexpr int PITCH = 144;                // tile row pitch in bytes (16 pairs + 2 pad)
constexpr int TILE_B = 32 * PITCH;        // bytes per tile (32 rows: the x2-down row pass produces 16 rows per block row)
constexpr int QROW_B = 80;                // fp32 table row: [q0..q7 | cq0..cq7 | 16 B pad] -> the 8 rows of a table sit on distinct bank groups
constexpr int QUADS_PER_IMAGE = 49;       // 7 token-row pairs x 7 token-column pairs
constexpr int TOKENS = 196, FEAT = 384, GRID_Y = 28;
constexpr int PLANE_ELEMS = (28 * 28 + 2 * 14 * 14) * 64;

struct __align__(16) WarpSmem {
    unsigned char T[2 * TILE_B];      // tile of pair p at T + p * TILE_B: rows of 16 p2 (+ pad), row-major
    unsigned char qt[24 * QROW_B];    // per (component, coefficient row): q[8] and cq[8] = -(2^23 + 2^15) * q (dequantisation bias)
    rgbnm_plan plan;
    int info[24];                     // per block of the quad: source block index, child flags, zeroing op (pack_info)
    int pad_[8];
};
static_assert(sizeof(WarpSmem) % 16 == 0, "warp slices must stay 16-byte aligned");
static_assert(CTAS_PER_SM * (WARPS * sizeof(WarpSmem) + 1024) <= 228 * 1024, "the resident CTAs of an SM must fit");

// ---- packed fp32x2 primitives: both halves IEEE round-to-nearest, like the scalar operators they replace ------------------
__device__ __forceinline__ p2 pk(float lo, float hi) { p2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo_of(p2 v) { return __uint_as_float(unsigned(v)); }
__device__ __forceinline__ float hi_of(p2 v) { return __uint_as_float(unsigned(v >> 32)); }
__device__ __forceinline__ p2 ffma2(p2 a, p2 b, p2 c) { p2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ p2 fadd2(p2 a, p2 b) { p2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ p2 fsub2(p2 a, p2 b) { p2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ p2 fmul2(p2 a, p2 b) { p2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ p2 bq(float s) { return pk(s, s); }      // broadcast operand (immediate or .F32 register form in SASS)
// rint_magic.  The first addition is written x * 1 + M: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (single
// rounding) even with --fmad=false, which would break bit-parity with the scalar twins wherever x is a product.
__device__ __forceinline__ p2 rint2(p2 x) { return fsub2(ffma2(x, bq(1.0f), bq(12582912.0f)), bq(12582912.0f)); }
__device__ __forceinline__ p2 clamp2(p2 x) { return pk(clampf(lo_of(x)), clampf(hi_of(x))); }
__device__ __forceinline__ p2 clamp_hi2(p2 x) { return pk(fminf(lo_of(x), CLAMP_HI), fminf(hi_of(x), CLAMP_HI)); }

// ---- shared memory through 32-bit shared-window addresses (no generic-address conversion, 64-bit register pairs moved as such) -----
__device__ __forceinline__ p2 lds64(uint32_t a) { p2 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts64(uint32_t a, p2 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void lds128(uint32_t a, p2& v0, p2& v1) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, p2 v0, p2 v1) {
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(a), "l"(v0), "l"(v1) : "memory");
}
__device__ __forceinline__ int lds32(uint32_t a) { int v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void lds_f4(uint32_t a, float (&o)[8], int at) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o[at]), "=f"(o[at + 1]), "=f"(o[at + 2]), "=f"(o[at + 3]) : "r"(a) : "memory");
}

// ---- 1-D transforms, packed twins of k0_common.cuh (same operation order, same compile-time constants) -------------------
template <int SCALE_NUM, int SCALE_DEN>
__device__ __forceinline__ void down2_1d_p(const p2 (&xl)[8], const p2 (&xr)[8], p2 (&o)[8]) {
    constexpr float B[8][8] = K0_KB_TABLE;
    constexpr float s = float(SCALE_NUM) / float(SCALE_DEN);
    o[0] = fmul2(fadd2(xl[0], xr[0]), bq(KE * s));
    o[2] = fmul2(fsub2(xl[1], xr[1]), bq(KE * s));
    o[4] = fmul2(fadd2(xl[2], xr[2]), bq(KE * s));
    o[6] = fmul2(fsub2(xl[3], xr[3]), bq(KE * s));
    p2 u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = (j & 1) ? fadd2(xl[j], xr[j]) : fsub2(xl[j], xr[j]);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        p2 acc = fmul2(u[0], bq(B[m][0] * s));
#pragma unroll
        for (int j = 1; j < 8; ++j) acc = ffma2(bq(B[m][j] * s), u[j], acc);
        o[2 * m + 1] = acc;
    }
}

__device__ __forceinline__ void a16_1d_p(const p2 (&xl)[8], const p2 (&xr)[8], p2 (&o)[16]) {
    constexpr float B[8][8] = K0_KB_TABLE;
    p2 u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const p2 sm = fadd2(xl[j], xr[j]), df = fsub2(xl[j], xr[j]);
        o[2 * j] = fmul2((j & 1) ? df : sm, bq(KE));
        u[j] = (j & 1) ? sm : df;
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        p2 acc = fmul2(u[0], bq(B[m][0]));
#pragma unroll
        for (int j = 1; j < 8; ++j) acc = ffma2(bq(B[m][j]), u[j], acc);
        o[2 * m + 1] = acc;
    }
}

// child0 / child1: which half of the source block the block of token 0 / 1 is (luma tokens sit two blocks apart, so their
// parities agree; adjacent chroma blocks are the two children of one source block)
template <int SCALE>
__device__ __forceinline__ void up2_1d_p(const p2 (&x)[8], int child0, int child1, p2 (&o)[8]) {
    constexpr float B[8][8] = K0_KB_TABLE;
    constexpr float s = float(SCALE);
    const p2 cs = pk(child0 ? -1.0f : 1.0f, child1 ? -1.0f : 1.0f);
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        p2 w = fmul2(x[1], bq(B[0][a] * s));
        w = ffma2(bq(B[1][a] * s), x[3], w);
        w = ffma2(bq(B[2][a] * s), x[5], w);
        w = ffma2(bq(B[3][a] * s), x[7], w);
        const p2 v = (a < 4) ? fmul2(x[2 * a], bq(KE * s)) : bq(0.0f);
        p2 r = ffma2(cs, w, v);
        if (a & 1) r = fmul2(r, cs);              // child ? -r : r  (a product with +-1 is exact)
        o[a] = r;
    }
}

// dequantise row `raw0` (token 0) and `raw1` (token 1): x[j] = (float(v0[j]) * q[j], float(v1[j]) * q[j]) exactly (k0_common.cuh dequant8)
__device__ __forceinline__ void dequant8_p(const int4& raw0, const int4& raw1, const float (&q)[8], const float (&cq)[8], bool clamp,
                                           p2 (&x)[8]) {
    const unsigned a[4] = {unsigned(raw0.x) ^ 0x80008000u, unsigned(raw0.y) ^ 0x80008000u, unsigned(raw0.z) ^ 0x80008000u,
                           unsigned(raw0.w) ^ 0x80008000u};
    const unsigned b[4] = {unsigned(raw1.x) ^ 0x80008000u, unsigned(raw1.y) ^ 0x80008000u, unsigned(raw1.z) ^ 0x80008000u,
                           unsigned(raw1.w) ^ 0x80008000u};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const p2 lo = pk(__uint_as_float(__byte_perm(a[p], 0x4B000000u, 0x7610)), __uint_as_float(__byte_perm(b[p], 0x4B000000u, 0x7610)));
        const p2 hi = pk(__uint_as_float(__byte_perm(a[p], 0x4B000000u, 0x7632)), __uint_as_float(__byte_perm(b[p], 0x4B000000u, 0x7632)));
        x[2 * p] = ffma2(lo, bq(q[2 * p]), bq(cq[2 * p]));
        x[2 * p + 1] = ffma2(hi, bq(q[2 * p + 1]), bq(cq[2 * p + 1]));
    }
    if (clamp) {     // datasets.py:288-290; idle for ordinary JPEGs (plan.clamp_in = 0)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = clamp2(x[j]);
    }
}

// source block index (row-major in its plane) | child flags | zeroing op + 1
__device__ __forceinline__ int pack_info(int blk, int child_r, int child_c, int zero) {
    return (blk & 0xffff) | (child_r << 16) | (child_c << 17) | ((zero + 1) << 18);
}

__device__ __forceinline__ int4 ldg128(const unsigned char* base, uint32_t off) {
    return __ldg(reinterpret_cast<const int4*>(base + off));
}

// per-lane constants of the passes (bits of the lane id -> shared-window addresses), set up once per warp
struct LaneK {
    uint32_t T;          // tile 0
    // tile columns are stored interleaved: column c (0..7) of half h (block column / chroma component) sits in slot 2c + h
    uint32_t rst;        // row-pass store, luma (all modes) / chroma x2-down: T + (bit3 * 8 + i8) * PITCH + bit4 * 8, columns at + 16 k
    uint32_t rst_c;      // row-pass store, chroma identity / x2-up: T + bit4 * TILE_B + i8 * PITCH + bit3 * 8 (also the chroma read-out base)
    uint32_t ccol;       // column-pass / P2b column: T + bit4 * TILE_B + slot(lane & 15) * 8
    uint32_t prow;       // P3 row: T + bit4 * TILE_B + (lane & 15) * PITCH
    uint32_t qrow;       // table row of the lane's coefficient row: qt + i8 * QROW_B
    uint32_t info;       // info[0]
};

__device__ __forceinline__ void load_q(uint32_t qrow, float (&q)[8], float (&cq)[8]) {
    lds_f4(qrow, q, 0);
    lds_f4(qrow + 16, q, 4);
    lds_f4(qrow + 32, cq, 0);
    lds_f4(qrow + 48, cq, 4);
}

// ---- R: row pass --------------------------------------------------------------------------------------------------------
// One x2-down round = 32 lane items: row i8 of the upper / lower (lane bit 3) source block pair of block column `half` (lane bit 4).
//   plane / wbytes: component plane and the byte size of one of its block rows; b0: block id of token 0 for half 0 (token 1 at
//   + tok_stride).
struct RowLoads {
    int4 l0, r0, l1, r1;
};
__device__ __forceinline__ RowLoads r_down2_load(uint32_t info, int lane, const unsigned char* __restrict__ plane, int wbytes, int b0,
                                                 int tok_stride) {
    const uint32_t lc = ((lane >> 3) & 1) * wbytes + (lane & 7) * 16;
    const uint32_t ia = info + 4 * (b0 + (lane >> 4));
    const uint32_t o0 = (uint32_t(lds32(ia)) & 0xffffu) * 128u + lc, o1 = (uint32_t(lds32(ia + 4 * tok_stride)) & 0xffffu) * 128u + lc;
    const int4* p0 = reinterpret_cast<const int4*>(plane + o0);        // right-hand block of the pair = + 128 bytes: immediate offset
    const int4* p1 = reinterpret_cast<const int4*>(plane + o1);
    RowLoads L;
    L.l0 = __ldg(p0);
    L.r0 = __ldg(p0 + 8);
    L.l1 = __ldg(p1);
    L.r1 = __ldg(p1 + 8);
    return L;
}
__device__ __forceinline__ void r_down2_compute(const RowLoads& L, uint32_t qrow, uint32_t dst, bool clamp) {
    float q[8], cq[8];
    load_q(qrow, q, cq);
    p2 xl[8], xr[8], o[8];
    dequant8_p(L.l0, L.l1, q, cq, clamp, xl);
    dequant8_p(L.r0, L.r1, q, cq, clamp, xr);
    down2_1d_p<1, 1>(xl, xr, o);
    // 8-byte stores into every other slot (tile columns are interleaved: slot 2k + half): a lane never writes both halves of a
    // 16-byte chunk, so ptxas cannot fuse stores into STS.128 -- which needs four consecutive registers and costs two moves per
    // pair -- and the wavefront count is the same (rows r and r + 8 of a half-warp share banks: 2 x 8 wavefronts either way)
#pragma unroll
    for (int k = 0; k < 8; ++k) sts64(dst + 16 * k, o[k]);
}

// identity / x2-up: one lane item = row i8 of one block pair (tokens 0 / 1)
__device__ __forceinline__ void r_small_load(uint32_t info, int lane, const unsigned char* __restrict__ plane, int b0, int tok_stride,
                                             int4& a0, int4& a1) {
    const int inf0 = lds32(info + 4 * b0), inf1 = lds32(info + 4 * (b0 + tok_stride));
    const uint32_t lc = (lane & 7) * 16;
    a0 = ldg128(plane, (uint32_t(inf0) & 0xffffu) * 128u + lc);
    a1 = ldg128(plane, (uint32_t(inf1) & 0xffffu) * 128u + lc);
}
template <int MODE>
__device__ __forceinline__ void r_small_compute(const int4& a0, const int4& a1, uint32_t info, int b0, int tok_stride, uint32_t qrow,
                                                uint32_t dst, bool clamp) {
    int inf0 = 0, inf1 = 0;
    if (MODE == MODE_UP2) { inf0 = lds32(info + 4 * b0); inf1 = lds32(info + 4 * (b0 + tok_stride)); }
    float q[8], cq[8];
    load_q(qrow, q, cq);
    p2 x[8];
    dequant8_p(a0, a1, q, cq, clamp, x);
    if (MODE == MODE_UP2) {
        p2 o[8];
        up2_1d_p<1>(x, (inf0 >> 17) & 1, (inf1 >> 17) & 1, o);
#pragma unroll
        for (int k = 0; k < 8; ++k) sts64(dst + 16 * k, o[k]);
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) sts64(dst + 16 * k, x[k]);
    }
}

// ---- ops on one block column (8 values x 2 tokens); packed twin of k0_fused.cu run_ops ----------------------------------------
// z0 / z1: index of the op that zeroes the block of token 0 / 1 (-1: none).  A zeroed half runs the ops before its zeroing op on
// the data of a valid block (values stay in range, table indices stay valid) and is then overwritten with 0, which is what the
// reference computes from that op on.
__device__ __forceinline__ float dc_op(float d, const rgbnm_plan_op& op, int code, int comp, int k, const rgbnm_k0_tables& tb,
                                       const float* __restrict__ stats, int img) {
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
        if (comp == 0 && tb.equalize_lut != nullptr)
            d = float(tb.equalize_lut[(size_t(img) * RGBNM_MAX_OPS + k) * 2048 + int(d) + 1024]);
    }
    return clampf(d);
}

__device__ __forceinline__ bool solarize_hit(const rgbnm_plan& pl, const rgbnm_k0_tables& tb, int img, int comp, int k, int r, int c) {
    position_at_op<GRID_Y>(pl, comp, k, r, c);
    if (comp != 0) { r *= 2; c *= 2; }
    const bool inside = r >= 0 && c >= 0 && r < GRID_Y && c < GRID_Y;
    return inside && tb.equalize_lut[(size_t(img) * RGBNM_MAX_OPS + k) * 2048 + r * GRID_Y + c] != 0;
}

__device__ __forceinline__ bool run_ops_p(p2 (&v)[8], const rgbnm_plan& pl, int comp, int c, int z0, int z1, const rgbnm_k0_tables& tb,
                                          const float* __restrict__ stats, int img, int fr, int fc0, int fc_step) {
    bool T = false;
    if (pl.flip && (c & 1)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmul2(v[i], bq(-1.0f));
    }
    if (pl.train) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = clamp2(v[i]);            // custom_transforms.py:1107-1108
    }
    for (int k = 0; k < pl.n_ops; ++k) {
        if (k == z0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = pk(0.0f, hi_of(v[i]));
        }
        if (k == z1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = pk(lo_of(v[i]), 0.0f);
        }
        const rgbnm_plan_op& op = pl.ops[k];
        const int code = op.code;
        if (code == RGBNM_OP_ROT90) {
            T = !T;
            const bool rows = op.p[0] > 0;
            const bool by_phys_col = (rows == T);   // logical row == physical column iff T
            if (by_phys_col) {
                if (c & 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = clamp_hi2(fmul2(v[i], bq(-1.0f)));
                }
            } else {
#pragma unroll
                for (int i = 1; i < 8; i += 2) v[i] = clamp_hi2(fmul2(v[i], bq(-1.0f)));
            }
        } else if (code == RGBNM_OP_SHARPNESS || code == RGBNM_OP_MIDFREQ) {
            if (comp == 0) {
                const float* F = tb.filters + op.p[0] * 64 + c;   // symmetric: F[i][c] == F[c][i]
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = rint2(clamp2(fmul2(v[i], bq(__ldg(F + 8 * i)))));
            }
        } else if (code == RGBNM_OP_INVERT) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = clamp_hi2(fmul2(v[i], bq(-1.0f)));
        } else if (code == RGBNM_OP_SOLARIZE) {
            if (tb.equalize_lut != nullptr) {
                if (solarize_hit(pl, tb, img, comp, k, fr, fc0)) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = pk(fminf(-lo_of(v[i]), CLAMP_HI), hi_of(v[i]));
                }
                if (solarize_hit(pl, tb, img, comp, k, fr, fc0 + fc_step)) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = pk(lo_of(v[i]), fminf(-hi_of(v[i]), CLAMP_HI));
                }
            }
        } else if (code == RGBNM_OP_FREQ_ENHANCE) {
            // every coefficient but DCT[0,0] (physical (row 0, column 0) whatever the transpose flag) * f, rounded, clamped
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i != 0 || c != 0) v[i] = clamp2(rint2(fmul2(v[i], bq(op.f))));
        } else if (c == 0) {
            // DC-only ops: one value of one lane in eight
            v[0] = pk(dc_op(lo_of(v[0]), op, code, comp, k, tb, stats, img), dc_op(hi_of(v[0]), op, code, comp, k, tb, stats, img));
        }
        // geometric / zeroing ops: handled by trace_back
    }
    return T;
}

__device__ __forceinline__ p2 to_range2(p2 v) {
    // packed twin of k0_fused.cu to_range: exact n / 2040 (one Newton step), then * 2 - 1
    constexpr float R = 1.0f / 2040.0f;
    const p2 n = fadd2(v, bq(1024.0f));
    const p2 q0 = fmul2(n, bq(R));
    const p2 e = ffma2(q0, bq(-2040.0f), n);
    const p2 z = ffma2(e, bq(R), q0);
    return ffma2(z, bq(2.0f), bq(-1.0f));
}

__device__ __forceinline__ unsigned bf2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<unsigned*>(&t);
}

// ---- block bookkeeping of one quad: lanes 0..23 trace one block each into ws.info
//      (block id = p*12 + [luma: tok*4 + bi*2 + bj | chroma: 8 + tok*2 + comp-1]) ------------------------------------------------
__device__ __forceinline__ void trace_quad(WarpSmem& ws, int lane, int tr, int tp, int mode, int wb, int wc) {
    const rgbnm_plan& pl = ws.plan;
    if (lane < 24) {
        const int p = lane >= 12, bb = lane - 12 * p;
        int comp, r, c;
        if (bb < 8) {
            comp = 0;
            r = 2 * (2 * tr + p) + ((bb >> 1) & 1);
            c = 2 * (2 * tp + (bb >> 2)) + (bb & 1);
        } else {
            comp = 1 + ((bb - 8) & 1);
            r = 2 * tr + p;
            c = 2 * tp + ((bb - 8) >> 1);
        }
        const Trace t = trace_back<GRID_Y>(pl, comp, r, c);
        const int ci = comp == 0 ? pl.crop_i : (pl.crop_i >> 1);
        const int cj = comp == 0 ? pl.crop_j : (pl.crop_j >> 1);
        int sr, sc, chr = 0, chc = 0;
        if (mode == MODE_DOWN2) { sr = ci + 2 * t.r; sc = cj + 2 * t.c; }
        else if (mode == MODE_IDENT) { sr = ci + t.r; sc = cj + t.c; }
        else { sr = ci + (t.r >> 1); sc = cj + (t.c >> 1); chr = t.r & 1; chc = t.c & 1; }
        const int W = comp == 0 ? wb : wc;
        ws.info[lane] = pack_info(sr * W + sc, chr, chc, t.zero);
    }
}

// first luma loads of a quad (x2-down: round (pair 0, block row 0); else the single round of both pairs: A.l0 / A.l1 = pair 0
// tokens 0 / 1, A.r0 / A.r1 = pair 1).
__device__ __forceinline__ void luma_first_loads(uint32_t info, int lane, const unsigned char* __restrict__ y_img, int wb, int mode,
                                                 RowLoads& A) {
    if (mode == MODE_DOWN2) {
        A = r_down2_load(info, lane, y_img, wb * 128, 0, 4);
    } else {
        const int b3 = (lane >> 3) & 1, b4 = lane >> 4;          // lane = (row i8, block row b3, block column b4)
        r_small_load(info, lane, y_img, b3 * 2 + b4, 4, A.l0, A.l1);
        r_small_load(info, lane, y_img, 12 + b3 * 2 + b4, 4, A.r0, A.r1);
    }
}

// ---- one quad -------------------------------------------------------------------------------------------------------------------
// MODE_T >= 0 / NOCLAMP_T: compile-time resize case / "dequantisation clamp proven idle" (the x2-down, no-clamp instance is the
// hot one in both the eval geometry and the training mix); MODE_T = -1: everything decided at run time.
// On entry ws.info is traced and A holds the quad's first luma loads.  `sched` / `grabbed`: ticket of the dynamic quad queue, drawn
// before the read-out phase (see the kernel).
// NOSUB: RGBNM_K0_LAYOUT_VIT16_NOSUB (`--no_subblock`, plainvit.py:173-216 with use_subblock = False): no A16 products; the luma part of
// the token is the 16 x 16 tile of the four un-converted blocks, row-major ('b c (h pdh) (w pdw) p1 p2 -> b c h w (pdh p1) (pdw p2)').
template <int OUT_MODE, int MODE_T, bool NOCLAMP_T, bool NOSUB>
__device__ __forceinline__ void process_quad(WarpSmem& ws, const LaneK& K, int lane, int img, int tr, int tp, int mode_rt,
                                             const unsigned char* __restrict__ y_img, const unsigned char* __restrict__ c_img,
                                             const rgbnm_k0_tables& tb, const float* __restrict__ stats, void* __restrict__ out_, int wb,
                                             int hc, int wc, RowLoads& A, unsigned* __restrict__ sched, unsigned& grabbed) {
    const rgbnm_plan& pl = ws.plan;
    const int mode = MODE_T >= 0 ? MODE_T : mode_rt;
    const bool clamp = NOCLAMP_T ? false : (pl.clamp_in != 0);
    const uint32_t info = K.info;
    const int b3 = (lane >> 3) & 1, b4 = lane >> 4;
    // ---- R, luma ----------------------------------------------------------------------------------------------------------------
    if (mode == MODE_DOWN2) {
        // rounds (pair, block row) = (0,0) (0,1) (1,0) (1,1); the loads of round r+2 are issued as round r is consumed
        const int wbytes = wb * 128;
        RowLoads B = r_down2_load(info, lane, y_img, wbytes, 2, 4);
        r_down2_compute(A, K.qrow, K.rst, clamp);
        A = r_down2_load(info, lane, y_img, wbytes, 12, 4);
        r_down2_compute(B, K.qrow, K.rst + 16 * PITCH, clamp);
        B = r_down2_load(info, lane, y_img, wbytes, 14, 4);
        r_down2_compute(A, K.qrow, K.rst + TILE_B, clamp);
        r_down2_compute(B, K.qrow, K.rst + TILE_B + 16 * PITCH, clamp);
    } else if (mode == MODE_UP2) {
        r_small_compute<MODE_UP2>(A.l0, A.l1, info, b3 * 2 + b4, 4, K.qrow, K.rst, clamp);
        r_small_compute<MODE_UP2>(A.r0, A.r1, info, 12 + b3 * 2 + b4, 4, K.qrow, K.rst + TILE_B, clamp);
    } else {
        r_small_compute<MODE_IDENT>(A.l0, A.l1, info, b3 * 2 + b4, 4, K.qrow, K.rst, clamp);
        r_small_compute<MODE_IDENT>(A.r0, A.r1, info, 12 + b3 * 2 + b4, 4, K.qrow, K.rst + TILE_B, clamp);
    }
    __syncwarp();

    // ---- C: three rounds (k = 0, 1: luma block rows; k = 2: chroma, whose row pass runs once the luma tokens are stored) ----------
    // lane = (pair p = b4, tile column c16)
    const int p = b4, c16 = lane & 15, half = c16 >> 3, c = c16 & 7;
    bool Tf = false;
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
        if (k == 2) {
            __syncwarp();                                                  // S complete (transposed blocks are written across columns)
            // dynamic schedule: ask for the next quad now, the answer is needed when this one is stored (atomic latency hidden,
            // and a quad is held by a busy warp for the last third of an iteration only)
            if (lane == 0) grabbed = atomicAdd(sched, 1u);
            if (OUT_MODE != RGBNM_K0_OUT_INT16_PLANES) {
                // ---- P2b: column half of the sub-block conversion, in place: S -> A16 . S ------------------------------------
                if (!NOSUB) {
                    p2 xl[8], xr[8], t[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { xl[i] = lds64(K.ccol + i * PITCH); xr[i] = lds64(K.ccol + (8 + i) * PITCH); }
                    a16_1d_p(xl, xr, t);
#pragma unroll
                    for (int i = 0; i < 16; ++i) sts64(K.ccol + i * PITCH, t[i]);
                }
                __syncwarp();
                // ---- P3: row half ((.) A16^T) + stores: lane = (pair p, tile row c16) ------------------------------------------
                {
                    p2 xl[8], xr[8], o[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        lds128(K.prow + 32 * j, xl[2 * j], xr[2 * j]);              // slots (2c, 2c + 1) = columns (c, c + 8)
                        lds128(K.prow + 32 * j + 16, xl[2 * j + 1], xr[2 * j + 1]);
                    }
                    if (NOSUB) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) { o[j] = xl[j]; o[8 + j] = xr[j]; }      // tile row as it stands
                    } else {
                        a16_1d_p(xl, xr, o);
                    }
                    const size_t off = (size_t(img) * TOKENS + (2 * tr + p) * 14 + 2 * tp) * FEAT + c16 * 16;
                    constexpr int HALF2 = 8;                       // element offset of the second 8 outputs of the row
                    if (OUT_MODE == RGBNM_K0_OUT_F32) {
                        float* f0 = reinterpret_cast<float*>(out_) + off;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float4* d0 = reinterpret_cast<float4*>(f0 + (j >> 1) * HALF2 + (j & 1) * 4);
                            float4* d1 = reinterpret_cast<float4*>(f0 + FEAT + (j >> 1) * HALF2 + (j & 1) * 4);
                            *d0 = make_float4(lo_of(o[4 * j]), lo_of(o[4 * j + 1]), lo_of(o[4 * j + 2]), lo_of(o[4 * j + 3]));
                            *d1 = make_float4(hi_of(o[4 * j]), hi_of(o[4 * j + 1]), hi_of(o[4 * j + 2]), hi_of(o[4 * j + 3]));
                        }
                    } else {
                        __nv_bfloat16* b0 = reinterpret_cast<__nv_bfloat16*>(out_) + off;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            uint4* d0 = reinterpret_cast<uint4*>(b0 + j * HALF2);
                            uint4* d1 = reinterpret_cast<uint4*>(b0 + FEAT + j * HALF2);
                            *d0 = make_uint4(bf2(lo_of(o[8 * j]), lo_of(o[8 * j + 1])), bf2(lo_of(o[8 * j + 2]), lo_of(o[8 * j + 3])),
                                               bf2(lo_of(o[8 * j + 4]), lo_of(o[8 * j + 5])), bf2(lo_of(o[8 * j + 6]), lo_of(o[8 * j + 7])));
                            *d1 = make_uint4(bf2(hi_of(o[8 * j]), hi_of(o[8 * j + 1])), bf2(hi_of(o[8 * j + 2]), hi_of(o[8 * j + 3])),
                                               bf2(hi_of(o[8 * j + 4]), hi_of(o[8 * j + 5])), bf2(hi_of(o[8 * j + 6]), hi_of(o[8 * j + 7])));
                        }
                    }
                }
                __syncwarp();
            }
            // ---- R, chroma: tile of pair p, rows = source rows, 16 columns = [Cb 8 | Cr 8] ----------------------------------------
            if (mode == MODE_DOWN2) {
                // lane bit 4 = component (the block-column `half` of r_down2_*); chroma block ids 8 + tok*2 + (comp-1), token stride 2
                const unsigned char* plane = c_img + size_t(b4) * hc * wc * 128;
                const RowLoads CA = r_down2_load(info, lane, plane, wc * 128, 8, 2);
                const RowLoads CB = r_down2_load(info, lane, plane, wc * 128, 20, 2);
                const uint32_t qr = K.qrow + (1 + b4) * 8 * QROW_B;
                r_down2_compute(CA, qr, K.rst, clamp);
                r_down2_compute(CB, qr, K.rst + TILE_B, clamp);
            } else {
                // lane = (row i8, component b3, pair b4)
                const unsigned char* plane = c_img + size_t(b3) * hc * wc * 128;
                int4 a0, a1;
                r_small_load(info, lane, plane, b4 * 12 + 8 + b3, 2, a0, a1);
                const uint32_t qr = K.qrow + (1 + b3) * 8 * QROW_B;
                if (mode == MODE_UP2) r_small_compute<MODE_UP2>(a0, a1, info, b4 * 12 + 8 + b3, 2, qr, K.rst_c, clamp);
                else r_small_compute<MODE_IDENT>(a0, a1, info, b4 * 12 + 8 + b3, 2, qr, K.rst_c, clamp);
            }
            __syncwarp();
        }
        // ---- the column pass proper ----------------------------------------------------------------------------------------
        const int b0 = p * 12 + (k < 2 ? k * 2 + half : 8 + half);
        const int inf0 = lds32(info + 4 * b0), inf1 = lds32(info + 4 * (b0 + (k < 2 ? 4 : 2)));
        const int z0 = ((inf0 >> 18) & 7) - 1, z1 = ((inf1 >> 18) & 7) - 1;
        const int comp = k < 2 ? 0 : 1 + half;
        p2 v[8];
        if (mode == MODE_DOWN2) {
            const uint32_t colp = K.ccol + (k < 2 ? k * 16 : 0) * PITCH;
            p2 xl[8], xr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { xl[i] = lds64(colp + i * PITCH); xr[i] = lds64(colp + (8 + i) * PITCH); }
            down2_1d_p<1, 2>(xl, xr, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = rint2(v[i]);                 // torch.round -> int16 (dct_ops.py:577-578)
        } else {
            const uint32_t colp = K.ccol + (k < 2 ? k * 8 : 0) * PITCH;
            if (mode == MODE_UP2) {
                p2 x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = lds64(colp + i * PITCH);
                up2_1d_p<2>(x, (inf0 >> 16) & 1, (inf1 >> 16) & 1, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = rint2(v[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = lds64(colp + i * PITCH);
            }
        }
        __syncwarp();                                                      // this round's tile reads are done: S may overwrite
        const int fr = k < 2 ? 2 * (2 * tr + p) + k : 2 * tr + p;
        const int fc0 = k < 2 ? 2 * (2 * tp) + half : 2 * tp;
        if (pl.train | pl.flip | pl.n_ops) Tf = run_ops_p(v, pl, comp, c, z0, z1, tb, stats, img, fr, fc0, k < 2 ? 2 : 1);

        if (OUT_MODE == RGBNM_K0_OUT_INT16_PLANES) {
            int16_t* o16 = reinterpret_cast<int16_t*>(out_) + size_t(img) * PLANE_ELEMS;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int blk;
                if (k < 2) blk = fr * GRID_Y + fc0 + 2 * h;
                else blk = GRID_Y * GRID_Y + ((comp - 1) * 14 + fr) * 14 + fc0 + h;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int li = Tf ? c : i, lj = Tf ? i : c;
                    o16[size_t(blk) * 64 + li * 8 + lj] = int16_t(int(h ? hi_of(v[i]) : lo_of(v[i])));
                }
            }
        } else {
            // S tile rows (k*8 ..) of pair p, slots 2 * column + half: element (i, c), or (c, i) for a transposed block -- one
            // store sequence with a run-time stride (two branches made ptxas copy every pair into a fixed register pair)
            const uint32_t S = K.ccol - c * 16 + (k < 2 ? k * 8 : 0) * PITCH + c * (Tf ? PITCH : 16);
            const uint32_t stride = Tf ? 16 : PITCH;
#pragma unroll
            for (int i = 0; i < 8; ++i) sts64(S + i * stride, to_range2(v[i]));
        }
    }
    __syncwarp();
    if (OUT_MODE == RGBNM_K0_OUT_INT16_PLANES) return;

    // ---- chroma out: lane = (pair p = b4, block row i8, column group b3): 16-byte loads give (Cb, Cr) of 4 columns -----------------
    {
        p2 cb[4], cr[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) lds128(K.rst_c - b3 * 8 + b3 * 64 + 16 * m, cb[m], cr[m]);     // row i8, slots 8 b3 + 2m, + 1
        const size_t off = (size_t(img) * TOKENS + (2 * tr + p) * 14 + 2 * tp) * FEAT + 256 + (lane & 7) * 8 + b3 * 4;
        if (OUT_MODE == RGBNM_K0_OUT_F32) {
            float* o = reinterpret_cast<float*>(out_) + off;
            *reinterpret_cast<float4*>(o) = make_float4(lo_of(cb[0]), lo_of(cb[1]), lo_of(cb[2]), lo_of(cb[3]));
            *reinterpret_cast<float4*>(o + 64) = make_float4(lo_of(cr[0]), lo_of(cr[1]), lo_of(cr[2]), lo_of(cr[3]));
            *reinterpret_cast<float4*>(o + FEAT) = make_float4(hi_of(cb[0]), hi_of(cb[1]), hi_of(cb[2]), hi_of(cb[3]));
            *reinterpret_cast<float4*>(o + FEAT + 64) = make_float4(hi_of(cr[0]), hi_of(cr[1]), hi_of(cr[2]), hi_of(cr[3]));
        } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_) + off;
            *reinterpret_cast<uint2*>(o) = make_uint2(bf2(lo_of(cb[0]), lo_of(cb[1])), bf2(lo_of(cb[2]), lo_of(cb[3])));
            *reinterpret_cast<uint2*>(o + 64) = make_uint2(bf2(lo_of(cr[0]), lo_of(cr[1])), bf2(lo_of(cr[2]), lo_of(cr[3])));
            *reinterpret_cast<uint2*>(o + FEAT) = make_uint2(bf2(hi_of(cb[0]), hi_of(cb[1])), bf2(hi_of(cb[2]), hi_of(cb[3])));
            *reinterpret_cast<uint2*>(o + FEAT + 64) = make_uint2(bf2(hi_of(cr[0]), hi_of(cr[1])), bf2(hi_of(cr[2]), hi_of(cr[3])));
        }
    }
    __syncwarp();
}

template <int OUT_MODE, bool NOSUB>
__global__ void __launch_bounds__(WARPS * 32, CTAS_PER_SM)
k0_vit2_kernel(const int16_t* __restrict__ y, const int16_t* __restrict__ cbcr, const int16_t* __restrict__ quant,
               const rgbnm_plan* __restrict__ plans, rgbnm_k0_tables tb, const float* __restrict__ stats_all, void* __restrict__ out_,
               int n_images, int hb, int wb, unsigned* __restrict__ sched) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSmem& ws = reinterpret_cast<WarpSmem*>(smem_raw)[warp];
    LaneK K;
    {
        const uint32_t base = uint32_t(__cvta_generic_to_shared(smem_raw)) + warp * uint32_t(sizeof(WarpSmem));
        const int i8 = lane & 7, b3 = (lane >> 3) & 1, b4 = lane >> 4;
        K.T = base;
        K.rst = base + (b3 * 8 + i8) * PITCH + b4 * 8;
        K.rst_c = base + b4 * TILE_B + i8 * PITCH + b3 * 8;
        K.ccol = base + b4 * TILE_B + (2 * (lane & 7) + b3) * 8;          // slot of tile column lane & 15
        K.prow = base + b4 * TILE_B + (lane & 15) * PITCH;
        K.qrow = base + uint32_t(offsetof(WarpSmem, qt)) + i8 * QROW_B;
        K.info = base + uint32_t(offsetof(WarpSmem, info));
    }
    const long long nq = (long long)n_images * QUADS_PER_IMAGE;
    const long long gw = (long long)blockIdx.x * WARPS + warp, tw = (long long)gridDim.x * WARPS;
    const int hc = hb >> 1, wc = wb >> 1;
    int cur = -1, mode = MODE_BAD;
    RowLoads A;
    A.l0 = A.r0 = A.l1 = A.r1 = make_int4(0, 0, 0, 0);
    // Quads are dealt DYNAMICALLY: a warp's first quad is its global index, every further one comes from a device counter
    // (sched[0]; quad = warps + ticket).  The work of a quad depends on the image's plan (x2-down reads 4x the coefficients of x2-up,
    // RandAugment ops add column-pass work) and a launch holds only ~5 quads per warp, so fixed ranges left warps idle for the
    // last quarter of the kernel.  The last warp to leave (sched[1] counts them) zeroes both words for the next launch.
    long long q = gw;
    while (q < nq) {
        const int img = int(q / QUADS_PER_IMAGE), rem = int(q - (long long)img * QUADS_PER_IMAGE);
        if (img != cur) {
            __syncwarp();
            const int* psrc = reinterpret_cast<const int*>(plans + img);
            if (lane < int(sizeof(rgbnm_plan) / 4)) reinterpret_cast<int*>(&ws.plan)[lane] = __ldg(psrc + lane);
            for (int k = lane; k < 192; k += 32) {
                const float qv = float(__ldg(quant + size_t(img) * 192 + k));
                const uint32_t a = K.T + uint32_t(offsetof(WarpSmem, qt)) + (k >> 3) * QROW_B + (k & 7) * 4;     // (component, row) = k / 8, column = k % 8
                sts32f(a, qv);
                sts32f(a + 32, -DEQ_BIAS * qv);
            }
            __syncwarp();
            cur = img;
            mode = mode_of(ws.plan.crop_size, GRID_Y);
        }
        unsigned grabbed = 0;
        if (mode != MODE_BAD) {
            const int tr = rem / 7, tp = rem - tr * 7;
            const unsigned char* y_img = reinterpret_cast<const unsigned char*>(y + size_t(img) * hb * wb * 64);
            const unsigned char* c_img = reinterpret_cast<const unsigned char*>(cbcr + size_t(img) * 2 * hc * wc * 64);
            trace_quad(ws, lane, tr, tp, mode, wb, wc);
            __syncwarp();
            luma_first_loads(K.info, lane, y_img, wb, mode, A);
            const float* stats = stats_all + size_t(img) * RGBNM_MAX_OPS * 2;
            if (OUT_MODE != RGBNM_K0_OUT_INT16_PLANES && mode == MODE_DOWN2 && ws.plan.clamp_in == 0)
                process_quad<OUT_MODE, MODE_DOWN2, true, NOSUB>(ws, K, lane, img, tr, tp, mode, y_img, c_img, tb, stats, out_, wb, hc, wc, A, sched, grabbed);
            else
                process_quad<OUT_MODE, -1, false, NOSUB>(ws, K, lane, img, tr, tp, mode, y_img, c_img, tb, stats, out_, wb, hc, wc, A, sched, grabbed);
        } else if (lane == 0) {
            grabbed = atomicAdd(sched, 1u);
        }
        q = tw + (long long)__shfl_sync(0xffffffffu, grabbed, 0);
    }
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(sched + 1, 1u) == unsigned(tw - 1)) {
            sched[0] = 0;
            sched[1] = 0;
            __threadfence();
        }
    }
}

}  // namespace k0v2

// Scheduler words of the dynamic quad queue: {next ticket, warps that left}.  A launch leaves its pair zeroed, launches of one stream
// serialise, and concurrent launches on different streams get different pairs (round robin over SCHED_SLOTS).
constexpr int SCHED_SLOTS = 256;
__device__ unsigned g_k0v2_sched[SCHED_SLOTS][2];

static int next_sched_slot(unsigned** out) {
    static std::atomic<unsigned> turn{0};
    static std::atomic<unsigned*> base[64];
    int dev = 0;
    RGBNM_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return RGBNM_ERR_ARG;
    unsigned* b = base[dev].load(std::memory_order_acquire);
    if (b == nullptr) {
        void* p = nullptr;
        RGBNM_CUDA_CHECK(cudaGetSymbolAddress(&p, g_k0v2_sched));
        b = static_cast<unsigned*>(p);
        base[dev].store(b, std::memory_order_release);
    }
    *out = b + 2 * (turn.fetch_add(1, std::memory_order_relaxed) % SCHED_SLOTS);
    return RGBNM_OK;
}

template <int OUT_MODE, bool NOSUB>
static int launch_vit2(const int16_t* y, const int16_t* cbcr, const int16_t* quant, const rgbnm_plan* plans, const rgbnm_k0_tables* tables,
                       const float* stats, void* out, int n, int hb, int wb, cudaStream_t st) {
    using namespace k0v2;
    static int sms = 0;
    const size_t smem = sizeof(WarpSmem) * WARPS;
    if (sms == 0) {
        int dev = 0, v = 0;
        RGBNM_CUDA_CHECK(cudaGetDevice(&dev));
        RGBNM_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        RGBNM_CUDA_CHECK(cudaFuncSetAttribute(k0_vit2_kernel<OUT_MODE, NOSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        RGBNM_CUDA_CHECK(cudaFuncSetAttribute(k0_vit2_kernel<OUT_MODE, NOSUB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                              cudaSharedmemCarveoutMaxShared));
        sms = v;
    }
    const long long nq = (long long)n * QUADS_PER_IMAGE;
    long long ctas = (nq + WARPS - 1) / WARPS;
    if (ctas > (long long)sms * CTAS_PER_SM) ctas = (long long)sms * CTAS_PER_SM;
    unsigned* sched = nullptr;
    if (int rc = next_sched_slot(&sched)) return rc;
    k0_vit2_kernel<OUT_MODE, NOSUB><<<int(ctas), WARPS * 32, smem, st>>>(y, cbcr, quant, plans, *tables, stats, out, n, hb, wb, sched);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

// entry used by rgbnm_k0_fused_ex for RGBNM_K0_LAYOUT_VIT16
int rgbnm_k0_vit2_launch(const int16_t* y, const int16_t* cbcr, const int16_t* quant, const rgbnm_plan* plans, const rgbnm_k0_tables* tables,
                         const float* stats, void* out, int out_mode, int nosub, int n, int hb, int wb, cudaStream_t st) {
#define RGBNM_VIT2_CASE(OM) \
    if (out_mode == OM) return nosub ? launch_vit2<OM, true>(y, cbcr, quant, plans, tables, stats, out, n, hb, wb, st) \
                                     : launch_vit2<OM, false>(y, cbcr, quant, plans, tables, stats, out, n, hb, wb, st)
    RGBNM_VIT2_CASE(RGBNM_K0_OUT_F32);
    RGBNM_VIT2_CASE(RGBNM_K0_OUT_BF16);
    if (out_mode == RGBNM_K0_OUT_INT16_PLANES)      // the int16 planes do not depend on the embedding layout
        return launch_vit2<RGBNM_K0_OUT_INT16_PLANES, false>(y, cbcr, quant, plans, tables, stats, out, n, hb, wb, st);
#undef RGBNM_VIT2_CASE
    return RGBNM_ERR_ARG;
}
