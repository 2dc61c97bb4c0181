// Kernels of the SwinV2 DCT forward path around the tensor-core contractions (sm_100a); SURVEY.md 8a row a33.
//   post-norm LayerNorm + residual      x = shortcut + norm(branch(x))            models/swinv2.py:302-306
//   window attention, forward           cosine attention + continuous relative position bias + shift mask
//                                                                                 models/swinv2.py:143-182, 244-300
//   patch-merging gather                x0 | x1 | x2 | x3 concat                  models/swinv2.py:346-362
//   token mean                          AdaptiveAvgPool1d(1)                      models/swinv2.py:697-699
// Window partition / cyclic shift / window reverse (swinv2.py:39-66, 283-300) are index maps: the attention kernel
// gathers each window's tokens from their image positions and scatters the result back, nothing is materialised.
// All four are HBM-bound by design; the attention arithmetic (64 x 64 x 32 per window and head) runs on warp-level
// tensor-core MMAs with S / P / O in registers (window_attn_mma_kernel); the first, CUDA-core fp32 version
// (window_attn_fwd_kernel) is kept behind RGBNM_WATTN_SIMT=1 for A/B runs.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "../../include/rgbnm_b200.h"
#include "common.cuh"

namespace swink {

constexpr int LN_WARPS = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float2 bf2_to_f2(unsigned w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ unsigned f2_to_bf2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<unsigned*>(&v);
}

// ------------------------------------------------------------------------------------------
// y = (res ? res : 0) + (x - mean) * rstd * gamma + beta     one warp per row, any even E <= 64 * MAXP
// ------------------------------------------------------------------------------------------
template <int MAXP>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_res_fwd_kernel(const unsigned* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                  const unsigned* __restrict__ res, unsigned* __restrict__ y, int rows, int E, float eps) {
    const int lane = threadIdx.x & 31;
    const int pairs = E >> 1;
    const float inv_e = 1.0f / float(E);
    const int stride = gridDim.x * LN_WARPS;
    int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    // software pipeline over rows: the next row's x / residual loads are in flight during this row's two reductions
    unsigned nx[MAXP], nr[MAXP];
    auto fetch = [&](int r) {
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
            const int p = k * 32 + lane;
            nx[k] = p < pairs ? __ldg(x + size_t(r) * pairs + p) : 0u;
            nr[k] = (res != nullptr && p < pairs) ? __ldg(res + size_t(r) * pairs + p) : 0u;
        }
    };
    if (row < rows) fetch(row);
    for (; row < rows; row += stride) {
        float2 v[MAXP];
        unsigned cr[MAXP];
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
            v[k] = bf2_to_f2(nx[k]);
            cr[k] = nr[k];
            s += v[k].x + v[k].y;
        }
        if (row + stride < rows) fetch(row + stride);
        const float mean = warp_sum(s) * inv_e;
        float q = 0.0f;
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
            if (k * 32 + lane < pairs) {
                const float a = v[k].x - mean, b = v[k].y - mean;
                q += a * a + b * b;
            }
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_e + eps);
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
            const int p = k * 32 + lane;
            if (p < pairs) {
                const float2 g = *reinterpret_cast<const float2*>(gamma + 2 * p);
                const float2 b = *reinterpret_cast<const float2*>(beta + 2 * p);
                const float2 r = bf2_to_f2(cr[k]);
                const float o0 = (v[k].x - mean) * rstd * g.x + b.x + r.x;
                const float o1 = (v[k].y - mean) * rstd * g.y + b.y + r.y;
                y[size_t(row) * pairs + p] = f2_to_bf2(o0, o1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Window attention forward.  One CTA (64 threads) = one (window, head); thread t = in-window token t = iy * 8 + ix.
//   q, k normalised (F.normalize, eps 1e-12); s_ij = <q_i, k_j> * scale_h + bias_h[i][j] (+ -100 across shift regions);
//   softmax; o_i = sum_j p_ij v_j.                                                 swinv2.py:152-177
// qkv bf16 [B * H * W][3 * C], columns (which, head, d) as produced by qkv.reshape(B_, N, 3, heads, -1) (swinv2.py:154).
// ------------------------------------------------------------------------------------------
constexpr int WS = 8, WT = 64, HD = 32;

__global__ void __launch_bounds__(WT)
window_attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, const float* __restrict__ bias,
                       const float* __restrict__ scale, int H, int W, int C, int shift) {
    __shared__ __align__(16) float Ks[WT][HD];
    __shared__ __align__(16) float Vs[WT][HD];
    __shared__ int region[WT];
    const int t = threadIdx.x, head = blockIdx.y;
    const int wpr = W / WS, wpi = (H / WS) * wpr;
    const int img = blockIdx.x / wpi, wrem = blockIdx.x - img * wpi;
    const int wy = wrem / wpr, wx = wrem - wy * wpr;
    // position in the cyclically shifted image, and the image token it was rolled from (torch.roll by -shift)
    const int sy = wy * WS + (t >> 3), sx = wx * WS + (t & 7);
    int py = sy + shift, px = sx + shift;
    if (py >= H) py -= H;
    if (px >= W) px -= W;
    const size_t tok = (size_t(img) * H + py) * W + px;
    int reg = 0;
    if (shift > 0) {     // img_mask regions of swinv2.py:227-238, in shifted coordinates
        const int hr = sy < H - WS ? 0 : (sy < H - shift ? 1 : 2);
        const int wr = sx < W - WS ? 0 : (sx < W - shift ? 1 : 2);
        reg = 3 * hr + wr;
    }
    region[t] = reg;

    const __nv_bfloat16* row = qkv + tok * (3 * size_t(C)) + head * HD;
    float q[HD];
    float qn = 0.0f, kn = 0.0f;
    float kreg[HD];
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(row) + c);
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(row + C) + c);
        const uint4 d = __ldg(reinterpret_cast<const uint4*>(row + 2 * C) + c);
        const unsigned aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 qa = bf2_to_f2(aw[e]), kb = bf2_to_f2(bw[e]), vd = bf2_to_f2(dw[e]);
            q[c * 8 + 2 * e] = qa.x;
            q[c * 8 + 2 * e + 1] = qa.y;
            kreg[c * 8 + 2 * e] = kb.x;
            kreg[c * 8 + 2 * e + 1] = kb.y;
            qn += qa.x * qa.x + qa.y * qa.y;
            kn += kb.x * kb.x + kb.y * kb.y;
            Vs[t][c * 8 + 2 * e] = vd.x;
            Vs[t][c * 8 + 2 * e + 1] = vd.y;
        }
    }
    // F.normalize: x / max(||x||, 1e-12); the logit scale is folded into q
    const float qs = __ldg(scale + head) / fmaxf(sqrtf(qn), 1e-12f);
    const float ks = 1.0f / fmaxf(sqrtf(kn), 1e-12f);
#pragma unroll
    for (int d = 0; d < HD; ++d) {
        q[d] *= qs;
        Ks[t][d] = kreg[d] * ks;
    }
    __syncthreads();

    float s[WT];
    const float4* brow = reinterpret_cast<const float4*>(bias + (size_t(head) * WT + t) * WT);
    float m = -3.0e38f;
#pragma unroll
    for (int j4 = 0; j4 < WT / 4; ++j4) {
        const float4 b4 = __ldg(brow + j4);
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = j4 * 4 + e;
            const float4* kr = reinterpret_cast<const float4*>(Ks[j]);
            float acc = 0.0f;
#pragma unroll
            for (int d4 = 0; d4 < HD / 4; ++d4) {
                const float4 kk = kr[d4];
                acc = fmaf(q[4 * d4], kk.x, acc);
                acc = fmaf(q[4 * d4 + 1], kk.y, acc);
                acc = fmaf(q[4 * d4 + 2], kk.z, acc);
                acc = fmaf(q[4 * d4 + 3], kk.w, acc);
            }
            acc += bb[e];
            if (shift > 0 && region[j] != reg) acc += -100.0f;      // attn_mask value of swinv2.py:242
            s[j] = acc;
            m = fmaxf(m, acc);
        }
    }
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < WT; ++j) {
        s[j] = __expf(s[j] - m);
        sum += s[j];
    }
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.0f;
#pragma unroll
    for (int j = 0; j < WT; ++j) {
        const float4* vr = reinterpret_cast<const float4*>(Vs[j]);
#pragma unroll
        for (int d4 = 0; d4 < HD / 4; ++d4) {
            const float4 vv = vr[d4];
            o[4 * d4] = fmaf(s[j], vv.x, o[4 * d4]);
            o[4 * d4 + 1] = fmaf(s[j], vv.y, o[4 * d4 + 1]);
            o[4 * d4 + 2] = fmaf(s[j], vv.z, o[4 * d4 + 2]);
            o[4 * d4 + 3] = fmaf(s[j], vv.w, o[4 * d4 + 3]);
        }
    }
    const float inv = 1.0f / sum;
    uint4* dst = reinterpret_cast<uint4*>(out + tok * size_t(C) + head * HD);
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
        dst[c] = make_uint4(f2_to_bf2(o[8 * c] * inv, o[8 * c + 1] * inv), f2_to_bf2(o[8 * c + 2] * inv, o[8 * c + 3] * inv),
                            f2_to_bf2(o[8 * c + 4] * inv, o[8 * c + 5] * inv), f2_to_bf2(o[8 * c + 6] * inv, o[8 * c + 7] * inv));
    }
}

// ------------------------------------------------------------------------------------------
// The same on warp-level tensor-core MMAs (mma.sync m16n8k16, bf16 x bf16 -> fp32).  64 x 64 x 32 problems sit below the
// 128-row tcgen05 tile (two windows per UMMA would waste half the MMA on cross-window scores and still pay a TMEM round
// trip per softmax), so this kernel keeps S, P and O in registers, flash-attention style:
//   CTA = 4 warps = one (window, head) per iteration, grid-strided over the windows of one head so that the head's
//   64 x 64 bias tile is staged in shared memory once; warp w owns query rows 16w .. 16w+15.
//   S = Q K^T on the RAW bf16 q / k (products exact in fp32); the cosine normalisation, logit scale, bias and shift mask
//   are applied to the fp32 accumulators: s = raw * (scale / |q_i|) * (1 / |k_j|) + bias_ij (+ -100) -- no extra bf16
//   rounding of normalised operands.  P (bf16) is re-used from the accumulator registers as the A operand of P V;
//   V sits transposed in shared memory so that every B fragment is one 32-bit load; all fragment loads are conflict-free
//   (row pitches 80 B / 144 B / 288 B).
// ------------------------------------------------------------------------------------------
constexpr int KS_LD = 40;     // bf16 per staged K row (32 + 8 pad)
constexpr int VT_LD = 72;     // bf16 per row of V^T (64 keys + 8 pad)
constexpr int BS_LD = 72;     // floats per staged bias row
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void mma16816(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// four 8x8 bf16 matrices from shared memory, one row address per lane (lanes 8m .. 8m+7 address the rows of matrix m);
// thread (g = lane / 4, t = lane % 4) receives elements [g][2t], [g][2t+1] of each matrix (.trans: [2t][g], [2t+1][g])
__device__ __forceinline__ void ldmatrix_x4(unsigned (&r)[4], const void* p) {
    const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(unsigned (&r)[4], const void* p) {
    const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
#ifndef RGBNM_WATTN_LDMATRIX
#define RGBNM_WATTN_LDMATRIX 1       // 0: 32-bit fragment loads + transposed V stores (first MMA version), kept for A/B builds
#endif
__device__ __forceinline__ float sumsq_bf2(unsigned w) {
    const float2 f = bf2_to_f2(w);
    return f.x * f.x + f.y * f.y;
}
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}

// image token of in-window position j of window `win` (cyclic shift folded in: torch.roll by -shift, swinv2.py:283-286)
__device__ __forceinline__ int window_token(int win, int j, int H, int W, int wpr, int wpi, int shift) {
    const int img = win / wpi, wrem = win - img * wpi;
    const int wy = wrem / wpr, wx = wrem - wy * wpr;
    int py = wy * WS + (j >> 3) + shift, px = wx * WS + (j & 7) + shift;
    if (py >= H) py -= H;
    if (px >= W) px -= W;
    return (img * H + py) * W + px;
}

struct WinLoad {          // one thread's share of a window's operands, in flight one window ahead of the arithmetic
    uint4 k0, k1, v0, v1;
    unsigned qa[2][4];
    int tok0, tok1;
};

__device__ __forceinline__ void window_fetch(WinLoad& L, const __nv_bfloat16* __restrict__ qkv, int win, int head, int C, int H, int W,
                                             int wpr, int wpi, int shift, int tid, int R0, int t) {
    const int r = tid >> 1, half = tid & 1;
    const __nv_bfloat16* row = qkv + size_t(window_token(win, r, H, W, wpr, wpi, shift)) * (3 * size_t(C)) + head * HD + half * 16;
    L.k0 = __ldg(reinterpret_cast<const uint4*>(row + C));
    L.k1 = __ldg(reinterpret_cast<const uint4*>(row + C) + 1);
    L.v0 = __ldg(reinterpret_cast<const uint4*>(row + 2 * C));
    L.v1 = __ldg(reinterpret_cast<const uint4*>(row + 2 * C) + 1);
    L.tok0 = window_token(win, R0, H, W, wpr, wpi, shift);
    L.tok1 = window_token(win, R0 + 8, H, W, wpr, wpi, shift);
    const unsigned* p0 = reinterpret_cast<const unsigned*>(qkv + size_t(L.tok0) * (3 * size_t(C)) + head * HD);
    const unsigned* p1 = reinterpret_cast<const unsigned*>(qkv + size_t(L.tok1) * (3 * size_t(C)) + head * HD);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        L.qa[ks][0] = __ldg(p0 + ks * 8 + t);
        L.qa[ks][1] = __ldg(p1 + ks * 8 + t);
        L.qa[ks][2] = __ldg(p0 + ks * 8 + 4 + t);
        L.qa[ks][3] = __ldg(p1 + ks * 8 + 4 + t);
    }
}

#ifndef RGBNM_WATTN_MINBLOCKS
#define RGBNM_WATTN_MINBLOCKS 5      // resident CTAs per SM the register allocation aims at; A/B in one box (profiles/r01_swin_ab_minblocks.log): 5 -> 7.17-7.29 ms, 6 -> 7.24-7.32, 8 -> 8.0
#endif
__global__ void __launch_bounds__(128, RGBNM_WATTN_MINBLOCKS)
window_attn_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, const float* __restrict__ bias,
                       const float* __restrict__ scale, int H, int W, int C, int shift, int n_windows) {
    __shared__ __align__(16) __nv_bfloat16 Ks[WT * KS_LD];
#if RGBNM_WATTN_LDMATRIX
    __shared__ __align__(16) __nv_bfloat16 Vs[WT * KS_LD];          // V row-major like K; B fragments through ldmatrix.trans
#else
    __shared__ __align__(16) __nv_bfloat16 Vt[HD * VT_LD];
#endif
    __shared__ __align__(16) float Bs[WT * BS_LD];
    __shared__ float rk[WT];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int head = blockIdx.y;
    for (int e = tid; e < WT * (WT / 4); e += 128) {
        const int row = e >> 4, c4 = e & 15;
        // scores are kept in log2 units (bias, logit scale and the mask value carry log2(e)): the softmax is one EX2 per element
        float4 b = __ldg(reinterpret_cast<const float4*>(bias + (size_t(head) * WT + row) * WT) + c4);
        b.x *= LOG2E; b.y *= LOG2E; b.z *= LOG2E; b.w *= LOG2E;
        *reinterpret_cast<float4*>(Bs + row * BS_LD + c4 * 4) = b;
    }
    const float sc = __ldg(scale + head) * LOG2E;
    const int wpr = W / WS, wpi = (H / WS) * wpr;
    const int R0 = warp * 16 + g, R1 = R0 + 8;             // in-window rows (2 * warp, g) and (2 * warp + 1, g)

    WinLoad nxt;
    int win = blockIdx.x;
    if (win < n_windows) window_fetch(nxt, qkv, win, head, C, H, W, wpr, wpi, shift, tid, R0, t);
    for (; win < n_windows; win += gridDim.x) {
        const WinLoad cur = nxt;
        __syncthreads();                       // the previous window's shared-memory reads are done (first pass: bias staged)
        // ---- K, V of the window -> shared memory (K row-major, V transposed); thread = (row, 16-dim half) ----
        {
            const int r = tid >> 1, half = tid & 1;
            uint4* kd = reinterpret_cast<uint4*>(Ks + r * KS_LD + half * 16);
            kd[0] = cur.k0;
            kd[1] = cur.k1;
            float n = sumsq_bf2(cur.k0.x) + sumsq_bf2(cur.k0.y) + sumsq_bf2(cur.k0.z) + sumsq_bf2(cur.k0.w) + sumsq_bf2(cur.k1.x) +
                      sumsq_bf2(cur.k1.y) + sumsq_bf2(cur.k1.z) + sumsq_bf2(cur.k1.w);
            n += __shfl_xor_sync(0xffffffffu, n, 1);
            if (half == 0) rk[r] = 1.0f / fmaxf(sqrtf(n), 1e-12f);
#if RGBNM_WATTN_LDMATRIX
            uint4* vd = reinterpret_cast<uint4*>(Vs + r * KS_LD + half * 16);
            vd[0] = cur.v0;
            vd[1] = cur.v1;
#else
            const unsigned vw[8] = {cur.v0.x, cur.v0.y, cur.v0.z, cur.v0.w, cur.v1.x, cur.v1.y, cur.v1.z, cur.v1.w};
            unsigned short* vt = reinterpret_cast<unsigned short*>(Vt);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                vt[(half * 16 + 2 * e) * VT_LD + r] = static_cast<unsigned short>(vw[e] & 0xffffu);
                vt[(half * 16 + 2 * e + 1) * VT_LD + r] = static_cast<unsigned short>(vw[e] >> 16);
            }
#endif
        }
        // the next window's operands start their trip now and land during this window's arithmetic
        if (win + int(gridDim.x) < n_windows) window_fetch(nxt, qkv, win + gridDim.x, head, C, H, W, wpr, wpi, shift, tid, R0, t);
        const float n0 = quad_sum(sumsq_bf2(cur.qa[0][0]) + sumsq_bf2(cur.qa[0][2]) + sumsq_bf2(cur.qa[1][0]) + sumsq_bf2(cur.qa[1][2]));
        const float n1 = quad_sum(sumsq_bf2(cur.qa[0][1]) + sumsq_bf2(cur.qa[0][3]) + sumsq_bf2(cur.qa[1][1]) + sumsq_bf2(cur.qa[1][3]));
        const float f0 = sc / fmaxf(sqrtf(n0), 1e-12f), f1 = sc / fmaxf(sqrtf(n1), 1e-12f);
        __syncthreads();

        // ---- S = Q K^T (raw), 8 key tiles x 2 k-steps ----
        float s[8][4];
#if RGBNM_WATTN_LDMATRIX
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
            // matrices m = 0..3: keys nt*8 .. +7 (rows), dims 8m .. 8m+7 -> {b0, b1} of k-step 0 and of k-step 1
            unsigned kb[4];
            ldmatrix_x4(kb, Ks + (nt * 8 + (lane & 7)) * KS_LD + (lane >> 3) * 8);
            mma16816(s[nt], cur.qa[0], kb[0], kb[1]);
            mma16816(s[nt], cur.qa[1], kb[2], kb[3]);
        }
#else
        const unsigned* ksw = reinterpret_cast<const unsigned*>(Ks);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int w = (nt * 8 + g) * (KS_LD / 2) + ks * 8 + t;
                mma16816(s[nt], cur.qa[ks], ksw[w], ksw[w + 4]);
            }
        }
#endif
        // ---- cosine normalisation, logit scale, relative position bias, shift mask; row maxima ----
        // shift regions (img_mask of swinv2.py:227-238, in shifted coordinates): only the last window row / column is
        // split, at in-window coordinate 8 - shift: region = 3 * hr + wr with hr, wr in {0 | 1, 2}
        const int wrem = win % wpi;
        const bool last_y = shift > 0 && (wrem / wpr) == (H / WS) - 1, last_x = shift > 0 && (wrem % wpr) == wpr - 1;
        const int cut = WS - shift;
        const int hr0 = last_y ? (2 * warp < cut ? 1 : 2) : 0, hr1 = last_y ? (2 * warp + 1 < cut ? 1 : 2) : 0;
        const int wrr = last_x ? (g < cut ? 1 : 2) : 0;                          // both rows of this thread sit at column g
        const int wc0 = last_x ? (2 * t < cut ? 1 : 2) : 0, wc1 = last_x ? (2 * t + 1 < cut ? 1 : 2) : 0;
        float m0 = -3.0e38f, m1 = -3.0e38f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int col = nt * 8 + 2 * t;                                      // key at in-window (nt, 2t) and (nt, 2t + 1)
            const float rk0 = rk[col], rk1 = rk[col + 1];
            const float2 b0 = *reinterpret_cast<const float2*>(Bs + R0 * BS_LD + col);
            const float2 b1 = *reinterpret_cast<const float2*>(Bs + R1 * BS_LD + col);
            s[nt][0] = fmaf(s[nt][0] * rk0, f0, b0.x);
            s[nt][1] = fmaf(s[nt][1] * rk1, f0, b0.y);
            s[nt][2] = fmaf(s[nt][2] * rk0, f1, b1.x);
            s[nt][3] = fmaf(s[nt][3] * rk1, f1, b1.y);
            if (last_y || last_x) {
                const int hc = last_y ? (nt < cut ? 1 : 2) : 0;
                if (hc != hr0 || wc0 != wrr) s[nt][0] += -100.0f * LOG2E;        // attn_mask value of swinv2.py:242
                if (hc != hr0 || wc1 != wrr) s[nt][1] += -100.0f * LOG2E;
                if (hc != hr1 || wc0 != wrr) s[nt][2] += -100.0f * LOG2E;
                if (hc != hr1 || wc1 != wrr) s[nt][3] += -100.0f * LOG2E;
            }
            m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
            m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
        }
        m0 = quad_max(m0);
        m1 = quad_max(m1);
        float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = ex2_fast(s[nt][0] - m0);
            s[nt][1] = ex2_fast(s[nt][1] - m0);
            s[nt][2] = ex2_fast(s[nt][2] - m1);
            s[nt][3] = ex2_fast(s[nt][3] - m1);
            l0 += s[nt][0] + s[nt][1];
            l1 += s[nt][2] + s[nt][3];
        }
        l0 = quad_sum(l0);
        l1 = quad_sum(l1);
        // ---- O = P V: 4 key steps x 4 dim tiles; P from the accumulator registers ----
        float o[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.0f;
#if !RGBNM_WATTN_LDMATRIX
        const unsigned* vtw = reinterpret_cast<const unsigned*>(Vt);
#endif
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const unsigned pa[4] = {f2_to_bf2(s[2 * kk][0], s[2 * kk][1]), f2_to_bf2(s[2 * kk][2], s[2 * kk][3]),
                                    f2_to_bf2(s[2 * kk + 1][0], s[2 * kk + 1][1]), f2_to_bf2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#if RGBNM_WATTN_LDMATRIX
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                // matrices m = 0..3 (transposed on load): keys kk*16 + 8*(m & 1) .. +7 (rows), dims (2*np + (m >> 1))*8 .. +7
                // -> {b0, b1} of dim tile 2*np and of dim tile 2*np + 1
                unsigned vb[4];
                const int m = lane >> 3;
                ldmatrix_x4_trans(vb, Vs + (kk * 16 + (m & 1) * 8 + (lane & 7)) * KS_LD + (2 * np + (m >> 1)) * 8);
                mma16816(o[2 * np], pa, vb[0], vb[1]);
                mma16816(o[2 * np + 1], pa, vb[2], vb[3]);
            }
#else
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int w = (nt * 8 + g) * (VT_LD / 2) + kk * 8 + t;
                mma16816(o[nt], pa, vtw[w], vtw[w + 4]);
            }
#endif
        }
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        unsigned* d0 = reinterpret_cast<unsigned*>(out + size_t(cur.tok0) * C + head * HD);
        unsigned* d1 = reinterpret_cast<unsigned*>(out + size_t(cur.tok1) * C + head * HD);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            d0[nt * 4 + t] = f2_to_bf2(o[nt][0] * i0, o[nt][1] * i0);
            d1[nt * 4 + t] = f2_to_bf2(o[nt][2] * i1, o[nt][3] * i1);
        }
    }
}

// out[b][h2][w2][q * C + c] = x[b][2 * h2 + (q & 1)][2 * w2 + (q >> 1)][c]      (x0 | x1 | x2 | x3, swinv2.py:353-358)
__global__ void __launch_bounds__(256)
patch_merge_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int H, int W, int C8, size_t total) {
    for (size_t e = blockIdx.x * size_t(blockDim.x) + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
        const int c = int(e % C8);
        size_t r = e / C8;
        const int q = int(r & 3);
        r >>= 2;
        const int W2 = W >> 1, H2 = H >> 1;
        const int w2 = int(r % W2);
        r /= W2;
        const int h2 = int(r % H2);
        const size_t b = r / H2;
        out[e] = __ldg(x + ((b * H + 2 * h2 + (q & 1)) * W + 2 * w2 + (q >> 1)) * C8 + c);
    }
}

// out[b][c] = mean_l x[b][l][c]; one CTA per (image, 64-column slab), 4 token groups
__global__ void __launch_bounds__(256)
token_mean_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int L, int C) {
    __shared__ float part[4][64];
    const int b = blockIdx.x, col = blockIdx.y * 64 + (threadIdx.x & 63), grp = threadIdx.x >> 6;
    float acc = 0.0f;
    if (col < C)
        for (int l = grp; l < L; l += 4) acc += __bfloat162float(x[(size_t(b) * L + l) * C + col]);
    part[grp][threadIdx.x & 63] = acc;
    __syncthreads();
    if (grp == 0 && col < C) {
        const int i = threadIdx.x & 63;
        out[size_t(b) * C + col] = __float2bfloat16_rn((part[0][i] + part[1][i] + part[2][i] + part[3][i]) / float(L));
    }
}

}  // namespace swink

static int rgbnm_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    }
    return sms;
}

extern "C" int rgbnm_layernorm_res_fwd(const void* x, const float* gamma, const float* beta, const void* res, void* y,
                                       int rows, int emb, float eps, void* stream) {
    using namespace swink;
    if (!x || !gamma || !beta || !y || rows < 0 || emb <= 0 || (emb & 1) || emb > 1536) return RGBNM_ERR_ARG;
    if (rows == 0) return RGBNM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int grid = (rows + LN_WARPS - 1) / LN_WARPS;
    const int cap = rgbnm_num_sms() * 8;
    if (grid > cap) grid = cap;
    const unsigned* xx = static_cast<const unsigned*>(x);
    const unsigned* rr = static_cast<const unsigned*>(res);
    unsigned* yy = static_cast<unsigned*>(y);
    if (emb <= 128) ln_res_fwd_kernel<2><<<grid, LN_WARPS * 32, 0, st>>>(xx, gamma, beta, rr, yy, rows, emb, eps);
    else if (emb <= 192) ln_res_fwd_kernel<3><<<grid, LN_WARPS * 32, 0, st>>>(xx, gamma, beta, rr, yy, rows, emb, eps);
    else if (emb <= 384) ln_res_fwd_kernel<6><<<grid, LN_WARPS * 32, 0, st>>>(xx, gamma, beta, rr, yy, rows, emb, eps);
    else if (emb <= 768) ln_res_fwd_kernel<12><<<grid, LN_WARPS * 32, 0, st>>>(xx, gamma, beta, rr, yy, rows, emb, eps);
    else ln_res_fwd_kernel<24><<<grid, LN_WARPS * 32, 0, st>>>(xx, gamma, beta, rr, yy, rows, emb, eps);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

extern "C" int rgbnm_window_attention_fwd(const void* qkv, void* out, const float* bias, const float* scale, int B, int H, int W,
                                          int C, int heads, int window, int shift, void* stream) {
    using namespace swink;
    if (!qkv || !out || !bias || !scale || B < 0) return RGBNM_ERR_ARG;
    // round 1: the SwinV2-T configuration (utils/configs.py:123-137): 8 x 8 windows, head dimension 32
    if (window != WS || heads <= 0 || C != heads * HD) return RGBNM_ERR_UNSUPPORTED;
    if (H <= 0 || W <= 0 || (H % WS) || (W % WS) || shift < 0 || shift >= WS) return RGBNM_ERR_ARG;
    if (B == 0) return RGBNM_OK;
    const int n_windows = B * (H / WS) * (W / WS);
    static const bool simt = (getenv("RGBNM_WATTN_SIMT") != nullptr);     // the CUDA-core kernel, kept for A/B runs
    if (simt) {
        const dim3 grid(n_windows, heads);
        window_attn_fwd_kernel<<<grid, WT, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), bias, scale, H, W, C, shift);
    } else {
        // a CTA keeps its head's bias tile in shared memory and walks windows: exactly one resident wave of CTAs
        static int occ = 0;
        if (occ == 0) {
            RGBNM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, window_attn_mma_kernel, 128, 0));
            if (occ < 1) occ = 1;
        }
        int ctas = rgbnm_num_sms() * occ / heads;
        if (ctas < 1) ctas = 1;
        if (ctas > n_windows) ctas = n_windows;
        const dim3 grid(ctas, heads);
        window_attn_mma_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), bias, scale, H, W, C, shift, n_windows);
    }
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

extern "C" int rgbnm_patch_merge_gather(const void* x, void* out, int B, int H, int W, int C, void* stream) {
    using namespace swink;
    if (!x || !out || B < 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || C <= 0 || (C % 8)) return RGBNM_ERR_ARG;
    if (B == 0) return RGBNM_OK;
    const size_t total = size_t(B) * (H / 2) * (W / 2) * 4 * (C / 8);
    size_t blocks = (total + 255) / 256;
    const size_t cap = size_t(rgbnm_num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    patch_merge_kernel<<<unsigned(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(x), static_cast<uint4*>(out), H, W, C / 8, total);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

extern "C" int rgbnm_token_mean_bf16(const void* x, void* out, int B, int L, int C, void* stream) {
    using namespace swink;
    if (!x || !out || B < 0 || L <= 0 || C <= 0) return RGBNM_ERR_ARG;
    if (B == 0) return RGBNM_OK;
    token_mean_kernel<<<dim3(B, (C + 63) / 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), L, C);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}
