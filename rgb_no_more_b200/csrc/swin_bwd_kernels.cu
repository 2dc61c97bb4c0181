// Backward kernels of the SwinV2 DCT path (training): first correct CUDA versions, fp32 on the CUDA
// cores -- the tensor-core forms come after parity.
//   post-norm residual with stochastic depth   y = res + s_b * LayerNorm(x)          models/swinv2.py:302-306 (drop_path)
//       forward with the per-image scale s_b, backward dx / dgamma / dbeta (dres = dy needs no kernel)
//   window attention backward                  dqkv, d(bias tile), d(logit scale)    models/swinv2.py:152-177
//   patch-merging scatter                      inverse of the forward gather         models/swinv2.py:353-358
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "../../include/rgbnm_b200.h"
#include "common.cuh"

namespace swinb {

constexpr int LN_WARPS = 8;
constexpr int WS = 8, WT = 64, HD = 32;

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

// y = (res ? res : 0) + s * ((x - mean) * rstd * gamma + beta),  s = row_scale ? row_scale[row / rows_per_scale] : 1
template <int MAXP>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_res_scaled_fwd_kernel(const unsigned* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                         const unsigned* __restrict__ res, const float* __restrict__ row_scale, int rows_per_scale,
                         unsigned* __restrict__ y, int rows, int E, float eps) {
    const int lane = threadIdx.x & 31, pairs = E >> 1;
    const float inv_e = 1.0f / float(E);
    for (int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5); row < rows; row += gridDim.x * LN_WARPS) {
        float2 v[MAXP];
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
            const int p = k * 32 + lane;
            v[k] = p < pairs ? bf2_to_f2(__ldg(x + size_t(row) * pairs + p)) : make_float2(0.0f, 0.0f);
            s += v[k].x + v[k].y;
        }
        const float mean = warp_sum(s) * inv_e;
        float q = 0.0f;
#pragma unroll
        for (int k = 0; k < MAXP; ++k)
            if (k * 32 + lane < pairs) { const float a = v[k].x - mean, b = v[k].y - mean; q += a * a + b * b; }
        const float rstd = rsqrtf(warp_sum(q) * inv_e + eps);
        const float sc = row_scale != nullptr ? __ldg(row_scale + row / rows_per_scale) : 1.0f;
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
            const int p = k * 32 + lane;
            if (p < pairs) {
                const float2 g = *reinterpret_cast<const float2*>(gamma + 2 * p), b = *reinterpret_cast<const float2*>(beta + 2 * p);
                float o0 = sc * ((v[k].x - mean) * rstd * g.x + b.x), o1 = sc * ((v[k].y - mean) * rstd * g.y + b.y);
                if (res != nullptr) { const float2 r = bf2_to_f2(__ldg(res + size_t(row) * pairs + p)); o0 += r.x; o1 += r.y; }
                y[size_t(row) * pairs + p] = f2_to_bf2(o0, o1);
            }
        }
    }
}

// Backward of the LayerNorm branch: g = s * dy;  dbeta += g;  dgamma += g * xhat;  dx = rstd * (g*gamma - mean(g*gamma) -
// xhat * mean(g*gamma*xhat)).  Statistics are recomputed from x.  dgamma / dbeta: per-warp registers -> shared -> one atomic
// per column and CTA.
template <int MAXP>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_res_bwd_kernel(const unsigned* __restrict__ dy, const unsigned* __restrict__ x, const float* __restrict__ gamma,
                  const float* __restrict__ row_scale, int rows_per_scale, unsigned* __restrict__ dx, float* __restrict__ dgamma,
                  float* __restrict__ dbeta, int rows, int E, float eps) {
    __shared__ float red[LN_WARPS][64 * MAXP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pairs = E >> 1;
    const float inv_e = 1.0f / float(E);
    float2 ag[MAXP], ab[MAXP];
#pragma unroll
    for (int k = 0; k < MAXP; ++k) ag[k] = ab[k] = make_float2(0.0f, 0.0f);
    for (int row = blockIdx.x * LN_WARPS + warp; row < rows; row += gridDim.x * LN_WARPS) {
        float2 v[MAXP], g[MAXP];
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
            const int p = k * 32 + lane;
            const bool in = p < pairs;
            v[k] = in ? bf2_to_f2(__ldg(x + size_t(row) * pairs + p)) : make_float2(0.0f, 0.0f);
            g[k] = in ? bf2_to_f2(__ldg(dy + size_t(row) * pairs + p)) : make_float2(0.0f, 0.0f);
            s += v[k].x + v[k].y;
        }
        const float mean = warp_sum(s) * inv_e;
        float q = 0.0f;
#pragma unroll
        for (int k = 0; k < MAXP; ++k)
            if (k * 32 + lane < pairs) { const float a = v[k].x - mean, b = v[k].y - mean; q += a * a + b * b; }
        const float rstd = rsqrtf(warp_sum(q) * inv_e + eps);
        const float sc = row_scale != nullptr ? __ldg(row_scale + row / rows_per_scale) : 1.0f;
        float m1 = 0.0f, m2 = 0.0f;
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
            const int p = k * 32 + lane;
            if (p < pairs) {
                const float2 gm = *reinterpret_cast<const float2*>(gamma + 2 * p);
                v[k].x = (v[k].x - mean) * rstd;                     // xhat
                v[k].y = (v[k].y - mean) * rstd;
                g[k].x *= sc;
                g[k].y *= sc;
                ab[k].x += g[k].x;
                ab[k].y += g[k].y;
                ag[k].x += g[k].x * v[k].x;
                ag[k].y += g[k].y * v[k].y;
                g[k].x *= gm.x;                                      // dxhat
                g[k].y *= gm.y;
                m1 += g[k].x + g[k].y;
                m2 += g[k].x * v[k].x + g[k].y * v[k].y;
            }
        }
        m1 = warp_sum(m1) * inv_e;
        m2 = warp_sum(m2) * inv_e;
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
            const int p = k * 32 + lane;
            if (p < pairs)
                dx[size_t(row) * pairs + p] = f2_to_bf2(rstd * (g[k].x - m1 - v[k].x * m2), rstd * (g[k].y - m1 - v[k].y * m2));
        }
    }
    // CTA reduction of the parameter gradients
    for (int pass = 0; pass < 2; ++pass) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
            const float2 a = pass == 0 ? ag[k] : ab[k];
            red[warp][2 * (k * 32 + lane)] = a.x;
            red[warp][2 * (k * 32 + lane) + 1] = a.y;
        }
        __syncthreads();
        float* out = pass == 0 ? dgamma : dbeta;
        for (int c = threadIdx.x; c < E; c += LN_WARPS * 32) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < LN_WARPS; ++w) t += red[w][c];
            atomicAdd(out + c, t);
        }
    }
}

// ------------------------------------------------------------------------------------------
// The same two kernels with G lanes per row (E = 2 * G * PPL exactly: 96 -> 8 lanes x 6 pairs, 192 -> 16 x 6, 384 -> 32 x 6,
// 768 -> 32 x 12).  At 96 / 192 channels a warp works on 4 / 2 rows at once: no idle lanes (a 96-wide row fills 48 of a warp's 64
// pair slots otherwise), reductions of 3 / 4 shuffle steps shared by the rows, gamma / beta held in registers.
// ------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int G, int PPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_res_scaled_fwd_g_kernel(const unsigned* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                           const unsigned* __restrict__ res, const float* __restrict__ row_scale, int rows_per_scale,
                           unsigned* __restrict__ y, int rows, float eps) {
    constexpr int RPW = 32 / G, PAIRS = G * PPL;
    const int lane = threadIdx.x & 31, grp = lane / G, j = lane % G;
    const float inv_e = 1.0f / float(2 * PAIRS);
    float2 gm[PPL], bt[PPL];
#pragma unroll
    for (int k = 0; k < PPL; ++k) {
        gm[k] = *reinterpret_cast<const float2*>(gamma + 2 * (k * G + j));
        bt[k] = *reinterpret_cast<const float2*>(beta + 2 * (k * G + j));
    }
    for (int base = (blockIdx.x * LN_WARPS + (threadIdx.x >> 5)) * RPW; base < rows; base += gridDim.x * LN_WARPS * RPW) {
        const int row = base + grp;
        const bool live = row < rows;
        const unsigned* xr = x + size_t(live ? row : 0) * PAIRS;
        float2 v[PPL];
        unsigned rr[PPL];
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            v[k] = bf2_to_f2(__ldg(xr + k * G + j));
            rr[k] = (res != nullptr) ? __ldg(res + size_t(live ? row : 0) * PAIRS + k * G + j) : 0u;
            s += v[k].x + v[k].y;
        }
        const float mean = group_sum<G>(s) * inv_e;
        float q = 0.0f;
#pragma unroll
        for (int k = 0; k < PPL; ++k) { const float a = v[k].x - mean, b = v[k].y - mean; q += a * a + b * b; }
        const float rstd = rsqrtf(group_sum<G>(q) * inv_e + eps);
        const float sc = (row_scale != nullptr && live) ? __ldg(row_scale + row / rows_per_scale) : 1.0f;
        if (live) {
#pragma unroll
            for (int k = 0; k < PPL; ++k) {
                const float2 r = bf2_to_f2(rr[k]);
                const float o0 = sc * ((v[k].x - mean) * rstd * gm[k].x + bt[k].x) + r.x;
                const float o1 = sc * ((v[k].y - mean) * rstd * gm[k].y + bt[k].y) + r.y;
                y[size_t(row) * PAIRS + k * G + j] = f2_to_bf2(o0, o1);
            }
        }
    }
}

template <int G, int PPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_res_bwd_g_kernel(const unsigned* __restrict__ dy, const unsigned* __restrict__ x, const float* __restrict__ gamma,
                    const float* __restrict__ row_scale, int rows_per_scale, unsigned* __restrict__ dx, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, float* __restrict__ dxsum, int rows, float eps) {
    constexpr int RPW = 32 / G, PAIRS = G * PPL, E = 2 * PAIRS;
    __shared__ float red[LN_WARPS * RPW][E];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / G, j = lane % G;
    const float inv_e = 1.0f / float(E);
    // dxsum (optional): column sums of the dx written here = the bias gradient of the Linear whose output this LayerNorm takes
    // (fc2 / proj of a block: no separate column-sum launch); summed from the bf16-rounded values, as the column-sum kernel would
    float2 gm[PPL], ag[PPL], ab[PPL], ax[PPL];
#pragma unroll
    for (int k = 0; k < PPL; ++k) {
        gm[k] = *reinterpret_cast<const float2*>(gamma + 2 * (k * G + j));
        ag[k] = ab[k] = ax[k] = make_float2(0.0f, 0.0f);
    }
    for (int base = (blockIdx.x * LN_WARPS + warp) * RPW; base < rows; base += gridDim.x * LN_WARPS * RPW) {
        const int row = base + grp;
        const bool live = row < rows;
        const size_t ro = size_t(live ? row : 0) * PAIRS;
        float2 v[PPL], g[PPL];
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            v[k] = bf2_to_f2(__ldg(x + ro + k * G + j));
            g[k] = live ? bf2_to_f2(__ldg(dy + ro + k * G + j)) : make_float2(0.0f, 0.0f);
            s += v[k].x + v[k].y;
        }
        const float mean = group_sum<G>(s) * inv_e;
        float q = 0.0f;
#pragma unroll
        for (int k = 0; k < PPL; ++k) { const float a = v[k].x - mean, b = v[k].y - mean; q += a * a + b * b; }
        const float rstd = rsqrtf(group_sum<G>(q) * inv_e + eps);
        const float sc = (row_scale != nullptr && live) ? __ldg(row_scale + row / rows_per_scale) : 1.0f;
        float m1 = 0.0f, m2 = 0.0f;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            v[k].x = (v[k].x - mean) * rstd;                         // xhat
            v[k].y = (v[k].y - mean) * rstd;
            g[k].x *= sc;
            g[k].y *= sc;
            ab[k].x += g[k].x;
            ab[k].y += g[k].y;
            ag[k].x += g[k].x * v[k].x;
            ag[k].y += g[k].y * v[k].y;
            g[k].x *= gm[k].x;                                       // dxhat
            g[k].y *= gm[k].y;
            m1 += g[k].x + g[k].y;
            m2 += g[k].x * v[k].x + g[k].y * v[k].y;
        }
        m1 = group_sum<G>(m1) * inv_e;
        m2 = group_sum<G>(m2) * inv_e;
        if (live) {
#pragma unroll
            for (int k = 0; k < PPL; ++k) {
                const unsigned w = f2_to_bf2(rstd * (g[k].x - m1 - v[k].x * m2), rstd * (g[k].y - m1 - v[k].y * m2));
                dx[ro + k * G + j] = w;
                const float2 r = bf2_to_f2(w);
                ax[k].x += r.x;
                ax[k].y += r.y;
            }
        }
    }
    for (int pass = 0; pass < (dxsum != nullptr ? 3 : 2); ++pass) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            const float2 a = pass == 0 ? ag[k] : (pass == 1 ? ab[k] : ax[k]);
            red[warp * RPW + grp][2 * (k * G + j)] = a.x;
            red[warp * RPW + grp][2 * (k * G + j) + 1] = a.y;
        }
        __syncthreads();
        float* out = pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum);
        for (int c = threadIdx.x; c < E; c += LN_WARPS * 32) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < LN_WARPS * RPW; ++w) t += red[w][c];
            atomicAdd(out + c, t);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Window attention backward, CUDA cores.  CTA = 64 threads, walks the windows of one head (blockIdx.y) so that the head's
// bias-gradient tile accumulates in shared memory and is flushed once.  Row pass (thread = query i): s, P, dP = dO . V^T,
// delta, dS -> shared; dq.  Column pass (thread = key j): dV = P^T dO, dk.  Cosine normalisation differentiated explicitly:
// d(x / |x|) -> (g - xn (xn . g)) / |x|.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int window_token(int win, int j, int H, int W, int wpr, int wpi, int shift) {
    const int img = win / wpi, wrem = win - img * wpi;
    const int wy = wrem / wpr, wx = wrem - wy * wpr;
    int py = wy * WS + (j >> 3) + shift, px = wx * WS + (j & 7) + shift;
    if (py >= H) py -= H;
    if (px >= W) px -= W;
    return (img * H + py) * W + px;
}

struct AttnBwdSmem {
    float Qn[WT][HD + 1], Kn[WT][HD + 1], V[WT][HD + 1], dO[WT][HD + 1];
    float P[WT][WT + 1], dS[WT][WT + 1];
    float dB[WT][WT + 1];
    float qinv[WT], kinv[WT];
    float red[2];
    int region[WT];
};

__global__ void __launch_bounds__(WT)
window_attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout, const float* __restrict__ bias,
                       const float* __restrict__ scale, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dbias,
                       float* __restrict__ dscale, int H, int W, int C, int shift, int n_windows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AttnBwdSmem& sm = *reinterpret_cast<AttnBwdSmem*>(smem_raw);
    const int t = threadIdx.x, head = blockIdx.y;
    const int wpr = W / WS, wpi = (H / WS) * wpr;
    const float sc = __ldg(scale + head);
    for (int j = 0; j < WT; ++j) sm.dB[t][j] = 0.0f;
    float dscale_acc = 0.0f;

    for (int win = blockIdx.x; win < n_windows; win += gridDim.x) {
        __syncthreads();
        const int tok = window_token(win, t, H, W, wpr, wpi, shift);
        {
            const int wrem = win % wpi;
            const int sy = (wrem / wpr) * WS + (t >> 3), sx = (wrem % wpr) * WS + (t & 7);
            int reg = 0;
            if (shift > 0) {
                const int hr = sy < H - WS ? 0 : (sy < H - shift ? 1 : 2);
                const int wr = sx < W - WS ? 0 : (sx < W - shift ? 1 : 2);
                reg = 3 * hr + wr;
            }
            sm.region[t] = reg;
        }
        const __nv_bfloat16* row = qkv + size_t(tok) * (3 * size_t(C)) + head * HD;
        const __nv_bfloat16* drow = dout + size_t(tok) * C + head * HD;
        float q[HD], qn2 = 0.0f, kn2 = 0.0f;
        for (int d = 0; d < HD; ++d) {
            q[d] = __bfloat162float(row[d]);
            const float k = __bfloat162float(row[C + d]);
            sm.Kn[t][d] = k;
            sm.V[t][d] = __bfloat162float(row[2 * C + d]);
            sm.dO[t][d] = __bfloat162float(drow[d]);
            qn2 += q[d] * q[d];
            kn2 += k * k;
        }
        const float qi = 1.0f / fmaxf(sqrtf(qn2), 1e-12f), ki = 1.0f / fmaxf(sqrtf(kn2), 1e-12f);
        sm.qinv[t] = qi;
        sm.kinv[t] = ki;
        for (int d = 0; d < HD; ++d) {
            q[d] *= qi;
            sm.Qn[t][d] = q[d];
            sm.Kn[t][d] *= ki;
        }
        __syncthreads();
        // ---- row pass ----
        float s[WT], cosv[WT];
        float m = -3.0e38f;
        const int reg = sm.region[t];
        for (int j = 0; j < WT; ++j) {
            float acc = 0.0f;
            for (int d = 0; d < HD; ++d) acc = fmaf(q[d], sm.Kn[j][d], acc);
            cosv[j] = acc;
            float v = fmaf(acc, sc, __ldg(bias + (size_t(head) * WT + t) * WT + j));
            if (shift > 0 && sm.region[j] != reg) v += -100.0f;
            s[j] = v;
            m = fmaxf(m, v);
        }
        float l = 0.0f;
        for (int j = 0; j < WT; ++j) { s[j] = __expf(s[j] - m); l += s[j]; }
        const float il = 1.0f / l;
        float delta = 0.0f;
        float dp[WT];
        for (int j = 0; j < WT; ++j) {
            s[j] *= il;                                            // P_ij
            float acc = 0.0f;
            for (int d = 0; d < HD; ++d) acc = fmaf(sm.dO[t][d], sm.V[j][d], acc);
            dp[j] = acc;
            delta = fmaf(s[j], acc, delta);
        }
        float dqn[HD];
        for (int d = 0; d < HD; ++d) dqn[d] = 0.0f;
        for (int j = 0; j < WT; ++j) {
            const float ds = s[j] * (dp[j] - delta);
            sm.P[t][j] = s[j];
            sm.dS[t][j] = ds;
            sm.dB[t][j] += ds;                                     // row t of the tile belongs to this thread
            dscale_acc = fmaf(ds, cosv[j], dscale_acc);
            const float dc = ds * sc;
            for (int d = 0; d < HD; ++d) dqn[d] = fmaf(dc, sm.Kn[j][d], dqn[d]);
        }
        {
            float dot = 0.0f;
            for (int d = 0; d < HD; ++d) dot = fmaf(q[d], dqn[d], dot);
            __nv_bfloat16* o = dqkv + size_t(tok) * (3 * size_t(C)) + head * HD;
            for (int d = 0; d < HD; ++d) o[d] = __float2bfloat16_rn((dqn[d] - q[d] * dot) * qi);
        }
        __syncthreads();
        // ---- column pass: thread = key t ----
        float dv[HD], dkn[HD];
        for (int d = 0; d < HD; ++d) dv[d] = dkn[d] = 0.0f;
        for (int i = 0; i < WT; ++i) {
            const float p = sm.P[i][t], dc = sm.dS[i][t] * sc;
            for (int d = 0; d < HD; ++d) {
                dv[d] = fmaf(p, sm.dO[i][d], dv[d]);
                dkn[d] = fmaf(dc, sm.Qn[i][d], dkn[d]);
            }
        }
        {
            float dot = 0.0f;
            for (int d = 0; d < HD; ++d) dot = fmaf(sm.Kn[t][d], dkn[d], dot);
            __nv_bfloat16* o = dqkv + size_t(tok) * (3 * size_t(C)) + head * HD;
            for (int d = 0; d < HD; ++d) {
                o[C + d] = __float2bfloat16_rn((dkn[d] - sm.Kn[t][d] * dot) * ki);
                o[2 * C + d] = __float2bfloat16_rn(dv[d]);
            }
        }
    }
    __syncthreads();
    for (int j = 0; j < WT; ++j) atomicAdd(dbias + (size_t(head) * WT + t) * WT + j, sm.dB[t][j]);
    dscale_acc = warp_sum(dscale_acc);
    if ((t & 31) == 0) atomicAdd(dscale + head, dscale_acc);
}

// ------------------------------------------------------------------------------------------
// Window attention backward on warp-level tensor-core MMAs (mma.sync m16n8k16, bf16 x bf16 -> fp32), the backward twin of
// swin_kernels.cu::window_attn_mma_kernel.  CTA = 4 warps = one (window, head) per iteration, grid-strided over the windows
// of one head (blockIdx.y): the head's bias tile is staged once and d(bias tile) accumulates in REGISTERS (a thread owns
// the same 32 (row, column) positions of every window), flushed with one atomic per element at the end.
//   stage      q, k, v, dO (64 x 32 bf16 each) -> shared memory, 1 / |q_i|, 1 / |k_j|
//   row phase  warp w = query rows 16w .. 16w+15:  S = Q K^T and dP = dO V^T (A fragments by ldmatrix, B = K / V rows),
//              cos = S / (|q||k|), logits = scale * cos + bias (+ -100 across shift regions), P = softmax, delta = sum_j P dP,
//              dS = P (dP - delta) -> d(bias), d(scale) += dS cos;  dqn = (scale dS / |k_j|) K   (A from the accumulator
//              registers, B = K through ldmatrix.trans);  dq = (dqn - qn (qn . dqn)) / |q|
//              P and E = scale dS / |q_i| (bf16) -> shared memory
//   col phase  warp w = keys 16w .. 16w+15:  dV = P^T dO, dkn = E^T Q (A = ldmatrix.trans of P / E, B = ldmatrix.trans of
//              dO / Q);  dk = (dkn - kn (kn . dkn)) / |k|
// The cosine normalisation is applied to fp32 accumulators of the RAW bf16 operands (no extra rounding), as in the forward.
// ------------------------------------------------------------------------------------------
constexpr int QLD = 40;       // bf16 per staged q / k / v / dO row (32 + 8 pad: 80 B, conflict-free ldmatrix)
constexpr int PLD = 72;       // bf16 per row of P / E (64 + 8 pad)
constexpr int BLD = 72;       // floats per staged bias row
constexpr float LOG2E_B = 1.4426950408889634f;

struct AttnBwdMmaSmem {
    __nv_bfloat16 Q[WT * QLD], K[WT * QLD], V[WT * QLD], dO[WT * QLD];
    __nv_bfloat16 P[WT * PLD], E[WT * PLD];
    float Bs[WT * BLD];
    float rq[WT], rk[WT];
};

__device__ __forceinline__ void mma16816(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(unsigned (&r)[4], const void* p) {
    const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(unsigned (&r)[4], const void* p) {
    const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ float sumsq_bf2(unsigned w) {
    const float2 f = bf2_to_f2(w);
    return f.x * f.x + f.y * f.y;
}
__device__ __forceinline__ float sumsq_u4(const uint4& a) { return sumsq_bf2(a.x) + sumsq_bf2(a.y) + sumsq_bf2(a.z) + sumsq_bf2(a.w); }
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

#ifndef RGBNM_WATTN_BWD_MINBLOCKS
#define RGBNM_WATTN_BWD_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(128, RGBNM_WATTN_BWD_MINBLOCKS)
window_attn_bwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout, const float* __restrict__ bias,
                           const float* __restrict__ scale, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dbias,
                           float* __restrict__ dscale, int H, int W, int C, int shift, int n_windows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AttnBwdMmaSmem& sm = *reinterpret_cast<AttnBwdMmaSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int head = blockIdx.y;
    for (int e = tid; e < WT * (WT / 4); e += 128) {
        const int row = e >> 4, c4 = e & 15;
        float4 b = __ldg(reinterpret_cast<const float4*>(bias + (size_t(head) * WT + row) * WT) + c4);
        b.x *= LOG2E_B; b.y *= LOG2E_B; b.z *= LOG2E_B; b.w *= LOG2E_B;               // logits are kept in log2 units
        *reinterpret_cast<float4*>(sm.Bs + row * BLD + c4 * 4) = b;
    }
    const float sc = __ldg(scale + head);
    const int wpr = W / WS, wpi = (H / WS) * wpr;
    const int R0 = warp * 16 + g, R1 = R0 + 8;
    float dB[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) dB[nt][0] = dB[nt][1] = dB[nt][2] = dB[nt][3] = 0.0f;
    float dscale_acc = 0.0f;
    const int m4 = lane >> 3, l8 = lane & 7;

    for (int win = blockIdx.x; win < n_windows; win += gridDim.x) {
        __syncthreads();                                   // the previous window's shared-memory reads are done
        // ---- stage: thread = (row, 16-dim half) ----
        const int sr = tid >> 1, half = tid & 1;
        const int stok = window_token(win, sr, H, W, wpr, wpi, shift);
        {
            const __nv_bfloat16* row = qkv + size_t(stok) * (3 * size_t(C)) + head * HD + half * 16;
            const __nv_bfloat16* drow = dout + size_t(stok) * C + head * HD + half * 16;
            const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(row)), q1 = __ldg(reinterpret_cast<const uint4*>(row) + 1);
            const uint4 k0 = __ldg(reinterpret_cast<const uint4*>(row + C)), k1 = __ldg(reinterpret_cast<const uint4*>(row + C) + 1);
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(row + 2 * C)), v1 = __ldg(reinterpret_cast<const uint4*>(row + 2 * C) + 1);
            const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(drow)), d1 = __ldg(reinterpret_cast<const uint4*>(drow) + 1);
            const int o = sr * QLD + half * 16;
            reinterpret_cast<uint4*>(sm.Q + o)[0] = q0; reinterpret_cast<uint4*>(sm.Q + o)[1] = q1;
            reinterpret_cast<uint4*>(sm.K + o)[0] = k0; reinterpret_cast<uint4*>(sm.K + o)[1] = k1;
            reinterpret_cast<uint4*>(sm.V + o)[0] = v0; reinterpret_cast<uint4*>(sm.V + o)[1] = v1;
            reinterpret_cast<uint4*>(sm.dO + o)[0] = d0; reinterpret_cast<uint4*>(sm.dO + o)[1] = d1;
            float nq = sumsq_u4(q0) + sumsq_u4(q1), nk = sumsq_u4(k0) + sumsq_u4(k1);
            nq += __shfl_xor_sync(0xffffffffu, nq, 1);
            nk += __shfl_xor_sync(0xffffffffu, nk, 1);
            if (half == 0) {
                sm.rq[sr] = 1.0f / fmaxf(sqrtf(nq), 1e-12f);           // F.normalize: x / max(|x|, 1e-12)
                sm.rk[sr] = 1.0f / fmaxf(sqrtf(nk), 1e-12f);
            }
        }
        __syncthreads();

        // ---- row phase ----
        unsigned qa[2][4], da[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int off = (warp * 16 + l8 + (m4 & 1) * 8) * QLD + ks * 16 + (m4 >> 1) * 8;
            ldmatrix_x4(qa[ks], sm.Q + off);
            ldmatrix_x4(da[ks], sm.dO + off);
        }
        float cs[8][4], dp[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            cs[nt][0] = cs[nt][1] = cs[nt][2] = cs[nt][3] = 0.0f;
            dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.0f;
            unsigned kb[4], vb[4];
            ldmatrix_x4(kb, sm.K + (nt * 8 + l8) * QLD + m4 * 8);
            ldmatrix_x4(vb, sm.V + (nt * 8 + l8) * QLD + m4 * 8);
            mma16816(cs[nt], qa[0], kb[0], kb[1]);
            mma16816(cs[nt], qa[1], kb[2], kb[3]);
            mma16816(dp[nt], da[0], vb[0], vb[1]);
            mma16816(dp[nt], da[1], vb[2], vb[3]);
        }
        const float rq0 = sm.rq[R0], rq1 = sm.rq[R1];
        const int wrem = win % wpi;
        const bool last_y = shift > 0 && (wrem / wpr) == (H / WS) - 1, last_x = shift > 0 && (wrem % wpr) == wpr - 1;
        const int cut = WS - shift;
        const int hr0 = last_y ? (2 * warp < cut ? 1 : 2) : 0, hr1 = last_y ? (2 * warp + 1 < cut ? 1 : 2) : 0;
        const int wrr = last_x ? (g < cut ? 1 : 2) : 0;
        const int wc0 = last_x ? (2 * t < cut ? 1 : 2) : 0, wc1 = last_x ? (2 * t + 1 < cut ? 1 : 2) : 0;
        float pr[8][4];
        float m0 = -3.0e38f, m1 = -3.0e38f;
        const float scl = sc * LOG2E_B;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int col = nt * 8 + 2 * t;
            const float rk0 = sm.rk[col], rk1 = sm.rk[col + 1];
            const float2 b0 = *reinterpret_cast<const float2*>(sm.Bs + R0 * BLD + col);
            const float2 b1 = *reinterpret_cast<const float2*>(sm.Bs + R1 * BLD + col);
            cs[nt][0] *= rq0 * rk0; cs[nt][1] *= rq0 * rk1; cs[nt][2] *= rq1 * rk0; cs[nt][3] *= rq1 * rk1;       // cosines
            pr[nt][0] = fmaf(cs[nt][0], scl, b0.x);
            pr[nt][1] = fmaf(cs[nt][1], scl, b0.y);
            pr[nt][2] = fmaf(cs[nt][2], scl, b1.x);
            pr[nt][3] = fmaf(cs[nt][3], scl, b1.y);
            if (last_y || last_x) {
                const int hc = last_y ? (nt < cut ? 1 : 2) : 0;
                if (hc != hr0 || wc0 != wrr) pr[nt][0] += -100.0f * LOG2E_B;
                if (hc != hr0 || wc1 != wrr) pr[nt][1] += -100.0f * LOG2E_B;
                if (hc != hr1 || wc0 != wrr) pr[nt][2] += -100.0f * LOG2E_B;
                if (hc != hr1 || wc1 != wrr) pr[nt][3] += -100.0f * LOG2E_B;
            }
            m0 = fmaxf(m0, fmaxf(pr[nt][0], pr[nt][1]));
            m1 = fmaxf(m1, fmaxf(pr[nt][2], pr[nt][3]));
        }
        m0 = quad_max(m0);
        m1 = quad_max(m1);
        float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            pr[nt][0] = ex2_fast(pr[nt][0] - m0);
            pr[nt][1] = ex2_fast(pr[nt][1] - m0);
            pr[nt][2] = ex2_fast(pr[nt][2] - m1);
            pr[nt][3] = ex2_fast(pr[nt][3] - m1);
            l0 += pr[nt][0] + pr[nt][1];
            l1 += pr[nt][2] + pr[nt][3];
        }
        const float i0 = 1.0f / quad_sum(l0), i1 = 1.0f / quad_sum(l1);
        float de0 = 0.0f, de1 = 0.0f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            pr[nt][0] *= i0; pr[nt][1] *= i0; pr[nt][2] *= i1; pr[nt][3] *= i1;
            de0 = fmaf(pr[nt][0], dp[nt][0], fmaf(pr[nt][1], dp[nt][1], de0));
            de1 = fmaf(pr[nt][2], dp[nt][2], fmaf(pr[nt][3], dp[nt][3], de1));
        }
        de0 = quad_sum(de0);
        de1 = quad_sum(de1);
        unsigned* Pw = reinterpret_cast<unsigned*>(sm.P);
        unsigned* Ew = reinterpret_cast<unsigned*>(sm.E);
        // dp becomes scale * dS / |k_j| (the A operand of dqn); P and E = scale * dS / |q_i| go to shared memory for the column phase
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int col = nt * 8 + 2 * t;
            const float rk0 = sm.rk[col], rk1 = sm.rk[col + 1];
            const float s0 = pr[nt][0] * (dp[nt][0] - de0), s1 = pr[nt][1] * (dp[nt][1] - de0);
            const float s2 = pr[nt][2] * (dp[nt][2] - de1), s3 = pr[nt][3] * (dp[nt][3] - de1);
            dB[nt][0] += s0; dB[nt][1] += s1; dB[nt][2] += s2; dB[nt][3] += s3;
            dscale_acc = fmaf(s0, cs[nt][0], fmaf(s1, cs[nt][1], fmaf(s2, cs[nt][2], fmaf(s3, cs[nt][3], dscale_acc))));
            Pw[(R0 * PLD + col) >> 1] = f2_to_bf2(pr[nt][0], pr[nt][1]);
            Pw[(R1 * PLD + col) >> 1] = f2_to_bf2(pr[nt][2], pr[nt][3]);
            Ew[(R0 * PLD + col) >> 1] = f2_to_bf2(s0 * sc * rq0, s1 * sc * rq0);
            Ew[(R1 * PLD + col) >> 1] = f2_to_bf2(s2 * sc * rq1, s3 * sc * rq1);
            dp[nt][0] = s0 * sc * rk0; dp[nt][1] = s1 * sc * rk1; dp[nt][2] = s2 * sc * rk0; dp[nt][3] = s3 * sc * rk1;
        }
        // dqn = (scale dS / |k|) K : 4 key steps x 4 dim tiles
        float acc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const unsigned pa[4] = {f2_to_bf2(dp[2 * kk][0], dp[2 * kk][1]), f2_to_bf2(dp[2 * kk][2], dp[2 * kk][3]),
                                    f2_to_bf2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]), f2_to_bf2(dp[2 * kk + 1][2], dp[2 * kk + 1][3])};
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                unsigned kb[4];
                ldmatrix_x4_trans(kb, sm.K + (kk * 16 + (m4 & 1) * 8 + l8) * QLD + (2 * np + (m4 >> 1)) * 8);
                mma16816(acc[2 * np], pa, kb[0], kb[1]);
                mma16816(acc[2 * np + 1], pa, kb[2], kb[3]);
            }
        }
        {
            // dq = (dqn - qn (qn . dqn)) / |q|; the accumulator positions (row, dims 8 nt + 2t, +1) are the A-fragment words of Q
            float dot0 = 0.0f, dot1 = 0.0f;
            float2 q0v[4], q1v[4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                q0v[nt] = bf2_to_f2(qa[nt >> 1][(nt & 1) * 2]);
                q1v[nt] = bf2_to_f2(qa[nt >> 1][(nt & 1) * 2 + 1]);
                q0v[nt].x *= rq0; q0v[nt].y *= rq0; q1v[nt].x *= rq1; q1v[nt].y *= rq1;
                dot0 = fmaf(q0v[nt].x, acc[nt][0], fmaf(q0v[nt].y, acc[nt][1], dot0));
                dot1 = fmaf(q1v[nt].x, acc[nt][2], fmaf(q1v[nt].y, acc[nt][3], dot1));
            }
            dot0 = quad_sum(dot0);
            dot1 = quad_sum(dot1);
            unsigned* o0 = reinterpret_cast<unsigned*>(dqkv + size_t(window_token(win, R0, H, W, wpr, wpi, shift)) * (3 * size_t(C)) + head * HD);
            unsigned* o1 = reinterpret_cast<unsigned*>(dqkv + size_t(window_token(win, R1, H, W, wpr, wpi, shift)) * (3 * size_t(C)) + head * HD);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                o0[nt * 4 + t] = f2_to_bf2((acc[nt][0] - q0v[nt].x * dot0) * rq0, (acc[nt][1] - q0v[nt].y * dot0) * rq0);
                o1[nt * 4 + t] = f2_to_bf2((acc[nt][2] - q1v[nt].x * dot1) * rq1, (acc[nt][3] - q1v[nt].y * dot1) * rq1);
            }
        }
        __syncthreads();                                   // P and E of all four warps are in shared memory

        // ---- column phase: warp w = keys 16w .. 16w+15 ----
        float dv[4][4], dk[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.0f;
            dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.0f;
        }
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) {
            // A = P^T / E^T: matrix m of the x4 load = queries 16 kq + 8 (m >> 1) .. +7 (stored rows), keys 16 w + 8 (m & 1) .. +7
            unsigned pa[4], ea[4];
            const int aoff = (kq * 16 + (m4 >> 1) * 8 + l8) * PLD + warp * 16 + (m4 & 1) * 8;
            ldmatrix_x4_trans(pa, sm.P + aoff);
            ldmatrix_x4_trans(ea, sm.E + aoff);
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                unsigned db[4], qb[4];
                const int boff = (kq * 16 + (m4 & 1) * 8 + l8) * QLD + (2 * np + (m4 >> 1)) * 8;
                ldmatrix_x4_trans(db, sm.dO + boff);
                ldmatrix_x4_trans(qb, sm.Q + boff);
                mma16816(dv[2 * np], pa, db[0], db[1]);
                mma16816(dv[2 * np + 1], pa, db[2], db[3]);
                mma16816(dk[2 * np], ea, qb[0], qb[1]);
                mma16816(dk[2 * np + 1], ea, qb[2], qb[3]);
            }
        }
        {
            const float rk0 = sm.rk[R0], rk1 = sm.rk[R1];
            const unsigned* Kw = reinterpret_cast<const unsigned*>(sm.K);
            float dot0 = 0.0f, dot1 = 0.0f;
            float2 k0v[4], k1v[4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                k0v[nt] = bf2_to_f2(Kw[(R0 * QLD + nt * 8 + 2 * t) >> 1]);
                k1v[nt] = bf2_to_f2(Kw[(R1 * QLD + nt * 8 + 2 * t) >> 1]);
                k0v[nt].x *= rk0; k0v[nt].y *= rk0; k1v[nt].x *= rk1; k1v[nt].y *= rk1;
                dot0 = fmaf(k0v[nt].x, dk[nt][0], fmaf(k0v[nt].y, dk[nt][1], dot0));
                dot1 = fmaf(k1v[nt].x, dk[nt][2], fmaf(k1v[nt].y, dk[nt][3], dot1));
            }
            dot0 = quad_sum(dot0);
            dot1 = quad_sum(dot1);
            unsigned* o0 = reinterpret_cast<unsigned*>(dqkv + size_t(window_token(win, R0, H, W, wpr, wpi, shift)) * (3 * size_t(C)) + head * HD);
            unsigned* o1 = reinterpret_cast<unsigned*>(dqkv + size_t(window_token(win, R1, H, W, wpr, wpi, shift)) * (3 * size_t(C)) + head * HD);
            const int ck = C >> 1;                           // 32-bit words per qkv third
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                o0[ck + nt * 4 + t] = f2_to_bf2((dk[nt][0] - k0v[nt].x * dot0) * rk0, (dk[nt][1] - k0v[nt].y * dot0) * rk0);
                o1[ck + nt * 4 + t] = f2_to_bf2((dk[nt][2] - k1v[nt].x * dot1) * rk1, (dk[nt][3] - k1v[nt].y * dot1) * rk1);
                o0[2 * ck + nt * 4 + t] = f2_to_bf2(dv[nt][0], dv[nt][1]);
                o1[2 * ck + nt * 4 + t] = f2_to_bf2(dv[nt][2], dv[nt][3]);
            }
        }
    }
    // ---- d(bias tile), d(scale) of this CTA's windows ----
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int col = nt * 8 + 2 * t;
        float* b0 = dbias + (size_t(head) * WT + R0) * WT + col;
        float* b1 = dbias + (size_t(head) * WT + R1) * WT + col;
        atomicAdd(b0, dB[nt][0]);
        atomicAdd(b0 + 1, dB[nt][1]);
        atomicAdd(b1, dB[nt][2]);
        atomicAdd(b1 + 1, dB[nt][3]);
    }
    dscale_acc = warp_sum(dscale_acc);
    if (lane == 0) atomicAdd(dscale + head, dscale_acc);
}

// dx[b][2*h2 + (q & 1)][2*w2 + (q >> 1)][c] = dy[b][h2][w2][q * C + c]   (inverse of patch_merge_kernel)
__global__ void __launch_bounds__(256)
patch_merge_scatter_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, int H, int W, int C8, size_t total) {
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
        dx[((b * H + 2 * h2 + (q & 1)) * W + 2 * w2 + (q >> 1)) * C8 + c] = __ldg(dy + e);
    }
}

}  // namespace swinb

static int swinb_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    }
    return sms;
}

#define SWINB_LN_DISPATCH(KERNEL, ...)                                                              \
    if (emb <= 128) KERNEL<2><<<grid, LN_WARPS * 32, 0, st>>>(__VA_ARGS__);                          \
    else if (emb <= 192) KERNEL<3><<<grid, LN_WARPS * 32, 0, st>>>(__VA_ARGS__);                     \
    else if (emb <= 384) KERNEL<6><<<grid, LN_WARPS * 32, 0, st>>>(__VA_ARGS__);                     \
    else KERNEL<12><<<grid, LN_WARPS * 32, 0, st>>>(__VA_ARGS__)

extern "C" int rgbnm_layernorm_res_scaled_fwd(const void* x, const float* gamma, const float* beta, const void* res,
                                              const float* row_scale, int rows_per_scale, void* y, int rows, int emb, float eps,
                                              void* stream) {
    using namespace swinb;
    if (!x || !gamma || !beta || !y || rows < 0 || emb <= 0 || (emb & 1) || emb > 768) return RGBNM_ERR_ARG;
    if (row_scale != nullptr && rows_per_scale <= 0) return RGBNM_ERR_ARG;
    if (rows == 0) return RGBNM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int grid = (rows + LN_WARPS - 1) / LN_WARPS;
    if (grid > swinb_num_sms() * 8) grid = swinb_num_sms() * 8;
#define SWINB_LN_FWD_G(GG, PP)                                                                                                         \
    ln_res_scaled_fwd_g_kernel<GG, PP><<<grid, LN_WARPS * 32, 0, st>>>(static_cast<const unsigned*>(x), gamma, beta,                   \
                                                                        static_cast<const unsigned*>(res), row_scale, rows_per_scale, \
                                                                        static_cast<unsigned*>(y), rows, eps)
    if (emb == 96 || emb == 192 || emb == 384 || emb == 768) {
        if (emb == 96) SWINB_LN_FWD_G(8, 6);
        else if (emb == 192) SWINB_LN_FWD_G(16, 6);
        else if (emb == 384) SWINB_LN_FWD_G(32, 6);
        else SWINB_LN_FWD_G(32, 12);
        RGBNM_CUDA_CHECK(cudaGetLastError());
        return RGBNM_OK;
    }
#undef SWINB_LN_FWD_G
    SWINB_LN_DISPATCH(ln_res_scaled_fwd_kernel, static_cast<const unsigned*>(x), gamma, beta, static_cast<const unsigned*>(res), row_scale,
                      rows_per_scale, static_cast<unsigned*>(y), rows, emb, eps);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

extern "C" int rgbnm_layernorm_res_bwd(const void* dy, const void* x, const float* gamma, const float* row_scale, int rows_per_scale,
                                       void* dx, float* dgamma, float* dbeta, int rows, int emb, float eps, void* stream) {
    return rgbnm_layernorm_res_bwd_ex(dy, x, gamma, row_scale, rows_per_scale, dx, dgamma, dbeta, nullptr, rows, emb, eps, stream);
}

extern "C" int rgbnm_layernorm_res_bwd_ex(const void* dy, const void* x, const float* gamma, const float* row_scale, int rows_per_scale,
                                          void* dx, float* dgamma, float* dbeta, float* dxsum, int rows, int emb, float eps,
                                          void* stream) {
    using namespace swinb;
    if (!dy || !x || !gamma || !dx || !dgamma || !dbeta || rows < 0 || emb <= 0 || (emb & 1) || emb > 768) return RGBNM_ERR_ARG;
    if (dxsum != nullptr && emb != 96 && emb != 192 && emb != 384 && emb != 768) return RGBNM_ERR_UNSUPPORTED;
    if (row_scale != nullptr && rows_per_scale <= 0) return RGBNM_ERR_ARG;
    if (rows == 0) return RGBNM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int grid = (rows + LN_WARPS - 1) / LN_WARPS;
    if (grid > swinb_num_sms() * 4) grid = swinb_num_sms() * 4;
#define SWINB_LN_BWD_G(GG, PP)                                                                                                     \
    ln_res_bwd_g_kernel<GG, PP><<<grid, LN_WARPS * 32, 0, st>>>(static_cast<const unsigned*>(dy), static_cast<const unsigned*>(x), \
                                                                 gamma, row_scale, rows_per_scale, static_cast<unsigned*>(dx),     \
                                                                 dgamma, dbeta, dxsum, rows, eps)
    if (emb == 96 || emb == 192 || emb == 384 || emb == 768) {
        if (emb == 96) SWINB_LN_BWD_G(8, 6);
        else if (emb == 192) SWINB_LN_BWD_G(16, 6);
        else if (emb == 384) SWINB_LN_BWD_G(32, 6);
        else SWINB_LN_BWD_G(32, 12);
        RGBNM_CUDA_CHECK(cudaGetLastError());
        return RGBNM_OK;
    }
#undef SWINB_LN_BWD_G
    SWINB_LN_DISPATCH(ln_res_bwd_kernel, static_cast<const unsigned*>(dy), static_cast<const unsigned*>(x), gamma, row_scale, rows_per_scale,
                      static_cast<unsigned*>(dx), dgamma, dbeta, rows, emb, eps);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

extern "C" int rgbnm_window_attention_bwd(const void* qkv, const void* dout, const float* bias, const float* scale, void* dqkv,
                                          float* dbias, float* dscale, int B, int H, int W, int C, int heads, int window, int shift,
                                          void* stream) {
    using namespace swinb;
    if (!qkv || !dout || !bias || !scale || !dqkv || !dbias || !dscale || B < 0) return RGBNM_ERR_ARG;
    if (window != WS || heads <= 0 || C != heads * HD) return RGBNM_ERR_UNSUPPORTED;
    if (H <= 0 || W <= 0 || (H % WS) || (W % WS) || shift < 0 || shift >= WS) return RGBNM_ERR_ARG;
    if (B == 0) return RGBNM_OK;
    static bool configured = false;
    if (!configured) {
        RGBNM_CUDA_CHECK(cudaFuncSetAttribute(window_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(AttnBwdSmem))));
        configured = true;
    }
    const int n_windows = B * (H / WS) * (W / WS);
    static const bool simt = [] { const char* e = std::getenv("RGBNM_WATTN_BWD_SIMT"); return e != nullptr && e[0] == '1'; }();
    if (!simt) {
        static bool configured_mma = false;
        if (!configured_mma) {
            RGBNM_CUDA_CHECK(cudaFuncSetAttribute(window_attn_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  int(sizeof(AttnBwdMmaSmem))));
            configured_mma = true;
        }
        // one resident wave of CTAs, each walking the windows of one head
        int per_head = swinb_num_sms() * RGBNM_WATTN_BWD_MINBLOCKS / heads;
        if (per_head < 1) per_head = 1;
        if (per_head > n_windows) per_head = n_windows;
        window_attn_bwd_mma_kernel<<<dim3(per_head, heads), 128, sizeof(AttnBwdMmaSmem), static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(dout), bias, scale, static_cast<__nv_bfloat16*>(dqkv),
            dbias, dscale, H, W, C, shift, n_windows);
        RGBNM_CUDA_CHECK(cudaGetLastError());
        return RGBNM_OK;
    }
    int ctas = swinb_num_sms() * 2 / heads;
    if (ctas < 1) ctas = 1;
    if (ctas > n_windows) ctas = n_windows;
    window_attn_bwd_kernel<<<dim3(ctas, heads), WT, sizeof(AttnBwdSmem), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(dout), bias, scale, static_cast<__nv_bfloat16*>(dqkv),
        dbias, dscale, H, W, C, shift, n_windows);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

extern "C" int rgbnm_patch_merge_scatter(const void* dy, void* dx, int B, int H, int W, int C, void* stream) {
    using namespace swinb;
    if (!dy || !dx || B < 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || C <= 0 || (C % 8)) return RGBNM_ERR_ARG;
    if (B == 0) return RGBNM_OK;
    const size_t total = size_t(B) * (H / 2) * (W / 2) * 4 * (C / 8);
    size_t blocks = (total + 255) / 256;
    if (blocks > size_t(swinb_num_sms()) * 16) blocks = size_t(swinb_num_sms()) * 16;
    patch_merge_scatter_kernel<<<unsigned(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(dy), static_cast<uint4*>(dx), H, W, C / 8, total);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}
