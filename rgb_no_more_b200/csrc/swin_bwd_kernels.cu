// Backward kernels of the SwinV2 DCT path (training): first correct CUDA versions, fp32 on the CUDA
// cores -- the tensor-core forms come after parity.
//   post-norm residual with stochastic depth   y = res + s_b * LayerNorm(x)          models/swinv2.py:302-306 (drop_path)
//       forward with the per-image scale s_b, backward dx / dgamma / dbeta (dres = dy needs no kernel)
//   window attention backward                  dqkv, d(bias tile), d(logit scale)    models/swinv2.py:152-177
//   patch-merging scatter                      inverse of the forward gather         models/swinv2.py:353-358
#include <cuda_bf16.h>
#include <cuda_runtime.h>

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
    SWINB_LN_DISPATCH(ln_res_scaled_fwd_kernel, static_cast<const unsigned*>(x), gamma, beta, static_cast<const unsigned*>(res), row_scale,
                      rows_per_scale, static_cast<unsigned*>(y), rows, emb, eps);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

extern "C" int rgbnm_layernorm_res_bwd(const void* dy, const void* x, const float* gamma, const float* row_scale, int rows_per_scale,
                                       void* dx, float* dgamma, float* dbeta, int rows, int emb, float eps, void* stream) {
    using namespace swinb;
    if (!dy || !x || !gamma || !dx || !dgamma || !dbeta || rows < 0 || emb <= 0 || (emb & 1) || emb > 768) return RGBNM_ERR_ARG;
    if (row_scale != nullptr && rows_per_scale <= 0) return RGBNM_ERR_ARG;
    if (rows == 0) return RGBNM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int grid = (rows + LN_WARPS - 1) / LN_WARPS;
    if (grid > swinb_num_sms() * 4) grid = swinb_num_sms() * 4;
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
