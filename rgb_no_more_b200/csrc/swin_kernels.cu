// Kernels of the SwinV2 DCT forward path around the tensor-core contractions (sm_100a); SURVEY.md 8a row a33.
//   post-norm LayerNorm + residual      x = shortcut + norm(branch(x))            models/swinv2.py:302-306
//   window attention, forward           cosine attention + continuous relative position bias + shift mask
//                                                                                 models/swinv2.py:143-182, 244-300
//   patch-merging gather                x0 | x1 | x2 | x3 concat                  models/swinv2.py:346-362
//   token mean                          AdaptiveAvgPool1d(1)                      models/swinv2.py:697-699
// Window partition / cyclic shift / window reverse (swinv2.py:39-66, 283-300) are index maps: the attention kernel
// gathers each window's tokens from their image positions and scatters the result back, nothing is materialised.
// All four are HBM-bound by design; the attention arithmetic (64 x 64 x 32 per window and head) runs on the
// CUDA cores in fp32 (round 1; a tensor-core version is the next step).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

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
    for (int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5); row < rows; row += gridDim.x * LN_WARPS) {
        const unsigned* xr = x + size_t(row) * pairs;
        float2 v[MAXP];
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
            const int p = k * 32 + lane;
            v[k] = p < pairs ? bf2_to_f2(__ldg(xr + p)) : make_float2(0.0f, 0.0f);
            s += v[k].x + v[k].y;
        }
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
                float o0 = (v[k].x - mean) * rstd * g.x + b.x;
                float o1 = (v[k].y - mean) * rstd * g.y + b.y;
                if (res != nullptr) {
                    const float2 r = bf2_to_f2(__ldg(res + size_t(row) * pairs + p));
                    o0 += r.x;
                    o1 += r.y;
                }
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
    const dim3 grid(unsigned(B) * (H / WS) * (W / WS), heads);
    window_attn_fwd_kernel<<<grid, WT, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), bias, scale, H, W, C, shift);
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
