// Memory-bound kernels of the DCT ViT around the tensor-core contractions (sm_100a):
//   LayerNorm forward / backward                models/plainvit.py:513,522,551 (eps 1e-5, fp32 statistics)
//   bias gradients (column sums)                nn.Linear backward
//   bf16 working copies of the fp32 master weights (row permutation for the fused qkv
//   projection, transposed copies for dgrad)    plainvit.py:441-447
//   gradient-norm + fused AdamW / decoupled weight decay step
//                                               train.py:163-172, utils/custom_optims.py:37-43
// All are HBM-bound: 4-byte-per-lane coalesced accesses (a warp reads 128 contiguous bytes), one
// warp per token row, grid sized as a multiple of the SM count.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/rgbnm_b200.h"
#include "common.cuh"

namespace vitk {

constexpr int LN_WARPS = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// kernel qkv order (which*(H*D) + h*D + d) -> reference order (h*3D + d*3 + which); plainvit.py:447
__device__ __forceinline__ int qkv_unperm(int r, int heads, int hd) {
    const int which = r / (heads * hd), rem = r - which * heads * hd;
    const int h = rem / hd, d = rem - h * hd;
    return h * (3 * hd) + d * 3 + which;
}
__device__ __forceinline__ float2 bf2_to_f2(unsigned w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ unsigned f2_to_bf2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<unsigned*>(&v);
}

// ------------------------------------------------------------------------------------------
// LayerNorm forward: y = (x - mean) * rstd * gamma + beta   (x, y bf16; statistics fp32)
// PPL = E / 64 bf16 pairs per lane.
// ------------------------------------------------------------------------------------------
template <int PPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const unsigned* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              unsigned* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, float eps) {
    constexpr int E = PPL * 64;
    const int lane = threadIdx.x & 31;
    float2 g[PPL], b[PPL];
#pragma unroll
    for (int k = 0; k < PPL; ++k) {
        g[k] = *reinterpret_cast<const float2*>(gamma + 2 * (k * 32 + lane));
        b[k] = *reinterpret_cast<const float2*>(beta + 2 * (k * 32 + lane));
    }
    // software-pipelined like the backward kernel: the next row's loads are in flight during this row's two reductions
    const int stride = gridDim.x * LN_WARPS;
    int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    unsigned nx[PPL];
    if (row < rows) {
#pragma unroll
        for (int k = 0; k < PPL; ++k) nx[k] = __ldg(x + size_t(row) * (E / 2) + k * 32 + lane);
    }
    for (; row < rows; row += stride) {
        float2 v[PPL];
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < PPL; ++k) { v[k] = bf2_to_f2(nx[k]); s += v[k].x + v[k].y; }
        if (row + stride < rows) {
#pragma unroll
            for (int k = 0; k < PPL; ++k) nx[k] = __ldg(x + size_t(row + stride) * (E / 2) + k * 32 + lane);
        }
        const float mean = warp_sum(s) * (1.0f / E);
        float q = 0.0f;
#pragma unroll
        for (int k = 0; k < PPL; ++k) { const float a = v[k].x - mean, c = v[k].y - mean; q += a * a + c * c; }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / E) + eps);
        unsigned* yr = y + size_t(row) * (E / 2);
#pragma unroll
        for (int k = 0; k < PPL; ++k)
            yr[k * 32 + lane] = f2_to_bf2((v[k].x - mean) * rstd * g[k].x + b[k].x, (v[k].y - mean) * rstd * g[k].y + b[k].y);
        if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
    }
}

// ------------------------------------------------------------------------------------------
// LayerNorm backward:  dx = rstd * (dy*g - mean_E(dy*g) - xhat * mean_E(dy*g*xhat)) [+ dres]
//                      dgamma += sum_rows dy * xhat,  dbeta += sum_rows dy   (fp32 atomics, one per column per CTA)
// ------------------------------------------------------------------------------------------
template <int PPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_kernel(const unsigned* __restrict__ dy, const unsigned* __restrict__ x, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, const float* __restrict__ gamma, const unsigned* __restrict__ dres,
              unsigned* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum,
              int rows, int dy_div) {
    constexpr int E = PPL * 64;
    __shared__ float red[LN_WARPS][E + 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2 g[PPL], dg[PPL], db[PPL], dc[PPL];
#pragma unroll
    for (int k = 0; k < PPL; ++k) {
        g[k] = *reinterpret_cast<const float2*>(gamma + 2 * (k * 32 + lane));
        dg[k] = make_float2(0.0f, 0.0f);
        db[k] = make_float2(0.0f, 0.0f);
        dc[k] = make_float2(0.0f, 0.0f);
    }
    // software-pipelined over the rows of this warp: the loads of the next row are issued before the current row's two
    // warp reductions, so the ~1 us global-load latency overlaps them (each warp walks ~10 rows; 16 warps per SM)
    const int stride = gridDim.x * LN_WARPS;
    int row = blockIdx.x * LN_WARPS + warp;
    unsigned n_dy[PPL], n_x[PPL], n_dr[PPL];
    float n_mean = 0.0f, n_rstd = 0.0f;
    auto issue = [&](int r) {
        const size_t o = size_t(r) * (E / 2);
        const size_t od = size_t(r / dy_div) * (E / 2);          // dy_div > 1: one dy row shared by dy_div consecutive rows (token mean)
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            n_dy[k] = __ldg(dy + od + k * 32 + lane);
            n_x[k] = __ldg(x + o + k * 32 + lane);
            n_dr[k] = dres != nullptr ? __ldg(dres + o + k * 32 + lane) : 0u;
        }
        n_mean = mean_in[r];
        n_rstd = rstd_in[r];
    };
    if (row < rows) issue(row);
    for (; row < rows; row += stride) {
        const size_t off = size_t(row) * (E / 2);
        const float mean = n_mean, rstd = n_rstd;
        unsigned c_dy[PPL], c_x[PPL], c_dr[PPL];
#pragma unroll
        for (int k = 0; k < PPL; ++k) { c_dy[k] = n_dy[k]; c_x[k] = n_x[k]; c_dr[k] = n_dr[k]; }
        if (row + stride < rows) issue(row + stride);
        float2 d[PPL], xh[PPL];
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            d[k] = bf2_to_f2(c_dy[k]);
            const float2 xv = bf2_to_f2(c_x[k]);
            xh[k] = make_float2((xv.x - mean) * rstd, (xv.y - mean) * rstd);
            dg[k].x += d[k].x * xh[k].x; dg[k].y += d[k].y * xh[k].y;
            db[k].x += d[k].x; db[k].y += d[k].y;
            d[k].x *= g[k].x; d[k].y *= g[k].y;
            s1 += d[k].x + d[k].y;
            s2 += d[k].x * xh[k].x + d[k].y * xh[k].y;
        }
        s1 = warp_sum(s1) * (1.0f / E);
        s2 = warp_sum(s2) * (1.0f / E);
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            float a = rstd * (d[k].x - s1 - xh[k].x * s2), c = rstd * (d[k].y - s1 - xh[k].y * s2);
            if (dres != nullptr) {
                const float2 r = bf2_to_f2(c_dr[k]);
                a += r.x; c += r.y;
            }
            const unsigned packed = f2_to_bf2(a, c);
            dx[off + k * 32 + lane] = packed;
            // column sums of the stored (bf16-rounded) dx: the bias gradient of the Linear that produced this residual stream
            const float2 r = bf2_to_f2(packed);
            dc[k].x += r.x; dc[k].y += r.y;
        }
    }
    // CTA-level reduction of the per-lane column partials, then one atomic per column per CTA
    for (int pass = 0; pass < (dxsum != nullptr ? 3 : 2); ++pass) {
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            const float2 v = pass == 0 ? dg[k] : (pass == 1 ? db[k] : dc[k]);
            red[warp][2 * (k * 32 + lane)] = v.x;
            red[warp][2 * (k * 32 + lane) + 1] = v.y;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < E; c += LN_WARPS * 32) {
            float s = 0.0f;
#pragma unroll
            for (int w = 0; w < LN_WARPS; ++w) s += red[w][c];
            atomicAdd((pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum)) + c, s);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Column sums of a bf16 matrix [rows][cols] (ld elements) -> fp32 out[cols] += sum_rows
// A warp covers 256 columns of one row with 16-byte loads; 4 independent rows in flight per thread.
// CTA = 8 warps on 8 interleaved row streams; grid (ceil(cols/256), row chunks).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ a, long long ld, int rows, int cols, float* __restrict__ out, int heads, int hd) {
    __shared__ float red[8][256 + 8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 256 + 8 * lane;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    if (c0 < cols) {
        const int stride = gridDim.y * 8;
        int r = blockIdx.y * 8 + w;
        for (; r + 3 * stride < rows; r += 4 * stride) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(a + size_t(r + u * stride) * ld + c0));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned wd[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float2 f = bf2_to_f2(wd[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
            }
        }
        for (; r < rows; r += stride) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(a + size_t(r) * ld + c0));
            const unsigned wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 f = bf2_to_f2(wd[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[w][8 * lane + j] = acc[j];
    __syncthreads();
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < cols) atomicAdd(out + (heads > 0 ? qkv_unperm(c, heads, hd) : c), s);
}

// ------------------------------------------------------------------------------------------
// fp32 master weight [N][K] -> bf16 W'[N][K] with W'[perm(n)] = W[n], and its transpose Wt[K][N]
// perm: identity, or the qkv regrouping  n = h*(3*D) + d*3 + which  ->  which*(H*D) + h*D + d
// (plainvit.py:447 "b n (h d qkv) -> (qkv) b h n d": qkv is the innermost factor of the reference layout)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int qkv_perm(int n, int heads, int hd) {
    const int which = n % 3, d = (n / 3) % hd, h = n / (3 * hd);
    return which * (heads * hd) + h * hd + d;
}

__global__ void __launch_bounds__(256)
weight_prep_kernel(const float* __restrict__ w, int N, int K, int heads, int hd, __nv_bfloat16* __restrict__ wb,
                   __nv_bfloat16* __restrict__ wt) {
    __shared__ float tile[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int n = n0 + i, k = k0 + tx;
        float v = 0.0f;
        if (n < N && k < K) {
            v = w[size_t(n) * K + k];
            const int np = heads > 0 ? qkv_perm(n, heads, hd) : n;
            wb[size_t(np) * K + k] = __float2bfloat16_rn(v);
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    if (wt != nullptr) {
        if (heads > 0) {
            // permuted rows are not contiguous: write element-wise (this kernel runs once per step on ~MBs)
            for (int i = ty; i < 32; i += 8) {
                const int k = k0 + i, n = n0 + tx;
                if (n < N && k < K) wt[size_t(k) * N + qkv_perm(n, heads, hd)] = __float2bfloat16_rn(tile[tx][i]);
            }
        } else {
            for (int i = ty; i < 32; i += 8) {
                const int k = k0 + i, n = n0 + tx;
                if (n < N && k < K) wt[size_t(k) * N + n] = __float2bfloat16_rn(tile[tx][i]);
            }
        }
    }
}

// All Linear layers of the model in ONE launch (49 launches of a few microseconds each otherwise dominate the refresh):
// block -> descriptor by a scan over the per-descriptor first-tile table, then the body of weight_prep_kernel.  The block
// owning tile 0 of a qkv descriptor also regroups that layer's bias into kernel order.
__global__ void __launch_bounds__(256)
weight_prep_batch_kernel(const rgbnm_wprep_desc* __restrict__ descs, int n_desc) {
    __shared__ float tile[32][33];
    __shared__ int s_first[256];
    __shared__ int s_d;
    // block -> descriptor: the first-tile table is fetched by all threads at once (one load latency), then counted
    for (int i = threadIdx.x; i < n_desc; i += 256) s_first[i] = descs[i].first_tile;
    if (threadIdx.x == 0) s_d = 0;
    __syncthreads();
    {
        int mine = 0;
        for (int i = threadIdx.x; i < n_desc; i += 256) mine += (int(blockIdx.x) >= s_first[i]) ? 1 : 0;
        if (mine) atomicAdd(&s_d, mine);
    }
    __syncthreads();
    const rgbnm_wprep_desc D = descs[s_d - 1];
    const int t = int(blockIdx.x) - D.first_tile;
    const int tiles_k = (D.k + 31) / 32;
    const int k0 = (t % tiles_k) * 32, n0 = (t / tiles_k) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* __restrict__ w = D.w;
    __nv_bfloat16* __restrict__ wb = static_cast<__nv_bfloat16*>(D.w_bf16);
    __nv_bfloat16* __restrict__ wt = static_cast<__nv_bfloat16*>(D.wt_bf16);
    const int N = D.n, K = D.k, heads = D.qkv_heads, hd = D.head_dim;
    for (int i = ty; i < 32; i += 8) {
        const int n = n0 + i, k = k0 + tx;
        float v = 0.0f;
        if (n < N && k < K) {
            v = w[size_t(n) * K + k];
            const int np = heads > 0 ? qkv_perm(n, heads, hd) : n;
            wb[size_t(np) * K + k] = __float2bfloat16_rn(v);
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    if (wt != nullptr) {
        for (int i = ty; i < 32; i += 8) {
            const int k = k0 + i, n = n0 + tx;
            if (n < N && k < K) wt[size_t(k) * N + (heads > 0 ? qkv_perm(n, heads, hd) : n)] = __float2bfloat16_rn(tile[tx][i]);
        }
    }
    if (t == 0 && heads > 0 && D.bias != nullptr) {
        for (int i = threadIdx.x; i < N; i += 256) D.bias_k[qkv_perm(i, heads, hd)] = D.bias[i];
    }
}

// bias (fp32, reference order) -> fp32 working copy in kernel order; gradient (kernel order) -> reference order
__global__ void perm_vec_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int heads, int hd, int inverse) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = qkv_perm(i, heads, hd);
    if (inverse) dst[i] = src[p];
    else dst[p] = src[i];
}
// weight gradient accumulated in kernel (permuted-row) order [N][K] -> add into reference order
__global__ void unperm_rows_add_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int K, int heads, int hd) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= size_t(N) * K) return;
    const int n = int(i / K), k = int(i % K);
    dst[i] += src[size_t(qkv_perm(n, heads, hd)) * K + k];
}

// ------------------------------------------------------------------------------------------
// sum of squares (gradient norm) and the fused optimiser step
// ------------------------------------------------------------------------------------------
// Deterministic: per-block partial sums go to a scratch array, the block that finishes last adds them up in index order.  The
// result feeds the clip factor of the optimiser step, which every data-parallel replica must evaluate to the SAME bits from the
// same all-reduced gradient -- an atomicAdd per block (arrival order) let replicas drift apart by an ulp per step
// (tests/test_ddp_gpu.py).
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, size_t n, float* __restrict__ out, float* __restrict__ partial, unsigned* __restrict__ counter) {
    __shared__ float red[8];
    __shared__ bool last;
    float s = 0.0f;
    const size_t n4 = n / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) {
        const float4 v = __ldg(g4 + i);
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[n4 * 4 + threadIdx.x]; s += v * v; }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int k = 0; k < 8; ++k) t += red[k];
        partial[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // fixed-shape tree over the partials: thread t sums partial[t], partial[t + 256], ... in order, then the block reduces
    float t = 0.0f;
    for (unsigned k = threadIdx.x; k < gridDim.x; k += 256) t += __ldcg(partial + k);
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.0f;
        for (int k = 0; k < 8; ++k) tot += red[k];
        *out += tot;
        *counter = 0;              // ready for the next launch on this stream
    }
}

// AdamW with weight_decay = 0 (pipeline_utils.py:536), preceded by clip_grad_norm_(max_norm) (train.py:163)
// and followed by the reference's separate decoupled decay p -= (lr / base_lr) * wd * p on the first
// `n_decay` elements of the flat buffer (custom_optims.py:37-43; Linear weights are laid out first).
// hyper (device, so a captured CUDA graph sees new values every replay):
//   [0] lr  [1] beta1  [2] beta2  [3] eps  [4] 1-beta1^t  [5] 1-beta2^t  [6] decay = lr/base_lr*wd  [7] grad_scale  [8] max_norm
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
             size_t n_decay, const float* __restrict__ gnorm_sq, const float* __restrict__ hyper) {
    const float lr = hyper[0], beta1 = hyper[1], beta2 = hyper[2], eps = hyper[3], bc1 = hyper[4], bc2 = hyper[5];
    const float decay = hyper[6], grad_scale = hyper[7], max_norm = hyper[8];
    float clip = 1.0f;
    if (max_norm > 0.0f) {
        const float total = sqrtf(*gnorm_sq) * grad_scale;
        clip = fminf(max_norm / (total + 1e-6f), 1.0f);
    }
    const float gs = grad_scale * clip;
    const float step = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const float gi = g[i] * gs;
        const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float pi = p[i] - step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
        if (i < n_decay) pi -= decay * pi;
        p[i] = pi;
    }
}

// RandomMixup_DCT on the embed input (cls_transforms.py:135-182): out[b] = lam0 * x[b] + lam1 * x[b-1 mod B], the batch
// rolled by one image.  lam lives in device memory (a captured CUDA graph replays with a fresh draw).  per_image = bf16
// elements per image, a multiple of 8.
__global__ void __launch_bounds__(256)
mixup_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, const float* __restrict__ lam, int batch, size_t vec_per_image) {
    const float l0 = lam[0], l1 = lam[1];
    const size_t total = size_t(batch) * vec_per_image;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const size_t b = i / vec_per_image, r = i - b * vec_per_image;
        const size_t j = (b == 0 ? size_t(batch) - 1 : b - 1) * vec_per_image + r;
        const uint4 a = __ldg(x + i), c = __ldg(x + j);
        const unsigned aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
        unsigned o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 fa = bf2_to_f2(aw[e]), fc = bf2_to_f2(cw[e]);
            o[e] = f2_to_bf2(fa.x * l0 + fc.x * l1, fa.y * l0 + fc.y * l1);
        }
        out[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

static int g_sms = 0;
static int sms() {
    if (g_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_sms;
}

}  // namespace vitk

extern "C" {

int rgbnm_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int rows,
                        int emb, float eps, void* stream) {
    using namespace vitk;
    if (!x || !gamma || !beta || !y || !mean || !rstd || rows < 0) return RGBNM_ERR_ARG;
    if (rows == 0) return RGBNM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = min((rows + LN_WARPS - 1) / LN_WARPS, sms() * 8);
    const unsigned* xi = static_cast<const unsigned*>(x);
    unsigned* yo = static_cast<unsigned*>(y);
    if (emb == 192) ln_fwd_kernel<3><<<grid, LN_WARPS * 32, 0, st>>>(xi, gamma, beta, yo, mean, rstd, rows, eps);
    else if (emb == 384) ln_fwd_kernel<6><<<grid, LN_WARPS * 32, 0, st>>>(xi, gamma, beta, yo, mean, rstd, rows, eps);
    else if (emb == 768) ln_fwd_kernel<12><<<grid, LN_WARPS * 32, 0, st>>>(xi, gamma, beta, yo, mean, rstd, rows, eps);
    else return RGBNM_ERR_UNSUPPORTED;
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

int rgbnm_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                        const void* dres, void* dx, float* dgamma, float* dbeta, float* dxsum, int rows, int emb, void* stream) {
    return rgbnm_layernorm_bwd_ex(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, dxsum, rows, emb, 1, stream);
}

int rgbnm_layernorm_bwd_ex(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                           const void* dres, void* dx, float* dgamma, float* dbeta, float* dxsum, int rows, int emb,
                           int rows_per_dy_row, void* stream) {
    using namespace vitk;
    if (!dy || !x || !mean || !rstd || !gamma || !dx || !dgamma || !dbeta || rows < 0 || rows_per_dy_row < 1) return RGBNM_ERR_ARG;
    const int dy_div = rows_per_dy_row;
    if (rows == 0) return RGBNM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = min((rows + LN_WARPS - 1) / LN_WARPS, sms() * 4);
    const unsigned* a = static_cast<const unsigned*>(dy);
    const unsigned* b = static_cast<const unsigned*>(x);
    const unsigned* r = static_cast<const unsigned*>(dres);
    unsigned* o = static_cast<unsigned*>(dx);
    if (emb == 192) ln_bwd_kernel<3><<<grid, LN_WARPS * 32, 0, st>>>(a, b, mean, rstd, gamma, r, o, dgamma, dbeta, dxsum, rows, dy_div);
    else if (emb == 384) ln_bwd_kernel<6><<<grid, LN_WARPS * 32, 0, st>>>(a, b, mean, rstd, gamma, r, o, dgamma, dbeta, dxsum, rows, dy_div);
    else if (emb == 768) ln_bwd_kernel<12><<<grid, LN_WARPS * 32, 0, st>>>(a, b, mean, rstd, gamma, r, o, dgamma, dbeta, dxsum, rows, dy_div);
    else return RGBNM_ERR_UNSUPPORTED;
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

int rgbnm_mixup_bf16(const void* x, void* out, const float* lam, int batch, long long per_image, void* stream) {
    if (!x || !out || !lam || batch <= 0 || per_image <= 0 || (per_image & 7) || x == out) return RGBNM_ERR_ARG;
    const size_t total = size_t(batch) * size_t(per_image / 8);
    size_t blocks = (total + 255) / 256;
    const int grid = int(blocks > size_t(vitk::sms()) * 16 ? size_t(vitk::sms()) * 16 : blocks);
    vitk::mixup_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(x), static_cast<uint4*>(out), lam,
                                                                           batch, size_t(per_image / 8));
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

int rgbnm_colsum_bf16(const void* a, long long ld, int rows, int cols, float* out, int qkv_heads, int head_dim, void* stream) {
    if (!a || !out || rows < 0 || cols <= 0 || (ld & 7) || (cols & 7)) return RGBNM_ERR_ARG;
    if (qkv_heads > 0 && cols != 3 * qkv_heads * head_dim) return RGBNM_ERR_ARG;
    if (rows == 0) return RGBNM_OK;
    const int gx = (cols + 255) / 256;
    int gy = (vitk::sms() * 4 + gx - 1) / gx;
    if (gy > (rows + 7) / 8) gy = (rows + 7) / 8;
    vitk::colsum_kernel<<<dim3(gx, gy), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(a), ld, rows, cols, out, qkv_heads, head_dim);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

int rgbnm_weight_prep(const float* w, int n, int k, int qkv_heads, int head_dim, void* w_bf16, void* wt_bf16, void* stream) {
    if (!w || !w_bf16 || n <= 0 || k <= 0) return RGBNM_ERR_ARG;
    if (qkv_heads > 0 && n != 3 * qkv_heads * head_dim) return RGBNM_ERR_ARG;
    vitk::weight_prep_kernel<<<dim3((k + 31) / 32, (n + 31) / 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w, n, k, qkv_heads, head_dim, static_cast<__nv_bfloat16*>(w_bf16), static_cast<__nv_bfloat16*>(wt_bf16));
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

int rgbnm_weight_prep_batch(const rgbnm_wprep_desc* descs_dev, int n_desc, int total_tiles, void* stream) {
    if (!descs_dev || n_desc <= 0 || n_desc > 256 || total_tiles <= 0) return RGBNM_ERR_ARG;
    vitk::weight_prep_batch_kernel<<<total_tiles, 256, 0, static_cast<cudaStream_t>(stream)>>>(descs_dev, n_desc);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

int rgbnm_qkv_perm_vec(const float* src, float* dst, int n, int heads, int head_dim, int inverse, void* stream) {
    if (!src || !dst || n != 3 * heads * head_dim) return RGBNM_ERR_ARG;
    vitk::perm_vec_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n, heads, head_dim, inverse);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

int rgbnm_qkv_unperm_rows_add(const float* src, float* dst, int n, int k, int heads, int head_dim, void* stream) {
    if (!src || !dst || n != 3 * heads * head_dim || k <= 0) return RGBNM_ERR_ARG;
    const size_t total = size_t(n) * k;
    vitk::unperm_rows_add_kernel<<<unsigned((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n, k, heads, head_dim);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

int rgbnm_sumsq_f32(const float* g, long long n, float* out, void* stream) {
    if (!g || !out || n < 0) return RGBNM_ERR_ARG;
    if (n == 0) return RGBNM_OK;
    long long blocks = (n / 4 + 255) / 256;
    const int grid = int(blocks < 1 ? 1 : (blocks > vitk::sms() * 8 ? vitk::sms() * 8 : blocks));
    // per-device scratch (partials + arrival counter), allocated on first use -- i.e. during the warm-up steps that precede any
    // CUDA-graph capture; launches that share it must be stream-ordered (one optimiser step at a time per device)
    static float* scratch[64] = {nullptr};
    int dev = 0;
    RGBNM_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return RGBNM_ERR_ARG;
    if (scratch[dev] == nullptr) {
        const size_t bytes = (size_t(vitk::sms()) * 8 + 1) * sizeof(float);
        RGBNM_CUDA_CHECK(cudaMalloc(&scratch[dev], bytes));
        RGBNM_CUDA_CHECK(cudaMemset(scratch[dev], 0, bytes));
    }
    float* partial = scratch[dev] + 1;
    unsigned* counter = reinterpret_cast<unsigned*>(scratch[dev]);
    vitk::sumsq_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(g, size_t(n), out, partial, counter);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

int rgbnm_adamw_step(float* p, const float* g, float* m, float* v, long long n, long long n_decay, const float* gnorm_sq,
                     const float* hyper, void* stream) {
    if (!p || !g || !m || !v || n < 0 || !hyper || !gnorm_sq) return RGBNM_ERR_ARG;
    if (n == 0) return RGBNM_OK;
    long long blocks = (n + 255) / 256;
    const int grid = int(blocks > vitk::sms() * 16 ? vitk::sms() * 16 : blocks);
    vitk::adamw_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, size_t(n), size_t(n_decay), gnorm_sq, hyper);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

}  // extern "C"
