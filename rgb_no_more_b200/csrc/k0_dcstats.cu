// K0 pre-pass: per-image DC statistics for the three RandAugment ops that need a
// whole-image reduction (SURVEY.md 7 "hard part 3"):
//   Brightness      mean(|DC_Y|) * m          utils/dct_ops.py:831-832
//   AutoContrast    min / max of DC_Y         utils/dct_ops.py:875-882
//   AutoSaturation  min / max of DC_CbCr (joint over both chroma planes)
// The reduction must see the DC plane *as it is when the op runs*, i.e. after the resize,
// the flip and every earlier op (translate / cutout zero blocks, posterize moves DCs, ...).
// Only the DC term of each block (1/64 of the data) is involved, so one small CTA per
// flagged image replays the plan on the 28x28 + 2x14x14 DC planes in shared memory and
// writes the resolved scalars to `stats`; the fused kernel then stays a pure per-block
// function.  Images whose plan has none of the three ops exit immediately.
#include <cuda_runtime.h>

#include "../../include/rgbnm_b200.h"
#include "common.cuh"
#include "k0_common.cuh"

namespace k0 {

// 1024 threads: the DC gather (1176 / 1536 post-resize blocks per image, up to 8 dependent-free 16-byte loads each) is one round of
// loads per thread plus a short second one -- with 256 threads it was five serial rounds of DRAM latency.
constexpr int STATS_THREADS = 1024;
constexpr int EQ_BINS_PER_THREAD = 2048 / STATS_THREADS;

// DC of the resized block at post-resize position (r, c) of plane `comp`, computed with the
// exact operation sequence of the fused kernel's row pass + column pass.
__device__ float resized_dc(const int16_t* __restrict__ plane, int W, const float* __restrict__ q,
                            const float* __restrict__ cq, int mode, bool clamp_in, int ci, int cj, int r, int c) {
    auto ld = [&](int brow, int bcol, int row, float (&x)[8]) {
        const int4 raw = __ldg(reinterpret_cast<const int4*>(plane + (size_t(brow) * W + bcol) * 64 + row * 8));
        dequant8(raw, q + row * 8, cq + row * 8, clamp_in, x);
    };
    if (mode == MODE_IDENT) {
        float x[8];
        ld(ci + r, cj + c, 0, x);
        return x[0];
    }
    if (mode == MODE_DOWN2) {
        // column 0 of R needs rows 0 and 8 of the 16x16 tile: first row of the top / bottom block pair
        float xl[8], xr[8], o[8], col_l[8], col_r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) col_l[i] = col_r[i] = 0.0f;
        ld(ci + 2 * r, cj + 2 * c, 0, xl);
        ld(ci + 2 * r, cj + 2 * c + 1, 0, xr);
        down2_1d<1, 1>(xl, xr, o);
        col_l[0] = o[0];
        ld(ci + 2 * r + 1, cj + 2 * c, 0, xl);
        ld(ci + 2 * r + 1, cj + 2 * c + 1, 0, xr);
        down2_1d<1, 1>(xl, xr, o);
        col_r[0] = o[0];
        float v[8];
        down2_1d<1, 2>(col_l, col_r, v);   // v[0] depends only on col_l[0], col_r[0]
        return rint_magic(v[0]);
    }
    // MODE_UP2
    float col[8], x[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        ld(ci + (r >> 1), cj + (c >> 1), i, x);
        up2_1d<1>(x, c & 1, o);
        col[i] = o[0];
    }
    float v[8];
    up2_1d<2>(col, r & 1, v);
    return rint_magic(v[0]);
}

__device__ __forceinline__ float block_reduce(float v, int op, float* scratch) {
    // op: 0 sum, 1 min, 2 max
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = op == 0 ? v + t : op == 1 ? fminf(v, t) : fmaxf(v, t);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = scratch[0];
    for (int w = 1; w < STATS_THREADS / 32; ++w) r = op == 0 ? r + scratch[w] : op == 1 ? fminf(r, scratch[w]) : fmaxf(r, scratch[w]);
    return r;
}

template <int LAYOUT>
__global__ void __launch_bounds__(STATS_THREADS)
k0_dcstats_kernel(const int16_t* __restrict__ y, const int16_t* __restrict__ cbcr, const int16_t* __restrict__ quant,
                  const rgbnm_plan* __restrict__ plans, rgbnm_k0_tables tb, float* __restrict__ stats_all, int hb, int wb) {
    constexpr int GRID_Y = Geo<LAYOUT>::GRID_Y, GRID_C = Geo<LAYOUT>::GRID_C;
    constexpr int NY = GRID_Y * GRID_Y;       // 784 / 1024
    constexpr int NC = 2 * GRID_C * GRID_C;   // 392 / 512
    constexpr int NDC = NY + NC;
    const int img = blockIdx.x;
    __shared__ rgbnm_plan pl;
    __shared__ __align__(16) float qf[192];
    __shared__ __align__(16) float cqf[192];
    __shared__ float dc[2][NDC];
    __shared__ float scratch[STATS_THREADS / 32];
    __shared__ int eq[2048];                 // Equalize: histogram of the luma DC plane, then the value -> value mapping
    __shared__ int iscratch[STATS_THREADS / 32 + 2];
    if (threadIdx.x < int(sizeof(rgbnm_plan) / 4))
        reinterpret_cast<int*>(&pl)[threadIdx.x] = __ldg(reinterpret_cast<const int*>(plans + img) + threadIdx.x);
    else if (threadIdx.x >= 32 && threadIdx.x < 32 + 192) {            // the tables travel together with the plan
        const int k = threadIdx.x - 32;
        qf[k] = float(__ldg(quant + size_t(img) * 192 + k));
        cqf[k] = -DEQ_BIAS * qf[k];
    }
    __syncthreads();
    if (!pl.needs_stats) return;

    const int hc = hb >> 1, wc = wb >> 1;
    const int mode = mode_of(pl.crop_size, GRID_Y);
    float* stats = stats_all + size_t(img) * RGBNM_MAX_OPS * 2;

    // post-resize, post-flip DC planes (flip only moves blocks: the DC term keeps its sign)
    for (int e = threadIdx.x; e < NDC; e += STATS_THREADS) {
        int comp, r, c;
        if (e < NY) { comp = 0; r = e / GRID_Y; c = e - r * GRID_Y; }
        else { const int f = e - NY; comp = 1 + f / (GRID_C * GRID_C); const int g = f % (GRID_C * GRID_C); r = g / GRID_C; c = g - r * GRID_C; }
        const int G = comp == 0 ? GRID_Y : GRID_C;
        const int sc = pl.flip ? G - 1 - c : c;
        const int16_t* plane = comp == 0 ? y + size_t(img) * hb * wb * 64
                                         : cbcr + (size_t(img) * 2 + (comp - 1)) * hc * wc * 64;
        float v = resized_dc(plane, comp == 0 ? wb : wc, qf + comp * 64, cqf + comp * 64, mode, pl.clamp_in != 0,
                             comp == 0 ? pl.crop_i : pl.crop_i >> 1, comp == 0 ? pl.crop_j : pl.crop_j >> 1, r, sc);
        if (pl.train) v = clampf(v);
        dc[0][e] = v;
    }
    __syncthreads();

    int cur = 0;
    for (int k = 0; k < pl.n_ops; ++k) {
        const rgbnm_plan_op op = pl.ops[k];
        const int code = op.code;
        float* src = dc[cur];
        float* dst = dc[cur ^ 1];
        // ---- reductions first (they read the plane as it stands before op k) ----
        float s0 = 0.0f, s1 = 0.0f;
        if (code == RGBNM_OP_BRIGHTNESS) {
            float a = 0.0f;
            for (int e = threadIdx.x; e < NY; e += STATS_THREADS) a += fabsf(src[e]);
            // integer-valued partial sums < 2^24: exact in fp32 whatever the order
            const float sum = block_reduce(a, 0, scratch);
            s0 = __fdiv_rn(sum, float(NY)) * op.f;      // torch.mean(|dc|) * m
        } else if (code == RGBNM_OP_AUTOCONTRAST || code == RGBNM_OP_AUTOSATURATION) {
            const int lo_e = code == RGBNM_OP_AUTOCONTRAST ? 0 : NY, hi_e = code == RGBNM_OP_AUTOCONTRAST ? NY : NDC;
            float mn = 3.0e38f, mx = -3.0e38f;
            for (int e = lo_e + threadIdx.x; e < hi_e; e += STATS_THREADS) { mn = fminf(mn, src[e]); mx = fmaxf(mx, src[e]); }
            s0 = block_reduce(mn, 1, scratch);
            s1 = block_reduce(mx, 2, scratch);
        } else if (code == RGBNM_OP_EQUALIZE) {
            // scale_channel_dct (dct_ops.py:916-940): hist = bincount(dc + 1024); cdf = cumsum(hist);
            // map[v] = round((cdf[v] - hist[first non-empty bin]) / (N - that) * 2039) - 1024, fp32 like torch's int / int
            for (int i = threadIdx.x; i < 2048; i += STATS_THREADS) eq[i] = 0;
            __syncthreads();
            for (int e = threadIdx.x; e < NY; e += STATS_THREADS) atomicAdd(&eq[int(src[e]) + 1024], 1);
            __syncthreads();
            int loc[EQ_BINS_PER_THREAD], sum = 0, first = 4096;
#pragma unroll
            for (int j = 0; j < EQ_BINS_PER_THREAD; ++j) {
                loc[j] = eq[threadIdx.x * EQ_BINS_PER_THREAD + j];
                if (loc[j] != 0 && first == 4096) first = threadIdx.x * EQ_BINS_PER_THREAD + j;
                sum += loc[j];
            }
            // block-wide exclusive scan of `sum` and minimum of `first`
            int incl = sum, fmin = first;
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if ((threadIdx.x & 31) >= o) incl += t;
            }
            for (int o = 16; o > 0; o >>= 1) fmin = min(fmin, __shfl_xor_sync(0xffffffffu, fmin, o));
            if ((threadIdx.x & 31) == 31) iscratch[threadIdx.x >> 5] = incl;
            __syncthreads();
            int base = 0;
            for (int w = 0; w < int(threadIdx.x >> 5); ++w) base += iscratch[w];
            __syncthreads();
            if ((threadIdx.x & 31) == 0) iscratch[threadIdx.x >> 5] = fmin;
            __syncthreads();
            for (int w = 0; w < STATS_THREADS / 32; ++w) fmin = min(fmin, iscratch[w]);
            const int h0 = eq[fmin];                       // every thread reads it before the table is overwritten
            const int mn = NY - h0;
            __syncthreads();
            int running = base + incl - sum;
#pragma unroll
            for (int j = 0; j < EQ_BINS_PER_THREAD; ++j) {
                const int b = threadIdx.x * EQ_BINS_PER_THREAD + j;
                running += loc[j];
                int val = b - 1024;                        // one distinct DC value (0 / 0 in the reference): unchanged
                if (mn > 0) val = int(rint_magic(__fdiv_rn(float(running - h0), float(mn)) * 2039.0f)) - 1024;
                eq[b] = val;
                if (tb.equalize_lut != nullptr)
                    tb.equalize_lut[(size_t(img) * RGBNM_MAX_OPS + k) * 2048 + b] = int16_t(max(-32768, min(32767, val)));
            }
            __syncthreads();
        }
        else if (code == RGBNM_OP_SOLARIZE) {
            // mask plane: luma DC above the threshold as the op runs (dct_ops.py:646); kept in `eq` and handed to the fused kernel
            for (int e = threadIdx.x; e < NY; e += STATS_THREADS) {
                const int m = src[e] > op.f ? 1 : 0;
                eq[e] = m;
                if (tb.equalize_lut != nullptr) tb.equalize_lut[(size_t(img) * RGBNM_MAX_OPS + k) * 2048 + e] = int16_t(m);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) { stats[2 * k] = s0; stats[2 * k + 1] = s1; }
        // ---- then apply op k to the DC planes ----
        for (int e = threadIdx.x; e < NDC; e += STATS_THREADS) {
            int comp, r, c;
            if (e < NY) { comp = 0; r = e / GRID_Y; c = e - r * GRID_Y; }
            else { const int f = e - NY; comp = 1 + f / (GRID_C * GRID_C); const int g = f % (GRID_C * GRID_C); r = g / GRID_C; c = g - r * GRID_C; }
            const int G = comp == 0 ? GRID_Y : GRID_C;
            const int base = comp == 0 ? 0 : NY + (comp - 1) * GRID_C * GRID_C;
            float v = src[e];
            if (code == RGBNM_OP_TRANSLATE_X) {
                const int nc = c - op.p[comp == 0 ? 0 : 1];
                v = (nc < 0 || nc >= G) ? 0.0f : src[base + r * G + nc];
            } else if (code == RGBNM_OP_TRANSLATE_Y) {
                const int nr = r - op.p[comp == 0 ? 0 : 1];
                v = (nr < 0 || nr >= G) ? 0.0f : src[base + nr * G + c];
            } else if (code == RGBNM_OP_ROT90) {
                const int sr = op.p[0] > 0 ? c : G - 1 - c, sc = op.p[0] > 0 ? G - 1 - r : r;
                v = src[base + sr * G + sc];
            } else if (code == RGBNM_OP_CUTOUT) {
                const int o = comp == 0 ? 0 : 4;
                if (r >= op.p[o] && r < op.p[o + 1] && c >= op.p[o + 2] && c < op.p[o + 3]) v = 0.0f;
            } else if (code == RGBNM_OP_GRAYSCALE) {
                if (comp != 0) v = 0.0f;
            } else if (code == RGBNM_OP_CHROMADROP) {
                if (comp == 1 + op.p[0]) v = 0.0f;
            } else if (code == RGBNM_OP_BRIGHTNESS) {
                if (comp == 0) v = rint_magic(v + s0);
            } else if (code == RGBNM_OP_CONTRAST) {
                if (comp == 0) v = rint_magic(v * op.f);
            } else if (code == RGBNM_OP_COLOR) {
                if (comp != 0) v = rint_magic(v * op.f);
            } else if (code == RGBNM_OP_AUTOCONTRAST || code == RGBNM_OP_AUTOSATURATION) {
                const bool mine = (code == RGBNM_OP_AUTOCONTRAST) ? (comp == 0) : (comp != 0);
                if (mine && s0 != s1) {
                    const float z = __fdiv_rn(v - s0, s1 - s0);
                    v = rint_magic(CLAMP_LO + z * (CLAMP_HI - CLAMP_LO));
                }
            } else if (code == RGBNM_OP_POSTERIZE) {
                v = float(tb.posterize_lut[op.p[0] * 2048 + int(v) + 1024]);
            } else if (code == RGBNM_OP_SHARPNESS || code == RGBNM_OP_MIDFREQ) {
                if (comp == 0) v = rint_magic(clampf(v * __ldg(tb.filters + op.p[0] * 64)));
            } else if (code == RGBNM_OP_SOLARIZE_ADD) {
                if (comp == 0 && v < 0.0f) v += float(op.p[0]);
            } else if (code == RGBNM_OP_INVERT) {
                v = -v;
            } else if (code == RGBNM_OP_EQUALIZE) {
                if (comp == 0) v = float(eq[int(v) + 1024]);
            } else if (code == RGBNM_OP_SOLARIZE) {
                if (eq[comp == 0 ? e : (2 * r) * GRID_Y + 2 * c] != 0) v = -v;
            }
            dst[e] = clampf(v);
        }
        __syncthreads();
        cur ^= 1;
    }
}

}  // namespace k0

extern "C" int rgbnm_k0_dcstats_ex(const int16_t* y, const int16_t* cbcr, const int16_t* quant, const rgbnm_plan* plans,
                                   const rgbnm_k0_tables* tables, float* stats, int n, int hb, int wb, int layout, void* stream) {
    using namespace k0;
    if (!y || !cbcr || !quant || !plans || !tables || !stats || n < 0) return RGBNM_ERR_ARG;
    if (hb < 2 || wb < 2 || hb > 255 || wb > 255 || (hb & 1) || (wb & 1)) return RGBNM_ERR_ARG;
    if (layout == RGBNM_K0_LAYOUT_VIT16_NOSUB) layout = RGBNM_K0_LAYOUT_VIT16;      // same planes, only the embedding tail differs
    if (layout != RGBNM_K0_LAYOUT_VIT16 && layout != RGBNM_K0_LAYOUT_SWIN4) return RGBNM_ERR_ARG;
    if (n == 0) return RGBNM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (layout == RGBNM_K0_LAYOUT_VIT16)
        k0_dcstats_kernel<RGBNM_K0_LAYOUT_VIT16><<<n, STATS_THREADS, 0, st>>>(y, cbcr, quant, plans, *tables, stats, hb, wb);
    else
        k0_dcstats_kernel<RGBNM_K0_LAYOUT_SWIN4><<<n, STATS_THREADS, 0, st>>>(y, cbcr, quant, plans, *tables, stats, hb, wb);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

extern "C" int rgbnm_k0_dcstats(const int16_t* y, const int16_t* cbcr, const int16_t* quant, const rgbnm_plan* plans,
                                const rgbnm_k0_tables* tables, float* stats, int n, int hb, int wb, void* stream) {
    return rgbnm_k0_dcstats_ex(y, cbcr, quant, plans, tables, stats, n, hb, wb, RGBNM_K0_LAYOUT_VIT16, stream);
}
