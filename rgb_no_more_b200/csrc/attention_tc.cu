// Attention core of the DCT ViT on tcgen05 tensor cores (sm_100a):
//     O = softmax(Q K^T / sqrt(emb_size)) V        per (image, head)     models/plainvit.py:450-461
// (the reference divides by sqrt(emb_size), not sqrt(head_dim): plainvit.py:455-457).
//
// Input  qkv [B][N=196][3*H*64] bf16 = [q | k | v], head-major (rows of the fused projection weight are
//        regrouped once per step, rgbnm_weight_prep);  output o [B][N][H*64] bf16 ('b n (h d)', :461) and
//        lse [B][H][N] fp32 (log-sum-exp of the scaled scores, kept for backward).
//
// Forward: one work item = one (image, head).  The 196 keys (padded to 208 = one UMMA N) are loaded once and
// shared by the two 128-query tiles, each owned by one softmax warpgroup:
//   TMA        Q0, Q1 [128 x 64], K, V [208 x 64] into a 2-stage ring: the next item's operands land while
//              this one computes (3-D tensor map over qkv; rows past token 195 are zero-filled)
//   MMA 1      S_w[128 x 208] = Q_w K^T         UMMA 128x208x16 x 4, fp32 in TMEM columns w*256 + [0, 208)
//   softmax    one thread per query row, two TMEM passes (max, then exp2 + row sum), P as packed bf16 back
//              into TMEM columns [0, 104) of the same slot (behind the read pointer)
//   MMA 2      O_w[128 x 64] = P_w V            A operand from TMEM, V as MN-major smem operand, fp32 in
//              TMEM columns w*256 + [128, 192)
//   epilogue   O / rowsum -> bf16 -> swizzled staging (the dead Q_w tile) -> TMA store (rows past 195 clipped)
// The two warpgroups run half an item apart, so one group's exponentials (the MUFU unit is the bound: 128 x 208
// ex2 per tile at 16/clk/SM) overlap the other group's MMAs, epilogue and barrier latencies.
// Warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = softmax / epilogue (warpgroup (warp-2)/4,
// TMEM lane quadrant warp % 4).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/rgbnm_b200.h"
#include "common.cuh"
#include "sm100.cuh"

int rgbnm_make_tmap_bf16_3d(CUtensorMap* map, const void* ptr, long long d0, long long d1, long long d2, long long ld1,
                            long long ld2, int box0, int box1);

namespace attn {
using namespace sm100;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^x on the FMA pipe (x <= 0 after the max subtraction): round-to-nearest split x = n + f, |f| <= 1/2, a cubic for
// 2^f (|rel err| < 1.1e-4, below the bf16 rounding of P) and n added into the exponent field.  Used for one element
// in four so the 16-lane MUFU unit and the FMA pipes finish together.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.0f);
    const float r = x + 12582912.0f;                  // 1.5 * 2^23: the integer part lands in the low mantissa bits
    const float f = x - (r - 12582912.0f);
    float p = fmaf(f, 0.0550086908f, 0.2422106266f);
    p = fmaf(p, f, 0.6932829022f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

constexpr int BM = 128;          // queries per tile
constexpr int NK = 208;          // keys padded to a multiple of 16 (UMMA N granularity at M = 128)
constexpr int HD = 64;           // head dim
constexpr int Q_BYTES = BM * HD * 2;      // 16384
constexpr int KV_BYTES = NK * HD * 2;     // 26624
constexpr int KV_SLOT = 27648;            // 1024-aligned

// ================================================= forward ==========================================================
constexpr int F_THREADS = 10 * 32;
constexpr int F_STAGE = 2 * Q_BYTES + 2 * KV_SLOT;          // Q0 | Q1 | K | V  = 88064
constexpr int F_OFF_BAR = 2 * F_STAGE;
constexpr int F_SMEM_TOTAL = F_OFF_BAR + 128 + 16 + 1024;
constexpr int F_TMEM_COLS = 512;
constexpr int COL_S = 0, COL_P = 0, COL_O = 128;            // inside a warpgroup's 256-column slot

struct Params {
    int B, N, H;
    int items;            // B * H
    float scale_log2e;    // scale * log2(e)
    float scale;
    float* lse;
};

__global__ void __launch_bounds__(F_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                const __grid_constant__ CUtensorMap tmO, const Params p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F_OFF_BAR);
    uint64_t* full_qk = bars + 0;        // [2] stage: Q0, Q1, K landed
    uint64_t* full_v = bars + 2;         // [2] stage: V landed
    uint64_t* stage_free = bars + 4;     // [2] stage: both warpgroups have stored their O tile (count 2)
    uint64_t* s_full = bars + 6;         // [2] warpgroup: S accumulator complete
    uint64_t* p_ready = bars + 8;        // [2] warpgroup: P written to TMEM (count 4 warps)
    uint64_t* o_full = bars + 10;        // [2] warpgroup: O accumulator complete
    uint64_t* o_drained = bars + 12;     // [2] warpgroup: O read out of TMEM (count 4 warps) -> slot reusable
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + F_OFF_BAR + 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmQ);
        prefetch_tensormap(&tmKV);
        prefetch_tensormap(&tmO);
        for (int s = 0; s < 2; ++s) {
            mbar_init(full_qk + s, 1); mbar_init(full_v + s, 1); mbar_init(stage_free + s, 2);
            mbar_init(s_full + s, 1); mbar_init(p_ready + s, 4); mbar_init(o_full + s, 1); mbar_init(o_drained + s, 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<F_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const int h = item % p.H, b = item / p.H;
                const int st = it & 1;
                unsigned char* sg = smem + st * F_STAGE;
                mbar_wait(stage_free + st, ((it >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(full_qk + st, 2 * Q_BYTES + KV_BYTES);
                tma_load_3d(sg, &tmQ, full_qk + st, h * HD, 0, b);
                tma_load_3d(sg + 2 * Q_BYTES, &tmKV, full_qk + st, (p.H + h) * HD, 0, b);
                tma_load_3d(sg + Q_BYTES, &tmQ, full_qk + st, h * HD, BM, b);
                mbar_arrive_expect_tx(full_v + st, KV_BYTES);
                tma_load_3d(sg + 2 * Q_BYTES + KV_SLOT, &tmKV, full_v + st, (2 * p.H + h) * HD, 0, b);
            }
        }
    } else if (warp == 1) {
        {
            // The whole warp walks the schedule (uniform control flow, descriptors in uniform registers); one elected lane
            // issues the tcgen05 instructions.  Issued from inside `if (lane == 0)`, each of the 13 small P.V UMMAs cost
            // ~200 cycles of single-thread ELECT / R2UR latency against the ~48 cycles it occupies the tensor pipe.
            constexpr uint32_t idesc_s = make_idesc_bf16(BM, NK, false, false);
            constexpr uint32_t idesc_o = make_idesc_bf16(BM, HD, false, true);      // B = V is MN-major (keys are the reduction)
            // Issue order  S0(0) S1(0) | PV0(i) S0(i+1) PV1(i) S1(i+1) | ...: a warpgroup gets its next score tile as soon
            // as its own O has left TMEM, without waiting for the other group's softmax, so the two groups drift half an
            // item apart and one group's exponentials cover the other group's MMA / barrier latencies.
            const int n_it = (p.items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
            const bool issuer = elect_one();
            const uint64_t dq0 = make_smem_desc_sw128(smem_u32(smem), 0, 1024);                           // stage 0: Q0
            const uint64_t dk0 = make_smem_desc_sw128(smem_u32(smem + 2 * Q_BYTES), 0, 1024);             // stage 0: K
            const uint64_t dv0 = make_smem_desc_sw128(smem_u32(smem + 2 * Q_BYTES + KV_SLOT), 0, 1024);   // stage 0: V
            auto issue_s = [&](int it, int w) {
                const int st = it & 1;
                if (w == 0) mbar_wait(full_qk + st, (it >> 1) & 1);
                mbar_wait(o_drained + w, (it & 1) ^ 1);              // the previous item's O_w has left this TMEM slot
                tc_fence_after();
                const uint64_t dq = dq0 + uint64_t(st) * (F_STAGE / 16) + uint64_t(w) * (Q_BYTES / 16);
                const uint64_t dk = dk0 + uint64_t(st) * (F_STAGE / 16);
                if (issuer) {
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) tc_mma_f16(tmem_base + w * 256 + COL_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                    tc_commit(s_full + w);
                }
                __syncwarp();
            };
            auto issue_pv = [&](int it, int w) {
                const int st = it & 1;
                if (w == 0) mbar_wait(full_v + st, (it >> 1) & 1);
                mbar_wait(p_ready + w, it & 1);
                tc_fence_after();
                const uint64_t dv = dv0 + uint64_t(st) * (F_STAGE / 16);
                if (issuer) {
#pragma unroll
                    for (int k = 0; k < NK / 16; ++k)
                        tc_mma_f16_ts(tmem_base + w * 256 + COL_O, tmem_base + w * 256 + COL_P + k * 8, dv + k * (2048 / 16), idesc_o, k != 0);
                    tc_commit(o_full + w);
                }
                __syncwarp();
            };
            if (n_it > 0) { issue_s(0, 0); issue_s(0, 1); }
            for (int it = 0; it < n_it; ++it) {
                issue_pv(it, 0);
                if (it + 1 < n_it) issue_s(it + 1, 0);
                issue_pv(it, 1);
                if (it + 1 < n_it) issue_s(it + 1, 1);
            }
        }
    } else {
        const int w = (warp - 2) >> 2;              // warpgroup = query tile
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int qrow = w * BM + row;
        const bool warp_live = (w * BM + quad * 32) < p.N;       // tile 1, rows 224..255: nothing to compute
        const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16) + w * 256;
        const bool leader = (quad == 0 && lane == 0);
        const int bar_id = 1 + w;
        int it = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const int h = item % p.H, b = item / p.H;
            const int st = it & 1;
            const uint32_t ph = it & 1;
            unsigned char* stg = smem + st * F_STAGE + w * Q_BYTES;     // Q_w: dead once S_w is complete
            mbar_wait(s_full + w, ph);
            tc_fence_after();
            float sum = 1.0f, mx = 0.0f;
            if (warp_live) {
                // ---- pass 1: row maximum over the 196 real keys -----------------------------------------
                mx = -INFINITY;
#pragma unroll 1
                for (int c = 0; c < 6; c += 2) {
                    uint32_t r0[32], r1[32];
                    tmem_ld32(lane_addr + COL_S + c * 32, r0);
                    tmem_ld32(lane_addr + COL_S + c * 32 + 32, r1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaxf(__uint_as_float(r0[j]), __uint_as_float(r1[j])));
                }
                {
                    uint32_t r[16];
                    tmem_ld16(lane_addr + COL_S + 192, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));       // keys 192..195
                }
                const float moff = mx * p.scale_log2e;
                // ---- pass 2: p = exp2(s * scale*log2e - max*scale*log2e), row sum, P (bf16) back into TMEM ----
                sum = 0.0f;
                uint32_t r[32];
                tmem_ld32(lane_addr + COL_S, r);
#pragma unroll 1
                for (int c = 0; c < 6; ++c) {
                    tmem_ld_wait();
                    float x[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = fmaf(__uint_as_float(r[j]), p.scale_log2e, -moff);
                    // next chunk's TMEM load overlaps this chunk's math (r is dead after the fma above)
                    if (c < 5) tmem_ld32(lane_addr + COL_S + (c + 1) * 32, r);
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float e0 = ex2_approx(x[2 * j]);
                        const float e1 = (j & 1) ? ex2_poly(x[2 * j + 1]) : ex2_approx(x[2 * j + 1]);
                        sum += e0 + e1;
                        pk[j] = pack_bf16(e0, e1);
                    }
                    tmem_st16(lane_addr + COL_P + c * 16, pk);
                }
                {
                    uint32_t t[16];
                    tmem_ld16(lane_addr + COL_S + 192, t);
                    tmem_ld_wait();
                    // keys 192..195 are real, 196..207 are padding: probability 0 (their V rows are zero-filled as well)
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = 0u;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float e0 = ex2_approx(fmaf(__uint_as_float(t[2 * j]), p.scale_log2e, -moff));
                        const float e1 = ex2_approx(fmaf(__uint_as_float(t[2 * j + 1]), p.scale_log2e, -moff));
                        sum += e0 + e1;
                        pk[j] = pack_bf16(e0, e1);
                    }
                    // P columns 96..103 hold keys 192..207; the x16 store also zeroes columns 104..111, which are free
                    tmem_st16(lane_addr + COL_P + 96, pk);
                }
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready + w);
            if (qrow < p.N) p.lse[(size_t(b) * p.H + h) * p.N + qrow] = mx * p.scale + __logf(sum);
            // ---- epilogue: O / sum -> bf16 -> staging (the Q_w tile, dead since MMA 1) -> TMA store ----
            mbar_wait(o_full + w, ph);
            tc_fence_after();
            if (warp_live) {
                const float inv = 1.0f / sum;
                unsigned char* rowp = stg + row * 128;
                uint32_t r0[32], r1[32];
                tmem_ld32(lane_addr + COL_O, r0);
                tmem_ld32(lane_addr + COL_O + 32, r1);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const int q = c * 4 + q4;
                        const float* x = reinterpret_cast<const float*>(c == 0 ? r0 : r1) + 8 * q4;
                        *reinterpret_cast<uint4*>(rowp + ((q ^ (row & 7)) << 4)) =
                            make_uint4(pack_bf16(x[0] * inv, x[1] * inv), pack_bf16(x[2] * inv, x[3] * inv),
                                       pack_bf16(x[4] * inv, x[5] * inv), pack_bf16(x[6] * inv, x[7] * inv));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_drained + w);
            fence_proxy_async();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (leader) {
                tma_store_3d(&tmO, stg, h * HD, w * BM, b);
                tma_store_commit();
                tma_store_wait_read<0>();
                mbar_arrive(stage_free + st);
            }
        }
        if (leader) tma_store_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<F_TMEM_COLS>(tmem_base);
}


// =================================================================================================
// Backward.  With P = softmax(S), S = Q K^T * scale, O = P V, Dr = rowsum(dO * O):
//     dV = P^T dO        dP = dO V^T        dS = P * (dP - Dr) * scale        dQ = dS K        dK = dS^T Q
// P is recomputed from S and the saved log-sum-exp.  Two kernels on one skeleton (score MMAs -> element-wise pass
// with one thread per accumulator row -> bf16 operands parked in TMEM -> output MMAs):
//   attn_bwd_kernel<false>  item = 128 queries x all keys:   S, dP as [q x key]; dS (TMEM A operand) . K -> dQ
//   attn_bwd_kernel<true>   item = 128 keys x all queries:   S^T = K Q^T, dP^T = V dO^T as [key x q];
//                           P^T . dO -> dV,  dS^T . Q -> dK   (both A operands straight from TMEM)
// The transposed formulation costs a second evaluation of S and dP, and in exchange no operand ever has to be
// transposed through shared memory and no partial dQ has to be reduced across CTAs.
// Schedule: operands of the next item are TMA-loaded into the second smem stage while this item computes; the
// element-wise pass is split over 8 warps (TMEM lane quadrant x column half).  Column half 0 owns score columns
// 0..95 and 192..207, half 1 owns columns 96..191; each parks its bf16 results behind its own read pointer, so the
// two warps of a quadrant never touch each other's columns.  The score MMAs are issued as two column groups and the
// output MMAs per half, so half 0 starts (and its part of the reduction is consumed) while half 1 is still busy;
// the first output accumulator lives in the 96 spare TMEM columns and can be written while scores are still read.
// =================================================================================================
constexpr int B_THREADS = 10 * 32;
constexpr int COL_DP = 208;                 // second score accumulator
constexpr int COL_DS = 208;                 // bf16 dS / dS^T of column half 0, aliasing consumed dP columns
constexpr int COL_P1 = 96, COL_DS1 = 304;   // column half 1 parks behind its own read pointers (S cols 96.., dP cols 304..)
constexpr int COL_OUT = 416;                // first output accumulator (dQ / dV): the 96 spare columns, aliases nothing
constexpr int COL_OUT2 = 352;               // second output accumulator (dK): dP columns that are consumed by then
constexpr int BWD_TMEM_COLS = 512;
constexpr int B_STAGE = 2 * Q_BYTES + 2 * KV_SLOT;           // A0 | A1 | B0 | B1 = 88064
constexpr int NKV = 224;                                     // per-query vector padded so the 32-wide tail chunk stays in bounds
constexpr int B_OFF_VEC = 2 * B_STAGE;                       // 2 x float2[NKV]: per-query {lse * log2e, Dr * scale}
constexpr int B_OFF_BAR = B_OFF_VEC + 2 * NKV * 8;
constexpr int B_SMEM_TOTAL = B_OFF_BAR + 128 + 16 + 1024;

struct BwdParams {
    int B, N, H, items, tiles;
    float scale, scale_log2e;
    const float* lse;     // [B][H][N]
    const float* dvec;    // [B][H][N]  rowsum(dO * O)
};

// Dr[b][h][q] = sum_d dO[b][q][h*64+d] * O[b][q][h*64+d]
// 8 lanes share one (token, head) segment of 128 bytes, 16 bytes each: consecutive threads read consecutive 16-byte
// chunks of dO / O (fully coalesced), then three shuffles reduce the partial dot products.
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O, float* __restrict__ dvec,
                     int B, int N, int H) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // 16-byte chunk index
    const long long seg = idx >> 3;                                                // (b*N + q) * H + h
    const bool live = seg < (long long)B * N * H;
    float s = 0.0f;
    if (live) {
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(dO) + idx), y = __ldg(reinterpret_cast<const uint4*>(O) + idx);
        const unsigned xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            s += __uint_as_float(xw[e] << 16) * __uint_as_float(yw[e] << 16);
            s += __uint_as_float(xw[e] & 0xffff0000u) * __uint_as_float(yw[e] & 0xffff0000u);
        }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (live && (idx & 7) == 0) {
        const int h = int(seg % H);
        const long long row = seg / H;
        const int q = int(row % N), b = int(row / N);
        dvec[((long long)b * H + h) * N + q] = s;
    }
}

// TMEM column of the bf16 A operand for reduction step k (16 score columns = 8 packed columns per step)
__device__ __forceinline__ uint32_t bwd_a_col(int base_half0, int base_half1, int k) {
    return k < 6 ? base_half0 + 8 * k : (k < 12 ? base_half1 + 8 * (k - 6) : base_half0 + 48);
}

// KV = false: rows are queries (A0 = Q tile, A1 = dO tile, B0 = K, B1 = V), one output dQ = dS . K.
// KV = true: rows are keys (A0 = K tile, A1 = V tile, B0 = Q, B1 = dO), two outputs dV = P^T . dO (into the V tile's
// staging) and dK = dS^T . Q.
template <bool KV>
__global__ void __launch_bounds__(B_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmTile, const __grid_constant__ CUtensorMap tmAll,
                const __grid_constant__ CUtensorMap tmTileDO, const __grid_constant__ CUtensorMap tmAllDO,
                const __grid_constant__ CUtensorMap tmOut, const BwdParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF_BAR);
    uint64_t* full = bars + 0;          // [2] stage: all four operand tiles landed
    uint64_t* stage_free = bars + 2;    // [2] stage: outputs stored, smem reusable
    uint64_t* s_full = bars + 4;        // [2] score accumulators complete: columns [0, 96) | [96, 208)
    uint64_t* p_ready = bars + 6;       // [3] bf16 operands parked in TMEM (count 4 warps each): half 0 chunks | half 1 chunks | tail
    uint64_t* o_full = bars + 9;        // output accumulator(s) complete
    uint64_t* acc_free = bars + 10;     // output accumulators read out (count 8 warps) -> TMEM reusable
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + B_OFF_BAR + 128);
    float2* vec = reinterpret_cast<float2*>(smem + B_OFF_VEC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmTile); prefetch_tensormap(&tmAll); prefetch_tensormap(&tmTileDO);
        prefetch_tensormap(&tmAllDO); prefetch_tensormap(&tmOut);
        for (int s = 0; s < 2; ++s) { mbar_init(full + s, 1); mbar_init(stage_free + s, 1); }
        mbar_init(s_full, 1); mbar_init(s_full + 1, 1); mbar_init(o_full, 1); mbar_init(acc_free, 8);
        for (int s = 0; s < 3; ++s) mbar_init(p_ready + s, 4);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<BWD_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const int t = item % p.tiles, bh = item / p.tiles;
                const int h = bh % p.H, b = bh / p.H;
                const int st = it & 1;
                unsigned char* sg = smem + st * B_STAGE;
                mbar_wait(stage_free + st, ((it >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(full + st, 2 * Q_BYTES + 2 * KV_BYTES);
                if (!KV) {
                    tma_load_3d(sg, &tmTile, full + st, h * HD, t * BM, b);                                   // Q tile
                    tma_load_3d(sg + 2 * Q_BYTES, &tmAll, full + st, (p.H + h) * HD, 0, b);                   // K
                    tma_load_3d(sg + Q_BYTES, &tmTileDO, full + st, h * HD, t * BM, b);                       // dO tile
                    tma_load_3d(sg + 2 * Q_BYTES + KV_SLOT, &tmAll, full + st, (2 * p.H + h) * HD, 0, b);     // V
                } else {
                    tma_load_3d(sg, &tmTile, full + st, (p.H + h) * HD, t * BM, b);                           // K tile
                    tma_load_3d(sg + 2 * Q_BYTES, &tmAll, full + st, h * HD, 0, b);                           // Q
                    tma_load_3d(sg + Q_BYTES, &tmTile, full + st, (2 * p.H + h) * HD, t * BM, b);             // V tile
                    tma_load_3d(sg + 2 * Q_BYTES + KV_SLOT, &tmAllDO, full + st, h * HD, 0, b);               // dO
                }
            }
        }
    } else if (warp == 1) {
        {
            // whole warp in uniform control flow, one elected lane issues (see the forward kernel)
            constexpr uint32_t idesc_o = make_idesc_bf16(BM, HD, false, true);
            constexpr uint32_t idesc_sa = make_idesc_bf16(BM, 96, false, false), idesc_sb = make_idesc_bf16(BM, NK - 96, false, false);
            const bool issuer = elect_one();
            const uint64_t d_a0 = make_smem_desc_sw128(smem_u32(smem), 0, 1024);                          // stage 0 tiles
            const uint64_t d_a1 = make_smem_desc_sw128(smem_u32(smem + Q_BYTES), 0, 1024);
            const uint64_t d_b0 = make_smem_desc_sw128(smem_u32(smem + 2 * Q_BYTES), 0, 1024);
            const uint64_t d_b1 = make_smem_desc_sw128(smem_u32(smem + 2 * Q_BYTES + KV_SLOT), 0, 1024);
            int it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const int st = it & 1;
                const uint32_t ph = it & 1;
                const uint64_t so = uint64_t(st) * (B_STAGE / 16);
                const uint64_t a0 = d_a0 + so, a1 = d_a1 + so, b0 = d_b0 + so, b1 = d_b1 + so;
                mbar_wait(full + st, (it >> 1) & 1);
                // KV: the second output accumulator aliases dP columns -> the previous item's outputs must have been read out.
                // !KV: nothing aliases; the score MMAs simply queue behind the previous item's output MMAs.
                if (KV) mbar_wait(acc_free, ph ^ 1);
                tc_fence_after();
                // scores:  !KV: S = Q K^T, dP = dO V^T      KV: S^T = K Q^T, dP^T = V dO^T.  Two column groups ([0, 96) then
                // [96, 208)) so the element-wise warps of half 0 start one MMA group earlier.
                if (issuer) {
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const uint32_t idesc_g = g == 0 ? idesc_sa : idesc_sb;
                        const uint64_t boff = uint64_t(g) * (96 * 128 / 16);       // B rows (K-major, 128 B each)
                        const uint32_t coff = g * 96;                              // accumulator columns
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k)
                            tc_mma_f16(tmem_base + COL_S + coff, a0 + 2 * k, b0 + boff + 2 * k, idesc_g, k != 0);
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k)
                            tc_mma_f16(tmem_base + COL_DP + coff, a1 + 2 * k, b1 + boff + 2 * k, idesc_g, k != 0);
                        tc_commit(s_full + g);
                    }
                }
                __syncwarp();
                // first output (dQ / dV) into the spare columns: reduction steps 0..5 as soon as half 0 has parked them
                if (!KV) mbar_wait(acc_free, ph ^ 1);          // the previous item's dQ has been read out of the spare columns
                mbar_wait(p_ready + 0, ph);
                tc_fence_after();
                const uint64_t bo = KV ? b1 : b0;                // reduction-side operand of the first output: dO (KV) / K
                constexpr int a0col = KV ? COL_P : COL_DS, a1col = KV ? COL_P1 : COL_DS1;
                if (issuer) {
#pragma unroll
                    for (int k = 0; k < 6; ++k)
                        tc_mma_f16_ts(tmem_base + COL_OUT, tmem_base + bwd_a_col(a0col, a1col, k), bo + k * (2048 / 16), idesc_o, k != 0);
                }
                __syncwarp();
                mbar_wait(p_ready + 1, ph);
                mbar_wait(p_ready + 2, ph);
                tc_fence_after();
                if (issuer) {
#pragma unroll
                    for (int k = 6; k < NK / 16; ++k)
                        tc_mma_f16_ts(tmem_base + COL_OUT, tmem_base + bwd_a_col(a0col, a1col, k), bo + k * (2048 / 16), idesc_o, true);
                    if (KV) {
                        // dK = dS^T . Q: its accumulator aliases dP columns, so only once every score column has been consumed
#pragma unroll
                        for (int k = 0; k < NK / 16; ++k)
                            tc_mma_f16_ts(tmem_base + COL_OUT2, tmem_base + bwd_a_col(COL_DS, COL_DS1, k), b0 + k * (2048 / 16), idesc_o, k != 0);
                    }
                    tc_commit(o_full);
                }
                __syncwarp();
            }
        }
    } else {
        const int quad = warp & 3;
        const int halfc = (warp - 2) >> 2;          // column half of the element-wise pass / of the 64 output columns
        const int row = quad * 32 + lane;
        const int tid = threadIdx.x - 64;
        const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
        const bool leader = (warp == 2 && lane == 0);
        // per-row (!KV) / per-query (KV) softmax statistics {lse * log2e, Dr * scale}: fetched one item ahead, so the global
        // load latency (it was ~10 % of the warps' time when loaded at the top of the item) hides behind the previous item
        auto fetch_stats = [&](int item) -> float2 {
            if (item >= p.items) return make_float2(0.0f, 0.0f);
            const int t = item % p.tiles, bh = item / p.tiles;
            const float* lse_bh = p.lse + size_t(bh) * p.N;
            const float* d_bh = p.dvec + size_t(bh) * p.N;
            if (!KV) {
                const int q = t * BM + row;
                return q < p.N ? make_float2(__ldg(lse_bh + q), __ldg(d_bh + q)) : make_float2(0.0f, 0.0f);
            }
            // padded queries: P = 0 (lse = +inf), dS = 0
            return tid < p.N ? make_float2(__ldg(lse_bh + tid), __ldg(d_bh + tid)) : make_float2(INFINITY, 0.0f);
        };      // raw values: the first arithmetic on them (= the scoreboard wait) happens one item later
        float2 next_stats = fetch_stats(blockIdx.x);
        int it = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const int t = item % p.tiles, bh = item / p.tiles;
            const int h = bh % p.H, b = bh / p.H;
            const int st = it & 1;
            const uint32_t ph = it & 1;
            const bool warp_live = (t * BM + quad * 32) < p.N;      // rows past the last token: outputs are clipped
            const float2 stats = make_float2(next_stats.x * 1.4426950408889634f, next_stats.y * p.scale);
            const float my_lse2 = stats.x, my_dds = stats.y;
            const float2* vq = vec + (it & 1) * NKV;
            if (KV) {
                float2* vw = vec + (it & 1) * NKV;
                if (tid < NKV) vw[tid] = stats;
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            next_stats = fetch_stats(item + int(gridDim.x));
            // chunks of 32 score columns: half 0 -> chunks 0,1,2 and the 16-column tail; half 1 -> chunks 3,4,5
            const uint32_t vq_s = smem_u32(vq);
#pragma unroll 1
            for (int ci = 0; ci < 3 + (halfc == 0 ? 1 : 0); ++ci) {
                const bool tail = (ci == 3);
                const int c = tail ? 6 : halfc * 3 + ci;
                if (ci == 0) { mbar_wait(s_full + halfc, ph); tc_fence_after(); }
                if (tail) {
                    // chunks 0..2 of this half are parked: the MMA warp may start on their reduction steps
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(p_ready + 0);
                    mbar_wait(s_full + 1, ph);                    // the tail columns belong to the second score group
                    tc_fence_after();
                }
                if (!warp_live) continue;
                uint32_t rs[32], rd[32];
                if (!tail) {
                    tmem_ld32(lane_addr + COL_S + c * 32, rs);
                    tmem_ld32(lane_addr + COL_DP + c * 32, rd);
                    tmem_ld_wait();
                } else {
                    uint32_t ta[16], tb[16];
                    tmem_ld16(lane_addr + COL_S + 192, ta);
                    tmem_ld16(lane_addr + COL_DP + 192, tb);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) { rs[j] = ta[j]; rd[j] = tb[j]; rs[16 + j] = 0u; rd[16 + j] = 0u; }
                }
                uint32_t pk[16], dk[16];
                const uint32_t vaddr = vq_s + c * 256;                      // 32 columns x float2
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float l2a = my_lse2, dda = my_dds, l2b = my_lse2, ddb = my_dds;
                    if (KV) {
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(l2a), "=f"(dda), "=f"(l2b), "=f"(ddb) : "r"(vaddr + j * 16));
                    }
                    const float xa = fmaf(__uint_as_float(rs[2 * j]), p.scale_log2e, -l2a);
                    const float xb = fmaf(__uint_as_float(rs[2 * j + 1]), p.scale_log2e, -l2b);
                    const float pa = ex2_approx(xa);
                    const float pb = (j & 1) ? ex2_poly(xb) : ex2_approx(xb);
                    const float da = pa * fmaf(__uint_as_float(rd[2 * j]), p.scale, -dda);
                    const float db = pb * fmaf(__uint_as_float(rd[2 * j + 1]), p.scale, -ddb);
                    pk[j] = pack_bf16(pa, pb);
                    dk[j] = pack_bf16(da, db);
                }
                // parked behind this half's own read pointer (tail: columns 48..63 of half 0's region, read long ago; its
                // x16 store spills zeros/garbage into 8 columns nobody reads)
                const uint32_t pcol = (halfc == 0 ? COL_P : COL_P1) + 16 * ci;
                const uint32_t dcol = (halfc == 0 ? COL_DS : COL_DS1) + 16 * ci;
                if (KV) tmem_st16(lane_addr + pcol, pk);
                tmem_st16(lane_addr + dcol, dk);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready + (halfc == 0 ? 2 : 1));
            // ---- epilogue: accumulators -> bf16 -> swizzled staging (dead A tiles) -> TMA store; this warp converts
            //      columns [32 * halfc, +32) of each output ----
            mbar_wait(o_full, ph);
            tc_fence_after();
            unsigned char* sg = smem + st * B_STAGE;
            if (warp_live) {
#pragma unroll
                for (int o = 0; o < (KV ? 2 : 1); ++o) {
                    // !KV: dQ -> staging A0 (Q tile).   KV: o = 0: dV (COL_O) -> staging A1 (V tile); o = 1: dK (COL_OUT2) -> A0 (K tile)
                    const uint32_t col0 = (o == 0) ? COL_OUT : COL_OUT2;
                    unsigned char* rowp = sg + ((KV && o == 0) ? Q_BYTES : 0) + row * 128;
                    uint32_t r[32];
                    tmem_ld32(lane_addr + col0 + halfc * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const int q = halfc * 4 + q4;
                        const float* x = reinterpret_cast<const float*>(r) + 8 * q4;
                        *reinterpret_cast<uint4*>(rowp + ((q ^ (row & 7)) << 4)) =
                            make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free);
            fence_proxy_async();
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (leader) {
                if (!KV) {
                    tma_store_3d(&tmOut, sg, h * HD, t * BM, b);                                       // dQ
                } else {
                    tma_store_3d(&tmOut, sg + Q_BYTES, (2 * p.H + h) * HD, t * BM, b);                 // dV
                    tma_store_3d(&tmOut, sg, (p.H + h) * HD, t * BM, b);                               // dK
                }
                tma_store_commit();
                tma_store_wait_read<0>();
                mbar_arrive(stage_free + st);
            }
        }
        if (leader) tma_store_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<BWD_TMEM_COLS>(tmem_base);
}

}  // namespace attn

extern "C" int rgbnm_attention_fwd(const void* qkv, void* o, float* lse, int B, int N, int H, int D, float scale, void* stream) {
    using namespace attn;
    if (!qkv || !o || !lse || B <= 0 || H <= 0) return RGBNM_ERR_ARG;
    if (D != HD || N != 196) return RGBNM_ERR_UNSUPPORTED;       // 14 x 14 tokens, head_size 64 (ViT-Ti/S/B)
    static bool configured = false;
    static int num_sms = 0;
    if (!configured) {
        RGBNM_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_TOTAL));
        int dev = 0;
        RGBNM_CUDA_CHECK(cudaGetDevice(&dev));
        RGBNM_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        configured = true;
    }
    const long long ldq = 3LL * H * HD;
    CUtensorMap tmQ, tmKV, tmO;
    int rc;
    if ((rc = rgbnm_make_tmap_bf16_3d(&tmQ, qkv, ldq, N, B, ldq, ldq * N, HD, BM))) return rc;
    if ((rc = rgbnm_make_tmap_bf16_3d(&tmKV, qkv, ldq, N, B, ldq, ldq * N, HD, NK))) return rc;
    if ((rc = rgbnm_make_tmap_bf16_3d(&tmO, o, (long long)H * HD, N, B, (long long)H * HD, (long long)H * HD * N, HD, BM))) return rc;
    Params p;
    p.B = B; p.N = N; p.H = H;
    p.items = B * H;
    p.scale = scale;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.lse = lse;
    const int grid = p.items < num_sms ? p.items : num_sms;
    attn_fwd_kernel<<<grid, F_THREADS, F_SMEM_TOTAL, static_cast<cudaStream_t>(stream)>>>(tmQ, tmKV, tmO, p);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

extern "C" int rgbnm_attention_bwd(const void* dout, const void* qkv, const void* o, const float* lse, void* dqkv, float* dvec,
                                   int B, int N, int H, int D, float scale, void* stream) {
    using namespace attn;
    if (!dout || !qkv || !o || !lse || !dqkv || !dvec || B <= 0 || H <= 0) return RGBNM_ERR_ARG;
    if (D != HD || N != 196) return RGBNM_ERR_UNSUPPORTED;
    static bool configured = false;
    static int num_sms = 0;
    if (!configured) {
        RGBNM_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM_TOTAL));
        RGBNM_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM_TOTAL));
        int dev = 0;
        RGBNM_CUDA_CHECK(cudaGetDevice(&dev));
        RGBNM_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        configured = true;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long ldq = 3LL * H * HD, ldo = (long long)H * HD;
    {
        const long long total = (long long)B * N * H * 8;          // 16-byte chunks: 8 per (token, head)
        attn_bwd_prep_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dout),
                                                                            static_cast<const __nv_bfloat16*>(o), dvec, B, N, H);
        RGBNM_CUDA_CHECK(cudaGetLastError());
    }
    CUtensorMap tmTile, tmAll, tmTileDO, tmAllDO, tmOut;
    int rc;
    if ((rc = rgbnm_make_tmap_bf16_3d(&tmTile, qkv, ldq, N, B, ldq, ldq * N, HD, BM))) return rc;
    if ((rc = rgbnm_make_tmap_bf16_3d(&tmAll, qkv, ldq, N, B, ldq, ldq * N, HD, NK))) return rc;
    if ((rc = rgbnm_make_tmap_bf16_3d(&tmTileDO, dout, ldo, N, B, ldo, ldo * N, HD, BM))) return rc;
    if ((rc = rgbnm_make_tmap_bf16_3d(&tmAllDO, dout, ldo, N, B, ldo, ldo * N, HD, NK))) return rc;
    if ((rc = rgbnm_make_tmap_bf16_3d(&tmOut, dqkv, ldq, N, B, ldq, ldq * N, HD, BM))) return rc;
    BwdParams p;
    p.B = B; p.N = N; p.H = H;
    p.tiles = (N + BM - 1) / BM;
    p.items = B * H * p.tiles;
    p.scale = scale;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.lse = lse;
    p.dvec = dvec;
    const int grid = p.items < num_sms ? p.items : num_sms;
    attn_bwd_kernel<false><<<grid, B_THREADS, B_SMEM_TOTAL, st>>>(tmTile, tmAll, tmTileDO, tmAllDO, tmOut, p);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    attn_bwd_kernel<true><<<grid, B_THREADS, B_SMEM_TOTAL, st>>>(tmTile, tmAll, tmTileDO, tmAllDO, tmOut, p);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}
