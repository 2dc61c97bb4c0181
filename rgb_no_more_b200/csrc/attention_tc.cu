// Attention core of the DCT ViT on tcgen05 tensor cores (sm_100a), forward:
//     O = softmax(Q K^T / sqrt(emb_size)) V        per (image, head)     models/plainvit.py:450-461
// (the reference divides by sqrt(emb_size), not sqrt(head_dim): plainvit.py:455-457).
//
// Input  qkv [B][N=196][3*H*64] bf16 = [q | k | v], head-major (rows of the fused projection weight are
//        regrouped once per step, rgbnm_weight_prep);  output o [B][N][H*64] bf16 ('b n (h d)', :461) and
//        lse [B][H][N] fp32 (log-sum-exp of the scaled scores, kept for backward).
//
// One work item = one 128-query tile of one (image, head): the whole key row (196 keys, padded to 208) fits one
// UMMA N, so the softmax is single-pass.  Per item:
//   TMA        Q tile [128 x 64], K [208 x 64], V [208 x 64] (3-D tensor map over qkv; rows past token 195 are
//              zero-filled by the TMA unit)
//   MMA 1      S[128 x 208] = Q K^T            UMMA 128x208x16 x 4, fp32 in TMEM columns [0, 208)
//   softmax    one thread per query row: tcgen05.ld S, max / exp2 / sum in registers, P as packed bf16 back
//              into TMEM columns [0, 104) (aliasing the S columns already consumed)
//   MMA 2      O[128 x 64] = P V               A operand from TMEM (tcgen05.mma ..., [tmem_a], ...), V as
//              MN-major smem operand, fp32 in TMEM columns [128, 192)
//   epilogue   O / rowsum -> bf16 -> swizzled staging (the dead Q tile) -> TMA store (rows past 195 clipped)
// Persistent CTAs, 256 TMEM columns each, 2 CTAs per SM so one CTA's softmax overlaps the other's MMAs.
// Warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = softmax / epilogue (TMEM lane quadrant = warp % 4).
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

constexpr int BM = 128;          // queries per tile
constexpr int NK = 208;          // keys padded to a multiple of 16 (UMMA N granularity at M = 128)
constexpr int HD = 64;           // head dim
constexpr int THREADS = 6 * 32;
constexpr int TMEM_COLS = 256;
constexpr int COL_S = 0, COL_P = 0, COL_O = 128;

constexpr int Q_BYTES = BM * HD * 2;      // 16384
constexpr int KV_BYTES = NK * HD * 2;     // 26624
constexpr int OFF_Q = 0, OFF_K = OFF_Q + Q_BYTES, OFF_V = OFF_K + 27648 /* 1024-aligned */, OFF_BAR = OFF_V + 27648;
constexpr int SMEM_TOTAL = OFF_BAR + 64 + 16 + 1024;

struct Params {
    int B, N, H;
    int items;            // B * H * tiles
    int tiles;            // ceil(N / 128)
    float scale_log2e;    // scale * log2(e)
    float scale;
    float* lse;
};

__global__ void __launch_bounds__(THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                const __grid_constant__ CUtensorMap tmO, const Params p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* full_qk = bars + 0;     // Q + K landed
    uint64_t* full_v = bars + 1;      // V landed
    uint64_t* s_full = bars + 2;      // S accumulator complete
    uint64_t* p_ready = bars + 3;     // P written to TMEM by all 128 softmax threads
    uint64_t* o_full = bars + 4;      // O accumulator complete
    uint64_t* item_done = bars + 5;   // smem + TMEM free for the next item
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 64);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmQ);
        prefetch_tensormap(&tmKV);
        prefetch_tensormap(&tmO);
        mbar_init(full_qk, 1);
        mbar_init(full_v, 1);
        mbar_init(s_full, 1);
        mbar_init(p_ready, 4);
        mbar_init(o_full, 1);
        mbar_init(item_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
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
                mbar_wait(item_done, (it & 1) ^ 1);
                mbar_arrive_expect_tx(full_qk, Q_BYTES + KV_BYTES);
                tma_load_3d(smem + OFF_Q, &tmQ, full_qk, h * HD, t * BM, b);
                tma_load_3d(smem + OFF_K, &tmKV, full_qk, (p.H + h) * HD, 0, b);
                mbar_arrive_expect_tx(full_v, KV_BYTES);
                tma_load_3d(smem + OFF_V, &tmKV, full_v, (2 * p.H + h) * HD, 0, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(BM, NK, false, false);
            constexpr uint32_t idesc_o = make_idesc_bf16(BM, HD, false, true);      // B = V is MN-major (keys are the reduction)
            const uint32_t sq = smem_u32(smem + OFF_Q), sk = smem_u32(smem + OFF_K), sv = smem_u32(smem + OFF_V);
            int it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                mbar_wait(full_qk, ph);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    tc_mma_f16(tmem_base + COL_S, make_smem_desc_sw128(sq + k * 32, 0, 1024),
                               make_smem_desc_sw128(sk + k * 32, 0, 1024), idesc_s, k != 0);
                tc_commit(s_full);
                mbar_wait(full_v, ph);
                mbar_wait(p_ready, ph);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < NK / 16; ++k)
                    tc_mma_f16_ts(tmem_base + COL_O, tmem_base + COL_P + k * 8,
                                  make_smem_desc_sw128(sv + k * 2048, 0, 1024), idesc_o, k != 0);
                tc_commit(o_full);
            }
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
        const bool leader = (warp == 2 && lane == 0);
        int it = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const int t = item % p.tiles, bh = item / p.tiles;
            const int h = bh % p.H, b = bh / p.H;
            const uint32_t ph = it & 1;
            mbar_wait(s_full, ph);
            tc_fence_after();
            // ---- pass 1: row maximum over the 196 real keys -----------------------------------------
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 6; ++c) {
                uint32_t r[32];
                tmem_ld32(lane_addr + COL_S + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
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
            float sum = 0.0f;
#pragma unroll 1
            for (int c = 0; c < 6; ++c) {
                uint32_t r[32];
                tmem_ld32(lane_addr + COL_S + c * 32, r);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float e0 = ex2_approx(fmaf(__uint_as_float(r[2 * j]), p.scale_log2e, -moff));
                    const float e1 = ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), p.scale_log2e, -moff));
                    // the sum uses the bf16-rounded probabilities, i.e. exactly what the P.V MMA sees
                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(e0, e1);
                    sum += __bfloat162float(b2.x) + __bfloat162float(b2.y);
                    pk[j] = *reinterpret_cast<const uint32_t*>(&b2);
                }
                tmem_st16(lane_addr + COL_P + c * 16, pk);
            }
            {
                uint32_t r[16];
                tmem_ld16(lane_addr + COL_S + 192, r);
                tmem_ld_wait();
                // keys 192..195 are real, 196..207 are padding: probability 0 (their V rows are zero-filled as well)
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[j] = 0u;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float e0 = ex2_approx(fmaf(__uint_as_float(r[2 * j]), p.scale_log2e, -moff));
                    const float e1 = ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), p.scale_log2e, -moff));
                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(e0, e1);
                    sum += __bfloat162float(b2.x) + __bfloat162float(b2.y);
                    pk[j] = *reinterpret_cast<const uint32_t*>(&b2);
                }
                // P columns 96..103 hold keys 192..207; the x16 store also zeroes columns 104..111, which are free
                tmem_st16(lane_addr + COL_P + 96, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);
            const int qrow = t * BM + row;
            if (qrow < p.N) p.lse[(size_t(b) * p.H + h) * p.N + qrow] = mx * p.scale + __logf(sum);
            // ---- epilogue: O / sum -> bf16 -> staging (the Q tile, dead since MMA 1) -> TMA store ----
            mbar_wait(o_full, ph);
            tc_fence_after();
            const float inv = 1.0f / sum;
            unsigned char* rowp = smem + OFF_Q + row * 128;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld32(lane_addr + COL_O + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int q = c * 4 + q4;
                    const float* x = reinterpret_cast<const float*>(r) + 8 * q4;
                    *reinterpret_cast<uint4*>(rowp + ((q ^ (row & 7)) << 4)) =
                        make_uint4(pack_bf16(x[0] * inv, x[1] * inv), pack_bf16(x[2] * inv, x[3] * inv),
                                   pack_bf16(x[4] * inv, x[5] * inv), pack_bf16(x[6] * inv, x[7] * inv));
                }
            }
            tc_fence_before();
            fence_proxy_async();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (leader) {
                tma_store_3d(&tmO, smem + OFF_Q, h * HD, t * BM, b);
                tma_store_commit();
                tma_store_wait_read<0>();
                mbar_arrive(item_done);
            }
        }
        if (leader) tma_store_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}


// =================================================================================================
// Backward.  With P = softmax(S), S = Q K^T * scale, O = P V, Dr = rowsum(dO * O):
//     dV = P^T dO        dP = dO V^T        dS = P * (dP - Dr) * scale        dQ = dS K        dK = dS^T Q
// P is recomputed from S and the saved log-sum-exp.  Two kernels, each a copy of the forward schedule
// (score MMAs -> one thread per accumulator row -> bf16 operand parked in TMEM -> second MMA):
//   attn_bwd_q_kernel   item = 128 queries x all keys:   S, dP as [q x key]; dS (TMEM A operand) . K -> dQ
//   attn_bwd_kv_kernel  item = 128 keys x all queries:   S^T = K Q^T, dP^T = V dO^T as [key x q];
//                       P^T . dO -> dV,  dS^T . Q -> dK   (both A operands straight from TMEM)
// The transposed formulation costs a second evaluation of S and dP, and in exchange no operand ever has to be
// transposed through shared memory and no partial dQ has to be reduced across CTAs.
// =================================================================================================
constexpr int COL_DP = 208;                 // second score accumulator
constexpr int COL_DS = 208;                 // bf16 dS / dS^T, aliasing the consumed dP columns
constexpr int COL_OUT2 = 336;               // second output accumulator (dK), inside the dP region
constexpr int BWD_TMEM_COLS = 512;
constexpr int B_OFF_A0 = 0, B_OFF_A1 = 16384, B_OFF_B0 = 32768, B_OFF_B1 = B_OFF_B0 + 27648, B_OFF_VEC = B_OFF_B1 + 27648;
constexpr int B_OFF_BAR = B_OFF_VEC + 2 * NK * 4;
constexpr int B_SMEM_TOTAL = B_OFF_BAR + 64 + 16 + 1024;

struct BwdParams {
    int B, N, H, items, tiles;
    float scale, scale_log2e;
    const float* lse;     // [B][H][N]
    const float* dvec;    // [B][H][N]  rowsum(dO * O)
};

// Dr[b][h][q] = sum_d dO[b][q][h*64+d] * O[b][q][h*64+d]
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O, float* __restrict__ dvec,
                     int B, int N, int H) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (b*N + q) * H + h
    if (idx >= (long long)B * N * H) return;
    const int h = int(idx % H);
    const long long row = idx / H;
    const int q = int(row % N), b = int(row / N);
    const uint4* a = reinterpret_cast<const uint4*>(dO + row * (long long)(H * HD) + h * HD);
    const uint4* c = reinterpret_cast<const uint4*>(O + row * (long long)(H * HD) + h * HD);
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint4 x = __ldg(a + k), y = __ldg(c + k);
        const unsigned xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            s += __uint_as_float(xw[e] << 16) * __uint_as_float(yw[e] << 16);
            s += __uint_as_float(xw[e] & 0xffff0000u) * __uint_as_float(yw[e] & 0xffff0000u);
        }
    }
    dvec[((long long)b * H + h) * N + q] = s;
}

// Shared skeleton of the two backward kernels.  KV = false: rows are queries (A0 = Q tile, A1 = dO tile, B0 = K, B1 = V),
// one output dQ = dS . K.  KV = true: rows are keys (A0 = K tile, A1 = V tile, B0 = Q, B1 = dO), two outputs
// dV = P^T . dO (into the V tile's staging) and dK = dS^T . Q.
template <bool KV>
__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmTile, const __grid_constant__ CUtensorMap tmAll,
                const __grid_constant__ CUtensorMap tmTileDO, const __grid_constant__ CUtensorMap tmAllDO,
                const __grid_constant__ CUtensorMap tmOut, const BwdParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF_BAR);
    uint64_t* full = bars + 0;        // all four operand tiles landed
    uint64_t* s_full = bars + 1;      // both score accumulators complete
    uint64_t* p_ready = bars + 2;     // bf16 operands parked in TMEM
    uint64_t* o_full = bars + 3;      // output accumulator(s) complete
    uint64_t* item_done = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + B_OFF_BAR + 64);
    float* vec_lse = reinterpret_cast<float*>(smem + B_OFF_VEC);          // KV: per-query lse * log2e (+inf past the last query)
    float* vec_d = vec_lse + NK;                                            // KV: per-query Dr

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmTile); prefetch_tensormap(&tmAll); prefetch_tensormap(&tmTileDO);
        prefetch_tensormap(&tmAllDO); prefetch_tensormap(&tmOut);
        mbar_init(full, 1); mbar_init(s_full, 1); mbar_init(p_ready, 4); mbar_init(o_full, 1); mbar_init(item_done, 1);
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
                mbar_wait(item_done, (it & 1) ^ 1);
                mbar_arrive_expect_tx(full, 2 * Q_BYTES + 2 * KV_BYTES);
                if (!KV) {
                    tma_load_3d(smem + B_OFF_A0, &tmTile, full, h * HD, t * BM, b);                   // Q tile
                    tma_load_3d(smem + B_OFF_A1, &tmTileDO, full, h * HD, t * BM, b);                 // dO tile
                    tma_load_3d(smem + B_OFF_B0, &tmAll, full, (p.H + h) * HD, 0, b);                 // K
                    tma_load_3d(smem + B_OFF_B1, &tmAll, full, (2 * p.H + h) * HD, 0, b);             // V
                } else {
                    tma_load_3d(smem + B_OFF_A0, &tmTile, full, (p.H + h) * HD, t * BM, b);           // K tile
                    tma_load_3d(smem + B_OFF_A1, &tmTile, full, (2 * p.H + h) * HD, t * BM, b);       // V tile
                    tma_load_3d(smem + B_OFF_B0, &tmAll, full, h * HD, 0, b);                         // Q
                    tma_load_3d(smem + B_OFF_B1, &tmAllDO, full, h * HD, 0, b);                       // dO
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(BM, NK, false, false);
            constexpr uint32_t idesc_o = make_idesc_bf16(BM, HD, false, true);
            const uint32_t a0 = smem_u32(smem + B_OFF_A0), a1 = smem_u32(smem + B_OFF_A1);
            const uint32_t b0 = smem_u32(smem + B_OFF_B0), b1 = smem_u32(smem + B_OFF_B1);
            int it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                mbar_wait(full, ph);
                tc_fence_after();
                // scores:  !KV: S = Q K^T, dP = dO V^T      KV: S^T = K Q^T, dP^T = V dO^T
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    tc_mma_f16(tmem_base + COL_S, make_smem_desc_sw128(a0 + k * 32, 0, 1024),
                               make_smem_desc_sw128(b0 + k * 32, 0, 1024), idesc_s, k != 0);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    tc_mma_f16(tmem_base + COL_DP, make_smem_desc_sw128(a1 + k * 32, 0, 1024),
                               make_smem_desc_sw128(b1 + k * 32, 0, 1024), idesc_s, k != 0);
                tc_commit(s_full);
                mbar_wait(p_ready, ph);
                tc_fence_after();
                if (!KV) {
                    // dQ = dS . K   (K as MN-major operand: keys are the reduction)
#pragma unroll
                    for (int k = 0; k < NK / 16; ++k)
                        tc_mma_f16_ts(tmem_base + COL_O, tmem_base + COL_DS + k * 8,
                                      make_smem_desc_sw128(b0 + k * 2048, 0, 1024), idesc_o, k != 0);
                } else {
                    // dV = P^T . dO,  dK = dS^T . Q
#pragma unroll
                    for (int k = 0; k < NK / 16; ++k)
                        tc_mma_f16_ts(tmem_base + COL_O, tmem_base + COL_P + k * 8,
                                      make_smem_desc_sw128(b1 + k * 2048, 0, 1024), idesc_o, k != 0);
#pragma unroll
                    for (int k = 0; k < NK / 16; ++k)
                        tc_mma_f16_ts(tmem_base + COL_OUT2, tmem_base + COL_DS + k * 8,
                                      make_smem_desc_sw128(b0 + k * 2048, 0, 1024), idesc_o, k != 0);
                }
                tc_commit(o_full);
            }
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int tid = threadIdx.x - 64;
        const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
        const bool leader = (warp == 2 && lane == 0);
        int it = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const int t = item % p.tiles, bh = item / p.tiles;
            const int h = bh % p.H, b = bh / p.H;
            const uint32_t ph = it & 1;
            const float* lse_bh = p.lse + (size_t(b) * p.H + h) * p.N;
            const float* d_bh = p.dvec + (size_t(b) * p.H + h) * p.N;
            float my_lse2 = 0.0f, my_d = 0.0f;
            if (!KV) {
                const int q = t * BM + row;
                if (q < p.N) { my_lse2 = lse_bh[q] * 1.4426950408889634f; my_d = d_bh[q]; }
            } else {
                for (int q = tid; q < NK; q += 128) {
                    vec_lse[q] = q < p.N ? lse_bh[q] * 1.4426950408889634f : INFINITY;     // padded queries: P = 0
                    vec_d[q] = q < p.N ? d_bh[q] : 0.0f;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait(s_full, ph);
            tc_fence_after();
            // one pass over the 208 columns: P and dS as packed bf16, written behind the read pointer
#pragma unroll 1
            for (int c = 0; c < 7; ++c) {
                uint32_t rs[32], rd[32];
                if (c < 6) {
                    tmem_ld32(lane_addr + COL_S + c * 32, rs);
                    tmem_ld32(lane_addr + COL_DP + c * 32, rd);
                } else {
                    uint32_t t16a[16], t16b[16];
                    tmem_ld16(lane_addr + COL_S + 192, t16a);
                    tmem_ld16(lane_addr + COL_DP + 192, t16b);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) { rs[j] = t16a[j]; rd[j] = t16b[j]; rs[16 + j] = 0u; rd[16 + j] = 0u; }
                }
                tmem_ld_wait();
                uint32_t pk[16], dk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float pv[2], dv[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int col = c * 32 + 2 * j + e;
                        float l2, dd;
                        if (!KV) { l2 = my_lse2; dd = my_d; }
                        else { l2 = vec_lse[col < NK ? col : NK - 1]; dd = vec_d[col < NK ? col : NK - 1]; }
                        const float pr = ex2_approx(fmaf(__uint_as_float(rs[2 * j + e]), p.scale_log2e, -l2));
                        pv[e] = pr;
                        dv[e] = pr * (__uint_as_float(rd[2 * j + e]) - dd) * p.scale;
                    }
                    pk[j] = pack_bf16(pv[0], pv[1]);
                    dk[j] = pack_bf16(dv[0], dv[1]);
                }
                if (c < 6) {
                    if (KV) tmem_st16(lane_addr + COL_P + c * 16, pk);
                    tmem_st16(lane_addr + COL_DS + c * 16, dk);
                } else {
                    // last 16 columns -> 8 packed columns; the x16 store spills zeros/garbage into 8 free columns
                    if (KV) tmem_st16(lane_addr + COL_P + 96, pk);
                    tmem_st16(lane_addr + COL_DS + 96, dk);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);
            // ---- epilogue: accumulators -> bf16 -> swizzled staging (dead A tiles) -> TMA store ----
            mbar_wait(o_full, ph);
            tc_fence_after();
#pragma unroll
            for (int o = 0; o < (KV ? 2 : 1); ++o) {
                // !KV: dQ -> staging A0 (Q tile).   KV: o = 0: dV (COL_O) -> staging A1 (V tile); o = 1: dK (COL_OUT2) -> A0 (K tile)
                const uint32_t col0 = (o == 0) ? COL_O : COL_OUT2;
                unsigned char* stg = smem + ((KV && o == 0) ? B_OFF_A1 : B_OFF_A0);
                unsigned char* rowp = stg + row * 128;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t r[32];
                    tmem_ld32(lane_addr + col0 + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const int q = c * 4 + q4;
                        const float* x = reinterpret_cast<const float*>(r) + 8 * q4;
                        *reinterpret_cast<uint4*>(rowp + ((q ^ (row & 7)) << 4)) =
                            make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (leader) {
                if (!KV) {
                    tma_store_3d(&tmOut, smem + B_OFF_A0, h * HD, t * BM, b);                          // dQ
                } else {
                    tma_store_3d(&tmOut, smem + B_OFF_A1, (2 * p.H + h) * HD, t * BM, b);              // dV
                    tma_store_3d(&tmOut, smem + B_OFF_A0, (p.H + h) * HD, t * BM, b);                  // dK
                }
                tma_store_commit();
                tma_store_wait_read<0>();
                mbar_arrive(item_done);
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
        RGBNM_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
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
    p.tiles = (N + BM - 1) / BM;
    p.items = B * H * p.tiles;
    p.scale = scale;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.lse = lse;
    const int grid = p.items < 2 * num_sms ? p.items : 2 * num_sms;
    attn_fwd_kernel<<<grid, THREADS, SMEM_TOTAL, static_cast<cudaStream_t>(stream)>>>(tmQ, tmKV, tmO, p);
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
        const long long total = (long long)B * N * H;
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
    attn_bwd_kernel<false><<<grid, THREADS, B_SMEM_TOTAL, st>>>(tmTile, tmAll, tmTileDO, tmAllDO, tmOut, p);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    attn_bwd_kernel<true><<<grid, THREADS, B_SMEM_TOTAL, st>>>(tmTile, tmAll, tmTileDO, tmAllDO, tmOut, p);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}
