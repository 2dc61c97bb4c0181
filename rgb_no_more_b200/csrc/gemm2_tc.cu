// Dense bf16 GEMMs of the DCT ViT, CTA-pair edition (sm_100a): tcgen05.mma.cta_group::2 over a
// 2-CTA cluster.  The round-1 single-CTA kernel (tools/legacy/gemm_tc_v1.cu, no longer built) re-read 128 x K of A and 192 x K of B from
// L2 for every 128 x 192 tile (76.8 FLOP per L2 byte) and is capped by the L2 -> SM fabric at
// ~0.95 PFLOP/s on this part (measured: 46.6 us for the 44.4 GFLOP qkv projection).  Here two SMs
// share one 256 x BN accumulator tile: each CTA loads its own 128 rows of A and only HALF of the B
// tile, the UMMA reads both halves across the pair, so the L2 traffic per FLOP drops by
// (128 + BN) / (128 + BN/2) -- 112 FLOP/B at BN = 192, 128 FLOP/B at BN = 256.
//
//   forward / dgrad   C[M,N] = A[M,K] . B[N,K]^T      K-major operands  (nn.Linear, plainvit.py:194,441-443,485-490)
//   wgrad             out[M,N] += A[T,M]^T . B[T,N]   MN-major operands, split over the tokens T,
//                                                     fp32 red.add into the flat gradient buffer
//
// Per CTA: warp 0 = TMA producer, warp 1 = TMEM owner (+ MMA issuer in the leader CTA), warp 2 =
// epilogue I/O (TMA stores of finished 128 x 64 sub-tiles, TMA loads of residual / pre-activation
// sub-tiles), warps 3..10 = epilogue math (TMEM lane quadrant = warp % 4, column half = (warp-3)/4).
// Pipelines: smem ring (TMA <-> MMA, `full` lives in the leader CTA and counts both CTAs' bytes),
// two TMEM accumulators (MMA <-> epilogue; `tempty` collects the epilogue warps of both CTAs),
// staging ring of 128 x 64 bf16 sub-tiles (epilogue math <-> epilogue I/O).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "../../include/rgbnm_b200.h"
#include "common.cuh"
#include "sm100.cuh"

int rgbnm_make_tmap_bf16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows);

namespace gemm2 {
using namespace sm100;

constexpr int BM_CTA = 128, BM = 256, BK = 64;
constexpr int EPI_WARPS = 8;
constexpr int FIRST_EPI_WARP = 3;
constexpr int THREADS = (FIRST_EPI_WARP + EPI_WARPS) * 32;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;      // clears the CTA-rank bit of a shared::cluster address -> rank 0's copy

enum Epi : int { EPI_STORE = 0, EPI_RESIDUAL = 1, EPI_GELU = 2, EPI_DGELU = 3, EPI_POSEMB = 4, EPI_ATOMIC = 5, EPI_F32 = 6,
                 EPI_GELU_ACT = 7, EPI_LNRES = 8, EPI_LN = 9 };

struct Params {
    int M, N;
    int m_tiles, n_tiles;     // tiles of 256 x BN
    int k_blocks;
    int splits;
    const float* bias;
    const float* posemb;
    int pos_period;
    float* out_f32;
    long long ldo;
    float alpha;
    int trans_out;            // atomic epilogue: out[n * ldo + m] instead of out[m * ldo + n]
    int vec_ok;               // atomic epilogue: 16-byte aligned rows -> red.global.add.v4.f32
    int perm_heads, perm_hd;  // atomic epilogue: M index is in kernel qkv order (q|k|v head-major) -> reference row h*3D + d*3 + which
    const float* ln_gamma;    // EPI_LNRES: C = aux + LayerNorm(acc + bias) * gamma + beta over the N columns of a row
    const float* ln_beta;
    float ln_eps;
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- PTX pieces that only the CTA-pair kernel needs ------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are counted on the LEADER CTA's mbarrier (issued by both CTAs of the pair)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t tmem_addr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {      // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(uint16_t(3)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(uint32_t(accumulate))
        : "memory");
}
__device__ __forceinline__ void red_add_f32(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- GELU(erf) (nn.GELU() default, plainvit.py:487) with ONE MUFU op per element ------------------------
// Phi(x) = (1 + erf(x / sqrt 2)) / 2 is evaluated as (1 + tanh(p(x))) / 2 with the odd quintic
// p(x) = x (C0 + C1 x^2 + C2 x^4) fitted to atanh(erf(x / sqrt 2)): |gelu - gelu_erf| <= 2.6e-5 and
// |gelu' - gelu_erf'| <= 1.1e-4 before the MUFU.TANH error (2^-11 relative), i.e. well inside the bf16
// rounding of the stored activation.  x^2 is clamped at 64 (|x| = 8) where tanh has saturated; the
// quintic would turn over beyond |x| ~ 11.  The previous form (Abramowitz-Stegun 7.1.26) needed
// MUFU.RCP + MUFU.EX2 per element: 16 issue-equivalent cycles per warp on a 16-lane/clk/SM unit, more
// than the whole tensor-core time of a K = 384 tile.
constexpr float GC0 = 7.97507857e-01f, GC1 = 3.70056692e-02f, GC2 = -3.51520268e-04f;
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gelu_fast(float x) {
    const float x2 = fminf(x * x, 64.0f);
    const float q = fmaf(x2, fmaf(x2, GC2, GC1), GC0);
    const float t = tanh_approx(x * q);
    const float h = 0.5f * x;
    return fmaf(h, t, h);
}
__device__ __forceinline__ float dgelu_fast(float x) {
    const float x2 = fminf(x * x, 64.0f);
    const float q = fmaf(x2, fmaf(x2, GC2, GC1), GC0);
    const float hdp = fmaf(x2, fmaf(x2, 2.5f * GC2, 1.5f * GC1), 0.5f * GC0);      // p'(x) / 2
    const float t = tanh_approx(x * q);
    const float s = fmaf(-t, t, 1.0f);
    const float cdf = fmaf(0.5f, t, 0.5f);
    return fmaf(x * hdp, s, cdf);
}

template <int BN, int STAGES, int EPI, bool B_MN>
struct Cfg {
    static constexpr bool STAGED = (EPI != EPI_ATOMIC && EPI != EPI_F32);
    static constexpr bool HAS_AUX = (EPI == EPI_RESIDUAL || EPI == EPI_DGELU || EPI == EPI_LNRES);
    static constexpr int NOUT = EPI == EPI_GELU ? 2 : (STAGED ? 1 : 0);
    static constexpr int NSUB = BN / 64;                        // 64-column sub-tiles of the CTA's 128 x BN output
    static constexpr int NBUF = !STAGED ? 0 : (NOUT == 2 ? 3 : 4);
    static constexpr int SUB_BYTES = BM_CTA * 128;              // one 128 x 64 bf16 sub-tile
    static constexpr int SLOT_BYTES = NOUT * SUB_BYTES;
    static constexpr int A_BYTES = BM_CTA * BK * 2;             // this CTA's half of the 256-row A tile
    static constexpr int B_BYTES = (BN / 2) * BK * 2;           // this CTA's half of the B tile
    static constexpr int OFF_A = 0, OFF_B = OFF_A + STAGES * A_BYTES, OFF_STG = OFF_B + STAGES * B_BYTES;
    static constexpr int OFF_BAR = OFF_STG + NBUF * SLOT_BYTES;
    static constexpr int NBARS = 2 * STAGES + 4 + 2 * (NBUF > 0 ? NBUF : 1);
    static constexpr int OFF_TMEM = OFF_BAR + NBARS * 8;
    static constexpr int OFF_LN = OFF_TMEM + 16;                // EPI_LNRES: row-statistics exchange between the two column halves
    static constexpr int LN_BYTES = (EPI == EPI_LNRES || EPI == EPI_LN) ? 2 * 8 * 32 * 8 : 0;
    static constexpr int TOTAL = OFF_LN + LN_BYTES + 1024;
    static constexpr int NACC = 2 * BN <= 512 ? 2 : 1;          // TMEM accumulator buffers (BN = 384: one, the epilogue is not overlapped)
    static_assert(BN % 64 == 0 && BN <= 512 && (BN / 2) % 8 == 0, "tile shape");
    static_assert(BN <= 256 || BN == 384, "BN = 384: two UMMAs per step (N = 256 + 128)");
    static_assert(!B_MN || (BN / 2) % 64 == 0, "MN-major B: each CTA needs whole 64-wide swizzle atoms");
    static_assert(TOTAL <= 232448, "shared memory");
};

template <int BN, int STAGES, int EPI, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
             const __grid_constant__ CUtensorMap tmAux, const Params p) {
    using L = Cfg<BN, STAGES, EPI, B_MN>;
    constexpr bool STAGED = L::STAGED, HAS_AUX = L::HAS_AUX;
    constexpr int NSUB = L::NSUB, NBUF = L::NBUF > 0 ? L::NBUF : 1;

    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* full = bars;                          // [STAGES]  leader's copy is the live one
    uint64_t* empty = bars + STAGES;                // [STAGES]  per CTA, multicast commit
    uint64_t* tfull = bars + 2 * STAGES;            // [2]       per CTA, multicast commit
    uint64_t* tempty = bars + 2 * STAGES + 2;       // [2]       leader's copy, 2 x EPI_WARPS arrivals
    uint64_t* stg_ready = bars + 2 * STAGES + 4;    // [NBUF]    slot free (and aux sub-tile landed)
    uint64_t* stg_done = stg_ready + (NBUF > 0 ? NBUF : 1);   // [NBUF]  slot written by all epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int total_tiles = p.m_tiles * p.n_tiles * p.splits;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmA);
        prefetch_tensormap(&tmB);
        if (STAGED) prefetch_tensormap(&tmC);
        if (EPI == EPI_GELU) prefetch_tensormap(&tmC2);
        if (HAS_AUX) prefetch_tensormap(&tmAux);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 2 * EPI_WARPS); }
        for (int s = 0; s < L::NBUF; ++s) { mbar_init(stg_ready + s, 1); mbar_init(stg_done + s, EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();                  // barriers of BOTH CTAs initialised before any remote arrive / TMA completion
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto k_range = [&](int sp, int& k0, int& k1) {
        k0 = int((long long)p.k_blocks * sp / p.splits);
        k1 = int((long long)p.k_blocks * (sp + 1) / p.splits);
    };

    if (warp == 0) {
        // ===================================== TMA producer (both CTAs) =========================
        // whole warp in uniform control flow, one elected lane issues the TMA instructions (operands stay in uniform
        // registers; from inside `if (lane == 0)` every UTMALDG paid an ELECT / R2UR sequence of single-thread latency)
        {
            const bool issuer = elect_one();
            int stage = 0, phase = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
                const int sp = tile % p.splits, rest = tile / p.splits;
                const int n_blk = rest % p.n_tiles, m_blk = rest / p.n_tiles;
                const int m0 = m_blk * BM + int(rank) * BM_CTA;          // this CTA's rows of A
                const int n0 = n_blk * BN + int(rank) * (BN / 2);        // this CTA's half of B
                int k0, k1;
                k_range(sp, k0, k1);
                for (int kb = k0; kb < k1; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1);
                    unsigned char* sa = smem + L::OFF_A + stage * L::A_BYTES;
                    unsigned char* sb = smem + L::OFF_B + stage * L::B_BYTES;
                    if (issuer) {
                    if (rank == 0) mbar_arrive_expect_tx(full + stage, 2 * (L::A_BYTES + L::B_BYTES));
                    if (!A_MN) {
                        tma_load_2d_pair(sa, &tmA, full + stage, kb * BK, m0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM_CTA / 64; ++j) tma_load_2d_pair(sa + j * (BK * 128), &tmA, full + stage, m0 + j * 64, kb * BK);
                    }
                    if (!B_MN && BN <= 256) {
                        tma_load_2d_pair(sb, &tmB, full + stage, kb * BK, n0);
                    } else if (!B_MN) {
                        // BN = 384, K-major: rows {r*128 .. +128} and {256 + r*64 .. +64} of the B tile (see the MN-major case below)
                        const int nb = n_blk * BN;
                        tma_load_2d_pair(sb, &tmB, full + stage, kb * BK, nb + int(rank) * 128);
                        tma_load_2d_pair(sb + 64 * 128, &tmB, full + stage, kb * BK, nb + int(rank) * 128 + 64);
                        tma_load_2d_pair(sb + 128 * 128, &tmB, full + stage, kb * BK, nb + 256 + int(rank) * 64);
                    } else if (BN <= 256) {
#pragma unroll
                        for (int j = 0; j < BN / 128; ++j) tma_load_2d_pair(sb + j * (BK * 128), &tmB, full + stage, n0 + j * 64, kb * BK);
                    } else {
                        // BN = 384 runs as two UMMAs per step (N = 256, then N = 128); a cta_group::2 UMMA takes the first half
                        // of its N from CTA 0 and the second half from CTA 1, so CTA r holds columns {r*128 .. +128} and
                        // {256 + r*64 .. +64} of the tile: accumulator column == tile column.
                        const int nb = n_blk * BN;
                        tma_load_2d_pair(sb, &tmB, full + stage, nb + int(rank) * 128, kb * BK);
                        tma_load_2d_pair(sb + BK * 128, &tmB, full + stage, nb + int(rank) * 128 + 64, kb * BK);
                        tma_load_2d_pair(sb + 2 * BK * 128, &tmB, full + stage, nb + 256 + int(rank) * 64, kb * BK);
                    }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer (leader CTA only) =====================
        // The whole warp walks the loop (uniform control flow, operands in uniform registers) and only the tcgen05
        // instructions themselves are issued by one elected lane.  With the loop inside `if (lane == 0)` every descriptor
        // lived in a per-thread register and each UMMA cost an ELECT / R2UR / PLOP3 / BRA.U.ANY sequence of ~200 cycles
        // of single-thread latency -- more than the 96..130 cycles the UMMA itself keeps the tensor pipe busy.
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN <= 256 ? BN : 256, A_MN, B_MN);
            constexpr uint32_t idesc_hi = make_idesc_bf16(BM, 128, A_MN, B_MN);          // BN = 384: second UMMA of a step
            // descriptors of stage 0, K step 0; a stage adds its byte size >> 4, a K step 2 (K-major: 32 B) or 128 (MN-major: 2048 B)
            const uint64_t da0 = A_MN ? make_smem_desc_sw128(smem_u32(smem + L::OFF_A), BK * 128, 1024)
                                      : make_smem_desc_sw128(smem_u32(smem + L::OFF_A), 0, 1024);
            const uint64_t db0 = B_MN ? make_smem_desc_sw128(smem_u32(smem + L::OFF_B), BK * 128, 1024)
                                      : make_smem_desc_sw128(smem_u32(smem + L::OFF_B), 0, 1024);
            constexpr uint64_t A_KSTEP = A_MN ? 2048 / 16 : 32 / 16, B_KSTEP = B_MN ? 2048 / 16 : 32 / 16;
            constexpr uint64_t B_HI = B_MN ? (2 * BK * 128) / 16 : (128 * 128) / 16;      // BN = 384: third 64-wide block of the B tile
            const bool issuer = elect_one();
            int stage = 0, phase = 0, it = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
                const int sp = tile % p.splits;
                int k0, k1;
                k_range(sp, k0, k1);
                const int acc = L::NACC == 2 ? (it & 1) : 0;
                mbar_wait(tempty + acc, ((L::NACC == 2 ? (it >> 1) : it) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = k0; kb < k1; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint64_t da = da0 + uint64_t(stage) * (L::A_BYTES / 16);
                    const uint64_t db = db0 + uint64_t(stage) * (L::B_BYTES / 16);
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            tc_mma_f16_pair(d_tmem, da + k * A_KSTEP, db + k * B_KSTEP, idesc, (kb != k0) || (k != 0));
                            if (BN > 256)       // columns [256, 384)
                                tc_mma_f16_pair(d_tmem + 256, da + k * A_KSTEP, db + B_HI + k * B_KSTEP, idesc_hi, (kb != k0) || (k != 0));
                        }
                        tc_commit_pair(empty + stage);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (issuer) tc_commit_pair(tfull + acc);
                __syncwarp();
            }
        }
    } else if (warp == 2) {
        // ===================================== epilogue I/O =====================================
        // Sub-tile g (running count over this CTA's tiles) lives in slot g % NBUF.  Hand-out of a slot
        // = (TMA load of the aux sub-tile, completing on stg_ready) or a plain arrive.
        // (whole warp in uniform control flow, one elected lane issues the TMA / mbarrier instructions)
        if (STAGED) {
            const bool issuer = elect_one();
            int h_tile = cluster_id, h_s = 0, h_g = 0;                  // hand-out cursor
            auto hand_out = [&]() {
                if (h_tile >= total_tiles) return;
                const int slot = h_g % NBUF;
                if (issuer) {
                    if (HAS_AUX) {
                        const int rest = h_tile / p.splits;
                        const int n_blk = rest % p.n_tiles, m_blk = rest / p.n_tiles;
                        mbar_arrive_expect_tx(stg_ready + slot, L::SUB_BYTES);
                        tma_load_2d(smem + L::OFF_STG + slot * L::SLOT_BYTES, &tmAux, stg_ready + slot, n_blk * BN + h_s * 64,
                                    m_blk * BM + int(rank) * BM_CTA);
                    } else {
                        mbar_arrive(stg_ready + slot);
                    }
                }
                ++h_g;
                if (++h_s == NSUB) { h_s = 0; h_tile += n_clusters; }
            };
            for (int i = 0; i < NBUF; ++i) hand_out();
            int g = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
                const int rest = tile / p.splits;
                const int n_blk = rest % p.n_tiles, m_blk = rest / p.n_tiles;
                const int m0 = m_blk * BM + int(rank) * BM_CTA;
                for (int s = 0; s < NSUB; ++s, ++g) {
                    const int slot = g % NBUF;
                    mbar_wait(stg_done + slot, (g / NBUF) & 1);
                    unsigned char* src = smem + L::OFF_STG + slot * L::SLOT_BYTES;
                    if (issuer) {
                        tma_store_2d(&tmC, src, n_blk * BN + s * 64, m0);
                        if (EPI == EPI_GELU) tma_store_2d(&tmC2, src + L::SUB_BYTES, n_blk * BN + s * 64, m0);
                        tma_store_commit();
                        // the store of sub-tile g-1 has left its slot -> that slot serves sub-tile g-1+NBUF.  (Measured: letting
                        // NBUF-2 stores stay in flight instead, at the price of one slot less run-ahead for the math warps, is slower.)
                        if (g >= 1) tma_store_wait_read<1>();
                    }
                    __syncwarp();
                    if (g >= 1) hand_out();
                }
            }
            if (issuer) tma_store_wait<0>();
            __syncwarp();
        }
    } else {
        // ===================================== epilogue math ====================================
        const int ew = warp - FIRST_EPI_WARP;
        const int quad = warp & 3;               // TMEM lane quadrant this warp may read
        const int half = ew >> 2;                // 32-column half of each 64-column sub-tile
        const int row = quad * 32 + lane;        // row inside this CTA's 128-row half tile
        int it = 0, g = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
            const int rest = tile / p.splits;
            const int n_blk = rest % p.n_tiles, m_blk = rest / p.n_tiles;
            const int acc = L::NACC == 2 ? (it & 1) : 0;
            const int n0 = n_blk * BN, m0 = m_blk * BM + int(rank) * BM_CTA;
            mbar_wait(tfull + acc, (L::NACC == 2 ? (it >> 1) : it) & 1);
            tc_fence_after();
            // EPI_LNRES (post-norm residual, models/swinv2.py:302-306): the whole row (N <= BN columns) sits in this CTA's
            // accumulator, thread = row; its columns are split between this warp and the warp of the other column half.
            // Two extra passes over TMEM give the row mean and the centred second moment (exchanged through shared memory).
            float ln_mean = 0.0f, ln_rstd = 1.0f;
            if (EPI == EPI_LNRES || EPI == EPI_LN) {
                // one extra pass over TMEM: shifted sums (shift = the first value this thread sees) -> (mean, M2) of this warp's
                // columns; the two halves are merged with the parallel-variance formula (no E[x^2] - mean^2 cancellation).
                // Exchange slots are double-buffered by tile parity: a warp may be one tile ahead of its partner.
                float2* sc = reinterpret_cast<float2*>(smem + L::OFF_LN) + (it & 1) * (8 * 32);
                float shift = 0.0f, s1 = 0.0f, s2 = 0.0f;
                int cnt = 0;
#pragma unroll 1
                for (int s = 0; s < NSUB; ++s) {
                    const int c = 2 * s + half;
                    const int col0 = n0 + c * 32;
                    if (col0 + 32 <= p.N) {                             // warp-uniform
                        uint32_t r[32];
                        tmem_ld32(tmem_base + (uint32_t(quad * 32) << 16) + acc * BN + c * 32, r);
                        float4 b4[8];
                        const bool hb = p.bias != nullptr;
                        if (hb) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) b4[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
                        }
                        tmem_ld_wait();
                        float x[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(r[j]);
                        if (hb) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) { x[4 * j] += b4[j].x; x[4 * j + 1] += b4[j].y; x[4 * j + 2] += b4[j].z; x[4 * j + 3] += b4[j].w; }
                        }
                        if (cnt == 0) shift = x[0];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float d = x[j] - shift;
                            s1 += d;
                            s2 = fmaf(d, d, s2);
                        }
                        cnt += 32;
                    }
                }
                const float na = float(cnt);
                const float mean_a = cnt ? shift + s1 / na : 0.0f;
                const float m2_a = cnt ? s2 - s1 * s1 / na : 0.0f;
                sc[ew * 32 + lane] = make_float2(mean_a, m2_a);
                named_bar_sync(1 + quad, 64);                           // the two warps of this TMEM lane quadrant
                const float2 o = sc[(ew ^ 4) * 32 + lane];
                const float n_all = float(p.N), nb = n_all - na;
                const float delta = o.x - mean_a;
                ln_mean = (na * mean_a + nb * o.x) / n_all;
                const float m2 = m2_a + o.y + delta * delta * (na * nb / n_all);
                ln_rstd = rsqrtf(m2 / n_all + p.ln_eps);
            }
#pragma unroll 1
            for (int s = 0; s < NSUB; ++s, ++g) {
                const int c = 2 * s + half;              // 32-column chunk of the tile
                uint32_t r[32];
                tmem_ld32(tmem_base + (uint32_t(quad * 32) << 16) + acc * BN + c * 32, r);
                unsigned char* slotp = nullptr;
                if (STAGED) {
                    const int slot = g % NBUF;
                    slotp = smem + L::OFF_STG + slot * L::SLOT_BYTES;
                    mbar_wait(stg_ready + slot, (g / NBUF) & 1);
                }
                // bias for this chunk: warp-uniform addresses, issued before the TMEM wait so the loads overlap it
                const int col0 = n0 + c * 32;
                constexpr bool ADD_BIAS = (EPI != EPI_ATOMIC && EPI != EPI_DGELU);
                float4 b4[8];
                const bool bias_vec = ADD_BIAS && p.bias != nullptr && col0 + 32 <= p.N && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
                if (bias_vec) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) b4[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
                }
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (ADD_BIAS) {
                    if (bias_vec) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) { v[4 * j] += b4[j].x; v[4 * j + 1] += b4[j].y; v[4 * j + 2] += b4[j].z; v[4 * j + 3] += b4[j].w; }
                    } else if (p.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
                    }
                }
                if ((EPI == EPI_LNRES || EPI == EPI_LN) && col0 + 32 <= p.N) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + col0 + j));
                        const float4 e4 = __ldg(reinterpret_cast<const float4*>(p.ln_beta + col0 + j));
                        v[j] = fmaf((v[j] - ln_mean) * ln_rstd, g4.x, e4.x);
                        v[j + 1] = fmaf((v[j + 1] - ln_mean) * ln_rstd, g4.y, e4.y);
                        v[j + 2] = fmaf((v[j + 2] - ln_mean) * ln_rstd, g4.z, e4.z);
                        v[j + 3] = fmaf((v[j + 3] - ln_mean) * ln_rstd, g4.w, e4.w);
                    }
                }
                if (EPI == EPI_POSEMB && col0 + 32 <= p.N) {
                    const float* pe = p.posemb + size_t((m0 + row) % p.pos_period) * p.N + col0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 e4 = __ldg(reinterpret_cast<const float4*>(pe + j));
                        v[j] += e4.x; v[j + 1] += e4.y; v[j + 2] += e4.z; v[j + 3] += e4.w;
                    }
                }
                if (EPI == EPI_ATOMIC || EPI == EPI_F32) {
                    int gr = m0 + row;
                    const bool in_range = gr < p.M;
                    if (EPI == EPI_ATOMIC && p.perm_heads > 0 && in_range) {
                        // undo the qkv row regrouping of rgbnm_weight_prep: kernel row = which*(H*D) + h*D + d
                        const int hd_all = p.perm_heads * p.perm_hd;
                        const int which = gr / hd_all, rem = gr - which * hd_all;
                        const int h = rem / p.perm_hd, d = rem - h * p.perm_hd;
                        gr = h * (3 * p.perm_hd) + d * 3 + which;
                    }
                    if (in_range) {
                        if (EPI == EPI_F32) {
                            float* o = p.out_f32 + size_t(gr) * p.ldo + col0;
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) o[j] = v[j];
                        } else if (p.trans_out) {
                            // lanes = consecutive m: every instruction adds 32 consecutive floats of one output row
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < p.N) red_add_f32(p.out_f32 + size_t(col0 + j) * p.ldo + gr, p.alpha * v[j]);
                        } else if (p.vec_ok) {
                            float* o = p.out_f32 + size_t(gr) * p.ldo + col0;
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                if (col0 + j < p.N) red_add_v4(o + j, p.alpha * v[j], p.alpha * v[j + 1], p.alpha * v[j + 2], p.alpha * v[j + 3]);
                        } else {
                            float* o = p.out_f32 + size_t(gr) * p.ldo + col0;
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) red_add_f32(o + j, p.alpha * v[j]);
                        }
                    }
                } else {
                    // slot layout = SWIZZLE_128B box {64 cols, 128 rows}: row r at r*128 bytes, 16-byte chunk q at q ^ (r & 7)
                    unsigned char* rowp = slotp + row * 128;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int q = half * 4 + t;
                        uint4* cell = reinterpret_cast<uint4*>(rowp + ((q ^ (row & 7)) << 4));
                        float* x = v + 8 * t;
                        if (HAS_AUX) {
                            const uint4 a = *cell;
                            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float lo = __uint_as_float(aw[e] << 16), hi = __uint_as_float(aw[e] & 0xffff0000u);
                                if (EPI == EPI_RESIDUAL || EPI == EPI_LNRES) { x[2 * e] += lo; x[2 * e + 1] += hi; }
                                else { x[2 * e] *= dgelu_fast(lo); x[2 * e + 1] *= dgelu_fast(hi); }
                            }
                        }
                        if (EPI == EPI_GELU_ACT) {             // inference: only the activation is stored
#pragma unroll
                            for (int e = 0; e < 8; ++e) x[e] = gelu_fast(x[e]);
                        }
                        *cell = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
                        if (EPI == EPI_GELU) {
                            uint4* cell2 = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(cell) + L::SUB_BYTES);
                            *cell2 = make_uint4(pack_bf16(gelu_fast(x[0]), gelu_fast(x[1])), pack_bf16(gelu_fast(x[2]), gelu_fast(x[3])),
                                                pack_bf16(gelu_fast(x[4]), gelu_fast(x[5])), pack_bf16(gelu_fast(x[6]), gelu_fast(x[7])));
                        }
                    }
                    fence_proxy_async();                   // generic-proxy writes -> visible to the TMA engine
                    __syncwarp();
                    if (lane == 0) mbar_arrive(stg_done + (g % NBUF));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(tempty + acc);   // accumulator `acc` of this CTA drained by this warp
        }
    }
    tc_fence_before();
    cluster_sync_all();                  // the peer may still read this CTA's smem / signal its barriers until here
    if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
static int g_num_sms = 0;

static int pick_splits(int tiles, int k_blocks, int clusters) {
    // fill whole waves of `clusters` work items; fewer splits = fewer atomics, so take the smallest
    // split count within 5 % of the best wave efficiency
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 40 && s <= k_blocks / 8; ++s) {
        const int items = tiles * s;
        const int waves = (items + clusters - 1) / clusters;
        const double eff = double(items) / (double(waves) * clusters);
        if (eff > best_eff + 0.05) { best_eff = eff; best = s; }
    }
    return best;
}

template <int BN, int STAGES, int EPI, bool A_MN, bool B_MN>
static int launch(const rgbnm_gemm_args& a, cudaStream_t st) {
    using L = Cfg<BN, STAGES, EPI, B_MN>;
    static bool configured = false;
    auto kfn = gemm2_kernel<BN, STAGES, EPI, A_MN, B_MN>;
    if (!configured) {
        RGBNM_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        configured = true;
    }
    if (g_num_sms == 0) {
        int dev = 0;
        RGBNM_CUDA_CHECK(cudaGetDevice(&dev));
        RGBNM_CUDA_CHECK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    CUtensorMap tmA, tmB, tmC, tmC2, tmAux;
    int rc;
    if (!A_MN) rc = rgbnm_make_tmap_bf16(&tmA, a.A, a.M, a.K, a.lda, BK, BM_CTA);
    else rc = rgbnm_make_tmap_bf16(&tmA, a.A, a.K, a.M, a.lda, 64, BK);
    if (rc) return rc;
    if (!B_MN) rc = rgbnm_make_tmap_bf16(&tmB, a.B, a.N, a.K, a.ldb, BK, BN <= 256 ? BN / 2 : 64);
    else rc = rgbnm_make_tmap_bf16(&tmB, a.B, a.K, a.N, a.ldb, 64, BK);
    if (rc) return rc;
    tmC = tmA; tmC2 = tmA; tmAux = tmA;
    if (L::NOUT >= 1) { rc = rgbnm_make_tmap_bf16(&tmC, a.C, a.M, a.N, a.ldc, 64, BM_CTA); if (rc) return rc; }
    if (L::NOUT == 2) { rc = rgbnm_make_tmap_bf16(&tmC2, a.C2, a.M, a.N, a.ldc, 64, BM_CTA); if (rc) return rc; }
    if (L::HAS_AUX) { rc = rgbnm_make_tmap_bf16(&tmAux, a.aux, a.M, a.N, a.ldaux, 64, BM_CTA); if (rc) return rc; }
    Params p;
    p.M = a.M; p.N = a.N;
    p.m_tiles = (a.M + BM - 1) / BM;
    p.n_tiles = (a.N + BN - 1) / BN;
    p.k_blocks = (a.K + BK - 1) / BK;
    const int clusters_max = g_num_sms / 2;
    p.splits = 1;
    if (EPI == EPI_ATOMIC) {
        p.splits = a.splits > 0 ? a.splits : pick_splits(p.m_tiles * p.n_tiles, p.k_blocks, clusters_max);
        if (p.splits > p.k_blocks) p.splits = p.k_blocks;
    }
    p.bias = a.bias; p.posemb = a.posemb; p.pos_period = a.pos_period > 0 ? a.pos_period : 1;
    p.out_f32 = a.out_f32; p.ldo = a.ldo; p.alpha = a.alpha;
    p.trans_out = a.trans_out;
    p.perm_heads = a.perm_heads; p.perm_hd = a.perm_head_dim;
    p.ln_gamma = a.ln_gamma; p.ln_beta = a.ln_beta; p.ln_eps = a.ln_eps;
    if (a.perm_heads > 0 && (a.perm_head_dim <= 0 || a.M != 3 * a.perm_heads * a.perm_head_dim)) return RGBNM_ERR_ARG;
    p.vec_ok = (!a.trans_out && (reinterpret_cast<uintptr_t>(a.out_f32) % 16 == 0) && (a.ldo % 4 == 0) && (a.N % 4 == 0)) ? 1 : 0;
    const int tiles = p.m_tiles * p.n_tiles * p.splits;
    const int clusters = tiles < clusters_max ? tiles : clusters_max;
    kfn<<<2 * clusters, THREADS, L::TOTAL, st>>>(tmA, tmB, tmC, tmC2, tmAux, p);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

}  // namespace gemm2

extern "C" int rgbnm_gemm_bf16(const rgbnm_gemm_args* args, void* stream) {
    using namespace gemm2;
    if (!args || !args->A || !args->B || args->M <= 0 || args->N <= 0 || args->K <= 0) return RGBNM_ERR_ARG;
    const rgbnm_gemm_args& a = *args;
    if ((a.lda % 8) || (a.ldb % 8)) return RGBNM_ERR_ARG;                      // TMA: 16-byte aligned row pitch
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool wide = (a.N % 256 == 0);
    // N = 384 with a long reduction: one 256 x 384 tile per CTA pair (157 FLOP per L2 byte).  Its single TMEM accumulator
    // leaves the epilogue exposed, which only pays off when the main loop is long (K >= 768).
    static const bool no384 = (getenv("RGBNM_GEMM_NO384") != nullptr);
    static const int tall_min_k = getenv("RGBNM_TALL_MINK") ? atoi(getenv("RGBNM_TALL_MINK")) : 768;
    const bool tall = (a.N % 384 == 0) && a.K >= tall_min_k && !no384;
    switch (a.epilogue) {
        case RGBNM_EPI_STORE:
            if (!a.C || (a.ldc % 8)) return RGBNM_ERR_ARG;
            return tall ? launch<384, 4, EPI_STORE, false, false>(a, st) : launch<192, 5, EPI_STORE, false, false>(a, st);
        case RGBNM_EPI_RESIDUAL:
            if (!a.C || !a.aux || (a.ldc % 8) || (a.ldaux % 8)) return RGBNM_ERR_ARG;
            return tall ? launch<384, 4, EPI_RESIDUAL, false, false>(a, st) : launch<192, 5, EPI_RESIDUAL, false, false>(a, st);
        case RGBNM_EPI_GELU:
            if (!a.C || !a.C2 || (a.ldc % 8)) return RGBNM_ERR_ARG;
            return wide ? launch<256, 4, EPI_GELU, false, false>(a, st) : launch<192, 4, EPI_GELU, false, false>(a, st);
        case RGBNM_EPI_LNRES:
            // the whole row must sit in one tile: N <= 384, whole 32-column chunks; gamma / beta read as float4
            if (!a.C || !a.aux || !a.ln_gamma || !a.ln_beta || (a.ldc % 8) || (a.ldaux % 8) || (a.N % 32) || a.N > 384 ||
                (reinterpret_cast<uintptr_t>(a.ln_gamma) & 15) || (reinterpret_cast<uintptr_t>(a.ln_beta) & 15))
                return RGBNM_ERR_ARG;
            if (a.N <= 128) return launch<128, 6, EPI_LNRES, false, false>(a, st);       // N = 96: a 128-wide tile wastes 25 %, not 50 %
            return a.N > 192 ? launch<384, 3, EPI_LNRES, false, false>(a, st) : launch<192, 5, EPI_LNRES, false, false>(a, st);
        case RGBNM_EPI_LN:
            if (!a.C || !a.ln_gamma || !a.ln_beta || (a.ldc % 8) || (a.N % 32) || a.N > 384 ||
                (reinterpret_cast<uintptr_t>(a.ln_gamma) & 15) || (reinterpret_cast<uintptr_t>(a.ln_beta) & 15))
                return RGBNM_ERR_ARG;
            if (a.N <= 128) return launch<128, 6, EPI_LN, false, false>(a, st);
            return a.N > 192 ? launch<384, 3, EPI_LN, false, false>(a, st) : launch<192, 5, EPI_LN, false, false>(a, st);
        case RGBNM_EPI_GELU_ACT:
            if (!a.C || (a.ldc % 8)) return RGBNM_ERR_ARG;
            return wide ? launch<256, 4, EPI_GELU_ACT, false, false>(a, st) : launch<192, 5, EPI_GELU_ACT, false, false>(a, st);
        case RGBNM_EPI_DGELU:
            if (!a.C || !a.aux || (a.ldc % 8) || (a.ldaux % 8)) return RGBNM_ERR_ARG;
            return wide ? launch<256, 4, EPI_DGELU, false, false>(a, st) : launch<192, 5, EPI_DGELU, false, false>(a, st);
        case RGBNM_EPI_POSEMB:
            if (!a.C || !a.posemb || (a.ldc % 8) || (a.N % 32)) return RGBNM_ERR_ARG;
            return launch<192, 5, EPI_POSEMB, false, false>(a, st);
        case RGBNM_EPI_WGRAD_ATOMIC:
            if (!a.out_f32) return RGBNM_ERR_ARG;
            // N = 384 (every wgrad of ViT-S once the longer side sits on M): one 256 x 384 tile per CTA pair, 157 FLOP per L2 byte
            if (a.N % 384 == 0 && !getenv("RGBNM_WGRAD_BN128")) return launch<384, 5, EPI_ATOMIC, true, true>(a, st);
            return launch<128, 6, EPI_ATOMIC, true, true>(a, st);
        case RGBNM_EPI_F32:
            if (!a.out_f32) return RGBNM_ERR_ARG;
            return launch<128, 6, EPI_F32, false, false>(a, st);
        default:
            return RGBNM_ERR_ARG;
    }
}
