// LEGACY, NOT BUILT: the round-1 single-CTA (cta_group::1) tcgen05 GEMM, superseded by rgb_no_more_b200/csrc/gemm2_tc.cu (CTA pairs).
// Kept for reference only (DESIGN.md section 4 explains why pairs won); its tensor-map helpers now live in csrc/tmap.cu.
// Dense bf16 GEMMs of the DCT ViT on the 5th-gen tensor cores (sm_100a): tcgen05.mma fed by
// TMA, fp32 accumulators in TMEM, fused epilogues.  One persistent, warp-specialised kernel
// template covers every dense contraction of models/plainvit.py forward and backward:
//
//   forward    Y[M,N]  = X[M,K] . W[N,K]^T            both operands K-major  (nn.Linear)
//              epilogues: +bias | +bias,+posemb (plainvit.py:194-198) | +bias,GELU(erf) with the
//              pre-activation kept for backward (plainvit.py:485-491) | +bias,+residual (:475-479)
//   dgrad      dX[M,K] = dY[M,N] . Wt[K,N]^T          K-major, Wt = transposed bf16 weight copy
//              epilogues: none | x GELU'(u)
//   wgrad      dW[N,K] += dY[M,N]^T . X[M,K]          both operands MN-major (token index is the
//              reduction), split along the tokens, fp32 red.add into the flat gradient buffer
//
// Tile 128 x BN x 64, UMMA 128 x BN x 16 (cta_group::1), 4-stage TMA->smem ring, two TMEM
// accumulator buffers so the epilogue of tile i overlaps the main loop of tile i+1.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = epilogue
// (TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/rgbnm_b200.h"
#include "common.cuh"
#include "sm100.cuh"

namespace gemm {
using namespace sm100;

constexpr int BM = 128, BK = 64;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (2 + EPI_WARPS) * 32;

struct Params {
    int M, N;                 // output extent (rows, cols) used for bounds in the fp32 / atomic epilogues
    int m_tiles, n_tiles;     // tiles of BM x BN
    int k_blocks;             // BK-blocks of the whole reduction
    int splits;               // split-K factor (atomic epilogue only), k_blocks % splits == 0
    const float* bias;        // [N] fp32 or nullptr
    const float* posemb;      // [pos_period][N] fp32 or nullptr
    int pos_period;
    float* out_f32;           // fp32 / atomic epilogues
    long long ldo;
    float alpha;              // scale applied to the accumulator (atomic epilogue)
};

template <int BN, int STAGES, int NOUT>
struct Smem {
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    static constexpr int STG_BYTES = BM * BN * 2;
    static constexpr int OFF_A = 0, OFF_B = OFF_A + STAGES * A_BYTES, OFF_STG = OFF_B + STAGES * B_BYTES;
    static constexpr int OFF_BIAS = OFF_STG + NOUT * STG_BYTES;
    static constexpr int OFF_BAR = OFF_BIAS + BN * 4;
    static constexpr int NBARS = 2 * STAGES + 4 + 1;
    static constexpr int OFF_TMEM = OFF_BAR + NBARS * 8;
    static constexpr int TOTAL = OFF_TMEM + 16 + 1024;   // + slack for the manual 1024-byte alignment
};

// GELU(erf) and its derivative (nn.GELU() default, plainvit.py:487).  erf by Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7, far below the bf16 output rounding): one MUFU.RCP + one MUFU.EX2 per element, and the
// exponential exp(-x^2/2) is shared with the Gaussian density of the derivative.
struct GeluParts { float cdf, pdf; };
__device__ __forceinline__ GeluParts gelu_parts(float x) {
    const float u = fabsf(x) * 0.70710678118654752f;
    float t, ex;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, u, 1.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(x * x * -0.72134752044448170f));   // exp(-x^2 / 2)
    float poly = fmaf(t, 1.061405429f, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float erf_abs = fmaf(-poly * t, ex, 1.0f);
    GeluParts g;
    g.cdf = fmaf(copysignf(0.5f, x), erf_abs, 0.5f);
    g.pdf = 0.3989422804014327f * ex;
    return g;
}
__device__ __forceinline__ float gelu_erf(float x) { return x * gelu_parts(x).cdf; }
__device__ __forceinline__ float dgelu_erf(float x) {
    const GeluParts g = gelu_parts(x);
    return fmaf(x, g.pdf, g.cdf);
}

__device__ __forceinline__ void bar_sync_epi(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(EPI_WARPS * 32) : "memory"); }

enum Epi : int {
    EPI_STORE = 0,       // C = acc (+bias)                                   bf16
    EPI_RESIDUAL = 1,    // C = acc + bias + AUX                              bf16, AUX bf16 [M,N]
    EPI_GELU = 2,        // C = acc + bias (pre-activation), C2 = gelu(C)     bf16 x 2
    EPI_DGELU = 3,       // C = acc * gelu'(AUX)                              bf16, AUX = pre-activation
    EPI_POSEMB = 4,      // C = acc + bias + posemb[row % period]             bf16
    EPI_ATOMIC = 5,      // out_f32[row][col] += alpha * acc                  fp32 red.add (split-K)
    EPI_F32 = 6          // out_f32[row][col] = acc + bias                    fp32 (logits)
};

template <int BN, int STAGES, int EPI, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
            const __grid_constant__ CUtensorMap tmAux, const Params p) {
    constexpr bool STAGED = (EPI != EPI_ATOMIC && EPI != EPI_F32);
    constexpr bool HAS_AUX = (EPI == EPI_RESIDUAL || EPI == EPI_DGELU);
    constexpr int NOUT = EPI == EPI_GELU ? 2 : (STAGED ? 1 : 0);
    using L = Smem<BN, STAGES, NOUT>;
    constexpr int NSUB = BN / 64;            // 64-column sub-tiles of the staging buffer / MN-major boxes
    constexpr int CHUNKS = BN / 32;          // 32-column TMEM loads per tile
    constexpr int CH_PER_HALF = CHUNKS / 2;
    static_assert(BN % 64 == 0 && CHUNKS % 2 == 0 && 2 * BN <= 512, "tile shape");

    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = bars + 2 * STAGES + 2;
    uint64_t* auxbar = bars + 2 * STAGES + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM);
    float* sbias = reinterpret_cast<float*>(smem + L::OFF_BIAS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb_per_split = p.k_blocks / p.splits;
    const int total_tiles = p.m_tiles * p.n_tiles * p.splits;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmA);
        prefetch_tensormap(&tmB);
        if (STAGED) prefetch_tensormap(&tmC);
        if (EPI == EPI_GELU) prefetch_tensormap(&tmC2);
        if (HAS_AUX) prefetch_tensormap(&tmAux);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, EPI_WARPS); }
        mbar_init(auxbar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int sp = tile % p.splits, rest = tile / p.splits;
                const int n_blk = rest % p.n_tiles, m_blk = rest / p.n_tiles;
                for (int kb = sp * kb_per_split; kb < (sp + 1) * kb_per_split; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1);
                    mbar_arrive_expect_tx(full + stage, L::A_BYTES + L::B_BYTES);
                    unsigned char* sa = smem + L::OFF_A + stage * L::A_BYTES;
                    unsigned char* sb = smem + L::OFF_B + stage * L::B_BYTES;
                    if (!A_MN) {
                        tma_load_2d(sa, &tmA, full + stage, kb * BK, m_blk * BM);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * (BK * 128), &tmA, full + stage, m_blk * BM + j * 64, kb * BK);
                    }
                    if (!B_MN) {
                        tma_load_2d(sb, &tmB, full + stage, kb * BK, n_blk * BN);
                    } else {
#pragma unroll
                        for (int j = 0; j < NSUB; ++j) tma_load_2d(sb + j * (BK * 128), &tmB, full + stage, n_blk * BN + j * 64, kb * BK);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
            int stage = 0, phase = 0, it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(tempty + acc, ((it >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < kb_per_split; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + L::OFF_A + stage * L::A_BYTES);
                    const uint32_t sb = smem_u32(smem + L::OFF_B + stage * L::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // K-major: 16 bf16 = 32 bytes further inside the 128-byte swizzle row;
                        // MN-major: 16 reduction rows = 2048 bytes further
                        const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * 2048, BK * 128, 1024)
                                                 : make_smem_desc_sw128(sa + k * 32, 0, 1024);
                        const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * 2048, BK * 128, 1024)
                                                 : make_smem_desc_sw128(sb + k * 32, 0, 1024);
                        tc_mma_f16(d_tmem, da, db, idesc, (kb | k) != 0);
                    }
                    tc_commit(empty + stage);                      // smem slot free once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull + acc);                            // accumulator complete
            }
        }
    } else {
        // ===================================== epilogue ==========================================
        const int ew = warp - 2;
        const int quad = warp & 3;               // TMEM lane quadrant this warp may read
        const int half = ew >> 2;                // which half of the tile's columns
        const int row = quad * 32 + lane;        // row inside the tile
        const bool leader = (ew == 0 && lane == 0);
        unsigned char* stg = smem + L::OFF_STG;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int rest = tile / p.splits;
            const int n_blk = rest % p.n_tiles, m_blk = rest / p.n_tiles;
            const int acc = it & 1;
            const int n0 = n_blk * BN, m0 = m_blk * BM;
            if (STAGED) {
                if (leader) {
                    tma_store_wait_read<0>();                      // previous tile's stores have left the staging buffer
                    if (HAS_AUX) {
                        mbar_arrive_expect_tx(auxbar, L::STG_BYTES);
#pragma unroll
                        for (int s = 0; s < NSUB; ++s) tma_load_2d(stg + s * (BM * 128), &tmAux, auxbar, n0 + s * 64, m0);
                    }
                }
            }
            if (p.bias != nullptr) {
                for (int j = threadIdx.x - 64; j < BN; j += EPI_WARPS * 32) sbias[j] = (n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.0f;
            }
            bar_sync_epi(1);                                       // staging buffer free, bias tile visible
            if (HAS_AUX) mbar_wait(auxbar, it & 1);
            mbar_wait(tfull + acc, (it >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < CH_PER_HALF; ++cc) {
                const int c = half * CH_PER_HALF + cc;             // 32-column chunk of the tile
                uint32_t r[32];
                tmem_ld32(tmem_base + (uint32_t(quad * 32) << 16) + acc * BN + c * 32, r);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (EPI != EPI_ATOMIC && EPI != EPI_DGELU) {
                    if (p.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(sbias + c * 32 + j);
                            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                        }
                    }
                }
                if (EPI == EPI_POSEMB) {
                    const float* pe = p.posemb + size_t((m0 + row) % p.pos_period) * p.N + n0 + c * 32;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 e4 = __ldg(reinterpret_cast<const float4*>(pe + j));
                        v[j] += e4.x; v[j + 1] += e4.y; v[j + 2] += e4.z; v[j + 3] += e4.w;
                    }
                }
                if (EPI == EPI_ATOMIC || EPI == EPI_F32) {
                    const int gr = m0 + row;
                    if (gr < p.M) {
                        float* o = p.out_f32 + size_t(gr) * p.ldo + n0 + c * 32;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (n0 + c * 32 + j < p.N) {
                                if (EPI == EPI_ATOMIC) atomicAdd(o + j, p.alpha * v[j]);
                                else o[j] = v[j];
                            }
                        }
                    }
                } else {
                    // staging layout = what a SWIZZLE_128B TMA box {64 cols, 128 rows} expects:
                    // sub-tile s (64 cols), row r at r*128 bytes, 16-byte chunk q stored at q ^ (r & 7)
                    const int s = c >> 1;
                    unsigned char* rowp = stg + s * (BM * 128) + row * 128;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int q = (c & 1) * 4 + t;
                        uint4* slot = reinterpret_cast<uint4*>(rowp + ((q ^ (row & 7)) << 4));
                        float* x = v + 8 * t;
                        if (HAS_AUX) {
                            const uint4 a = *slot;
                            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float lo = __uint_as_float(aw[e] << 16), hi = __uint_as_float(aw[e] & 0xffff0000u);
                                if (EPI == EPI_RESIDUAL) { x[2 * e] += lo; x[2 * e + 1] += hi; }
                                else { x[2 * e] *= dgelu_erf(lo); x[2 * e + 1] *= dgelu_erf(hi); }
                            }
                        }
                        *slot = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
                        if (EPI == EPI_GELU) {
                            uint4* slot2 = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(slot) + L::STG_BYTES);
                            *slot2 = make_uint4(pack_bf16(gelu_erf(x[0]), gelu_erf(x[1])), pack_bf16(gelu_erf(x[2]), gelu_erf(x[3])),
                                                pack_bf16(gelu_erf(x[4]), gelu_erf(x[5])), pack_bf16(gelu_erf(x[6]), gelu_erf(x[7])));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + acc);              // this warp is done with the accumulator
            if (STAGED) fence_proxy_async();                   // staging writes -> visible to the TMA engine
            bar_sync_epi(2);                                       // every warp is done with sbias / the staging tile
            if (STAGED) {
                if (leader) {
#pragma unroll
                    for (int s = 0; s < NSUB; ++s) {
                        tma_store_2d(&tmC, stg + s * (BM * 128), n0 + s * 64, m0);
                        if (EPI == EPI_GELU) tma_store_2d(&tmC2, stg + L::STG_BYTES + s * (BM * 128), n0 + s * 64, m0);
                    }
                    tma_store_commit();
                }
            }
        }
        if (STAGED && leader) tma_store_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || !sym) return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

}  // namespace gemm

// 2-D bf16 row-major tensor [rows][cols] with leading dimension ld (elements); box = {box_cols, box_rows}, 128-byte swizzle.
int rgbnm_make_tmap_bf16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows) {
    gemm::EncodeTiledFn enc = gemm::get_encode();
    if (!enc) return RGBNM_ERR_CUDA;
    cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(ld) * 2};
    cuuint32_t box[2] = {cuuint32_t(box_cols), cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rgbnm_set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled");
        return RGBNM_ERR_CUDA;
    }
    return RGBNM_OK;
}

// 3-D bf16 tensor [d2][d1][d0] (d0 contiguous; pitches ld1, ld2 in elements); box = {box0, box1, 1}, 128-byte swizzle.
int rgbnm_make_tmap_bf16_3d(CUtensorMap* map, const void* ptr, long long d0, long long d1, long long d2, long long ld1,
                            long long ld2, int box0, int box1) {
    gemm::EncodeTiledFn enc = gemm::get_encode();
    if (!enc) return RGBNM_ERR_CUDA;
    cuuint64_t dims[3] = {cuuint64_t(d0), cuuint64_t(d1), cuuint64_t(d2)};
    cuuint64_t strides[2] = {cuuint64_t(ld1) * 2, cuuint64_t(ld2) * 2};
    cuuint32_t box[3] = {cuuint32_t(box0), cuuint32_t(box1), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rgbnm_set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(3d)");
        return RGBNM_ERR_CUDA;
    }
    return RGBNM_OK;
}

namespace gemm {

static int g_num_sms = 0;

template <int BN, int STAGES, int EPI, bool A_MN, bool B_MN>
static int launch(const rgbnm_gemm_args& a, cudaStream_t st) {
    constexpr int NOUT = EPI == EPI_GELU ? 2 : ((EPI != EPI_ATOMIC && EPI != EPI_F32) ? 1 : 0);
    using L = Smem<BN, STAGES, NOUT>;
    static bool configured = false;
    auto kfn = gemm_kernel<BN, STAGES, EPI, A_MN, B_MN>;
    if (!configured) {
        RGBNM_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        configured = true;
    }
    if (g_num_sms == 0) {
        int dev = 0;
        RGBNM_CUDA_CHECK(cudaGetDevice(&dev));
        RGBNM_CUDA_CHECK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    // logical problem: C[M,N] = sum_k A[m,k] B[n,k]
    CUtensorMap tmA, tmB, tmC, tmC2, tmAux;
    int rc;
    // K-major operand stored [rows = M|N][cols = K]; MN-major operand stored [rows = K][cols = M|N]
    if (!A_MN) rc = rgbnm_make_tmap_bf16(&tmA, a.A, a.M, a.K, a.lda, BK, BM);
    else rc = rgbnm_make_tmap_bf16(&tmA, a.A, a.K, a.M, a.lda, 64, BK);
    if (rc) return rc;
    if (!B_MN) rc = rgbnm_make_tmap_bf16(&tmB, a.B, a.N, a.K, a.ldb, BK, BN);
    else rc = rgbnm_make_tmap_bf16(&tmB, a.B, a.K, a.N, a.ldb, 64, BK);
    if (rc) return rc;
    tmC = tmA; tmC2 = tmA; tmAux = tmA;
    if (NOUT >= 1) { rc = rgbnm_make_tmap_bf16(&tmC, a.C, a.M, a.N, a.ldc, 64, BM); if (rc) return rc; }
    if (NOUT == 2) { rc = rgbnm_make_tmap_bf16(&tmC2, a.C2, a.M, a.N, a.ldc, 64, BM); if (rc) return rc; }
    if (EPI == EPI_RESIDUAL || EPI == EPI_DGELU) { rc = rgbnm_make_tmap_bf16(&tmAux, a.aux, a.M, a.N, a.ldaux, 64, BM); if (rc) return rc; }
    Params p;
    p.M = a.M; p.N = a.N;
    p.m_tiles = (a.M + BM - 1) / BM;
    p.n_tiles = (a.N + BN - 1) / BN;
    p.k_blocks = (a.K + BK - 1) / BK;
    p.splits = (EPI == EPI_ATOMIC && a.splits > 1) ? a.splits : 1;
    while (p.k_blocks % p.splits) --p.splits;
    p.bias = a.bias; p.posemb = a.posemb; p.pos_period = a.pos_period > 0 ? a.pos_period : 1;
    p.out_f32 = a.out_f32; p.ldo = a.ldo; p.alpha = a.alpha;
    const int tiles = p.m_tiles * p.n_tiles * p.splits;
    const int grid = tiles < g_num_sms ? tiles : g_num_sms;
    kfn<<<grid, THREADS, L::TOTAL, st>>>(tmA, tmB, tmC, tmC2, tmAux, p);
    RGBNM_CUDA_CHECK(cudaGetLastError());
    return RGBNM_OK;
}

}  // namespace gemm

// Single-CTA (cta_group::1) kernel: superseded by gemm2_tc.cu, kept selectable (RGBNM_GEMM_V1=1) for A/B measurements.
int rgbnm_gemm_bf16_v1(const rgbnm_gemm_args* args, void* stream) {
    using namespace gemm;
    if (!args || !args->A || !args->B || args->M <= 0 || args->N <= 0 || args->K <= 0) return RGBNM_ERR_ARG;
    const rgbnm_gemm_args& a = *args;
    if ((a.lda % 8) || (a.ldb % 8)) return RGBNM_ERR_ARG;                      // TMA: 16-byte aligned row pitch
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (a.epilogue) {
        case RGBNM_EPI_STORE:
            if (!a.C || (a.ldc % 8)) return RGBNM_ERR_ARG;
            return launch<192, 4, EPI_STORE, false, false>(a, st);
        case RGBNM_EPI_RESIDUAL:
            if (!a.C || !a.aux || (a.ldc % 8) || (a.ldaux % 8)) return RGBNM_ERR_ARG;
            return launch<192, 4, EPI_RESIDUAL, false, false>(a, st);
        case RGBNM_EPI_GELU:
            if (!a.C || !a.C2 || (a.ldc % 8)) return RGBNM_ERR_ARG;
            return launch<192, 3, EPI_GELU, false, false>(a, st);
        case RGBNM_EPI_DGELU:
            if (!a.C || !a.aux || (a.ldc % 8) || (a.ldaux % 8)) return RGBNM_ERR_ARG;
            return launch<192, 4, EPI_DGELU, false, false>(a, st);
        case RGBNM_EPI_POSEMB:
            if (!a.C || !a.posemb || (a.ldc % 8) || (a.N % 4)) return RGBNM_ERR_ARG;
            return launch<192, 4, EPI_POSEMB, false, false>(a, st);
        case RGBNM_EPI_WGRAD_ATOMIC:
            if (!a.out_f32) return RGBNM_ERR_ARG;
            return launch<192, 4, EPI_ATOMIC, true, true>(a, st);
        case RGBNM_EPI_F32:
            if (!a.out_f32) return RGBNM_ERR_ARG;
            return launch<192, 4, EPI_F32, false, false>(a, st);
        default:
            return RGBNM_ERR_ARG;
    }
}
