// Microbenchmark: issue rate of tcgen05.mma (kind::f16, M=128, K=16) for SS vs TS (A from TMEM) operands and several N.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_rate tools/micro/mma_rate.cu -I rgb_no_more_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100.cuh"
using namespace sm100;

template <int N, bool TS, bool BMN>
__global__ void __launch_bounds__(128, 1) k(long long* out, int reps) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, N, false, BMN);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
        for (int rep = 0; rep < 2; ++rep) {
            long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                const int k = r & 3;
                const uint64_t db = BMN ? make_smem_desc_sw128(sb + k * 2048, 0, 1024) : make_smem_desc_sw128(sb + k * 32, 0, 1024);
                if (TS) tc_mma_f16_ts(tb + 256, tb + k * 8, db, idesc, r != 0);
                else tc_mma_f16(tb + 256, make_smem_desc_sw128(sa + k * 32, 0, 1024), db, idesc, r != 0);
            }
            tc_commit(&bar);
            mbar_wait(&bar, rep & 1);
            long long t1 = clock64();
            out[rep] = t1 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tb);
}

template <int N, bool TS, bool BMN>
void run(const char* name, long long* d, int reps) {
    cudaFuncSetAttribute(k<N, TS, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<N, TS, BMN><<<1, 128, 100 * 1024>>>(d, reps);
    long long h[2];
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-28s N=%3d reps=%d: %lld cycles total, %.1f cycles/MMA (2nd run %.1f)  %s\n", name, N, reps, h[0], double(h[0]) / reps,
           double(h[1]) / reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* d;
    cudaMalloc(&d, 64);
    const int R = 256;
    run<64, false, false>("SS  B K-major", d, R);
    run<64, false, true>("SS  B MN-major", d, R);
    run<64, true, true>("TS  B MN-major", d, R);
    run<64, true, false>("TS  B K-major", d, R);
    run<128, false, false>("SS  B K-major", d, R);
    run<128, true, true>("TS  B MN-major", d, R);
    run<128, true, false>("TS  B K-major", d, R);
    run<208, false, false>("SS  B K-major", d, R);
    run<208, true, false>("TS  B K-major", d, R);
    run<256, false, false>("SS  B K-major", d, R);
    run<256, true, false>("TS  B K-major", d, R);
    run<256, true, true>("TS  B MN-major", d, R);
    return 0;
}
