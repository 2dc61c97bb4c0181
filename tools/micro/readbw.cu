// HBM read-bandwidth probe: streams `bytes` with 16-byte loads (grid-stride), 148 x k CTAs; prints GB/s.  For the K0 roofline
// discussion in DESIGN.md: what a pure read stream reaches on this box next to the driver-measured copy figure.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rd(const int4* __restrict__ p, size_t n, int* out) {
    int acc = 0;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const int4 v = __ldg(p + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678) *out = acc;
}
int main() {
    const size_t bytes = size_t(1) << 30;
    int4* p; int* o;
    cudaMalloc(&p, bytes); cudaMalloc(&o, 4); cudaMemset(p, 1, bytes);
    for (int ctas : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) for (int thr : {256, 512}) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        rd<<<ctas, thr>>>(p, bytes / 16, o);
        cudaEventRecord(a);
        for (int r = 0; r < 5; ++r) rd<<<ctas, thr>>>(p, bytes / 16, o);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("read %d x %d: %.1f GB/s\n", ctas, thr, bytes * 5 / ms / 1e6);
    }
    // 154 MB working set read once per launch (as K0's eval-geometry input), 4 distinct buffers cycled
    const size_t small = 154u << 20;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        for (int r = 0; r < 20; ++r) rd<<<148 * 8, 512>>>(p + (r % 4) * (small / 16), small / 16, o);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("154 MiB per launch: %.1f us per launch, %.1f GB/s\n", ms * 1000 / 20, small * 20 / ms / 1e6);
    }
    return 0;
}
