// Library-wide ABI helpers (version, last CUDA error text).
#include <cuda_runtime.h>

#include <cstring>

#include "../../include/rgbnm_b200.h"
#include "common.cuh"

static thread_local char g_last_cuda_error[512] = "";

void rgbnm_set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s (%s)", where, cudaGetErrorName(e),
             cudaGetErrorString(e));
}

extern "C" {
const char* rgbnm_last_cuda_error(void) { return g_last_cuda_error; }
int rgbnm_abi_version(void) { return 5; }
}
