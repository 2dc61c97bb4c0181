// Shared host/device helpers for the rgbnm CUDA sources.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>

void rgbnm_set_cuda_error(cudaError_t e, const char* where);

#define RGBNM_CUDA_CHECK(expr)                       \
    do {                                             \
        cudaError_t _e = (expr);                     \
        if (_e != cudaSuccess) {                     \
            rgbnm_set_cuda_error(_e, #expr);         \
            return RGBNM_ERR_CUDA;                   \
        }                                            \
    } while (0)
