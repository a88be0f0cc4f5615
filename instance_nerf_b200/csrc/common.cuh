// Shared helpers for libinerf_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "../../include/inerf_b200.h"

#define INERF_REQUIRE(p) do { if ((p) == nullptr) return INERF_ERR_NULL; } while (0)
#define INERF_LAUNCH_CHECK() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; } while (0)

static inline unsigned int div_up(unsigned long long a, unsigned int b) { return (unsigned int)((a + b - 1) / b); }

// B200: 148 SMs.  Grid-stride kernels are sized to a multiple of this.
constexpr int kNumSMs = 148;
// SM count of the CURRENT device (persistent kernels launch one CTA per SM); falls back to the B200 figure on error
static inline int device_sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return kNumSMs;
    return n;
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
