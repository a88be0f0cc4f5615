// libinerf_probe.so -- measurement helpers, NOT part of the product library (bench.py and tests/dev_l2_peak.py load it to
// measure the denominators of the gather roofline on the box they run on; SURVEY.md section 8d: "the builder must
// additionally measure an L2-resident read peak on the same box").
//
//   inerf_probe_stream   every thread streams 16-byte loads over a buffer `reps` times (grid-stride, coalesced): with a
//                        buffer that fits the 126 MB L2 all passes after the first are served from L2
//   inerf_probe_gather   every thread issues `per_thread` 8-byte ld.global.nc gathers at pseudo-random entries of a
//                        table, 8 independent loads in flight per thread, at full occupancy: the L1TEX / L2 rate for the
//                        access shape of the hashed levels (one sector per lane, no reuse)
// Both return 0 or a cudaError_t; results are folded into `sink` so the loads cannot be optimised away.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__global__ void __launch_bounds__(1024) k_stream(const uint4* __restrict__ buf, uint64_t n16, uint32_t reps, uint32_t* __restrict__ sink) {
    uint32_t acc = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint32_t r = 0; r < reps; r++) {
        uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n16; i += 4 * stride) {   // 4 independent 16-byte loads in flight
            const uint4 a = __ldg(buf + i), b = __ldg(buf + i + stride), c = __ldg(buf + i + 2 * stride), d = __ldg(buf + i + 3 * stride);
            acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
        }
        for (; i < n16; i += stride) { const uint4 a = __ldg(buf + i); acc ^= a.x ^ a.y ^ a.z ^ a.w; }
    }
    if (acc == 0x9E3779B9u) sink[0] = acc;   // practically never true: keeps the loads alive
}

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(1024) k_gather(const uint2* __restrict__ table, uint32_t n_entries, uint32_t per_thread, uint32_t* __restrict__ sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0, s = hash32(tid * 2654435761u + 12345u);
    for (uint32_t i = 0; i < per_thread; i += 8) {
        uint32_t idx[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { s = s * 1664525u + 1013904223u; idx[j] = (uint32_t)(((uint64_t)hash32(s) * n_entries) >> 32); }
        uint2 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __ldg(table + idx[j]);
#pragma unroll
        for (int j = 0; j < 8; j++) acc ^= v[j].x ^ v[j].y;
    }
    if (acc == 0x9E3779B9u) sink[0] = acc;
}

// random float2 (8-byte) or float4 (16-byte) REDs into a table: the scatter rate of the hash-grid backward
template <int V>
__global__ void __launch_bounds__(1024) k_red(float* __restrict__ table, uint32_t n_entries, uint32_t per_thread) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = hash32(tid * 2654435761u + 777u);
    for (uint32_t i = 0; i < per_thread; i++) {
        s = s * 1664525u + 1013904223u;
        const uint32_t idx = (uint32_t)(((uint64_t)hash32(s) * n_entries) >> 32);
        if (V == 2) atomicAdd(reinterpret_cast<float2*>(table) + idx, make_float2(1e-9f, 1e-9f));
        else atomicAdd(reinterpret_cast<float4*>(table) + (idx >> 1), make_float4(1e-9f, 1e-9f, 1e-9f, 1e-9f));
    }
}

}  // namespace

extern "C" int inerf_probe_red(void* table, uint32_t n_entries, uint32_t per_thread, uint32_t blocks, int vec, void* stream) {
    if (vec == 4) k_red<4><<<blocks, 1024, 0, (cudaStream_t)stream>>>((float*)table, n_entries, per_thread);
    else k_red<2><<<blocks, 1024, 0, (cudaStream_t)stream>>>((float*)table, n_entries, per_thread);
    return (int)cudaGetLastError();
}

extern "C" int inerf_probe_stream(const void* buf, uint64_t bytes, uint32_t reps, uint32_t blocks, uint32_t* sink, void* stream) {
    k_stream<<<blocks, 1024, 0, (cudaStream_t)stream>>>((const uint4*)buf, bytes / 16, reps, sink);
    return (int)cudaGetLastError();
}

extern "C" int inerf_probe_gather(const void* table, uint32_t n_entries, uint32_t per_thread, uint32_t blocks, uint32_t* sink, void* stream) {
    k_gather<<<blocks, 1024, 0, (cudaStream_t)stream>>>((const uint2*)table, n_entries, per_thread, sink);
    return (int)cudaGetLastError();
}
