// Occupancy-grid EMA update + bitfield packing without host round trips (sm_100a).
//
// Replaces the tail of NeRFMaskRenderer.update_extra_state
// (nerf/mask_renderer.py:532-540): boolean-mask indexing, torch.maximum, a
// .item() sync for the mean, then packbits (raymarching.cu:267-289).  Here:
//   launch 1: grid = max(grid * decay, tmp) where both >= 0, and the sum of
//             clamp(grid, 0) reduced in double into sum_out[0];
//   launch 2: thresh = min(sum / n_cells, density_thresh) read on the device,
//             bits packed 32 cells per thread.
// Both are pure streaming kernels: 16.125 B per cell, float4 loads.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_occ_ema(float* __restrict__ grid, const float* __restrict__ tmp, uint32_t n, float decay,
                                                 double* __restrict__ sum_out) {
    double local = 0.0;
    const uint32_t n4 = n >> 2;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        float4 g = reinterpret_cast<float4*>(grid)[i];
        const float4 t = __ldg(reinterpret_cast<const float4*>(tmp) + i);
        if (g.x >= 0 && t.x >= 0) g.x = fmaxf(__fmul_rn(g.x, decay), t.x);
        if (g.y >= 0 && t.y >= 0) g.y = fmaxf(__fmul_rn(g.y, decay), t.y);
        if (g.z >= 0 && t.z >= 0) g.z = fmaxf(__fmul_rn(g.z, decay), t.z);
        if (g.w >= 0 && t.w >= 0) g.w = fmaxf(__fmul_rn(g.w, decay), t.w);
        reinterpret_cast<float4*>(grid)[i] = g;
        local += (double)fmaxf(g.x, 0.f) + (double)fmaxf(g.y, 0.f) + (double)fmaxf(g.z, 0.f) + (double)fmaxf(g.w, 0.f);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3u)) {
        const uint32_t i = (n & ~3u) + threadIdx.x;
        float g = grid[i];
        const float t = tmp[i];
        if (g >= 0 && t >= 0) g = fmaxf(__fmul_rn(g, decay), t);
        grid[i] = g;
        local += (double)fmaxf(g, 0.f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ double warp_part[8];
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int i = 0; i < 8; i++) s += warp_part[i];
        atomicAdd(sum_out, s);
    }
}

__global__ void __launch_bounds__(256) k_occ_pack(const float* __restrict__ grid, uint32_t n_cells, const double* __restrict__ sum_in,
                                                  float density_thresh, uint8_t* __restrict__ bitfield, float* __restrict__ mean_out) {
    const float mean = (float)(sum_in[0] / (double)n_cells);
    const float thresh = fminf(mean, density_thresh);
    if (mean_out && blockIdx.x == 0 && threadIdx.x == 0) mean_out[0] = mean;
    const uint32_t nwords = n_cells >> 5;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += gridDim.x * blockDim.x) {
        const float4* g = reinterpret_cast<const float4*>(grid) + (size_t)w * 8;
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 v = __ldg(g + i);
            bits |= (uint32_t)(v.x > thresh) << (i * 4 + 0);
            bits |= (uint32_t)(v.y > thresh) << (i * 4 + 1);
            bits |= (uint32_t)(v.z > thresh) << (i * 4 + 2);
            bits |= (uint32_t)(v.w > thresh) << (i * 4 + 3);
        }
        reinterpret_cast<uint32_t*>(bitfield)[w] = bits;
    }
}

}  // namespace

extern "C" int inerf_occupancy_ema(float* density_grid, const float* tmp_grid, uint32_t n_cells, float decay, double* sum_out,
                                   void* stream) {
    if (n_cells == 0) return INERF_OK;
    INERF_REQUIRE(density_grid); INERF_REQUIRE(tmp_grid); INERF_REQUIRE(sum_out);
    if (((uintptr_t)density_grid & 15u) || ((uintptr_t)tmp_grid & 15u)) return INERF_ERR_ALIGN;
    const unsigned int blocks = min(div_up(max(n_cells >> 2, 1u), 256), (unsigned int)(kNumSMs * 8));
    k_occ_ema<<<blocks, 256, 0, (cudaStream_t)stream>>>(density_grid, tmp_grid, n_cells, decay, sum_out);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_occupancy_pack(const float* density_grid, uint32_t n_cells, const double* sum_in, float density_thresh,
                                    uint8_t* bitfield, float* mean_out, void* stream) {
    if (n_cells == 0) return INERF_OK;
    if (n_cells & 31u) return INERF_ERR_SIZE;  // C * H^3 is a multiple of 32 for every H >= 4
    INERF_REQUIRE(density_grid); INERF_REQUIRE(sum_in); INERF_REQUIRE(bitfield);
    if (((uintptr_t)density_grid & 15u) || ((uintptr_t)bitfield & 3u)) return INERF_ERR_ALIGN;
    const unsigned int blocks = min(div_up(n_cells >> 5, 256), (unsigned int)(kNumSMs * 8));
    k_occ_pack<<<blocks, 256, 0, (cudaStream_t)stream>>>(density_grid, n_cells, sum_in, density_thresh, bitfield, mean_out);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}
