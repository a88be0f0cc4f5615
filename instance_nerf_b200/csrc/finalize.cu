// inerf_frame_to_u8: a rendered frame -> what MaskTrainer.test / evaluate_one_epoch write to disk (nerf/utils.py:1425-1431,
// 1461-1485, 1624-1634): rgb and depth quantised as numpy's (x * 255).astype(uint8) does for x in [0, 1] (fp32 product, truncation;
// out-of-range values saturate here where numpy wraps), instance label = argmax_k softmax(logits)[k] = argmax_k logits[k]
// (lowest index on exact ties).  The reference copies image fp32 + depth fp32 + argmax int64 = 24 B / pixel to the host and
// quantises there; this leaves 5 B / pixel to copy.  One warp per ray: the K logits of a ray are one coalesced row.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint8_t quant_u8(float v) {
    const float s = __fmul_rn(v, 255.0f);
    return (uint8_t)(int)fminf(fmaxf(s, 0.0f), 255.0f);   // NaN -> 0
}

__global__ void __launch_bounds__(256) k_frame_to_u8(const float* __restrict__ image, const float* __restrict__ depth, const float* __restrict__ logits,
                                                     uint32_t N, uint32_t K, uint8_t* __restrict__ rgb, uint8_t* __restrict__ depth_u8,
                                                     uint8_t* __restrict__ label) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < N; r += warps) {
        if (logits != nullptr) {
            const float* x = logits + (size_t)r * K;
            float best = -INFINITY;
            uint32_t arg = 0xffffffffu;
            for (uint32_t k = lane; k < K; k += 32u) {
                const float v = __ldg(x + k);
                if (v > best || arg == 0xffffffffu) { best = v; arg = k; }   // first element always taken (handles -inf / NaN rows)
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const uint32_t oa = __shfl_xor_sync(0xffffffffu, arg, o);
                if (oa != 0xffffffffu && (arg == 0xffffffffu || ob > best || (ob == best && oa < arg))) { best = ob; arg = oa; }
            }
            if (lane == 0) label[r] = (uint8_t)arg;
        }
        if (lane < 3) rgb[(size_t)r * 3 + lane] = quant_u8(__ldg(image + (size_t)r * 3 + lane));
        if (lane == 3 && depth_u8 != nullptr) depth_u8[r] = quant_u8(__ldg(depth + r));
    }
}

}  // namespace

extern "C" int inerf_frame_to_u8(const float* image, const float* depth, const float* logits, uint32_t N, uint32_t K, uint8_t* rgb,
                                 uint8_t* depth_u8, uint8_t* label, void* stream) {
    if (N == 0) return INERF_OK;
    INERF_REQUIRE(image); INERF_REQUIRE(rgb);
    if ((depth == nullptr) != (depth_u8 == nullptr) || (logits == nullptr) != (label == nullptr)) return INERF_ERR_NULL;
    if (logits != nullptr && (K == 0 || K > 256)) return INERF_ERR_SIZE;
    const uint32_t blocks = min(div_up((unsigned long long)N * 32ull, 256), (unsigned int)(device_sm_count() * 8));
    k_frame_to_u8<<<blocks, 256, 0, (cudaStream_t)stream>>>(image, depth, logits, N, K, rgb, depth_u8, label);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}
