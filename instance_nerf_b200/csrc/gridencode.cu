// Multiresolution hash / tiled grid encoder, forward and backward (sm_100a).
//
// Replaces gridencoder/src/gridencoder.cu:50-500 of the reference.
//
// Thread mapping: consecutive threads take consecutive LEVELS of the same point
// (the reference puts the level on blockIdx.y).  A warp therefore covers 32/L
// whole points and, in the [B, L*C] layout the network consumes, writes one
// contiguous 128-byte span per warp; the reference writes [L, B, C] and then
// pays a permute+copy over the whole activation (grid.py:57, :75).
//
// fp16 semantics follow the reference exactly: interpolation weights in fp32,
// every product rounded to half, accumulated in half (c10::Half `+=`):
// gridencoder.cu:161-185.  Results are bit-identical for both dtypes.
#include "common.cuh"
#include "field_device.cuh"   // level geometry / corner indices shared with the fused kernels (bit-identical to grid_index below)

namespace {

template <typename T> struct Acc;
template <> struct Acc<float> {
    __device__ static __forceinline__ float mul_add(float acc, float w, float g) { return fmaf(w, g, acc); }
    __device__ static __forceinline__ float load(const float* p) { return __ldg(p); }
};
template <> struct Acc<__half> {
    // acc = half(float(acc) + float(half(w * float(g))))
    __device__ static __forceinline__ __half mul_add(__half acc, float w, __half g) {
        return __float2half_rn(__fadd_rn(__half2float(acc), __half2float(__float2half_rn(__fmul_rn(w, __half2float(g))))));
    }
    __device__ static __forceinline__ __half load(const __half* p) { return __ldg(p); }
};

template <uint32_t D>
__device__ __forceinline__ uint32_t fast_hash(const uint32_t pos_grid[D]) {
    constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t result = 0;
#pragma unroll
    for (uint32_t i = 0; i < D; ++i) result ^= pos_grid[i] * primes[i];
    return result;
}

// gridencoder.cu:66-84
template <uint32_t D, uint32_t C>
__device__ __forceinline__ uint32_t grid_index(uint32_t gridtype, bool align_corners, uint32_t hashmap_size, uint32_t resolution,
                                               const uint32_t pos_grid[D]) {
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) {
        index += pos_grid[d] * stride;
        stride *= align_corners ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) index = fast_hash<D>(pos_grid);
    return (index % hashmap_size) * C;
}

struct LevelGeom {
    uint32_t hashmap_size, resolution;
    float scale;
};
__device__ __forceinline__ LevelGeom level_geom(const int32_t* __restrict__ offsets, uint32_t level, float S, uint32_t H) {
    LevelGeom g;
    g.hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    g.scale = __fmaf_rn(exp2f(__fmul_rn((float)level, S)), (float)H, -1.0f);  // exp2f(level * S) * H - 1.0f, one FMA (:138)
    g.resolution = (uint32_t)ceilf(g.scale) + 1;
    return g;
}

template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) k_grid_fwd(const float* __restrict__ inputs, const T* __restrict__ grid,
                                                  const int32_t* __restrict__ offsets, T* __restrict__ outputs, uint32_t B, uint32_t L,
                                                  float S, uint32_t H, T* __restrict__ dy_dx, uint32_t gridtype, bool align_corners,
                                                  uint32_t interp, int out_layout) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = (uint32_t)(tid / L), level = (uint32_t)(tid % L);
    if (b >= B) return;
    grid += (size_t)(uint32_t)offsets[level] * C;
    T* out = out_layout == 0 ? outputs + ((size_t)level * B + b) * C : outputs + ((size_t)b * L + level) * C;
    T* dd = dy_dx ? dy_dx + (size_t)b * D * L * C + (size_t)level * D * C : nullptr;

    float x[D];
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) { x[d] = __ldg(inputs + (size_t)b * D + d); oob |= (x[d] < 0 || x[d] > 1); }
    if (oob) {  // :110-135
#pragma unroll
        for (uint32_t c = 0; c < C; c++) out[c] = T(0.f);
        if (dd)
            for (uint32_t i = 0; i < D * C; i++) dd[i] = T(0.f);
        return;
    }
    const LevelGeom lg = level_geom(offsets, level, S, H);
    float pos[D], pos_deriv[D];
    uint32_t pos_grid[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        pos[d] = __fmaf_rn(x[d], lg.scale, align_corners ? 0.0f : 0.5f);
        const float fl = floorf(pos[d]);
        pos_grid[d] = (uint32_t)fl;
        pos[d] = __fsub_rn(pos[d], (float)pos_grid[d]);
        pos_deriv[d] = 1.0f;
        if (interp == 1) {
            pos_deriv[d] = 6 * pos[d] * (1.0f - pos[d]);
            pos[d] = pos[d] * pos[d] * (3.0f - 2.0f * pos[d]);
        }
    }
    T results[C];
#pragma unroll
    for (uint32_t c = 0; c < C; c++) results[c] = T(0.f);
#pragma unroll
    for (uint32_t idx = 0; idx < (1u << D); idx++) {
        float w = 1;
        uint32_t pgl[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            if ((idx & (1u << d)) == 0) { w = __fmul_rn(w, __fsub_rn(1.0f, pos[d])); pgl[d] = pos_grid[d]; }
            else { w = __fmul_rn(w, pos[d]); pgl[d] = pos_grid[d] + 1; }
        }
        const uint32_t index = grid_index<D, C>(gridtype, align_corners, lg.hashmap_size, lg.resolution, pgl);
#pragma unroll
        for (uint32_t c = 0; c < C; c++) results[c] = Acc<T>::mul_add(results[c], w, Acc<T>::load(grid + index + c));
    }
#pragma unroll
    for (uint32_t c = 0; c < C; c++) out[c] = results[c];

    if (dd) {  // :198-241 (pos_deriv is 1 for linear interpolation in every dimension)
#pragma unroll
        for (uint32_t gd = 0; gd < D; gd++) {
            float rg[C];
#pragma unroll
            for (uint32_t c = 0; c < C; c++) rg[c] = 0.f;
#pragma unroll
            for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
                float w = lg.scale;
                uint32_t pgl[D];
#pragma unroll
                for (uint32_t nd = 0; nd < D - 1; nd++) {
                    const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                    if ((idx & (1u << nd)) == 0) { w *= 1 - pos[d]; pgl[d] = pos_grid[d]; }
                    else { w *= pos[d]; pgl[d] = pos_grid[d] + 1; }
                }
                pgl[gd] = pos_grid[gd];
                const uint32_t il = grid_index<D, C>(gridtype, align_corners, lg.hashmap_size, lg.resolution, pgl);
                pgl[gd] = pos_grid[gd] + 1;
                const uint32_t ir = grid_index<D, C>(gridtype, align_corners, lg.hashmap_size, lg.resolution, pgl);
#pragma unroll
                for (uint32_t c = 0; c < C; c++)
                    rg[c] += w * ((float)Acc<T>::load(grid + ir + c) - (float)Acc<T>::load(grid + il + c)) * pos_deriv[gd];
            }
#pragma unroll
            for (uint32_t c = 0; c < C; c++) dd[gd * C + c] = T(rg[c]);
        }
    }
}

__device__ __forceinline__ void atomic_add_vec(float* p, const float* v, uint32_t n) {
    if (n == 2 && (((uintptr_t)p & 7u) == 0)) { atomicAdd(reinterpret_cast<float2*>(p), make_float2(v[0], v[1])); return; }
    for (uint32_t i = 0; i < n; i++) atomicAdd(p + i, v[i]);
}

// gridencoder.cu:245-337: scatter w * grad to the 2^D corners.
template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) k_grid_bwd(const T* __restrict__ grad, const float* __restrict__ inputs,
                                                  const int32_t* __restrict__ offsets, T* __restrict__ grad_grid, uint32_t B, uint32_t L,
                                                  float S, uint32_t H, uint32_t gridtype, bool align_corners, uint32_t interp,
                                                  int grad_layout) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = (uint32_t)(tid / L), level = (uint32_t)(tid % L);
    if (b >= B) return;
    grad_grid += (size_t)(uint32_t)offsets[level] * C;
    const T* g = grad_layout == 0 ? grad + ((size_t)level * B + b) * C : grad + ((size_t)b * L + level) * C;

    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) { x[d] = __ldg(inputs + (size_t)b * D + d); if (x[d] < 0 || x[d] > 1) return; }
    const LevelGeom lg = level_geom(offsets, level, S, H);
    float pos[D];
    uint32_t pos_grid[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        pos[d] = __fmaf_rn(x[d], lg.scale, align_corners ? 0.0f : 0.5f);
        pos_grid[d] = (uint32_t)floorf(pos[d]);
        pos[d] = __fsub_rn(pos[d], (float)pos_grid[d]);
        if (interp == 1) pos[d] = pos[d] * pos[d] * (3.0f - 2.0f * pos[d]);
    }
    float gc[C];
#pragma unroll
    for (uint32_t c = 0; c < C; c++) gc[c] = (float)g[c];
#pragma unroll
    for (uint32_t idx = 0; idx < (1u << D); idx++) {
        float w = 1;
        uint32_t pgl[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            if ((idx & (1u << d)) == 0) { w = __fmul_rn(w, __fsub_rn(1.0f, pos[d])); pgl[d] = pos_grid[d]; }
            else { w = __fmul_rn(w, pos[d]); pgl[d] = pos_grid[d] + 1; }
        }
        const uint32_t index = grid_index<D, C>(gridtype, align_corners, lg.hashmap_size, lg.resolution, pgl);
        if constexpr (sizeof(T) == 2) {
            if constexpr (C % 2 == 0) {
#pragma unroll
                for (uint32_t c = 0; c < C; c += 2) {
                    const __half2 v = __halves2half2(__float2half_rn(__fmul_rn(w, gc[c])), __float2half_rn(__fmul_rn(w, gc[c + 1])));
                    atomicAdd(reinterpret_cast<__half2*>(grad_grid + index + c), v);
                }
            } else {
                atomicAdd(reinterpret_cast<__half*>(grad_grid + index), __float2half_rn(__fmul_rn(w, gc[0])));
            }
        } else {
            float v[C];
#pragma unroll
            for (uint32_t c = 0; c < C; c++) v[c] = __fmul_rn(w, gc[c]);
            if constexpr (C % 2 == 0) {
#pragma unroll
                for (uint32_t c = 0; c < C; c += 2) atomic_add_vec(reinterpret_cast<float*>(grad_grid) + index + c, v + c, 2);
            } else {
                atomic_add_vec(reinterpret_cast<float*>(grad_grid) + index, v, C);
            }
        }
    }
}

// ---- fast path: D = 3, C = 2, hash grid, linear interpolation, [B, L*C] layout, L % 4 == 0 (the configuration of every
// encoder on the instance-field path, network_mask.py:34,76) ------------------------------------------------------------
// Thread = (sample, 4 consecutive levels); the 32 lanes of a warp hold 32 CONSECUTIVE samples at the same levels.  The
// sample stream is sorted by ray and by t, so neighbouring lanes read neighbouring cells and one L1 sector serves several
// lanes on the coarse and middle levels (the generic kernels above put the 16 levels of two samples in one warp: every
// lane of a load hits a different table).  Table entries are read as one 4-byte (fp16) / 8-byte (fp32) word per corner,
// a thread's 4 levels x 2 channels leave as one (fp16) or two (fp32) 16-byte stores.  Same arithmetic, same bits.
constexpr uint32_t kFastSamples = 64;   // samples per 256-thread CTA (x 4 level quarters)

template <typename T>
__global__ void __launch_bounds__(256) k_grid_fwd3x2(const float* __restrict__ inputs, const T* __restrict__ grid,
                                                     const int32_t* __restrict__ offsets, T* __restrict__ outputs, uint32_t B, uint32_t L,
                                                     float S, uint32_t H) {
    __shared__ field::LevelGeom lg[64];
    field::init_levels(lg, offsets, L, S, H, threadIdx.x);
    __syncthreads();
    const uint32_t LQ = L / 4, quarter = threadIdx.x / kFastSamples;
    const uint32_t b = blockIdx.x * kFastSamples + (threadIdx.x % kFastSamples);
    if (b >= B) return;
    float x[3];
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < 3; d++) { x[d] = __ldg(inputs + (size_t)b * 3 + d); oob |= (x[d] < 0 || x[d] > 1); }
    T* out = outputs + ((size_t)b * L + quarter * LQ) * 2;
    for (uint32_t l0 = 0; l0 < LQ; l0 += 4) {
        T res[8];
#pragma unroll
        for (uint32_t li = 0; li < 4; li++) {
            uint32_t idx[8];
            float w[8];
            const field::LevelGeom g = lg[quarter * LQ + l0 + li];
            const float xs[3] = {oob ? 0.f : x[0], oob ? 0.f : x[1], oob ? 0.f : x[2]};
            field::level_corners(xs, g, idx, w);
            if constexpr (sizeof(T) == 2) {
                const uint32_t* base = reinterpret_cast<const uint32_t*>(grid) + g.offset;
                uint32_t v[8];
#pragma unroll
                for (uint32_t c = 0; c < 8; c++) v[c] = __ldg(base + idx[c]);
                __half2 acc = __float2half2_rn(0.f);
#pragma unroll
                for (uint32_t c = 0; c < 8; c++) {   // product rounded to half, accumulated in half (gridencoder.cu:161-185)
                    const float2 f = __half22float2(field::bits_h2(v[c]));
                    acc = __hadd2(acc, __floats2half2_rn(__fmul_rn(w[c], f.x), __fmul_rn(w[c], f.y)));
                }
                if (oob) acc = __float2half2_rn(0.f);
                res[2 * li] = __low2half(acc); res[2 * li + 1] = __high2half(acc);
            } else {
                const float2* base = reinterpret_cast<const float2*>(grid) + g.offset;
                float2 v[8];
#pragma unroll
                for (uint32_t c = 0; c < 8; c++) v[c] = __ldg(base + idx[c]);
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (uint32_t c = 0; c < 8; c++) { a0 = fmaf(w[c], v[c].x, a0); a1 = fmaf(w[c], v[c].y, a1); }
                res[2 * li] = oob ? 0.f : a0; res[2 * li + 1] = oob ? 0.f : a1;
            }
        }
        if constexpr (sizeof(T) == 2) {
            *reinterpret_cast<uint4*>(out + l0 * 2) = *reinterpret_cast<const uint4*>(res);
        } else {
            reinterpret_cast<float4*>(out + l0 * 2)[0] = reinterpret_cast<const float4*>(res)[0];
            reinterpret_cast<float4*>(out + l0 * 2)[1] = reinterpret_cast<const float4*>(res)[1];
        }
    }
}

// Backward fast path, same mapping.  On levels whose resolution is at most kRunRes consecutive samples of a ray mostly fall
// into the SAME cell: their atomics would serialise on a handful of L2 addresses (the generic kernel: 4 % of HBM peak).
// Those levels reduce each run of equal cells inside the warp first (segmented shuffle reduction in fp32 towards the run's
// first lane) and only run heads issue atomics; the fp16 table gradient then receives ONE rounding of the run's fp32 sum
// instead of a sum of individually rounded halves (closer to the exact gradient than the reference's own ordering noise).
constexpr float kRunRes = 512.0f;

template <typename T>
__global__ void __launch_bounds__(256) k_grid_bwd3x2(const T* __restrict__ grad, const float* __restrict__ inputs,
                                                     const int32_t* __restrict__ offsets, T* __restrict__ grad_grid, uint32_t B, uint32_t L,
                                                     float S, uint32_t H) {
    __shared__ field::LevelGeom lg[64];
    field::init_levels(lg, offsets, L, S, H, threadIdx.x);
    __syncthreads();
    const uint32_t LQ = L / 4, quarter = threadIdx.x / kFastSamples, lane = threadIdx.x & 31u;
    const uint32_t b = blockIdx.x * kFastSamples + (threadIdx.x % kFastSamples);
    float x[3] = {0.f, 0.f, 0.f};
    bool ok = b < B;
    if (ok) {
#pragma unroll
        for (uint32_t d = 0; d < 3; d++) { x[d] = __ldg(inputs + (size_t)b * 3 + d); ok &= !(x[d] < 0 || x[d] > 1); }
    }
    if (!ok) { x[0] = x[1] = x[2] = 0.f; }
    const T* g_row = grad + ((size_t)(b < B ? b : 0) * L + quarter * LQ) * 2;
    for (uint32_t li = 0; li < LQ; li++) {
        const field::LevelGeom g = lg[quarter * LQ + li];
        float g0 = 0.f, g1 = 0.f;
        if (ok) { g0 = (float)g_row[2 * li]; g1 = (float)g_row[2 * li + 1]; }
        uint32_t idx[8];
        float w[8];
        field::level_corners(x, g, idx, w);
        float a0[8], a1[8];
#pragma unroll
        for (uint32_t c = 0; c < 8; c++) { a0[c] = __fmul_rn(w[c], g0); a1[c] = __fmul_rn(w[c], g1); }
        bool head = true;
        if (g.scale <= kRunRes) {   // warp-uniform: every lane of a warp works on the same level
            const uint32_t p0 = __shfl_up_sync(0xffffffffu, idx[0], 1), p7 = __shfl_up_sync(0xffffffffu, idx[7], 1);
            head = lane == 0 || p0 != idx[0] || p7 != idx[7];
            const uint32_t heads = __ballot_sync(0xffffffffu, head);
            const uint32_t after = lane == 31 ? 0u : (heads >> (lane + 1));
            const uint32_t next_head = after ? (lane + 1 + (uint32_t)__ffs(after) - 1u) : 32u;
#pragma unroll
            for (uint32_t off = 1; off < 32; off <<= 1) {
                const bool take = lane + off < next_head;
#pragma unroll
                for (uint32_t c = 0; c < 8; c++) {
                    const float o0 = __shfl_down_sync(0xffffffffu, a0[c], off), o1 = __shfl_down_sync(0xffffffffu, a1[c], off);
                    if (take) { a0[c] += o0; a1[c] += o1; }
                }
            }
        }
        if (head) {
#pragma unroll
            for (uint32_t c = 0; c < 8; c++) {
                if (a0[c] == 0.f && a1[c] == 0.f) continue;
                if constexpr (sizeof(T) == 2)
                    atomicAdd(reinterpret_cast<__half2*>(grad_grid) + g.offset + idx[c], __floats2half2_rn(a0[c], a1[c]));
                else
                    atomicAdd(reinterpret_cast<float2*>(grad_grid) + g.offset + idx[c], make_float2(a0[c], a1[c]));
            }
        }
    }
}

template <typename T>
bool fast_path_ok(uint32_t D, uint32_t C, uint32_t L, uint32_t gridtype, bool ac, uint32_t interp, int layout, const void* table, const void* act) {
    return D == 3 && C == 2 && L % 4 == 0 && L <= 64 && gridtype == 0 && !ac && interp == 0 && layout == 1 &&
           ((uintptr_t)table & (2 * sizeof(T) - 1)) == 0 && ((uintptr_t)act & 15u) == 0 && ((L / 4) * 2 * sizeof(T)) % 16 == 0;
}

// gridencoder.cu:340-366
template <typename T>
__global__ void k_input_bwd(const T* __restrict__ grad, const T* __restrict__ dy_dx, T* __restrict__ grad_inputs, uint32_t B, uint32_t D,
                            uint32_t C, uint32_t L, int grad_layout) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    dy_dx += (size_t)b * L * D * C;
    float result = 0;
    for (uint32_t l = 0; l < L; l++)
        for (uint32_t ch = 0; ch < C; ch++) {
            const T gv = grad_layout == 0 ? grad[((size_t)l * B + b) * C + ch] : grad[((size_t)b * L + l) * C + ch];
            result += (float)gv * (float)dy_dx[l * D * C + d * C + ch];
        }
    grad_inputs[t] = T(result);
}

template <typename T, uint32_t D>
int launch_fwd_c(const float* inputs, const T* emb, const int32_t* offsets, T* out, uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H,
                 T* dy_dx, uint32_t gridtype, bool ac, uint32_t interp, int layout, cudaStream_t st) {
    const unsigned int blocks = div_up((unsigned long long)B * L, 256);
    switch (C) {
        case 1: k_grid_fwd<T, D, 1><<<blocks, 256, 0, st>>>(inputs, emb, offsets, out, B, L, S, H, dy_dx, gridtype, ac, interp, layout); break;
        case 2: k_grid_fwd<T, D, 2><<<blocks, 256, 0, st>>>(inputs, emb, offsets, out, B, L, S, H, dy_dx, gridtype, ac, interp, layout); break;
        case 4: k_grid_fwd<T, D, 4><<<blocks, 256, 0, st>>>(inputs, emb, offsets, out, B, L, S, H, dy_dx, gridtype, ac, interp, layout); break;
        case 8: k_grid_fwd<T, D, 8><<<blocks, 256, 0, st>>>(inputs, emb, offsets, out, B, L, S, H, dy_dx, gridtype, ac, interp, layout); break;
        default: return INERF_ERR_UNSUPPORTED;
    }
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}
template <typename T, uint32_t D>
int launch_bwd_c(const T* grad, const float* inputs, const int32_t* offsets, T* gg, uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H,
                 uint32_t gridtype, bool ac, uint32_t interp, int layout, cudaStream_t st) {
    const unsigned int blocks = div_up((unsigned long long)B * L, 256);
    switch (C) {
        case 1: k_grid_bwd<T, D, 1><<<blocks, 256, 0, st>>>(grad, inputs, offsets, gg, B, L, S, H, gridtype, ac, interp, layout); break;
        case 2: k_grid_bwd<T, D, 2><<<blocks, 256, 0, st>>>(grad, inputs, offsets, gg, B, L, S, H, gridtype, ac, interp, layout); break;
        case 4: k_grid_bwd<T, D, 4><<<blocks, 256, 0, st>>>(grad, inputs, offsets, gg, B, L, S, H, gridtype, ac, interp, layout); break;
        case 8: k_grid_bwd<T, D, 8><<<blocks, 256, 0, st>>>(grad, inputs, offsets, gg, B, L, S, H, gridtype, ac, interp, layout); break;
        default: return INERF_ERR_UNSUPPORTED;
    }
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

// gridencoder.cu:504-603: total-variation gradient on the grid vertices hit by `inputs`, added into `grad` (the embeddings'
// gradient, after loss.backward()).  Thread = (point, level), levels on consecutive threads as in the encode kernels.
// Per channel: r = sum over the 2D face neighbours n of (v - v_n), s = sum of (v - v_n)^2, grad[v] += weight / (2D) * r * rsqrt(s + 1e-9).
template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) k_grad_tv(const float* __restrict__ inputs, const T* __restrict__ grid, T* __restrict__ grad,
                                                 const int32_t* __restrict__ offsets, float weight, uint32_t B, uint32_t L, float S, uint32_t H,
                                                 uint32_t gridtype, bool align_corners) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)B * L) return;
    const uint32_t b = (uint32_t)(t / L), level = (uint32_t)(t - (uint64_t)b * L);
    const float* x = inputs + (size_t)b * D;
#pragma unroll
    for (uint32_t d = 0; d < D; d++)
        if (x[d] < 0.f || x[d] > 1.f) return;
    const LevelGeom g = level_geom(offsets, level, S, H);
    grid += (size_t)offsets[level] * C;
    grad += (size_t)offsets[level] * C;
    uint32_t pos_grid[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) pos_grid[d] = (uint32_t)floorf(__fmaf_rn(x[d], g.scale, align_corners ? 0.0f : 0.5f));
    const uint32_t index = grid_index<D, C>(gridtype, align_corners, g.hashmap_size, g.resolution, pos_grid);
    float results[C], idelta[C], centre[C];
#pragma unroll
    for (uint32_t ch = 0; ch < C; ch++) { results[ch] = 0.f; idelta[ch] = 0.f; centre[ch] = (float)grid[index + ch]; }
    const float w = (float)(T)(weight / (float)(2 * D));   // `scalar_t w = weight / (2 * D)`
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        const uint32_t cur = pos_grid[d];
#pragma unroll
        for (int side = 0; side < 2; side++) {
            if (side == 0 ? cur < g.resolution : cur > 0) {
                pos_grid[d] = side == 0 ? cur + 1 : cur - 1;
                const uint32_t nb = grid_index<D, C>(gridtype, align_corners, g.hashmap_size, g.resolution, pos_grid);
#pragma unroll
                for (uint32_t ch = 0; ch < C; ch++) {
                    const float gv = (float)(T)(centre[ch] - (float)grid[nb + ch]);     // the reference keeps every intermediate in scalar_t
                    results[ch] = (float)(T)(results[ch] + gv);
                    idelta[ch] = (float)(T)(idelta[ch] + (float)(T)(gv * gv));
                }
            }
        }
        pos_grid[d] = cur;
    }
#pragma unroll
    for (uint32_t ch = 0; ch < C; ch++) {
        const float v = (float)(T)((float)(T)(w * results[ch]) * rsqrtf(idelta[ch] + 1e-9f));
        if constexpr (std::is_same<T, float>::value) atomicAdd(grad + index + ch, v);
        else atomicAdd(grad + index + ch, __float2half_rn(v));
    }
}

template <typename T, uint32_t D>
int launch_tv_c(const float* inputs, const T* emb, T* grad, const int32_t* offsets, float weight, uint32_t B, uint32_t C, uint32_t L, float S,
                uint32_t H, uint32_t gridtype, bool ac, cudaStream_t st) {
    const unsigned blocks = div_up((unsigned long long)B * L, 256);
    switch (C) {
        case 1: k_grad_tv<T, D, 1><<<blocks, 256, 0, st>>>(inputs, emb, grad, offsets, weight, B, L, S, H, gridtype, ac); break;
        case 2: k_grad_tv<T, D, 2><<<blocks, 256, 0, st>>>(inputs, emb, grad, offsets, weight, B, L, S, H, gridtype, ac); break;
        case 4: k_grad_tv<T, D, 4><<<blocks, 256, 0, st>>>(inputs, emb, grad, offsets, weight, B, L, S, H, gridtype, ac); break;
        case 8: k_grad_tv<T, D, 8><<<blocks, 256, 0, st>>>(inputs, emb, grad, offsets, weight, B, L, S, H, gridtype, ac); break;
        default: return INERF_ERR_UNSUPPORTED;
    }
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}
template <typename T>
int tv_t(const float* inputs, const void* emb, void* grad, const int32_t* offsets, float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
         float S, uint32_t H, uint32_t gridtype, bool ac, cudaStream_t st) {
    switch (D) {
        case 2: return launch_tv_c<T, 2>(inputs, (const T*)emb, (T*)grad, offsets, weight, B, C, L, S, H, gridtype, ac, st);
        case 3: return launch_tv_c<T, 3>(inputs, (const T*)emb, (T*)grad, offsets, weight, B, C, L, S, H, gridtype, ac, st);
        default: return INERF_ERR_UNSUPPORTED;
    }
}

template <typename T>
int fwd_t(const float* inputs, const void* emb, const int32_t* offsets, void* out, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
          uint32_t H, void* dy_dx, uint32_t gridtype, bool ac, uint32_t interp, int layout, cudaStream_t st) {
    if (dy_dx == nullptr && fast_path_ok<T>(D, C, L, gridtype, ac, interp, layout, emb, out)) {
        k_grid_fwd3x2<T><<<div_up(B, kFastSamples), 256, 0, st>>>(inputs, (const T*)emb, offsets, (T*)out, B, L, S, H);
        INERF_LAUNCH_CHECK();
        return INERF_OK;
    }
    switch (D) {
        case 2: return launch_fwd_c<T, 2>(inputs, (const T*)emb, offsets, (T*)out, B, C, L, S, H, (T*)dy_dx, gridtype, ac, interp, layout, st);
        case 3: return launch_fwd_c<T, 3>(inputs, (const T*)emb, offsets, (T*)out, B, C, L, S, H, (T*)dy_dx, gridtype, ac, interp, layout, st);
        default: return INERF_ERR_UNSUPPORTED;
    }
}
template <typename T>
int bwd_t(const void* grad, const float* inputs, const int32_t* offsets, void* gg, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
          uint32_t H, const void* dy_dx, void* grad_inputs, uint32_t gridtype, bool ac, uint32_t interp, int layout, cudaStream_t st) {
    int e;
    if (fast_path_ok<T>(D, C, L, gridtype, ac, interp, layout, gg, grad)) {
        k_grid_bwd3x2<T><<<div_up(B, kFastSamples), 256, 0, st>>>((const T*)grad, inputs, offsets, (T*)gg, B, L, S, H);
        e = INERF_OK;
        INERF_LAUNCH_CHECK();
    } else
    switch (D) {
        case 2: e = launch_bwd_c<T, 2>((const T*)grad, inputs, offsets, (T*)gg, B, C, L, S, H, gridtype, ac, interp, layout, st); break;
        case 3: e = launch_bwd_c<T, 3>((const T*)grad, inputs, offsets, (T*)gg, B, C, L, S, H, gridtype, ac, interp, layout, st); break;
        default: return INERF_ERR_UNSUPPORTED;
    }
    if (e) return e;
    if (dy_dx && grad_inputs) {
        k_input_bwd<T><<<div_up((unsigned long long)B * D, 256), 256, 0, st>>>((const T*)grad, (const T*)dy_dx, (T*)grad_inputs, B, D, C, L, layout);
        INERF_LAUNCH_CHECK();
    }
    return INERF_OK;
}

}  // namespace

extern "C" int inerf_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets, void* outputs, uint32_t B,
                                         uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, void* dy_dx, uint32_t gridtype,
                                         int align_corners, uint32_t interp, int dtype, int out_layout, void* stream) {
    if (L == 0 || L > 64 || gridtype > 1 || interp > 1 || (out_layout != 0 && out_layout != 1)) return INERF_ERR_SIZE;
    if (B == 0) return INERF_OK;
    INERF_REQUIRE(inputs); INERF_REQUIRE(embeddings); INERF_REQUIRE(offsets); INERF_REQUIRE(outputs);
    if (dtype == INERF_F32)
        return fwd_t<float>(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners != 0, interp, out_layout, (cudaStream_t)stream);
    if (dtype == INERF_F16)
        return fwd_t<__half>(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners != 0, interp, out_layout, (cudaStream_t)stream);
    return INERF_ERR_UNSUPPORTED;
}

extern "C" int inerf_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings, const int32_t* offsets,
                                          void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                          const void* dy_dx, void* grad_inputs, uint32_t gridtype, int align_corners, uint32_t interp,
                                          int dtype, int grad_layout, void* stream) {
    (void)embeddings;  // not read by the reference kernel either (gridencoder.cu:245-337)
    if (L == 0 || L > 64 || gridtype > 1 || interp > 1 || (grad_layout != 0 && grad_layout != 1)) return INERF_ERR_SIZE;
    if (B == 0) return INERF_OK;
    INERF_REQUIRE(grad); INERF_REQUIRE(inputs); INERF_REQUIRE(offsets); INERF_REQUIRE(grad_embeddings);
    if (dtype == INERF_F32)
        return bwd_t<float>(grad, inputs, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs, gridtype, align_corners != 0, interp, grad_layout, (cudaStream_t)stream);
    if (dtype == INERF_F16)
        return bwd_t<__half>(grad, inputs, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs, gridtype, align_corners != 0, interp, grad_layout, (cudaStream_t)stream);
    return INERF_ERR_UNSUPPORTED;
}

extern "C" int inerf_grad_total_variation(const float* inputs, const void* embeddings, void* grad, const int32_t* offsets, float weight,
                                          uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                          int align_corners, int dtype, void* stream) {
    if (L == 0 || L > 64 || gridtype > 1) return INERF_ERR_SIZE;
    if (B == 0) return INERF_OK;
    INERF_REQUIRE(inputs); INERF_REQUIRE(embeddings); INERF_REQUIRE(grad); INERF_REQUIRE(offsets);
    if (dtype == INERF_F32) return tv_t<float>(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners != 0, (cudaStream_t)stream);
    if (dtype == INERF_F16) return tv_t<__half>(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners != 0, (cudaStream_t)stream);
    return INERF_ERR_UNSUPPORTED;
}
