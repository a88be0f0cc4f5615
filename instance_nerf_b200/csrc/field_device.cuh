// Device-side building blocks of the fused instance field (shared by field_fused.cu and render_fused.cu):
// hash-grid gathers for both tables straight into UMMA operand tiles, SH degree 4, and the tcgen05 MLP
// chain with its TMEM epilogues.
//
// Arithmetic contract (nerf/network_mask.py:119-158 under torch autocast, the reference's `-O` preset):
//   * hash encode: weights fp32, every product rounded to fp16, accumulated in fp16 (gridencoder.cu:161-185);
//   * every nn.Linear: fp16 operands, fp32 accumulation (TMEM), ONE rounding of the output to fp16;
//   * sigma = exp(float(h0)) in fp32 (activation.py:5-18), rgb = fp16(sigmoid(float(h))), logits = fp16 values.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace field {

constexpr int kTile = 128;        // samples per tile = TMEM lanes = UMMA M
constexpr int kThreads = 256;     // 8 warps: gathers split by level half, epilogues split by column half
constexpr uint32_t kLBO = 128;    // K-direction core-matrix stride (bytes)

// ---- packed weight blob (fp16, UMMA B-operand layout, see inerf_field_pack_weights) ----
struct WeightLayout {
    uint32_t s0, s1, c0, c1, c2, m0, m1, m2, total, Kp;
};
__host__ __device__ inline WeightLayout weight_layout(uint32_t K) {
    WeightLayout w;
    w.Kp = (K + 15u) & ~15u;
    w.s0 = 0;                       // [64 x 32]
    w.s1 = w.s0 + 64 * 32 * 2;      // [16 x 64]
    w.c0 = w.s1 + 16 * 64 * 2;      // [64 x 32]   (31 inputs + 1 zero column)
    w.c1 = w.c0 + 64 * 32 * 2;      // [64 x 64]
    w.c2 = w.c1 + 64 * 64 * 2;      // [16 x 64]   (3 outputs + 13 zero rows)
    w.m0 = w.c2 + 16 * 64 * 2;      // [64 x 48]   (47 inputs + 1 zero column)
    w.m1 = w.m0 + 64 * 48 * 2;      // [64 x 64]
    w.m2 = w.m1 + 64 * 64 * 2;      // [Kp x 64]
    w.total = w.m2 + w.Kp * 64 * 2;
    return w;
}

// ---- shared-memory plan -------------------------------------------------------------------
struct LevelGeom {
    float scale;
    uint32_t resolution, size, offset;
};
struct Smem {
    // activations (A operands), fp16, K-major no-swizzle: sbo = (Kdim / 8) * 128
    static constexpr uint32_t A_es = 0;                        // [128 x 32] sigma-table features
    static constexpr uint32_t A_ci = A_es + kTile * 32 * 2;    // [128 x 32] SH16 | geo15 | 0
    static constexpr uint32_t A_mi = A_ci + kTile * 32 * 2;    // [128 x 48] mask-table features 32 | geo15 | 0
    static constexpr uint32_t A_h1 = A_mi + kTile * 48 * 2;    // [128 x 64] hidden (sigma, then colour)
    static constexpr uint32_t A_h2 = A_h1 + kTile * 64 * 2;    // [128 x 64] hidden (mask)
    static constexpr uint32_t W = A_h2 + kTile * 64 * 2;       // weight blob
    static __host__ __device__ uint32_t misc(uint32_t K) { return W + weight_layout(K).total; }
    // misc: LevelGeom[16] (256 B) | mbarrier (8 B) | tmem slot (4 B) | ...
    static __host__ __device__ uint32_t bytes(uint32_t K) { return misc(K) + 16 * sizeof(LevelGeom) + 64; }
};
constexpr uint32_t sbo_of(uint32_t kdim) { return (kdim / 8) * 128; }

// TMEM column plan (fp32 accumulators); 256 columns allocated -> two CTAs per SM.
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t D_a = 0;     // 64: sigma hidden / colour hidden
constexpr uint32_t D_b = 64;    // 64: mask hidden
constexpr uint32_t D_c = 128;   // 16: sigma-net output / rgb
constexpr uint32_t D_d = 144;   // up to 64: logits

__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// level geometry exactly as the op-level kernel (gridencode.cu level_geom / gridencoder.cu:137-139)
__device__ __forceinline__ void init_levels(LevelGeom* lg, const int32_t* __restrict__ offsets, uint32_t L, float S, uint32_t H) {
    if (threadIdx.x < L) {
        const uint32_t l = threadIdx.x;
        LevelGeom g;
        g.offset = (uint32_t)offsets[l];
        g.size = (uint32_t)(offsets[l + 1] - offsets[l]);
        g.scale = __fmaf_rn(exp2f(__fmul_rn((float)l, S)), (float)H, -1.0f);
        g.resolution = (uint32_t)ceilf(g.scale) + 1;
        lg[l] = g;
    }
}

__device__ __forceinline__ __half acc_half(__half acc, float w, __half g) {
    return __float2half_rn(__fadd_rn(__half2float(acc), __half2float(__float2half_rn(__fmul_rn(w, __half2float(g))))));
}

// Encode 8 consecutive levels [l0, l0+8) of one point from BOTH tables (they share geometry, so the 8 corner
// indices are computed once) and store the 16+16 fp16 features as 2+2 16-byte core-matrix rows.
__device__ __forceinline__ void encode8(const float x01[3], bool oob, uint32_t l0, const LevelGeom* __restrict__ lg,
                                        const __half2* __restrict__ tab_s, const __half2* __restrict__ tab_m,
                                        uint8_t* smem, uint32_t row) {
    uint32_t fs[8], fm[8];  // packed half2 per level
#pragma unroll
    for (uint32_t li = 0; li < 8; li++) {
        const LevelGeom g = lg[l0 + li];
        float pos[3];
        uint32_t pg[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            pos[d] = __fmaf_rn(x01[d], g.scale, 0.5f);
            pg[d] = (uint32_t)floorf(pos[d]);
            pos[d] = __fsub_rn(pos[d], (float)pg[d]);
        }
        const uint32_t r1 = g.resolution + 1;
        const bool dense = (r1 <= g.size) && (r1 * r1 <= g.size) && ((uint64_t)r1 * r1 * r1 <= g.size);
        const bool pow2 = (g.size & (g.size - 1)) == 0;
        uint32_t idx[8];
        float w[8];
#pragma unroll
        for (uint32_t c = 0; c < 8; c++) {
            const uint32_t cx = pg[0] + (c & 1u), cy = pg[1] + ((c >> 1) & 1u), cz = pg[2] + ((c >> 2) & 1u);
            float ww = (c & 1u) ? pos[0] : __fsub_rn(1.0f, pos[0]);
            ww = __fmul_rn(ww, (c & 2u) ? pos[1] : __fsub_rn(1.0f, pos[1]));
            ww = __fmul_rn(ww, (c & 4u) ? pos[2] : __fsub_rn(1.0f, pos[2]));
            w[c] = ww;
            uint32_t index = dense ? cx + cy * r1 + cz * r1 * r1 : (cx ^ (cy * 2654435761u) ^ (cz * 805459861u));
            index = pow2 ? (index & (g.size - 1)) : (index % g.size);
            idx[c] = g.offset + index;
        }
        __half2 vs[8], vm[8];
#pragma unroll
        for (uint32_t c = 0; c < 8; c++) { vs[c] = __ldg(tab_s + idx[c]); vm[c] = __ldg(tab_m + idx[c]); }
        __half s0 = __float2half_rn(0.f), s1 = s0, m0 = s0, m1 = s0;
#pragma unroll
        for (uint32_t c = 0; c < 8; c++) {
            s0 = acc_half(s0, w[c], __low2half(vs[c])); s1 = acc_half(s1, w[c], __high2half(vs[c]));
            m0 = acc_half(m0, w[c], __low2half(vm[c])); m1 = acc_half(m1, w[c], __high2half(vm[c]));
        }
        fs[li] = oob ? 0u : pack_half2(s0, s1);
        fm[li] = oob ? 0u : pack_half2(m0, m1);
    }
    // features k = 2*level + ch: levels l0..l0+3 -> one 16-byte chunk, l0+4..l0+7 -> the next
    const uint32_t k0 = l0 * 2;
    *reinterpret_cast<uint4*>(smem + Smem::A_es + umma::tile_off(row, k0, kLBO, sbo_of(32))) = make_uint4(fs[0], fs[1], fs[2], fs[3]);
    *reinterpret_cast<uint4*>(smem + Smem::A_es + umma::tile_off(row, k0 + 8, kLBO, sbo_of(32))) = make_uint4(fs[4], fs[5], fs[6], fs[7]);
    *reinterpret_cast<uint4*>(smem + Smem::A_mi + umma::tile_off(row, k0, kLBO, sbo_of(48))) = make_uint4(fm[0], fm[1], fm[2], fm[3]);
    *reinterpret_cast<uint4*>(smem + Smem::A_mi + umma::tile_off(row, k0 + 8, kLBO, sbo_of(48))) = make_uint4(fm[4], fm[5], fm[6], fm[7]);
}

// SH degree 4 (same polynomials as shencode.cu) rounded to fp16 into A_ci columns 0..15
__device__ __forceinline__ void sh16_to_smem(float x, float y, float z, uint8_t* smem, uint32_t row) {
    float o[16];
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
    uint32_t p[8];
#pragma unroll
    for (int i = 0; i < 8; i++) p[i] = pack_half2(__float2half_rn(o[2 * i]), __float2half_rn(o[2 * i + 1]));
    *reinterpret_cast<uint4*>(smem + Smem::A_ci + umma::tile_off(row, 0, kLBO, sbo_of(32))) = make_uint4(p[0], p[1], p[2], p[3]);
    *reinterpret_cast<uint4*>(smem + Smem::A_ci + umma::tile_off(row, 8, kLBO, sbo_of(32))) = make_uint4(p[4], p[5], p[6], p[7]);
}

// ---- MMA issue (one thread) -------------------------------------------------------------------
// D[128 x N] (+)= A[128 x Kdim] * B[N x Kdim]^T ; A at smem+a_off (sbo_of(Kdim)), B at smem+b_off (sbo_of(Kdim))
__device__ __forceinline__ void issue_gemm(uint32_t smem_base, uint32_t a_off, uint32_t b_off, uint32_t Kdim, uint32_t N,
                                           uint32_t tmem_d) {
    const uint32_t idesc = umma::make_idesc_f16(128, N);
    const uint32_t sbo = sbo_of(Kdim);
    for (uint32_t k = 0; k < Kdim / 16; k++) {
        const uint64_t da = umma::make_desc(smem_base + a_off + k * 2 * kLBO, kLBO, sbo);
        const uint64_t db = umma::make_desc(smem_base + b_off + k * 2 * kLBO, kLBO, sbo);
        umma::mma_f16(tmem_d, da, db, idesc, k > 0);
    }
}

// ---- epilogues -----------------------------------------------------------------------------
// 64-wide hidden layer: TMEM -> fp16 round -> ReLU -> next A tile.  Warp w: rows 32*(w&3).., columns 32*(w>>2)..
__device__ __forceinline__ void epilogue_hidden(uint32_t tmem_d, uint8_t* smem, uint32_t a_off) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row = (warp & 3u) * 32u + lane, c0 = (warp >> 2) * 32u;
    const uint32_t taddr = tmem_d + (((warp & 3u) * 32u) << 16) + c0;
    uint32_t v[2][16];
    umma::tmem_ld16(taddr, v[0]);
    umma::tmem_ld16(taddr + 16, v[1]);
    umma::tmem_ld_wait();
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint32_t p[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const __half a = __float2half_rn(fmaxf(__uint_as_float(v[h][2 * i]), 0.f));
            const __half b = __float2half_rn(fmaxf(__uint_as_float(v[h][2 * i + 1]), 0.f));
            p[i] = pack_half2(a, b);
        }
        const uint32_t k = c0 + h * 16;
        *reinterpret_cast<uint4*>(smem + a_off + umma::tile_off(row, k, kLBO, sbo_of(64))) = make_uint4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<uint4*>(smem + a_off + umma::tile_off(row, k + 8, kLBO, sbo_of(64))) = make_uint4(p[4], p[5], p[6], p[7]);
    }
}

// sigma-net output (16 columns): h0 -> sigma (returned, valid for warps 0..3), geo15 -> A_ci[:,16:32] and A_mi[:,32:48]
__device__ __forceinline__ float epilogue_sigma(uint32_t tmem_base, uint8_t* smem, float density_scale) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float sigma = 0.f;
    if (warp < 4) {
        const uint32_t row = warp * 32u + lane;
        uint32_t v[16];
        umma::tmem_ld16(tmem_base + D_c + ((warp * 32u) << 16), v);
        umma::tmem_ld_wait();
        __half h[16];
#pragma unroll
        for (int i = 0; i < 16; i++) h[i] = __float2half_rn(__uint_as_float(v[i]));
        sigma = __fmul_rn(expf(__half2float(h[0])), density_scale);
        const __half z = __float2half_rn(0.f);
        const uint4 g0 = make_uint4(pack_half2(h[1], h[2]), pack_half2(h[3], h[4]), pack_half2(h[5], h[6]), pack_half2(h[7], h[8]));
        const uint4 g1 = make_uint4(pack_half2(h[9], h[10]), pack_half2(h[11], h[12]), pack_half2(h[13], h[14]), pack_half2(h[15], z));
        *reinterpret_cast<uint4*>(smem + Smem::A_ci + umma::tile_off(row, 16, kLBO, sbo_of(32))) = g0;
        *reinterpret_cast<uint4*>(smem + Smem::A_ci + umma::tile_off(row, 24, kLBO, sbo_of(32))) = g1;
        *reinterpret_cast<uint4*>(smem + Smem::A_mi + umma::tile_off(row, 32, kLBO, sbo_of(48))) = g0;
        *reinterpret_cast<uint4*>(smem + Smem::A_mi + umma::tile_off(row, 40, kLBO, sbo_of(48))) = g1;
    }
    return sigma;
}

// colour output: 3 of 16 columns -> fp16 round -> sigmoid (fp32) -> fp16 round.  Valid for warps 0..3.
__device__ __forceinline__ void epilogue_rgb(uint32_t tmem_base, float rgb[3]) {
    const uint32_t warp = threadIdx.x >> 5;
    uint32_t v[16];
    umma::tmem_ld16(tmem_base + D_c + (((warp & 3u) * 32u) << 16), v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float h = __half2float(__float2half_rn(__uint_as_float(v[i])));
        rgb[i] = __half2float(__float2half_rn(__fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-h)))));
    }
}

// The whole MLP chain for one 128-sample tile whose encoder features / SH are already in shared memory
// (written by all threads, followed by fence_async_smem + __syncthreads by the caller).
// After return: D_c holds the colour pre-activations, D_d the logits (if with_masks); returns sigma for warps 0..3.
// `phase` is the running mbarrier parity, updated in place.
// `on_sigma(sigma)` runs in every thread right after the sigma-net epilogue (sigma is valid in warps 0..3) and
// before the next block-wide barrier, so whatever it writes to shared memory is visible after the chain.
template <typename OnSigma>
__device__ __forceinline__ float mlp_chain(uint8_t* smem, uint32_t tmem_base, uint64_t* bar, uint32_t& phase, uint32_t K,
                                           float density_scale, bool with_masks, OnSigma&& on_sigma) {
    const uint32_t sbase = umma::smem_u32(smem);
    const WeightLayout wl = weight_layout(K);
    const bool issuer = threadIdx.x == 0;
    // sigma layer 0
    if (issuer) {
        umma::fence_after_sync();
        issue_gemm(sbase, Smem::A_es, Smem::W + wl.s0, 32, 64, tmem_base + D_a);
        umma::commit(bar);
    }
    __syncwarp();
    umma::mbar_wait(bar, phase); phase ^= 1u;
    umma::fence_after_sync();
    epilogue_hidden(tmem_base + D_a, smem, Smem::A_h1);
    umma::fence_async_smem(); umma::fence_before_sync();
    __syncthreads();
    // sigma layer 1
    if (issuer) {
        umma::fence_after_sync();
        issue_gemm(sbase, Smem::A_h1, Smem::W + wl.s1, 64, 16, tmem_base + D_c);
        umma::commit(bar);
    }
    __syncwarp();
    umma::mbar_wait(bar, phase); phase ^= 1u;
    umma::fence_after_sync();
    const float sigma = epilogue_sigma(tmem_base, smem, density_scale);
    on_sigma(sigma);
    umma::fence_async_smem(); umma::fence_before_sync();
    __syncthreads();
    // colour layer 0 + mask layer 0
    if (issuer) {
        umma::fence_after_sync();
        issue_gemm(sbase, Smem::A_ci, Smem::W + wl.c0, 32, 64, tmem_base + D_a);
        if (with_masks) issue_gemm(sbase, Smem::A_mi, Smem::W + wl.m0, 48, 64, tmem_base + D_b);
        umma::commit(bar);
    }
    __syncwarp();
    umma::mbar_wait(bar, phase); phase ^= 1u;
    umma::fence_after_sync();
    epilogue_hidden(tmem_base + D_a, smem, Smem::A_h1);
    if (with_masks) epilogue_hidden(tmem_base + D_b, smem, Smem::A_h2);
    umma::fence_async_smem(); umma::fence_before_sync();
    __syncthreads();
    // colour layer 1 + mask layer 1
    if (issuer) {
        umma::fence_after_sync();
        issue_gemm(sbase, Smem::A_h1, Smem::W + wl.c1, 64, 64, tmem_base + D_a);
        if (with_masks) issue_gemm(sbase, Smem::A_h2, Smem::W + wl.m1, 64, 64, tmem_base + D_b);
        umma::commit(bar);
    }
    __syncwarp();
    umma::mbar_wait(bar, phase); phase ^= 1u;
    umma::fence_after_sync();
    epilogue_hidden(tmem_base + D_a, smem, Smem::A_h1);
    if (with_masks) epilogue_hidden(tmem_base + D_b, smem, Smem::A_h2);
    umma::fence_async_smem(); umma::fence_before_sync();
    __syncthreads();
    // colour layer 2 + mask layer 2
    if (issuer) {
        umma::fence_after_sync();
        issue_gemm(sbase, Smem::A_h1, Smem::W + wl.c2, 64, 16, tmem_base + D_c);
        if (with_masks) issue_gemm(sbase, Smem::A_h2, Smem::W + wl.m2, 64, wl.Kp, tmem_base + D_d);
        umma::commit(bar);
    }
    __syncwarp();
    umma::mbar_wait(bar, phase); phase ^= 1u;
    umma::fence_after_sync();
    return sigma;
}

// CTA prologue shared by both kernels: weights -> smem, level geometry, mbarrier, TMEM.
__device__ __forceinline__ uint32_t cta_setup(uint8_t* smem, const inerf_field_desc& desc, LevelGeom*& lg, uint64_t*& bar) {
    const WeightLayout wl = weight_layout(desc.K);
    const uint32_t misc = Smem::misc(desc.K);
    lg = reinterpret_cast<LevelGeom*>(smem + misc);
    bar = reinterpret_cast<uint64_t*>(smem + misc + 16 * sizeof(LevelGeom));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + misc + 16 * sizeof(LevelGeom) + 8);
    const uint4* wsrc = reinterpret_cast<const uint4*>(desc.weights);
    uint4* wdst = reinterpret_cast<uint4*>(smem + Smem::W);
    for (uint32_t i = threadIdx.x; i < wl.total / 16; i += blockDim.x) wdst[i] = __ldg(wsrc + i);
    init_levels(lg, desc.offsets, desc.L, desc.S, desc.H);
    if (threadIdx.x == 0) { umma::mbar_init(bar, 1); umma::mbar_fence_init(); }
    if (threadIdx.x < 32) umma::tmem_alloc<kTmemCols>(tmem_slot);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    return *tmem_slot;
}
__device__ __forceinline__ void cta_teardown(uint32_t tmem_base) {
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace field
