// inerf_render_fused: whole-frame instance-field inference in ONE persistent launch.
//
// Replaces the host loop of NeRFMaskRenderer.run_cuda (nerf/mask_renderer.py:322-381): per iteration the
// reference launches march_rays, two full-table fp16 casts, two grid encodes, 8 GEMMs + elementwise kernels,
// composite_rays_with_masks and a boolean-index compaction, with two device->host syncs.  Here each CTA owns 128
// ray slots (thread = slot = TMEM lane); every pass each live slot marches to its next occupied sample
// (raymarching.cu:1008-1062 semantics), the 128 samples go through the hash gathers + tcgen05 MLP chain of
// field_device.cuh, and each thread composites its own sample into register accumulators
// (raymarching.cu:1216-1255: T = 1 - sum(w), stop once T < T_thresh).  Finished slots pull the next ray id from a
// global counter.  No sample stream, no per-sample sigma/rgb/logit tensor and no alive list ever touch HBM.
//
// Per-ray sample positions equal the reference's for a continuous march from `near`; the reference re-derives t
// from the composited depth deltas between its n_step-sized chunks, which can differ in the last ulp after long
// empty-space skips (documented in DESIGN.md; maps agree to the 1e-3 tolerance of the north star).
#include "field_device.cuh"
#include "march_device.cuh"

namespace {

using namespace field;

struct RenderParams {
    const float* rays_o;
    const float* rays_d;
    const float* nears;
    const float* fars;
    const uint8_t* bitfield;
    uint32_t N, C, H, max_steps;
    float dt_gamma, T_thresh;
    float* weights_sum;
    float* depth;
    float* image;
    float* mask_out;
    int32_t* work_counter;
};

constexpr uint32_t kScratchBytes = kTile * 3 * 4 + kTile * 4 + kTile * 4;

template <int NCH>  // 16-column logit chunks owned per thread: 1 -> K <= 32, 2 -> K <= 64
__global__ void __launch_bounds__(kThreads, 2) k_render_fused(inerf_field_desc desc, RenderParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    LevelGeom* lg;
    uint64_t* bar;
    const uint32_t tmem_base = cta_setup(smem, desc, lg, bar);
    uint32_t phase = 0;
    const uint32_t K = desc.K, Kp = weight_layout(K).Kp;
    const bool with_masks = p.mask_out != nullptr;
    const float inv2b = __fdiv_rn(1.0f, __fmul_rn(2.0f, desc.bound));
    const __half2* tab_s = reinterpret_cast<const __half2*>(desc.table_sigma);
    const __half2* tab_m = reinterpret_cast<const __half2*>(desc.table_mask);

    const uint32_t scratch = (Smem::bytes(K) + 15u) & ~15u;
    float* xyz_s = reinterpret_cast<float*>(smem + scratch);
    float* w_s = xyz_s + kTile * 3;
    int32_t* flush_s = reinterpret_cast<int32_t*>(w_s + kTile);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row = threadIdx.x & (kTile - 1), half = threadIdx.x >> 7;
    const bool owner = half == 0;
    // logit chunks of this thread: warps 0..3 take [0, ceil(chunks/2)), warps 4..7 the rest
    const uint32_t chunks = Kp / 16;
    const uint32_t c_begin = half ? (chunks + 1) / 2 : 0, c_end = half ? chunks : (chunks + 1) / 2;

    // ---- per-slot ray state (owner threads) ----
    march::Walk wk;
    int32_t ray = -1;
    bool done = false;
    float t = 0.f, last_t = 0.f, t_depth = 0.f, ws = 0.f, dep = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    uint32_t nsteps = 0;
    float macc[NCH][16];
#pragma unroll
    for (int c = 0; c < NCH; c++)
#pragma unroll
        for (int i = 0; i < 16; i++) macc[c][i] = 0.f;

    auto store_macc = [&](int32_t r) {
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            if (c_begin + c < c_end) {
                float* out = p.mask_out + (size_t)r * K + (c_begin + c) * 16;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    if ((c_begin + c) * 16 + i < K) out[i] = macc[c][i];
                    macc[c][i] = 0.f;
                }
            }
        }
    };

    while (true) {
        bool have = false;
        float x = 0.f, y = 0.f, z = 0.f, dt = 0.f, delta1 = 0.f;
        if (owner) {
            int32_t flush_ray = -1;
            while (true) {
                if (ray < 0) {
                    const uint32_t idx = (uint32_t)atomicAdd(p.work_counter, 1);
                    if (idx >= p.N) break;
                    ray = (int32_t)idx;
                    wk.init(p.rays_o + (size_t)idx * 3, p.rays_d + (size_t)idx * 3, p.bitfield, desc.bound, p.dt_gamma, p.max_steps,
                            p.C, p.H, __ldg(p.fars + idx));
                    t = __ldg(p.nears + idx);
                    last_t = t; t_depth = t;
                    ws = dep = cr = cg = cb = 0.f;
                    nsteps = 0; done = false;
                }
                if (!done && nsteps < p.max_steps && wk.next_sample(t, x, y, z, dt)) {
                    delta1 = __fsub_rn(t, last_t);
                    last_t = t;
                    nsteps++;
                    have = true;
                    break;
                }
                // ray finished: write its pixel
                p.weights_sum[ray] = ws;
                p.depth[ray] = dep;
                p.image[(size_t)ray * 3] = cr; p.image[(size_t)ray * 3 + 1] = cg; p.image[(size_t)ray * 3 + 2] = cb;
                if (with_masks) {
                    if (nsteps > 0) {
                        store_macc(ray);
                        flush_ray = ray;      // the partner thread holds the other half of the classes
                    } else {
                        float* out = p.mask_out + (size_t)ray * K;
                        for (uint32_t k = 0; k < K; k++) out[k] = 0.f;
                    }
                }
                ray = -1;
            }
            xyz_s[row * 3] = x; xyz_s[row * 3 + 1] = y; xyz_s[row * 3 + 2] = z;
            flush_s[row] = flush_ray;
        }
        const int any = __syncthreads_or(have ? 1 : 0);
        if (!owner && with_masks) {
            const int32_t fr = flush_s[row];
            if (fr >= 0) store_macc(fr);
        }
        if (!any) break;

        // ---- gathers for the 128 samples of this pass ----
        float x01[3];
        bool oob = false;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            x01[d] = __fmul_rn(__fadd_rn(xyz_s[row * 3 + d], desc.bound), inv2b);
            oob |= (x01[d] < 0.f || x01[d] > 1.f);
        }
        encode8(x01, oob, half * 8, lg, tab_s, tab_m, smem, row);
        if (owner) sh16_to_smem(have ? wk.dx : 0.f, have ? wk.dy : 0.f, have ? wk.dz : 0.f, smem, row);
        umma::fence_async_smem();
        umma::fence_before_sync();
        __syncthreads();

        float weight = 0.f;
        mlp_chain(smem, tmem_base, bar, phase, K, desc.density_scale, with_masks, [&](float sigma) {
            if (owner) {
                if (have) {
                    const float alpha = 1.0f - __expf(-sigma * dt);
                    const float T = 1.0f - ws;
                    weight = alpha * T;
                    ws += weight;
                    t_depth += delta1;
                    dep = fmaf(weight, t_depth, dep);
                    if (T < p.T_thresh) done = true;
                }
                w_s[row] = weight;
            }
        });

        if (owner) {
            float rgb[3];
            epilogue_rgb(tmem_base, rgb);
            cr = fmaf(weight, rgb[0], cr); cg = fmaf(weight, rgb[1], cg); cb = fmaf(weight, rgb[2], cb);
        }
        if (with_masks) {
            const float wgt = w_s[row];
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                if (c_begin + c < c_end) {
                    uint32_t v[16];
                    umma::tmem_ld16(tmem_base + D_d + (((warp & 3u) * 32u) << 16) + (c_begin + c) * 16, v);
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) macc[c][i] = fmaf(wgt, __half2float(__float2half_rn(__uint_as_float(v[i]))), macc[c][i]);
                }
            }
        }
        umma::fence_before_sync();
        __syncthreads();  // TMEM, operand tiles and the scratch arrays are reused by the next pass
    }
    cta_teardown(tmem_base);
    (void)lane;
}

}  // namespace

namespace field { int validate(const inerf_field_desc* d); }

extern "C" int inerf_render_fused(const inerf_field_desc* desc, const float* rays_o, const float* rays_d, const float* nears,
                                  const float* fars, const uint8_t* bitfield, uint32_t N, uint32_t C, uint32_t H, float dt_gamma,
                                  uint32_t max_steps, float T_thresh, float* weights_sum, float* depth, float* image,
                                  float* mask_out, int32_t* work_counter, void* stream) {
    if (int e = field::validate(desc)) return e;
    if (C == 0 || C > 16 || H == 0 || H > 1024 || (H & (H - 1)) || max_steps == 0) return INERF_ERR_SIZE;
    if (N == 0) return INERF_OK;
    INERF_REQUIRE(rays_o); INERF_REQUIRE(rays_d); INERF_REQUIRE(nears); INERF_REQUIRE(fars); INERF_REQUIRE(bitfield);
    INERF_REQUIRE(weights_sum); INERF_REQUIRE(depth); INERF_REQUIRE(image); INERF_REQUIRE(work_counter);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce = cudaMemsetAsync(work_counter, 0, sizeof(int32_t), st);
    if (ce != cudaSuccess) return (int)ce;
    RenderParams p{rays_o, rays_d, nears, fars, bitfield, N, C, H, max_steps, dt_gamma, T_thresh, weights_sum, depth, image, mask_out, work_counter};
    const uint32_t smem_bytes = ((field::Smem::bytes(desc->K) + 15u) & ~15u) + kScratchBytes;
    static bool attr_set = false;
    if (!attr_set) {
        ce = cudaFuncSetAttribute(k_render_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (ce != cudaSuccess) return (int)ce;
        ce = cudaFuncSetAttribute(k_render_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (ce != cudaSuccess) return (int)ce;
        attr_set = true;
    }
    const uint32_t want = (N + field::kTile - 1) / field::kTile;
    const uint32_t grid = want < 2u * kNumSMs ? want : 2u * kNumSMs;
    if (desc->K <= 32) k_render_fused<1><<<grid, field::kThreads, smem_bytes, st>>>(*desc, p);
    else k_render_fused<2><<<grid, field::kThreads, smem_bytes, st>>>(*desc, p);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}
