// inerf_field_forward: NeRFNetwork.forward (nerf/network_mask.py:119-158) as ONE persistent kernel.
//
// Per 128-sample tile: both hash encoders gathered straight into UMMA operand tiles in shared memory
// (no [B,32] activation ever touches HBM), SH degree 4, then sigma-net, colour-net and mask-net as a chain
// of tcgen05.mma (fp16 x fp16 -> fp32 in TMEM) with TMEM->register epilogues that apply ReLU / exp / sigmoid
// and write the next layer's operand tile.  Weights (< 40 KB fp16) stay resident in shared memory for the
// life of the CTA.  (The first version -- two lock-step 256-thread CTAs per SM that overlapped gathers and MMAs only across
// CTAs -- took 3.78 ms for 7.8 M samples against 2.05 ms for the warp-specialised kernel below and was removed)
#include <vector>

#include "field_device.cuh"
#include "occupancy_device.cuh"

namespace {

using namespace field;

// ---- gather and MLP chain as concurrent roles of ONE CTA per SM -----------------------------------------------------------
// The structure is the render kernel's (render_fused.cu) without its marcher and compositor: 16 gather warps (thread = sample
// row x 4 of the 16 levels, 64 registers) fill one of two operand stages while 8 chain warps (96 registers) run the tcgen05
// chain of the previous tile and write sigma / rgb / logits (and the saved mask-net input row); the gather warps never wait
// for an epilogue.
constexpr uint32_t kWsChainT = 256, kWsGatherT = 512, kWsThreads = kWsChainT + kWsGatherT, kWsStages = 2;

struct WsCtrl {
    uint64_t a_full[kWsStages], a_empty[kWsStages], mma_bar;
    uint32_t tmem_slot;
    LevelGeom lg[16];
};
struct WsSmem {
    static constexpr uint32_t A = 0;
    static constexpr uint32_t W = A + kWsStages * kStageBytes;
    static __host__ __device__ uint32_t ctrl(uint32_t K) { return (W + weight_layout(K).total + 15u) & ~15u; }
    static __host__ __device__ uint32_t bytes(uint32_t K) { return ctrl(K) + (uint32_t)sizeof(WsCtrl) + 16u; }
};

__global__ void __launch_bounds__(kWsThreads, 1) k_field_forward_ws(inerf_field_desc desc, const float* __restrict__ xyzs,
                                                                    const float* __restrict__ dirs, uint32_t B_rows,
                                                                    float* __restrict__ sigmas, float* __restrict__ rgbs,
                                                                    float* __restrict__ masks, uint4* __restrict__ x0_save,
                                                                    uint4* __restrict__ xs_save) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t K = desc.K, Kp = weight_layout(K).Kp;
    WsCtrl* ctl = reinterpret_cast<WsCtrl*>(smem + WsSmem::ctrl(K));
    const uint32_t tid = threadIdx.x;
    load_weights(smem, WsSmem::W, desc.weights, K);
    init_levels(ctl->lg, desc.offsets, desc.L, desc.S, desc.H, tid);
    if (tid == 0) {
        for (uint32_t i = 0; i < kWsStages; i++) { umma::mbar_init(&ctl->a_full[i], kWsGatherT / 32); umma::mbar_init(&ctl->a_empty[i], 1); }
        umma::mbar_init(&ctl->mma_bar, 1);
        umma::mbar_fence_init();
    }
    if (tid < 32) umma::tmem_alloc<kTmemCols>(&ctl->tmem_slot);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = ctl->tmem_slot;
    const uint32_t B = desc.n_valid ? min(B_rows, (uint32_t)max(0, __ldg(desc.n_valid))) : B_rows;   // rows past *n_valid: padding
    const uint32_t num_tiles = (B + kTile - 1) / kTile;
    const bool with_masks = masks != nullptr;

    if (tid < kWsChainT) {
        // ------------------------------------------------------------------------------------------------ chain --
        umma::reg_alloc<96>();
        const uint32_t warp = tid >> 5, lane = tid & 31;
        uint32_t phase = 0, it = 0;
        auto chain_sync = [] { umma::named_sync<1, kWsChainT>(); };
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
            const uint32_t sa = it % kWsStages;
            umma::mbar_wait(&ctl->a_full[sa], (it / kWsStages) & 1u);
            const uint32_t a_es = WsSmem::A + sa * kStageBytes;
            const ChainBufs bufs{a_es, a_es + kBytesEs, a_es + kBytesEs + kBytesCi, WsSmem::W};
            const uint32_t orow = tile * kTile + (warp & 3u) * 32u + lane;
            const float sigma = mlp_chain(smem, bufs, tmem_base, &ctl->mma_bar, phase, K, desc.density_scale, with_masks, tid, &ctl->a_empty[sa],
                                          chain_sync, [&](float) {
                // training: keep the mask-net input row (32 mask-table features | 15 geo | 0, fp16) for the backward pass.  The row's
                // owner wrote the geo columns itself a moment ago, and the stage is only released after the layer-0 MMAs.
                if (x0_save != nullptr && warp < 4 && orow < B) {
#pragma unroll
                    for (uint32_t c = 0; c < 6; c++)
                        x0_save[(size_t)orow * 6 + c] =
                            *reinterpret_cast<const uint4*>(smem + bufs.a_mi + umma::tile_off(tid, c * 8, kLBO, sbo_of(48)));
                }
                // stage-1 training: keep the inputs of BOTH nets -- 32 sigma-table features | SH16, geo15, 0 (fp16 [.,64])
                if (xs_save != nullptr && warp < 4 && orow < B) {
#pragma unroll
                    for (uint32_t c = 0; c < 4; c++) {
                        xs_save[(size_t)orow * 8 + c] = *reinterpret_cast<const uint4*>(smem + bufs.a_es + umma::tile_off(tid, c * 8, kLBO, sbo_of(32)));
                        xs_save[(size_t)orow * 8 + 4 + c] = *reinterpret_cast<const uint4*>(smem + bufs.a_ci + umma::tile_off(tid, c * 8, kLBO, sbo_of(32)));
                    }
                }
            });
            if (warp < 4) {
                float rgb[3];
                epilogue_rgb(tmem_base, rgb, tid);
                if (orow < B) {
                    sigmas[orow] = sigma;
                    rgbs[(size_t)orow * 3] = rgb[0]; rgbs[(size_t)orow * 3 + 1] = rgb[1]; rgbs[(size_t)orow * 3 + 2] = rgb[2];
                }
            }
            if (with_masks) {
                const uint32_t chunks = Kp / 16, c_begin = (warp >> 2) ? (chunks + 1) / 2 : 0, c_end = (warp >> 2) ? chunks : (chunks + 1) / 2;
                for (uint32_t c = c_begin; c < c_end; c++) {
                    uint32_t v[16];
                    umma::tmem_ld16(tmem_base + D_d + (((warp & 3u) * 32u) << 16) + c * 16, v);
                    umma::tmem_ld_wait();
                    if (orow < B) {
                        float* out = masks + (size_t)orow * K + c * 16;
                        if ((K & 3u) == 0 && c * 16 + 16 <= K) {
#pragma unroll
                            for (int i = 0; i < 4; i++)
                                reinterpret_cast<float4*>(out)[i] = make_float4(__half2float(__float2half_rn(__uint_as_float(v[4 * i]))),
                                                                                __half2float(__float2half_rn(__uint_as_float(v[4 * i + 1]))),
                                                                                __half2float(__float2half_rn(__uint_as_float(v[4 * i + 2]))),
                                                                                __half2float(__float2half_rn(__uint_as_float(v[4 * i + 3]))));
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; i++)
                                if (c * 16 + i < K) out[i] = __half2float(__float2half_rn(__uint_as_float(v[i])));
                        }
                    }
                }
            }
            umma::fence_before_sync();
            chain_sync();   // TMEM and the chain's hidden tiles are reused by the next tile
        }
    } else {
        // ----------------------------------------------------------------------------------------------- gather --
        umma::reg_dealloc<64>();
        const uint32_t gt = tid - kWsChainT, row = gt & (kTile - 1), quarter = gt >> 7;
        const float inv2b = __fdiv_rn(1.0f, __fmul_rn(2.0f, desc.bound));
        const uint2* table = reinterpret_cast<const uint2*>(desc.table_packed);
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
            const uint32_t sa = it % kWsStages;
            if (it >= kWsStages) umma::mbar_wait(&ctl->a_empty[sa], ((it / kWsStages) - 1u) & 1u);
            const uint32_t s = tile * kTile + row;
            if (s < B) {   // rows past the end keep stale operands: their outputs are never stored
                const uint32_t a_es = WsSmem::A + sa * kStageBytes, a_ci = a_es + kBytesEs, a_mi = a_ci + kBytesCi;
                float x01[3];
                bool oob = false;
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    x01[d] = __fmul_rn(__fadd_rn(__ldg(xyzs + (size_t)s * 3 + d), desc.bound), inv2b);  // grid.py:149
                    oob |= (x01[d] < 0.f || x01[d] > 1.f);
                }
                if (quarter == 0) sh16_to_smem(__ldg(dirs + (size_t)s * 3), __ldg(dirs + (size_t)s * 3 + 1), __ldg(dirs + (size_t)s * 3 + 2), smem, a_ci, row);
                encode4(x01, oob, quarter * 4, ctl->lg, table, smem, a_es, a_mi, row);
            }
            umma::fence_async_smem();
            __syncwarp();
            if ((gt & 31u) == 0) umma::mbar_arrive(&ctl->a_full[sa]);   // one arrival per gather warp
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc<kTmemCols>(tmem_base);
}

// ---- occupancy-grid density sweep (update_extra_state, nerf/mask_renderer.py:466-527) ------------------------------------
// tmp_grid[c, cell] = density(jittered point of the cell) * density_scale for every sample of the sweep, in ONE launch with
// the role structure above: the gather warps make each point from its cell id (Morton decode, cascade scale, jitter --
// occupancy_device.cuh) instead of reading a sample stream, the chain warps run the sigma-net only and scatter the result to
// the cell.  The reference goes through meshgrid / morton3D / rand_like / two table casts / encode / 2 GEMMs / index_put per
// (block, cascade) from 5 nested Python loops.
__global__ void __launch_bounds__(kWsThreads, 1) k_occupancy_density(inerf_field_desc desc, occ::OccPoints pts, uint32_t n, float* __restrict__ tmp_grid) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t K = desc.K;
    WsCtrl* ctl = reinterpret_cast<WsCtrl*>(smem + WsSmem::ctrl(K));
    const uint32_t tid = threadIdx.x;
    load_weights(smem, WsSmem::W, desc.weights, K);
    init_levels(ctl->lg, desc.offsets, desc.L, desc.S, desc.H, tid);
    if (tid == 0) {
        for (uint32_t i = 0; i < kWsStages; i++) { umma::mbar_init(&ctl->a_full[i], kWsGatherT / 32); umma::mbar_init(&ctl->a_empty[i], 1); }
        umma::mbar_init(&ctl->mma_bar, 1);
        umma::mbar_fence_init();
    }
    if (tid < 32) umma::tmem_alloc<kTmemCols>(&ctl->tmem_slot);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = ctl->tmem_slot;
    const uint32_t num_tiles = (n + kTile - 1) / kTile;

    if (tid < kWsChainT) {
        umma::reg_alloc<96>();
        uint32_t phase = 0, it = 0;
        auto chain_sync = [] { umma::named_sync<1, kWsChainT>(); };
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
            const uint32_t sa = it % kWsStages;
            umma::mbar_wait(&ctl->a_full[sa], (it / kWsStages) & 1u);
            const uint32_t a_es = WsSmem::A + sa * kStageBytes;
            const ChainBufs bufs{a_es, a_es + kBytesEs, a_es + kBytesEs + kBytesCi, WsSmem::W};
            const float sigma = sigma_chain(smem, bufs, tmem_base, &ctl->mma_bar, phase, K, tid, &ctl->a_empty[sa], chain_sync);
            const uint32_t s = tile * kTile + tid;
            if (tid < kTile && s < n) {
                const uint32_t c = s / pts.per_cascade;
                const uint32_t m = pts.cells ? (uint32_t)__ldg(pts.cells + s) : s - c * pts.per_cascade;
                tmp_grid[(size_t)c * pts.G * pts.G * pts.G + m] = __fmul_rn(sigma, desc.density_scale);   // `sigmas *= self.density_scale`
            }
            umma::fence_before_sync();
            chain_sync();   // TMEM and the hidden tile are reused by the next tile
        }
    } else {
        umma::reg_dealloc<64>();
        const uint32_t gt = tid - kWsChainT, row = gt & (kTile - 1), quarter = gt >> 7;
        const float inv2b = __fdiv_rn(1.0f, __fmul_rn(2.0f, desc.bound));
        const uint2* table = reinterpret_cast<const uint2*>(desc.table_packed);
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
            const uint32_t sa = it % kWsStages;
            if (it >= kWsStages) umma::mbar_wait(&ctl->a_empty[sa], ((it / kWsStages) - 1u) & 1u);
            const uint32_t s = tile * kTile + row;
            if (s < n) {
                const uint32_t a_es = WsSmem::A + sa * kStageBytes, a_mi = a_es + kBytesEs + kBytesCi;
                float xyz[3], x01[3];
                uint32_t flat;
                occ::occ_point(pts, s, xyz, flat);
                bool oob = false;
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    x01[d] = __fmul_rn(__fadd_rn(xyz[d], desc.bound), inv2b);  // grid.py:149
                    oob |= (x01[d] < 0.f || x01[d] > 1.f);
                }
                encode4(x01, oob, quarter * 4, ctl->lg, table, smem, a_es, a_mi, row);
            }
            umma::fence_async_smem();
            __syncwarp();
            if ((gt & 31u) == 0) umma::mbar_arrive(&ctl->a_full[sa]);   // one arrival per gather warp
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc<kTmemCols>(tmem_base);
}

// fp32 / fp16 embeddings of both encoders -> interleaved fp16 (sigma.c0, sigma.c1, mask.c0, mask.c1), one pass
template <typename T>
__global__ void k_pack_tables(const T* __restrict__ es, const T* __restrict__ em, uint64_t n, uint2* __restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        __half2 a, b;
        if constexpr (std::is_same<T, float>::value) {
            const float2 fa = reinterpret_cast<const float2*>(es)[i], fb = reinterpret_cast<const float2*>(em)[i];
            a = __floats2half2_rn(fa.x, fa.y); b = __floats2half2_rn(fb.x, fb.y);
        } else {
            a = reinterpret_cast<const __half2*>(es)[i]; b = reinterpret_cast<const __half2*>(em)[i];
        }
        out[i] = make_uint2(h2_bits(a), h2_bits(b));
    }
}

inline uint16_t f2h_bits(float f) {
    const __half h = __float2half_rn(f);
    uint16_t b;
    memcpy(&b, &h, 2);
    return b;
}

// W is [n_out, n_in] row-major fp32 -> B-operand tile [Npad x Kpad] with zero padding
void pack_layer(uint8_t* dst, const float* W, uint32_t n_out, uint32_t n_in, uint32_t Npad, uint32_t Kpad) {
    const uint32_t sbo = sbo_of(Kpad);
    for (uint32_t n = 0; n < Npad; n++)
        for (uint32_t k = 0; k < Kpad; k++) {
            const float v = (n < n_out && k < n_in) ? W[(size_t)n * n_in + k] : 0.f;
            const uint16_t b = f2h_bits(v);
            memcpy(dst + umma::tile_off(n, k, kLBO, sbo), &b, 2);
        }
}

int validate_desc(const inerf_field_desc* d) {
    INERF_REQUIRE(d);
    INERF_REQUIRE(d->table_packed); INERF_REQUIRE(d->offsets); INERF_REQUIRE(d->weights);
    if (d->L != 16 || d->K == 0 || d->K > 64 || d->H == 0 || !(d->bound > 0.f)) return INERF_ERR_UNSUPPORTED;
    if (((uintptr_t)d->weights & 15u) || ((uintptr_t)d->table_packed & 7u)) return INERF_ERR_ALIGN;
    return INERF_OK;
}

}  // namespace

namespace field {
int validate(const inerf_field_desc* d) { return validate_desc(d); }
}

extern "C" int inerf_field_pack_tables(const void* emb_sigma, const void* emb_mask, int dtype, uint64_t n_entries, void* packed,
                                       void* stream) {
    INERF_REQUIRE(emb_sigma); INERF_REQUIRE(emb_mask); INERF_REQUIRE(packed);
    if (n_entries == 0) return INERF_OK;
    if (((uintptr_t)packed & 7u) || ((uintptr_t)emb_sigma & 7u) || ((uintptr_t)emb_mask & 7u)) return INERF_ERR_ALIGN;
    const unsigned grid = 8 * kNumSMs;
    if (dtype == INERF_F32) k_pack_tables<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)emb_sigma, (const float*)emb_mask, n_entries, (uint2*)packed);
    else if (dtype == INERF_F16) k_pack_tables<__half><<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)emb_sigma, (const __half*)emb_mask, n_entries, (uint2*)packed);
    else return INERF_ERR_UNSUPPORTED;
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" size_t inerf_field_weights_bytes(uint32_t K) { return (K == 0 || K > 64) ? 0 : field::weight_layout(K).total; }

extern "C" int inerf_field_pack_weights(const float* sigma0, const float* sigma1, const float* color0, const float* color1,
                                        const float* color2, const float* mask0, const float* mask1, const float* mask2, uint32_t K,
                                        void* packed_host) {
    if (K == 0 || K > 64) return INERF_ERR_SIZE;
    INERF_REQUIRE(sigma0); INERF_REQUIRE(sigma1); INERF_REQUIRE(color0); INERF_REQUIRE(color1); INERF_REQUIRE(color2);
    INERF_REQUIRE(mask0); INERF_REQUIRE(mask1); INERF_REQUIRE(mask2); INERF_REQUIRE(packed_host);
    const field::WeightLayout wl = field::weight_layout(K);
    uint8_t* p = static_cast<uint8_t*>(packed_host);
    memset(p, 0, wl.total);
    pack_layer(p + wl.s0, sigma0, 64, 32, 64, 32);
    pack_layer(p + wl.s1, sigma1, 16, 64, 16, 64);
    pack_layer(p + wl.c0, color0, 64, 31, 64, 32);
    pack_layer(p + wl.c1, color1, 64, 64, 64, 64);
    pack_layer(p + wl.c2, color2, 3, 64, 16, 64);
    pack_layer(p + wl.m0, mask0, 64, 47, 64, 48);
    pack_layer(p + wl.m1, mask1, 64, 64, 64, 64);
    pack_layer(p + wl.m2, mask2, K, 64, wl.Kp, 64);
    return INERF_OK;
}

static int field_forward_impl(const inerf_field_desc* desc, const float* xyzs, const float* dirs, uint32_t B, float* sigmas, float* rgbs,
                              float* masks, void* x0_save, void* stream, void* xs_save = nullptr) {
    if (int e = validate_desc(desc)) return e;
    if (B == 0) return INERF_OK;
    INERF_REQUIRE(xyzs); INERF_REQUIRE(dirs); INERF_REQUIRE(sigmas); INERF_REQUIRE(rgbs);
    if (x0_save && (((uintptr_t)x0_save & 15u) || masks == nullptr)) return INERF_ERR_ALIGN;
    if (xs_save && ((uintptr_t)xs_save & 15u)) return INERF_ERR_ALIGN;
    const uint32_t num_tiles = (B + field::kTile - 1) / field::kTile;
    // per-device attribute, cheap to set: no process-global "already done" flag
    cudaError_t e = cudaFuncSetAttribute(k_field_forward_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return (int)e;
    const uint32_t sms = (uint32_t)device_sm_count();
    const uint32_t grid_ws = num_tiles < sms ? num_tiles : sms;
    k_field_forward_ws<<<grid_ws, kWsThreads, WsSmem::bytes(desc->K), (cudaStream_t)stream>>>(*desc, xyzs, dirs, B, sigmas, rgbs, masks,
                                                                                             (uint4*)x0_save, (uint4*)xs_save);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_occupancy_density(const inerf_field_desc* desc, uint32_t C, uint32_t G, const int32_t* cells, uint32_t per_cascade,
                                       const float* noise, uint64_t seed, float* tmp_grid, void* stream) {
    if (int e = validate_desc(desc)) return e;
    occ::OccPoints pts;
    if (int e = occ::make_points(&pts, C, G, desc->bound, cells, per_cascade, noise, seed)) return e;
    INERF_REQUIRE(tmp_grid);
    const uint32_t n = C * per_cascade, num_tiles = (n + field::kTile - 1) / field::kTile;
    cudaError_t ce = cudaFuncSetAttribute(k_occupancy_density, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (ce != cudaSuccess) return (int)ce;
    const uint32_t sms = (uint32_t)device_sm_count();
    k_occupancy_density<<<num_tiles < sms ? num_tiles : sms, kWsThreads, WsSmem::bytes(desc->K), (cudaStream_t)stream>>>(*desc, pts, n, tmp_grid);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_field_forward(const inerf_field_desc* desc, const float* xyzs, const float* dirs, uint32_t B, float* sigmas,
                                   float* rgbs, float* masks, void* stream) {
    return field_forward_impl(desc, xyzs, dirs, B, sigmas, rgbs, masks, nullptr, stream);
}

extern "C" int inerf_field_forward_train(const inerf_field_desc* desc, const float* xyzs, const float* dirs, uint32_t B, float* sigmas,
                                         float* rgbs, float* masks, void* x0_save, void* stream) {
    INERF_REQUIRE(x0_save);
    return field_forward_impl(desc, xyzs, dirs, B, sigmas, rgbs, masks, x0_save, stream);
}

// Stage-1 (RGB-sigma) training forward: sigma / rgb as inerf_field_forward with the instance head off, plus the saved inputs of
// the sigma and colour nets for inerf_field_backward_rgb.
extern "C" int inerf_field_forward_train_rgb(const inerf_field_desc* desc, const float* xyzs, const float* dirs, uint32_t B, float* sigmas,
                                             float* rgbs, void* xs_save, void* stream) {
    INERF_REQUIRE(xs_save);
    return field_forward_impl(desc, xyzs, dirs, B, sigmas, rgbs, nullptr, nullptr, stream, xs_save);
}
