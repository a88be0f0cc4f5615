// inerf_render_fused: whole-frame instance-field inference in ONE persistent, warp-specialised launch.
//
// Replaces the host loop of NeRFMaskRenderer.run_cuda (nerf/mask_renderer.py:322-381): per iteration the
// reference launches march_rays, two full-table fp16 casts, two grid encodes, 8 GEMMs + elementwise kernels,
// composite_rays_with_masks and a boolean-index compaction, with two device->host syncs.
//
// One CTA per SM, 640 threads in three roles that run concurrently and hand 128-sample tiles to each other
// through mbarrier-guarded rings in shared memory (nothing per-sample ever touches HBM):
//
//   march  (4 warps, thread = ray slot)   occupancy-grid DDA (raymarching.cu:1008-1062 semantics), one sample per
//                                          slot per tile into a 4-deep FIFO; runs AHEAD of compositing
//                                          (speculatively: a ray killed by T < T_thresh drops its queued samples);
//                                          finished slots pull the next ray id from a global counter.
//   gather (8 warps, thread = slot x half  16 levels x 8 corners, one 8-byte gather per corner from the interleaved
//           of the levels)                 fp16 table (both encoders at once), fp16 trilinear blend, SH degree 4 ->
//                                          UMMA operand tiles, double buffered.
//   chain  (8 warps, thread = TMEM lane x  sigma / colour / mask MLPs as tcgen05.mma chains (fp32 accumulators in
//           column half)                   TMEM), ReLU / exp / sigmoid epilogues, then the per-ray alpha compositing
//                                          of colour, depth and K instance logits in registers
//                                          (raymarching.cu:1216-1255: T = 1 - sum(w), stop once T < T_thresh).
//
// Per-ray sample positions equal the reference's for a continuous march from `near`; the reference re-derives t
// from the composited depth deltas between its n_step-sized chunks, which can differ in the last ulp after long
// empty-space skips (maps agree to the 1e-3 tolerance of the north star).
#include "field_device.cuh"
#include "march_device.cuh"

namespace {

using namespace field;

constexpr uint32_t kChainT = 256, kGatherT = 256, kMarchT = 128;
constexpr uint32_t kThreadsR = kChainT + kGatherT + kMarchT;
constexpr uint32_t DQ = 4;   // sample FIFO depth (tiles)
constexpr uint32_t DA = 2;   // gathered operand stages

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

struct QStage {   // one FIFO stage: one sample per ray slot
    float x[kTile], y[kTile], z[kTile], dt[kTile], d1[kTile], dx[kTile], dy[kTile], dz[kTile];
    int32_t ray[kTile];   // -1 = no sample in this slot
    int32_t end[kTile];   // 1 = last sample of its ray
};

struct Ctrl {
    uint64_t q_full[DQ], q_empty[DQ], a_full[DA], a_empty[DA], mma_bar;
    uint32_t tmem_slot;
    int32_t n_done;       // marchers that ran out of rays
    int32_t last_tile;    // last tile holding a real sample (valid once n_done == 128)
    int32_t a_flag[DA];   // 1 = terminal tile
    int32_t kill[kTile];  // ray id terminated early by the compositor (marcher drops it)
    float w_s[kTile];
    int32_t fin_s[kTile];
    LevelGeom lg[16];
};

struct RSmem {
    static constexpr uint32_t A = 0;                                 // DA stages of (es | ci | mi)
    static constexpr uint32_t H1 = A + DA * kStageBytes;
    static constexpr uint32_t H2 = H1 + kBytesH;
    static constexpr uint32_t W = H2 + kBytesH;
    static __host__ __device__ uint32_t ctrl(uint32_t K) { return (W + weight_layout(K).total + 15u) & ~15u; }
    static __host__ __device__ uint32_t queue(uint32_t K) { return (ctrl(K) + (uint32_t)sizeof(Ctrl) + 15u) & ~15u; }
    static __host__ __device__ uint32_t bytes(uint32_t K) { return queue(K) + DQ * (uint32_t)sizeof(QStage); }
};

__device__ __forceinline__ int32_t ld_vol(const int32_t* p) { return *reinterpret_cast<const volatile int32_t*>(p); }
__device__ __forceinline__ void st_vol(int32_t* p, int32_t v) { *reinterpret_cast<volatile int32_t*>(p) = v; }

// ------------------------------------------------------------------------------------------------ march --
__device__ __forceinline__ void march_role(const inerf_field_desc& desc, const RenderParams& p, Ctrl* ctl, QStage* q, uint32_t r) {
    march::Walk wk;
    int32_t ray = -1, my_last = -1;
    bool exhausted = false, have_pending = false;
    uint32_t nsteps = 0;
    float t = 0.f, last_t = 0.f, px = 0.f, py = 0.f, pz = 0.f, pdt = 0.f, pd1 = 0.f;
    for (uint32_t tile = 0;; tile++) {
        const uint32_t s = tile % DQ;
        bool quit = false;
        if (tile >= DQ) {
            const uint32_t par = ((tile / DQ) - 1u) & 1u;
            while (!umma::mbar_try_wait(&ctl->q_empty[s], par)) {
                if (exhausted && ld_vol(&ctl->n_done) == (int32_t)kMarchT) {
                    __threadfence_block();
                    if ((int32_t)tile > ld_vol(&ctl->last_tile) + 1) { quit = true; break; }
                }
            }
        }
        if (!quit && exhausted && ld_vol(&ctl->n_done) == (int32_t)kMarchT) {
            __threadfence_block();
            if ((int32_t)tile > ld_vol(&ctl->last_tile) + 1) quit = true;
        }
        if (quit) break;

        if (ray >= 0 && ld_vol(&ctl->kill[r]) == ray) { ray = -1; have_pending = false; }
        if (!have_pending && !exhausted) {
            while (true) {
                const uint32_t idx = (uint32_t)atomicAdd(p.work_counter, 1);
                if (idx >= p.N) { exhausted = true; break; }
                wk.init(p.rays_o + (size_t)idx * 3, p.rays_d + (size_t)idx * 3, p.bitfield, desc.bound, p.dt_gamma, p.max_steps, p.C, p.H,
                        __ldg(p.fars + idx));
                t = __ldg(p.nears + idx);
                last_t = t;
                if (wk.next_sample(t, px, py, pz, pdt)) {
                    pd1 = __fsub_rn(t, last_t);
                    last_t = t;
                    nsteps = 1;
                    ray = (int32_t)idx;
                    have_pending = true;
                    break;
                }
                // no sample at all: the outputs of this ray stay at the zeros the host wrote
            }
            if (exhausted) {
                atomicMax(&ctl->last_tile, my_last);
                __threadfence_block();
                atomicAdd(&ctl->n_done, 1);
            }
        }
        QStage& qs = q[s];
        if (have_pending) {
            float nx = 0.f, ny = 0.f, nz = 0.f, ndt = 0.f;
            const bool more = nsteps < p.max_steps && wk.next_sample(t, nx, ny, nz, ndt);   // look one sample ahead
            qs.x[r] = px; qs.y[r] = py; qs.z[r] = pz; qs.dt[r] = pdt; qs.d1[r] = pd1;
            qs.dx[r] = wk.dx; qs.dy[r] = wk.dy; qs.dz[r] = wk.dz;
            qs.ray[r] = ray;
            qs.end[r] = more ? 0 : 1;
            my_last = (int32_t)tile;
            if (more) {
                pd1 = __fsub_rn(t, last_t);
                last_t = t;
                nsteps++;
                px = nx; py = ny; pz = nz; pdt = ndt;
            } else {
                have_pending = false;
                ray = -1;
            }
        } else {
            qs.ray[r] = -1;
        }
        umma::mbar_arrive(&ctl->q_full[s]);
    }
}

// ----------------------------------------------------------------------------------------------- gather --
__device__ __forceinline__ void gather_role(const inerf_field_desc& desc, uint8_t* smem, Ctrl* ctl, QStage* q, uint32_t gt) {
    const uint32_t row = gt & (kTile - 1), half = gt >> 7;
    const float inv2b = __fdiv_rn(1.0f, __fmul_rn(2.0f, desc.bound));
    const uint2* table = reinterpret_cast<const uint2*>(desc.table_packed);
    for (uint32_t tile = 0;; tile++) {
        const uint32_t sq = tile % DQ, sa = tile % DA;
        umma::mbar_wait(&ctl->q_full[sq], (tile / DQ) & 1u);
        const int32_t nd = ld_vol(&ctl->n_done);
        __threadfence_block();
        const int32_t lt = ld_vol(&ctl->last_tile);
        const bool stop = nd == (int32_t)kMarchT && (int32_t)tile > lt;
        const QStage& qs = q[sq];
        int32_t ray = -1;
        float x = 0.f, y = 0.f, z = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
        if (!stop) {
            ray = qs.ray[row];
            if (ray >= 0) {
                x = qs.x[row]; y = qs.y[row]; z = qs.z[row];
                if (half == 0) { dx = qs.dx[row]; dy = qs.dy[row]; dz = qs.dz[row]; }
            }
        }
        umma::mbar_arrive(&ctl->q_empty[sq]);
        if (tile >= DA) umma::mbar_wait(&ctl->a_empty[sa], ((tile / DA) - 1u) & 1u);
        if (gt == 0) ctl->a_flag[sa] = stop ? 1 : 0;
        if (stop) { umma::mbar_arrive(&ctl->a_full[sa]); break; }
        if (ray >= 0 && ld_vol(&ctl->kill[row]) != ray) {
            const uint32_t a_es = RSmem::A + sa * kStageBytes, a_ci = a_es + kBytesEs, a_mi = a_ci + kBytesCi;
            float x01[3];
            x01[0] = __fmul_rn(__fadd_rn(x, desc.bound), inv2b);
            x01[1] = __fmul_rn(__fadd_rn(y, desc.bound), inv2b);
            x01[2] = __fmul_rn(__fadd_rn(z, desc.bound), inv2b);
            const bool oob = x01[0] < 0.f || x01[0] > 1.f || x01[1] < 0.f || x01[1] > 1.f || x01[2] < 0.f || x01[2] > 1.f;
            encode8(x01, oob, half * 8, ctl->lg, table, smem, a_es, a_mi, row);
            if (half == 0) sh16_to_smem(dx, dy, dz, smem, a_ci, row);
        }
        umma::fence_async_smem();
        umma::mbar_arrive(&ctl->a_full[sa]);
    }
}

// ------------------------------------------------------------------------------------------------ chain --
template <int NCH>  // 16-column logit chunks owned per thread: 1 -> K <= 32, 2 -> K <= 64
__device__ __forceinline__ void chain_role(const inerf_field_desc& desc, const RenderParams& p, uint8_t* smem, Ctrl* ctl, QStage* q,
                                           uint32_t tmem_base, uint32_t ct) {
    const uint32_t K = desc.K, Kp = weight_layout(K).Kp;
    const bool with_masks = p.mask_out != nullptr;
    const uint32_t warp = ct >> 5, lane = ct & 31;
    const uint32_t row = ct & (kTile - 1), half = ct >> 7;
    const bool owner = half == 0;
    const uint32_t chunks = Kp / 16;
    const uint32_t c_begin = half ? (chunks + 1) / 2 : 0, c_end = half ? chunks : (chunks + 1) / 2;
    const bool vec_ok = (K & 3u) == 0;

    uint32_t phase = 0, my_samples = 0;
    int32_t cur_ray = -1, dead = -1;
    float t_depth = 0.f, ws = 0.f, dep = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    float macc[NCH][16];
#pragma unroll
    for (int c = 0; c < NCH; c++)
#pragma unroll
        for (int i = 0; i < 16; i++) macc[c][i] = 0.f;

    auto chain_sync = [] { umma::named_sync<1, kChainT>(); };

    for (uint32_t tile = 0;; tile++) {
        const uint32_t sq = tile % DQ, sa = tile % DA;
        umma::mbar_wait(&ctl->a_full[sa], (tile / DA) & 1u);
        if (ld_vol(&ctl->a_flag[sa])) break;
        const uint32_t a_es = RSmem::A + sa * kStageBytes;
        const ChainBufs bufs{a_es, a_es + kBytesEs, a_es + kBytesEs + kBytesCi, RSmem::H1, RSmem::H2, RSmem::W};
        float weight = 0.f;
        mlp_chain(smem, bufs, tmem_base, &ctl->mma_bar, phase, K, desc.density_scale, with_masks, ct, &ctl->a_empty[sa], chain_sync,
                  [&](float sigma) {
                      if (owner) {
                          const QStage& qs = q[sq];
                          const int32_t ray = qs.ray[row];
                          int32_t fin = -1;
                          if (ray >= 0 && ray != dead) {
                              if (ray != cur_ray) { cur_ray = ray; t_depth = __ldg(p.nears + ray); }
                              const float dt = qs.dt[row], d1 = qs.d1[row];
                              const int32_t endf = qs.end[row];
                              const float alpha = 1.0f - __expf(-sigma * dt);
                              const float T = 1.0f - ws;
                              weight = alpha * T;
                              ws += weight;
                              t_depth += d1;
                              dep = fmaf(weight, t_depth, dep);
                              my_samples++;
                              const bool term = T < p.T_thresh;
                              if (term || endf) fin = ray;
                              if (term) {
                                  dead = ray;
                                  if (!endf) st_vol(&ctl->kill[row], ray);
                              }
                          }
                          umma::mbar_arrive(&ctl->q_empty[sq]);
                          ctl->w_s[row] = weight;
                          ctl->fin_s[row] = fin;
                      }
                  });

        if (owner) {
            float rgb[3];
            epilogue_rgb(tmem_base, rgb, ct);
            cr = fmaf(weight, rgb[0], cr); cg = fmaf(weight, rgb[1], cg); cb = fmaf(weight, rgb[2], cb);
        }
        if (with_masks) {
            const float wgt = ctl->w_s[row];
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
        const int32_t fin = ctl->fin_s[row];
        if (fin >= 0) {   // this ray is complete: write its pixel and reset the slot's accumulators
            if (owner) {
                p.weights_sum[fin] = ws;
                p.depth[fin] = dep;
                p.image[(size_t)fin * 3] = cr; p.image[(size_t)fin * 3 + 1] = cg; p.image[(size_t)fin * 3 + 2] = cb;
                ws = dep = cr = cg = cb = 0.f;
            }
            if (with_masks) {
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    if (c_begin + c < c_end) {
                        const uint32_t k0 = (c_begin + c) * 16;
                        float* out = p.mask_out + (size_t)fin * K + k0;
                        if (vec_ok && k0 + 16 <= K) {
#pragma unroll
                            for (int i = 0; i < 4; i++)
                                reinterpret_cast<float4*>(out)[i] = make_float4(macc[c][4 * i], macc[c][4 * i + 1], macc[c][4 * i + 2], macc[c][4 * i + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; i++)
                                if (k0 + i < K) out[i] = macc[c][i];
                        }
#pragma unroll
                        for (int i = 0; i < 16; i++) macc[c][i] = 0.f;
                    }
                }
            }
        }
        umma::fence_before_sync();
        chain_sync();   // TMEM, the chain's operand tiles and w_s / fin_s are reused by the next tile
    }
    // samples composited by this launch -> work_counter[2..3] (u64), one atomic per owner warp
    if (owner) {
        uint32_t tot = my_samples;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0 && tot) atomicAdd(reinterpret_cast<unsigned long long*>(p.work_counter + 2), (unsigned long long)tot);
    }
}

template <int NCH>
__global__ void __launch_bounds__(kThreadsR, 1) k_render_fused(inerf_field_desc desc, RenderParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t K = desc.K;
    Ctrl* ctl = reinterpret_cast<Ctrl*>(smem + RSmem::ctrl(K));
    QStage* q = reinterpret_cast<QStage*>(smem + RSmem::queue(K));
    const uint32_t tid = threadIdx.x;

    load_weights(smem, RSmem::W, desc.weights, K);
    init_levels(ctl->lg, desc.offsets, desc.L, desc.S, desc.H, tid);
    if (tid == 0) {
        for (uint32_t i = 0; i < DQ; i++) { umma::mbar_init(&ctl->q_full[i], kMarchT); umma::mbar_init(&ctl->q_empty[i], kGatherT + kTile); }
        for (uint32_t i = 0; i < DA; i++) { umma::mbar_init(&ctl->a_full[i], kGatherT); umma::mbar_init(&ctl->a_empty[i], 1); ctl->a_flag[i] = 0; }
        umma::mbar_init(&ctl->mma_bar, 1);
        ctl->n_done = 0;
        ctl->last_tile = -1;
        umma::mbar_fence_init();
    }
    if (tid < kTile) ctl->kill[tid] = -1;
    if (tid < 32) umma::tmem_alloc<kTmemCols>(&ctl->tmem_slot);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = ctl->tmem_slot;

    if (tid < kChainT) chain_role<NCH>(desc, p, smem, ctl, q, tmem_base, tid);
    else if (tid < kChainT + kGatherT) gather_role(desc, smem, ctl, q, tid - kChainT);
    else march_role(desc, p, ctl, q, tid - kChainT - kGatherT);

    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

extern "C" int inerf_render_fused(const inerf_field_desc* desc, const float* rays_o, const float* rays_d, const float* nears,
                                  const float* fars, const uint8_t* bitfield, uint32_t N, uint32_t C, uint32_t H, float dt_gamma,
                                  uint32_t max_steps, float T_thresh, float* weights_sum, float* depth, float* image,
                                  float* mask_out, int32_t* work_counter, void* stream) {
    if (int e = field::validate(desc)) return e;
    if (C == 0 || C > 16 || H == 0 || H > 1024 || (H & (H - 1)) || max_steps == 0 || N >= 0x7fffffffu) return INERF_ERR_SIZE;
    if (N == 0) return INERF_OK;
    INERF_REQUIRE(rays_o); INERF_REQUIRE(rays_d); INERF_REQUIRE(nears); INERF_REQUIRE(fars); INERF_REQUIRE(bitfield);
    INERF_REQUIRE(weights_sum); INERF_REQUIRE(depth); INERF_REQUIRE(image); INERF_REQUIRE(work_counter);
    if ((uintptr_t)work_counter & 7u) return INERF_ERR_ALIGN;
    if (mask_out && ((uintptr_t)mask_out & 15u)) return INERF_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce = cudaMemsetAsync(work_counter, 0, 4 * sizeof(int32_t), st);
    if (ce != cudaSuccess) return (int)ce;
    // rays without a single sample are never touched by the kernel: their outputs are these zeros
    if ((ce = cudaMemsetAsync(weights_sum, 0, (size_t)N * 4, st)) != cudaSuccess) return (int)ce;
    if ((ce = cudaMemsetAsync(depth, 0, (size_t)N * 4, st)) != cudaSuccess) return (int)ce;
    if ((ce = cudaMemsetAsync(image, 0, (size_t)N * 12, st)) != cudaSuccess) return (int)ce;
    if (mask_out && (ce = cudaMemsetAsync(mask_out, 0, (size_t)N * desc->K * 4, st)) != cudaSuccess) return (int)ce;
    RenderParams p{rays_o, rays_d, nears, fars, bitfield, N, C, H, max_steps, dt_gamma, T_thresh, weights_sum, depth, image, mask_out, work_counter};
    const uint32_t smem_bytes = RSmem::bytes(desc->K);
    static bool attr_set = false;
    if (!attr_set) {
        ce = cudaFuncSetAttribute(k_render_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce != cudaSuccess) return (int)ce;
        ce = cudaFuncSetAttribute(k_render_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce != cudaSuccess) return (int)ce;
        attr_set = true;
    }
    const uint32_t want = (N + field::kTile - 1) / field::kTile;
    const uint32_t grid = want < (uint32_t)kNumSMs ? want : (uint32_t)kNumSMs;
    if (desc->K <= 32) k_render_fused<1><<<grid, kThreadsR, smem_bytes, st>>>(*desc, p);
    else k_render_fused<2><<<grid, kThreadsR, smem_bytes, st>>>(*desc, p);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}
