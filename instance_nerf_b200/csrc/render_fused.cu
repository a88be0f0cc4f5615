// inerf_render_fused: whole-frame instance-field inference in ONE persistent, warp-specialised launch.
//
// Replaces the host loop of NeRFMaskRenderer.run_cuda (nerf/mask_renderer.py:322-381): per iteration the
// reference launches march_rays, two full-table fp16 casts, two grid encodes, 8 GEMMs + elementwise kernels,
// composite_rays_with_masks and a boolean-index compaction, with two device->host syncs.
//
// One CTA per SM, 896 threads in three roles that run concurrently and hand 128-sample tiles to each other
// through mbarrier-guarded rings in shared memory (nothing per-sample ever touches HBM):
//
//   march  (4 warps, thread = ray slot)   occupancy-grid DDA (raymarching.cu:1008-1062 semantics) into a 24-deep
//                                          per-slot sample ring; runs AHEAD of compositing (speculatively: a ray
//                                          killed by T < T_thresh drops its queued samples); a warp whose 32 slots
//                                          are all free pulls the next 32 consecutive rays from a global counter.
//   gather (16 warps, thread = slot x 4    16 levels x 8 corners, one 8-byte gather per corner from the interleaved
//           of the 16 levels)              fp16 table (both encoders at once), fp16 trilinear blend, SH degree 4 ->
//                                          tcgen05.st into the TENSOR-MEMORY operand stage of the tile, double buffered.
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

constexpr uint32_t kChainT = 256, kGatherT = 512, kMarchT = 128;
// register budget per role (setmaxnreg): 896 threads start at 72; 256*96 + 512*64 + 128*56 = 896*72
constexpr uint32_t kRegsChain = 96, kRegsGather = 64, kRegsMarch = 56;
// setmaxnreg moves registers inside the CTA's OWN allocation (threads x the launch register count, 896 x 72 here), not inside
// the SM's 64 K file: a role budget that sums to more than that never gets its registers and the kernel hangs in setmaxnreg.inc
static_assert(kChainT * kRegsChain + kGatherT * kRegsGather + kMarchT * kRegsMarch <= (kChainT + kGatherT + kMarchT) * 72,
              "register budget of the three roles");
constexpr uint32_t kThreadsR = kChainT + kGatherT + kMarchT;
// samples queued per ray slot.  Shared memory holds only the weights, the rings, the coarse occupancy bitmap and control words (every
// MMA A operand is in tensor memory): 32 x 3 KB of rings keep the CTA inside the 164 KB carve-out (92 KB of L1 left).  Measured on
// B200 (c2), with the operand stages still in shared memory: 10 / 14 / 20 / 24 deep -> tile fill 0.905 / 0.921 / 0.943 / 0.955,
// 8.92 / 8.77 / 8.71 / 8.68 ms; with them in tensor memory: 24 / 32 / 40 deep -> 8.22 / 8.15 / 8.34 ms (40 needs the 196 KB carve-out)
constexpr uint32_t RING = 32;
constexpr uint32_t DT = 4;    // tile descriptors in flight (gather may run DA tiles ahead of the chain)
constexpr uint32_t DA = 2;    // gathered operand stages
constexpr int kMarchCells = 8;   // occupancy cells a marcher lane may evaluate per warp iteration
// (the alternatives measured on B200 and rejected -- ring depth, cells / skip steps per iteration, tile hold-back, one ray per
// free lane instead of 32-ray patches, a software-pipelined gather -- are recorded with their numbers in DESIGN.md 4.1)

struct RenderParams {
    const float* rays_o;
    const float* rays_d;
    const float* nears;
    const float* fars;
    const float* noises;     // optional per-ray jitter of the first step (perturb, raymarching.cu:1004); NULL = none
    const uint8_t* bitfield;
    uint32_t N, C, H, max_steps;
    float dt_gamma, T_thresh;
    float* weights_sum;
    float* depth;
    float* image;
    float* mask_out;
    int32_t* work_counter;
    uint32_t coarse_bytes;   // size of the shared-memory coarse occupancy bitmap (0 = disabled)
};

// Per-slot sample rings (SoA over the 128 ray slots).  dt == 0 marks an END entry (its ray has no more samples).
struct Rings {
    float x[RING][kTile], y[RING][kTile], z[RING][kTile], dt[RING][kTile], d1[RING][kTile];
    int32_t ray[RING][kTile];
};

struct Ctrl {
    uint64_t a_full[DA], a_empty[DA], mma_bar;
    uint64_t sel_ready[DT];       // tile descriptor st is published (4 selector warps arrive)
    uint32_t sel_any[DT][4];      // per selector warp: some slot of its 32 had a sample for that tile
    uint32_t tmem_slot;
    int32_t n_done;               // marcher lanes that ran out of rays
    int32_t a_flag[DA];           // 1 = terminal tile
    uint32_t tail[kTile];         // entries produced per slot (marcher)
    uint32_t chead[kTile];        // entries retired per slot (compositor)
    int32_t kill[kTile];          // ray id terminated early by the compositor (marcher drops it)
    int32_t tsel[DT][kTile];      // per tile: ring entry taken for each row, -1 = bubble
    float w_s[kTile];
    int32_t fin_s[kTile];
    LevelGeom lg[16];
};

struct RSmem {
    static constexpr uint32_t W = 0;   // weights first: the DA operand stages (es | ci | mi) and the hidden activations live in tensor memory
    static __host__ __device__ uint32_t ctrl(uint32_t K) { return (W + weight_layout(K).total + 15u) & ~15u; }
    static __host__ __device__ uint32_t rings(uint32_t K) { return (ctrl(K) + (uint32_t)sizeof(Ctrl) + 15u) & ~15u; }
    static __host__ __device__ uint32_t coarse(uint32_t K) { return rings(K) + (uint32_t)sizeof(Rings); }
    // + coarse occupancy bitmap: C*H^3/64 bits
    static __host__ __device__ uint32_t bytes(uint32_t K, uint32_t coarse_bytes) { return coarse(K) + coarse_bytes; }
};

__device__ __forceinline__ int32_t ld_vol(const int32_t* p) { return *reinterpret_cast<const volatile int32_t*>(p); }
__device__ __forceinline__ uint32_t ld_vol(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void st_vol(int32_t* p, int32_t v) { *reinterpret_cast<volatile int32_t*>(p) = v; }
__device__ __forceinline__ void st_vol(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }
// OR-reduction of a predicate over the kGatherT threads of the gather group (named barrier 2)
__device__ __forceinline__ bool gather_any(bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
        "barrier.cta.red.or.pred p, 2, %2, q;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(r) : "r"((uint32_t)pred), "n"(kGatherT) : "memory");
    return r != 0;
}

// ------------------------------------------------------------------------------------------------ march --
// Per-lane state machine; every warp iteration each lane does ONE bounded unit of work (fetch a ray / evaluate up to
// kMarchCells occupancy cells), so a lane crossing a long empty stretch never holds back the other 31.  The sample
// sequence per ray is the reference's (raymarching.cu:1008-1062).
__device__ __forceinline__ void march_role(const inerf_field_desc& desc, const RenderParams& p, Ctrl* ctl, Rings* rg, const uint32_t* coarse,
                                           uint32_t r) {
    enum : int { NEED_RAY = 0, EVAL = 1, DONE = 3 };
    march::Walk wk;
    wk.coarse = coarse;
    int state = NEED_RAY;
    int32_t ray = -1;
    uint32_t nsteps = 0, tail = 0;
    float t = 0.f, last_t = 0.f, tt = 0.f;
    while (true) {
        if (state != DONE && ray >= 0 && ld_vol(&ctl->kill[r]) == ray) { ray = -1; state = NEED_RAY; }
        const uint32_t chead_seen = ld_vol(&ctl->chead[r]);
        const bool room = tail - chead_seen < RING;
        bool worked = false;
        // Rays are taken 32 at a time: a warp's lanes always march 32 CONSECUTIVE rays (neighbouring pixels), sample for
        // sample in step, so the gather warp that consumes these 32 slots reads neighbouring cells -- one L1 sector serves
        // several lanes at the coarse and middle levels (measured 87 -> ~41 sectors per sample).  A lane
        // whose ray ends early idles until its 31 neighbours have finished marching (not compositing): ~5 % bubble rows.
        const bool acquire = __all_sync(0xffffffffu, state == NEED_RAY || state == DONE) && state == NEED_RAY;
        if (state == NEED_RAY && acquire) {
            uint32_t base = 0;
            const uint32_t takers = __activemask();
            const uint32_t leader = __ffs(takers) - 1u;
            if ((r & 31u) == leader) base = (uint32_t)atomicAdd(p.work_counter, 32);
            base = __shfl_sync(takers, base, leader);
            const uint32_t idx = base + (r & 31u);
            if (idx >= p.N) {
                state = DONE;
                atomicAdd(&ctl->n_done, 1);
            } else {
                wk.init(p.rays_o + (size_t)idx * 3, p.rays_d + (size_t)idx * 3, p.bitfield, desc.bound, p.dt_gamma, p.max_steps, p.C, p.H,
                        __ldg(p.fars + idx));
                // k_first_hit already walked this ray's leading empty space (same DDA, same t sequence): resume at the first
                // occupied cell.  The value travels in depth[idx], which goes back to 0 (the output of a ray with no sample).
                t = *reinterpret_cast<const volatile float*>(p.depth + idx);
                p.depth[idx] = 0.f;
                last_t = __ldg(p.nears + idx);
                if (p.noises) last_t = __fmaf_rn(wk.step_size(last_t), __ldg(p.noises + idx), last_t);   // :1004, as k_first_hit started
                nsteps = 0;
                ray = (int32_t)idx;
                state = EVAL;
            }
            worked = true;
        } else if (state == EVAL && room) {
            worked = true;
            // up to kMarchCells occupancy cells per warp iteration (the loop's vote / ring bookkeeping is paid once), emitting a
            // sample for every occupied one while the ring has room: every lane with a ray does the same amount of work per
            // iteration whether it is crossing empty space or a solid
            uint32_t free_entries = RING - (tail - chead_seen);
#pragma unroll 1
            for (int c = 0; c < kMarchCells && free_entries; c++) {
                const uint32_t e = tail % RING;
                if (!(t < wk.far) || nsteps >= p.max_steps) {
                    if (nsteps > 0) {   // END marker; a ray without samples is never seen by the compositor (outputs pre-zeroed)
                        rg->dt[e][r] = 0.f;
                        rg->ray[e][r] = ray;
                        __threadfence_block();
                        st_vol(&ctl->tail[r], ++tail);
                    }
                    ray = -1;
                    state = NEED_RAY;
                    break;
                }
                float x, y, z, dt;
                if (wk.eval_cell(t, x, y, z, dt, tt)) {
                    rg->x[e][r] = x; rg->y[e][r] = y; rg->z[e][r] = z;
                    rg->dt[e][r] = dt;
                    const float d1 = __fsub_rn(t, last_t);
                    rg->d1[e][r] = d1;
                    rg->ray[e][r] = ray;
                    last_t = t;
                    nsteps++;
                    // perturb, as the reference's loop really behaves: its first iteration marches ONE jittered sample per ray
                    // (n_step = N / n_alive = 1, mask_renderer.py:333) and composite_rays then stores rays_t = near + deltas[1]
                    // (raymarching.cu:1216-1251 starts from the UNjittered rays_t), so the march resumes at near + d1
                    if (p.noises != nullptr && nsteps == 1) { t = __fadd_rn(__ldg(p.nears + ray), d1); last_t = t; }
                    __threadfence_block();
                    st_vol(&ctl->tail[r], ++tail);
                    free_entries--;
                    continue;
                }
                do { t = __fadd_rn(t, wk.step_size(t)); } while (t < tt);
            }
        }
        if (__all_sync(0xffffffffu, state == DONE)) break;
        if (!__any_sync(0xffffffffu, worked)) __nanosleep(256);
    }
}

// ----------------------------------------------------------------------------------------------- gather --
__device__ __forceinline__ void gather_role(const inerf_field_desc& desc, const RenderParams& p, uint8_t* smem, Ctrl* ctl, Rings* rg,
                                            uint32_t gt, uint32_t tmem_base) {
    const uint32_t row = gt & (kTile - 1), quarter = gt >> 7;   // thread = (ray slot, 4 of the 16 levels)
    const float inv2b = __fdiv_rn(1.0f, __fmul_rn(2.0f, desc.bound));
    const uint2* table = reinterpret_cast<const uint2*>(desc.table_packed);
    uint32_t ghead = 0;   // entries of this row already handed to a tile (quarter-0 threads)
    // Tile descriptors are made ONE TILE AHEAD: the four selector warps (quarter 0) describe tile t + 1 before they gather tile t
    // and publish it through an mbarrier, so nobody waits for the selection at the top of an iteration (the group-wide barrier
    // that used to sit there cost every gather warp its latency once per tile).  Only when a descriptor came out empty does the
    // group fall back to the blocking selection below.
    auto preselect = [&](uint32_t t) {
        const uint32_t st1 = t % DT;
        int32_t sel = -1;
        if (ld_vol(&ctl->tail[row]) != ghead) { sel = (int32_t)(ghead % RING); ghead++; }
        ctl->tsel[st1][row] = sel;
        const bool any = __any_sync(0xffffffffu, sel >= 0);
        if ((gt & 31u) == 0) ctl->sel_any[st1][gt >> 5] = any ? 1u : 0u;
        __syncwarp();
        if ((gt & 31u) == 0) umma::mbar_arrive(&ctl->sel_ready[st1]);
    };
    if (quarter == 0) preselect(0);
    for (uint32_t tile = 0;; tile++) {
        const uint32_t st = tile % DT, sa = tile % DA;
        bool stop = false;
        umma::mbar_wait(&ctl->sel_ready[st], (tile / DT) & 1u);
        if (!(ld_vol(&ctl->sel_any[st][0]) | ld_vol(&ctl->sel_any[st][1]) | ld_vol(&ctl->sel_any[st][2]) | ld_vol(&ctl->sel_any[st][3]))) {
            while (true) {   // empty descriptor: blocking selection (start of the frame, marchers behind, end of the frame)
                int32_t sel = -1;
                bool not_finished = false;
                if (quarter == 0) {
                    const int32_t nd = ld_vol(&ctl->n_done);
                    __threadfence_block();
                    if (ld_vol(&ctl->tail[row]) != ghead) sel = (int32_t)(ghead % RING);
                    not_finished = sel >= 0 || nd != (int32_t)kMarchT;
                    ctl->tsel[st][row] = sel;
                }
                if (gather_any(sel >= 0)) {
                    if (sel >= 0) ghead++;
                    break;
                }
                if (!gather_any(not_finished)) { stop = true; break; }   // all marchers done and every ring drained
                __nanosleep(128);
            }
        }
        if (quarter == 0 && !stop) preselect(tile + 1);
        if (tile >= DA) umma::mbar_wait(&ctl->a_empty[sa], ((tile / DA) - 1u) & 1u);
        if (gt == 0) ctl->a_flag[sa] = stop ? 1 : 0;
        if (stop) {
            __syncwarp();
            if ((gt & 31u) == 0) umma::mbar_arrive(&ctl->a_full[sa]);
            break;
        }
        const int32_t e = ctl->tsel[st][row];
        uint32_t fs[4] = {0u, 0u, 0u, 0u}, fm[4] = {0u, 0u, 0u, 0u}, sh[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (e >= 0) {
            __threadfence_block();
            const int32_t ray = rg->ray[e][row];
            if (rg->dt[e][row] != 0.f && ld_vol(&ctl->kill[row]) != ray) {
                float x01[3];
                x01[0] = __fmul_rn(__fadd_rn(rg->x[e][row], desc.bound), inv2b);
                x01[1] = __fmul_rn(__fadd_rn(rg->y[e][row], desc.bound), inv2b);
                x01[2] = __fmul_rn(__fadd_rn(rg->z[e][row], desc.bound), inv2b);
                const bool oob = x01[0] < 0.f || x01[0] > 1.f || x01[1] < 0.f || x01[1] > 1.f || x01[2] < 0.f || x01[2] > 1.f;
                if (quarter == 0) {
                    const float* d = p.rays_d + (size_t)ray * 3;
                    sh16_pack(__ldg(d), __ldg(d + 1), __ldg(d + 2), sh);
                }
                encode4_regs(x01, oob, quarter * 4, ctl->lg, table, fs, fm);
            }
        }
        // every lane of the warp stores its row (bubble rows: zeros) into the tensor-memory stage: warp w of the CTA owns TMEM
        // lanes 32 (w % 4) .., and gather warp gw holds rows 32 (gw % 4) ..
        __syncwarp();
        {
            const uint32_t stage = tmem_base + T_stage0 + sa * kStageCols + ((((gt >> 5) & 3u) * 32u) << 16);
            umma::tmem_st4(stage + S_es + quarter * 4, fs);
            umma::tmem_st4(stage + S_mi + quarter * 4, fm);
            if (quarter == 0) umma::tmem_st8(stage + S_ci, sh);
            umma::tmem_st_wait();
            umma::fence_before_sync();
        }
        __syncwarp();
        if ((gt & 31u) == 0) umma::mbar_arrive(&ctl->a_full[sa]);
    }
}


// ------------------------------------------------------------------------------------------------ chain --
template <int NCH>  // 16-column logit chunks owned per thread: 1 -> K <= 32, 2 -> K <= 64
__device__ __forceinline__ void chain_role(const inerf_field_desc& desc, const RenderParams& p, uint8_t* smem, Ctrl* ctl, Rings* rg,
                                           uint32_t tmem_base, uint32_t ct) {
    const uint32_t K = desc.K, Kp = weight_layout(K).Kp;
    const bool with_masks = p.mask_out != nullptr;
    const uint32_t warp = ct >> 5, lane = ct & 31;
    const uint32_t row = ct & (kTile - 1), half = ct >> 7;
    const bool owner = half == 0;
    const uint32_t chunks = Kp / 16;
    const uint32_t c_begin = half ? (chunks + 1) / 2 : 0, c_end = half ? chunks : (chunks + 1) / 2;
    const bool vec_ok = (K & 3u) == 0;

    uint32_t phase = 0, my_samples = 0, chead = 0;
    int32_t cur_ray = -1, dead = -1;
    float t_depth = 0.f, ws = 0.f, dep = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    float macc[NCH][16];
#pragma unroll
    for (int c = 0; c < NCH; c++)
#pragma unroll
        for (int i = 0; i < 16; i++) macc[c][i] = 0.f;

    auto chain_sync = [] { umma::named_sync<1, kChainT>(); };

    uint32_t tile_count = 0;
    for (uint32_t tile = 0;; tile++) {
        const uint32_t st = tile % DT, sa = tile % DA;
        umma::mbar_wait(&ctl->a_full[sa], (tile / DA) & 1u);
        tile_count = tile;
        if (ld_vol(&ctl->a_flag[sa])) break;
        const ChainBufs bufs{0u, 0u, 0u, RSmem::W};   // no shared-memory operand tiles: inputs and hidden activations are in TMEM
        float weight = 0.f;
        umma::fence_after_sync();
        mlp_chain<true>(smem, bufs, tmem_base, &ctl->mma_bar, phase, K, desc.density_scale, with_masks, ct, &ctl->a_empty[sa], chain_sync,
                  [&](float sigma) {
                      if (owner) {
                          const int32_t e = ctl->tsel[st][row];
                          int32_t fin = -1;
                          if (e >= 0) {
                              const int32_t ray = rg->ray[e][row];
                              const float dt = rg->dt[e][row], d1 = rg->d1[e][row];
                              st_vol(&ctl->chead[row], ++chead);   // entry retired: the marcher may reuse it
                              if (ray != dead) {
                                  if (dt == 0.f) {
                                      if (ray == cur_ray) fin = ray;   // END marker: the ray is complete
                                  } else {
                                      if (ray != cur_ray) { cur_ray = ray; t_depth = __ldg(p.nears + ray); }
                                      const float alpha = 1.0f - __expf(-sigma * dt);
                                      const float T = 1.0f - ws;
                                      weight = alpha * T;
                                      ws += weight;
                                      t_depth += d1;
                                      dep = fmaf(weight, t_depth, dep);
                                      my_samples++;
                                      if (T < p.T_thresh) {   // raymarching.cu:1252: terminated, later samples of this ray are dropped
                                          fin = ray;
                                          dead = ray;
                                          st_vol(&ctl->kill[row], ray);
                                      }
                                  }
                              }
                          }
                          ctl->w_s[row] = weight;
                          ctl->fin_s[row] = fin;
                      }
                  }, tmem_base + T_stage0 + sa * kStageCols);

        if (owner) {
            float rgb[3];
            epilogue_rgb(tmem_base, rgb, ct);
            if (weight != 0.f) {   // bubble rows hold stale operands: never let them touch the accumulators
                cr = fmaf(weight, rgb[0], cr); cg = fmaf(weight, rgb[1], cg); cb = fmaf(weight, rgb[2], cb);
            }
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
                    if (wgt != 0.f) {
#pragma unroll
                        for (int i = 0; i < 16; i++) macc[c][i] = fmaf(wgt, __half2float(__float2half_rn(__uint_as_float(v[i]))), macc[c][i]);
                    }
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
                cur_ray = -1;
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
    // tiles evaluated by this CTA -> work_counter[1] (rows = 128 x tiles; rows / samples = 1 + bubble share)
    if (ct == 0) atomicAdd(p.work_counter + 1, (int32_t)tile_count);
    // samples composited by this launch -> work_counter[2..3] (u64), one atomic per owner warp
    if (owner) {
        uint32_t tot = my_samples;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0 && tot) atomicAdd(reinterpret_cast<unsigned long long*>(p.work_counter + 2), (unsigned long long)tot);
    }
}

// Pre-pass: walk every ray from `near` through its leading empty space to the first occupied cell, one thread per ray at
// full occupancy (the walk is a serial chain of ~10^4 dependent instructions per ray; inside the persistent kernel it
// would run on 4 marcher warps per SM and starve the gather / MLP pipeline at every ray start: 12 % bubble rows at c2).
// t_first[i] = parameter at which the render kernel's marcher resumes (same eval_cell / step sequence, so the samples
// are unchanged); >= far when the ray never meets an occupied cell.
__global__ void __launch_bounds__(256) k_first_hit(RenderParams p, float bound, float* __restrict__ t_first) {
    const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
    if (idx >= p.N) return;
    march::Walk wk;
    wk.init(p.rays_o + (size_t)idx * 3, p.rays_d + (size_t)idx * 3, p.bitfield, bound, p.dt_gamma, p.max_steps, p.C, p.H, __ldg(p.fars + idx));
    float t = __ldg(p.nears + idx), tt = 0.f;
    if (p.noises) t = __fmaf_rn(wk.step_size(t), __ldg(p.noises + idx), t);   // raymarching.cu:1004
    while (t < wk.far) {
        const float t0 = t;
        float x, y, z, dt;
        if (wk.eval_cell(t, x, y, z, dt, tt)) { t = t0; break; }
        do { t = __fadd_rn(t, wk.step_size(t)); } while (t < tt);
    }
    t_first[idx] = t;
}

template <int NCH>
__global__ void __launch_bounds__(kThreadsR, 1) k_render_fused(inerf_field_desc desc, RenderParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t K = desc.K;
    Ctrl* ctl = reinterpret_cast<Ctrl*>(smem + RSmem::ctrl(K));
    Rings* rg = reinterpret_cast<Rings*>(smem + RSmem::rings(K));
    const uint32_t tid = threadIdx.x;

    // (rows of a tile without a sample -- bubbles -- get zero operand rows from the gather warps; their outputs are never used)
    load_weights(smem, RSmem::W, desc.weights, K);
    init_levels(ctl->lg, desc.offsets, desc.L, desc.S, desc.H, tid);
    if (tid == 0) {
        for (uint32_t i = 0; i < DA; i++) { umma::mbar_init(&ctl->a_full[i], kGatherT / 32); umma::mbar_init(&ctl->a_empty[i], 1); ctl->a_flag[i] = 0; }
        umma::mbar_init(&ctl->mma_bar, 1);
        for (uint32_t i = 0; i < DT; i++) umma::mbar_init(&ctl->sel_ready[i], 4);
        ctl->n_done = 0;
        umma::mbar_fence_init();
    }
    if (tid < kTile) { ctl->kill[tid] = -1; ctl->tail[tid] = 0; ctl->chead[tid] = 0; }
    // coarse occupancy: bit b = (8-byte word b of the bitfield != 0), i.e. any of the 64 cells of that 4x4x4 block occupied
    uint32_t* coarse = p.coarse_bytes ? reinterpret_cast<uint32_t*>(smem + RSmem::coarse(K)) : nullptr;
    if (coarse) {
        const uint32_t n_blocks = p.coarse_bytes * 8u;   // multiple of 32
        const unsigned long long* words = reinterpret_cast<const unsigned long long*>(p.bitfield);
        for (uint32_t base = (tid >> 5) * 32u; base < n_blocks; base += (kThreadsR / 32u) * 32u) {
            const uint32_t bits = __ballot_sync(0xffffffffu, __ldg(words + base + (tid & 31u)) != 0ull);
            if ((tid & 31u) == 0) coarse[base >> 5] = bits;
        }
    }
    if (tid < 32) umma::tmem_alloc<kTmemCols>(&ctl->tmem_slot);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = ctl->tmem_slot;

    if (tid < kChainT) {
        umma::reg_alloc<kRegsChain>();
        chain_role<NCH>(desc, p, smem, ctl, rg, tmem_base, tid);
    } else if (tid < kChainT + kGatherT) {
        umma::reg_dealloc<kRegsGather>();
        gather_role(desc, p, smem, ctl, rg, tid - kChainT, tmem_base);
    } else {
        umma::reg_dealloc<kRegsMarch>();
        march_role(desc, p, ctl, rg, coarse, tid - kChainT - kGatherT);
    }

    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

extern "C" int inerf_render_fused_perturb(const inerf_field_desc* desc, const float* rays_o, const float* rays_d, const float* nears,
                                          const float* fars, const float* noises, const uint8_t* bitfield, uint32_t N, uint32_t C, uint32_t H,
                                          float dt_gamma, uint32_t max_steps, float T_thresh, float* weights_sum, float* depth, float* image,
                                          float* mask_out, int32_t* work_counter, void* stream) {
    if (int e = field::validate(desc)) return e;
    if (C == 0 || C > 16 || H == 0 || H > 1024 || (H & (H - 1)) || max_steps == 0 || N >= 0x7fffffffu) return INERF_ERR_SIZE;
    if (N == 0) return INERF_OK;
    INERF_REQUIRE(rays_o); INERF_REQUIRE(rays_d); INERF_REQUIRE(nears); INERF_REQUIRE(fars); INERF_REQUIRE(bitfield);
    INERF_REQUIRE(weights_sum); INERF_REQUIRE(depth); INERF_REQUIRE(image); INERF_REQUIRE(work_counter);
    if (((uintptr_t)work_counter & 7u) || ((uintptr_t)bitfield & 7u)) return INERF_ERR_ALIGN;
    if (mask_out && ((uintptr_t)mask_out & 15u)) return INERF_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce = cudaMemsetAsync(work_counter, 0, 4 * sizeof(int32_t), st);
    if (ce != cudaSuccess) return (int)ce;
    // rays without a single sample are never touched by the kernel: their outputs are these zeros
    if ((ce = cudaMemsetAsync(weights_sum, 0, (size_t)N * 4, st)) != cudaSuccess) return (int)ce;
    if ((ce = cudaMemsetAsync(image, 0, (size_t)N * 12, st)) != cudaSuccess) return (int)ce;
    if (mask_out && (ce = cudaMemsetAsync(mask_out, 0, (size_t)N * desc->K * 4, st)) != cudaSuccess) return (int)ce;
    // coarse bitmap: one bit per 64 cells; needs C*H^3 to be a multiple of 2048 and to fit the spare shared memory
    const uint64_t cells = (uint64_t)C * H * H * H;
    uint32_t coarse_bytes = 0;
    if (cells % 2048 == 0 && cells / 512 <= 48 * 1024) coarse_bytes = (uint32_t)(cells / 512);
    RenderParams p{rays_o, rays_d, nears, fars, noises, bitfield, N, C, H, max_steps, dt_gamma, T_thresh, weights_sum, depth, image, mask_out, work_counter, coarse_bytes};
    const uint32_t smem_bytes = RSmem::bytes(desc->K, coarse_bytes);
    // the attribute is per device and cheap to set: no process-global "already done" flag (a second GPU in the same process
    // would otherwise launch without the opt-in)
    if (desc->K <= 32) ce = cudaFuncSetAttribute(k_render_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    else ce = cudaFuncSetAttribute(k_render_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (ce != cudaSuccess) return (int)ce;
    k_first_hit<<<(N + 255u) / 256u, 256, 0, st>>>(p, desc->bound, depth);   // depth[] carries t_first into the render kernel
    INERF_LAUNCH_CHECK();
    const uint32_t want = (N + field::kTile - 1) / field::kTile;
    const uint32_t sms = (uint32_t)device_sm_count();
    const uint32_t grid = want < sms ? want : sms;
    if (desc->K <= 32) k_render_fused<1><<<grid, kThreadsR, smem_bytes, st>>>(*desc, p);
    else k_render_fused<2><<<grid, kThreadsR, smem_bytes, st>>>(*desc, p);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_render_fused(const inerf_field_desc* desc, const float* rays_o, const float* rays_d, const float* nears,
                                  const float* fars, const uint8_t* bitfield, uint32_t N, uint32_t C, uint32_t H, float dt_gamma,
                                  uint32_t max_steps, float T_thresh, float* weights_sum, float* depth, float* image,
                                  float* mask_out, int32_t* work_counter, void* stream) {
    return inerf_render_fused_perturb(desc, rays_o, rays_d, nears, fars, nullptr, bitfield, N, C, H, dt_gamma, max_steps, T_thresh, weights_sum,
                                      depth, image, mask_out, work_counter, stream);
}
