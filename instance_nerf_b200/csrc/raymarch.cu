// Ray generation helpers and occupancy-grid marching for sm_100a.
//
// Replaces raymarching/src/raymarching.cu:91-490 and :958-1073 of the reference
// (near/far, Morton, packbits, march_rays_train, march_rays).  The arithmetic is
// spelled with explicit round-to-nearest intrinsics (__fmaf_rn / __fmul_rn /
// __fadd_rn / __fdiv_rn) in exactly the places where nvcc's default -fmad=true
// contracts the reference's expressions, so the sample stream is bit-identical
// to the reference's and independent of compiler flags.
//
// Unlike the reference (atomicAdd reservation, racy order) training marching is
// deterministic: count -> exclusive scan -> write, rays in id order.
#include "march_device.cuh"

namespace {

using namespace march;

// ---- utils -------------------------------------------------------------------

// slab test of raymarching.cu:91-145; a miss at the y or z stage returns near = far = FLT_MAX
__device__ __forceinline__ void near_far_one(float ox, float oy, float oz, float dx, float dy, float dz, const float* __restrict__ aabb,
                                             float min_near, float& near_out, float& far_out) {
    const float rdx = __fdiv_rn(1.0f, dx), rdy = __fdiv_rn(1.0f, dy), rdz = __fdiv_rn(1.0f, dz);
    const float kMax = 3.402823466e+38f;
    float near = __fmul_rn(__fsub_rn(aabb[0], ox), rdx), far = __fmul_rn(__fsub_rn(aabb[3], ox), rdx), tmp;
    if (near > far) { tmp = near; near = far; far = tmp; }
    float near_y = __fmul_rn(__fsub_rn(aabb[1], oy), rdy), far_y = __fmul_rn(__fsub_rn(aabb[4], oy), rdy);
    if (near_y > far_y) { tmp = near_y; near_y = far_y; far_y = tmp; }
    if (near > far_y || near_y > far) { near_out = kMax; far_out = kMax; return; }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = __fmul_rn(__fsub_rn(aabb[2], oz), rdz), far_z = __fmul_rn(__fsub_rn(aabb[5], oz), rdz);
    if (near_z > far_z) { tmp = near_z; near_z = far_z; far_z = tmp; }
    if (near > far_z || near_z > far) { near_out = kMax; far_out = kMax; return; }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
    near_out = near;
    far_out = far;
}

__global__ void k_near_far_from_aabb(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                     const float* __restrict__ aabb, uint32_t N, float min_near,
                                     float* __restrict__ nears, float* __restrict__ fars) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float near, far;
    near_far_one(rays_o[n * 3], rays_o[n * 3 + 1], rays_o[n * 3 + 2], rays_d[n * 3], rays_d[n * 3 + 1], rays_d[n * 3 + 2], aabb, min_near,
                 near, far);
    nears[n] = near;
    fars[n] = far;
}

// nerf/utils.py:56-140 get_rays: pixel centre (+0.5), camera direction ((i-cx)/fx, (j-cy)/fy, 1) normalised, rotated by the
// pose; rays_o = camera centre.  One thread per (pose, ray); `inds` (int64 [N], shared by every pose: utils.py:104,108)
// selects pixels, NULL = all H*W pixels in row-major order.  The reference builds the full H*W meshgrid and ~10 torch
// kernels even for 4096 rays.  With `aabb` the slab test of near_far_from_aabb runs in the same thread on the same
// fp32 values (bit-identical to calling it afterwards).  Divisions / sqrt are IEEE round-to-nearest, the 3x3 rotation is
// a plain multiply-add chain in k order (torch's matmul may order / contract differently: parity to ~1 ulp, tests/).
__global__ void k_get_rays(const float* __restrict__ poses, float fx, float fy, float cx, float cy, uint32_t W, uint32_t N, uint32_t B,
                           const long long* __restrict__ inds, float* __restrict__ rays_o, float* __restrict__ rays_d,
                           const float* __restrict__ aabb, float min_near, float* __restrict__ nears, float* __restrict__ fars) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (uint64_t)N * B) return;
    const uint32_t b = (uint32_t)(g / N), n = (uint32_t)(g - (uint64_t)b * N);
    const uint32_t pix = inds ? (uint32_t)inds[n] : n;
    const uint32_t row = pix / W, col = pix - row * W;
    const float* P = poses + (size_t)b * 16;
    const float xs = __fdiv_rn(__fsub_rn(__fadd_rn((float)col, 0.5f), cx), fx);
    const float ys = __fdiv_rn(__fsub_rn(__fadd_rn((float)row, 0.5f), cy), fy);
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(xs, xs), __fmul_rn(ys, ys)), 1.0f));
    const float vx = __fdiv_rn(xs, nrm), vy = __fdiv_rn(ys, nrm), vz = __fdiv_rn(1.0f, nrm);
    float dxyz[3], oxyz[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        dxyz[r] = __fadd_rn(__fadd_rn(__fmul_rn(vx, P[r * 4]), __fmul_rn(vy, P[r * 4 + 1])), __fmul_rn(vz, P[r * 4 + 2]));
        oxyz[r] = P[r * 4 + 3];
    }
    float* po = rays_o + g * 3;
    float* pd = rays_d + g * 3;
    po[0] = oxyz[0]; po[1] = oxyz[1]; po[2] = oxyz[2];
    pd[0] = dxyz[0]; pd[1] = dxyz[1]; pd[2] = dxyz[2];
    if (aabb) {
        float near, far;
        near_far_one(oxyz[0], oxyz[1], oxyz[2], dxyz[0], dxyz[1], dxyz[2], aabb, min_near, near, far);
        nears[g] = near;
        fars[g] = far;
    }
}

// raymarching.cu:162-198
__global__ void k_sph_from_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float radius,
                               uint32_t N, float* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float B = ox * dx + oy * dy + oz * dz;
    const float Cc = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-B + sqrtf(B * B - A * Cc)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2f(sqrtf(x * x + z * z), y);
    const float phi = atan2f(z, x);
    const float kRPi = 0.3183098861837907f;
    coords[n * 2] = 2 * theta * kRPi - 1;
    coords[n * 2 + 1] = phi * kRPi;
}

__global__ void k_morton3D(const int32_t* __restrict__ coords, uint32_t N, int32_t* __restrict__ indices) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    indices[n] = (int32_t)morton3D_enc(coords[n * 3], coords[n * 3 + 1], coords[n * 3 + 2]);
}
__global__ void k_morton3D_invert(const int32_t* __restrict__ indices, uint32_t N, int32_t* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int32_t ind = indices[n];
    coords[n * 3 + 0] = (int32_t)morton3D_dec((uint32_t)(ind >> 0));
    coords[n * 3 + 1] = (int32_t)morton3D_dec((uint32_t)(ind >> 1));
    coords[n * 3 + 2] = (int32_t)morton3D_dec((uint32_t)(ind >> 2));
}

// One thread packs 32 cells (4 bytes out, 8 x float4 in): coalesced 128-bit
// loads instead of the reference's 1 byte per thread (raymarching.cu:267-289).
__global__ void k_packbits(const float* __restrict__ grid, uint32_t N, float thresh, uint8_t* __restrict__ bitfield) {
    const uint32_t nwords = N >> 2;
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
    // tail bytes (N not a multiple of 4)
    if (blockIdx.x == 0 && threadIdx.x < (N & 3u)) {
        const uint32_t n = (N & ~3u) + threadIdx.x;
        uint8_t bits = 0;
        for (int i = 0; i < 8; i++) bits |= (grid[(size_t)n * 8 + i] > thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

// ---- training marching ---------------------------------------------------------

__global__ void __launch_bounds__(128) k_march_count(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                     const uint8_t* __restrict__ grid, float bound, float dt_gamma,
                                                     uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                                     const float* __restrict__ nears, const float* __restrict__ fars,
                                                     const float* __restrict__ noises, int32_t* __restrict__ rays,
                                                     float* __restrict__ t_scratch) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    Walk w;
    w.init(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, grid, bound, dt_gamma, max_steps, C, H, fars[n]);
    float t0 = nears[n];
    t0 = __fmaf_rn(w.step_size(t0), noises[n], t0);  // :351
    rays[n * 3 + 2] = (int32_t)w.run<false>(t0, max_steps, nullptr, nullptr, nullptr, t_scratch + (size_t)n * max_steps);
}

// Second phase when the count pass recorded every sample's t (t_scratch [N, max_steps]): no second walk.  One warp per ray,
// lane j rebuilds sample j from t_j alone -- position, dt = step(t_j), and delta1 = (t_j + dt_j) - (t_{j-1} + dt_{j-1}), the
// very expressions of the walk (raymarching.cu:373-389) -- and the 32 B / sample leave as coalesced stores.
__global__ void __launch_bounds__(256) k_march_expand(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float bound,
                                                      float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                                      const float* __restrict__ nears, const float* __restrict__ noises,
                                                      const int32_t* __restrict__ rays, const float* __restrict__ t_scratch,
                                                      float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (n >= N) return;
    const uint32_t offset = (uint32_t)rays[n * 3 + 1], num = (uint32_t)rays[n * 3 + 2];
    if (num == 0 || offset + num > M) return;  // :415-416
    Walk w;
    w.init(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, nullptr, bound, dt_gamma, max_steps, C, H, 0.f);
    float last = nears[n];
    last = __fmaf_rn(w.step_size(last), noises[n], last);   // last_t starts at the (jittered) start parameter
    const float* ts = t_scratch + (size_t)n * max_steps;
    for (uint32_t base = 0; base < num; base += 32) {
        const uint32_t j = base + lane;
        const bool valid = j < num;
        const float t = valid ? ts[j] : 0.f;
        const float dt = w.step_size(t);
        const float tn = __fadd_rn(t, dt);
        const float up = __shfl_up_sync(0xffffffffu, tn, 1);
        const float lt = lane == 0 ? last : up;
        if (valid) {
            const size_t s = (size_t)offset + j;
            xyzs[s * 3] = clampf(__fmaf_rn(t, w.dx, w.ox), -bound, bound);
            xyzs[s * 3 + 1] = clampf(__fmaf_rn(t, w.dy, w.oy), -bound, bound);
            xyzs[s * 3 + 2] = clampf(__fmaf_rn(t, w.dz, w.oz), -bound, bound);
            dirs[s * 3] = w.dx; dirs[s * 3 + 1] = w.dy; dirs[s * 3 + 2] = w.dz;
            reinterpret_cast<float2*>(deltas)[s] = make_float2(dt, __fsub_rn(tn, lt));
        }
        last = __shfl_sync(0xffffffffu, tn, min(31u, num - base - 1u));
    }
}

// Single-CTA exclusive scan of the counts (N is at most a few million rays; the
// scan reads 4 B and writes 8 B per ray, microseconds next to the marching).
__global__ void __launch_bounds__(1024) k_march_scan(int32_t* __restrict__ rays, uint32_t N, int32_t* __restrict__ counter) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = (uint32_t)counter[0];
    __syncthreads();
    for (uint32_t base = 0; base < N; base += 1024) {
        const uint32_t n = base + threadIdx.x;
        const uint32_t v = n < N ? (uint32_t)rays[n * 3 + 2] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += up;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= (uint32_t)o) wi += up;
            }
            warp_sums[lane] = wi - ws;  // exclusive warp offsets
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t excl = carry + warp_sums[warp] + incl - v;
        if (n < N) { rays[n * 3] = (int32_t)n; rays[n * 3 + 1] = (int32_t)excl; }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;  // last thread holds the chunk's inclusive total
        __syncthreads();
    }
    if (threadIdx.x == 0) { counter[0] = (int32_t)carry_s; counter[1] += (int32_t)N; }
}

// ---- training marching, one WARP per ray (small batches) -------------------------------------------------------------
// With a few thousand rays the thread-per-ray kernels above run one serial chain of ~4*10^4 dependent instructions per
// ray on a fraction of a warp slot per SM (c3: 4096 rays = 128 warps on 148 SMs, 120-150 us per pass).  Here the 32
// lanes of a warp share one ray.  The reference's walk visits a SUBSEQUENCE of the fixed step sequence
// t_{n+1} = t_n + clamp(t_n * dt_gamma, dt_min, dt_max) (both the sample step and the empty-space skip use it,
// raymarching.cu:373-398): a sample is taken at an evaluation point whose cell is occupied, an empty cell moves the
// evaluation point to the first t_m >= its exit parameter.  So every lane takes one candidate t_{n+lane}, evaluates its
// cell (occupancy, exit parameter) in parallel, and the warp then replays the reference's pointer chase over the 32
// results with ballots.  Same samples bit for bit (tests/test_ops_vs_ref_gpu.py); writes are coalesced.
__global__ void __launch_bounds__(256) k_march_train_warp(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                          const uint8_t* __restrict__ grid, float bound, float dt_gamma, uint32_t max_steps,
                                                          uint32_t N, uint32_t C, uint32_t H, const float* __restrict__ nears,
                                                          const float* __restrict__ fars, const float* __restrict__ noises,
                                                          int32_t* __restrict__ rays, float* __restrict__ t_scratch) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (n >= N) return;
    float* ts = t_scratch + (size_t)n * max_steps;   // the parameter t of every sample, for k_march_expand
    const uint32_t limit = max_steps;
    Walk w;
    w.init(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, grid, bound, dt_gamma, max_steps, C, H, fars[n]);
    float t = nears[n];
    t = __fmaf_rn(w.step_size(t), noises[n], t);  // :351
    for (uint32_t j = 0; j < lane; j++) t = __fadd_rn(t, w.step_size(t));   // lane j holds candidate t_{n0 + j}
    uint32_t count = 0;
    float pending = -1.0f;                          // exit parameter of an empty cell whose target lies in a later batch
    bool have_pending = false;
    const uint32_t lt_mask = (1u << lane) - 1u;
    while (true) {
        // ---- every lane: its cell ----
        float tn = t, x, y, z, dt, tt = 0.f;
        const bool in_range = t < w.far;
        bool occ = false;
        if (in_range) occ = w.eval_cell(tn, x, y, z, dt, tt);   // occupied: tn = t + dt
        const uint32_t occ_b = __ballot_sync(0xffffffffu, occ);
        const uint32_t in_b = __ballot_sync(0xffffffffu, in_range);
        // ---- replay the reference's walk over the 32 candidates ----
        uint32_t cur = 0, emit = 0;
        bool done = false;
        if (have_pending) {
            const uint32_t ge = __ballot_sync(0xffffffffu, !(t < pending));
            if (ge) { cur = __ffs(ge) - 1u; have_pending = false; } else cur = 32u;
        }
        while (cur < 32u) {
            if (!((in_b >> cur) & 1u) || count >= limit) { done = true; break; }
            if ((occ_b >> cur) & 1u) {
                emit |= 1u << cur;
                count++;
                cur++;
            } else {
                const float ttc = __shfl_sync(0xffffffffu, tt, cur);
                const uint32_t ge = __ballot_sync(0xffffffffu, !(t < ttc)) & ~((2u << cur) - 1u);   // lanes above cur
                if (ge) cur = __ffs(ge) - 1u;
                else { pending = ttc; have_pending = true; cur = 32u; }
            }
        }
        if ((emit >> lane) & 1u) ts[(count - __popc(emit)) + __popc(emit & lt_mask)] = t;
        if (done || !(in_b >> 31)) break;   // the walk ended, or candidates past this batch are all beyond `far`
        // ---- next batch: every lane advances 32 steps ----
#pragma unroll 4
        for (int j = 0; j < 32; j++) t = __fadd_rn(t, w.step_size(t));
    }
    if (lane == 0) rays[n * 3 + 2] = (int32_t)count;
}

// ---- inference marching (raymarching.cu:958-1063) -------------------------------

__global__ void __launch_bounds__(128) k_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* __restrict__ rays_alive,
                                                    const float* __restrict__ rays_t, const float* __restrict__ rays_o,
                                                    const float* __restrict__ rays_d, float bound, float dt_gamma,
                                                    uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* __restrict__ grid,
                                                    const float* __restrict__ fars, float* __restrict__ xyzs,
                                                    float* __restrict__ dirs, float* __restrict__ deltas,
                                                    const float* __restrict__ noises) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int32_t index = rays_alive[n];
    Walk w;
    w.init(rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, grid, bound, dt_gamma, max_steps, C, H, fars[index]);
    float t = rays_t[index];
    t = __fmaf_rn(w.step_size(t), noises[n], t);  // :1004
    const size_t s = (size_t)n * n_step;
    w.run<true>(t, n_step, xyzs + s * 3, dirs + s * 3, deltas + s * 2);
}

// Stable compaction of rays_alive >= 0 (replaces the boolean-index + sync at mask_renderer.py:370).
__global__ void __launch_bounds__(1024) k_compact_alive(const int32_t* __restrict__ in, uint32_t n, int32_t* __restrict__ out,
                                                        int32_t* __restrict__ n_out) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const int32_t v = i < n ? in[i] : -1;
        const bool keep = v >= 0;
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        const uint32_t prefix = __popc(ballot & ((1u << lane) - 1u));
        if (lane == 0) warp_sums[warp] = __popc(ballot);
        __syncthreads();
        if (warp == 0) {
            const uint32_t ws = warp_sums[lane];
            uint32_t wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= (uint32_t)o) wi += up;
            }
            warp_sums[lane] = wi - ws;
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t pos = carry + warp_sums[warp] + prefix;
        if (keep) out[pos] = v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = pos + (keep ? 1u : 0u);
        __syncthreads();
    }
    if (threadIdx.x == 0) n_out[0] = (int32_t)carry_s;
}

}  // namespace

// ---- C ABI ---------------------------------------------------------------------

extern "C" int inerf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                                        float min_near, float* nears, float* fars, void* stream) {
    if (N == 0) return INERF_OK;
    INERF_REQUIRE(rays_o); INERF_REQUIRE(rays_d); INERF_REQUIRE(aabb); INERF_REQUIRE(nears); INERF_REQUIRE(fars);
    k_near_far_from_aabb<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_get_rays(const float* poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W,
                              const long long* inds, uint32_t N, float* rays_o, float* rays_d, const float* aabb, float min_near,
                              float* nears, float* fars, void* stream) {
    if (H == 0 || W == 0 || (uint64_t)H * W > 0x7fffffffull || fx == 0.f || fy == 0.f) return INERF_ERR_SIZE;
    if (inds == nullptr && N != H * W) return INERF_ERR_SIZE;
    if ((uint64_t)N * B == 0) return INERF_OK;
    INERF_REQUIRE(poses); INERF_REQUIRE(rays_o); INERF_REQUIRE(rays_d);
    if (aabb) { INERF_REQUIRE(nears); INERF_REQUIRE(fars); }
    k_get_rays<<<div_up((uint64_t)N * B, 256), 256, 0, (cudaStream_t)stream>>>(poses, fx, fy, cx, cy, W, N, B, inds, rays_o, rays_d, aabb,
                                                                              min_near, nears, fars);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords, void* stream) {
    if (N == 0) return INERF_OK;
    INERF_REQUIRE(rays_o); INERF_REQUIRE(rays_d); INERF_REQUIRE(coords);
    k_sph_from_ray<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, radius, N, coords);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream) {
    if (N == 0) return INERF_OK;
    INERF_REQUIRE(coords); INERF_REQUIRE(indices);
    k_morton3D<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(coords, N, indices);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream) {
    if (N == 0) return INERF_OK;
    INERF_REQUIRE(indices); INERF_REQUIRE(coords);
    k_morton3D_invert<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(indices, N, coords);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, void* stream) {
    if (N == 0) return INERF_OK;
    INERF_REQUIRE(grid); INERF_REQUIRE(bitfield);
    if (((uintptr_t)grid & 15u) || ((uintptr_t)bitfield & 3u)) return INERF_ERR_ALIGN;
    const unsigned int blocks = min(div_up(max(N >> 2, 1u), 256), (unsigned int)(kNumSMs * 8));
    k_packbits<<<blocks, 256, 0, (cudaStream_t)stream>>>(grid, N, density_thresh, bitfield);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

// One warp per ray pays off while thread-per-ray cannot fill the machine (a serial walk per thread); it evaluates every
// step of the empty stretches instead of one per cell, so the thread-per-ray kernel wins once there are enough rays.
// The warp kernel reads the bitfield as 8-byte words (Walk::occupied_cached).
constexpr uint32_t kWarpMarchMaxRays = 24576u;   // crossover measured on B200
static bool use_warp_march(uint32_t N, const uint8_t* grid, uint32_t C, uint32_t H) {
    return N <= kWarpMarchMaxRays && ((uintptr_t)grid & 7u) == 0 && ((uint64_t)C * H * H * H) % 64u == 0;
}

static int check_march_args(uint32_t C, uint32_t H, uint32_t max_steps) {
    if (C == 0 || C > 16 || H == 0 || H > 1024 || max_steps == 0) return INERF_ERR_SIZE;
    if (H & (H - 1)) return INERF_ERR_UNSUPPORTED;  // the float rewrite of the double sub-expression needs H = 2^k
    return INERF_OK;
}

extern "C" size_t inerf_march_scratch_floats(uint32_t N, uint32_t max_steps) { return (size_t)N * max_steps; }

extern "C" int inerf_march_rays_train_count_t(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                              float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                              const float* nears, const float* fars, int32_t* rays, int32_t* counter,
                                              const float* noises, float* t_scratch, void* stream) {
    if (int e = check_march_args(C, H, max_steps)) return e;
    INERF_REQUIRE(counter);
    if (N == 0) return INERF_OK;
    INERF_REQUIRE(rays_o); INERF_REQUIRE(rays_d); INERF_REQUIRE(grid); INERF_REQUIRE(nears); INERF_REQUIRE(fars);
    INERF_REQUIRE(rays); INERF_REQUIRE(noises); INERF_REQUIRE(t_scratch);
    if (use_warp_march(N, grid, C, H))
        k_march_train_warp<<<div_up(N, 8), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars,
                                                                         noises, rays, t_scratch);
    else
        k_march_count<<<div_up(N, 128), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H,
                                                                      nears, fars, noises, rays, t_scratch);
    INERF_LAUNCH_CHECK();
    k_march_scan<<<1, 1024, 0, (cudaStream_t)stream>>>(rays, N, counter);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_march_rays_train_expand(const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                                             uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears, float* xyzs, float* dirs,
                                             float* deltas, const int32_t* rays, const float* noises, const float* t_scratch,
                                             void* stream) {
    if (int e = check_march_args(C, H, max_steps)) return e;
    if (N == 0 || M == 0) return INERF_OK;
    INERF_REQUIRE(rays_o); INERF_REQUIRE(rays_d); INERF_REQUIRE(nears); INERF_REQUIRE(rays); INERF_REQUIRE(noises);
    INERF_REQUIRE(xyzs); INERF_REQUIRE(dirs); INERF_REQUIRE(deltas); INERF_REQUIRE(t_scratch);
    if ((uintptr_t)deltas & 7u) return INERF_ERR_ALIGN;
    k_march_expand<<<div_up(N, 8), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, bound, dt_gamma, max_steps, N, C, H, M, nears, noises, rays,
                                                                 t_scratch, xyzs, dirs, deltas);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                                      uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                                      const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays,
                                      int32_t* counter, const float* noises, float* t_scratch, void* stream) {
    if (int e = inerf_march_rays_train_count_t(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, rays,
                                               counter, noises, t_scratch, stream)) return e;
    return inerf_march_rays_train_expand(rays_o, rays_d, bound, dt_gamma, max_steps, N, C, H, M, nears, xyzs, dirs, deltas, rays, noises,
                                         t_scratch, stream);
}

extern "C" int inerf_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                                const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                                uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars, float* xyzs,
                                float* dirs, float* deltas, const float* noises, void* stream) {
    (void)nears;  // passed but unused by the reference kernel as well (raymarching.cu:995)
    if (int e = check_march_args(C, H, max_steps)) return e;
    if (n_alive == 0 || n_step == 0) return INERF_OK;
    INERF_REQUIRE(rays_alive); INERF_REQUIRE(rays_t); INERF_REQUIRE(rays_o); INERF_REQUIRE(rays_d); INERF_REQUIRE(grid);
    INERF_REQUIRE(fars); INERF_REQUIRE(xyzs); INERF_REQUIRE(dirs); INERF_REQUIRE(deltas); INERF_REQUIRE(noises);
    k_march_rays<<<div_up(n_alive, 128), 128, 0, (cudaStream_t)stream>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound,
                                                                       dt_gamma, max_steps, C, H, grid, fars, xyzs, dirs, deltas, noises);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_compact_alive(const int32_t* rays_alive, uint32_t n_alive, int32_t* out, int32_t* n_out, void* stream) {
    INERF_REQUIRE(n_out);
    if (n_alive) { INERF_REQUIRE(rays_alive); INERF_REQUIRE(out); }
    if (rays_alive == out && n_alive) return INERF_ERR_UNSUPPORTED;
    k_compact_alive<<<1, 1024, 0, (cudaStream_t)stream>>>(rays_alive, n_alive, out, n_out);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}
