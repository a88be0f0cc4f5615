// Alpha compositing of colour, depth and per-sample instance logits (sm_100a).
//
// Replaces raymarching/src/raymarching.cu:500-951 (training forward/backward,
// with and without masks) and :1076-1280 (inference) of the reference.
//
// Mapping: ONE WARP PER RAY instead of the reference's one thread per ray.
//   * the sample stream (sigma, deltas, rgb) is read 32 samples at a time with
//     coalesced loads, alpha is evaluated for 32 samples in parallel;
//   * the K running sums live in registers, where the reference read-modify-writes
//     mask_out / grad_masks_acc in global memory 2K times per sample
//     (raymarching.cu:766-768, 894-896);
//   * training kernels: the transmittance recurrence, the depth prefix and the
//     "still to come" terms of grad_sigma are warp scans over 32-sample chunks
//     (see the block comment below); the early-stop decision `T < T_thresh` is one
//     ballot and matches the reference sample for sample.
#include "common.cuh"

namespace {

constexpr int kMaxKPL = 4;  // classes per lane -> K <= 128
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Rows of the [M, K] logit stream in flight per warp: the recurrence (phase A) never waits on them, so phase B issues
// kRowsInFlight independent 128-byte row loads back to back (64 warps/SM x 8 x 128 B = 64 KB in flight per SM).
constexpr int kRowsInFlight = 8;

// ---- training backward for K > 64 (K == 0: raymarching.cu:601-682; K > 0: :828-940) ----
// Many-class path only (the scan kernel below keeps a ray's K logit gradients in registers, which stops at K = 64): lane k
// owns class k (+32, ...).  Phase A replays the recurrence in the reference's order (every lane redundantly) and leaves in
// lane j: w_j, T_j (after sample j) and the colour / weights_sum part of grad_sigma_j.  Phase B streams the logit rows:
// macc, grad_masks = g_m * w_j, and the K-term sum of grad_sigma_j (one warp reduction per row).
template <int KPL>
__global__ void __launch_bounds__(256) k_composite_train_bwd(
    const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_image, const float* __restrict__ grad_mask_out,
    const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ masks,
    const float* __restrict__ deltas, const int32_t* __restrict__ rays, const float* __restrict__ weights_sum,
    const float* __restrict__ image, const float* __restrict__ mask_out, uint32_t M, uint32_t N, uint32_t K, float T_thresh,
    float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs, float* __restrict__ grad_masks) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) return;

    const float gr = grad_image[(size_t)index * 3], gg = grad_image[(size_t)index * 3 + 1], gb = grad_image[(size_t)index * 3 + 2];
    const float r_final = image[(size_t)index * 3], g_final = image[(size_t)index * 3 + 1], b_final = image[(size_t)index * 3 + 2];
    const float ws_term = grad_weights_sum[index] * (1 - weights_sum[index]);
    float gm[KPL > 0 ? KPL : 1], mfin[KPL > 0 ? KPL : 1], macc[KPL > 0 ? KPL : 1];
#pragma unroll
    for (int i = 0; i < (KPL > 0 ? KPL : 1); i++) {
        const uint32_t k = lane + 32u * i;
        const bool ok = KPL > 0 && k < K;
        gm[i] = ok ? grad_mask_out[(size_t)index * K + k] : 0.f;
        mfin[i] = ok ? mask_out[(size_t)index * K + k] : 0.f;
        macc[i] = 0.f;
    }
    float T = 1.0f, r = 0, g = 0, b = 0;

    for (uint32_t base = 0; base < num_steps; base += 32) {
        const uint32_t s = offset + base + lane;
        const bool valid = base + lane < num_steps;
        float alpha = 0.f, d0 = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
        if (valid) {
            d0 = __ldg(deltas + (size_t)s * 2);
            alpha = 1.0f - __expf(-__ldg(sigmas + s) * d0);
            cr = __ldg(rgbs + (size_t)s * 3); cg = __ldg(rgbs + (size_t)s * 3 + 1); cb = __ldg(rgbs + (size_t)s * 3 + 2);
        }
        float my_w = 0.f, my_T = 0.f, my_gs = 0.f;  // lane j keeps the results of sample base + j
        const uint32_t cnt = min(32u, num_steps - base);
        uint32_t m = cnt;
        bool done = false;
        for (uint32_t j = 0; j < cnt; j++) {
            const float a = __shfl_sync(kFull, alpha, j);
            const float weight = a * T;
            const float rj = __shfl_sync(kFull, cr, j), gj = __shfl_sync(kFull, cg, j), bj = __shfl_sync(kFull, cb, j);
            r = fmaf(weight, rj, r);
            g = fmaf(weight, gj, g);
            b = fmaf(weight, bj, b);
            T *= 1.0f - a;
            const float gs = gr * (T * rj - (r_final - r)) + gg * (T * gj - (g_final - g)) + gb * (T * bj - (b_final - b)) + ws_term;
            if (lane == j) { my_w = weight; my_T = T; my_gs = gs; }
            if (T < T_thresh) { done = true; m = j + 1; break; }
        }
        if (KPL > 0) {
            const size_t row0 = (size_t)(offset + base) * K;
            float my_part = 0.f;
            for (uint32_t j0 = 0; j0 < m; j0 += kRowsInFlight) {
                float v[kRowsInFlight][KPL > 0 ? KPL : 1];
#pragma unroll
                for (int u = 0; u < kRowsInFlight; u++)
#pragma unroll
                    for (int i = 0; i < KPL; i++) {
                        const uint32_t k = lane + 32u * i;
                        v[u][i] = (j0 + u < m && k < K) ? __ldg(masks + row0 + (size_t)(j0 + u) * K + k) : 0.f;
                    }
#pragma unroll
                for (int u = 0; u < kRowsInFlight; u++) {
                    if (j0 + u < m) {   // warp-uniform
                        const float weight = __shfl_sync(kFull, my_w, j0 + u), Tj = __shfl_sync(kFull, my_T, j0 + u);
                        float part = 0.f;
#pragma unroll
                        for (int i = 0; i < KPL; i++) {
                            const uint32_t k = lane + 32u * i;
                            if (k < K) {
                                macc[i] = fmaf(weight, v[u][i], macc[i]);
                                grad_masks[row0 + (size_t)(j0 + u) * K + k] = gm[i] * weight;
                                part += gm[i] * (Tj * v[u][i] - (mfin[i] - macc[i]));
                            }
                        }
                        part = warp_sum(part);
                        if (lane == j0 + u) my_part = part;
                    }
                }
            }
            my_gs += my_part;
        }
        if (lane < m) {
            grad_sigmas[s] = d0 * my_gs;
            grad_rgbs[(size_t)s * 3] = gr * my_w; grad_rgbs[(size_t)s * 3 + 1] = gg * my_w; grad_rgbs[(size_t)s * 3 + 2] = gb * my_w;
        }
        if (done) break;
    }
}

// ---- training forward / backward, scan formulation -------------------------------------------------------------------------
// Replaying the reference's per-sample recurrence in every lane and broadcasting each sample costs 5-11 SHFL per sample, and
// the shuffle unit retires one warp instruction per clock per SM -- that, not HBM, bounded the first version of these kernels
// (forward 0.45, backward 0.38 of HBM peak at K = 32).  Here lane j of a 32-sample chunk owns sample j:
//   T_j      = T_carry * prod_{i<=j}(1 - alpha_i)          one inclusive product scan (5 SHFL per CHUNK)
//   stop     = first j with T_j < T_thresh                  one ballot; samples after it get weight 0 (raymarching.cu:573)
//   w_j      = alpha_j * T_{j-1},  t_j = t_carry + sum_{i<=j} delta1_i   (one add scan)
//   sums     per-lane partials of w, w*rgb, w*t across the chunks of a ray, ONE warp reduction per ray
// and in the backward the two "what is still to come" terms of grad_sigma become scans as well:
//   sum_c g_c (T_j c_jc - (C_c - Crun_jc)) = T_j q_j - (Q - Qrun_j),   q_j = g . rgb_j,        Qrun = scan(w q)
//   sum_k g_k (T_j m_jk - (M_k - Mrun_jk)) = T_j p_j - (G - Prun_j),   p_j = g_mask . logits_j, Prun = scan(w p)
// with p_j formed by the lane that owns row j from its own 4K-byte row (no per-row warp reduction, no broadcast).
// Same mathematics as the reference; sums are re-associated (differences ~1e-7 relative, tests/test_ops_vs_ref_gpu.py).
__device__ __forceinline__ float scan_mul(float v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_up_sync(kFull, v, o); if (lane >= (uint32_t)o) v *= u; }
    return v;
}
__device__ __forceinline__ float scan_add(float v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_up_sync(kFull, v, o); if (lane >= (uint32_t)o) v += u; }
    return v;
}

template <int KPL>
__global__ void __launch_bounds__(256) k_composite_train_fwd_scan(
    const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ masks,
    const float* __restrict__ deltas, const int32_t* __restrict__ rays, uint32_t M, uint32_t N, uint32_t K,
    float T_thresh, float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image,
    float* __restrict__ mask_out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
    float macc[KPL > 0 ? KPL : 1];
#pragma unroll
    for (int i = 0; i < (KPL > 0 ? KPL : 1); i++) macc[i] = 0.f;
    float pr = 0.f, pg = 0.f, pb = 0.f, pd = 0.f, pws = 0.f;   // per-lane partial sums
    float Tc = 1.0f, tc = 0.f;
    if (num_steps != 0 && offset + num_steps <= M) {
        for (uint32_t base = 0; base < num_steps; base += 32) {
            const uint32_t s = offset + base + lane;
            const bool valid = base + lane < num_steps;
            float alpha = 0.f, d1 = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
            if (valid) {
                const float2 dl = __ldg(reinterpret_cast<const float2*>(deltas) + s);
                alpha = 1.0f - __expf(-__ldg(sigmas + s) * dl.x);
                d1 = dl.y;
                cr = __ldg(rgbs + (size_t)s * 3); cg = __ldg(rgbs + (size_t)s * 3 + 1); cb = __ldg(rgbs + (size_t)s * 3 + 2);
            }
            const float P = scan_mul(1.0f - alpha, lane);
            const float Pex = __shfl_up_sync(kFull, P, 1);
            const float T_after = Tc * P, T_before = lane == 0 ? Tc : Tc * Pex;
            const uint32_t cnt = min(32u, num_steps - base);
            const uint32_t term = __ballot_sync(kFull, valid && T_after < T_thresh);
            const uint32_t m = term ? (uint32_t)__ffs(term) : cnt;      // samples of this chunk that contribute
            const float my_w = lane < m ? alpha * T_before : 0.f;
            const float t = tc + scan_add(d1, lane);
            pws += my_w;
            pr = fmaf(my_w, cr, pr); pg = fmaf(my_w, cg, pg); pb = fmaf(my_w, cb, pb);
            pd = fmaf(my_w, t, pd);
            if (KPL > 0) {
                const float* mbase = masks + (size_t)(offset + base) * K;
                for (uint32_t j0 = 0; j0 < m; j0 += kRowsInFlight) {
                    float v[kRowsInFlight][KPL > 0 ? KPL : 1];
#pragma unroll
                    for (int u = 0; u < kRowsInFlight; u++)
#pragma unroll
                        for (int i = 0; i < KPL; i++) {
                            const uint32_t k = lane + 32u * i;
                            v[u][i] = (j0 + u < m && k < K) ? __ldg(mbase + (size_t)(j0 + u) * K + k) : 0.f;
                        }
#pragma unroll
                    for (int u = 0; u < kRowsInFlight; u++) {
                        if (j0 + u < m) {   // warp-uniform
                            const float weight = __shfl_sync(kFull, my_w, j0 + u);
#pragma unroll
                            for (int i = 0; i < KPL; i++) macc[i] = fmaf(weight, v[u][i], macc[i]);
                        }
                    }
                }
            }
            if (term) break;
            Tc = __shfl_sync(kFull, T_after, 31);
            tc = __shfl_sync(kFull, t, 31);
        }
    }
    const float ws = warp_sum(pws), d = warp_sum(pd), r = warp_sum(pr), g = warp_sum(pg), b = warp_sum(pb);
    if (lane == 0) {
        weights_sum[index] = ws;
        depth[index] = d;
        image[(size_t)index * 3] = r; image[(size_t)index * 3 + 1] = g; image[(size_t)index * 3 + 2] = b;
    }
    if (KPL > 0) {
#pragma unroll
        for (int i = 0; i < KPL; i++) {
            const uint32_t k = lane + 32u * i;
            if (k < K) mask_out[(size_t)index * K + k] = macc[i];
        }
    }
}

// Logit rows [32 samples x K] of a chunk are one CONTIGUOUS block of the [M, K] stream, but the lane that owns sample j needs
// row j: read straight from global memory, every 16-byte load of the warp touches 32 different 128-byte lines (32 L1
// wavefronts for 512 useful bytes) and the gradient rows go back the same way -- the L1 wavefront rate, not HBM, bounded the
// first version of this kernel (0.43 of HBM peak at K = 32).  The block is therefore staged through shared memory: coalesced
// 512-byte warp loads -> rotated [row][chunk] layout (chunk index rotated by the row, so both the coalesced side and the
// row-owner side are bank-conflict free for K = 32) -> lane j reads its row, overwrites it with the gradient row -> coalesced
// warp stores.  Only the rows that contribute (j < m) move.
template <int KPL> struct BwdCfg { static constexpr int kWarps = KPL == 2 ? 4 : 8; };   // 32 KB of staging per CTA either way

template <int KPL>   // K <= 32 * KPL, KPL <= 2: every lane keeps the ray's K logit gradients in registers
__global__ void __launch_bounds__(32 * BwdCfg<KPL>::kWarps) k_composite_train_bwd_scan(
    const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_image, const float* __restrict__ grad_mask_out,
    const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ masks,
    const float* __restrict__ deltas, const int32_t* __restrict__ rays, const float* __restrict__ weights_sum,
    const float* __restrict__ image, const float* __restrict__ mask_out, uint32_t M, uint32_t N, uint32_t K, float T_thresh,
    float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs, float* __restrict__ grad_masks, bool zero_tail) {
    constexpr int KMAX = KPL > 0 ? 32 * KPL : 1;
    __shared__ float4 stage_all[KPL > 0 ? BwdCfg<KPL>::kWarps * 32 * (KMAX / 4) : 1];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;   // whole warps leave together; only __syncwarp is used below
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) return;
    float4* stage = stage_all + (KPL > 0 ? (threadIdx.x >> 5) * 32 * (KMAX / 4) : 0);

    const float gr = grad_image[(size_t)index * 3], gg = grad_image[(size_t)index * 3 + 1], gb = grad_image[(size_t)index * 3 + 2];
    const float Qfin = gr * image[(size_t)index * 3] + gg * image[(size_t)index * 3 + 1] + gb * image[(size_t)index * 3 + 2];
    const float ws_term = grad_weights_sum[index] * (1 - weights_sum[index]);
    float gm[KMAX];
    float Gfin = 0.f;
    if (KPL > 0) {
#pragma unroll
        for (int k = 0; k < KMAX; k++) {
            gm[k] = (uint32_t)k < K ? __ldg(grad_mask_out + (size_t)index * K + k) : 0.f;
            Gfin = fmaf(gm[k], (uint32_t)k < K ? __ldg(mask_out + (size_t)index * K + k) : 0.f, Gfin);
        }
    }
    const bool vec = KPL > 0 && (K & 3u) == 0 && (((uintptr_t)masks | (uintptr_t)grad_masks) & 15u) == 0;
    const uint32_t C4 = K >> 2;   // 16-byte chunks per row (vec path)
    const bool pow2 = (C4 & (C4 - 1u)) == 0u;            // K = 16 / 32 / 64: shifts and masks instead of divisions
    const uint32_t sh = 31u - (uint32_t)__clz((int)(C4 | 1u));
    // staging slot of chunk c of row r: the chunk index is rotated by the row
    auto slot = [&](uint32_t r, uint32_t c) { return r * C4 + (pow2 ? ((c + r) & (C4 - 1u)) : ((c + r) % C4)); };
    float Tc = 1.0f, Qc = 0.f, Pc = 0.f;

    for (uint32_t base = 0; base < num_steps; base += 32) {
        const uint32_t s = offset + base + lane;
        const bool valid = base + lane < num_steps;
        float alpha = 0.f, d0 = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
        if (valid) {
            d0 = __ldg(deltas + (size_t)s * 2);
            alpha = 1.0f - __expf(-__ldg(sigmas + s) * d0);
            cr = __ldg(rgbs + (size_t)s * 3); cg = __ldg(rgbs + (size_t)s * 3 + 1); cb = __ldg(rgbs + (size_t)s * 3 + 2);
        }
        const float P = scan_mul(1.0f - alpha, lane);
        const float Pex = __shfl_up_sync(kFull, P, 1);
        const float T_after = Tc * P, T_before = lane == 0 ? Tc : Tc * Pex;
        const uint32_t cnt = min(32u, num_steps - base);
        const uint32_t term = __ballot_sync(kFull, valid && T_after < T_thresh);
        const uint32_t m = term ? (uint32_t)__ffs(term) : cnt;
        const bool inc = lane < m;
        const float w = inc ? alpha * T_before : 0.f;
        const float q = gr * cr + gg * cg + gb * cb;
        const float Qrun = Qc + scan_add(w * q, lane);
        float gs = T_after * q - (Qfin - Qrun) + ws_term;
        if (KPL > 0) {
            float p = 0.f;
            if (vec) {
                const float4* src = reinterpret_cast<const float4*>(masks + (size_t)(offset + base) * K);
                float4* dst = reinterpret_cast<float4*>(grad_masks + (size_t)(offset + base) * K);
                const uint32_t n16 = m * C4;   // contiguous 16-byte chunks of the rows that contribute
                for (uint32_t i = lane; i < n16; i += 32) {
                    const uint32_t r = pow2 ? i >> sh : i / C4, c = i - r * C4;
                    stage[slot(r, c)] = __ldg(src + i);
                }
                __syncwarp();
                if (inc) {
#pragma unroll
                    for (int k4 = 0; k4 < KMAX / 4; k4++) {
                        if ((uint32_t)k4 < C4) {
                            const uint32_t at = slot(lane, (uint32_t)k4);
                            const float4 v = stage[at];
                            p = fmaf(gm[4 * k4], v.x, p); p = fmaf(gm[4 * k4 + 1], v.y, p);
                            p = fmaf(gm[4 * k4 + 2], v.z, p); p = fmaf(gm[4 * k4 + 3], v.w, p);
                            stage[at] = make_float4(gm[4 * k4] * w, gm[4 * k4 + 1] * w, gm[4 * k4 + 2] * w, gm[4 * k4 + 3] * w);
                        }
                    }
                }
                __syncwarp();
                for (uint32_t i = lane; i < n16; i += 32) {
                    const uint32_t r = pow2 ? i >> sh : i / C4, c = i - r * C4;
                    dst[i] = stage[slot(r, c)];
                }
                __syncwarp();   // the staging block is reused by the next chunk
            } else if (inc) {
                const float* mrow = masks + (size_t)s * K;
                float* grow = grad_masks + (size_t)s * K;
#pragma unroll
                for (int k = 0; k < KMAX; k++) {
                    if ((uint32_t)k < K) {
                        p = fmaf(gm[k], __ldg(mrow + k), p);
                        grow[k] = gm[k] * w;
                    }
                }
            }
            const float Prun = Pc + scan_add(w * p, lane);
            gs += T_after * p - (Gfin - Prun);
            Pc = __shfl_sync(kFull, Prun, 31);
        }
        if (inc) {
            if (grad_sigmas) grad_sigmas[s] = d0 * gs;
            if (grad_rgbs) { grad_rgbs[(size_t)s * 3] = gr * w; grad_rgbs[(size_t)s * 3 + 1] = gg * w; grad_rgbs[(size_t)s * 3 + 2] = gb * w; }
        }
        if (term) {
            // Samples behind the one that terminated the ray have zero gradients (raymarching.cu:905-907 leaves them as the
            // caller's zero fill).  With zero_tail the kernel writes those zeros itself, so the caller's buffers need no
            // memset (148 B / sample of fills at K = 32 for the three gradient streams).
            if (zero_tail) {
                for (uint32_t r0 = base + m; r0 < num_steps; r0 += 32) {
                    const uint32_t r = r0 + lane;
                    if (r < num_steps) {
                        const size_t z = (size_t)offset + r;
                        if (grad_sigmas) grad_sigmas[z] = 0.f;
                        if (grad_rgbs) { grad_rgbs[z * 3] = 0.f; grad_rgbs[z * 3 + 1] = 0.f; grad_rgbs[z * 3 + 2] = 0.f; }
                    }
                }
                if (KPL > 0) {
                    if (vec) {
                        float4* dst = reinterpret_cast<float4*>(grad_masks + (size_t)(offset + base + m) * K);
                        const uint32_t n16 = (num_steps - base - m) * C4;
                        for (uint32_t i = lane; i < n16; i += 32) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    } else {
                        float* dst = grad_masks + (size_t)(offset + base + m) * K;
                        const uint32_t nf = (num_steps - base - m) * K;
                        for (uint32_t i = lane; i < nf; i += 32) dst[i] = 0.f;
                    }
                }
            }
            break;
        }
        Tc = __shfl_sync(kFull, T_after, 31);
        Qc = __shfl_sync(kFull, Qrun, 31);
    }
}

// ---- inference (K == 0: raymarching.cu:1076-1163; K > 0: :1175-1271) ----
template <int KPL>
__global__ void __launch_bounds__(256) k_composite_infer(
    uint32_t n_alive, uint32_t n_step, uint32_t K, float T_thresh, int32_t* __restrict__ rays_alive, float* __restrict__ rays_t,
    const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ masks,
    const float* __restrict__ deltas, float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image,
    float* __restrict__ mask_out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= n_alive) return;
    const int32_t index = rays_alive[n];
    const size_t s0 = (size_t)n * n_step;

    float t = rays_t[index], weight_sum = weights_sum[index], d = depth[index];
    float r = image[(size_t)index * 3], g = image[(size_t)index * 3 + 1], b = image[(size_t)index * 3 + 2];
    float macc[KPL > 0 ? KPL : 1];
#pragma unroll
    for (int i = 0; i < (KPL > 0 ? KPL : 1); i++) {
        const uint32_t k = lane + 32u * i;
        macc[i] = (KPL > 0 && k < K) ? mask_out[(size_t)index * K + k] : 0.f;
    }
    uint32_t step = 0;
    bool stop = false;
    for (uint32_t base = 0; base < n_step && !stop; base += 32) {
        const bool valid = base + lane < n_step;
        const size_t s = s0 + base + lane;
        float d0 = 0.f, d1 = 0.f, alpha = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
        if (valid) {
            const float2 dl = __ldg(reinterpret_cast<const float2*>(deltas) + s);
            d0 = dl.x; d1 = dl.y;
            alpha = 1.0f - __expf(-__ldg(sigmas + s) * d0);
            cr = __ldg(rgbs + s * 3); cg = __ldg(rgbs + s * 3 + 1); cb = __ldg(rgbs + s * 3 + 2);
        }
        const uint32_t cnt = min(32u, n_step - base);
        for (uint32_t j = 0; j < cnt; j++) {
            if (__shfl_sync(kFull, d0, j) == 0.f) { stop = true; break; }  // ray finished (:1116)
            const float a = __shfl_sync(kFull, alpha, j);
            const float T = 1 - weight_sum;
            const float weight = a * T;
            weight_sum += weight;
            t += __shfl_sync(kFull, d1, j);
            d = fmaf(weight, t, d);
            r = fmaf(weight, __shfl_sync(kFull, cr, j), r);
            g = fmaf(weight, __shfl_sync(kFull, cg, j), g);
            b = fmaf(weight, __shfl_sync(kFull, cb, j), b);
            if (KPL > 0) {
                const float* mrow = masks + (s0 + base + j) * K;
#pragma unroll
                for (int i = 0; i < KPL; i++) {
                    const uint32_t k = lane + 32u * i;
                    if (k < K) macc[i] = fmaf(weight, __ldg(mrow + k), macc[i]);
                }
            }
            if (T < T_thresh) { stop = true; break; }
            step++;
        }
    }
    if (lane == 0) {
        if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
        weights_sum[index] = weight_sum;
        depth[index] = d;
        image[(size_t)index * 3] = r; image[(size_t)index * 3 + 1] = g; image[(size_t)index * 3 + 2] = b;
    }
    if (KPL > 0) {
#pragma unroll
        for (int i = 0; i < KPL; i++) {
            const uint32_t k = lane + 32u * i;
            if (k < K) mask_out[(size_t)index * K + k] = macc[i];
        }
    }
}

template <typename F>
int dispatch_kpl(uint32_t K, F&& f) {
    if (K == 0) return f(std::integral_constant<int, 0>{});
    if (K <= 32) return f(std::integral_constant<int, 1>{});
    if (K <= 64) return f(std::integral_constant<int, 2>{});
    if (K <= 32 * kMaxKPL) return f(std::integral_constant<int, kMaxKPL>{});
    return INERF_ERR_UNSUPPORTED;
}

int composite_train_fwd(const float* sigmas, const float* rgbs, const float* masks, const float* deltas, const int32_t* rays,
                        uint32_t M, uint32_t N, uint32_t K, float T_thresh, float* weights_sum, float* depth, float* image,
                        float* mask_out, void* stream) {
    if (N == 0) return INERF_OK;
    INERF_REQUIRE(rays); INERF_REQUIRE(weights_sum); INERF_REQUIRE(depth); INERF_REQUIRE(image);
    if (M) { INERF_REQUIRE(sigmas); INERF_REQUIRE(rgbs); INERF_REQUIRE(deltas); }
    if (K) { INERF_REQUIRE(mask_out); if (M) INERF_REQUIRE(masks); }
    if ((uintptr_t)deltas & 7u) return INERF_ERR_ALIGN;
    return dispatch_kpl(K, [&](auto kpl) {
        k_composite_train_fwd_scan<decltype(kpl)::value><<<div_up((unsigned long long)N * 32, 256), 256, 0, (cudaStream_t)stream>>>(
            sigmas, rgbs, masks, deltas, rays, M, N, K, T_thresh, weights_sum, depth, image, mask_out);
        INERF_LAUNCH_CHECK();
        return (int)INERF_OK;
    });
}

int composite_train_bwd(const float* grad_weights_sum, const float* grad_image, const float* grad_mask_out, const float* sigmas,
                        const float* rgbs, const float* masks, const float* deltas, const int32_t* rays, const float* weights_sum,
                        const float* image, const float* mask_out, uint32_t M, uint32_t N, uint32_t K, float T_thresh,
                        float* grad_sigmas, float* grad_rgbs, float* grad_masks, void* stream, bool dense = false) {
    if (N == 0 || M == 0) return INERF_OK;
    INERF_REQUIRE(grad_weights_sum); INERF_REQUIRE(grad_image); INERF_REQUIRE(sigmas); INERF_REQUIRE(rgbs); INERF_REQUIRE(deltas);
    INERF_REQUIRE(rays); INERF_REQUIRE(weights_sum); INERF_REQUIRE(image);
    if (!dense) { INERF_REQUIRE(grad_sigmas); INERF_REQUIRE(grad_rgbs); }
    if (K) { INERF_REQUIRE(grad_mask_out); INERF_REQUIRE(masks); INERF_REQUIRE(mask_out); INERF_REQUIRE(grad_masks); }
    return dispatch_kpl(K, [&](auto kpl) {
        if constexpr (decltype(kpl)::value <= 2) {   // K <= 64: the ray's logit gradients fit in registers
            constexpr unsigned threads = 32 * BwdCfg<decltype(kpl)::value>::kWarps;
            k_composite_train_bwd_scan<decltype(kpl)::value><<<div_up((unsigned long long)N * 32, threads), threads, 0, (cudaStream_t)stream>>>(
                grad_weights_sum, grad_image, grad_mask_out, sigmas, rgbs, masks, deltas, rays, weights_sum, image, mask_out, M, N, K,
                T_thresh, grad_sigmas, grad_rgbs, grad_masks, dense);
        } else {
            if (dense) return (int)INERF_ERR_UNSUPPORTED;   // K > 64: the thread-per-ray kernel keeps the reference's contract
            k_composite_train_bwd<decltype(kpl)::value><<<div_up((unsigned long long)N * 32, 256), 256, 0, (cudaStream_t)stream>>>(
                grad_weights_sum, grad_image, grad_mask_out, sigmas, rgbs, masks, deltas, rays, weights_sum, image, mask_out, M, N, K,
                T_thresh, grad_sigmas, grad_rgbs, grad_masks);
        }
        INERF_LAUNCH_CHECK();
        return (int)INERF_OK;
    });
}

int composite_infer(uint32_t n_alive, uint32_t n_step, uint32_t K, float T_thresh, int32_t* rays_alive, float* rays_t,
                    const float* sigmas, const float* rgbs, const float* masks, const float* deltas, float* weights_sum,
                    float* depth, float* image, float* mask_out, void* stream) {
    if (n_alive == 0 || n_step == 0) return INERF_OK;
    INERF_REQUIRE(rays_alive); INERF_REQUIRE(rays_t); INERF_REQUIRE(sigmas); INERF_REQUIRE(rgbs); INERF_REQUIRE(deltas);
    INERF_REQUIRE(weights_sum); INERF_REQUIRE(depth); INERF_REQUIRE(image);
    if (K) { INERF_REQUIRE(masks); INERF_REQUIRE(mask_out); }
    if ((uintptr_t)deltas & 7u) return INERF_ERR_ALIGN;
    return dispatch_kpl(K, [&](auto kpl) {
        k_composite_infer<decltype(kpl)::value><<<div_up((unsigned long long)n_alive * 32, 256), 256, 0, (cudaStream_t)stream>>>(
            n_alive, n_step, K, T_thresh, rays_alive, rays_t, sigmas, rgbs, masks, deltas, weights_sum, depth, image, mask_out);
        INERF_LAUNCH_CHECK();
        return (int)INERF_OK;
    });
}

}  // namespace

extern "C" int inerf_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                                  uint32_t M, uint32_t N, float T_thresh, float* weights_sum, float* depth,
                                                  float* image, void* stream) {
    return composite_train_fwd(sigmas, rgbs, nullptr, deltas, rays, M, N, 0, T_thresh, weights_sum, depth, image, nullptr, stream);
}
extern "C" int inerf_composite_rays_with_masks_train_forward(const float* sigmas, const float* rgbs, const float* masks,
                                                             const float* deltas, const int32_t* rays, uint32_t M, uint32_t N,
                                                             uint32_t K, float T_thresh, float* weights_sum, float* depth,
                                                             float* image, float* mask_out, void* stream) {
    if (K == 0) return INERF_ERR_SIZE;
    return composite_train_fwd(sigmas, rgbs, masks, deltas, rays, M, N, K, T_thresh, weights_sum, depth, image, mask_out, stream);
}
extern "C" int inerf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                                   const float* rgbs, const float* deltas, const int32_t* rays,
                                                   const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                                   float T_thresh, float* grad_sigmas, float* grad_rgbs, void* stream) {
    return composite_train_bwd(grad_weights_sum, grad_image, nullptr, sigmas, rgbs, nullptr, deltas, rays, weights_sum, image, nullptr,
                               M, N, 0, T_thresh, grad_sigmas, grad_rgbs, nullptr, stream);
}
extern "C" int inerf_composite_rays_with_masks_train_backward(const float* grad_weights_sum, const float* grad_image,
                                                              const float* grad_mask_out, const float* sigmas, const float* rgbs,
                                                              const float* masks, const float* deltas, const int32_t* rays,
                                                              const float* weights_sum, const float* image, const float* mask_out,
                                                              uint32_t M, uint32_t N, uint32_t K, float T_thresh, float* grad_sigmas,
                                                              float* grad_rgbs, float* grad_masks_acc, float* grad_masks, void* stream) {
    (void)grad_masks_acc;
    if (K == 0) return INERF_ERR_SIZE;
    return composite_train_bwd(grad_weights_sum, grad_image, grad_mask_out, sigmas, rgbs, masks, deltas, rays, weights_sum, image,
                               mask_out, M, N, K, T_thresh, grad_sigmas, grad_rgbs, grad_masks, stream);
}
extern "C" int inerf_composite_rays_with_masks_train_backward_dense(const float* grad_weights_sum, const float* grad_image,
                                                                    const float* grad_mask_out, const float* sigmas, const float* rgbs,
                                                                    const float* masks, const float* deltas, const int32_t* rays,
                                                                    const float* weights_sum, const float* image, const float* mask_out,
                                                                    uint32_t M, uint32_t N, uint32_t K, float T_thresh, float* grad_sigmas,
                                                                    float* grad_rgbs, float* grad_masks, void* stream) {
    if (K > 64) return INERF_ERR_UNSUPPORTED;
    return composite_train_bwd(grad_weights_sum, grad_image, K ? grad_mask_out : nullptr, sigmas, rgbs, K ? masks : nullptr, deltas, rays,
                               weights_sum, image, K ? mask_out : nullptr, M, N, K, T_thresh, grad_sigmas, grad_rgbs, K ? grad_masks : nullptr,
                               stream, true);
}
extern "C" int inerf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                                    const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum, float* depth,
                                    float* image, void* stream) {
    return composite_infer(n_alive, n_step, 0, T_thresh, rays_alive, rays_t, sigmas, rgbs, nullptr, deltas, weights_sum, depth, image,
                           nullptr, stream);
}
extern "C" int inerf_composite_rays_with_masks(uint32_t n_alive, uint32_t n_step, uint32_t K, float T_thresh, int32_t* rays_alive,
                                               float* rays_t, const float* sigmas, const float* rgbs, const float* masks,
                                               const float* deltas, float* weights_sum, float* depth, float* image, float* mask_out,
                                               void* stream) {
    if (K == 0) return INERF_ERR_SIZE;
    return composite_infer(n_alive, n_step, K, T_thresh, rays_alive, rays_t, sigmas, rgbs, masks, deltas, weights_sum, depth, image,
                           mask_out, stream);
}
