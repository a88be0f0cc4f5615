// Instance-stage loss tail in two launches: MaskTrainer.train_step's cross-entropy on
// labelled pixels (nerf/utils.py:1310-1314, criterion = CrossEntropyLoss(reduction='none'), main_nerf_mask.py:177) plus the
// depth-aware label smoothness on p x p patches (label_regularization, nerf/utils.py:1262-1285), forward and gradient.
// The reference spends ~150 tiny elementwise / reduction launches on these [N, K] tensors.
//
//   loss = sum_{labelled n} (logsumexp(x_n) - x_n[gt_n]) / n_labelled
//        + reg_w * ( sum diff_x^2 w_x / (K sum w_x) + sum diff_y^2 w_y / (K sum w_y) ),   w = exp(-(ddepth)^2)
//   (the reference divides by torch.sum(weight.expand_as(diff)), i.e. K times the per-pixel-pair sum)
// No gradient flows to depth: the compositor drops grad_depth (raymarching.py:342).
#include "common.cuh"

namespace {

struct LossAcc {   // accumulated by k_loss_reduce, consumed by k_loss_grad
    float ce_sum, n_lab, sx, wx, sy, wy;
};

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; i++) t += red[i];
    return t;
}

// one block per patch, one thread per pixel of the patch
__global__ void k_loss_reduce(const float* __restrict__ logits, const float* __restrict__ depth, const long long* __restrict__ labels, uint32_t K,
                              uint32_t p, int use_reg, LossAcc* __restrict__ acc) {
    __shared__ float red[32];
    const uint32_t pp = p * p, t = threadIdx.x;
    const uint32_t n = blockIdx.x * pp + t;
    const uint32_t i = t / p, j = t % p;
    const float* x = logits + (size_t)n * K;
    float ce = 0.f, lab = 0.f;
    const long long gt = labels[n];
    if (gt >= 0) {
        float m = -INFINITY;
        for (uint32_t k = 0; k < K; k++) m = fmaxf(m, x[k]);
        float s = 0.f;
        for (uint32_t k = 0; k < K; k++) s += expf(x[k] - m);
        ce = logf(s) + m - x[gt];
        lab = 1.f;
    }
    float sx = 0.f, wx = 0.f, sy = 0.f, wy = 0.f;
    if (use_reg) {
        if (j + 1 < p) {
            const float dd = depth[n + 1] - depth[n];
            wx = expf(-(dd * dd));
            const float* xr = x + K;
            float a = 0.f;
            for (uint32_t k = 0; k < K; k++) { const float df = xr[k] - x[k]; a = fmaf(df, df, a); }
            sx = a * wx;
        }
        if (i + 1 < p) {
            const float dd = depth[n + p] - depth[n];
            wy = expf(-(dd * dd));
            const float* xd = x + (size_t)p * K;
            float a = 0.f;
            for (uint32_t k = 0; k < K; k++) { const float df = xd[k] - x[k]; a = fmaf(df, df, a); }
            sy = a * wy;
        }
    }
    ce = block_sum(ce, red); lab = block_sum(lab, red);
    sx = block_sum(sx, red); wx = block_sum(wx, red); sy = block_sum(sy, red); wy = block_sum(wy, red);
    if (t == 0) {
        atomicAdd(&acc->ce_sum, ce); atomicAdd(&acc->n_lab, lab);
        atomicAdd(&acc->sx, sx); atomicAdd(&acc->wx, wx); atomicAdd(&acc->sy, sy); atomicAdd(&acc->wy, wy);
    }
}

// loss value (block 0, thread 0) and d loss / d logits * upstream
__global__ void k_loss_grad(const float* __restrict__ logits, const float* __restrict__ depth, const long long* __restrict__ labels, uint32_t K,
                            uint32_t p, float reg_w, const LossAcc* __restrict__ acc, const float* __restrict__ upstream, float* __restrict__ loss_out,
                            float* __restrict__ grad) {
    const uint32_t pp = p * p, t = threadIdx.x;
    const uint32_t n = blockIdx.x * pp + t;
    const uint32_t i = t / p, j = t % p;
    const LossAcc A = *acc;
    const float n_lab = fmaxf(A.n_lab, 1.f);
    const float Wx = (float)K * A.wx, Wy = (float)K * A.wy;
    if (loss_out != nullptr && blockIdx.x == 0 && t == 0) {
        float l = A.ce_sum / n_lab;
        if (reg_w > 0.f) l += reg_w * (A.sx / Wx + A.sy / Wy);
        loss_out[0] = l;
    }
    if (grad == nullptr) return;
    const float up = upstream ? upstream[0] : 1.f;
    const float* x = logits + (size_t)n * K;
    float* g = grad + (size_t)n * K;
    const long long gt = labels[n];
    float m = -INFINITY, inv = 0.f;
    if (gt >= 0) {
        for (uint32_t k = 0; k < K; k++) m = fmaxf(m, x[k]);
        float s = 0.f;
        for (uint32_t k = 0; k < K; k++) s += expf(x[k] - m);
        inv = 1.f / s;
    }
    float wl = 0.f, wr = 0.f, wu = 0.f, wd = 0.f;   // pair weights to the left / right / up / down neighbour
    if (reg_w > 0.f) {
        const float d0 = depth[n];
        if (j > 0) { const float dd = d0 - depth[n - 1]; wl = expf(-(dd * dd)); }
        if (j + 1 < p) { const float dd = depth[n + 1] - d0; wr = expf(-(dd * dd)); }
        if (i > 0) { const float dd = d0 - depth[n - p]; wu = expf(-(dd * dd)); }
        if (i + 1 < p) { const float dd = depth[n + p] - d0; wd = expf(-(dd * dd)); }
    }
    const float cx = reg_w > 0.f ? 2.f * reg_w / Wx : 0.f, cy = reg_w > 0.f ? 2.f * reg_w / Wy : 0.f;
    for (uint32_t k = 0; k < K; k++) {
        float v = 0.f;
        if (gt >= 0) v = (expf(x[k] - m) * inv - (k == (uint32_t)gt ? 1.f : 0.f)) / n_lab;
        if (reg_w > 0.f) {
            const float xc = x[k];
            float r = 0.f;
            if (j > 0) r += cx * wl * (xc - x[(long long)k - (long long)K]);
            if (j + 1 < p) r -= cx * wr * (x[k + K] - xc);
            if (i > 0) r += cy * wu * (xc - x[(long long)k - (long long)p * K]);
            if (i + 1 < p) r -= cy * wd * (x[k + (size_t)p * K] - xc);
            v += r;
        }
        g[k] = v * up;
    }
}

}  // namespace

extern "C" int inerf_mask_loss(const float* logits, const float* depth, const long long* labels, uint32_t N, uint32_t K, uint32_t patch,
                               float reg_weight, float* acc6, float* loss_out, void* stream) {
    INERF_REQUIRE(logits); INERF_REQUIRE(labels); INERF_REQUIRE(acc6); INERF_REQUIRE(loss_out);
    if (reg_weight > 0.f) INERF_REQUIRE(depth);
    if (patch == 0 || patch > 32 || K == 0 || N == 0 || N % (patch * patch)) return INERF_ERR_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(acc6, 0, sizeof(LossAcc), st);
    if (e != cudaSuccess) return (int)e;
    const uint32_t pp = patch * patch;
    k_loss_reduce<<<N / pp, pp, 0, st>>>(logits, depth, labels, K, patch, reg_weight > 0.f, (LossAcc*)acc6);
    INERF_LAUNCH_CHECK();
    k_loss_grad<<<1, 1, 0, st>>>(logits, depth, labels, K, patch, reg_weight, (const LossAcc*)acc6, nullptr, loss_out, nullptr);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_mask_loss_backward(const float* logits, const float* depth, const long long* labels, uint32_t N, uint32_t K, uint32_t patch,
                                        float reg_weight, const float* acc6, const float* upstream, float* grad_logits, void* stream) {
    INERF_REQUIRE(logits); INERF_REQUIRE(labels); INERF_REQUIRE(acc6); INERF_REQUIRE(grad_logits);
    if (reg_weight > 0.f) INERF_REQUIRE(depth);
    if (patch == 0 || patch > 32 || K == 0 || N == 0 || N % (patch * patch)) return INERF_ERR_SIZE;
    const uint32_t pp = patch * patch;
    k_loss_grad<<<N / pp, pp, 0, (cudaStream_t)stream>>>(logits, depth, labels, K, patch, reg_weight, (const LossAcc*)acc6, upstream, nullptr,
                                                          grad_logits);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}


// ---- Adam step of the trainable parameters (main_nerf_mask.py:182: torch.optim.Adam(betas=(0.9, 0.99), eps=1e-15)), one pass ----
// What torch runs per step on the 13.3 M-entry table: zero_grad (write g), the AMP finite check (read + write g), fused Adam
// (read p g m v, write p m v) = 9 passes over 53 MB.  Here: read p g m v, write p m v and g = 0 in ONE kernel (the finite check
// stays with GradScaler).  Arithmetic is torch's `adam_math` (fused_adam_utils.cuh) for amsgrad = maximize = weight_decay = 0:
//   g /= grad_scale;  m = lerp(m, g, 1 - b1);  v = b2 v + (1 - b2) g g;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// `step` is a device float (t - 1 on entry), so the launch can live in a CUDA graph; inerf_adam_advance bumps it afterwards.
// found_inf != 0 (non-finite gradients somewhere in the model): parameters and moments are left alone, g is still cleared.
namespace {
__global__ void __launch_bounds__(256) k_adam_step(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   uint64_t n, float lr, float b1, float b2, float eps, const float* __restrict__ step,
                                                   const float* __restrict__ grad_scale, const float* __restrict__ found_inf, float grad_div) {
    const bool skip = found_inf != nullptr && *found_inf != 0.f;
    const float t = *step + 1.0f;
    const float bc1 = 1.0f - (float)pow((double)b1, (double)t), bc2 = 1.0f - (float)pow((double)b2, (double)t);
    const float step_size = lr / bc1, bc2_sqrt = sqrtf(bc2);
    const float scale = (grad_scale ? *grad_scale : 1.0f) * grad_div;   // grad_div = 1 leaves the scale bit-identical
    const uint64_t n4 = n >> 2, stride = (uint64_t)gridDim.x * blockDim.x;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg = gg / scale;
        mm = mm + (1.0f - b1) * (gg - mm);                    // lerp(m, g, 1 - b1), weight < 0.5
        vv = b2 * vv + (1.0f - b2) * gg * gg;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        pp -= step_size * mm / denom;
    };
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 g4 = reinterpret_cast<const float4*>(g)[i];
        if (!skip) {
            float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
            upd(p4.x, g4.x, m4.x, v4.x); upd(p4.y, g4.y, m4.y, v4.y); upd(p4.z, g4.z, m4.z, v4.z); upd(p4.w, g4.w, m4.w, v4.w);
            reinterpret_cast<float4*>(p)[i] = p4; reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4;
        }
        reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (uint64_t i = (n4 << 2) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gg = g[i];
        if (!skip) { float pp = p[i], mm = m[i], vv = v[i]; upd(pp, gg, mm, vv); p[i] = pp; m[i] = mm; v[i] = vv; }
        g[i] = 0.f;
    }
}
__global__ void k_adam_advance(float* __restrict__ step, const float* __restrict__ found_inf) {
    if (found_inf == nullptr || *found_inf == 0.f) *step += 1.0f;
}
}  // namespace

extern "C" int inerf_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, uint64_t n, float lr, float beta1, float beta2,
                               float eps, const float* step, const float* grad_scale, const float* found_inf, float grad_div, void* stream) {
    if (n == 0) return INERF_OK;
    if (!(grad_div > 0.f)) return INERF_ERR_SIZE;
    INERF_REQUIRE(param); INERF_REQUIRE(grad); INERF_REQUIRE(exp_avg); INERF_REQUIRE(exp_avg_sq); INERF_REQUIRE(step);
    if ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15u) != 0) return INERF_ERR_ALIGN;
    const uint64_t want = ((n >> 2) + 255) / 256;
    const unsigned int blocks = (unsigned int)(want < 1 ? 1 : (want > (uint64_t)kNumSMs * 16 ? (uint64_t)kNumSMs * 16 : want));
    k_adam_step<<<blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, grad_scale, found_inf, grad_div);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_adam_advance(float* step, const float* found_inf, void* stream) {
    INERF_REQUIRE(step);
    k_adam_advance<<<1, 1, 0, (cudaStream_t)stream>>>(step, found_inf);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}
