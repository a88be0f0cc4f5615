// Real spherical-harmonics encoding of the view direction (sm_100a).
//
// Replaces shencoder/src/shencoder.cu:28-129, 359-438 of the reference for the
// degrees the instance-field path uses (degree <= 4 -> 16 outputs,
// encoding.py:54-57).  One thread per direction; the 16 outputs of a thread are
// written as four float4 stores.  Polynomials are the closed forms of the real
// SH basis (Condon-Shortley sign convention as in the reference), evaluated in
// fp32 (sphere_harmonics.py:16 forces float32).
#include "common.cuh"

namespace {

// Evaluates the first degree^2 basis functions into o[].
__device__ __forceinline__ void sh_eval(float x, float y, float z, uint32_t degree, float* o) {
    o[0] = 0.28209479177387814f;
    if (degree <= 1) return;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    if (degree <= 2) return;
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    if (degree <= 3) return;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

// d(basis)/d(x,y,z) for degree <= 4, used only when directions require grad.
__device__ __forceinline__ void sh_grad(float x, float y, float z, uint32_t degree, float* gx, float* gy, float* gz) {
    const uint32_t n = degree * degree;
    for (uint32_t i = 0; i < n; i++) { gx[i] = 0.f; gy[i] = 0.f; gz[i] = 0.f; }
    if (degree <= 1) return;
    gy[1] = -0.48860251190291987f; gz[2] = 0.48860251190291987f; gx[3] = -0.48860251190291987f;
    if (degree <= 2) return;
    const float c4 = 1.0925484305920792f, c6 = 0.94617469575755997f, c8 = 0.54627421529603959f;
    gx[4] = c4 * y; gy[4] = c4 * x;
    gy[5] = -c4 * z; gz[5] = -c4 * y;
    gz[6] = 2.0f * c6 * z;
    gx[7] = -c4 * z; gz[7] = -c4 * x;
    gx[8] = 2.0f * c8 * x; gy[8] = -2.0f * c8 * y;
    if (degree <= 3) return;
    const float x2 = x * x, y2 = y * y, z2 = z * z;
    const float c9 = 0.59004358992664352f, c10 = 2.8906114426405538f, c11 = 0.45704579946446572f, c12 = 0.3731763325901154f,
                c14 = 1.4453057213202769f;
    gx[9] = -6.0f * c9 * x * y;            gy[9] = c9 * (-3.0f * x2 + 3.0f * y2);
    gx[10] = c10 * y * z;                  gy[10] = c10 * x * z;              gz[10] = c10 * x * y;
    gy[11] = c11 * (1.0f - 5.0f * z2);     gz[11] = -10.0f * c11 * y * z;
    gz[12] = c12 * (15.0f * z2 - 3.0f);
    gx[13] = c11 * (1.0f - 5.0f * z2);     gz[13] = -10.0f * c11 * x * z;
    gx[14] = 2.0f * c14 * x * z;           gy[14] = -2.0f * c14 * y * z;      gz[14] = c14 * (x2 - y2);
    gx[15] = c9 * (-3.0f * x2 + 3.0f * y2); gy[15] = 6.0f * c9 * x * y;
}

__global__ void __launch_bounds__(256) k_sh_fwd(const float* __restrict__ inputs, float* __restrict__ outputs, uint32_t B, uint32_t degree,
                                                float* __restrict__ dy_dx) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float x = __ldg(inputs + (size_t)b * 3), y = __ldg(inputs + (size_t)b * 3 + 1), z = __ldg(inputs + (size_t)b * 3 + 2);
    const uint32_t n = degree * degree;
    float o[16];
    sh_eval(x, y, z, degree, o);
    float* out = outputs + (size_t)b * n;
    if (degree == 4) {
#pragma unroll
        for (int i = 0; i < 4; i++) reinterpret_cast<float4*>(out)[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    } else {
        for (uint32_t i = 0; i < n; i++) out[i] = o[i];
    }
    if (dy_dx) {
        float gx[16], gy[16], gz[16];
        sh_grad(x, y, z, degree, gx, gy, gz);
        float* dd = dy_dx + (size_t)b * 3 * n;
        for (uint32_t i = 0; i < n; i++) { dd[i] = gx[i]; dd[n + i] = gy[i]; dd[2 * n + i] = gz[i]; }
    }
}

// shencoder.cu:359-382: grad_inputs[b, d] = sum_k grad[b, k] * dy_dx[b, d, k]
__global__ void k_sh_bwd(const float* __restrict__ grad, const float* __restrict__ dy_dx, uint32_t B, uint32_t n,
                         float* __restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 3) return;
    const uint32_t b = t / 3, d = t - b * 3;
    float r = 0;
    for (uint32_t k = 0; k < n; k++) r += grad[(size_t)b * n + k] * dy_dx[(size_t)b * 3 * n + d * n + k];
    grad_inputs[t] = r;
}

}  // namespace

extern "C" int inerf_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t degree, float* dy_dx,
                                       void* stream) {
    if (D != 3 || degree == 0 || degree > 8) return INERF_ERR_SIZE;
    if (degree > 4) return INERF_ERR_UNSUPPORTED;  // the instance-field path uses degree 4 (encoding.py:45-57)
    if (B == 0) return INERF_OK;
    INERF_REQUIRE(inputs); INERF_REQUIRE(outputs);
    if (degree == 4 && ((uintptr_t)outputs & 15u)) return INERF_ERR_ALIGN;
    k_sh_fwd<<<div_up(B, 256), 256, 0, (cudaStream_t)stream>>>(inputs, outputs, B, degree, dy_dx);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}

extern "C" int inerf_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t degree,
                                        const float* dy_dx, float* grad_inputs, void* stream) {
    (void)inputs;
    if (D != 3 || degree == 0 || degree > 8) return INERF_ERR_SIZE;
    if (degree > 4) return INERF_ERR_UNSUPPORTED;
    if (B == 0) return INERF_OK;
    INERF_REQUIRE(grad); INERF_REQUIRE(dy_dx); INERF_REQUIRE(grad_inputs);
    k_sh_bwd<<<div_up((unsigned long long)B * 3, 256), 256, 0, (cudaStream_t)stream>>>(grad, dy_dx, B, degree * degree, grad_inputs);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}
