// inerf_project_labels: 3D instance-mask volume -> per-view 2D label maps by marching every pixel ray through the label grid.
//
// The reference's scripts/project_3d_masks.py:135-266 turns the labelled voxels of the predicted 3D masks into a PyTorch3D point
// cloud and rasterises it once per camera pose (PointsRasterizer, radius 0.015, points_per_pixel 1: the nearest labelled point
// along a pixel wins).  Here the same volume is a grid of cells centred on those points (grid_pts_coord, :72-83: point i sits at
// lo + i / n * (hi - lo)) and every pixel's ray -- the rays inerf_get_rays builds for the NeRF render of the same pose -- walks
// it with an exact 3D DDA (Amanatides-Woo): the first cell with a non-zero label is the pixel's label, its entry parameter the
// depth.  No PyTorch3D, no per-pose point-cloud copies; one thread per ray, the volume (a few MB) stays L2-resident.
// Every arithmetic step is a single IEEE fp32 operation in a fixed order (no contraction), so oracle/raymarch_oracle.c
// (oracle_project_labels) reproduces labels and depths bit for bit.
#include "common.cuh"

namespace {

struct Volume {
    const int32_t* labels;   // [nx, ny, nz], C order (project_3d_masks.py:186-188)
    int32_t n[3];
    float lo[3], scale[3];   // grid coordinate g = (p - lo) * scale + 0.5, scale = n / (hi - lo): cell i covers [i, i + 1)
};

__global__ void __launch_bounds__(256) k_project_labels(const float* __restrict__ rays_o, const float* __restrict__ rays_d, uint32_t N,
                                                        Volume v, int32_t* __restrict__ out_label, float* __restrict__ out_t) {
    const uint32_t r = blockIdx.x * 256u + threadIdx.x;
    if (r >= N) return;
    float go[3], gd[3], inv[3];
    float tn = 0.f, tf = 3.402823466e+38f;
    bool miss = false;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        go[a] = __fadd_rn(__fmul_rn(__fsub_rn(__ldg(rays_o + (size_t)r * 3 + a), v.lo[a]), v.scale[a]), 0.5f);
        gd[a] = __fmul_rn(__ldg(rays_d + (size_t)r * 3 + a), v.scale[a]);
        const float nf = (float)v.n[a];
        if (gd[a] != 0.f) {
            inv[a] = __fdiv_rn(1.0f, gd[a]);
            float t0 = __fmul_rn(__fsub_rn(0.f, go[a]), inv[a]), t1 = __fmul_rn(__fsub_rn(nf, go[a]), inv[a]);
            if (t0 > t1) { const float s = t0; t0 = t1; t1 = s; }
            tn = fmaxf(tn, t0);
            tf = fminf(tf, t1);
        } else {
            inv[a] = 0.f;
            if (go[a] < 0.f || go[a] >= nf) miss = true;
        }
    }
    int32_t label = 0;
    float t_hit = 0.f;
    if (!miss && tn < tf) {
        int32_t cell[3], step[3];
        float tmax[3], tdelta[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float p = __fadd_rn(go[a], __fmul_rn(tn, gd[a]));
            int32_t c = (int32_t)floorf(p);
            c = c < 0 ? 0 : (c > v.n[a] - 1 ? v.n[a] - 1 : c);
            cell[a] = c;
            step[a] = gd[a] > 0.f ? 1 : -1;
            if (gd[a] != 0.f) {
                const float edge = (float)(c + (gd[a] > 0.f ? 1 : 0));
                tmax[a] = __fmul_rn(__fsub_rn(edge, go[a]), inv[a]);
                tdelta[a] = fabsf(inv[a]);
            } else {
                tmax[a] = 3.402823466e+38f;
                tdelta[a] = 0.f;
            }
        }
        float t = tn;
        const int32_t limit = v.n[0] + v.n[1] + v.n[2] + 3;
        for (int32_t it = 0; it < limit; it++) {
            const int32_t l = __ldg(v.labels + ((size_t)cell[0] * v.n[1] + cell[1]) * v.n[2] + cell[2]);
            if (l != 0) { label = l; t_hit = t; break; }
            const int a = (tmax[0] <= tmax[1] && tmax[0] <= tmax[2]) ? 0 : (tmax[1] <= tmax[2] ? 1 : 2);
            t = tmax[a];
            // select by axis without dynamic register indexing
            if (a == 0) { cell[0] += step[0]; tmax[0] = __fadd_rn(tmax[0], tdelta[0]); }
            else if (a == 1) { cell[1] += step[1]; tmax[1] = __fadd_rn(tmax[1], tdelta[1]); }
            else { cell[2] += step[2]; tmax[2] = __fadd_rn(tmax[2], tdelta[2]); }
            if (cell[0] < 0 || cell[0] >= v.n[0] || cell[1] < 0 || cell[1] >= v.n[1] || cell[2] < 0 || cell[2] >= v.n[2]) break;
        }
    }
    out_label[r] = label;
    if (out_t) out_t[r] = t_hit;
}

}  // namespace

extern "C" int inerf_project_labels(const float* rays_o, const float* rays_d, uint32_t N, const int32_t* labels, uint32_t nx, uint32_t ny,
                                    uint32_t nz, const float* bbox_host, int32_t* out_label, float* out_t, void* stream) {
    if (N == 0) return INERF_OK;
    INERF_REQUIRE(rays_o); INERF_REQUIRE(rays_d); INERF_REQUIRE(labels); INERF_REQUIRE(bbox_host); INERF_REQUIRE(out_label);
    if (nx == 0 || ny == 0 || nz == 0 || nx > 4096 || ny > 4096 || nz > 4096) return INERF_ERR_SIZE;
    Volume v;
    v.labels = labels;
    const uint32_t n[3] = {nx, ny, nz};
    for (int a = 0; a < 3; a++) {
        const float ext = bbox_host[3 + a] - bbox_host[a];
        if (!(ext > 0.f)) return INERF_ERR_SIZE;
        v.n[a] = (int32_t)n[a];
        v.lo[a] = bbox_host[a];
        v.scale[a] = (float)n[a] / ext;
    }
    k_project_labels<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, N, v, out_label, out_t);
    INERF_LAUNCH_CHECK();
    return INERF_OK;
}
