"""TEST / MEASUREMENT INFRASTRUCTURE (not the product, not bench.py's value).

"The kernel to beat on the same box" (last paragraph): the reference's OWN CUDA path for a c2 frame,
i.e. the alive-ray loop of NeRFMaskRenderer.run_cuda (nerf/mask_renderer.py:322-381) driven through the reference's own
kernels built unmodified into oracle/_ref (march_rays, grid_encode_forward x2 with the per-call table cast of
gridencoder/grid.py:43-44, sh_encode_forward, composite_rays_with_masks) with the MLPs as torch nn.Linear under fp16
autocast (network_mask.py:119-158) -- timed with CUDA events next to this repo's one-launch renderer on the same frame,
and compared map by map.  Writes gpurun_out/ref_gpu_path.json.

    python tests/dev_ref_gpu_path.py [--frames 3]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402  (workload constants + scene builder)
from oracle import ref_loader  # noqa: E402


class RefPath:
    """The reference operator sequence on the reference kernels; weights / tables are the product model's parameters."""

    def __init__(self, model, ref):
        self.m, self.ref = model, ref
        self.S = float(np.log2(model.encoder.per_level_scale))

    def grid(self, enc, x):
        x01 = (x + self.m.bound) / (2 * self.m.bound)                      # grid.py:148
        table = enc.embeddings.to(torch.half)                              # grid.py:43-44, every call
        B = x01.shape[0]
        out = torch.empty(16, B, 2, device=x.device, dtype=torch.half)
        self.ref.gridencoder.grid_encode_forward(x01.contiguous(), table, enc.offsets, out, B, 3, 2, 16, self.S, 16, None, 0, False, 0)
        return out.permute(1, 0, 2).reshape(B, 32)                         # grid.py:57

    def sh(self, d):
        out = torch.empty(d.shape[0], 16, device=d.device, dtype=torch.float32)
        self.ref.shencoder.sh_encode_forward(d.contiguous(), out, d.shape[0], 3, 4, None)
        return out

    @staticmethod
    def mlp(net, h):
        for l in range(len(net)):
            h = net[l](h)
            if l != len(net) - 1:
                h = F.relu(h, inplace=True)
        return h

    def field(self, x, d):
        m = self.m
        with torch.autocast("cuda", dtype=torch.float16):
            h = self.mlp(m.sigma_net, self.grid(m.encoder, x))
            sigma = torch.exp(h[..., 0].float())
            geo = h[..., 1:]
            rgb = torch.sigmoid(self.mlp(m.color_net, torch.cat([self.sh(d), geo], dim=-1)))
            logits = self.mlp(m.mask_net, torch.cat([self.grid(m.encoder_mask, x), geo], dim=-1))
        return sigma, rgb, logits

    @torch.no_grad()
    def render(self, o, d, dt_gamma, max_steps, T_thresh):
        m, rm = self.m, self.ref.raymarching
        N, K, dev = o.shape[0], m.num_instances, o.device
        nears = torch.empty(N, device=dev); fars = torch.empty(N, device=dev)
        rm.near_far_from_aabb(o, d, m.aabb_infer, N, m.min_near, nears, fars)
        ws = torch.zeros(N, device=dev); depth = torch.zeros(N, device=dev); image = torch.zeros(N, 3, device=dev)
        logits_out = torch.zeros(N, K, device=dev)
        rays_alive = torch.arange(N, dtype=torch.int32, device=dev)
        rays_t = nears.clone()
        step, n_samples = 0, 0
        while step < max_steps:
            n_alive = rays_alive.shape[0]
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            M = n_alive * n_step
            M += 128 - (M % 128)
            xyzs = torch.zeros(M, 3, device=dev); dirs = torch.zeros(M, 3, device=dev); deltas = torch.zeros(M, 2, device=dev)
            noises = torch.zeros(n_alive, device=dev)
            rm.march_rays(n_alive, n_step, rays_alive, rays_t, o, d, m.bound, dt_gamma, max_steps, m.cascade, m.grid_size,
                          m.density_bitfield, nears, fars, xyzs, dirs, deltas, noises)
            sigmas, rgbs, masks = self.field(xyzs, dirs)
            sigmas = m.density_scale * sigmas
            rm.composite_rays_with_masks(n_alive, n_step, K, T_thresh, rays_alive, rays_t, sigmas.float().contiguous(),
                                         rgbs.float().contiguous(), masks.float().contiguous(), deltas, ws, depth, image, logits_out)
            rays_alive = rays_alive[rays_alive >= 0]
            n_samples += M
            step += n_step
        image = image + (1 - ws).unsqueeze(-1) * 1
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        return dict(image=image, depth=depth, instance_mask_logits=logits_out, evaluated=n_samples)


def time_frames(fn, frames, warm=1):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(frames):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / frames


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3)
    args = ap.parse_args()
    ref = ref_loader.load()
    assert ref is not None, "oracle/_ref is not built (oracle/build_ref.sh)"
    dev = torch.device("cuda:0")
    model, scene, poses = bench.build_scene_and_model(dev)
    rays = [tuple(t.to(dev) for t in bench.frame_rays(poses, i)) for i in range(max(args.frames, 2))]
    kw = dict(dt_gamma=bench.DT_GAMMA, max_steps=bench.MAX_STEPS, T_thresh=bench.T_THRESH)
    rp = RefPath(model, ref)

    o, d = rays[0]
    r0 = rp.render(o, d, **kw)
    with torch.no_grad():
        r1 = model.render(o[None], d[None], staged=True, render_mask=True, perturb=False, bg_color=1, **kw)
    err = dict(
        image=float((r1["image"][0] - r0["image"]).abs().max()),
        depth=float((r1["depth"][0] - r0["depth"]).abs().max()),
        instance_prob=float((torch.softmax(r1["instance_mask_logits"][0], -1) - torch.softmax(r0["instance_mask_logits"], -1)).abs().max()),
        instance_argmax_agree=float((r1["instance_mask_logits"][0].argmax(-1) == r0["instance_mask_logits"].argmax(-1)).float().mean()),
    )
    ms_ref = time_frames(lambda i: rp.render(*rays[i % len(rays)], **kw), args.frames)

    def ours(i):
        oo, dd = rays[i % len(rays)]
        with torch.no_grad():
            model.render(oo[None], dd[None], staged=True, render_mask=True, perturb=False, bg_color=1, **kw)
    ms_ours = time_frames(ours, max(args.frames, 5), warm=2)
    N = o.shape[0]
    out = dict(workload=bench.WORKLOAD, rays=N, gpu=torch.cuda.get_device_name(0),
               reference_cuda_path=dict(ms_per_frame=ms_ref, mrays_per_s=N / ms_ref / 1e3, evaluated_samples_frame0=r0["evaluated"],
                                        what="oracle/_ref kernels (reference sources, unmodified, sm_100a) + torch nn.Linear fp16 autocast, "
                                             "run_cuda alive-ray loop (mask_renderer.py:322-381)"),
               this_repo=dict(ms_per_frame=ms_ours, mrays_per_s=N / ms_ours / 1e3, what="NeRFNetwork.render -> inerf_render_fused"),
               speedup=ms_ref / ms_ours, max_abs_diff_frame0=err)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_gpu_path.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
