"""TEST / MEASUREMENT INFRASTRUCTURE ONLY (never the product, never bench.py's `value`).

"The kernel to beat on the same box" (SURVEY.md section 8d last paragraph): the reference's OWN CUDA path --
its kernels built unmodified for sm_100a into oracle/_ref (oracle/build_ref.sh) -- driven by a restatement of the Python
glue that surrounds them in the reference, because /root/reference itself does not travel to the GPU box:

  RefPath.render       NeRFMaskRenderer.run_cuda inference branch, nerf/mask_renderer.py:322-381 (alive-ray loop: march_rays,
                       two grid encodes with the per-call fp16 table cast of gridencoder/grid.py:43-44, sh_encode, nn.Linear
                       under fp16 autocast = network_mask.py:119-158, composite_rays_with_masks, boolean-index compaction)
  RefPath.train_step   one optimisation step of the reference's instance stage: Trainer.train_one_epoch's body
                       (nerf/utils.py:919-936) around MaskTrainer.train_step (:1287-1373) -> run_cuda training branch
                       (mask_renderer.py:256-320) -> raymarching.py:161-364 / grid.py:24-89 autograd wrappers (restated below as
                       _GridEncode / _CompositeMasksTrain on the reference kernels), torch.optim.Adam, GradScaler.

Used by tests/test_ref_cuda_path_gpu.py (full-frame parity of the product against this path) and, in a SUBPROCESS, by
bench.py's `ref_cuda_path` key (`python -m oracle.ref_cuda_path --json`).  The model object only supplies parameters and
buffers (the product's NeRFNetwork holds the same state-dict keys as the reference's); none of its kernels run here.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402


class _GridEncode(torch.autograd.Function):
    """gridencoder/grid.py:24-89 on the reference kernels (table cast to fp16 on every call under autocast, [L,B,C] kernel
    layout + permute, zeros_like(table) gradient)."""

    @staticmethod
    def forward(ctx, ref, inputs, embeddings, offsets, S, H):
        inputs = inputs.contiguous()
        B, D = inputs.shape
        L, C = offsets.shape[0] - 1, embeddings.shape[1]
        if torch.is_autocast_enabled() and C % 2 == 0:
            embeddings = embeddings.to(torch.half)
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
        ref.gridencoder.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, None, 0, False, 0)
        ctx.save_for_backward(inputs, embeddings, offsets)
        ctx.dims = (ref, B, D, C, L, S, H)
        return outputs.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets = ctx.saved_tensors
        ref, B, D, C, L, S, H = ctx.dims
        grad = grad.to(embeddings.dtype).view(B, L, C).permute(1, 0, 2).contiguous()
        grad_embeddings = torch.zeros_like(embeddings)
        ref.gridencoder.grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, None, None, 0, False, 0)
        return None, None, grad_embeddings, None, None, None


class _CompositeMasksTrain(torch.autograd.Function):
    """raymarching/raymarching.py:296-364 on the reference kernels."""

    @staticmethod
    def forward(ctx, ref, sigmas, rgbs, masks, deltas, rays, T_thresh):
        sigmas, rgbs, masks = sigmas.float().contiguous(), rgbs.float().contiguous(), masks.float().contiguous()
        M, N, K = sigmas.shape[0], rays.shape[0], masks.shape[1]
        dev = sigmas.device
        ws = torch.empty(N, device=dev); depth = torch.empty(N, device=dev); image = torch.empty(N, 3, device=dev)
        mask_out = torch.empty(N, K, device=dev)
        ref.raymarching.composite_rays_with_masks_train_forward(sigmas, rgbs, masks, deltas, rays, M, N, K, T_thresh, ws, depth, image, mask_out)
        ctx.save_for_backward(sigmas, rgbs, masks, deltas, rays, ws, depth, image, mask_out)
        ctx.dims = (ref, M, N, K, T_thresh)
        return ws, depth, image, mask_out

    @staticmethod
    def backward(ctx, g_ws, g_depth, g_image, g_mask_out):
        sigmas, rgbs, masks, deltas, rays, ws, depth, image, mask_out = ctx.saved_tensors
        ref, M, N, K, T_thresh = ctx.dims
        g_sig, g_rgb, g_masks = torch.zeros_like(sigmas), torch.zeros_like(rgbs), torch.zeros_like(masks)
        g_acc = torch.zeros_like(g_mask_out)
        ref.raymarching.composite_rays_with_masks_train_backward(g_ws.contiguous(), g_image.contiguous(), g_mask_out.contiguous(), sigmas, rgbs,
                                                                 masks, deltas, rays, ws, image, mask_out, M, N, K, T_thresh, g_sig, g_rgb,
                                                                 g_acc, g_masks)
        return None, g_sig, g_rgb, g_masks, None, None, None


class RefPath:
    """The reference operator sequence on the reference kernels; weights / tables / buffers are `model`'s."""

    def __init__(self, model, ref=None):
        self.m = model
        self.ref = ref if ref is not None else ref_loader.load()
        if self.ref is None:
            raise RuntimeError("oracle/_ref is not built (oracle/build_ref.sh where /root/reference exists)")
        self.S = float(np.log2(model.encoder.per_level_scale))
        self.H = int(model.encoder.base_resolution)
        self.local_step = 0
        self.step_counter = torch.zeros(16, 2, dtype=torch.int32, device=model.density_bitfield.device)

    # ---- field (network_mask.py:119-158) -------------------------------------------------------------------------
    def grid(self, enc, x):
        x01 = (x + self.m.bound) / (2 * self.m.bound)                      # grid.py:148
        return _GridEncode.apply(self.ref, x01, enc.embeddings, enc.offsets, self.S, self.H)

    def sh(self, d):
        out = torch.empty(d.shape[0], 16, device=d.device, dtype=torch.float32)
        self.ref.shencoder.sh_encode_forward(d.float().contiguous(), out, d.shape[0], 3, 4, None)
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
        h = self.mlp(m.sigma_net, self.grid(m.encoder, x))
        sigma = torch.exp(h[..., 0].float())                               # trunc_exp forward (activation.py:5-18)
        geo = h[..., 1:]
        rgb = torch.sigmoid(self.mlp(m.color_net, torch.cat([self.sh(d), geo], dim=-1)))
        logits = self.mlp(m.mask_net, torch.cat([self.grid(m.encoder_mask, x), geo], dim=-1))
        return sigma, rgb, logits

    # ---- inference (mask_renderer.py:322-381) -----------------------------------------------------------------------
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
            with torch.autocast("cuda", dtype=torch.float16):
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

    # ---- training (mask_renderer.py:256-320, raymarching.py:161-235, utils.py:1287-1373, 919-936) ----------------------
    def render_train(self, o, d, dt_gamma, max_steps, T_thresh, noises=None):
        m, rm = self.m, self.ref.raymarching
        N, dev = o.shape[0], o.device
        nears = torch.empty(N, device=dev); fars = torch.empty(N, device=dev)
        rm.near_far_from_aabb(o, d, m.aabb_train, N, m.min_near, nears, fars)
        counter = self.step_counter[self.local_step % 16]
        counter.zero_()
        self.local_step += 1
        M = N * max_steps                                                   # force_all_rays (patch training): raymarching.py:197
        xyzs = torch.zeros(M, 3, device=dev); dirs = torch.zeros(M, 3, device=dev); deltas = torch.zeros(M, 2, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        if noises is None:
            noises = torch.rand(N, device=dev)
        rm.march_rays_train(o, d, m.density_bitfield, m.bound, dt_gamma, max_steps, N, m.cascade, m.grid_size, M, nears, fars, xyzs, dirs,
                            deltas, rays, counter, noises)
        total = counter[0].item()                                           # D2H copy (raymarching.py:224)
        total += 128 - total % 128
        xyzs, dirs, deltas = xyzs[:total], dirs[:total], deltas[:total]
        torch.cuda.empty_cache()                                            # raymarching.py:231
        sigmas, rgbs, masks = self.field(xyzs, dirs)
        sigmas = m.density_scale * sigmas
        ws, depth, image, mask_out = _CompositeMasksTrain.apply(self.ref, sigmas, rgbs, masks, deltas, rays, T_thresh)
        image = image + (1 - ws).unsqueeze(-1) * 1
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        return dict(image=image, depth=depth, instance_mask_logits=mask_out, weights_sum=ws, total=total)

    def make_optimizer(self, lr=1e-2):
        m = self.m
        for mod in (m.encoder, m.sigma_net, m.encoder_dir, m.color_net):    # utils.py:1242-1246
            mod.requires_grad_(False)
        params = [p for mod in (m.encoder_mask, m.mask_net) for p in mod.parameters()]
        self.optimizer = torch.optim.Adam(params, lr=lr, betas=(0.9, 0.99), eps=1e-15)   # main_nerf_mask.py:182
        self.scaler = torch.amp.GradScaler("cuda", enabled=True)

    def train_step(self, data, patch=8, reg_weight=0.1, dt_gamma=1 / 128, max_steps=1024, T_thresh=1e-4):
        """-> loss tensor.  utils.py:929-936 around :1287-1373."""
        from oracle import host_oracle
        K = self.m.num_instances
        self.optimizer.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16):
            out = self.render_train(data["rays_o"].view(-1, 3), data["rays_d"].view(-1, 3), dt_gamma, max_steps, T_thresh, data.get("noises"))
            logits, gt = out["instance_mask_logits"].view(-1, K), data["masks"].view(-1)
            labeled = gt != -1
            if labeled.sum() > 0:                                           # host sync, as the reference (utils.py:1311)
                loss = F.cross_entropy(logits[labeled], gt[labeled], reduction="none").mean()
            else:
                loss = torch.zeros((), device=logits.device)
            if reg_weight > 0:
                loss = loss + host_oracle.label_regularization(out["depth"], logits, patch, K) * reg_weight
        self.scaler.scale(loss).backward()
        self.scaler.step(self.optimizer)
        self.scaler.update()
        self.last_total = out["total"]
        return loss


def _time(fn, n, warm):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(warm + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def measure(frames=3, train_steps=10):
    """-> dict with the reference CUDA path's c2 render time and c3 train-step time on cuda:0 (bench.py's workloads)."""
    import bench
    from oracle import host_oracle
    dev = torch.device("cuda:0")
    model, scene, poses = bench.build_scene_and_model(dev)
    rp = RefPath(model)
    kw = dict(dt_gamma=bench.DT_GAMMA, max_steps=bench.MAX_STEPS, T_thresh=bench.T_THRESH)
    rays = []
    for i in range(max(2, min(frames, 4))):
        r = host_oracle.get_rays(poses[i:i + 1], bench.intrinsics(), bench.H_IMG, bench.W_IMG)
        rays.append((r["rays_o"].reshape(-1, 3).contiguous().to(dev), r["rays_d"].reshape(-1, 3).contiguous().to(dev)))
    evaluated = []

    def frame(i):
        evaluated.append(rp.render(*rays[i % len(rays)], **kw)["evaluated"])
    ms_render = _time(frame, frames, 1)
    N = rays[0][0].shape[0]
    out = {"gpu": torch.cuda.get_device_name(0),
           "what": "oracle/_ref = the reference's raymarching.cu / gridencoder.cu / shencoder.cu compiled unmodified for sm_100a, driven by the "
                   "reference's run_cuda loop and autograd wrappers (restated in oracle/ref_cuda_path.py), MLPs = torch nn.Linear under fp16 autocast",
           "render": {"workload": bench.WORKLOAD, "ms_per_frame": ms_render, "mrays_per_s": N / ms_render / 1e3, "frames": frames,
                      "evaluated_samples_per_frame": float(np.mean(evaluated))}}
    if train_steps > 0:
        batches = bench.train_batches(dev, scene, poses, 4096, 0, 1, n=4)
        rp.make_optimizer()
        totals = []

        def step(i):
            rp.train_step(batches[i % len(batches)], patch=8, reg_weight=0.1, **kw)
            totals.append(rp.last_total)
        ms_train = _time(step, train_steps, 3)
        out["train_step"] = {"workload": "c3: 4096 rays (8x8 patches) x max_steps 1024, K=32, fp16 autocast, CE + label smoothness, backward, Adam",
                             "ms": ms_train, "steps": train_steps, "samples": float(np.mean(totals))}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--train-steps", type=int, default=10)
    ap.add_argument("--json", action="store_true", help="print exactly one JSON line on stdout")
    args = ap.parse_args()
    out = measure(args.frames, args.train_steps)
    print(json.dumps(out) if args.json else json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
