"""Ray generation behind the reference's `nerf.utils.get_rays` signature (nerf/utils.py:56-140).

Pixel selection (uniform / patch / error-map sampling) is index bookkeeping on a handful of integers and stays in torch;
the per-ray arithmetic (pixel centre, normalise, rotate by the pose, optional fused near/far slab test) is ONE launch of
`inerf_get_rays` instead of the reference's full H*W meshgrid + ~10 elementwise / gather / matmul kernels per call.
CUDA tensors only (no CPU fallback)."""
from __future__ import annotations

import torch

from .._lib import call, ptr, stream_ptr


@torch.no_grad()
def sample_pixels(H: int, W: int, N: int, device, error_map=None, patch_size: int = 1, B: int = 1, generator=None):
    """utils.py:79-118 -> (inds int64 [N] or [B,N], extras dict).  Patch order is patch-major, then row, then column
    (the label regulariser relies on it, utils.py:1267-1273)."""
    N = min(N, H * W)
    extras = {}
    if patch_size > 1:
        num_patch = N // (patch_size ** 2)
        rows = torch.randint(0, H - patch_size, size=[num_patch], device=device, generator=generator)
        cols = torch.randint(0, W - patch_size, size=[num_patch], device=device, generator=generator)
        p = torch.arange(patch_size, device=device)
        inds = ((rows[:, None, None] + p[None, :, None]) * W + (cols[:, None, None] + p[None, None, :])).reshape(-1)
    elif error_map is None:
        inds = torch.randint(0, H * W, size=[N], device=device, generator=generator)
    else:
        coarse = torch.multinomial(error_map.to(device), N, replacement=False)          # [B, N] in [0, 128*128)
        cx, cy = coarse // 128, coarse % 128
        sx, sy = H / 128, W / 128
        ix = (cx * sx + torch.rand(B, N, device=device, generator=generator) * sx).long().clamp(max=H - 1)
        iy = (cy * sy + torch.rand(B, N, device=device, generator=generator) * sy).long().clamp(max=W - 1)
        inds = ix * W + iy
        extras["inds_coarse"] = coarse
    return inds, extras


@torch.no_grad()
def get_rays(poses, intrinsics, H, W, N=-1, error_map=None, patch_size=1, generator=None, aabb=None, min_near=0.2, inds=None):
    """poses [B,4,4] cam2world (CUDA), intrinsics (fx, fy, cx, cy) -> dict(rays_o [B,N,3], rays_d [B,N,3], inds [B,N] if N > 0
    [, inds_coarse]).  Extras (optional): `aabb` [6] fuses near_far_from_aabb and adds `nears`, `fars` [B,N]; `inds` (int64 [N]
    or [B,N]) injects the pixel choice instead of drawing it (parity tests against the reference's draws)."""
    if not poses.is_cuda:
        raise RuntimeError("get_rays: CUDA tensors only (no CPU fallback)")
    dev = poses.device
    B = poses.shape[0]
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    poses = poses.to(torch.float32).contiguous()
    results = {}
    if inds is not None:
        inds = inds.to(device=dev, dtype=torch.int64)
        N = inds.shape[-1]
        results["inds"] = inds.expand([B, N]) if inds.dim() == 1 else inds
    elif N > 0:
        inds, extras = sample_pixels(H, W, N, dev, error_map, patch_size, B, generator)
        results.update(extras)
        N = inds.shape[-1]
        results["inds"] = inds.expand([B, N]) if inds.dim() == 1 else inds
    else:
        N = H * W
    rays_o = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    rays_d = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    nears = fars = None
    if aabb is not None:
        aabb = aabb.to(dev, torch.float32).contiguous()
        nears = torch.empty(B, N, dtype=torch.float32, device=dev)
        fars = torch.empty(B, N, dtype=torch.float32, device=dev)
    st = stream_ptr(dev)
    if inds is not None and inds.dim() == 2:      # error-map sampling draws different pixels per pose: one launch per pose
        for b in range(B):
            call("inerf_get_rays", ptr(poses[b:b + 1]), 1, fx, fy, cx, cy, int(H), int(W), ptr(inds[b].contiguous()), N, ptr(rays_o[b]),
                 ptr(rays_d[b]), ptr(aabb), float(min_near), ptr(nears[b]) if nears is not None else None,
                 ptr(fars[b]) if fars is not None else None, st)
    else:
        call("inerf_get_rays", ptr(poses), B, fx, fy, cx, cy, int(H), int(W), ptr(inds.contiguous()) if inds is not None else None, N,
             ptr(rays_o), ptr(rays_d), ptr(aabb), float(min_near), ptr(nears), ptr(fars), st)
    results["rays_o"] = rays_o
    results["rays_d"] = rays_d
    if nears is not None:
        results["nears"], results["fars"] = nears, fars
    return results
