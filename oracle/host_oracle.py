"""TEST INFRASTRUCTURE ONLY -- CPU (torch / numpy) restatement of the HOST-side pieces of the reference's instance-field
path: ray generation, the MaskTrainer loss tail, and the sampling front of the occupancy-grid update.  Nothing in the
product path may import this module; only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it.

Restated (paths relative to /root/reference/instance_nerf/):
  get_rays              nerf/utils.py:56-140 (full frame, uniform, 8x8 patch and error-map sampling)
  mask_cross_entropy    nerf/utils.py:1310-1314, 1349 (labelled pixels only, label -1 = unlabelled, mean)
  label_regularization  nerf/utils.py:1262-1285
  mask3d_loss           nerf/utils.py:1250-1260 (+ its weighting at :1367-1369)
  occupancy_update      nerf/mask_renderer.py:454-548 (full sweep and partial update) with the random draws injected
  mark_untrained_grid   nerf/mask_renderer.py:389-452

Pinned by tests/golden/ref_host.npz = outputs of the reference's own functions imported unmodified from /root/reference
(tests/golden/make_golden_host.py; tests/test_oracle_cpu.py::test_host_oracle_*).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import raymarch_oracle as ro


# ------------------------------------------------------------------------------------------------- rays --
@torch.no_grad()
def get_rays(poses: torch.Tensor, intr, H: int, W: int, N: int = -1, error_map=None, patch_size: int = 1, generator=None):
    """poses [B,4,4] cam2world -> dict(rays_o [B,N,3], rays_d [B,N,3], inds [B,N] if N > 0 [, inds_coarse]).
    `generator` replaces the global RNG of the reference (a fresh torch.Generator().manual_seed(s) reproduces what the
    reference draws after torch.manual_seed(s)); draw order as in the reference."""
    device = poses.device
    B = poses.shape[0]
    fx, fy, cx, cy = intr
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=device), torch.linspace(0, H - 1, H, device=device), indexing="ij")
    i = i.t().reshape([1, H * W]).expand([B, H * W]) + 0.5
    j = j.t().reshape([1, H * W]).expand([B, H * W]) + 0.5
    results = {}
    if N > 0:
        N = min(N, H * W)
        if patch_size > 1:
            num_patch = N // (patch_size ** 2)
            inds_x = torch.randint(0, H - patch_size, size=[num_patch], device=device, generator=generator)
            inds_y = torch.randint(0, W - patch_size, size=[num_patch], device=device, generator=generator)
            inds = torch.stack([inds_x, inds_y], dim=-1)
            pi, pj = torch.meshgrid(torch.arange(patch_size, device=device), torch.arange(patch_size, device=device), indexing="ij")
            offsets = torch.stack([pi.reshape(-1), pj.reshape(-1)], dim=-1)
            inds = (inds.unsqueeze(1) + offsets.unsqueeze(0)).view(-1, 2)
            inds = inds[:, 0] * W + inds[:, 1]
            inds = inds.expand([B, N])
        elif error_map is None:
            inds = torch.randint(0, H * W, size=[N], device=device, generator=generator).expand([B, N])
        else:
            inds_coarse = torch.multinomial(error_map.to(device), N, replacement=False, generator=generator)
            inds_x, inds_y = inds_coarse // 128, inds_coarse % 128
            sx, sy = H / 128, W / 128
            inds_x = (inds_x * sx + torch.rand(B, N, device=device, generator=generator) * sx).long().clamp(max=H - 1)
            inds_y = (inds_y * sy + torch.rand(B, N, device=device, generator=generator) * sy).long().clamp(max=W - 1)
            inds = inds_x * W + inds_y
            results["inds_coarse"] = inds_coarse
        i = torch.gather(i, -1, inds)
        j = torch.gather(j, -1, inds)
        results["inds"] = inds
    zs = torch.ones_like(i)
    xs = (i - cx) / fx * zs
    ys = (j - cy) / fy * zs
    directions = torch.stack((xs, ys, zs), dim=-1)
    directions = directions / torch.norm(directions, dim=-1, keepdim=True)
    rays_d = directions @ poses[:, :3, :3].transpose(-1, -2)
    rays_o = poses[..., :3, 3][..., None, :].expand_as(rays_d)
    results["rays_o"] = rays_o
    results["rays_d"] = rays_d
    return results


# ------------------------------------------------------------------------------------------------- loss --
def mask_cross_entropy(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """logits [N,K], labels int64 [N] (-1 = unlabelled) -> scalar: mean CE over the labelled pixels, 0 if there is none."""
    labeled = labels != -1
    if labeled.sum() > 0:
        return F.cross_entropy(logits[labeled], labels[labeled], reduction="none").mean()
    return torch.zeros((), dtype=logits.dtype)


def label_regularization(depth: torch.Tensor, pred_masks: torch.Tensor, patch: int, K: int) -> torch.Tensor:
    pm = pred_masks.view(-1, patch, patch, K).permute(0, 3, 1, 2).contiguous()
    diff_x = pm[:, :, :, 1:] - pm[:, :, :, :-1]
    diff_y = pm[:, :, 1:, :] - pm[:, :, :-1, :]
    depth = depth.view(-1, patch, patch)
    ddx = depth[:, :, 1:] - depth[:, :, :-1]
    ddy = depth[:, 1:, :] - depth[:, :-1, :]
    wx = torch.exp(-(ddx * ddx)).unsqueeze(1).expand_as(diff_x)
    wy = torch.exp(-(ddy * ddy)).unsqueeze(1).expand_as(diff_y)
    return torch.sum(diff_x * diff_x * wx) / torch.sum(wx) + torch.sum(diff_y * diff_y * wy) / torch.sum(wy)


def mask_train_loss(logits, depth, labels, patch, K, reg_weight, mask3d_logits=None, mask3d_labels=None, mask3d_weight=0.0):
    """The whole loss of MaskTrainer.train_step (nerf/utils.py:1310-1369) given what render / the 3D-mask query return."""
    loss = mask_cross_entropy(logits.view(-1, K), labels.view(-1))
    if reg_weight > 0:
        loss = loss + label_regularization(depth, logits, patch, K) * reg_weight
    if mask3d_weight > 0:
        loss = loss + F.cross_entropy(mask3d_logits, mask3d_labels, reduction="none").mean() * mask3d_weight
    return loss


# ---------------------------------------------------------------------------------------- occupancy grid --
def sweep_points(coords: np.ndarray, cas: int, G: int, bound: float, noise: np.ndarray) -> np.ndarray:
    """`xyzs = 2 * coords.float() / (G - 1) - 1; cas_xyzs = xyzs * (bound_c - hgs); cas_xyzs += (noise * 2 - 1) * hgs`
    (mask_renderer.py:480-487) in fp32, the Python scalars rounded to fp32 where they meet the tensor."""
    b = min(2 ** cas, bound)
    hgs = b / G
    xyz = (np.float32(2) * coords.astype(np.float32) / np.float32(G - 1) - np.float32(1)).astype(np.float32)
    out = (xyz * np.float32(b - hgs)).astype(np.float32)
    return (out + ((noise.astype(np.float32) * np.float32(2) - np.float32(1)) * np.float32(hgs)).astype(np.float32)).astype(np.float32)


def partial_cells(density_grid: np.ndarray, uniform_coords: np.ndarray, occ_picks: np.ndarray):
    """mask_renderer.py:498-513 for ONE cascade with the two randint draws injected: `uniform_coords` [N,3] and
    `occ_picks` [N] (positions in nonzero(density_grid > 0)).  -> (Morton indices [2N], coords [2N,3])."""
    indices = ro.morton3D(uniform_coords.astype(np.int32)).astype(np.int64)
    occ = np.nonzero(density_grid > 0)[0]
    occ_indices = occ[occ_picks]
    occ_coords = ro.morton3D_invert(occ_indices.astype(np.int32))
    return np.concatenate([indices, occ_indices]), np.concatenate([uniform_coords.astype(np.int32), occ_coords], axis=0)


def mark_untrained_grid(density_grid: np.ndarray, poses: np.ndarray, intrinsic, bound: float, G: int) -> np.ndarray:
    """mask_renderer.py:389-452 in fp32 numpy -> new density grid [C, G^3] (cells no camera sees = -1)."""
    C = density_grid.shape[0]
    fx, fy, cx, cy = intrinsic
    ii = np.arange(G, dtype=np.int32)
    X, Y, Z = np.meshgrid(ii, ii, ii, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], -1)
    indices = ro.morton3D(coords).astype(np.int64)
    world = (np.float32(2) * coords.astype(np.float32) / np.float32(G - 1) - np.float32(1)).astype(np.float32)
    poses = poses.astype(np.float32)
    count = np.zeros_like(density_grid)
    for cas in range(C):
        b = min(2 ** cas, bound)
        hgs = b / G
        cw = (world * np.float32(b - hgs)).astype(np.float32)
        cam = cw[None] - poses[:, None, :3, 3]
        cam = np.einsum("bni,bij->bnj", cam, poses[:, :3, :3]).astype(np.float32)
        mz = cam[:, :, 2] > 0
        mx = np.abs(cam[:, :, 0]) < (np.float32(cx / fx) * cam[:, :, 2]).astype(np.float32) + np.float32(hgs * 2)
        my = np.abs(cam[:, :, 1]) < (np.float32(cy / fy) * cam[:, :, 2]).astype(np.float32) + np.float32(hgs * 2)
        count[cas, indices] += (mz & mx & my).sum(0)
    out = density_grid.copy()
    out[count == 0] = -1
    return out
