"""3D-mask projection: the labelled voxels of the predicted 3D instance masks rendered into per-view 2D label maps
(scripts/project_3d_masks.py:108-266 of the reference, the producer of the multi-view pseudo-labels config c4 is shaped after).

The reference builds a PyTorch3D point cloud from the labelled voxels and rasterises it once per camera pose (nearest point per
pixel).  Here the label volume stays a grid and every pixel ray -- the ray the NeRF render of that pose uses (`nerf.utils.get_rays`
-> inerf_get_rays) -- walks it with an exact 3D DDA (inerf_project_labels, csrc/project.cu); frames are independent, so a
multi-view job shards by whole frames over the GPUs like the c4 render (`parallel.shard_frames`).  No CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import call, ptr, stream_ptr


def generate_predicted_grid(masks, scores):
    """project_3d_masks.py:108-131: masks bool [X, Y, Z, M] (one 3D mask per detection), scores [M] -> int32 [X, Y, Z] with
    label m + 1 where detection m wins (0 = no detection).  Voxels claimed by several detections keep every detection whose
    score equals the best one there; their labels ADD, as in the reference (`instance_masks * instance_id` summed, :127-129)."""
    masks = torch.as_tensor(masks)
    scores = torch.as_tensor(scores, dtype=torch.float64, device=masks.device)
    X, Y, Z, M = masks.shape
    inst = masks.reshape(-1, M).to(torch.float64)
    overlap = inst.sum(-1) > 1
    sc = inst * scores[None, :]
    best = sc.max(-1, keepdim=True).values
    keep = torch.where(overlap[:, None], (sc >= best).to(torch.float64), inst)
    ids = torch.arange(1, M + 1, dtype=torch.float64, device=masks.device)
    return (keep * ids[None, :]).sum(-1).to(torch.int32).reshape(X, Y, Z)


def voxel_points(shape, room_bbox):
    """grid_pts_coord / grid2world (project_3d_masks.py:69-88): world position of voxel (i, j, k) = lo + (i, j, k) / shape * (hi - lo)
    -- the centres of the cells inerf_project_labels walks.  -> float64 [X*Y*Z, 3]"""
    bb = np.asarray(room_bbox, dtype=np.float64).reshape(2, 3)
    idx = np.stack(np.meshgrid(*[np.arange(s, dtype=np.float64) for s in shape], indexing="ij"), -1).reshape(-1, 3)
    return idx / np.asarray(shape, dtype=np.float64)[None] * (bb[1] - bb[0]) + bb[0]


def project_rays(rays_o, rays_d, labels, room_bbox, want_depth=False):
    """First labelled voxel along each ray.  rays [N, 3] CUDA float32, labels int32 [X, Y, Z] CUDA -> int32 [N] (, t float32 [N])."""
    if not (rays_o.is_cuda and labels.is_cuda):
        raise RuntimeError("project_rays: CUDA tensors only (there is no CPU fallback)")
    o = rays_o.float().contiguous().view(-1, 3)
    d = rays_d.float().contiguous().view(-1, 3)
    lab = labels.to(torch.int32).contiguous()
    N = o.shape[0]
    out = torch.empty(N, dtype=torch.int32, device=o.device)
    t = torch.empty(N, dtype=torch.float32, device=o.device) if want_depth else None
    bb = np.ascontiguousarray(np.asarray(room_bbox, dtype=np.float32).reshape(6))
    call("inerf_project_labels", ptr(o), ptr(d), N, ptr(lab), lab.shape[0], lab.shape[1], lab.shape[2], bb.ctypes.data, ptr(out), ptr(t),
         stream_ptr(o.device))
    return (out, t) if want_depth else out


def project_masks(labels, room_bbox, poses, intrinsics, H, W, rank=0, world=1):
    """Label maps of the frames this rank owns (frames rank, rank + world, ...): -> (frame indices, int32 [n, H, W])."""
    from .nerf.utils import get_rays
    from .parallel import shard_frames
    dev = labels.device
    mine = list(shard_frames(poses.shape[0], rank, world))
    out = torch.empty(len(mine), H, W, dtype=torch.int32, device=dev)
    for j, f in enumerate(mine):
        r = get_rays(poses[f][None].to(dev), intrinsics, H, W, -1)
        out[j] = project_rays(r["rays_o"], r["rays_d"], labels, room_bbox).view(H, W)
    return mine, out
