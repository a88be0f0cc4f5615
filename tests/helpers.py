"""Shared test fixtures: small synthetic scenes, rays, and the canonical form of the reference's sample stream."""
import functools

import numpy as np
import torch

from instance_nerf_b200 import synthetic
from oracle import host_oracle


@functools.lru_cache(maxsize=4)
def scene_arrays(K=16, bound=8.0, seed=0):
    sc = synthetic.RoomScene(K, bound, seed)
    cascade = 1 + int(np.ceil(np.log2(bound)))
    grid = sc.density_grid(cascade)
    bits = synthetic.packbits_np(grid, min(float(grid.mean()), 10.0))
    return sc, cascade, grid, bits


def make_rays(sc, H, W, n_poses=1, seed=1, pose_index=0):
    poses = synthetic.camera_poses(sc, max(n_poses, pose_index + 1), seed)[pose_index:pose_index + n_poses]
    r = host_oracle.get_rays(torch.from_numpy(poses), synthetic.intrinsics(H, W), H, W)
    return r["rays_o"].reshape(-1, 3).contiguous(), r["rays_d"].reshape(-1, 3).contiguous()


def adversarial_rays(bound):
    """Axis-parallel directions (rd = inf), rays grazing cell faces, rays starting outside / missing the box."""
    o, d = [], []
    for ax in range(3):
        for sgn in (1.0, -1.0):
            dd = [0.0, 0.0, 0.0]; dd[ax] = sgn
            o.append([0.1, -0.2, 0.3]); d.append(dd)
            oo = [0.0, 0.0, 0.0]; oo[(ax + 1) % 3] = 1.0 / 64  # exactly on a cell face of cascade 0
            o.append(oo); d.append(dd)
    o.append([20.0, 20.0, 20.0]); d.append([0.0, 1.0, 0.0])      # misses the aabb
    o.append([-20.0, 0.3, 0.1]); d.append([1.0, 0.0, 0.0])       # starts outside, hits
    o.append([0.0, 0.0, 0.0]); d.append([0.57735027, 0.57735027, 0.57735027])
    o.append([bound, bound, bound]); d.append([-0.57735027, -0.57735027, -0.57735027])
    return torch.tensor(o, dtype=torch.float32), torch.tensor(d, dtype=torch.float32)


def canonicalize(rays, xyzs, dirs, deltas, M=None):
    """sort `rays` by ray id, offsets := exclusive prefix sum of counts, samples
    re-packed in that order.  Inputs may be torch (any device) or numpy; returns numpy."""
    rays, xyzs, dirs, deltas = (t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t) for t in (rays, xyzs, dirs, deltas))
    order = np.argsort(rays[:, 0], kind="stable")
    rays = rays[order]
    counts = rays[:, 2].astype(np.int64)
    old_off = rays[:, 1].astype(np.int64)
    new_off = np.concatenate([[0], np.cumsum(counts)[:-1]])
    total = int(counts.sum())
    cx = np.zeros((total, 3), np.float32); cd = np.zeros((total, 3), np.float32); cl = np.zeros((total, 2), np.float32)
    limit = xyzs.shape[0] if M is None else M
    for o, n, c in zip(old_off, new_off, counts):
        if c and o + c <= limit:
            cx[n:n + c] = xyzs[o:o + c]; cd[n:n + c] = dirs[o:o + c]; cl[n:n + c] = deltas[o:o + c]
    out_rays = np.stack([rays[:, 0], new_off, counts], -1).astype(np.int32)
    return out_rays, cx, cd, cl


def bits_equal(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        return np.array_equal(a.view(np.uint32 if a.itemsize == 4 else np.uint16), b.view(np.uint32 if b.itemsize == 4 else np.uint16))
    return np.array_equal(a, b)


def record_metric(name, **kv):
    """Append a measured figure of a GPU test to gpurun_out/test_metrics.jsonl (brought back by gpurun; best effort)."""
    import json
    import os
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", "test_metrics.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **kv}) + "\n")
    except OSError:
        pass
