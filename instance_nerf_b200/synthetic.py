"""Deterministic synthetic "3D-FRONT-shaped" scenes, cameras and rays.

Everything here is host-side numpy / torch-CPU set-up code shared by tests and
bench.py: a room box with K-1 axis-aligned furniture boxes, its analytic
occupancy grid in the reference's layout (density_grid [C, 128^3] in Morton
order, nerf/renderer.py:91-93) and pinhole cameras on a seeded walk.  Rays come
from the product's `nerf.utils.get_rays` (CUDA) or, on the CPU side of a test,
from the pinned restatement `oracle.host_oracle.get_rays`.
"""
from __future__ import annotations

import math

import numpy as np
import torch


# ---------------------------------------------------------------- Morton (host) --
def _expand_bits(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64)
    v = (v * np.uint64(0x00010001)) & np.uint64(0xFF0000FF)
    v = (v * np.uint64(0x00000101)) & np.uint64(0x0F00F00F)
    v = (v * np.uint64(0x00000011)) & np.uint64(0xC30C30C3)
    v = (v * np.uint64(0x00000005)) & np.uint64(0x49249249)
    return v


def morton3D_np(x, y, z) -> np.ndarray:
    """10-bit-per-axis Morton code (raymarching.cu:56-71), numpy."""
    return (_expand_bits(np.asarray(x)) | (_expand_bits(np.asarray(y)) << np.uint64(1)) | (_expand_bits(np.asarray(z)) << np.uint64(2))).astype(np.int64)


# ---------------------------------------------------------------------- scene --
class RoomScene:
    """Room shell (floor, ceiling, 4 walls, 0.1 thick) + K-1 furniture boxes resting on the floor.
    solids: list of (instance_id, lo[3], hi[3]); instance 0 = structure / background."""

    def __init__(self, num_instances: int = 16, bound: float = 8.0, seed: int = 0):
        rng = np.random.RandomState(seed)
        self.bound = float(bound)
        self.K = int(num_instances)
        rx, ry, rz, th = 3.0, 1.4, 2.5, 0.1
        self.room = (np.array([-rx, -ry, -rz]), np.array([rx, ry, rz]))
        solids = [
            (0, [-rx, -ry - th, -rz], [rx, -ry, rz]),            # floor
            (0, [-rx, ry, -rz], [rx, ry + th, rz]),              # ceiling
            (0, [-rx - th, -ry, -rz], [-rx, ry, rz]),            # walls
            (0, [rx, -ry, -rz], [rx + th, ry, rz]),
            (0, [-rx, -ry, -rz - th], [rx, ry, -rz]),
            (0, [-rx, -ry, rz], [rx, ry, rz + th]),
        ]
        for k in range(1, self.K):
            size = rng.uniform(0.3, 1.5, size=3)
            size[1] = min(size[1], 2 * ry - 0.2)
            cx = rng.uniform(-rx + size[0] / 2, rx - size[0] / 2)
            cz = rng.uniform(-rz + size[2] / 2, rz - size[2] / 2)
            lo = [cx - size[0] / 2, -ry, cz - size[2] / 2]
            hi = [cx + size[0] / 2, -ry + size[1], cz + size[2] / 2]
            solids.append((k, lo, hi))
        self.solids = [(i, np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)) for i, lo, hi in solids]

    def inside(self, pts: np.ndarray) -> np.ndarray:
        """instance id + 1 of the solid containing each point (0 = free space); later solids win."""
        out = np.zeros(pts.shape[0], dtype=np.int32)
        for inst, lo, hi in self.solids:
            m = np.all((pts >= lo) & (pts <= hi), axis=1)
            out[m] = inst + 1
        return out

    def density_grid(self, cascade: int, H: int = 128, value: float = 50.0) -> np.ndarray:
        """[cascade, H^3] float32 in Morton order: `value` inside solids, 0 in free space.  Cell centres as in
        update_extra_state (mask_renderer.py:480-487); a cell is solid if its centre or any corner is."""
        ii = np.arange(H)
        X, Y, Z = np.meshgrid(ii, ii, ii, indexing="ij")
        idx = morton3D_np(X.ravel(), Y.ravel(), Z.ravel())
        grid = np.zeros((cascade, H ** 3), dtype=np.float32)
        unit = 2 * ii.astype(np.float64) / (H - 1) - 1
        for c in range(cascade):
            b = min(2 ** c, self.bound)
            hgs = b / H
            ctr = unit * (b - hgs)                      # per-axis cell centres
            occ = np.zeros((H, H, H), dtype=bool)
            for _, lo, hi in self.solids:               # per-axis interval overlap, then outer product
                m = [(ctr + hgs >= lo[a]) & (ctr - hgs <= hi[a]) for a in range(3)]
                occ |= m[0][:, None, None] & m[1][None, :, None] & m[2][None, None, :]
            grid[c, idx] = np.where(occ.ravel(), value, 0.0).astype(np.float32)
        return grid

    def first_hit_labels(self, rays_o: np.ndarray, rays_d: np.ndarray) -> np.ndarray:
        """Analytic first-hit instance id per ray (slab test against every solid); -1 = miss (unlabelled)."""
        N = rays_o.shape[0]
        best_t = np.full(N, np.inf)
        label = np.full(N, -1, dtype=np.int64)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / rays_d.astype(np.float64)
            for inst, lo, hi in self.solids:
                t0 = (lo - rays_o) * inv
                t1 = (hi - rays_o) * inv
                tn = np.nanmax(np.minimum(t0, t1), axis=1)
                tf = np.nanmin(np.maximum(t0, t1), axis=1)
                hit = (tf >= np.maximum(tn, 0.0)) & (tf > 0)
                t = np.where(tn > 0, tn, tf)
                upd = hit & (t < best_t)
                best_t[upd] = t[upd]
                label[upd] = inst
        return label


def packbits_np(grid: np.ndarray, thresh: float) -> np.ndarray:
    """raymarching.cu:267-289 on the host: bit i of byte n = grid_flat[8n+i] > thresh."""
    bits = (grid.reshape(-1, 8) > thresh).astype(np.uint8)
    return (bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8)


# -------------------------------------------------------------------- cameras --
def look_at_pose(eye: np.ndarray, target: np.ndarray, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """cam2world [4,4] with the camera looking along +z (the convention of get_rays: dir = (x, y, 1))."""
    f = target - eye
    f = f / np.linalg.norm(f)
    r = np.cross(np.asarray(up, dtype=np.float64), f)
    if np.linalg.norm(r) < 1e-6:
        r = np.array([1.0, 0.0, 0.0])
    r = r / np.linalg.norm(r)
    u = np.cross(f, r)
    pose = np.eye(4)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = r, u, f, eye
    return pose


def camera_poses(scene: RoomScene, n: int, seed: int = 1) -> np.ndarray:
    """n poses on a seeded random walk inside the room at height 0, looking at random box centres."""
    rng = np.random.RandomState(seed)
    lo, hi = scene.room
    eye = np.array([0.0, 0.0, 0.0])
    poses = []
    boxes = [s for s in scene.solids if s[0] > 0] or scene.solids
    for _ in range(n):
        eye = eye + rng.normal(0, 0.3, size=3)
        eye[1] = 0.0
        eye = np.clip(eye, lo + 0.4, hi - 0.4)
        inst, blo, bhi = boxes[rng.randint(len(boxes))]
        tgt = (blo + bhi) / 2
        if np.linalg.norm(tgt - eye) < 0.2:
            tgt = tgt + np.array([0.5, 0.0, 0.5])
        poses.append(look_at_pose(eye.copy(), tgt))
    return np.stack(poses).astype(np.float32)


def intrinsics(H: int, W: int, fovy_deg: float = 60.0):
    fl = 0.5 * H / math.tan(math.radians(fovy_deg) / 2)
    return (fl, fl, W / 2, H / 2)


# ---------------------------------------------------------------------- model --
def randomize_tables(model, seed: int = 0, scale: float = 0.5):
    """Embeddings re-drawn U(-scale, scale): the default +-1e-4 init (grid.py:139-140) gives a constant field."""
    g = torch.Generator().manual_seed(seed)
    for enc in (model.encoder, model.encoder_mask):
        with torch.no_grad():
            enc.embeddings.copy_((torch.rand(enc.embeddings.shape, generator=g) * 2 - 1) * scale)


def install_scene(model, scene: RoomScene, thresh: float | None = None):
    """Analytic occupancy into the model's density_grid / density_bitfield buffers."""
    grid = scene.density_grid(model.cascade, model.grid_size)
    mean = float(np.clip(grid, 0, None).mean())
    t = min(mean, model.density_thresh) if thresh is None else thresh
    bits = packbits_np(grid, t)
    with torch.no_grad():
        model.density_grid.copy_(torch.from_numpy(grid))
        model.density_bitfield.copy_(torch.from_numpy(bits))
    model.mean_density = mean
    return grid, bits
