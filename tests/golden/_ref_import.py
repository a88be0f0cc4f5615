"""Shared prelude of the golden generators that import the REFERENCE's Python (only where /root/reference exists).

* sys.path: the repo, tests/, oracle/_ref (prebuilt _raymarching / _gridencoder / _shencoder: import only) and
  /root/reference/instance_nerf;
* empty stub modules for the viz / logging packages the reference imports at module scope but that are not installed;
* the ops the reference implements in CUDA only (near_far_from_aabb, GridEncoder, SHEncoder, morton3D, morton3D_invert,
  packbits: raymarching.py:34-45, 51-108, grid.py:54, sphere_harmonics.py:32) are monkey-patched with the oracle's CPU
  restatements, which are pinned separately against the reference kernels (tests/golden/ref_kernels.npz).
Everything else that runs is the reference's own code.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/instance_nerf"
for p in (REF, os.path.join(ROOT, "oracle", "_ref"), os.path.join(ROOT, "tests"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

for name in ("trimesh", "mcubes", "tensorboardX", "torch_ema", "lpips", "torchmetrics", "torchmetrics.functional", "imageio", "matplotlib",
             "matplotlib.pyplot", "h5py", "wandb", "cv2"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
sys.modules["torch_ema"].__dict__.setdefault("ExponentialMovingAverage", object)
sys.modules["torchmetrics.functional"].__dict__.setdefault("structural_similarity_index_measure", None)

from oracle import field_oracle as fo  # noqa: E402
from oracle import raymarch_oracle as ro  # noqa: E402
import raymarching.raymarching as ref_rm  # noqa: E402
import gridencoder.grid as ref_grid  # noqa: E402
import shencoder.sphere_harmonics as ref_sh  # noqa: E402
import raymarching  # noqa: E402


def _near_far(rays_o, rays_d, aabb, min_near=0.2):
    return fo.near_far(rays_o, rays_d, aabb, min_near)


def _grid_forward(self, inputs, bound=1):
    x01 = (inputs + bound) / (2 * bound)
    prefix = list(x01.shape[:-1])
    out = fo.grid_encode(x01.view(-1, 3), self.embeddings.detach(), self.offsets.numpy(), self.per_level_scale, self.base_resolution)
    return out.view(prefix + [self.output_dim])


def _sh_forward(self, inputs, size=1):
    prefix = list(inputs.shape[:-1])
    return fo.sh_encode((inputs / size).reshape(-1, 3), self.degree).reshape(prefix + [self.output_dim])


def _morton3D(coords):
    return torch.from_numpy(ro.morton3D(coords.numpy().astype(np.int32))).to(torch.int32)


def _morton3D_invert(indices):
    return torch.from_numpy(ro.morton3D_invert(indices.numpy().astype(np.int32))).to(torch.int32)


def _packbits(grid, thresh, bitfield=None):
    return torch.from_numpy(ro.packbits(grid.contiguous().numpy().astype(np.float32), float(thresh)))


for mod in (ref_rm, raymarching):
    mod.near_far_from_aabb = _near_far
    mod.morton3D = _morton3D
    mod.morton3D_invert = _morton3D_invert
    mod.packbits = _packbits
ref_grid.GridEncoder.forward = _grid_forward
ref_sh.SHEncoder.forward = _sh_forward
