"""Generates tests/golden/ref_run.npz by running the REFERENCE's own Python renderer
(NeRFNetwork / NeRFMaskRenderer.run from /root/reference/instance_nerf, imported unmodified) on the CPU.

Only runs where /root/reference exists (this container).  The three ops the reference implements in CUDA only
(near_far_from_aabb, GridEncoder, SHEncoder: raymarching.py:34-45, grid.py:54, sphere_harmonics.py:32) are
monkey-patched with the oracle's restatements -- those are pinned separately against the reference kernels
(tests/golden/ref_kernels.npz); everything else (network wiring, run(), staged render()) is the reference's code.

    python tests/golden/make_golden_cpu.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/instance_nerf"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))   # prebuilt _raymarching/_gridencoder/_shencoder (import only)
sys.path.insert(0, REF)

# viz / logging packages the reference imports at module scope but that are not installed here: empty stubs
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
import raymarching.raymarching as ref_rm  # noqa: E402
import gridencoder.grid as ref_grid  # noqa: E402
import shencoder.sphere_harmonics as ref_sh  # noqa: E402


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


ref_rm.near_far_from_aabb = _near_far
import raymarching  # noqa: E402
raymarching.near_far_from_aabb = _near_far
ref_grid.GridEncoder.forward = _grid_forward
ref_sh.SHEncoder.forward = _sh_forward

from nerf.network_mask import NeRFNetwork  # noqa: E402  (the reference's class)
from helpers import make_rays, scene_arrays  # noqa: E402

K, bound, H, W, T = 4, 2.0, 12, 16, 32
torch.manual_seed(0)
model = NeRFNetwork(encoding="hashgrid", bound=bound, cuda_ray=False, num_instances=K).eval()
g = torch.Generator().manual_seed(7)
with torch.no_grad():
    for enc in (model.encoder, model.encoder_mask):
        enc.embeddings.copy_((torch.rand(enc.embeddings.shape, generator=g) * 2 - 1) * 0.5)
sc, _, _, _ = scene_arrays(16, 8.0, 0)
o, d = make_rays(sc, H, W)
o = o * 0.25   # bring the camera inside the bound-2 volume
with torch.no_grad():
    res = model.render(o[None], d[None], staged=True, max_ray_batch=100, render_mask=True, num_steps=T, upsample_steps=0, perturb=False, bg_color=1)

small = {k: v.detach().numpy() for k, v in model.state_dict().items() if "embeddings" not in k}
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_run.npz"), rays_o=o.numpy(), rays_d=d.numpy(),
                    image=res["image"].numpy(), depth=res["depth"].numpy(), logits=res["instance_mask_logits"].numpy(),
                    cfg=np.array([K, bound, H, W, T, 7]), **{"sd_" + k: v for k, v in small.items()})
print("wrote tests/golden/ref_run.npz", res["image"].shape, float(res["image"].mean()), float(res["depth"].mean()))
