"""Generates tests/golden/ref_host.npz: outputs of the REFERENCE's own host-side functions on the instance-field path,
imported unmodified from /root/reference/instance_nerf and run on the CPU (only where /root/reference exists):

  get_rays                         nerf/utils.py:56-140        full frame, uniform, 8x8 patches, error-map sampling
  MaskTrainer.train_step's loss    nerf/utils.py:1287-1373     CE + label_regularization + mask3d_loss (render stubbed to
                                                               return fixed maps; the 3D-mask query runs the reference net)
  mark_untrained_grid              nerf/mask_renderer.py:389-452
  update_extra_state               nerf/mask_renderer.py:454-548  full sweep and partial update, every random draw recorded

CUDA-only ops are replaced by the oracle's CPU restatements (tests/golden/_ref_import.py).  The occupancy grid is 16^3 per
cascade (the attribute the reference hard-codes to 128, renderer.py:74, is overridden after construction) to keep the
fixture small; nothing in the code under test depends on the value.

    python tests/golden/make_golden_host.py
"""
import json
import os
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

import _ref_import as ri  # noqa: F401  (sys.path, stubs, CPU patches)
from nerf.network_mask import NeRFNetwork  # noqa: E402  (the reference's class)
from nerf.utils import MaskTrainer, get_rays  # noqa: E402
from helpers import scene_arrays  # noqa: E402
from instance_nerf_b200 import synthetic  # noqa: E402

out = {}
sc, _, _, _ = scene_arrays(16, 8.0, 0)
poses = torch.from_numpy(synthetic.camera_poses(sc, 5, 1)).float()
out["poses"] = poses.numpy()

# ------------------------------------------------------------------------------------------------ get_rays --
cases = {"full": dict(B=2, H=12, W=16, N=-1, patch=1, seed=0, emap=False),
         "uniform": dict(B=2, H=12, W=16, N=50, patch=1, seed=12, emap=False),
         "patch": dict(B=1, H=40, W=48, N=256, patch=8, seed=11, emap=False),
         "emap": dict(B=2, H=160, W=200, N=64, patch=1, seed=13, emap=True)}
for name, c in cases.items():
    intr = synthetic.intrinsics(c["H"], c["W"])
    emap = None
    if c["emap"]:
        emap = ((torch.arange(c["B"] * 128 * 128) % 97).float() + 1).view(c["B"], -1) / 97.0
    torch.manual_seed(c["seed"])
    r = get_rays(poses[:c["B"]], intr, c["H"], c["W"], c["N"], error_map=emap, patch_size=c["patch"])
    out[f"rays_{name}_cfg"] = np.array([c["B"], c["H"], c["W"], c["N"], c["patch"], c["seed"], int(c["emap"])])
    for k, v in r.items():
        out[f"rays_{name}_{k}"] = v.contiguous().numpy()

# ---------------------------------------------------------------------------------------------- loss tail --
K, bound = 4, 2.0
torch.manual_seed(0)
model = NeRFNetwork(encoding="hashgrid", bound=bound, cuda_ray=True, num_instances=K, density_scale=1, density_thresh=0.5)
g = torch.Generator().manual_seed(7)
with torch.no_grad():
    for enc in (model.encoder, model.encoder_mask):
        enc.embeddings.copy_((torch.rand(enc.embeddings.shape, generator=g) * 2 - 1) * 0.5)
small = {k: v.detach().numpy() for k, v in model.state_dict().items() if "embeddings" not in k and "density" not in k}
out.update({"sd_" + k: v for k, v in small.items()})
out["model_cfg"] = np.array([K, bound, 7])
# the reference network's complete state-dict layout (names, shapes, dtypes): what a reference checkpoint contains
out["ref_sd_layout"] = np.array(json.dumps([[k, list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()]))

gl = torch.Generator().manual_seed(21)
N, patch = 256, 8
logits = (torch.randn(1, N, K, generator=gl) * 2).requires_grad_(True)
depth = torch.rand(1, N, generator=gl)
labels = torch.randint(-1, K, (1, N), generator=gl)
coords3d = (torch.rand(200, 3, generator=gl) * 2 - 1) * bound
labels3d = torch.randint(0, K, (200,), generator=gl)
render_stub = SimpleNamespace(render=lambda *a, **kw: {"instance_mask_logits": logits, "depth": depth}, density=model.density, mask=model.mask)
for tag, reg_w, m3_w in (("ce", 0.0, 0.0), ("reg", 0.1, 0.0), ("all", 0.1, 0.5)):
    me = SimpleNamespace(model=render_stub, opt=SimpleNamespace(patch_size=patch, label_regularization_weight=reg_w, mask3d_loss_weight=m3_w),
                         criterion=nn.CrossEntropyLoss(reduction="none"), num_instances=K, error_map=None, device=torch.device("cpu"))
    me.label_regularization = lambda d, p, me=me: MaskTrainer.label_regularization(me, d, p)
    me.mask3d_loss = lambda data, me=me: MaskTrainer.mask3d_loss(me, data)
    data = {"rays_o": None, "rays_d": None, "masks": labels, "mask3d_coords": coords3d, "mask3d_labels": labels3d}
    pred, _, loss = MaskTrainer.train_step(me, data)
    (grad,) = torch.autograd.grad(loss, logits)
    out[f"loss_{tag}"] = np.array([float(loss), reg_w, m3_w])
    out[f"loss_{tag}_grad"] = grad.numpy()
    out[f"loss_{tag}_pred"] = pred.numpy()
with torch.no_grad():
    m3_logits = model.mask(coords3d, geo_feat=model.density(coords3d)["geo_feat"])
out.update(loss_logits=logits.detach().numpy(), loss_depth=depth.numpy(), loss_labels=labels.numpy(), loss_patch=np.array(patch),
           m3_coords=coords3d.numpy(), m3_labels=labels3d.numpy(), m3_logits=m3_logits.numpy(),
           m3_loss=np.array(float(nn.functional.cross_entropy(m3_logits, labels3d))))
# all labels unlabelled -> CE branch returns 0 (utils.py:1313-1314)
me.opt.label_regularization_weight, me.opt.mask3d_loss_weight = 0.0, 0.0
_, _, loss0 = MaskTrainer.train_step(me, {"rays_o": None, "rays_d": None, "masks": torch.full_like(labels, -1)})
out["loss_unlabelled"] = np.array(float(loss0))

# ------------------------------------------------------------------------------------------ occupancy grid --
G = 16
model.grid_size = G
C = model.cascade
model.density_grid = torch.zeros(C, G ** 3)
model.density_bitfield = torch.zeros(C * G ** 3 // 8, dtype=torch.uint8)
cam = poses.clone()
cam[:, :3, 3] *= 0.25           # cameras inside the bound-2 volume
intr = synthetic.intrinsics(48, 64)
model.mark_untrained_grid(cam, intr)
out.update(occ_cfg=np.array([C, G, bound]), mark_poses=cam.numpy(), mark_intr=np.array(intr, np.float64), mark_grid=model.density_grid.numpy().copy())

rec = {}
_randint, _rand_like, _density = torch.randint, torch.rand_like, model.density


def randint(*a, **kw):
    v = _randint(*a, **kw)
    rec.setdefault("randint", []).append(v.numpy().copy())
    return v


def rand_like(*a, **kw):
    v = _rand_like(*a, **kw)
    rec.setdefault("noise", []).append(v.numpy().copy())
    return v


def density(x):
    r = _density(x)
    rec.setdefault("points", []).append(x.detach().numpy().copy())
    rec.setdefault("sigma", []).append(r["sigma"].detach().numpy().copy())
    return r


torch.randint, torch.rand_like, model.density = randint, rand_like, density
try:
    for tag, it in (("full", 0), ("partial", 16)):
        rec.clear()
        model.iter_density = it
        model.local_step = 0
        out[f"occ_{tag}_grid_in"] = model.density_grid.numpy().copy()
        torch.manual_seed(31 + it)
        with torch.no_grad():
            model.update_extra_state()
        out[f"occ_{tag}_noise"] = np.stack(rec["noise"])          # [C, n, 3] in the reference's draw order
        out[f"occ_{tag}_points"] = np.stack(rec["points"])        # [C, n, 3]
        out[f"occ_{tag}_sigma"] = np.stack(rec["sigma"])          # [C, n]
        if tag == "partial":
            out["occ_partial_coords"] = np.stack(rec["randint"][0::2])     # [C, N, 3] uniform cells
            out["occ_partial_picks"] = np.stack(rec["randint"][1::2])      # [C, N] positions in nonzero(density_grid > 0)
        out[f"occ_{tag}_grid_out"] = model.density_grid.numpy().copy()
        out[f"occ_{tag}_bits"] = model.density_bitfield.numpy().copy()
        out[f"occ_{tag}_mean"] = np.array(model.mean_density)
finally:
    torch.randint, torch.rand_like = _randint, _rand_like

path = os.path.join(ri.ROOT, "tests", "golden", "ref_host.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if k.startswith("occ_")})
