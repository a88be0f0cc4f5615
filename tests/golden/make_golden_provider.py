"""Generates tests/golden/ref_provider.npz: what the REFERENCE's own providers (nerf/provider.py, imported unmodified, CPU)
make of a small synthetic scene directory -- NeRFMaskDataset (instance maps + 3D-mask voxels; train / val / test splits) and
NeRFDataset (RGB frames, colmap layout) -- plus one collated training batch of each with the drawn pixel indices.

The scene directory is rebuilt by the test from the arrays stored in the same file (tests/provider_scene.py writes it for both
sides).  h5py is not installed here: instance maps are written as npy payloads under the reference's `.hdf5` names and a
stand-in `h5py.File` (provider_scene.FakeH5File) reads them; everything else that runs is the reference's code.

    python tests/golden/make_golden_provider.py
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

import _ref_import as ri  # noqa: F401  (sys.path, stubs)
import provider_scene as ps  # noqa: E402  (tests/provider_scene.py)

sys.modules["h5py"].File = ps.FakeH5File
import nerf.provider as ref_provider  # noqa: E402  (the reference's module)

ref_provider.h5py.File = ps.FakeH5File
ref_provider.tqdm.tqdm = lambda it, **kw: it

out = {}
scene = ps.make_scene(seed=0)
out.update({"scene_" + k: v for k, v in ps.pack_scene(scene).items()})

with tempfile.TemporaryDirectory() as root:
    ps.write_scene(scene, root)
    for split in ("train", "val", "test"):
        opt = ps.options(root, mask3d=os.path.join(root, "mask3d.npy"))
        ds = ref_provider.NeRFMaskDataset(opt, "cpu", type=split)
        tag = f"mask_{split}_"
        out[tag + "poses"] = ds.poses.numpy()
        out[tag + "intrinsics"] = np.asarray(ds.intrinsics, dtype=np.float64)
        out[tag + "meta"] = np.array([ds.H, ds.W, ds.num_instances, ds.num_rays], dtype=np.int64)
        out[tag + "radius"] = np.float64(ds.radius)
        out[tag + "offset"] = np.asarray(ds.offset, dtype=np.float64)
        if ds.masks is not None:
            out[tag + "masks"] = ds.masks.numpy()
        if ds.mask3d is not None:
            out[tag + "mask3d_coords"] = ds.mask3d_coords.numpy()
            out[tag + "mask3d_labels"] = ds.mask3d_labels.numpy()
        loader = ds.dataloader()
        out[tag + "loader"] = np.array([len(loader), int(loader.has_gt)], dtype=np.int64)
        if split == "train":
            torch.manual_seed(3)
            b = ds.collate([2])
            # the reference does not return the pixel indices for masks: recover them by re-drawing with the same seed
            torch.manual_seed(3)
            from nerf.utils import get_rays
            r = get_rays(ds.poses[[2]], ds.intrinsics, ds.H, ds.W, ds.num_rays, None, opt.patch_size)
            assert torch.equal(r["rays_o"], b["rays_o"])
            out[tag + "batch_inds"] = r["inds"].numpy()
            for k in ("rays_o", "rays_d", "masks"):
                out[tag + "batch_" + k] = b[k].numpy()
            out[tag + "batch_file_name"] = np.array(b["file_name"])
        elif split == "val":
            b = ds.collate([0])
            out[tag + "batch_rays_d"] = b["rays_d"].numpy()
            out[tag + "batch_masks"] = b["masks"].numpy()

    for split in ("train", "val"):
        opt = ps.options(root, mask3d=None, rgb=True)
        ds = ref_provider.NeRFDataset(opt, "cpu", type=split)
        tag = f"rgb_{split}_"
        out[tag + "poses"] = ds.poses.numpy()
        out[tag + "images"] = ds.images.numpy()
        out[tag + "intrinsics"] = np.asarray(ds.intrinsics, dtype=np.float64)
        out[tag + "radius"] = np.float64(ds.radius)
        if split == "train":
            torch.manual_seed(5)
            b = ds.collate([1])
            torch.manual_seed(5)
            from nerf.utils import get_rays
            r = get_rays(ds.poses[[1]], ds.intrinsics, ds.H, ds.W, ds.num_rays, None, opt.patch_size)
            out[tag + "batch_inds"] = r["inds"].numpy()
            out[tag + "batch_rays_d"] = b["rays_d"].numpy()
            out[tag + "batch_images"] = b["images"].numpy()

# pose conversion + orbit poses (pure functions)
rng = np.random.default_rng(1)
P = rng.normal(size=(4, 4)).astype(np.float32)
out["ngp_in"] = P
out["ngp_out"] = ref_provider.nerf_matrix_to_ngp(P, scale=0.4, offset=[0.1, -0.2, 0.3])
torch.manual_seed(9)
out["rand_poses"] = ref_provider.rand_poses(6, "cpu", radius=2.5).numpy()

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_provider.npz")
np.savez_compressed(path, **out)
print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")
