"""Scene providers behind the reference's dataset API (nerf/provider.py:21-31, 60-94, 97-347 NeRFDataset, 350-638
NeRFMaskDataset): the `transforms.json` scene description, camera poses in the NGP frame, RGB frames (stage 1) or per-frame
instance-id maps plus the optional 3D-mask voxels (instance stage), and `collate` -- the function that turns a frame index
into the ray batch `Trainer.train_step` / `MaskTrainer.train_step` consume.

What differs from the reference: the frames are staged to the GPU once and every batch is built there -- poses (64 B) ->
`inerf_get_rays` (one launch; nerf/utils.py here) -> one gather of the labels / pixels at the drawn indices; the reference
builds an H*W meshgrid per batch and, unless `--preload`, copies the whole frame host->device first.  Same attribute names
(`poses, masks / images, intrinsics, H, W, radius, num_instances, offset, error_map, mask3d_coords, mask3d_labels`), same
batch keys, same `dataloader()` contract (`loader._data`, `loader.has_gt`).

Frame files: instance-id maps are read from `.hdf5` (dataset `cp_instance_id_segmaps`, needs h5py -- not in this image: the
import error is raised as is), `.npy`, `.npz` (same key) or `.png` (first channel; the branch the reference intends at
provider.py:522-523 but mis-spells).  RGB frames go through cv2 exactly as provider.py:223-238.
"""
from __future__ import annotations

import glob
import json
import os

import numpy as np
import torch
from torch.utils.data import DataLoader

from .utils import get_rays

SEGMAP_KEY = "cp_instance_id_segmaps"


def nerf_matrix_to_ngp(pose, scale=0.33, offset=(0, 0, 0)):
    """provider.py:21-31: camera-to-world of the NeRF convention -> the NGP frame (axes cycled y, z, x; y and z columns
    negated; translation scaled and offset)."""
    pose = np.asarray(pose)
    out = np.empty((4, 4), dtype=np.float32)
    perm = (1, 2, 0)
    for r, src in enumerate(perm):
        out[r, 0] = pose[src, 0]
        out[r, 1] = -pose[src, 1]
        out[r, 2] = -pose[src, 2]
        out[r, 3] = pose[src, 3] * scale + offset[r]
    out[3] = (0, 0, 0, 1)
    return out


def rand_poses(size, device, radius=1, theta_range=(np.pi / 3, 2 * np.pi / 3), phi_range=(0, 2 * np.pi), generator=None):
    """provider.py:60-94: `size` orbit cameras at `radius` looking at the origin.  -> [size, 4, 4] float32 on `device`."""
    def unit(v):
        return v / (torch.norm(v, dim=-1, keepdim=True) + 1e-10)

    thetas = torch.rand(size, device=device, generator=generator) * (theta_range[1] - theta_range[0]) + theta_range[0]
    phis = torch.rand(size, device=device, generator=generator) * (phi_range[1] - phi_range[0]) + phi_range[0]
    centers = torch.stack([radius * torch.sin(thetas) * torch.sin(phis), radius * torch.cos(thetas),
                           radius * torch.sin(thetas) * torch.cos(phis)], dim=-1)
    forward = -unit(centers)
    up = torch.tensor([0.0, -1.0, 0.0], device=device).expand(size, 3)
    right = unit(torch.cross(forward, up, dim=-1))
    up = unit(torch.cross(right, forward, dim=-1))
    poses = torch.eye(4, dtype=torch.float32, device=device).repeat(size, 1, 1)
    poses[:, :3, :3] = torch.stack((right, up, forward), dim=-1)
    poses[:, :3, 3] = centers
    return poses


def read_segmap(path):
    """One frame's instance-id map [H, W] (integer)."""
    low = path.lower()
    if low.endswith("hdf5") or low.endswith(".h5"):
        import h5py   # absent in this image: the ImportError is the error message
        with h5py.File(path, "r") as f:
            return np.array(f[SEGMAP_KEY][:])
    if low.endswith(".npy"):
        return np.load(path)
    if low.endswith(".npz"):
        return np.load(path)[SEGMAP_KEY]
    if low.endswith(".png"):
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is None:
            raise RuntimeError(f"cannot read {path}")
        return img[..., 0] if img.ndim == 3 else img
    raise RuntimeError(f"unknown instance-map format: {path}")


def _intrinsics(transform, H, W, downscale):
    """provider.py:322-338 / 564-580 -> np.array([fl_x, fl_y, cx, cy])"""
    if "fl_x" in transform or "fl_y" in transform:
        fl_x = (transform["fl_x"] if "fl_x" in transform else transform["fl_y"]) / downscale
        fl_y = (transform["fl_y"] if "fl_y" in transform else transform["fl_x"]) / downscale
    elif "camera_angle_x" in transform or "camera_angle_y" in transform:
        fl_x = W / (2 * np.tan(transform["camera_angle_x"] / 2)) if "camera_angle_x" in transform else None
        fl_y = H / (2 * np.tan(transform["camera_angle_y"] / 2)) if "camera_angle_y" in transform else None
        fl_x = fl_y if fl_x is None else fl_x
        fl_y = fl_x if fl_y is None else fl_y
    else:
        raise RuntimeError("Failed to load focal length, please check the transforms.json!")
    cx = (transform["cx"] / downscale) if "cx" in transform else (W / 2)
    cy = (transform["cy"] / downscale) if "cy" in transform else (H / 2)
    return np.array([fl_x, fl_y, cx, cy])


class _SceneDataset:
    """What the two datasets share: option plumbing, scene frame (scale / offset from `room_bbox`), pose conversion, the
    error map, the random-pose branch and the batch assembly on the device."""

    payload_name = None   # 'images' / 'masks'

    def _init_common(self, opt, device, type, downscale):
        self.opt = opt
        self.device = torch.device(device)
        self.type = type
        self.downscale = downscale
        self.root_path = opt.path
        self.preload = opt.preload
        self.scale = opt.scale
        self.offset = opt.offset
        self.bound = opt.bound
        self.fp16 = opt.fp16
        self.training = type in ("train", "all", "trainval")
        self.num_rays = opt.num_rays if self.training else -1
        self.rand_pose = opt.rand_pose

    def _scene_frame(self, transform):
        if "h" in transform and "w" in transform:
            self.H, self.W = int(transform["h"]) // self.downscale, int(transform["w"]) // self.downscale
        else:
            self.H = self.W = None   # taken from the first frame read
        self.room_bbox = None
        if "room_bbox" in transform:
            self.room_bbox = np.array(transform["room_bbox"])
            self.offset = -(self.room_bbox[0] + self.room_bbox[1]) * 0.5 * self.scale

    def _pose(self, frame):
        return nerf_matrix_to_ngp(np.array(frame["transform_matrix"], dtype=np.float32), scale=self.scale, offset=self.offset)

    def _finish(self, transform, poses, payload):
        self.poses = torch.from_numpy(np.stack(poses, axis=0))
        setattr(self, self.payload_name, None if payload is None else torch.from_numpy(np.stack(payload, axis=0)))
        self.radius = self.poses[:, :3, 3].norm(dim=-1).mean(0).item()
        n = 0 if payload is None else len(payload)
        self.error_map = torch.ones([n, 128 * 128], dtype=torch.float) if (self.training and self.opt.error_map) else None
        self.intrinsics = _intrinsics(transform, self.H, self.W, self.downscale)

    def _stage(self, payload_dtype):
        """Poses, frames and the error map live on the device from here on (180 GB of HBM: a 3D-FRONT scene is a few hundred
        MB); `collate` never touches host memory.  `preload=False` keeps the reference's host-resident behaviour."""
        if not self.preload:
            return
        self.poses = self.poses.to(self.device)
        payload = getattr(self, self.payload_name)
        if payload is not None:
            setattr(self, self.payload_name, payload.to(payload_dtype).to(self.device))
        if self.error_map is not None:
            self.error_map = self.error_map.to(self.device)

    # ---- batches ------------------------------------------------------------------------------------------------------
    def _rand_pose_batch(self, B):
        """provider.py:347-362 / 591-606: a low-resolution full frame from a random orbit pose (no ground truth)."""
        poses = rand_poses(B, self.device, radius=self.radius)
        s = np.sqrt(self.H * self.W / self.num_rays)
        rH, rW = int(self.H / s), int(self.W / s)
        rays = get_rays(poses, self.intrinsics / s, rH, rW, -1)
        return {"H": rH, "W": rW, "rays_o": rays["rays_o"], "rays_d": rays["rays_d"]}

    def _ray_batch(self, index, inds=None):
        poses = self.poses[index].to(self.device)
        error_map = None if self.error_map is None else self.error_map[index]
        rays = get_rays(poses, self.intrinsics, self.H, self.W, self.num_rays, error_map, self.opt.patch_size, inds=inds)
        results = {"H": self.H, "W": self.W, "rays_o": rays["rays_o"], "rays_d": rays["rays_d"]}
        return results, rays, error_map

    def _loader(self, has_gt):
        size = len(self.poses)
        if self.training and self.rand_pose > 0:
            size += size // self.rand_pose   # indices past the frames select the random-pose branch
        loader = DataLoader(list(range(size)), batch_size=1, collate_fn=self.collate, shuffle=self.training, num_workers=0)
        loader._data = self
        loader.has_gt = has_gt
        return loader


class NeRFDataset(_SceneDataset):
    """Stage-1 provider (provider.py:97-347): RGB(A) frames."""

    payload_name = "images"

    def __init__(self, opt, device, type="train", downscale=1, n_test=10, n_test_per_pose=2):
        self._init_common(opt, device, type, downscale)
        root = self.root_path
        if os.path.exists(os.path.join(root, "transforms.json")):
            self.mode = "colmap"    # one file, first frame held out for validation
        elif os.path.exists(os.path.join(root, "transforms_train.json")):
            self.mode = "blender"   # one file per split
        else:
            raise NotImplementedError(f"[NeRFDataset] Cannot find transforms*.json under {root}")
        transform = self._load_transform(type)
        self._scene_frame(transform)
        frames = transform["frames"]
        poses, images = [], []
        if self.mode == "colmap" and type == "test":
            images = None
            poses = [self._pose(f) for f in frames]
        else:
            if self.mode == "colmap":
                frames = frames[1:] if type == "train" else (frames[:1] if type == "val" else frames)
            import cv2
            for f in frames:
                f_path = os.path.join(root, f["file_path"])
                if self.mode == "blender" and "." not in os.path.basename(f_path):
                    f_path += ".png"
                if not os.path.exists(f_path):
                    continue
                image = cv2.imread(f_path, cv2.IMREAD_UNCHANGED)
                if self.H is None or self.W is None:
                    self.H, self.W = image.shape[0] // downscale, image.shape[1] // downscale
                image = cv2.cvtColor(image, cv2.COLOR_BGR2RGB if image.shape[-1] == 3 else cv2.COLOR_BGRA2RGBA)
                if image.shape[0] != self.H or image.shape[1] != self.W:
                    image = cv2.resize(image, (self.W, self.H), interpolation=cv2.INTER_AREA)
                poses.append(self._pose(f))
                images.append(image.astype(np.float32) / 255)
        self._finish(transform, poses, images)
        self._stage(torch.half if (self.fp16 and getattr(opt, "color_space", "srgb") != "linear") else torch.float)

    def _load_transform(self, type):
        root = self.root_path

        def load(name):
            with open(os.path.join(root, name), "r") as f:
                return json.load(f)

        if self.mode == "colmap":
            return load("transforms.json")
        if type == "all":
            transform = None
            for path in glob.glob(os.path.join(root, "*.json")):
                t = load(os.path.basename(path))
                if transform is None:
                    transform = t
                else:
                    transform["frames"].extend(t["frames"])
            return transform
        if type == "trainval":
            transform = load("transforms_train.json")
            transform["frames"].extend(load("transforms_val.json")["frames"])
            return transform
        return load(f"transforms_{type}.json")

    def collate(self, index, inds=None):
        B = len(index)
        if self.rand_pose == 0 or index[0] >= len(self.poses):
            return self._rand_pose_batch(B)
        results, rays, error_map = self._ray_batch(index, inds)
        if self.images is not None:
            images = self.images[index].to(self.device)
            if self.training:
                C = images.shape[-1]
                images = torch.gather(images.view(B, -1, C), 1, rays["inds"][..., None].expand(-1, -1, C))
            results["images"] = images
        if error_map is not None:
            results["index"] = index
            results["inds_coarse"] = rays["inds_coarse"]
        return results

    def dataloader(self):
        return self._loader(self.images is not None)


class NeRFMaskDataset(_SceneDataset):
    """Instance-stage provider (provider.py:350-638): per-frame instance-id maps (+ the 3D-mask voxel constraints)."""

    payload_name = "masks"

    def __init__(self, opt, device, type="train", downscale=1, n_test_per_pose=10, n_test_poses=10):
        self._init_common(opt, device, type, downscale)
        self.mask3d = opt.mask3d if (self.training or type == "val") else None
        with open(os.path.join(self.root_path, "transforms.json"), "r") as f:
            transform = json.load(f)
        self._scene_frame(transform)
        if "num_room_objects" not in transform:
            raise RuntimeError("Failed to load number of instances, please check the transforms.json!")
        self.num_instances = transform["num_instances"] + 1   # + background
        if self.mask3d is not None:
            self._load_mask3d(self.mask3d)
        frames = transform["frames"]
        self.frames = frames
        poses, masks = [], []
        if type == "test":
            masks = None
            poses = [self._pose(f) for f in frames]
        else:
            if type == "val":
                frames = frames[:1]
                self.frames = self.frames[:1]
            for f in frames:
                f_path = os.path.join(self.root_path, f["file_path"])
                if not os.path.exists(f_path):
                    print(f"Warning: {f_path} does not exist, skipping.")
                    continue
                mask = read_segmap(f_path)
                if mask.max() >= self.num_instances:
                    raise AssertionError(f"Instance id {mask.max()} exceeds the number of instances {self.num_instances - 1}")
                if self.H is None or self.W is None:
                    self.H, self.W = mask.shape[0] // downscale, mask.shape[1] // downscale
                if mask.shape[0] != self.H or mask.shape[1] != self.W:
                    import cv2
                    mask = cv2.resize(mask, (self.W, self.H), interpolation=cv2.INTER_AREA)
                poses.append(self._pose(f))
                masks.append(mask)
        self._finish(transform, poses, masks)
        self._stage(torch.long)

    def _load_mask3d(self, path):
        """provider.py:448-473: labelled voxels of the 3D mask volume -> (coords [M, 3] in the NGP-scaled scene frame, labels [M])."""
        if self.room_bbox is None:
            raise AssertionError("3d mask requires room_bbox in transforms.json!")
        vol = np.load(path)
        if vol.max() >= self.num_instances:
            raise RuntimeError(f"3d mask has too many instances {vol.max()}, only {self.num_instances - 1} instances are loaded!")
        if vol.ndim != 3:
            raise AssertionError(f"3d mask should be [W, L, H], got {vol.shape}")
        axes = [np.linspace(self.room_bbox[0][a], self.room_bbox[1][a], vol.shape[a]) * self.scale + self.offset[a] for a in range(3)]
        keep = np.flatnonzero(vol.reshape(-1) > 0)
        ijk = np.unravel_index(keep, vol.shape)
        coords = np.stack([axes[a][ijk[a]] for a in range(3)], -1)   # only the labelled voxels are ever materialised
        self.mask3d_labels = torch.from_numpy(vol.reshape(-1)[keep]).to(torch.long).to(self.device)
        self.mask3d_coords = torch.from_numpy(coords).to(torch.float).to(self.device)

    def collate(self, index, inds=None):
        B = len(index)
        if self.rand_pose == 0 or index[0] >= len(self.poses):
            return self._rand_pose_batch(B)
        results, rays, error_map = self._ray_batch(index, inds)
        results["file_name"] = self.frames[index[0]]["file_path"][9:-5]
        if self.masks is not None:
            masks = self.masks[index].to(self.device)
            if self.training:
                masks = torch.gather(masks.view(B, -1), 1, rays["inds"])
            results["masks"] = masks
        if self.mask3d is not None:
            results["mask3d_coords"] = self.mask3d_coords
            results["mask3d_labels"] = self.mask3d_labels
        if error_map is not None:
            results["index"] = index
            results["inds_coarse"] = rays["inds_coarse"]
        return results

    def dataloader(self):
        return self._loader(self.masks is not None)
