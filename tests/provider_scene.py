"""A small synthetic scene directory in the layout the reference's providers read (transforms.json, per-frame instance maps,
RGB frames, a 3D-mask volume) -- shared by tests/golden/make_golden_provider.py (reference side) and tests/test_provider.py."""
import json
import os
from types import SimpleNamespace

import numpy as np

H, W, N_FRAMES, K = 24, 32, 4, 5


class FakeH5File:
    """Stand-in for h5py.File (h5py is not installed in this image): the `.hdf5` files of the test scene hold an npy payload
    with the `cp_instance_id_segmaps` array."""

    def __init__(self, path, mode="r"):
        with open(path, "rb") as f:
            self._d = {"cp_instance_id_segmaps": np.load(f)}

    def __enter__(self):
        return self._d

    def __exit__(self, *a):
        return False


def make_scene(seed=0):
    rng = np.random.default_rng(seed)
    frames, masks, images = [], [], []
    for i in range(N_FRAMES):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        T = np.eye(4)
        T[:3, :3] = q
        T[:3, 3] = rng.normal(size=3) * 2
        frames.append(T)
        m = np.zeros((H, W), np.uint8)
        for k in range(1, K + 1):
            y, x = rng.integers(0, H - 6), rng.integers(0, W - 8)
            m[y:y + rng.integers(3, 7), x:x + rng.integers(3, 9)] = k
        masks.append(m)
        images.append(rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8))
    vol = np.zeros((6, 5, 7), np.int64)
    for k in range(1, K + 1):
        vol[rng.integers(0, 5), rng.integers(0, 4), rng.integers(0, 6)] = k
    vol[1:3, 1:3, 2:4] = 2
    bbox = np.array([[-3.0, -1.5, -2.5], [3.5, 1.3, 2.0]])
    return dict(frames=np.stack(frames), masks=np.stack(masks), images=np.stack(images), vol=vol, bbox=bbox)


def pack_scene(s):
    return {k: np.asarray(v) for k, v in s.items()}


def unpack_scene(npz, prefix="scene_"):
    return {k[len(prefix):]: npz[k] for k in npz.files if k.startswith(prefix)}


def write_scene(s, root):
    import cv2
    os.makedirs(os.path.join(root, "segmaps"), exist_ok=True)
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    frames = []
    for i in range(len(s["frames"])):
        rel = f"./segmaps/{i:04d}.hdf5"
        with open(os.path.join(root, rel), "wb") as f:
            np.save(f, s["masks"][i])
        cv2.imwrite(os.path.join(root, f"images/{i:04d}.png"), s["images"][i][..., ::-1])   # stored BGR -> read back as RGB
        frames.append({"file_path": rel, "transform_matrix": np.asarray(s["frames"][i]).tolist()})
    base = {"h": H, "w": W, "fl_x": 30.0, "fl_y": 31.0, "cx": W / 2 - 0.5, "cy": H / 2 + 0.25, "num_room_objects": K, "num_instances": K,
            "room_bbox": np.asarray(s["bbox"]).tolist()}
    with open(os.path.join(root, "transforms.json"), "w") as f:
        json.dump(dict(base, frames=frames), f)
    np.save(os.path.join(root, "mask3d.npy"), s["vol"])
    # RGB layout (colmap mode) in a sibling directory: same poses, png frames
    rgb = os.path.join(root, "rgb")
    os.makedirs(rgb, exist_ok=True)
    rgb_frames = [{"file_path": f"../images/{i:04d}.png", "transform_matrix": np.asarray(s["frames"][i]).tolist()} for i in range(len(s["frames"]))]
    with open(os.path.join(rgb, "transforms.json"), "w") as f:
        json.dump(dict(base, frames=rgb_frames), f)


def options(root, mask3d, rgb=False):
    return SimpleNamespace(path=os.path.join(root, "rgb") if rgb else root, preload=True, scale=0.33, offset=[0, 0, 0], bound=8.0, fp16=False,
                           num_rays=128, rand_pose=-1, mask3d=mask3d, error_map=False, patch_size=8, color_space="srgb")
