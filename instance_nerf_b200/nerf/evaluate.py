"""Evaluation / test loop of the instance stage (MaskTrainer.eval_step, test_step, test, evaluate_one_epoch: nerf/utils.py:1375-1431,
1435-1496, 1553-1640) on top of the one-launch renderer.

A test job is the c4 workload shape (many poses, whole frames): frames are sharded round-robin over the ranks
(`parallel.shard_frames`), each frame is rendered by `model.render(..., staged=True, render_mask=True)`, finalised ON THE DEVICE
(`inerf_frame_to_u8`: rgb / depth quantisation + instance argmax, 5 B / pixel) and copied to pinned host memory while the next
frame renders; the reference copies fp32 image + fp32 depth + int64 labels (24 B / pixel) and quantises with numpy.
PNG output goes through cv2 as in the reference; mp4 output needs imageio (not in this image: requested -> ImportError).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .._lib import call, ptr, stream_ptr
from ..parallel import shard_frames


def frame_to_u8(image, depth=None, logits=None):
    """image [N,3] (, depth [N], logits [N,K]) float32 CUDA -> (rgb uint8 [N,3], depth uint8 [N] | None, label uint8 [N] | None)."""
    if not image.is_cuda:
        raise RuntimeError("frame_to_u8: CUDA tensors only (no CPU fallback)")
    image = image.float().contiguous().view(-1, 3)
    N, dev = image.shape[0], image.device
    rgb = torch.empty(N, 3, dtype=torch.uint8, device=dev)
    d8 = lab = None
    K = 0
    if depth is not None:
        depth = depth.float().contiguous().view(-1)
        d8 = torch.empty(N, dtype=torch.uint8, device=dev)
    if logits is not None:
        K = logits.shape[-1]
        logits = logits.float().contiguous().view(-1, K)
        lab = torch.empty(N, dtype=torch.uint8, device=dev)
    call("inerf_frame_to_u8", ptr(image), ptr(depth), ptr(logits), N, K, ptr(rgb), ptr(d8), ptr(lab), stream_ptr(dev))
    return rgb, d8, lab


def instance_palette(n=256, seed=7):
    """Colours for the `_mask_rgb.png` visualisation.  The reference indexes matplotlib's `gist_ncar` (nerf/utils.py:1236-1238);
    matplotlib is not a dependency here, so this is a fixed pseudo-random palette (label 0 = black) -- visualisation only."""
    rng = np.random.default_rng(seed)
    pal = rng.integers(32, 256, size=(n, 3), dtype=np.uint8)
    pal[0] = 0
    return pal


class FrameWriter:
    """Double-buffered device -> pinned-host hand-off of finalised frames: the copy of frame i overlaps the render of frame i+1."""

    def __init__(self, n_pixels, device, depth=2):
        self.dev = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.slots = [dict(host=torch.empty(n_pixels, 5, dtype=torch.uint8, pin_memory=True),
                           dev=torch.empty(n_pixels, 5, dtype=torch.uint8, device=device), done=torch.cuda.Event(), tag=None) for _ in range(depth)]
        self.k = 0

    def push(self, tag, rgb, d8, lab):
        """Queue one frame; returns the (tag, host array [n, 5]) of the frame whose slot is being recycled, if any."""
        s = self.slots[self.k % len(self.slots)]
        self.k += 1
        ready = None
        if s["tag"] is not None:
            s["done"].synchronize()
            ready = (s["tag"], s["host"].numpy().copy())
        s["dev"][:, 0:3] = rgb
        s["dev"][:, 3] = d8
        s["dev"][:, 4] = lab
        self.copy_stream.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.copy_stream):
            s["host"].copy_(s["dev"], non_blocking=True)
            s["done"].record(self.copy_stream)
        s["tag"] = tag
        return ready

    def flush(self):
        out = []
        for j in range(len(self.slots)):
            s = self.slots[(self.k + j) % len(self.slots)]
            if s["tag"] is not None:
                s["done"].synchronize()
                out.append((s["tag"], s["host"].numpy().copy()))
                s["tag"] = None
        return out


@torch.no_grad()
def test(model, loader, save_path, name="ngp", fp16=True, write_video=False, rank=0, world=1, render_kw=None, palette=None):
    """MaskTrainer.test (nerf/utils.py:1435-1496): render every pose of `loader` (a provider's `dataloader()`, batch size 1), write
    `{name}_{i:04d}_rgb.png / _depth.png / _mask.png / _mask_rgb.png`.  Rank r renders frames r, r + world, ...; returns the
    frame indices this rank wrote."""
    import cv2
    if write_video:
        import imageio  # noqa: F401  (not installed here: the ImportError is the message)
        raise NotImplementedError("mp4 output: write PNGs and encode them outside the render path")
    os.makedirs(save_path, exist_ok=True)
    kw = dict(staged=True, render_mask=True, perturb=False, bg_color=None)
    kw.update(render_kw or {})
    palette = instance_palette() if palette is None else palette
    ds = loader._data
    mine = shard_frames(len(ds.poses), rank, world)
    was_training = model.training
    model.eval()
    writer = None

    def emit(item):
        (i, H, W), arr = item
        cv2.imwrite(os.path.join(save_path, f"{name}_{i:04d}_rgb.png"), cv2.cvtColor(arr[:, 0:3].reshape(H, W, 3), cv2.COLOR_RGB2BGR))
        cv2.imwrite(os.path.join(save_path, f"{name}_{i:04d}_depth.png"), arr[:, 3].reshape(H, W))
        cv2.imwrite(os.path.join(save_path, f"{name}_{i:04d}_mask.png"), arr[:, 4].reshape(H, W))
        cv2.imwrite(os.path.join(save_path, f"{name}_{i:04d}_mask_rgb.png"), cv2.cvtColor(palette[arr[:, 4]].reshape(H, W, 3), cv2.COLOR_RGB2BGR))

    for i in mine:
        data = ds.collate([i])
        H, W = data["H"], data["W"]
        with torch.autocast("cuda", dtype=torch.float16, enabled=fp16):
            out = model.render(data["rays_o"], data["rays_d"], **kw)
        rgb, d8, lab = frame_to_u8(out["image"][0], out["depth"][0], out["instance_mask_logits"][0])
        if writer is None:
            writer = FrameWriter(H * W, rgb.device)
        ready = writer.push((i, H, W), rgb, d8, lab)
        if ready is not None:
            emit(ready)
    if writer is not None:
        for item in writer.flush():
            emit(item)
    model.train(was_training)
    return list(mine)
