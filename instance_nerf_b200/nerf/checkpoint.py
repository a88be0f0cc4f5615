"""Checkpoint interop with the reference trainer (Trainer.save_checkpoint / load_checkpoint, nerf/utils.py:1100-1228).

The file layout is the reference's: `torch.save` of a dict with `epoch`, `global_step`, `stats`, `model` (the state dict; the
networks here keep the reference's parameter / buffer names and shapes, so it loads either way round), `mean_count` /
`mean_density` for cuda_ray models and, for `full=True`, `optimizer` / `lr_scheduler` / `scaler`.  A bare state dict (no
`model` key) is accepted as the reference accepts it (:1169-1172).  "Best" checkpoints of the reference drop `density_grid`
(:1150-1152): loading is non-strict for exactly that key and reports everything else.
"""
from __future__ import annotations

import torch

from .._lib import invalidate_param_caches


def save_checkpoint(path, model, epoch=0, global_step=0, stats=None, optimizer=None, lr_scheduler=None, scaler=None, full=False, best=False):
    state = {"epoch": epoch, "global_step": global_step, "stats": stats if stats is not None else {}}
    if getattr(model, "cuda_ray", False):
        state["mean_count"] = model.mean_count          # host ints / floats, as the reference stores them
        state["mean_density"] = model.mean_density
    if full:
        if optimizer is not None:
            state["optimizer"] = optimizer.state_dict()
        if lr_scheduler is not None:
            state["lr_scheduler"] = lr_scheduler.state_dict()
        if scaler is not None:
            state["scaler"] = scaler.state_dict()
    sd = model.state_dict()
    if best and "density_grid" in sd:
        sd = {k: v for k, v in sd.items() if k != "density_grid"}
    state["model"] = sd
    torch.save(state, path)
    return state


def load_checkpoint(path, model, model_only=False, optimizer=None, lr_scheduler=None, scaler=None, map_location=None):
    """-> dict(missing_keys, unexpected_keys, epoch, global_step, stats).  Raises on shape mismatches (as load_state_dict does)."""
    ck = torch.load(path, map_location=map_location if map_location is not None else next(model.parameters()).device, weights_only=False)
    out = {"missing_keys": [], "unexpected_keys": [], "epoch": None, "global_step": None, "stats": None}
    if "model" not in ck:
        model.load_state_dict(ck)
        invalidate_param_caches()
        return out
    res = model.load_state_dict(ck["model"], strict=False)
    out["missing_keys"], out["unexpected_keys"] = list(res.missing_keys), list(res.unexpected_keys)
    invalidate_param_caches()
    if getattr(model, "cuda_ray", False):
        if "mean_count" in ck:
            model.mean_count = ck["mean_count"]
        if "mean_density" in ck:
            model.mean_density = ck["mean_density"]
    if model_only:
        return out
    out.update(epoch=ck.get("epoch"), global_step=ck.get("global_step"), stats=ck.get("stats"))
    for name, obj in (("optimizer", optimizer), ("lr_scheduler", lr_scheduler), ("scaler", scaler)):
        if obj is not None and name in ck:
            obj.load_state_dict(ck[name])
    return out
