"""Host side of the raymarching ops: the reference's module functions
(raymarching/raymarching.py:19-474) re-implemented over libinerf_b200's C ABI.

Same names, argument order, defaults, dtypes and return values, so
`import instance_nerf_b200.raymarching as raymarching` drops into
nerf/renderer.py / nerf/mask_renderer.py.  Differences, all deliberate:

* kernels run on torch's CURRENT stream (the reference uses the legacy default
  stream, raymarching.cu:154);
* `march_rays_train` is deterministic (count -> scan -> write): `rays` is sorted
  by ray id and the sample buffers are sized from the counted total instead of
  zero-filling N*max_steps rows (raymarching.py:205-207) and slicing;
* non-zero return codes raise RuntimeError (the reference checks nothing).
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .._lib import call, lib, ptr, stream_ptr

__all__ = [
    "near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits",
    "march_rays_train", "composite_rays_train", "composite_rays_with_masks_train",
    "march_rays", "composite_rays", "composite_rays_with_masks", "compact_alive", "count_samples",
]


def _f32(t: torch.Tensor) -> torch.Tensor:
    """custom_fwd(cast_inputs=torch.float32) of the reference: fp32, CUDA, contiguous."""
    if not t.is_cuda:
        t = t.cuda()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """raymarching.py:19-49 -> (nears [N], fars [N])"""
    rays_o = _f32(rays_o).view(-1, 3)
    rays_d = _f32(rays_d).view(-1, 3)
    aabb = _f32(aabb)
    N = rays_o.shape[0]
    nears = torch.empty(N, dtype=torch.float32, device=rays_o.device)
    fars = torch.empty(N, dtype=torch.float32, device=rays_o.device)
    call("inerf_near_far_from_aabb", ptr(rays_o), ptr(rays_d), ptr(aabb), N, float(min_near), ptr(nears), ptr(fars),
         stream_ptr(rays_o.device))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    """raymarching.py:52-80 -> coords [N, 2]"""
    rays_o = _f32(rays_o).view(-1, 3)
    rays_d = _f32(rays_d).view(-1, 3)
    N = rays_o.shape[0]
    coords = torch.empty(N, 2, dtype=torch.float32, device=rays_o.device)
    call("inerf_sph_from_ray", ptr(rays_o), ptr(rays_d), float(radius), N, ptr(coords), stream_ptr(rays_o.device))
    return coords


def morton3D(coords):
    """raymarching.py:83-104: int32 [N, 3] in [0, 128) -> int32 [N]"""
    if not coords.is_cuda:
        coords = coords.cuda()
    coords = coords.int().contiguous()
    N = coords.shape[0]
    indices = torch.empty(N, dtype=torch.int32, device=coords.device)
    call("inerf_morton3D", ptr(coords), N, ptr(indices), stream_ptr(coords.device))
    return indices


def morton3D_invert(indices):
    """raymarching.py:106-126: int32 [N] -> int32 [N, 3]"""
    if not indices.is_cuda:
        indices = indices.cuda()
    indices = indices.int().contiguous()
    N = indices.shape[0]
    coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
    call("inerf_morton3D_invert", ptr(indices), N, ptr(coords), stream_ptr(indices.device))
    return coords


def packbits(grid, thresh, bitfield=None):
    """raymarching.py:129-155: float [C, H^3] -> uint8 [C*H^3/8]"""
    grid = _f32(grid)
    N = grid.shape[0] * grid.shape[1] // 8
    if bitfield is None:
        bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
    call("inerf_packbits", ptr(grid), N, float(thresh), ptr(bitfield), stream_ptr(grid.device))
    return bitfield


def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                     perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024, noises=None, zero_fill=True):
    """raymarching.py:161-235 -> (xyzs [M,3], dirs [M,3], deltas [M,2], rays int32 [N,3]).

    `noises` (extra, optional) injects the per-ray jitter in [0,1) instead of
    drawing torch.rand(N) (raymarching.py:213-216); parity tests use it."""
    rays_o = _f32(rays_o).view(-1, 3)
    rays_d = _f32(rays_d).view(-1, 3)
    nears, fars = _f32(nears), _f32(fars)
    if not density_bitfield.is_cuda:
        density_bitfield = density_bitfield.cuda()
    density_bitfield = density_bitfield.contiguous()
    dev = rays_o.device
    N = rays_o.shape[0]
    st = stream_ptr(dev)

    if step_counter is None:
        step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
    if noises is None:
        noises = torch.rand(N, dtype=torch.float32, device=dev) if perturb else torch.zeros(N, dtype=torch.float32, device=dev)
    else:
        noises = _f32(noises)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)

    # the count pass records t per sample ([N, max_steps] scratch, never zero-filled) and the write pass becomes a parallel
    # expansion (one warp per ray, coalesced stores) instead of a second walk
    t_scratch = torch.empty(max(1, int(lib().inerf_march_scratch_floats(N, int(max_steps)))), dtype=torch.float32, device=dev)
    call("inerf_march_rays_train_count_t", ptr(rays_o), ptr(rays_d), ptr(density_bitfield), float(bound), float(dt_gamma),
         int(max_steps), N, int(C), int(H), ptr(nears), ptr(fars), ptr(rays), ptr(step_counter), ptr(noises), ptr(t_scratch), st)

    if not force_all_rays and mean_count > 0:
        # budgeted mode: M fixed from the running mean, no host sync; rays past the budget are dropped
        if align > 0:
            mean_count += align - mean_count % align
        M = int(mean_count)
    else:
        M = int(step_counter[0].item())  # D2H sync, as raymarching.py:224
        if align > 0:
            M += align - M % align       # a full extra `align` rows when already aligned (reference quirk)
    # rows past the last sample must be zero: they are fed through the network (padding) -- unless the caller bounds every
    # consumer by the device-side total (zero_fill=False: the CUDA-graph training step, inerf_field_desc.n_valid)
    alloc = torch.zeros if zero_fill else torch.empty
    xyzs = alloc(M, 3, dtype=torch.float32, device=dev)
    dirs = alloc(M, 3, dtype=torch.float32, device=dev)
    deltas = alloc(M, 2, dtype=torch.float32, device=dev)
    call("inerf_march_rays_train_expand", ptr(rays_o), ptr(rays_d), float(bound), float(dt_gamma), int(max_steps), N, int(C), int(H),
         M, ptr(nears), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(rays), ptr(noises), ptr(t_scratch), st)
    return xyzs, dirs, deltas, rays


def count_samples(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter, dt_gamma=0, max_steps=1024, t_scratch=None):
    """The count pass of march_rays_train alone (extra; no host sync): adds the number of samples these rays march to
    `step_counter[0]` (int32 [2], device) -- the cost estimate `parallel.balance_frames` plans multi-view jobs with.
    Returns the scratch buffer so that callers can reuse it."""
    rays_o = _f32(rays_o).view(-1, 3)
    rays_d = _f32(rays_d).view(-1, 3)
    nears, fars = _f32(nears).view(-1), _f32(fars).view(-1)
    dev = rays_o.device
    N = rays_o.shape[0]
    need = max(1, int(lib().inerf_march_scratch_floats(N, int(max_steps))))
    if t_scratch is None or t_scratch.numel() < need:
        t_scratch = torch.empty(need, dtype=torch.float32, device=dev)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    noises = torch.zeros(N, dtype=torch.float32, device=dev)
    call("inerf_march_rays_train_count_t", ptr(rays_o), ptr(rays_d), ptr(density_bitfield.contiguous()), float(bound), float(dt_gamma),
         int(max_steps), N, int(C), int(H), ptr(nears), ptr(fars), ptr(rays), ptr(step_counter), ptr(noises), ptr(t_scratch), stream_ptr(dev))
    return t_scratch


class _composite_rays_train(Function):
    """raymarching.py:238-291"""

    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4, dense=False):
        ctx.dense = bool(dense)
        sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
        rays = rays.contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        weights_sum = torch.empty(N, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        image = torch.empty(N, 3, dtype=torch.float32, device=dev)
        call("inerf_composite_rays_train_forward", ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(rays), M, N, float(T_thresh),
             ptr(weights_sum), ptr(depth), ptr(image), stream_ptr(dev))
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, image)
        ctx.dims = (M, N, T_thresh)
        return weights_sum, depth, image

    @staticmethod
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        # grad_depth is dropped, as in the reference (raymarching.py:275)
        sigmas, rgbs, deltas, rays, weights_sum, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        grad_weights_sum, grad_image = _f32(grad_weights_sum), _f32(grad_image)
        if ctx.dense:   # every row belongs to a ray: the kernel writes the zeros behind a ray's termination itself
            grad_sigmas = torch.empty_like(sigmas) if ctx.needs_input_grad[0] else None
            grad_rgbs = torch.empty_like(rgbs) if ctx.needs_input_grad[1] else None
            if grad_sigmas is not None or grad_rgbs is not None:
                call("inerf_composite_rays_with_masks_train_backward_dense", ptr(grad_weights_sum), ptr(grad_image), None, ptr(sigmas),
                     ptr(rgbs), None, ptr(deltas), ptr(rays), ptr(weights_sum), ptr(image), None, M, N, 0, float(T_thresh),
                     ptr(grad_sigmas), ptr(grad_rgbs), None, stream_ptr(sigmas.device))
            return grad_sigmas, grad_rgbs, None, None, None, None
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        call("inerf_composite_rays_train_backward", ptr(grad_weights_sum), ptr(grad_image), ptr(sigmas), ptr(rgbs), ptr(deltas),
             ptr(rays), ptr(weights_sum), ptr(image), M, N, float(T_thresh), ptr(grad_sigmas), ptr(grad_rgbs),
             stream_ptr(sigmas.device))
        return grad_sigmas, grad_rgbs, None, None, None, None


def composite_rays_train(sigmas, rgbs, deltas, rays, T_thresh=1e-4, dense=False):
    """`dense=True` (not in the reference): the caller guarantees that every row of the sample stream belongs to a ray of `rays`
    (the count -> scan -> expand marcher's output), so the backward needs no zero-filled gradient buffers and skips the
    gradients nobody asked for."""
    return _composite_rays_train.apply(sigmas, rgbs, deltas, rays, T_thresh, dense)


class _composite_rays_with_masks_train(Function):
    """raymarching.py:297-364"""

    @staticmethod
    def forward(ctx, sigmas, rgbs, masks, deltas, rays, T_thresh=1e-4, dense=False):
        ctx.dense = bool(dense) and masks.shape[1] <= 64
        sigmas, rgbs, masks, deltas = _f32(sigmas), _f32(rgbs), _f32(masks), _f32(deltas)
        rays = rays.contiguous()
        M, N, K = sigmas.shape[0], rays.shape[0], masks.shape[1]
        dev = sigmas.device
        weights_sum = torch.empty(N, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        image = torch.empty(N, 3, dtype=torch.float32, device=dev)
        mask_out = torch.empty(N, K, dtype=torch.float32, device=dev)
        call("inerf_composite_rays_with_masks_train_forward", ptr(sigmas), ptr(rgbs), ptr(masks), ptr(deltas), ptr(rays),
             M, N, K, float(T_thresh), ptr(weights_sum), ptr(depth), ptr(image), ptr(mask_out), stream_ptr(dev))
        ctx.save_for_backward(sigmas, rgbs, masks, deltas, rays, weights_sum, image, mask_out)
        ctx.dims = (M, N, K, T_thresh)
        return weights_sum, depth, image, mask_out

    @staticmethod
    def backward(ctx, grad_weights_sum, grad_depth, grad_image, grad_mask_out):
        sigmas, rgbs, masks, deltas, rays, weights_sum, image, mask_out = ctx.saved_tensors
        M, N, K, T_thresh = ctx.dims
        grad_weights_sum, grad_image, grad_mask_out = _f32(grad_weights_sum), _f32(grad_image), _f32(grad_mask_out)
        if ctx.dense:
            grad_sigmas = torch.empty_like(sigmas) if ctx.needs_input_grad[0] else None
            grad_rgbs = torch.empty_like(rgbs) if ctx.needs_input_grad[1] else None
            grad_masks = torch.empty_like(masks)
            call("inerf_composite_rays_with_masks_train_backward_dense", ptr(grad_weights_sum), ptr(grad_image), ptr(grad_mask_out),
                 ptr(sigmas), ptr(rgbs), ptr(masks), ptr(deltas), ptr(rays), ptr(weights_sum), ptr(image), ptr(mask_out),
                 M, N, K, float(T_thresh), ptr(grad_sigmas), ptr(grad_rgbs), ptr(grad_masks), stream_ptr(sigmas.device))
            return grad_sigmas, grad_rgbs, grad_masks, None, None, None, None
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        grad_masks = torch.zeros_like(masks)
        call("inerf_composite_rays_with_masks_train_backward", ptr(grad_weights_sum), ptr(grad_image), ptr(grad_mask_out),
             ptr(sigmas), ptr(rgbs), ptr(masks), ptr(deltas), ptr(rays), ptr(weights_sum), ptr(image), ptr(mask_out),
             M, N, K, float(T_thresh), ptr(grad_sigmas), ptr(grad_rgbs), None, ptr(grad_masks), stream_ptr(sigmas.device))
        return grad_sigmas, grad_rgbs, grad_masks, None, None, None, None


def composite_rays_with_masks_train(sigmas, rgbs, masks, deltas, rays, T_thresh=1e-4, dense=False):
    return _composite_rays_with_masks_train.apply(sigmas, rgbs, masks, deltas, rays, T_thresh, dense)


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1,
               perturb=False, dt_gamma=0, max_steps=1024, noises=None):
    """raymarching.py:370-421 -> (xyzs, dirs, deltas) with n_alive*n_step (+pad) rows, zero where unused"""
    rays_o = _f32(rays_o).view(-1, 3)
    rays_d = _f32(rays_d).view(-1, 3)
    dev = rays_o.device
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
    if noises is None:
        noises = torch.rand(n_alive, dtype=torch.float32, device=dev) if perturb else torch.zeros(n_alive, dtype=torch.float32, device=dev)
    call("inerf_march_rays", int(n_alive), int(n_step), ptr(rays_alive), ptr(rays_t), ptr(rays_o), ptr(rays_d), float(bound),
         float(dt_gamma), int(max_steps), int(C), int(H), ptr(density_bitfield), ptr(near), ptr(far), ptr(xyzs), ptr(dirs),
         ptr(deltas), ptr(noises), stream_ptr(dev))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """raymarching.py:424-446: in-place accumulation into weights_sum / depth / image, kills rays"""
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    call("inerf_composite_rays", int(n_alive), int(n_step), float(T_thresh), ptr(rays_alive), ptr(rays_t), ptr(sigmas), ptr(rgbs),
         ptr(deltas), ptr(weights_sum), ptr(depth), ptr(image), stream_ptr(sigmas.device))
    return tuple()


def composite_rays_with_masks(n_alive, n_step, n_instance, rays_alive, rays_t, sigmas, rgbs, masks, deltas, weights_sum, depth,
                              image, mask_out, T_thresh=1e-2):
    """raymarching.py:449-474"""
    sigmas, rgbs, masks, deltas = _f32(sigmas), _f32(rgbs), _f32(masks), _f32(deltas)
    call("inerf_composite_rays_with_masks", int(n_alive), int(n_step), int(n_instance), float(T_thresh), ptr(rays_alive),
         ptr(rays_t), ptr(sigmas), ptr(rgbs), ptr(masks), ptr(deltas), ptr(weights_sum), ptr(depth), ptr(image), ptr(mask_out),
         stream_ptr(sigmas.device))
    return tuple()


def compact_alive(rays_alive, n_alive):
    """Device-side `rays_alive[rays_alive >= 0]` (mask_renderer.py:370) -> (compacted int32 [n_alive], count int32 [1])."""
    out = torch.empty_like(rays_alive)
    n_out = torch.empty(1, dtype=torch.int32, device=rays_alive.device)
    call("inerf_compact_alive", ptr(rays_alive), int(n_alive), ptr(out), ptr(n_out), stream_ptr(rays_alive.device))
    return out, n_out
