"""`_backend` shims: the reference's pybind FFI, function for function, over libinerf_b200's C ABI.

The reference's op wrappers do `import _raymarching as _backend` (raymarching/raymarching.py:9-12),
`import _gridencoder as _backend` (gridencoder/grid.py:9-12), `import _shencoder as _backend`
(shencoder/sphere_harmonics.py:9-12) and call `_backend.<fn>(tensors..., sizes..., outputs...)` with the
prototypes of raymarching/src/raymarching.h:5-22, gridencoder/src/gridencoder.h:12-15, shencoder/src/shencoder.h:9-10.
The three objects below accept exactly those argument lists (torch CUDA tensors in the reference's layouts,
caller-allocated outputs, nothing returned), so a reference checkout switches to the B200 kernels with

    import instance_nerf_b200.backend as b
    sys.modules["_raymarching"], sys.modules["_gridencoder"], sys.modules["_shencoder"] = b.raymarching, b.gridencoder, b.shencoder

before `import raymarching`.  Differences from the reference modules: kernels run on torch's current
stream; a failed launch or a bad argument raises RuntimeError; `march_rays_train` writes the canonical (sorted) stream.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from ._lib import call, lib, ptr, stream_ptr

_DT = {torch.float32: 0, torch.float16: 1}


def _st(t):
    return stream_ptr(t.device)


# ---------------------------------------------------------------------------------- _raymarching (13 functions) --
def _near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars):
    call("inerf_near_far_from_aabb", ptr(rays_o), ptr(rays_d), ptr(aabb), int(N), float(min_near), ptr(nears), ptr(fars), _st(rays_o))


def _sph_from_ray(rays_o, rays_d, radius, N, coords):
    call("inerf_sph_from_ray", ptr(rays_o), ptr(rays_d), float(radius), int(N), ptr(coords), _st(rays_o))


def _morton3D(coords, N, indices):
    call("inerf_morton3D", ptr(coords), int(N), ptr(indices), _st(coords))


def _morton3D_invert(indices, N, coords):
    call("inerf_morton3D_invert", ptr(indices), int(N), ptr(coords), _st(indices))


def _packbits(grid, N, density_thresh, bitfield):
    call("inerf_packbits", ptr(grid), int(N), float(density_thresh), ptr(bitfield), _st(grid))


def _march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises):
    # the reference's argument list + the caller-owned workspace the single-walk marcher needs (include/inerf_b200.h)
    scratch = torch.empty(max(1, int(lib().inerf_march_scratch_floats(int(N), int(max_steps)))), dtype=torch.float32, device=rays_o.device)
    call("inerf_march_rays_train", ptr(rays_o), ptr(rays_d), ptr(grid), float(bound), float(dt_gamma), int(max_steps), int(N), int(C), int(H),
         int(M), ptr(nears), ptr(fars), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(rays), ptr(counter), ptr(noises), ptr(scratch), _st(rays_o))


def _composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image):
    call("inerf_composite_rays_train_forward", ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(rays), int(M), int(N), float(T_thresh),
         ptr(weights_sum), ptr(depth), ptr(image), _st(sigmas))


def _composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N, T_thresh, grad_sigmas,
                                   grad_rgbs):
    call("inerf_composite_rays_train_backward", ptr(grad_weights_sum), ptr(grad_image), ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(rays),
         ptr(weights_sum), ptr(image), int(M), int(N), float(T_thresh), ptr(grad_sigmas), ptr(grad_rgbs), _st(sigmas))


def _composite_rays_with_masks_train_forward(sigmas, rgbs, masks, deltas, rays, M, N, K, T_thresh, weights_sum, depth, image, mask_out):
    call("inerf_composite_rays_with_masks_train_forward", ptr(sigmas), ptr(rgbs), ptr(masks), ptr(deltas), ptr(rays), int(M), int(N), int(K),
         float(T_thresh), ptr(weights_sum), ptr(depth), ptr(image), ptr(mask_out), _st(sigmas))


def _composite_rays_with_masks_train_backward(grad_weights_sum, grad_image, grad_mask_out, sigmas, rgbs, masks, deltas, rays, weights_sum, image,
                                              mask_out, M, N, K, T_thresh, grad_sigmas, grad_rgbs, grad_masks_acc, grad_masks):
    call("inerf_composite_rays_with_masks_train_backward", ptr(grad_weights_sum), ptr(grad_image), ptr(grad_mask_out), ptr(sigmas), ptr(rgbs),
         ptr(masks), ptr(deltas), ptr(rays), ptr(weights_sum), ptr(image), ptr(mask_out), int(M), int(N), int(K), float(T_thresh),
         ptr(grad_sigmas), ptr(grad_rgbs), ptr(grad_masks_acc), ptr(grad_masks), _st(sigmas))


def _march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars, xyzs, dirs, deltas,
                noises):
    call("inerf_march_rays", int(n_alive), int(n_step), ptr(rays_alive), ptr(rays_t), ptr(rays_o), ptr(rays_d), float(bound), float(dt_gamma),
         int(max_steps), int(C), int(H), ptr(grid), ptr(nears), ptr(fars), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(noises), _st(rays_o))


def _composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
    call("inerf_composite_rays", int(n_alive), int(n_step), float(T_thresh), ptr(rays_alive), ptr(rays_t), ptr(sigmas), ptr(rgbs), ptr(deltas),
         ptr(weights_sum), ptr(depth), ptr(image), _st(sigmas))


def _composite_rays_with_masks(n_alive, n_step, K, T_thresh, rays_alive, rays_t, sigmas, rgbs, masks, deltas, weights_sum, depth, image,
                               mask_out):
    call("inerf_composite_rays_with_masks", int(n_alive), int(n_step), int(K), float(T_thresh), ptr(rays_alive), ptr(rays_t), ptr(sigmas),
         ptr(rgbs), ptr(masks), ptr(deltas), ptr(weights_sum), ptr(depth), ptr(image), ptr(mask_out), _st(sigmas))


raymarching = SimpleNamespace(
    near_far_from_aabb=_near_far_from_aabb, sph_from_ray=_sph_from_ray, morton3D=_morton3D, morton3D_invert=_morton3D_invert,
    packbits=_packbits, march_rays_train=_march_rays_train, composite_rays_train_forward=_composite_rays_train_forward,
    composite_rays_train_backward=_composite_rays_train_backward,
    composite_rays_with_masks_train_forward=_composite_rays_with_masks_train_forward,
    composite_rays_with_masks_train_backward=_composite_rays_with_masks_train_backward, march_rays=_march_rays,
    composite_rays=_composite_rays, composite_rays_with_masks=_composite_rays_with_masks)


# ------------------------------------------------------------------------------------- _gridencoder (3 functions) --
def _grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners, interp):
    """outputs is [L, B, C] (gridencoder/grid.py:47), dtype of `embeddings` (fp32, or fp16 under autocast)."""
    call("inerf_grid_encode_forward", ptr(inputs), ptr(embeddings), ptr(offsets), ptr(outputs), int(B), int(D), int(C), int(L), float(S), int(H),
         ptr(dy_dx), int(gridtype), int(bool(align_corners)), int(interp), _DT[embeddings.dtype], 0, _st(inputs))


def _grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs, gridtype, align_corners, interp):
    """grad is [L, B, C] (gridencoder/grid.py:75); grad_embeddings pre-zeroed by the caller (grid.py:77)."""
    call("inerf_grid_encode_backward", ptr(grad), ptr(inputs), ptr(embeddings), ptr(offsets), ptr(grad_embeddings), int(B), int(D), int(C), int(L),
         float(S), int(H), ptr(dy_dx), ptr(grad_inputs), int(gridtype), int(bool(align_corners)), int(interp), _DT[grad_embeddings.dtype], 0,
         _st(inputs))


def _grad_total_variation(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners):
    call("inerf_grad_total_variation", ptr(inputs.float().contiguous()), ptr(embeddings), ptr(grad), ptr(offsets), float(weight), int(B), int(D), int(C),
         int(L), float(S), int(H), int(gridtype), int(bool(align_corners)), _DT[embeddings.dtype], _st(inputs))


gridencoder = SimpleNamespace(grid_encode_forward=_grid_encode_forward, grid_encode_backward=_grid_encode_backward,
                              grad_total_variation=_grad_total_variation)


# --------------------------------------------------------------------------------------- _shencoder (2 functions) --
def _sh_encode_forward(inputs, outputs, B, D, C, dy_dx):
    call("inerf_sh_encode_forward", ptr(inputs), ptr(outputs), int(B), int(D), int(C), ptr(dy_dx), _st(inputs))


def _sh_encode_backward(grad, inputs, B, D, C, dy_dx, grad_inputs):
    call("inerf_sh_encode_backward", ptr(grad), ptr(inputs), int(B), int(D), int(C), ptr(dy_dx), ptr(grad_inputs), _st(inputs))


shencoder = SimpleNamespace(sh_encode_forward=_sh_encode_forward, sh_encode_backward=_sh_encode_backward)
