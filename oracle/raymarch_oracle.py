"""TEST INFRASTRUCTURE ONLY -- numpy front-end of oracle/raymarch_oracle.c (built by oracle/Makefile into
oracle/_build/liboracle.so).  Same argument meaning as the reference's pybind functions
(raymarching/src/raymarching.h:7-22) on host arrays; outputs are returned instead of pre-allocated."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build():
    r = subprocess.run(["make", "-C", _HERE, "CC=gcc"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the oracle failed:\n" + r.stdout + r.stderr)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    o, d, aabb = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3), _f32(aabb)
    N = o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().oracle_near_far_from_aabb(_p(o), _p(d), _p(aabb), ctypes.c_uint32(N), ctypes.c_float(min_near), _p(nears), _p(fars))
    return nears, fars


def morton3D(coords):
    c = np.ascontiguousarray(coords, dtype=np.int32)
    out = np.empty(c.shape[0], np.int32)
    lib().oracle_morton3D(_p(c), ctypes.c_uint32(c.shape[0]), _p(out))
    return out


def morton3D_invert(indices):
    i = np.ascontiguousarray(indices, dtype=np.int32)
    out = np.empty((i.shape[0], 3), np.int32)
    lib().oracle_morton3D_invert(_p(i), ctypes.c_uint32(i.shape[0]), _p(out))
    return out


def packbits(grid, thresh):
    g = _f32(grid)
    N = g.size // 8
    out = np.empty(N, np.uint8)
    lib().oracle_packbits(_p(g), ctypes.c_uint32(N), ctypes.c_float(thresh), _p(out))
    return out


def march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, C, H, nears, fars, noises, M=None):
    """-> (xyzs [total,3], dirs, deltas [total,2], rays int32 [N,3] canonical, counter int32 [2])"""
    o, d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    grid = np.ascontiguousarray(grid, dtype=np.uint8)
    nears, fars, noises = _f32(nears), _f32(fars), _f32(noises)
    N = o.shape[0]
    cap = N * max_steps if M is None else M
    # two calls would double the work; allocate the worst case lazily: count first with M = 0
    rays = np.empty((N, 3), np.int32)
    counter = np.zeros(2, np.int32)
    dummy = np.zeros(1, np.float32)
    args = lambda M_, x, dd, dl, r, c: (_p(o), _p(d), _p(grid), ctypes.c_float(bound), ctypes.c_float(dt_gamma), ctypes.c_uint32(max_steps),
                                       ctypes.c_uint32(N), ctypes.c_uint32(C), ctypes.c_uint32(H), ctypes.c_uint32(M_), _p(nears), _p(fars),
                                       _p(x), _p(dd), _p(dl), _p(r), _p(c), _p(noises))
    lib().oracle_march_rays_train(*args(0, dummy, dummy, dummy, rays, counter))
    total = int(counter[0])
    rows = min(total, cap) if M is not None else total
    xyzs, dirs, deltas = np.zeros((max(rows, 1), 3), np.float32), np.zeros((max(rows, 1), 3), np.float32), np.zeros((max(rows, 1), 2), np.float32)
    counter[:] = 0
    lib().oracle_march_rays_train(*args(rows, xyzs, dirs, deltas, rays, counter))
    return xyzs[:rows], dirs[:rows], deltas[:rows], rays, counter


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars, noises):
    o, d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    grid = np.ascontiguousarray(grid, dtype=np.uint8)
    alive = np.ascontiguousarray(rays_alive, dtype=np.int32)
    rays_t, nears, fars, noises = _f32(rays_t), _f32(nears), _f32(fars), _f32(noises)
    M = n_alive * n_step
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    lib().oracle_march_rays(ctypes.c_uint32(n_alive), ctypes.c_uint32(n_step), _p(alive), _p(rays_t), _p(o), _p(d), ctypes.c_float(bound),
                            ctypes.c_float(dt_gamma), ctypes.c_uint32(max_steps), ctypes.c_uint32(C), ctypes.c_uint32(H), _p(grid), _p(nears),
                            _p(fars), _p(xyzs), _p(dirs), _p(deltas), _p(noises))
    return xyzs, dirs, deltas


def composite_rays_train_forward(sigmas, rgbs, masks, deltas, rays, T_thresh=1e-4):
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    rays = np.ascontiguousarray(rays, dtype=np.int32)
    K = 0 if masks is None else masks.shape[1]
    masks = None if masks is None else _f32(masks)
    M, N = sigmas.shape[0], rays.shape[0]
    ws, depth, image = np.empty(N, np.float32), np.empty(N, np.float32), np.empty((N, 3), np.float32)
    mask_out = np.empty((N, K), np.float32) if K else None
    lib().oracle_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(masks), _p(deltas), _p(rays), ctypes.c_uint32(M), ctypes.c_uint32(N),
                                              ctypes.c_uint32(K), ctypes.c_float(T_thresh), _p(ws), _p(depth), _p(image), _p(mask_out))
    return ws, depth, image, mask_out


def composite_rays_train_backward(grad_ws, grad_image, grad_mask_out, sigmas, rgbs, masks, deltas, rays, ws, image, mask_out, T_thresh=1e-4):
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    rays = np.ascontiguousarray(rays, dtype=np.int32)
    K = 0 if masks is None else masks.shape[1]
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gr = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    gm = np.zeros((M, K), np.float32) if K else None
    f = lambda a: None if a is None else _f32(a)
    grad_ws, grad_image, grad_mask_out, masks, ws, image, mask_out = map(f, (grad_ws, grad_image, grad_mask_out, masks, ws, image, mask_out))
    lib().oracle_composite_rays_train_backward(_p(grad_ws), _p(grad_image), _p(grad_mask_out), _p(sigmas), _p(rgbs), _p(masks), _p(deltas),
                                               _p(rays), _p(ws), _p(image), _p(mask_out), ctypes.c_uint32(M), ctypes.c_uint32(N),
                                               ctypes.c_uint32(K), ctypes.c_float(T_thresh), _p(gs), _p(gr), _p(gm))
    return gs, gr, gm


def composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, masks, deltas, weights_sum, depth, image, mask_out):
    """In place on the given numpy arrays (as the reference kernel)."""
    K = 0 if masks is None else masks.shape[1]
    for a in (rays_t, weights_sum, depth, image):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    assert rays_alive.dtype == np.int32
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    masks = None if masks is None else _f32(masks)
    lib().oracle_composite_rays(ctypes.c_uint32(n_alive), ctypes.c_uint32(n_step), ctypes.c_uint32(K), ctypes.c_float(T_thresh), _p(rays_alive),
                                _p(rays_t), _p(sigmas), _p(rgbs), _p(masks), _p(deltas), _p(weights_sum), _p(depth), _p(image), _p(mask_out))


def project_labels(rays_o, rays_d, labels, bbox):
    """oracle_project_labels: labels int32 [nx, ny, nz], bbox [6] (min xyz, max xyz) -> (label int32 [N], t float32 [N])."""
    o, d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    bb = _f32(bbox).reshape(6)
    N = o.shape[0]
    out_l, out_t = np.zeros(N, np.int32), np.zeros(N, np.float32)
    lib().oracle_project_labels(_p(o), _p(d), ctypes.c_uint32(N), _p(lab), ctypes.c_uint32(lab.shape[0]), ctypes.c_uint32(lab.shape[1]),
                                ctypes.c_uint32(lab.shape[2]), _p(bb), _p(out_l), _p(out_t))
    return out_l, out_t
