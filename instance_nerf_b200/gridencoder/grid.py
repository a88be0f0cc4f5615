"""GridEncoder over libinerf_b200 (mirrors gridencoder/grid.py:24-185 of the reference).

Same constructor, parameter / buffer names (`embeddings`, `offsets` -> same
state-dict keys), same forward contract.  Differences:

* the kernel writes [B, L*C] directly (no [L,B,C] -> permute -> reshape copy,
  grid.py:57,75);
* under autocast the fp16 copy of the table is cached and refreshed only when
  the parameter's version counter changes, instead of casting the whole table on
  every call (grid.py:43-44: 51 MB read + 25 MB write per call per encoder).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from .._lib import PARAM_EPOCH, call, invalidate_param_caches, ptr, stream_ptr

_gridtype_to_id = {"hash": 0, "tiled": 1}
_interp_to_id = {"linear": 0, "smoothstep": 1}
_DT = {torch.float32: 0, torch.float16: 1}


class _grid_encode(Function):
    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False, interpolation=0, embeddings_lowp=None):
        inputs = inputs.float().contiguous()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)
        table = embeddings
        if torch.is_autocast_enabled() and C % 2 == 0:
            table = embeddings_lowp if embeddings_lowp is not None else embeddings.to(torch.half)
        table = table.contiguous()
        outputs = torch.empty(B, L * C, device=inputs.device, dtype=table.dtype)
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=table.dtype) if calc_grad_inputs else None
        call("inerf_grid_encode_forward", ptr(inputs), ptr(table), ptr(offsets), ptr(outputs), B, D, C, L, S, H, ptr(dy_dx),
             gridtype, int(align_corners), interpolation, _DT[table.dtype], 1, stream_ptr(inputs.device))
        ctx.save_for_backward(inputs, offsets, dy_dx)
        ctx.dims = (B, D, C, L, S, H, gridtype, interpolation, align_corners, table.dtype, tuple(embeddings.shape))
        return outputs

    @staticmethod
    def backward(ctx, grad):
        inputs, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, interpolation, align_corners, dtype, shape = ctx.dims
        grad = grad.to(dtype).contiguous()
        grad_embeddings = torch.zeros(shape, dtype=dtype, device=grad.device)
        grad_inputs = torch.zeros(B, D, dtype=dtype, device=grad.device) if dy_dx is not None else None
        call("inerf_grid_encode_backward", ptr(grad), ptr(inputs), None, ptr(offsets), ptr(grad_embeddings), B, D, C, L, S, H,
             ptr(dy_dx), ptr(grad_inputs), gridtype, int(align_corners), interpolation, _DT[dtype], 1, stream_ptr(grad.device))
        if grad_inputs is not None:
            grad_inputs = grad_inputs.to(inputs.dtype)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None, None


def grid_encode(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False, interpolation=0, embeddings_lowp=None):
    return _grid_encode.apply(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs, gridtype,
                              align_corners, interpolation, embeddings_lowp)


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype="hash", align_corners=False, interpolation="linear"):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.interpolation = interpolation
        self.interp_id = _interp_to_id[interpolation]
        self.align_corners = align_corners

        # level table sizes: grid.py:118-129
        self.max_params = 2 ** log2_hashmap_size
        offsets, offset = [], 0
        for i in range(num_levels):
            resolution = int(np.ceil(base_resolution * per_level_scale ** i))
            n = min(self.max_params, (resolution if align_corners else resolution + 1) ** input_dim)
            n = int(np.ceil(n / 8) * 8)
            offsets.append(offset)
            offset += n
        offsets.append(offset)
        self.register_buffer("offsets", torch.from_numpy(np.array(offsets, dtype=np.int32)))
        self.n_params = offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
        self.reset_parameters()
        self._lowp = None
        self._lowp_key = None

    def reset_parameters(self):
        std = 1e-4
        self.embeddings.data.uniform_(-std, std)
        invalidate_param_caches()   # written through `.data`: the version counter the fp16 caches key on did not move

    def __repr__(self):
        return (f"GridEncoder(B200): input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} gridtype={self.gridtype} "
                f"align_corners={self.align_corners} interpolation={self.interpolation}")

    def half_table(self) -> torch.Tensor:
        """Persistent fp16 shadow of `embeddings`, refreshed when the parameter changes."""
        e = self.embeddings
        key = (e.data_ptr(), e._version, e.device, PARAM_EPOCH[0])
        if self._lowp is None or self._lowp_key != key:
            self._lowp = e.detach().to(torch.half).contiguous()
            self._lowp_key = key
        return self._lowp

    def forward(self, inputs, bound=1):
        inputs = (inputs + bound) / (2 * bound)  # map to [0, 1]  (grid.py:149)
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        lowp = None
        if torch.is_autocast_enabled() and self.level_dim % 2 == 0 and not (torch.is_grad_enabled() and self.embeddings.requires_grad):
            lowp = self.half_table()
        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              inputs.requires_grad, self.gridtype_id, self.align_corners, self.interp_id, lowp)
        return outputs.view(prefix_shape + [self.output_dim])

    @torch.no_grad()
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        """grid.py:163-185 / gridencoder.cu:504-642: adds the TV gradient of the vertices hit by `inputs` (random points when
        None) into `self.embeddings.grad`; call after loss.backward() and before optimizer.step().  Always fp32 (the reference
        disables autocast here)."""
        if inputs is None:
            inputs = torch.rand(B, self.input_dim, device=self.embeddings.device)
        else:
            inputs = ((inputs + bound) / (2 * bound)).view(-1, self.input_dim)
            B = inputs.shape[0]
        if self.embeddings.grad is None:
            raise ValueError("grad is None, should be called after loss.backward() and before optimizer.step()!")
        inputs = inputs.float().contiguous()
        emb, grad = self.embeddings, self.embeddings.grad
        call("inerf_grad_total_variation", ptr(inputs), ptr(emb.detach()), ptr(grad), ptr(self.offsets), float(weight), B, self.input_dim,
             emb.shape[1], self.num_levels, float(np.log2(self.per_level_scale)), self.base_resolution, self.gridtype_id, int(self.align_corners),
             0 if emb.dtype == torch.float32 else 1, stream_ptr(emb.device))
