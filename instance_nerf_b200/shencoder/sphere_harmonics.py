"""SHEncoder over libinerf_b200 (mirrors shencoder/sphere_harmonics.py:14-86)."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd import Function

from .._lib import call, ptr, stream_ptr


class _sh_encoder(Function):
    @staticmethod
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        inputs = inputs.float().contiguous()  # always fp32 (sphere_harmonics.py:16)
        B, input_dim = inputs.shape
        output_dim = degree ** 2
        outputs = torch.empty(B, output_dim, dtype=torch.float32, device=inputs.device)
        dy_dx = torch.empty(B, input_dim * output_dim, dtype=torch.float32, device=inputs.device) if calc_grad_inputs else None
        call("inerf_sh_encode_forward", ptr(inputs), ptr(outputs), B, input_dim, degree, ptr(dy_dx), stream_ptr(inputs.device))
        ctx.save_for_backward(inputs, dy_dx)
        ctx.dims = (B, input_dim, degree)
        return outputs

    @staticmethod
    def backward(ctx, grad):
        inputs, dy_dx = ctx.saved_tensors
        if dy_dx is None:
            return None, None, None
        B, input_dim, degree = ctx.dims
        grad = grad.float().contiguous()
        grad_inputs = torch.zeros_like(inputs)
        call("inerf_sh_encode_backward", ptr(grad), ptr(inputs), B, input_dim, degree, ptr(dy_dx), ptr(grad_inputs),
             stream_ptr(inputs.device))
        return grad_inputs, None, None


def sh_encode(inputs, degree, calc_grad_inputs=False):
    return _sh_encoder.apply(inputs, degree, calc_grad_inputs)


class SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = degree ** 2
        assert self.input_dim == 3, "SH encoder only support input dim == 3"
        assert 0 < self.degree <= 8, "SH encoder only supports degree in [1, 8]"

    def __repr__(self):
        return f"SHEncoder(B200): input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        inputs = inputs / size
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        outputs = sh_encode(inputs, self.degree, inputs.requires_grad)
        return outputs.reshape(prefix_shape + [self.output_dim])
