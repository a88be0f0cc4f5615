"""Adam for the trainable parameters of the instance stage, one kernel pass per tensor (csrc/loss.cu: k_adam_step).

Same update rule as the optimizer the reference builds (main_nerf_mask.py:182: torch.optim.Adam(betas=(0.9, 0.99),
eps=1e-15), no weight decay, no amsgrad), but the pass also divides by GradScaler's scale and clears the gradient, so a
training step spends one read of (p, g, m, v) + one write of (p, m, v, g) on the 13.3 M-entry hash table where
zero_grad + fused torch Adam spend two more passes.  Speaks GradScaler's protocol (`_step_supports_amp_scaling`:
`grad_scale` / `found_inf` device scalars, skipped step on non-finite gradients) and keeps its step count on the
device, so it can be captured in a CUDA graph.  Because the step leaves every gradient at zero, callers do NOT call
`zero_grad()` between steps (MaskTrainStep does not).
"""
from __future__ import annotations

import torch

from .._lib import call, ptr, stream_ptr


class FusedAdam(torch.optim.Optimizer):
    _step_supports_amp_scaling = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.grads_cleared_by_step = True
        # data-parallel training: gradients arrive as the SUM over ranks; the division by the world size is folded into this
        # pass (parallel.FlatGradBucket sets it) instead of a separate scaling pass over the 53 MB bucket
        self.grad_div = 1.0

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise RuntimeError("FusedAdam does not take a closure")
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        advanced = []          # distinct step counters touched by this call: ONE advance launch each (normally one in total)
        for group in self.param_groups:
            lr, (b1, b2), eps = float(group["lr"]), group["betas"], float(group["eps"])
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_cuda:
                    raise RuntimeError("FusedAdam: fp32 CUDA parameters only (there is no CPU fallback)")
                st = self.state[p]
                if not st:
                    # every parameter of a device steps together: they share one device-side step counter
                    shared = next((s_["step"] for s_ in self.state.values() if "step" in s_ and s_["step"].device == p.device), None)
                    st["step"] = shared if shared is not None else torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                s = stream_ptr(p.device)
                call("inerf_adam_step", ptr(p), ptr(p.grad), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(), lr, float(b1), float(b2), eps,
                     ptr(st["step"]), ptr(grad_scale), ptr(found_inf), float(self.grad_div), s)
                if not any(t is st["step"] for t in advanced):
                    advanced.append(st["step"])
        for t in advanced:
            call("inerf_adam_advance", ptr(t), ptr(found_inf), stream_ptr(t.device))
        return None
