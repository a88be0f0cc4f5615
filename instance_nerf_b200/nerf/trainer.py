"""Instance-stage training step (mirrors MaskTrainer of the reference: nerf/utils.py:1231-1373 train_step /
label_regularization / mask3d_loss, and the step loop nerf/utils.py:919-939), reduced to what touches the hot path: freeze the
RGB-sigma nets, render a ray batch with `render_mask=True`, cross-entropy on labelled pixels + depth-aware label smoothness on
the 8x8 patches + optional 3D-mask cross-entropy, AMP backward, (optional data-parallel gradient all-reduce,) Adam.

Data providers: nerf/provider.py; checkpoints: nerf/checkpoint.py; the test loop and image writers: nerf/evaluate.py.  Logging,
metrics and the epoch loop of the reference's Trainer are out of scope.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn

import torch.distributed as dist

from .._lib import call, ptr, stream_ptr
from ..parallel import FlatGradBucket


class _MaskLoss(torch.autograd.Function):
    """Cross-entropy on labelled pixels + depth-aware label smoothness, value and gradient in 3 launches
    (inerf_mask_loss / inerf_mask_loss_backward) instead of ~150 elementwise / reduction kernels."""

    @staticmethod
    def forward(ctx, logits, depth, labels, patch, reg_weight):
        logits = logits.float().contiguous()
        N, K = logits.shape
        depth = depth.detach().float().contiguous().view(-1)
        labels = labels.contiguous().view(-1).long()
        acc = torch.empty(6, dtype=torch.float32, device=logits.device)
        loss = torch.empty(1, dtype=torch.float32, device=logits.device)
        call("inerf_mask_loss", ptr(logits), ptr(depth), ptr(labels), N, K, int(patch), float(reg_weight), ptr(acc), ptr(loss),
             stream_ptr(logits.device))
        ctx.save_for_backward(logits, depth, labels, acc)
        ctx.cfg = (N, K, int(patch), float(reg_weight))
        return loss[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        logits, depth, labels, acc = ctx.saved_tensors
        N, K, patch, reg_weight = ctx.cfg
        grad = torch.empty_like(logits)
        call("inerf_mask_loss_backward", ptr(logits), ptr(depth), ptr(labels), N, K, patch, reg_weight, ptr(acc),
             ptr(g.float().contiguous().view(1)), ptr(grad), stream_ptr(logits.device))
        return grad, None, None, None, None


class _TrainStepBase:
    """What the instance-stage and the stage-1 steps share: AMP (GradScaler), FusedAdam, the optional data-parallel flat
    gradient bucket, and the CUDA-graph replay of the whole step over a fixed-size sample stream."""

    def _setup(self, model, params, fp16, data_parallel, fused_adam, cuda_graph):
        dev = next(model.parameters()).device
        # cuda_graph: the whole step (march -> field -> composite -> loss -> backward -> Adam) replays as ONE CUDA graph; at
        # 4096 rays the eager step is bound by ~60 launches of host work, not by the GPU
        # (fp16 is required: the budget-overflow guard below turns the gradients non-finite and relies on an ENABLED
        # GradScaler to skip that optimizer step; without a scaler the replay would write inf into the parameters)
        self.cuda_graph = bool(cuda_graph and fp16 and dev.type == "cuda" and fused_adam)
        self.world = dist.get_world_size() if (data_parallel and dist.is_available() and dist.is_initialized()) else 1
        params = [{"params": g["params"], "lr": float(g["lr"])} for g in params]
        if fused_adam and dev.type == "cuda":
            # main_nerf_mask.py:182's Adam as ONE pass per tensor that also unscales and clears the gradient (nerf/optim.py)
            from .optim import FusedAdam
            self.optimizer = FusedAdam(params, betas=(0.9, 0.99), eps=1e-15)
        else:
            kw = dict(fused=True) if (fused_adam and dev.type == "cuda") else {}
            self.optimizer = torch.optim.Adam(params, betas=(0.9, 0.99), eps=1e-15, **kw)   # main_nerf_mask.py:182
        self._own_zero = getattr(self.optimizer, "grads_cleared_by_step", False)
        self.scaler = torch.amp.GradScaler("cuda", enabled=fp16 and dev.type == "cuda")
        # Data parallel: every gradient is a view of ONE flat buffer; the step all-reduces that buffer in place (SUM) and the
        # optimizer divides by the world size in its own pass.  One extra slot carries the "a rank overflowed its sample
        # budget" flag of the graph-captured step through the same collective.
        self.bucket = FlatGradBucket([p for g in params for p in g["params"]], extra=1).attach() if data_parallel else None
        if self.bucket is not None and hasattr(self.optimizer, "grad_div"):
            self.optimizer.grad_div = float(self.world)
        self._fold_average = self.bucket is not None and hasattr(self.optimizer, "grad_div")
        self.global_step = 0
        self._graph = None          # (torch.cuda.CUDAGraph, static inputs, static loss, budget, counter view)
        self._graph_key = None
        self._samples_seen = 0      # largest marched total of any batch so far
        self._eager_steps = 0
        self.graph_replays = 0
        self.last_total = 0         # marched samples of the last replayed step
        self.graph_captures = 0

    def step(self, data):
        """One optimisation step (nerf/utils.py:929-936); returns the loss tensor (no host sync in the eager path)."""
        if self.cuda_graph:
            return self._step_graphed(data)
        return self._step_eager(data)

    def _forward_backward(self, data, guard=None, flag=None):
        """render -> loss -> backward: gradients land in `.grad` (views of the flat buffer when data parallel)."""
        self.model.train()
        if not self._own_zero:   # FusedAdam leaves every gradient at zero
            self.optimizer.zero_grad(set_to_none=False)
        with torch.autocast("cuda", dtype=torch.float16, enabled=self.fp16):
            _, _, loss = self.train_step(data)
        if guard is not None:
            loss = guard(loss)
        self.scaler.scale(loss).backward()
        if self.bucket is not None:     # the "a rank overflowed its budget" slot rides in the gradient all-reduce; eager steps never overflow
            if flag is not None:
                self.bucket.extra.copy_(flag())
            else:
                self.bucket.extra.zero_()
        return loss.detach()

    def _exchange(self):
        """Data parallel: ONE in-place all_reduce(SUM) of the flat gradient buffer on the current stream."""
        if self.bucket is not None:
            if self._fold_average:
                self.bucket.all_reduce()      # 1/world is folded into FusedAdam's pass
            else:
                self.bucket.sync()

    def _optimizer_step(self):
        self.scaler.step(self.optimizer)
        self.scaler.update()

    def _step_eager(self, data, guard=None, flag=None):
        self.global_step += 1
        loss = self._forward_backward(data, guard, flag)
        self._exchange()
        self._optimizer_step()
        return loss

    # ---- CUDA-graph replay of the whole step -------------------------------------------------------------------------
    # The sample stream of a step has a data-dependent length (the marched total M).  The graph is captured with a FIXED
    # stream of `budget` rows (model.sample_budget -> march_rays_train's no-sync budget mode); the marcher leaves the real
    # total in step_counter[slot][0] on the device, the field forward / backward kernels stop there (inerf_field_desc.n_valid)
    # and the compositing kernels only touch rows that rays reference, so padding costs nothing and no ray is dropped while
    # M <= budget.  If a batch ever marches more than `budget` samples, a device-side guard turns that step's loss into +inf:
    # GradScaler sees non-finite gradients and skips the optimizer step, so a truncated batch never updates the parameters;
    # the host reads the total after every replay (the one sync of the step, where the reference reads loss.item(),
    # nerf/utils.py:937), re-runs that batch eagerly and re-captures with a larger budget.
    #
    # Data parallel: the step is TWO graphs -- (march ... backward) and (inf check + Adam + scaler update) -- with the NCCL
    # all-reduce of the flat gradient buffer issued eagerly between them.  A collective captured INSIDE the graph replayed at
    # 35-67 ms per step on 2 B200s (NCCL 2.28.9 / torch 2.11) against 10.9 ms for the eager step, so it stays outside; the cost
    # is two extra launches.  Whether any rank overflowed its budget rides through the collective in one extra float, so all
    # ranks take the same redo / re-capture decision and the collective sequence never diverges.
    GRAPH_WARMUP_STEPS = 3
    GRAPH_HEADROOM = 1.125

    def _counter_slot(self):
        # run_cuda uses step_counter[local_step % 16] and then increments local_step; the captured graph keeps ONE slot
        return self.model.step_counter[self.model.local_step % 16]

    def _capture(self, data):
        from .._lib import PARAM_EPOCH
        model = self.model
        budget = int(self._samples_seen * self.GRAPH_HEADROOM) + 4096
        budget += (-budget) % 4096
        static = {k: v.clone() for k, v in data.items() if torch.is_tensor(v)}
        counter = self._counter_slot()
        limit = torch.tensor(budget, dtype=torch.int32, device=counter.device)
        inf = torch.tensor(float("inf"), dtype=torch.float32, device=counter.device)
        one = torch.ones((), dtype=torch.float32, device=counter.device)

        def guard(loss):   # x inf, not where(.., inf, loss): the GRADIENTS must turn non-finite for GradScaler to skip the step
            return loss.float() * torch.where(counter[0] > limit, inf, one)

        def flag():        # data parallel: 1.0 on a rank that overflowed; summed over ranks by the gradient all-reduce itself
            return (counter[0] > limit).float().view(1)

        model.sample_budget = budget
        PARAM_EPOCH[0] += 1                       # the packed fp16 tables / weight blobs must be rebuilt INSIDE the graph
        slot = model.local_step
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        g2 = None
        try:
            if self.bucket is None:
                with torch.cuda.graph(g):
                    loss = self._step_eager(static, guard)
            else:
                # thread_local: NCCL's watchdog thread may query the events of earlier collectives while this thread captures
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self.global_step += 1
                    loss = self._forward_backward(static, guard, flag)
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, pool=g.pool(), capture_error_mode="thread_local"):
                    self._optimizer_step()
        finally:
            model.sample_budget = 0
            model._n_valid_ptr = None
        model.local_step = slot                   # capture advanced it; replays always use the captured slot
        self.global_step -= 1
        # `limit`, `inf`, `one` are read by the captured kernels on every replay: they must outlive this function (a tensor freed
        # here goes back to the caching allocator and the graph would read whatever lands in its block next)
        self._graph = (g, static, loss, budget, counter, g2, (limit, inf, one))
        self._graph_key = tuple((k, tuple(v.shape), v.dtype) for k, v in static.items())
        self.graph_captures += 1

    def _step_graphed(self, data):
        from .._lib import PARAM_EPOCH
        tensors = {k: v for k, v in data.items() if torch.is_tensor(v)}
        key = tuple((k, tuple(v.shape), v.dtype) for k, v in tensors.items())
        if self._graph is not None and key != self._graph_key:
            self._graph = None
        if self._graph is None:
            if self._eager_steps < self.GRAPH_WARMUP_STEPS or key != self._graph_key:
                slot = self._counter_slot()
                loss = self._step_eager(data)
                self._samples_seen = max(self._samples_seen, int(slot[0].item()))
                self._eager_steps += 1
                self._graph_key = key
                return loss
            self._capture(data)
        g, static, loss, budget, counter, g2, _keep = self._graph
        for k, v in tensors.items():
            static[k].copy_(v, non_blocking=True)
        g.replay()
        if g2 is not None:                        # data parallel: eager all-reduce between the two halves of the step
            self._exchange()
            g2.replay()
        self.global_step += 1
        self.graph_replays += 1
        PARAM_EPOCH[0] += 1                       # parameters changed without their version counters moving
        # the replay always writes the captured counter slot; keep the model's per-step bookkeeping (step_counter ring +
        # local_step, which update_extra_state averages into mean_count, mask_renderer.py:543-546) moving as the eager path does
        model = self.model
        model.step_counter[model.local_step % 16].copy_(counter)
        model.local_step += 1
        if self.bucket is not None:               # one host read: own total + how many ranks overflowed (rode in the all-reduce)
            total, n_over = (int(v) for v in torch.stack([counter[0].double(), self.bucket.extra[0].double()]).tolist())
        else:
            total = int(counter[0].item())
            n_over = int(total > budget)
        self.last_total = total
        self._samples_seen = max(self._samples_seen, total)
        if n_over > 0:                            # the guard skipped this step on EVERY rank (inf gradients are summed): redo it exactly, then re-capture
            self._graph = None
            self.global_step -= 1
            model.local_step -= 1                 # the eager redo records this step again
            return self._step_eager(data)
        return loss


class MaskTrainStep(_TrainStepBase):
    def __init__(self, model, lr=1e-2, fp16=True, patch_size=8, label_regularization_weight=0.0, dt_gamma=1 / 128, max_steps=1024,
                 T_thresh=1e-4, data_parallel=False, fused_adam=True, fused_loss=True, cuda_graph=False, mask3d_loss_weight=0.0):
        self.model = model
        self.opt = SimpleNamespace(patch_size=patch_size, label_regularization_weight=label_regularization_weight,
                                   mask3d_loss_weight=mask3d_loss_weight)
        self.render_kw = dict(dt_gamma=dt_gamma, max_steps=max_steps, T_thresh=T_thresh)
        self.fp16 = fp16
        self.fused_loss = fused_loss
        self.num_instances = model.num_instances
        # freeze rgb and density (nerf/utils.py:1242-1246)
        model.encoder.requires_grad_(False)
        model.sigma_net.requires_grad_(False)
        model.encoder_dir.requires_grad_(False)
        model.color_net.requires_grad_(False)
        self.criterion = nn.CrossEntropyLoss(reduction="none")          # main_nerf_mask.py:177
        params = [{"params": [p for p in g["params"] if p.requires_grad], "lr": g["lr"]} for g in model.get_params(lr)]
        params = [g for g in params if g["params"]]
        self._setup(model, params, fp16, data_parallel, fused_adam, cuda_graph and fused_loss)

    def label_regularization(self, depth, pred_masks):
        """nerf/utils.py:1262-1285: squared differences of neighbouring logits inside each patch, weighted by exp(-ddepth^2)."""
        p = self.opt.patch_size
        pm = pred_masks.view(-1, p, p, self.num_instances).permute(0, 3, 1, 2).contiguous()
        diff_x = pm[:, :, :, 1:] - pm[:, :, :, :-1]
        diff_y = pm[:, :, 1:, :] - pm[:, :, :-1, :]
        depth = depth.view(-1, p, p)
        ddx = depth[:, :, 1:] - depth[:, :, :-1]
        ddy = depth[:, 1:, :] - depth[:, :-1, :]
        wx = torch.exp(-(ddx * ddx)).unsqueeze(1).expand_as(diff_x)
        wy = torch.exp(-(ddy * ddy)).unsqueeze(1).expand_as(diff_y)
        return torch.sum(diff_x * diff_x * wx) / torch.sum(wx) + torch.sum(diff_y * diff_y * wy) / torch.sum(wy)

    def train_step(self, data):
        """nerf/utils.py:1287-1373 -> (pred_masks [B,N] argmax, gt_masks, loss)"""
        rays_o, rays_d, gt_masks = data["rays_o"], data["rays_d"], data["masks"]
        outputs = self.model.render(rays_o, rays_d, render_mask=True, staged=False, bg_color=1, perturb=True,
                                    force_all_rays=self.opt.patch_size != 1, noises=data.get("noises"), **self.render_kw)
        pred = outputs["instance_mask_logits"]
        flat = pred.view(-1, self.num_instances)
        gt = gt_masks.view(-1)
        p = self.opt.patch_size
        if self.fused_loss and flat.is_cuda and flat.shape[0] % (p * p) == 0:
            loss = _MaskLoss.apply(flat, outputs["depth"], gt, p, self.opt.label_regularization_weight)
        else:
            labeled = gt != -1
            # same value as the reference's boolean-index form, without the host sync of `labeled.sum() > 0`
            ce = self.criterion(flat.float(), torch.where(labeled, gt, torch.zeros_like(gt)))
            loss = (ce * labeled).sum() / labeled.sum().clamp(min=1)
            if self.opt.label_regularization_weight > 0:
                loss = loss + self.label_regularization(outputs["depth"], pred) * self.opt.label_regularization_weight
        if self.opt.mask3d_loss_weight > 0:                     # 3d mask constraints (nerf/utils.py:1367-1369)
            loss = loss + self.mask3d_loss(data).mean() * self.opt.mask3d_loss_weight
        return pred.argmax(dim=-1), gt_masks, loss

    @torch.no_grad()
    def eval_step(self, data, bg_color=1):
        """nerf/utils.py:1375-1408 -> (pred_rgb [B,H,W,3], pred_depth [B,H,W], pred_masks [B,H,W] argmax, gt_masks, loss)"""
        rays_o, rays_d, gt_masks = data["rays_o"], data["rays_d"], data["masks"]
        B, H, W = gt_masks.shape
        outputs = self.model.render(rays_o, rays_d, render_mask=True, staged=True, bg_color=bg_color, perturb=False, **self.render_kw)
        logits = outputs["instance_mask_logits"].reshape(B, H, W, -1)
        flat, gt = logits.view(-1, self.num_instances).float(), gt_masks.view(-1)
        labeled = gt != -1
        loss = (self.criterion(flat, torch.where(labeled, gt, torch.zeros_like(gt))) * labeled).sum() / labeled.sum().clamp(min=1)
        if self.opt.label_regularization_weight > 0:
            loss = loss + self.label_regularization(outputs["depth"], logits) * self.opt.label_regularization_weight
        if self.opt.mask3d_loss_weight > 0:
            loss = loss + self.mask3d_loss(data).mean() * self.opt.mask3d_loss_weight
        return outputs["image"].reshape(B, H, W, 3), outputs["depth"].reshape(B, H, W), logits.argmax(dim=-1), gt_masks, loss

    @torch.no_grad()
    def test_step(self, data, bg_color=None, perturb=False):
        """nerf/utils.py:1410-1431 -> (pred_rgb [B,H,W,3], pred_depth [B,H,W], pred_masks [B,H,W])"""
        B, H, W = data["rays_o"].shape[0], data["H"], data["W"]
        outputs = self.model.render(data["rays_o"], data["rays_d"], render_mask=True, staged=True, bg_color=bg_color, perturb=perturb,
                                    **self.render_kw)
        return (outputs["image"].reshape(-1, H, W, 3), outputs["depth"].reshape(-1, H, W),
                outputs["instance_mask_logits"].reshape(B, H, W, -1).argmax(dim=-1))

    def mask3d_loss(self, data):
        """nerf/utils.py:1250-1260: cross-entropy of the instance head queried at labelled 3D points (`mask3d_coords` [N,3],
        `mask3d_labels` [N]) -> [N].  On the fused path the query is the same one-launch field forward / backward the ray
        samples take (the direction input only feeds the colour net, whose output is not used)."""
        coords, labels = data["mask3d_coords"].view(-1, 3), data["mask3d_labels"].view(-1).long()
        model = self.model
        dirs = torch.zeros_like(coords)
        dirs[:, 2] = 1.0
        if hasattr(model, "fused_train_available") and torch.is_grad_enabled() and model.fused_train_available(coords, dirs):
            saved, model._n_valid_ptr = model._n_valid_ptr, None     # every row of this batch is live (no marcher total applies)
            try:
                _, _, logits = model.forward_fused_train(coords, dirs)
            finally:
                model._n_valid_ptr = saved
        else:
            logits = model.mask(coords, geo_feat=model.density(coords)["geo_feat"])
        return self.criterion(logits.float(), labels)


class RGBTrainStep(_TrainStepBase):
    """Stage-1 RGB-sigma training step (mirrors Trainer.train_step, nerf/utils.py:536-632, and the step loop :919-939):
    render a ray batch, MSE against the ground-truth colours (alpha-blended onto a random per-pixel background when the
    images carry alpha), AMP backward through compositing, both MLPs, SH and the sigma hash table, Adam.  With fp16 autocast
    and the standard architecture the field is one launch forward (inerf_field_forward_train_rgb) and one launch backward
    (inerf_field_backward_rgb), and `cuda_graph=True` replays the whole step as the instance stage does.  The LPIPS patch
    term, CLIP loss and the error-map resampling of the reference's Trainer are outside the hot path."""

    def __init__(self, model, lr=1e-2, fp16=True, patch_size=1, dt_gamma=1 / 128, max_steps=1024, T_thresh=1e-4, data_parallel=False,
                 fused_adam=True, cuda_graph=False):
        self.model = model
        self.opt = SimpleNamespace(patch_size=patch_size)
        self.render_kw = dict(dt_gamma=dt_gamma, max_steps=max_steps, T_thresh=T_thresh)
        self.fp16 = fp16
        self.criterion = nn.MSELoss(reduction="none")                                   # main_nerf.py:104
        params = [{"params": list(g["params"]), "lr": g["lr"]} for g in model.get_params(lr)]   # main_nerf.py:120: Adam(betas=(0.9, 0.99), eps=1e-15)
        params = [g for g in params if g["params"]]
        self._setup(model, params, fp16, data_parallel, fused_adam, cuda_graph)

    def train_step(self, data):
        """-> (pred_rgb [B,N,3], gt_rgb [B,N,3], loss)"""
        rays_o, rays_d, images = data["rays_o"], data["rays_d"], data["images"]
        C = images.shape[-1]
        if C == 3 or self.model.bg_radius > 0:
            bg_color = 1
        else:
            bg_color = data["bg_color"] if "bg_color" in data else torch.rand_like(images[..., :3])
        gt_rgb = images[..., :3] * images[..., 3:] + bg_color * (1 - images[..., 3:]) if C == 4 else images
        bg = bg_color if isinstance(bg_color, int) else bg_color.reshape(-1, 3)
        outputs = self.model.render(rays_o, rays_d, staged=False, bg_color=bg, perturb=True, force_all_rays=self.opt.patch_size != 1,
                                    noises=data.get("noises"), **self.render_kw)
        pred_rgb = outputs["image"]
        loss = self.criterion(pred_rgb, gt_rgb).mean(-1).mean()
        return pred_rgb, gt_rgb, loss
