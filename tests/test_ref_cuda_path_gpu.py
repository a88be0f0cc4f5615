"""Full-frame parity of the product against the REFERENCE's own CUDA path on the headline configuration (SURVEY.md section 8d,
BASELINE.md B2): frame 0 of c2 -- 640x480, K = 32, bound 8, 4 x 128^3 occupancy grid, 16-level 2^19 hash grids -- rendered
(a) by the reference's kernels built unmodified into oracle/_ref, driven by the reference's run_cuda alive-ray loop with the MLPs
as nn.Linear under fp16 autocast (oracle/ref_cuda_path.py), and (b) by NeRFNetwork.render -> inerf_render_fused (one launch).
North-star tolerance: image / depth / instance probabilities within 1e-3, instance argmax identical."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_c2_frame_matches_reference_cuda_path(ref, cuda):
    import bench
    from instance_nerf_b200.nerf.utils import get_rays
    from oracle.ref_cuda_path import RefPath
    model, scene, poses = bench.build_scene_and_model(cuda)
    r = get_rays(poses[0:1].to(cuda), bench.intrinsics(), bench.H_IMG, bench.W_IMG)
    o, d = r["rays_o"].view(-1, 3), r["rays_d"].view(-1, 3)
    kw = dict(dt_gamma=bench.DT_GAMMA, max_steps=bench.MAX_STEPS, T_thresh=bench.T_THRESH)
    want = RefPath(model, ref).render(o, d, **kw)
    with torch.no_grad():
        got = model.render(o[None], d[None], staged=True, render_mask=True, perturb=False, bg_color=1, **kw)
    assert o.shape[0] == 307200 and want["evaluated"] > 30e6
    e_img = float((got["image"][0] - want["image"]).abs().max())
    e_dep = float((got["depth"][0] - want["depth"]).abs().max())
    e_prob = float((torch.softmax(got["instance_mask_logits"][0], -1) - torch.softmax(want["instance_mask_logits"], -1)).abs().max())
    agree = float((got["instance_mask_logits"][0].argmax(-1) == want["instance_mask_logits"].argmax(-1)).float().mean())
    print(f"[c2 frame vs reference CUDA path] max abs diff image {e_img:.2e} depth {e_dep:.2e} instance-prob {e_prob:.2e}, argmax agreement {agree:.6f}")
    from helpers import record_metric
    record_metric("c2_frame_vs_reference_cuda_path", image=e_img, depth=e_dep, instance_prob=e_prob, argmax_agreement=agree,
                  evaluated_samples_reference=int(want["evaluated"]))
    assert e_img <= 1e-3 and e_dep <= 1e-3 and e_prob <= 1e-3
    assert agree == 1.0


def test_reference_train_step_runs_and_matches_product_loss(ref, cuda):
    """The reference-kernel training step (oracle/ref_cuda_path.RefPath.train_step, the baseline bench.py times next to the
    product's step) computes the same first-step loss as MaskTrainStep on identical weights, rays, labels and jitter."""
    import bench
    from instance_nerf_b200.nerf.trainer import MaskTrainStep
    from oracle.ref_cuda_path import RefPath
    model, scene, poses = bench.build_scene_and_model(cuda)
    batch = bench.train_batches(cuda, scene, poses, 1024, 0, 1, n=1)[0]
    batch["noises"] = torch.rand(1024, device=cuda, generator=torch.Generator(device=cuda).manual_seed(2))
    kw = dict(dt_gamma=bench.DT_GAMMA, max_steps=bench.MAX_STEPS, T_thresh=bench.T_THRESH)
    rp = RefPath(model, ref)
    rp.make_optimizer()
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    l_ref = float(rp.train_step(batch, patch=8, reg_weight=0.1, **kw))
    model.load_state_dict(state)
    tr = MaskTrainStep(model, lr=1e-2, fp16=True, patch_size=8, label_regularization_weight=0.1, **kw)
    l_ours = float(tr.step(batch))
    print(f"[train step] first-step loss: reference kernels {l_ref:.6f}, product {l_ours:.6f}")
    from helpers import record_metric
    record_metric("train_step_first_loss", reference_kernels=l_ref, product=l_ours)
    assert abs(l_ours - l_ref) <= 2e-3 * max(1.0, abs(l_ref))
