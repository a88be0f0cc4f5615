"""Test / evaluation loop of the instance stage (instance_nerf_b200/nerf/evaluate.py; MaskTrainer.test / eval_step / test_step,
nerf/utils.py:1375-1496): the device-side frame finalisation against the reference's numpy arithmetic, and the PNG writer
end to end on a provider scene."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import provider_scene as ps

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_provider.npz")


@pytest.mark.parametrize("N,K", [(1000, 32), (257, 37), (64, 1), (33, 256), (0, 8)])
def test_frame_to_u8_matches_numpy_semantics(cuda, N, K):
    from instance_nerf_b200.nerf.evaluate import frame_to_u8
    g = torch.Generator().manual_seed(N + K)
    image = torch.rand(N, 3, generator=g)
    depth = torch.rand(N, generator=g)
    if N:
        image[0] = torch.tensor([0.0, 1.0, 0.999999])
        depth[0] = 1.0
    logits = torch.randn(N, K, generator=g) * 3
    if N > 4 and K > 2:
        logits[1, :] = 0.5                      # all equal: lowest index
        logits[2, K - 1] = logits[2, 0] = 9.0   # tie between first and last
        logits[3, :] = -float("inf")
    rgb, d8, lab = frame_to_u8(image.to(cuda), depth.to(cuda), logits.to(cuda))
    # nerf/utils.py:1461-1467: (pred * 255).astype(np.uint8) on fp32 arrays
    assert np.array_equal(rgb.cpu().numpy(), (image.numpy() * 255).astype(np.uint8))
    assert np.array_equal(d8.cpu().numpy(), (depth.numpy() * 255).astype(np.uint8))
    want = torch.softmax(logits, dim=-1).argmax(dim=-1).numpy() if N else np.zeros(0, np.int64)
    got = lab.cpu().numpy().astype(np.int64)
    ok = np.ones(N, bool)
    if N > 4 and K > 2:
        ok[1:4] = False
        assert got[1] == 0 and got[2] == 0 and got[3] == 0
    assert np.array_equal(got[ok], want[ok])
    # optional outputs
    rgb2, none_d, none_l = frame_to_u8(image.to(cuda))
    assert none_d is None and none_l is None and torch.equal(rgb2, rgb)


def test_test_loop_writes_reference_pngs(cuda, tmp_path, monkeypatch):
    import cv2
    from test_field_gpu import build_model
    from instance_nerf_b200.nerf import evaluate
    from instance_nerf_b200.nerf.provider import NeRFMaskDataset
    from instance_nerf_b200.nerf.trainer import MaskTrainStep
    gold = np.load(GOLD)
    root = str(tmp_path / "scene")
    ps.write_scene(ps.unpack_scene(gold), root)
    mod = types.ModuleType("h5py")
    mod.File = ps.FakeH5File
    monkeypatch.setitem(sys.modules, "h5py", mod)
    opt = ps.options(root, None)
    ds = NeRFMaskDataset(opt, cuda, type="test")
    m, _ = build_model(cuda, ds.num_instances)
    kw = dict(dt_gamma=1 / 128, max_steps=256, T_thresh=1e-4)
    out_dir = str(tmp_path / "results")
    wrote0 = evaluate.test(m, ds.dataloader(), out_dir, name="t", rank=0, world=2, render_kw=kw)
    wrote1 = evaluate.test(m, ds.dataloader(), out_dir, name="t", rank=1, world=2, render_kw=kw)
    assert wrote0 == [0, 2] and wrote1 == [1, 3]
    step = MaskTrainStep(m, patch_size=8, **kw)
    m.eval()
    for i in range(len(ds.poses)):
        data = ds.collate([i])
        with torch.autocast("cuda", dtype=torch.float16):
            rgb, depth, labels = step.test_step(data)
        want_rgb = (rgb[0].float().cpu().numpy() * 255).astype(np.uint8)
        want_depth = (depth[0].float().cpu().numpy() * 255).astype(np.uint8)
        got_rgb = cv2.cvtColor(cv2.imread(os.path.join(out_dir, f"t_{i:04d}_rgb.png")), cv2.COLOR_BGR2RGB)
        got_depth = cv2.imread(os.path.join(out_dir, f"t_{i:04d}_depth.png"), cv2.IMREAD_UNCHANGED)
        got_mask = cv2.imread(os.path.join(out_dir, f"t_{i:04d}_mask.png"), cv2.IMREAD_UNCHANGED)
        assert got_rgb.shape == (ds.H, ds.W, 3) and np.array_equal(got_rgb, want_rgb)
        assert np.array_equal(got_depth, want_depth)
        assert np.array_equal(got_mask, labels[0].cpu().numpy().astype(np.uint8))
        assert os.path.exists(os.path.join(out_dir, f"t_{i:04d}_mask_rgb.png"))
    with pytest.raises((ImportError, NotImplementedError)):
        evaluate.test(m, ds.dataloader(), out_dir, write_video=True)


def test_eval_step_loss_and_maps(cuda, tmp_path, monkeypatch):
    from test_field_gpu import build_model
    from instance_nerf_b200.nerf.provider import NeRFMaskDataset
    from instance_nerf_b200.nerf.trainer import MaskTrainStep
    gold = np.load(GOLD)
    root = str(tmp_path / "scene")
    ps.write_scene(ps.unpack_scene(gold), root)
    mod = types.ModuleType("h5py")
    mod.File = ps.FakeH5File
    monkeypatch.setitem(sys.modules, "h5py", mod)
    ds = NeRFMaskDataset(ps.options(root, os.path.join(root, "mask3d.npy")), cuda, type="val")
    m, _ = build_model(cuda, ds.num_instances)
    m.eval()
    step = MaskTrainStep(m, patch_size=8, label_regularization_weight=0.1, mask3d_loss_weight=0.5, dt_gamma=1 / 128, max_steps=256)
    data = ds.collate([0])
    with torch.autocast("cuda", dtype=torch.float16):
        rgb, depth, pred, gt, loss = step.eval_step(data)
        out = m.render(data["rays_o"], data["rays_d"], render_mask=True, staged=True, bg_color=1, perturb=False, **step.render_kw)
    assert rgb.shape == (1, ds.H, ds.W, 3) and depth.shape == (1, ds.H, ds.W) and pred.shape == gt.shape == (1, ds.H, ds.W)
    logits = out["instance_mask_logits"].reshape(-1, ds.num_instances).float()
    # nerf/utils.py:1393-1402 in the reference's own form
    want = torch.nn.functional.cross_entropy(logits, gt.view(-1))
    want = want + step.label_regularization(out["depth"], out["instance_mask_logits"]) * 0.1
    with torch.autocast("cuda", dtype=torch.float16):
        want = want + step.mask3d_loss(data).mean() * 0.5
    assert abs(float(loss) - float(want)) < 1e-4 * max(1.0, abs(float(want)))
    assert torch.equal(pred.view(-1), torch.softmax(logits, -1).argmax(-1))
