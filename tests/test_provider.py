"""Scene providers (instance_nerf_b200/nerf/provider.py) against the reference's own NeRFMaskDataset / NeRFDataset
(nerf/provider.py:97-638, imported unmodified and run on the CPU by tests/golden/make_golden_provider.py ->
tests/golden/ref_provider.npz) on the same synthetic scene directory.  Host-side loading is checked on the CPU; batch
assembly (inerf_get_rays + label gather on the device) on the GPU with the reference's drawn pixel indices injected."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import provider_scene as ps

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_provider.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def scene_dir(gold, tmp_path_factory):
    root = str(tmp_path_factory.mktemp("scene"))
    ps.write_scene(ps.unpack_scene(gold), root)
    return root


@pytest.fixture()
def fake_h5py(monkeypatch):
    """h5py is not installed in this image: the scene's `.hdf5` files hold npy payloads read by a stand-in File class."""
    mod = types.ModuleType("h5py")
    mod.File = ps.FakeH5File
    monkeypatch.setitem(sys.modules, "h5py", mod)


def test_pose_conversion_and_orbit_poses(gold):
    from instance_nerf_b200.nerf.provider import nerf_matrix_to_ngp, rand_poses
    got = nerf_matrix_to_ngp(gold["ngp_in"], scale=0.4, offset=[0.1, -0.2, 0.3])
    assert got.dtype == np.float32 and np.array_equal(got, gold["ngp_out"])
    torch.manual_seed(9)
    np.testing.assert_allclose(rand_poses(6, "cpu", radius=2.5).numpy(), gold["rand_poses"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("split", ["train", "val", "test"])
def test_mask_dataset_loads_like_the_reference(gold, scene_dir, fake_h5py, split):
    from instance_nerf_b200.nerf.provider import NeRFMaskDataset
    ds = NeRFMaskDataset(ps.options(scene_dir, os.path.join(scene_dir, "mask3d.npy")), "cpu", type=split)
    tag = f"mask_{split}_"
    assert np.array_equal(ds.poses.numpy(), gold[tag + "poses"])
    assert np.array_equal(np.asarray(ds.intrinsics, dtype=np.float64), gold[tag + "intrinsics"])
    assert [ds.H, ds.W, ds.num_instances, ds.num_rays] == gold[tag + "meta"].tolist()
    assert ds.radius == float(gold[tag + "radius"])
    assert np.array_equal(np.asarray(ds.offset, dtype=np.float64), gold[tag + "offset"])
    if split == "test":
        assert ds.masks is None and ds.mask3d is None
    else:
        assert ds.masks.dtype == torch.long and np.array_equal(ds.masks.numpy(), gold[tag + "masks"])
        assert np.array_equal(ds.mask3d_coords.numpy(), gold[tag + "mask3d_coords"])
        assert np.array_equal(ds.mask3d_labels.numpy(), gold[tag + "mask3d_labels"])
    loader = ds.dataloader()
    assert [len(loader), int(loader.has_gt)] == gold[tag + "loader"].tolist() and loader._data is ds


def test_loader_contract_rand_pose_and_error_map(scene_dir, fake_h5py):
    """dataloader(): `size // rand_pose` extra indices select the random-pose branch (provider.py:627-631); the error map is
    [frames, 128 * 128] of ones for training sets only (:546-549); evaluation sets keep num_rays = -1."""
    from instance_nerf_b200.nerf.provider import NeRFMaskDataset
    opt = ps.options(scene_dir, None)
    opt.rand_pose, opt.error_map = 2, True
    ds = NeRFMaskDataset(opt, "cpu", type="train")
    loader = ds.dataloader()
    assert len(loader) == len(ds.poses) + len(ds.poses) // 2 and loader.has_gt and loader.batch_size == 1
    assert ds.error_map.shape == (len(ds.poses), 128 * 128) and bool((ds.error_map == 1).all())
    dv = NeRFMaskDataset(opt, "cpu", type="val")
    assert dv.error_map is None and dv.num_rays == -1 and len(dv.dataloader()) == 1
    dt = NeRFMaskDataset(opt, "cpu", type="test")
    assert dt.masks is None and not dt.dataloader().has_gt and len(dt.dataloader()) == len(dt.poses)


def test_segmap_formats_and_errors(tmp_path, scene_dir, gold):
    from instance_nerf_b200.nerf import provider
    m = ps.unpack_scene(gold)["masks"][0]
    np.save(tmp_path / "a.npy", m)
    np.savez(tmp_path / "b.npz", **{provider.SEGMAP_KEY: m})
    import cv2
    cv2.imwrite(str(tmp_path / "c.png"), np.stack([m, m, m], -1))
    for name in ("a.npy", "b.npz", "c.png"):
        assert np.array_equal(provider.read_segmap(str(tmp_path / name)), m), name
    with pytest.raises(RuntimeError):
        provider.read_segmap(str(tmp_path / "d.tiff"))
    if "h5py" not in sys.modules:
        with pytest.raises(ImportError):      # no silent fallback for the reference's native format
            provider.read_segmap(os.path.join(scene_dir, "segmaps/0000.hdf5"))
    # a scene without the instance count is rejected as the reference rejects it (provider.py:433-436)
    import json
    bad = tmp_path / "bad"
    bad.mkdir()
    t = json.load(open(os.path.join(scene_dir, "transforms.json")))
    del t["num_room_objects"]
    json.dump(t, open(bad / "transforms.json", "w"))
    with pytest.raises(RuntimeError):
        provider.NeRFMaskDataset(ps.options(str(bad), None), "cpu", type="test")


@pytest.mark.parametrize("split", ["train", "val"])
def test_rgb_dataset_loads_like_the_reference(gold, scene_dir, split):
    from instance_nerf_b200.nerf.provider import NeRFDataset
    ds = NeRFDataset(ps.options(scene_dir, None, rgb=True), "cpu", type=split)
    tag = f"rgb_{split}_"
    assert ds.mode == "colmap"
    assert np.array_equal(ds.poses.numpy(), gold[tag + "poses"])
    assert np.array_equal(ds.images.numpy(), gold[tag + "images"])
    assert np.array_equal(np.asarray(ds.intrinsics, dtype=np.float64), gold[tag + "intrinsics"])
    assert ds.radius == float(gold[tag + "radius"])


@pytest.mark.gpu
def test_mask_batches_on_device_match_reference(cuda, gold, scene_dir, fake_h5py):
    from instance_nerf_b200.nerf.provider import NeRFMaskDataset
    opt = ps.options(scene_dir, os.path.join(scene_dir, "mask3d.npy"))
    ds = NeRFMaskDataset(opt, cuda, type="train")
    assert ds.masks.is_cuda and ds.poses.is_cuda and ds.mask3d_coords.is_cuda
    inds = torch.from_numpy(gold["mask_train_batch_inds"]).to(cuda)
    b = ds.collate([2], inds=inds)
    assert b["H"] == ds.H and b["W"] == ds.W and b["file_name"] == str(gold["mask_train_batch_file_name"])
    np.testing.assert_allclose(b["rays_o"].cpu().numpy(), gold["mask_train_batch_rays_o"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(b["rays_d"].cpu().numpy(), gold["mask_train_batch_rays_d"], rtol=0, atol=2e-6)
    assert np.array_equal(b["masks"].cpu().numpy(), gold["mask_train_batch_masks"])
    assert b["mask3d_coords"] is ds.mask3d_coords and b["mask3d_labels"] is ds.mask3d_labels
    # own draw: 8x8 patches inside the frame, labels = the frame's map at those pixels
    torch.manual_seed(0)
    b2 = ds.collate([1])
    assert b2["rays_o"].shape == (1, opt.num_rays, 3) and b2["masks"].shape == (1, opt.num_rays)
    # validation split: the whole first frame
    dv = NeRFMaskDataset(opt, cuda, type="val")
    bv = dv.collate([0])
    np.testing.assert_allclose(bv["rays_d"].cpu().numpy(), gold["mask_val_batch_rays_d"], rtol=0, atol=2e-6)
    assert np.array_equal(bv["masks"].cpu().numpy(), gold["mask_val_batch_masks"])
    # random-pose branch (provider.py:591-606): a low-resolution full frame, no labels
    opt_r = ps.options(scene_dir, None)
    opt_r.rand_pose = 0
    br = NeRFMaskDataset(opt_r, cuda, type="train").collate([0])
    assert set(br) == {"H", "W", "rays_o", "rays_d"} and br["rays_o"].shape == (1, br["H"] * br["W"], 3)


@pytest.mark.gpu
def test_rgb_batches_on_device_match_reference(cuda, gold, scene_dir):
    from instance_nerf_b200.nerf.provider import NeRFDataset
    ds = NeRFDataset(ps.options(scene_dir, None, rgb=True), cuda, type="train")
    inds = torch.from_numpy(gold["rgb_train_batch_inds"]).to(cuda)
    b = ds.collate([1], inds=inds)
    np.testing.assert_allclose(b["rays_d"].cpu().numpy(), gold["rgb_train_batch_rays_d"], rtol=0, atol=2e-6)
    assert np.array_equal(b["images"].cpu().numpy(), gold["rgb_train_batch_images"])


@pytest.mark.gpu
def test_provider_feeds_the_training_step(cuda, gold, scene_dir, fake_h5py):
    """dataloader() -> MaskTrainStep.step: the batch keys are the ones the step consumes (nerf/utils.py:1287-1373)."""
    from test_field_gpu import build_model
    from instance_nerf_b200.nerf.provider import NeRFMaskDataset
    from instance_nerf_b200.nerf.trainer import MaskTrainStep
    opt = ps.options(scene_dir, os.path.join(scene_dir, "mask3d.npy"))
    ds = NeRFMaskDataset(opt, cuda, type="train")
    m, _ = build_model(cuda, ds.num_instances)
    m.train()
    step = MaskTrainStep(m, patch_size=opt.patch_size, label_regularization_weight=0.1, mask3d_loss_weight=0.5)
    torch.manual_seed(0)
    losses = []
    for i, data in enumerate(ds.dataloader()):
        losses.append(float(step.step(data)))
    assert len(losses) == len(ds.poses) and all(np.isfinite(losses))
