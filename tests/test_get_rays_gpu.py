"""GPU parity of inerf_get_rays (nerf/utils.py:56-140) against the torch restatement of the reference's get_rays.

Tolerance: fp32, IEEE divisions / sqrt; only the 3x3 rotation may be ordered / contracted differently by torch's matmul,
so directions agree to 2e-6 absolute (unit vectors, ~8 ulp worst case observed << that); origins are exact copies.  The fused
near/far is bit-identical to near_far_from_aabb run on the kernel's own rays."""
import numpy as np
import pytest
import torch

from helpers import bits_equal, scene_arrays
from instance_nerf_b200 import synthetic
from oracle import host_oracle

pytestmark = pytest.mark.gpu


def _poses(n):
    sc, *_ = scene_arrays(16, 8.0, 0)
    return torch.from_numpy(synthetic.camera_poses(sc, n, 1))


@pytest.mark.parametrize("H,W,B", [(48, 64, 1), (33, 57, 3), (480, 640, 2)])
def test_full_frame(cuda, H, W, B):
    from instance_nerf_b200.nerf.utils import get_rays
    poses = _poses(B)
    intr = synthetic.intrinsics(H, W)
    want = host_oracle.get_rays(poses, intr, H, W)
    got = get_rays(poses.to(cuda), intr, H, W)
    assert got["rays_o"].shape == (B, H * W, 3) and "inds" not in got
    assert bits_equal(got["rays_o"], want["rays_o"].contiguous())
    torch.testing.assert_close(got["rays_d"].cpu(), want["rays_d"], rtol=0, atol=2e-6)
    n = got["rays_d"].norm(dim=-1)
    assert float((n - 1).abs().max()) < 1e-6


@pytest.mark.parametrize("patch", [1, 8])
def test_sampled_pixels_and_fused_near_far(cuda, patch):
    from instance_nerf_b200 import raymarching as rm
    from instance_nerf_b200.nerf.utils import get_rays
    H, W, B, N = 120, 160, 2, 4096
    poses = _poses(B)
    intr = synthetic.intrinsics(H, W)
    g = torch.Generator(device=cuda).manual_seed(3)
    aabb = torch.tensor([-8.0] * 3 + [8.0] * 3, device=cuda)
    got = get_rays(poses.to(cuda), intr, H, W, N=N, patch_size=patch, generator=g, aabb=aabb, min_near=0.2)
    inds = got["inds"]
    assert inds.shape == (B, N) and int(inds.min()) >= 0 and int(inds.max()) < H * W
    if patch > 1:   # patch-major, then row, then column (utils.py:83-100)
        i0 = inds[0].view(-1, patch, patch).cpu()
        assert torch.equal(i0[:, :, 1:] - i0[:, :, :-1], torch.ones_like(i0[:, :, 1:]))
        assert torch.equal(i0[:, 1:, :] - i0[:, :-1, :], torch.full_like(i0[:, 1:, :], W))
        assert int((i0[:, 0, 0] // W).max()) < H - patch and int((i0[:, 0, 0] % W).max()) < W - patch
    full = host_oracle.get_rays(poses, intr, H, W)
    want_d = torch.gather(full["rays_d"], 1, inds.cpu()[..., None].expand(B, N, 3))
    torch.testing.assert_close(got["rays_d"].cpu(), want_d, rtol=0, atol=2e-6)
    n1, f1 = rm.near_far_from_aabb(got["rays_o"].view(-1, 3), got["rays_d"].view(-1, 3), aabb, 0.2)
    assert bits_equal(got["nears"].view(-1), n1) and bits_equal(got["fars"].view(-1), f1)


def test_errors(cuda):
    from instance_nerf_b200._lib import InerfError, call, ptr
    from instance_nerf_b200.nerf.utils import get_rays
    with pytest.raises(RuntimeError):
        get_rays(_poses(1), synthetic.intrinsics(8, 8), 8, 8)          # CPU tensor: no fallback
    p = _poses(1).to(cuda).float().contiguous()
    o = torch.empty(64, 3, device=cuda)
    with pytest.raises(InerfError):                                     # inds NULL but N != H*W
        call("inerf_get_rays", ptr(p), 1, 1.0, 1.0, 4.0, 4.0, 8, 8, None, 63, ptr(o), ptr(o), None, 0.2, None, None, 0)


@pytest.mark.parametrize("case", ["full", "uniform", "patch", "emap"])
def test_get_rays_matches_reference_golden(cuda, case):
    """inerf_get_rays against the REFERENCE's own get_rays (nerf/utils.py:56-140, imported unmodified and run on the CPU by
    tests/golden/make_golden_host.py): full frame, uniform, 8x8 patches and the error-map branch.  The pixel choice is injected
    (the reference drew it from torch's CPU generator); origins are exact copies, directions agree to 2e-6 (the 3x3 rotation
    may be ordered / contracted differently by the CPU matmul)."""
    import os
    from instance_nerf_b200.nerf.utils import get_rays
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_host.npz"))
    B, H, W, N, patch, seed, emap = (int(v) for v in g[f"rays_{case}_cfg"])
    poses = torch.from_numpy(g["poses"])[:B].to(cuda)
    inds = torch.from_numpy(g[f"rays_{case}_inds"]).to(cuda) if N > 0 else None
    if inds is not None and not emap:
        inds = inds[0]                       # one pixel set shared by every pose (utils.py:104, 108)
    got = get_rays(poses, synthetic.intrinsics(H, W), H, W, N=N, patch_size=patch, inds=inds)
    assert bits_equal(got["rays_o"], g[f"rays_{case}_rays_o"])
    torch.testing.assert_close(got["rays_d"].cpu(), torch.from_numpy(g[f"rays_{case}_rays_d"]), rtol=0, atol=2e-6)
    if N > 0:
        assert torch.equal(got["inds"].cpu(), torch.from_numpy(g[f"rays_{case}_inds"]))
    # the product's own sampler (CUDA generator: other draws than the CPU reference) keeps the reference's structure
    if case == "emap":
        em = ((torch.arange(B * 128 * 128) % 97).float() + 1).view(B, -1) / 97.0
        r = get_rays(poses, synthetic.intrinsics(H, W), H, W, N=N, error_map=em.to(cuda))
        assert r["inds"].shape == (B, N) and r["inds_coarse"].shape == (B, N) and int(r["inds"].max()) < H * W
        cx, cy = r["inds_coarse"] // 128, r["inds_coarse"] % 128
        ix, iy = r["inds"] // W, r["inds"] % W
        assert bool(((ix.float() - cx * (H / 128)).abs() <= H / 128 + 1).all()) and bool(((iy.float() - cy * (W / 128)).abs() <= W / 128 + 1).all())
