"""Generates tests/golden/ref_kernels.npz: outputs of the REFERENCE's own CUDA kernels (oracle/_ref, built from
/root/reference by oracle/build_ref.sh) on small seeded inputs.  Run on a GPU box:

    gpurun -- python tests/golden/make_golden_gpu.py        # writes gpurun_out/golden/ref_kernels.npz
    cp gpurun_out/golden/ref_kernels.npz tests/golden/

The CPU test-suite (tests/test_oracle_cpu.py) pins the oracle (oracle/*.c, oracle/*.py) against this file."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import adversarial_rays, canonicalize, make_rays, scene_arrays  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle.field_oracle import level_offsets  # noqa: E402

ref = ref_loader.load()
dev = torch.device("cuda:0")
out = {}

for tag, bound, dt_gamma, max_steps in (("b8", 8.0, 1 / 128, 1024), ("b3", 3.0, 1 / 128, 256)):
    sc, cascade, grid, bits = scene_arrays(16, bound, 0)
    o, d = make_rays(sc, 12, 16)
    ao, ad = adversarial_rays(bound)
    o, d = torch.cat([o, ao]).to(dev), torch.cat([d, ad]).to(dev)
    N = o.shape[0]
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, device=dev)
    nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
    ref.raymarching.near_far_from_aabb(o, d, aabb, N, 0.2, nears, fars)
    noises = torch.rand(N, generator=torch.Generator().manual_seed(2)).to(dev)
    bits_t = torch.from_numpy(bits).to(dev)
    M = N * max_steps
    x, dd, dl = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    ref.raymarching.march_rays_train(o, d, bits_t, bound, dt_gamma, max_steps, N, cascade, 128, M, nears, fars, x, dd, dl, rays, counter, noises)
    torch.cuda.synchronize()
    c_rays, cx, cd, cl = canonicalize(rays, x, dd, dl)
    out.update({f"{tag}_rays_o": o.cpu().numpy(), f"{tag}_rays_d": d.cpu().numpy(), f"{tag}_nears": nears.cpu().numpy(),
                f"{tag}_fars": fars.cpu().numpy(), f"{tag}_noises": noises.cpu().numpy(), f"{tag}_rays": c_rays, f"{tag}_xyzs": cx,
                f"{tag}_deltas": cl, f"{tag}_counter": counter.cpu().numpy(), f"{tag}_cfg": np.array([bound, dt_gamma, max_steps, cascade])})
    if tag == "b8":
        # inference marching of 3 steps from perturbed starts
        alive = torch.arange(0, N, 2, dtype=torch.int32, device=dev)
        rays_t = (nears + 0.37).contiguous()
        n_alive, n_step = alive.shape[0], 3
        ix, idd, idl = torch.zeros(n_alive * n_step, 3, device=dev), torch.zeros(n_alive * n_step, 3, device=dev), torch.zeros(n_alive * n_step, 2, device=dev)
        ref.raymarching.march_rays(n_alive, n_step, alive, rays_t, o, d, bound, dt_gamma, max_steps, cascade, 128, bits_t, nears, fars, ix, idd, idl,
                                   torch.zeros(n_alive, device=dev))
        out.update({"inf_alive": alive.cpu().numpy(), "inf_rays_t": rays_t.cpu().numpy(), "inf_xyzs": ix.cpu().numpy(), "inf_deltas": idl.cpu().numpy()})
        # compositing forward / backward with masks (K = 5) and without
        total = int(counter[0])
        g = torch.Generator().manual_seed(5)
        K = 5
        sig = torch.exp(torch.randn(total, generator=g) * 1.5 + 2.0).to(dev)
        rgb = torch.rand(total, 3, generator=g).to(dev)
        msk = (torch.randn(total, K, generator=g) * 2).to(dev)
        crays = torch.from_numpy(c_rays).to(dev)
        cdl = torch.from_numpy(cl).to(dev)
        ws, dp, im, mo = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, 3, device=dev), torch.empty(N, K, device=dev)
        ref.raymarching.composite_rays_with_masks_train_forward(sig, rgb, msk, cdl, crays, total, N, K, 1e-4, ws, dp, im, mo)
        gws, gim, gmo = torch.randn(N, generator=g).to(dev), torch.randn(N, 3, generator=g).to(dev), torch.randn(N, K, generator=g).to(dev)
        gs, gr, gm, acc = torch.zeros(total, device=dev), torch.zeros(total, 3, device=dev), torch.zeros(total, K, device=dev), torch.zeros(N, K, device=dev)
        ref.raymarching.composite_rays_with_masks_train_backward(gws, gim, gmo, sig, rgb, msk, cdl, crays, ws, im, mo, total, N, K, 1e-4, gs, gr, acc, gm)
        ws2, dp2, im2 = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, 3, device=dev)
        ref.raymarching.composite_rays_train_forward(sig, rgb, cdl, crays, total, N, 1e-4, ws2, dp2, im2)
        out.update({k: v.cpu().numpy() for k, v in dict(cmp_sig=sig, cmp_rgb=rgb, cmp_msk=msk, cmp_ws=ws, cmp_depth=dp, cmp_image=im, cmp_mask_out=mo,
                                                        cmp_gws=gws, cmp_gim=gim, cmp_gmo=gmo, cmp_gs=gs, cmp_gr=gr, cmp_gm=gm, cmp_ws_plain=ws2,
                                                        cmp_image_plain=im2).items()})

# morton / packbits
g = torch.Generator().manual_seed(0)
coords = torch.randint(0, 128, (512, 3), generator=g, dtype=torch.int32).to(dev)
idx = torch.empty(512, dtype=torch.int32, device=dev)
ref.raymarching.morton3D(coords, 512, idx)
pg = (torch.rand(1, 4096, generator=g) * 20 - 2).to(dev)
pb = torch.empty(512, dtype=torch.uint8, device=dev)
ref.raymarching.packbits(pg, 512, 5.0, pb)
out.update({"morton_coords": coords.cpu().numpy(), "morton_idx": idx.cpu().numpy(), "pack_grid": pg.cpu().numpy(), "pack_bits": pb.cpu().numpy()})

# hash-grid encode (fp32) with the real level geometry of bound = 2 and a seeded table; SH degree 4
offsets, pls = level_offsets(desired_resolution=2048 * 2)
T = int(offsets[-1])
table = ((torch.rand(T, 2, generator=torch.Generator().manual_seed(3)) - 0.5)).to(dev)
x01 = torch.rand(300, 3, generator=g).to(dev)
x01[:3] = torch.tensor([[0.0, 1.0, 0.5], [1.0, 1.0, 1.0], [1.2, 0.5, 0.5]]).to(dev)
enc = torch.empty(16, 300, 2, device=dev)
off_t = torch.from_numpy(offsets).to(dev)
ref.gridencoder.grid_encode_forward(x01, table, off_t, enc, 300, 3, 2, 16, float(np.log2(pls)), 16, None, 0, False, 0)
ggrad = torch.randn(16, 300, 2, generator=g).to(dev)
gtab = torch.zeros_like(table)
ref.gridencoder.grid_encode_backward(ggrad, x01, table, off_t, gtab, 300, 3, 2, 16, float(np.log2(pls)), 16, None, None, 0, False, 0)
nz = torch.nonzero(gtab.abs().sum(1)).squeeze(1)
dirs = torch.nn.functional.normalize(torch.randn(256, 3, generator=g), dim=-1).to(dev)
sh = torch.empty(256, 16, device=dev)
ref.shencoder.sh_encode_forward(dirs, sh, 256, 3, 4, None)
# per-level scales as CUDA's exp2f produces them (torch.exp2 on the device is the same exp2f), for the bounds the tests use
for b in (1, 2, 3, 8):
    _, pls_b = level_offsets(desired_resolution=2048 * b)
    S_b = torch.tensor(float(np.log2(pls_b)), device=dev, dtype=torch.float32)
    out[f"enc_scales_b{b}"] = (torch.exp2(torch.arange(16, device=dev, dtype=torch.float32) * S_b) * 16 - 1).cpu().numpy()
out.update({"enc_x01": x01.cpu().numpy(), "enc_out": enc.permute(1, 0, 2).reshape(300, 32).cpu().numpy(), "enc_table_seed": np.array([3]),
            "enc_grad": ggrad.permute(1, 0, 2).reshape(300, 32).cpu().numpy(), "enc_gtab_rows": nz.cpu().numpy().astype(np.int64),
            "enc_gtab_vals": gtab[nz].cpu().numpy(), "sh_dirs": dirs.cpu().numpy(), "sh_out": sh.cpu().numpy()})

os.makedirs(os.path.join(ROOT, "gpurun_out", "golden"), exist_ok=True)
path = os.path.join(ROOT, "gpurun_out", "golden", "ref_kernels.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if v.size > 4000})
