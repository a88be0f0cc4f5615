import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import numpy as np, torch
from test_field_gpu import build_model
from instance_nerf_b200 import synthetic
from oracle import host_oracle
from instance_nerf_b200.nerf.trainer import MaskTrainStep
cuda = torch.device("cuda:0")
p = torch.nn.Parameter(torch.zeros(10, device=cuda)); p.grad = torch.ones_like(p)
opt = torch.optim.Adam([p], fused=True); v0 = p._version; opt.step(); print("fused adam version bump:", v0, "->", p._version)
K = 16
m, sc = build_model(cuda, K, density_scale=10.0)
H, W = 96, 128
poses = torch.from_numpy(synthetic.camera_poses(sc, 1, 1))
r = host_oracle.get_rays(poses, synthetic.intrinsics(H, W), H, W, N=1024, patch_size=8, generator=torch.Generator().manual_seed(0))
o, d = r["rays_o"].reshape(-1, 3).contiguous(), r["rays_d"].reshape(-1, 3).contiguous()
labels = torch.from_numpy(sc.first_hit_labels(o.numpy().astype(np.float64), d.numpy().astype(np.float64)))
print("labels:", np.unique(labels.numpy(), return_counts=True))
data = {"rays_o": o[None].to(cuda), "rays_d": d[None].to(cuda), "masks": labels[None].to(cuda)}
tr = MaskTrainStep(m, lr=1e-2, fp16=True, label_regularization_weight=0.1)
for i in range(12):
    w_before = m.mask_net[2].weight.detach().clone()
    loss = tr.step(data)
    g = m.mask_net[2].weight.grad
    tg = m.encoder_mask.embeddings.grad
    print(i, "loss", float(loss), "scale", tr.scaler.get_scale(), "gw2 finite", bool(torch.isfinite(g).all()), float(g.abs().max()),
          "table grad finite", bool(torch.isfinite(tg).all()), float(tg.abs().max()), "dw", float((m.mask_net[2].weight - w_before).abs().max()),
          "ver", m.mask_net[2].weight._version)
