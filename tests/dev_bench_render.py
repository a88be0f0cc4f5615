"""Development timing of the fused renderer (not the official bench)."""
import sys, time
import torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from test_field_gpu import build_model
from helpers import make_rays

cuda = torch.device("cuda:0")
for ds in (10.0,):
    m, sc = build_model(cuda, 32, density_scale=ds)
    o, d = make_rays(sc, 480, 640)
    o, d = o.to(cuda), d.to(cuda)
    kw = dict(dt_gamma=1 / 128, max_steps=1024, T_thresh=1e-4)
    with torch.no_grad():
        for _ in range(3):
            r = m.render(o[None], d[None], staged=True, render_mask=True, perturb=False, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r = m.render(o[None], d[None], staged=True, render_mask=True, perturb=False, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        # sample count via training marcher
        m.train()
        from instance_nerf_b200 import raymarching as rm
        nears, fars = rm.near_far_from_aabb(o, d, m.aabb_train, m.min_near)
        c = torch.zeros(2, dtype=torch.int32, device=cuda)
        rm.march_rays_train(o, d, m.bound, m.density_bitfield, m.cascade, 128, nears, fars, c, -1, False, 128, True, 1 / 128, 1024)
        m.eval()
        print(f"density_scale={ds}: {ms:.2f} ms/frame, {o.shape[0] / ms / 1e3:.2f} Mrays/s, full-march samples/ray={c[0].item() / o.shape[0]:.1f}, ws mean={r['image'].mean().item():.3f}")
