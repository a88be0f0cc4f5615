"""Development timing: ray ORDER of a full frame fed to the fused renderer (row-major vs pw x ph pixel patches per 32-ray group)."""
import sys
import torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from test_field_gpu import build_model
from helpers import make_rays

cuda = torch.device("cuda:0")
H, W = 480, 640
m, sc = build_model(cuda, 32, density_scale=10.0)
o, d = make_rays(sc, H, W)
o, d = o.to(cuda), d.to(cuda)
kw = dict(dt_gamma=1 / 128, max_steps=1024, T_thresh=1e-4)


def patch_perm(pw, ph):
    y, x = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    key = ((y // ph) * (W // pw) + (x // pw)) * (pw * ph) + (y % ph) * pw + (x % pw)
    return torch.argsort(key.reshape(-1)).to(cuda)


ref = None
for name, perm in (("row-major", None), ("8x4", patch_perm(8, 4)), ("16x2", patch_perm(16, 2)), ("4x8", patch_perm(4, 8)), ("32x1", patch_perm(32, 1))):
    oo, dd = (o, d) if perm is None else (o[perm].contiguous(), d[perm].contiguous())
    with torch.no_grad():
        for _ in range(3):
            r = m.render(oo[None], dd[None], staged=True, render_mask=True, perturb=False, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            r = m.render(oo[None], dd[None], staged=True, render_mask=True, perturb=False, **kw)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    img = r["image"][0]
    if perm is not None:
        full = torch.empty_like(img); full[perm] = img; img = full
    if ref is None:
        ref = img
    tiles = int(m._work_counter[1].item())
    samples = int(m._work_counter[2:4].view(torch.int64).item())
    print(f"{name:10s} {ms:.3f} ms/frame  {o.shape[0] / ms / 1e3:.2f} Mrays/s  fill {samples / (tiles * 128):.3f}  max|img - row-major| {float((img - ref).abs().max()):.2e}")
