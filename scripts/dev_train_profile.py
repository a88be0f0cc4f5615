"""Development: where does the training step's wall time go (host vs device)?"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from instance_nerf_b200 import synthetic
from oracle import host_oracle
from instance_nerf_b200.nerf.trainer import MaskTrainStep
dev = torch.device("cuda:0")
model, scene, poses = bench.build_scene_and_model(dev)
tr = MaskTrainStep(model, label_regularization_weight=0.1)
g = torch.Generator().manual_seed(100)
r = host_oracle.get_rays(poses[0][None], synthetic.intrinsics(480, 640), 480, 640, N=4096, patch_size=8, generator=g)
o, d = r["rays_o"].reshape(-1, 3).contiguous(), r["rays_d"].reshape(-1, 3).contiguous()
labels = torch.from_numpy(scene.first_hit_labels(o.numpy().astype(np.float64), d.numpy().astype(np.float64)))
data = {"rays_o": o[None].to(dev), "rays_d": d[None].to(dev), "masks": labels[None].to(dev)}
for _ in range(5): tr.step(data)
torch.cuda.synchronize()
def timed(fn, name):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name:28s} host {1e3*(t1-t0):7.2f} ms   +drain {1e3*(t2-t1):7.2f} ms"); return out
for rep in range(2):
    model.train()
    timed(lambda: tr.optimizer.zero_grad(set_to_none=False), "zero_grad")
    def fwd():
        with torch.autocast("cuda", dtype=torch.float16):
            return tr.train_step(data)
    _, _, loss = timed(fwd, "train_step (fwd)")
    timed(lambda: tr.scaler.scale(loss).backward(), "backward")
    timed(lambda: tr.scaler.step(tr.optimizer), "scaler.step")
    timed(lambda: tr.scaler.update(), "scaler.update")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): tr.step(data)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
