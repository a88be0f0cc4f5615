"""Development: data-parallel graphed step, per-phase timing and overflow bookkeeping (torchrun, N >= 2)."""
import os, sys, time, json
from datetime import timedelta
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from instance_nerf_b200.nerf.trainer import MaskTrainStep

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev, timeout=timedelta(seconds=120))
n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
model, scene, poses = bench.build_scene_and_model(dev)
tr = MaskTrainStep(model, lr=1e-2, fp16=True, patch_size=8, label_regularization_weight=0.1, dt_gamma=bench.DT_GAMMA, max_steps=bench.MAX_STEPS,
                   T_thresh=bench.T_THRESH, data_parallel=True, cuda_graph=True)
batches = bench.train_batches(dev, scene, poses, n_rays, rank, world, n=4)
log = []
for i in range(24):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = tr.step(batches[i % 4])
    e1.record(); torch.cuda.synchronize()
    log.append(dict(i=i, ms=round(e0.elapsed_time(e1), 3), wall_ms=round((time.perf_counter() - t0) * 1e3, 3), total=tr.last_total, seen=tr._samples_seen,
                    captures=tr.graph_captures, replays=tr.graph_replays, extra=float(tr.bucket.extra[0]), loss=round(float(loss), 4),
                    budget=(tr._graph[3] if tr._graph else None)))
# phases of one replayed step
if tr._graph is not None:
    g, static, loss, budget, counter, g2, _keep = tr._graph
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for rep in range(3):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        ev[0].record(); g.replay(); ev[1].record(); tr._exchange(); ev[2].record(); g2.replay(); ev[3].record(); torch.cuda.synchronize()
        log.append(dict(phase_ms=[round(ev[k].elapsed_time(ev[k + 1]), 3) for k in range(3)]))
if True:
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/dp_debug_rank{rank}.json", "w") as f:
        json.dump(log, f, indent=0)
dist.destroy_process_group()
