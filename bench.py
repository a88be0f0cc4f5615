#!/usr/bin/env python
"""bench.py -- instance-field render throughput (BASELINE.json metric: Mrays/s) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-cpu-baseline]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY.md section 8d "c2"): 640x480 frames (307 200 rays each) of the synthetic
3D-FRONT-shaped room, bound 8 -> 4-cascade 128^3 occupancy grid, two 16-level 2^19-entry hash tables, 32 instance classes,
dt_gamma 1/128, max_steps 1024, T_thresh 1e-4; seeded random weights / tables (no network, no dataset).

A "step" = one frame rendered through NeRFNetwork.render (the call MaskTrainer.test_step makes, nerf/utils.py:1421): near/far
+ ONE fused launch (march + hash gathers x2 + tcgen05 MLP + composite) + the bg / depth tail.  Every step renders a different
camera pose, and L2 (126 MB) is flushed between timed steps by writing a 512 MB buffer (outside the event pair).
At N > 1 every rank renders its own frame each step (weak scaling, rays sharded by frame) and the finished tiles are
gathered with ONE NCCL all_gather per step, issued asynchronously so it overlaps the next frame's render (the last
step waits for its gather inside the timed region).

Keys: value = device-timed Mrays/s with rays resident in HBM; e2e = same through the public API from pinned HOST rays to
pinned HOST results: every step copies that frame's rays host->device and every rank reads ITS OWN finished frame
(image | depth | K logits) back to pinned host memory on a side stream, double buffered, so the copy of frame i overlaps the
render of frame i + 1; ONE event pair brackets the whole e2e region, L2 flushes and the final copy join included; roofline = the fused kernel's algorithmic bytes / its own
CUDA-event time vs the measured HBM peak; cpu_baseline = the CPU oracle port of the reference's PyTorch (non-cuda_ray)
renderer on a bounded sample of the same frame, timed on this box's host cores.

--impl reference runs ONLY that CPU path (the reference arm): no CUDA code of this repo is touched.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "instance_field_render_Mrays_per_s"
UNIT = "Mrays/s"
W_IMG, H_IMG, K_INST, BOUND = 640, 480, 32, 8.0
DT_GAMMA, MAX_STEPS, T_THRESH = 1.0 / 128, 1024, 1e-4
N_POSES = 16
WORKLOAD = "c2: 640x480 instance-field render, 4x128^3 occupancy grid, 2x(16-level 2^19 hash grid), K=32, synthetic room"


def build_scene_and_model(device=None):
    import torch
    from instance_nerf_b200 import synthetic
    from instance_nerf_b200.nerf.network_mask import NeRFNetwork

    torch.manual_seed(0)
    model = NeRFNetwork(bound=BOUND, cuda_ray=True, num_instances=K_INST, density_scale=1, density_thresh=10)
    synthetic.randomize_tables(model, 0)
    scene = synthetic.RoomScene(K_INST, BOUND, 0)
    synthetic.install_scene(model, scene)
    poses = torch.from_numpy(synthetic.camera_poses(scene, N_POSES, 1))
    if device is not None:
        model = model.to(device)
    return model.eval(), scene, poses


def frame_rays(poses, i):
    from instance_nerf_b200 import synthetic
    r = synthetic.get_rays(poses[i:i + 1], synthetic.intrinsics(H_IMG, W_IMG), H_IMG, W_IMG)
    return r["rays_o"].reshape(-1, 3).contiguous(), r["rays_d"].reshape(-1, 3).contiguous()


# ------------------------------------------------------------------------------------------------ clocks --
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.t.join(timeout=2)
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------- CPU reference arm --
def make_cpu_reference(n_rays: int, poses=None, model=None):
    """The reference's PyTorch (non-cuda_ray) renderer -- NeRFMaskRenderer.run / staged render, mask_renderer.py:89-231,
    565-584 -- as restated in oracle/field_oracle.py (pinned against the reference's own outputs, tests/golden/ref_run.npz),
    set up on `n_rays` rays strided over frame 0 of the workload.  -> (render thunk, n_rays, cores)"""
    import torch
    from oracle import field_oracle as fo

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if model is None:
        model, _, poses = build_scene_and_model(None)
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    field = fo.OracleField(sd, BOUND, K_INST, density_scale=1.0)
    o, d = frame_rays(poses.cpu(), 0)
    stride = max(1, o.shape[0] // n_rays)
    o, d = o[::stride][:n_rays].contiguous(), d[::stride][:n_rays].contiguous()

    def render():
        return field.render(o[None], d[None], max_ray_batch=4096, render_mask=True, num_steps=128, bg_color=1)

    return render, o.shape[0], cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    render, n_rays, cores = make_cpu_reference(8192)
    for _ in range(args.warmup):
        render()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        render()
    total = time.perf_counter() - t0
    value = n_rays * args.steps / total / 1e6
    sample = f"{n_rays} rays strided over frame 0 per step, 128 uniform samples/ray (the reference's non-cuda_ray sampler), K=32"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "rays_per_step": n_rays, "samples_per_ray": 128},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------- GPU arm --
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from instance_nerf_b200 import _lib
    _lib.lib()  # fail loudly if libinerf_b200.so is missing

    model, scene, poses = build_scene_and_model(dev)
    N = H_IMG * W_IMG
    K = K_INST
    # Work units of one step = `world` whole frames.  Rank r renders the row blocks r, r + world, ... of EVERY frame of the step
    # (H / world rows of each), not one frame of its own: cost is proportional to marched samples, which vary by +-25 % from
    # pose to pose, and the per-step tile gather is a synchronisation point -- interleaved rows give every rank the same mix
    # (SURVEY.md section 8e).  Rows stay whole, so the marcher's 32-ray patches remain runs of neighbouring pixels.
    from instance_nerf_b200 import parallel
    block_rows = 4
    if H_IMG % (world * block_rows):
        raise SystemExit(f"bench.py: {H_IMG} rows do not split into {block_rows}-row blocks over {world} ranks")
    mine = parallel.shard_rows(H_IMG, W_IMG, rank, world, block_rows=block_rows)    # this rank's pixel ids inside a frame
    n_groups = max(1, N_POSES // world)
    host_rays = []
    for g in range(n_groups):
        os_, ds_ = [], []
        for j in range(world):
            o, d = frame_rays(poses, (g * world + j) % N_POSES)
            os_.append(o[mine])
            ds_.append(d[mine])
        host_rays.append((torch.cat(os_).contiguous().pin_memory(), torch.cat(ds_).contiguous().pin_memory()))
    dev_rays = [(o.to(dev), d.to(dev)) for o, d in host_rays]
    kw = dict(staged=True, render_mask=True, perturb=False, dt_gamma=DT_GAMMA, max_steps=MAX_STEPS, T_thresh=T_THRESH, bg_color=1)
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    # image 3 | depth 1 | logits K; double buffered so the NCCL gather of frame i overlaps the render of frame i + 1
    tiles = [torch.empty(N, 4 + K, dtype=torch.float32, device=dev) for _ in range(2)]
    gathered = [torch.empty(world * N, 4 + K, dtype=torch.float32, device=dev) for _ in range(2)] if world > 1 else None
    handles = [None, None]
    copy_done = [None, None]
    copy_stream = torch.cuda.Stream(device=dev)

    # kernel-only timing hook around the fused launch (events on the launching stream)
    kern_events, samples_seen = [], []
    launches = {"n": 0}
    fused = model._render_fused

    def timed_fused(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fused(*a, **k)
        e1.record()
        kern_events.append((e0, e1))
        samples_seen.append(model._work_counter.clone())
        launches["n"] += 3          # inerf_near_far_from_aabb + k_first_hit + k_render_fused
        return out

    model._render_fused = timed_fused

    def step(i, e2e=False, out_host=None):
        p = i % n_groups
        if e2e:
            o = host_rays[p][0].to(dev, non_blocking=True)
            d = host_rays[p][1].to(dev, non_blocking=True)
        else:
            o, d = dev_rays[p]
        b = i & 1
        if handles[b] is not None:        # the gather that last used this buffer pair must have finished (stream-side wait)
            handles[b].wait()
            handles[b] = None
        if copy_done[b] is not None:      # ... and so must the device->host copy of the frame rendered two steps ago
            torch.cuda.current_stream().wait_event(copy_done[b])
            copy_done[b] = None
        tile = tiles[b]
        with torch.no_grad():
            r = model.render(o[None], d[None], **kw)
        tile[:, 0:3] = r["image"][0]
        tile[:, 3] = r["depth"][0]
        tile[:, 4:] = r["instance_mask_logits"][0]
        if world > 1:
            handles[b] = dist.all_gather_into_tensor(gathered[b], tile, async_op=True)
        if e2e:
            # every rank reads ITS OWN frame back over its own PCIe link, on a side stream, so the copy of frame i overlaps the
            # render of frame i + 1 (double-buffered device tiles and pinned host buffers); the NCCL gather still runs
            ready = torch.cuda.Event()
            ready.record()
            copy_stream.wait_event(ready)
            with torch.cuda.stream(copy_stream):
                out_host[b].copy_(tile, non_blocking=True)
                copy_done[b] = torch.cuda.Event()
                copy_done[b].record()

    def drain():
        for b in range(2):
            if handles[b] is not None:
                handles[b].wait()
                handles[b] = None
            if copy_done[b] is not None:
                torch.cuda.current_stream().wait_event(copy_done[b])
                copy_done[b] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    drain()
    barrier()
    kern_events.clear(); samples_seen.clear(); launches["n"] = 0

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()

    # ---- device-resident timed region: K steps, L2 flushed between them -----------------------------------------
    ev = []
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush_buf.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(args.warmup + i)
        if i == args.steps - 1:
            drain()
        e1.record()
        ev.append((e0, e1))
    barrier()
    wall_s = time.perf_counter() - t_wall0
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    kern_ms = [a.elapsed_time(b) for a, b in kern_events]
    n_samples = [int(s[2:4].view(torch.int64).item()) for s in samples_seen]
    n_tiles = [int(s[1].item()) for s in samples_seen]
    gpu_launches = launches["n"]

    # ---- end-to-end region: pinned host rays -> render -> pinned host results ----------------------------------------
    out_host = [torch.empty(N, 4 + K, dtype=torch.float32).pin_memory() for _ in range(2)]
    for i in range(min(2, args.warmup)):
        step(i, True, out_host)
    drain()
    barrier()
    # one event pair around the whole region: L2 flushes and the final join of the copy stream are INSIDE it
    e2e0, e2e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e0.record()
    for i in range(args.steps):
        flush_buf.fill_(i & 0xFF)
        step(args.warmup + i, True, out_host)
    drain()
    e2e1.record()
    barrier()
    e2e_ms = e2e0.elapsed_time(e2e1)
    clk = clocks.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        peaks = {}
        ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(ppath):
            peaks = json.load(open(ppath))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # algorithmic bytes of the fused kernel (DESIGN.md "roofline"): per composited sample 2 tables x 16 levels x 8 corners x
        # 4 B (fp16 pair) = 1024 B of gathers; per ray 32 B in (o, d, near, far) + (5 + K) * 4 B out.  Nothing else touches HBM.
        avg_samples = sum(n_samples) / max(1, len(n_samples))
        alg_bytes = avg_samples * 1024.0 + N * (32.0 + (5 + K) * 4.0)
        avg_kern_ms = sum(kern_ms) / max(1, len(kern_ms))
        achieved = alg_bytes / (avg_kern_ms * 1e-3) / 1e9
        mlp_flops = avg_samples * (6144 + 12544 + 14208 + 128 * K)
        ms_per_step = total_ms / args.steps
        value = world * N / (ms_per_step * 1e-3) / 1e6
        h2d = 2 * N * 3 * 4
        d2h = out_host[0].numel() * 4
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": N, "samples_per_ray": avg_samples / N,
                       "tile_fill": (sum(n_samples) / max(1, 128 * sum(n_tiles))), "l2": "flushed between timed steps (512 MB write)",
                       "parallelism": (f"{world} frames per step, every frame's rows interleaved over {world} ranks (balanced by samples), async NCCL "
                                       f"all_gather of the finished row blocks overlapped with the next step") if world > 1 else "single GPU",
                       "wall_s_timed_region_incl_flush": wall_s},
            "e2e": {"value": world * N / (e2e_ms / args.steps * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": gpu_launches,
            "roofline": {"kernel": "k_render_fused", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": None, "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "kernel_ms": avg_kern_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "mlp_tflops": mlp_flops / (avg_kern_ms * 1e-3) / 1e12},
            "clocks": clk,
        }
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                line["roofline"]["traffic"] = json.load(open(tpath)).get("k_render_fused_dram_bytes_per_launch")
            except Exception:
                pass
        if world == 1 and not args.no_train:
            # second half of BASELINE.json's metric ("train-step ms @ 4096 rays"), measured in the same run (config c3)
            try:
                tr = measure_train(dev, 0, 1, 20, 5, 4096, model, scene, poses)
            except Exception as e:   # the headline line must survive a failed graph capture: redo the figure eagerly and say so
                print(f"[bench] graphed train step failed ({type(e).__name__}: {e}); measuring the eager step", file=sys.stderr)
                torch.cuda.synchronize()
                os.environ["INERF_NO_GRAPH"] = "1"
                tr = measure_train(dev, 0, 1, 20, 5, 4096, model, scene, poses)
            line["train_step"] = {"ms": tr["ms"], "ms_median": tr["ms_median"], "unit": "ms", "rays": 4096, "samples": tr["samples"],
                                  "workload": "c3: 4096 rays (8x8 patches) x max_steps 1024, K=32, fp16 autocast, CE + label smoothness, backward, Adam",
                                  "cuda_graph": tr["cuda_graph"], "graph_replays": tr["graph_replays"], "graph_captures": tr["graph_captures"]}
        if world == 1 and not args.no_cpu_baseline:
            render, n_cpu, cores = make_cpu_reference(32768)
            t0 = time.perf_counter()
            render()
            secs = time.perf_counter() - t0
            v = n_cpu / secs / 1e6
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n_cpu} rays strided over frame 0, 128 uniform samples/ray (reference non-cuda_ray sampler), {secs:.1f} s"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- train arm --
def measure_train(dev, rank, world, steps, warmup, n_rays, model=None, scene=None, poses=None):
    """Times `steps` full optimisation steps (MaskTrainStep.step: render -> loss -> backward -> [all_reduce] -> Adam) with CUDA
    events, L2 flushed between steps.  -> dict(ms, samples, loss, clocks)"""
    import numpy as np
    import torch
    import torch.distributed as dist
    from instance_nerf_b200 import synthetic
    from instance_nerf_b200.nerf.trainer import MaskTrainStep

    if model is None:
        model, scene, poses = build_scene_and_model(dev)
    trainer = MaskTrainStep(model, lr=1e-2, fp16=True, patch_size=8, label_regularization_weight=0.1, dt_gamma=DT_GAMMA, max_steps=MAX_STEPS,
                            T_thresh=T_THRESH, data_parallel=world > 1, cuda_graph=(world == 1 and not os.environ.get("INERF_NO_GRAPH")))
    intr = synthetic.intrinsics(H_IMG, W_IMG)
    batches = []
    for i in range(8):
        g = torch.Generator().manual_seed(100 + i * world + rank)
        r = synthetic.get_rays(poses[(i * world + rank) % N_POSES][None], intr, H_IMG, W_IMG, N=n_rays, patch_size=8, generator=g)
        o, d = r["rays_o"].reshape(-1, 3).contiguous(), r["rays_d"].reshape(-1, 3).contiguous()
        labels = torch.from_numpy(scene.first_hit_labels(o.numpy().astype(np.float64), d.numpy().astype(np.float64)))
        batches.append({"rays_o": o[None].to(dev), "rays_d": d[None].to(dev), "masks": labels[None].to(dev)})
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # every distinct batch goes through once before timing: the sample count differs per batch, and the first time a larger
    # sample stream shows up the caching allocator grows (cudaMalloc + sync) -- steady-state training never sees that
    for i in range(max(warmup, 2 * len(batches))):   # two passes: the second one replays a graph whose budget fits every batch
        trainer.step(batches[i % 8])
    barrier()
    model.step_counter.zero_()
    model.local_step = 0
    clocks = ClockSampler(dev.index or 0)
    if rank == 0:
        clocks.start()
    ev, totals = [], []
    for i in range(steps):
        flush_buf.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = trainer.step(batches[(warmup + i) % 8])
        e1.record()
        ev.append((e0, e1))
        totals.append(trainer.last_total)
    barrier()
    times = sorted(a.elapsed_time(b) for a, b in ev)
    total_ms = sum(times)
    clk = clocks.stop() if rank == 0 else None
    n_samples = float(model.step_counter[: min(16, steps), 0].float().mean().item())
    if trainer.cuda_graph and trainer.graph_replays:
        n_samples = float(sum(totals)) / max(1, len(totals))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t[0])
    model.eval()
    for p in model.parameters():
        p.requires_grad_(True)
    return {"ms": total_ms / steps, "ms_median": times[len(times) // 2], "samples": n_samples, "loss": float(loss.item()), "clocks": clk,
            "cuda_graph": bool(trainer.cuda_graph), "graph_captures": trainer.graph_captures, "graph_replays": trainer.graph_replays}


def run_train_arm(args):
    """BASELINE.json configs[2] / [4]: instance-field training step (MaskTrainer.train_step + backward + Adam, nerf/utils.py:
    929-936, 1287-1373), `--rays` rays per GPU per step as 8x8 patches, max_steps 1024, fp16 autocast (the `-O` preset).
    At N > 1: data-parallel, one all_reduce of the mask-table + mask-net gradients per step."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_rays = args.rays
    r = measure_train(dev, rank, world, args.steps, args.warmup, n_rays)
    if rank == 0:
        ms = r["ms"]
        line = {"metric": "instance_field_train_step_ms", "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f16 autocast / f32 params",
                "data": "synthetic",
                "config": {"workload": f"c3: instance-field training step, {n_rays} rays/GPU (8x8 patches) x max_steps 1024, K=32, hash+MLP backward, Adam",
                           "rays_per_gpu": n_rays, "samples_per_step_per_gpu": r["samples"], "l2": "flushed between timed steps (512 MB write)",
                           "parallelism": f"dp{world}, flat-bucket all_reduce" if world > 1 else "single GPU"},
                "ms_per_step_median": r["ms_median"], "rays_per_s": world * n_rays / (ms * 1e-3), "loss": r["loss"], "clocks": r["clocks"]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  NCCL / CUDA libraries print banners ("NCCL version ...") straight to fd 1, so
    fd 1 is pointed at stderr for the life of the process and the JSON line goes out through a private duplicate."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


_JSON_OUT = None


def emit(line: dict):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the train-step figure appended to the render line at N = 1")
    ap.add_argument("--workload", default="render", choices=["render", "train"], help="render = the headline metric (default); train = train-step ms")
    ap.add_argument("--rays", type=int, default=4096, help="train workload: rays per GPU per step (4096 = config c3, 65536 = c5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "train":
        run_train_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
