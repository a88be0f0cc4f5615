"""TEST / MEASUREMENT INFRASTRUCTURE (not the product, not bench.py's value).

Per-op roofline table (SURVEY.md section 8d "algorithmic bytes per unit"): every op-level kernel of libinerf_b200 timed
with CUDA events on a c5-sized training stream (65 536 rays of the synthetic room, K = 32), next to the reference's own
kernel for the same op (oracle/_ref, built unmodified from the reference sources for sm_100a) on the same inputs.
achieved GB/s = algorithmic bytes / time; frac = achieved / measured HBM peak (MEASURED_PEAKS.json, else the
B200_PROFILING.md fallback).  Writes gpurun_out/op_bench.json (copied to profiles/ by hand).

    python tests/dev_op_bench.py [--rays 65536]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
from helpers import make_rays  # noqa: E402
from instance_nerf_b200 import raymarching as rm  # noqa: E402
from instance_nerf_b200._lib import call, ptr, stream_ptr  # noqa: E402
from oracle import ref_loader  # noqa: E402


ITERS, WARM = [10], [3]     # --iters / --warm (1 / 1 under ncu: two launches per kernel are enough there)


def timeit(fn, iters=None, warm=None):
    iters, warm = iters or ITERS[0], WARM[0] if warm is None else warm
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gbs_burst", "hbm_copy_gbs"):
                if k in j:
                    return float(j[k]), "MEASURED_PEAKS.json:" + k
        except Exception:
            pass
    return 6550.0, "B200_PROFILING.md fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warm", type=int, default=3)
    args = ap.parse_args()
    ITERS[0], WARM[0] = args.iters, args.warm
    ref = ref_loader.load()
    assert ref is not None, "oracle/_ref is not built"
    dev = torch.device("cuda:0")
    st = stream_ptr(dev)
    K, bound = bench.K_INST, bench.BOUND
    model, scene, poses = bench.build_scene_and_model(dev)
    side = int(round(args.rays ** 0.5))
    o, d = make_rays(scene, side, args.rays // side)
    o, d = o.to(dev), d.to(dev)
    N = o.shape[0]
    bits, C, H = model.density_bitfield, model.cascade, model.grid_size
    nears, fars = rm.near_far_from_aabb(o, d, model.aabb_train, model.min_near)
    noises = torch.rand(N, generator=torch.Generator().manual_seed(2)).to(dev)
    peak, peak_src = hbm_peak()
    rows = []

    def row(op, unit, units, bytes_per_unit, ms_ours, ms_ref, extra_bytes=0, note=""):
        b = units * bytes_per_unit + extra_bytes
        r = dict(op=op, unit=unit, units=int(units), algorithmic_bytes=int(b), ours_ms=ms_ours, ours_gbs=b / ms_ours / 1e6,
                 ours_frac_of_hbm_peak=b / ms_ours / 1e6 / peak, ref_ms=ms_ref,
                 ref_gbs=(b / ms_ref / 1e6) if ms_ref else None, speedup_vs_ref_kernel=(ms_ref / ms_ours) if ms_ref else None, note=note)
        rows.append(r)
        print(json.dumps(r), flush=True)

    # ---------------------------------------------------------------- march_rays_train ----------------------------
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, bound, bits, C, H, nears, fars, counter, -1, True, -1, True, bench.DT_GAMMA,
                                                   bench.MAX_STEPS, noises=noises)
    M = int(counter[0].item())
    print(f"# {N} rays, {M} samples ({M / N:.1f}/ray), K={K}", flush=True)
    rays_w = torch.empty_like(rays)

    from instance_nerf_b200._lib import lib
    t_scr = torch.empty(int(lib().inerf_march_scratch_floats(N, bench.MAX_STEPS)), dtype=torch.float32, device=dev)

    def ours_march():   # what raymarching.march_rays_train issues: count + record t, scan, expand
        counter.zero_()
        call("inerf_march_rays_train_count_t", ptr(o), ptr(d), ptr(bits), bound, bench.DT_GAMMA, bench.MAX_STEPS, N, C, H, ptr(nears),
             ptr(fars), ptr(rays_w), ptr(counter), ptr(noises), ptr(t_scr), st)
        call("inerf_march_rays_train_expand", ptr(o), ptr(d), bound, bench.DT_GAMMA, bench.MAX_STEPS, N, C, H, M, ptr(nears), ptr(xyzs),
             ptr(dirs), ptr(deltas), ptr(rays_w), ptr(noises), ptr(t_scr), st)

    rx = torch.empty(M, 3, device=dev); rd = torch.empty(M, 3, device=dev); rl = torch.empty(M, 2, device=dev)
    rrays = torch.empty(N, 3, dtype=torch.int32, device=dev); rcounter = torch.zeros(2, dtype=torch.int32, device=dev)

    def ref_march():   # raymarching.py:205-208 zero-fills the sample buffers on every call (here at the exact M, not N*max_steps)
        rx.zero_(); rd.zero_(); rl.zero_(); rcounter.zero_()
        ref.raymarching.march_rays_train(o, d, bits, bound, bench.DT_GAMMA, bench.MAX_STEPS, N, C, H, M, nears, fars, rx, rd, rl, rrays, rcounter, noises)

    row("march_rays_train (count+scan+expand)", "sample", M, 32, timeit(ours_march), timeit(ref_march), extra_bytes=N * 48,
        note="ref zero-fills exactly M rows here; its wrapper zero-fills N*max_steps rows")

    # ---------------------------------------------------------------- grid encode fwd / bwd -----------------------
    x01 = ((xyzs + bound) / (2 * bound)).contiguous()
    enc = model.encoder_mask
    S = float(np.log2(enc.per_level_scale))
    for dtype, code, s in ((torch.float16, 1, 2), (torch.float32, 0, 4)):
        table = enc.embeddings.detach().to(dtype).contiguous()
        out1 = torch.empty(M, 32, device=dev, dtype=dtype)
        out0 = torch.empty(16, M, 2, device=dev, dtype=dtype)
        t1 = timeit(lambda: call("inerf_grid_encode_forward", ptr(x01), ptr(table), ptr(enc.offsets), ptr(out1), M, 3, 2, 16, S, 16, None, 0, 0, 0, code, 1, st))

        def ref_fwd():   # grid.py:43-44 + :57: per-call table cast (autocast) and the [L,B,C] -> [B,L*C] permute copy
            tb = enc.embeddings.to(dtype) if dtype == torch.float16 else table
            ref.gridencoder.grid_encode_forward(x01, tb, enc.offsets, out0, M, 3, 2, 16, S, 16, None, 0, False, 0)
            return out0.permute(1, 0, 2).reshape(M, 32)
        t0 = timeit(ref_fwd)
        row(f"grid_encode_forward {str(dtype)[6:]}", "sample", M, 12 + 256 * s + 32 * s, t1, t0,
            note="ref includes its per-call table cast (fp16) and permute copy")
        grad = torch.randn(M, 32, device=dev).to(dtype)
        g1 = torch.zeros_like(table)
        t1 = timeit(lambda: call("inerf_grid_encode_backward", ptr(grad), ptr(x01), None, ptr(enc.offsets), ptr(g1), M, 3, 2, 16, S, 16, None, None,
                                 0, 0, 0, code, 1, st))
        g0 = torch.zeros_like(table)

        def ref_bwd():   # grid.py:75: permute + contiguous copy of the incoming gradient
            gp = grad.view(M, 16, 2).permute(1, 0, 2).contiguous()
            ref.gridencoder.grid_encode_backward(gp, x01, table, enc.offsets, g0, M, 3, 2, 16, S, 16, None, None, 0, False, 0)
        t0 = timeit(ref_bwd)
        row(f"grid_encode_backward {str(dtype)[6:]}", "sample", M, 12 + 32 * s + 512 * s, t1, t0, note="atomic RMW counted as read+write")
        del out0, out1, grad, g0, g1

    # ---------------------------------------------------------------- SH ------------------------------------------
    sh1 = torch.empty(M, 16, device=dev); sh0 = torch.empty(M, 16, device=dev)
    t1 = timeit(lambda: call("inerf_sh_encode_forward", ptr(dirs), ptr(sh1), M, 3, 4, None, st))
    t0 = timeit(lambda: ref.shencoder.sh_encode_forward(dirs, sh0, M, 3, 4, None))
    row("sh_encode_forward deg4", "sample", M, 12 + 64, t1, t0)
    del sh0, sh1

    # ---------------------------------------------------------------- fused field forward -------------------------
    sig = torch.empty(M, device=dev); rgb = torch.empty(M, 3, device=dev); msk = torch.empty(M, K, device=dev)
    with torch.no_grad():
        t1 = timeit(lambda: model.forward_fused(xyzs, dirs))
    row("field_forward fused (2x hash encode + SH + 3 MLPs, tcgen05)", "sample", M, 1024 + 24 + (4 + K) * 4, t1, None,
        note="no single reference kernel: the reference runs 2 encodes + SH + 8 cuBLAS GEMMs + elementwise; see ref_gpu_path.json")
    del sig, rgb, msk

    # ---------------------------------------------------------------- fused field backward (instance head) --------
    import ctypes
    desc = model._field_desc()
    _, wb = model._packed_weights(want_bwd=True)
    x0 = torch.empty(M, 48, dtype=torch.float16, device=dev)
    sig = torch.empty(M, device=dev); rgb = torch.empty(M, 3, device=dev); msk = torch.empty(M, K, device=dev)
    call("inerf_field_forward_train", ctypes.byref(desc), ptr(xyzs), ptr(dirs), M, ptr(sig), ptr(rgb), ptr(msk), ptr(x0), st)
    gl = (torch.randn(M, K, device=dev) * 1e-3).contiguous()
    gt = torch.zeros_like(model.encoder_mask.embeddings)
    gw0 = torch.zeros(64, 47, device=dev); gw1 = torch.zeros(64, 64, device=dev); gw2 = torch.zeros(K, 64, device=dev)
    t1 = timeit(lambda: call("inerf_field_backward_mask", ctypes.byref(desc), ptr(wb), ptr(xyzs), ptr(x0), ptr(gl), M, ptr(gt), ptr(gw0), ptr(gw1),
                             ptr(gw2), st))
    row("field_backward_mask fused (mask-net dX/dW on tcgen05 + fp32 table scatter)", "sample", M, 12 + 96 + 4 * K + 2048, t1, None,
        note="table-gradient atomics counted as read+write of 16 levels x 8 corners x 8 B; the reference runs 6 GEMMs + grid backward through autograd")
    del x0, sig, rgb, msk, gl, gt

    # ---------------------------------------------------------------- composite train fwd / bwd -------------------
    g = torch.Generator().manual_seed(5)
    sigmas = torch.exp(torch.randn(M, generator=g) * 1.5 - 1.0).to(dev)
    rgbs = torch.rand(M, 3, generator=g).to(dev)
    masks = torch.randn(M, K, device=dev)
    ws = torch.empty(N, device=dev); dp = torch.empty(N, device=dev); im = torch.empty(N, 3, device=dev); mo = torch.empty(N, K, device=dev)
    T = bench.T_THRESH
    t1 = timeit(lambda: call("inerf_composite_rays_with_masks_train_forward", ptr(sigmas), ptr(rgbs), ptr(masks), ptr(deltas), ptr(rays), M, N, K, T,
                             ptr(ws), ptr(dp), ptr(im), ptr(mo), st))
    ws0 = torch.empty(N, device=dev); dp0 = torch.empty(N, device=dev); im0 = torch.empty(N, 3, device=dev); mo0 = torch.empty(N, K, device=dev)
    t0 = timeit(lambda: ref.raymarching.composite_rays_with_masks_train_forward(sigmas, rgbs, masks, deltas, rays, M, N, K, T, ws0, dp0, im0, mo0))
    row(f"composite_rays_with_masks_train_forward K={K}", "sample", M, 24 + 4 * K, t1, t0, extra_bytes=N * (12 + 20 + 4 * K),
        note="all samples counted; samples after a ray's early stop are not read")
    gws = torch.randn(N, device=dev); gim = torch.randn(N, 3, device=dev); gmo = torch.randn(N, K, device=dev)
    gs = torch.zeros(M, device=dev); gr = torch.zeros(M, 3, device=dev); gm = torch.zeros(M, K, device=dev)

    def ours_cbwd():
        gs.zero_(); gr.zero_(); gm.zero_()     # raymarching.py:323-326: grads beyond the early stop stay zero
        call("inerf_composite_rays_with_masks_train_backward", ptr(gws), ptr(gim), ptr(gmo), ptr(sigmas), ptr(rgbs), ptr(masks), ptr(deltas),
             ptr(rays), ptr(ws), ptr(im), ptr(mo), M, N, K, T, ptr(gs), ptr(gr), None, ptr(gm), st)
    acc0 = torch.zeros(N, K, device=dev)

    def ref_cbwd():
        gs.zero_(); gr.zero_(); gm.zero_(); acc0.zero_()
        ref.raymarching.composite_rays_with_masks_train_backward(gws, gim, gmo, sigmas, rgbs, masks, deltas, rays, ws0, im0, mo0, M, N, K, T,
                                                                 gs, gr, acc0, gm)
    row(f"composite_rays_with_masks_train_backward K={K}", "sample", M, 24 + 4 * K + 16 + 4 * K, timeit(ours_cbwd), timeit(ref_cbwd),
        extra_bytes=N * 2 * (20 + 4 * K), note="both timings include zero-filling the three gradient buffers")
    del masks, gm, gs, gr

    # ---------------------------------------------------------------- occupancy EMA + pack ------------------------
    cells = model.density_grid.numel()
    tmp = (torch.rand(model.density_grid.shape, device=dev) * 40 - 8)
    t1 = timeit(lambda: model.ema_update_(tmp, 0.95))

    def ref_ema():   # mask_renderer.py:532-540 in torch + the reference packbits kernel
        grid = model.density_grid
        valid = (grid >= 0) & (tmp >= 0)
        grid[valid] = torch.maximum(grid[valid] * 0.95, tmp[valid])
        mean = torch.mean(grid.clamp(min=0)).item()
        ref.raymarching.packbits(grid, model.density_bitfield.shape[0], min(mean, 10.0), model.density_bitfield)
    t0 = timeit(ref_ema)
    row("occupancy EMA + mean + packbits", "cell", cells, 16.125, t1, t0, note="ref = the torch boolean-index sequence + reference packbits")

    out = dict(gpu=torch.cuda.get_device_name(0), rays=N, samples=M, K=K, hbm_peak_gbs=peak, peak_source=peak_src, rows=rows)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "op_bench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
