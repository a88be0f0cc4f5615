"""World-size-2 gloo tests of the multi-GPU host logic (instance_nerf_b200/parallel.py): ray sharding + tile gather for
rendering, flat-bucket gradient all-reduce for training.  The render function is a deterministic stand-in (the kernels need
a GPU); what is tested is that sharded == unsharded."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from instance_nerf_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_render(o, d, K=5):
    """Per-ray deterministic function of the ray: stands in for model.render."""
    o, d = o[0], d[0]
    base = (o * 0.3 + d * 0.7)
    image = torch.sigmoid(base)
    depth = base.sum(-1).abs()
    logits = torch.stack([(base * (k + 1)).sum(-1) for k in range(K)], -1)
    return {"image": image[None], "depth": depth[None], "instance_mask_logits": logits[None]}


def _worker(rank, world, port, H, W, n_frames, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        o, d = torch.randn(H * W, 3, generator=g), torch.randn(H * W, 3, generator=g)
        # latency mode: one frame over both ranks, ragged shards (H = 21 rows, blocks of 4)
        full = parallel.render_frame_sharded(_fake_render, o, d, H, W, block_rows=4, dst=None)
        ref = _fake_render(o[None], d[None])
        assert torch.allclose(full["image"], ref["image"][0]) and torch.allclose(full["depth"], ref["depth"][0])
        assert torch.allclose(full["instance_mask_logits"], ref["instance_mask_logits"][0])
        only0 = parallel.render_frame_sharded(_fake_render, o, d, H, W, block_rows=4, dst=0)
        assert (only0 is not None) == (rank == 0)
        # throughput mode: frames round-robin, odd frame count
        frames = [torch.randn(H * W, 3, generator=torch.Generator().manual_seed(10 + f)) for f in range(n_frames)]
        outs = parallel.render_frames_sharded(lambda f: _fake_render(frames[f][None], d[None]), n_frames, dst=None)
        assert sorted(outs) == list(range(n_frames))
        for f in range(n_frames):
            assert torch.allclose(outs[f]["image"], _fake_render(frames[f][None], d[None])["image"][0])
        # training: per-rank gradients averaged through one flat bucket; a parameter without grad counts as zeros
        torch.manual_seed(0)
        table = torch.nn.Parameter(torch.zeros(1000, 2))
        w = torch.nn.Parameter(torch.zeros(7, 3))
        frozen = torch.nn.Parameter(torch.zeros(4), requires_grad=False)
        table.grad = torch.full_like(table, float(rank + 1))
        bucket = parallel.FlatGradBucket([table, w, frozen], extra=1).attach()
        # gradients are VIEWS of the flat buffer: what was there is kept, a parameter without a gradient reads as zeros,
        # in-place accumulation (what autograd and the backward kernel do) lands in the buffer
        assert bucket.attached() and table.grad.data_ptr() == bucket.flat.data_ptr()
        assert w.grad.data_ptr() == bucket.flat.data_ptr() + 4 * 2000 and float(w.grad.abs().sum()) == 0.0
        if rank == 0:
            w.grad += 4
        bucket.extra[0] = float(rank == 1)            # a flag riding in the same collective
        _, nbytes = bucket.all_reduce()
        assert nbytes == (2000 + 24 + 4) * 4          # slices padded to 16 bytes
        assert torch.allclose(table.grad, torch.full_like(table, 3.0)) and torch.allclose(w.grad, torch.full_like(w, 4.0))
        assert float(bucket.extra[0]) == 1.0
        table.grad.zero_(); w.grad.zero_()
        table.grad += float(rank + 1)
        assert bucket.sync() == nbytes and torch.allclose(table.grad, torch.full_like(table, 1.5))
        # async tile gather to one rank, double buffered
        tg = parallel.TileGather(6, 3, torch.device("cpu"), dst=0)
        for step in range(3):
            got = tg.submit(torch.full((6, 3), float(10 * step + rank)))
            tg.drain()
            if rank == 0:
                assert [float(t[0, 0]) for t in got] == [10.0 * step, 10.0 * step + 1]
            else:
                assert got is None
        assert tg.bytes_sent_per_step == (0 if rank == 0 else 6 * 3 * 4)
        # parameter broadcast
        m = torch.nn.Linear(3, 2)
        with torch.no_grad():
            m.weight.fill_(float(rank))
        parallel.broadcast_parameters(m, 0)
        assert float(m.weight.abs().sum()) == 0.0
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_sharded_render_and_grad_bucket_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 21, 16, 5, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_sharding_helpers_cover_everything_once():
    for world in (1, 2, 3, 8):
        seen = torch.cat([parallel.shard_rows(30, 7, r, world, 4) for r in range(world)])
        assert sorted(seen.tolist()) == list(range(30 * 7))
        fr = sum((parallel.shard_frames(11, r, world) for r in range(world)), [])
        assert sorted(fr) == list(range(11))
        n = sum(len(range(*parallel.shard_batch(4096, r, world).indices(4096))) for r in range(world))
        assert n == 4096
    # single process: collectives are no-ops
    t = torch.arange(12.0).reshape(3, 4)
    assert parallel.gather_tiles(t)[0] is t


def test_balance_frames_equal_counts_and_near_equal_totals():
    """parallel.balance_frames: every frame exactly once, counts differ by at most one, totals within a few percent where plain
    round-robin over a drifting camera path is off by much more; deterministic under ties."""
    import numpy as np
    from instance_nerf_b200.parallel import balance_frames, shard_frames
    rng = np.random.default_rng(0)
    n, world = 200, 8
    costs = (100 + 25 * np.sin(np.arange(n) / 7.0) + rng.normal(0, 8, n)).clip(min=1).tolist()
    plan = balance_frames(costs, world)
    assert sorted(f for p in plan for f in p) == list(range(n))
    assert max(len(p) for p in plan) - min(len(p) for p in plan) <= 1
    assert all(p == sorted(p) for p in plan)
    tot = [sum(costs[f] for f in p) for p in plan]
    rr = [sum(costs[f] for f in shard_frames(n, r, world)) for r in range(world)]
    assert max(tot) / (sum(tot) / world) < 1.01 < max(rr) / (sum(rr) / world) + 0.02
    assert max(tot) - min(tot) <= max(rr) - min(rr)
    assert balance_frames([1.0] * 10, 4) == balance_frames([1.0] * 10, 4) == [[0, 7, 8], [1, 6, 9], [2, 5], [3, 4]]
    assert balance_frames([], 3) == [[], [], []] and balance_frames([5.0, 1.0], 1) == [[0, 1]]
