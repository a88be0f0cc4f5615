"""Data-parallel instance-field training on 2 GPUs over NCCL (BASELINE.json configs[4]): gradients are views
of one flat buffer, ONE in-place all_reduce(SUM) inside the step (captured in the CUDA graph), 1/world folded into FusedAdam.
Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py -m gpu`); skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    from datetime import timedelta
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=timedelta(seconds=90))
    try:
        import bench
        from instance_nerf_b200 import parallel
        from instance_nerf_b200.nerf.trainer import MaskTrainStep
        kw = dict(lr=1e-2, fp16=True, patch_size=8, label_regularization_weight=0.1, dt_gamma=bench.DT_GAMMA, max_steps=bench.MAX_STEPS,
                  T_thresh=bench.T_THRESH, data_parallel=True)
        losses = {}
        finals = {}
        for mode in ("eager", "graph"):
            model, scene, poses = bench.build_scene_and_model(dev, K=16)
            parallel.broadcast_parameters(model, 0)
            batches = bench.train_batches(dev, scene, poses, 2048, rank, world, n=3)
            tr = MaskTrainStep(model, cuda_graph=(mode == "graph"), **kw)
            assert tr.bucket.attached() and tr.optimizer.grad_div == float(world)
            ls = [float(tr.step(batches[i % 3])) for i in range(12)]
            assert np.isfinite(ls).all()
            if mode == "graph":
                assert tr.graph_captures == 1 and tr.graph_replays == 12 - MaskTrainStep.GRAPH_WARMUP_STEPS
            losses[mode] = ls
            # every rank holds the same parameters after the same all-reduced updates: bit for bit
            flat = torch.cat([p.detach().reshape(-1) for p in (model.encoder_mask.embeddings, *[l.weight for l in model.mask_net])])
            ck = torch.stack([flat.double().sum(), flat.double().abs().sum(), (flat.double() * torch.arange(flat.numel(), device=dev).double()).sum()])
            both = [torch.empty_like(ck) for _ in range(world)]
            dist.all_gather(both, ck)
            assert torch.equal(both[0], both[1]), (mode, both)
            finals[mode] = flat
            if mode == "graph":
                # overflow on ONE rank only: its guard poisons the summed gradients, every rank skips the step, every rank redoes
                # the batch eagerly and re-captures -- the collective sequence stays aligned
                total = tr._samples_seen
                tr._graph = None
                if rank == 1:
                    tr._samples_seen = total // 4
                caps = tr.graph_captures
                l1 = float(tr.step(batches[0]))
                assert tr.graph_captures == caps + 1 and tr._graph is None and np.isfinite(l1)
                l2 = float(tr.step(batches[0]))      # same batch: the re-captured budget (1.125 x its total) fits
                assert tr.graph_captures == caps + 2 and tr._graph is not None and np.isfinite(l2)
        np.testing.assert_allclose(losses["graph"], losses["eager"], rtol=5e-3, atol=5e-4)
        assert np.mean(losses["eager"][-3:]) < np.mean(losses["eager"][:3])
        w_e, w_g = finals["eager"][-(64 * 64):], finals["graph"][-(64 * 64):]
        torch.testing.assert_close(w_g, w_e, rtol=5e-2, atol=5e-4)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(420)
def test_dp_training_two_gpus(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert all(ret.get(r) for r in range(2))
