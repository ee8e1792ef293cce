"""N > 1 host logic on CPU: world_size-2 gloo process group (127.0.0.1), no GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rlinf_b200 import dist_utils as D

        # env sharding on group boundaries
        start, per = D.shard_envs(4096, world, rank, group_size=8)
        assert (start, per) == (rank * 2048, 2048)
        # per-rank batch arithmetic of run_training (embodied_fsdp_actor_worker.py:523-546)
        per_rank, accum, n_global = D.per_rank_batch(262144, world, 65536, 2048 * 512)
        assert (per_rank, accum, n_global) == (131072, 2, 8)
        # per-rank shuffle permutation: seed + rank, CPU mt19937 (bit-exact with the reference)
        g = torch.Generator()
        g.manual_seed(D.shuffle_seed(1234, rank))
        perm = torch.randperm(1000, generator=g)
        gathered = [torch.empty_like(perm) for _ in range(world)]
        dist.all_gather(gathered, perm)
        assert not torch.equal(gathered[0], gathered[1])
        g2 = torch.Generator()
        g2.manual_seed(1234 + 1)
        assert torch.equal(gathered[1], torch.randperm(1000, generator=g2))
        # gradient all-reduce: SUM in place, 1/world folded into the optimiser's grad_scale
        grads = torch.full((1000,), float(rank + 1))
        scale = D.allreduce_flat_grads(grads, world)
        assert scale == 0.5 and torch.equal(grads, torch.full((1000,), 3.0))
        assert torch.allclose(grads * scale, torch.full((1000,), 1.5))  # = mean over ranks (DDP/FSDP)
        # metrics: AVG of means, SUM of explained-variance statistics
        mean_vec = torch.arange(24, dtype=torch.float32) + rank
        ev = torch.tensor([10.0, 1.0, 2.0, 3.0, 4.0]) * (rank + 1)
        out = D.reduce_metric_pack(mean_vec, ev, torch.tensor(2.0 + rank), world)
        assert torch.allclose(out[:24], torch.arange(24, dtype=torch.float32) + 0.5)
        assert torch.allclose(out[24:29], torch.tensor([30.0, 3.0, 6.0, 9.0, 12.0]))
        assert abs(out[29].item() - 2.5) < 1e-6
        # parameter broadcast from rank 0
        params = torch.full((77,), float(rank))
        D.broadcast_params(params, world, src=0)
        assert torch.equal(params, torch.zeros(77))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_shard_errors():
    from rlinf_b200 import dist_utils as D

    with pytest.raises(ValueError):
        D.shard_envs(100, 3, 0)
    with pytest.raises(ValueError):
        D.shard_envs(64, 2, 0, group_size=5)
    with pytest.raises(AssertionError):
        D.per_rank_batch(1000, 2, 300, 4000)


def test_bench_sharding_arithmetic_for_1_2_4_8_ranks():
    """The headline config (B=4096, T=512, 8 mini-batches) splits evenly for every N the scaling run uses: per-rank
    envs, per-rank share of the global batch, micro-batch count (the N=2 run once failed on exactly this)."""
    from rlinf_b200.config import synthetic_ppo_config
    from rlinf_b200.dist_utils import per_rank_batch, shard_envs

    B, T = 4096, 512
    for world in (1, 2, 4, 8):
        cfg = synthetic_ppo_config(B=B, T=T, world_size=world)
        covered = []
        for rank in range(world):
            start, n = shard_envs(cfg.env.train.total_num_envs, world, rank, cfg.algorithm.get("group_size", 1))
            covered.append((start, n))
            per_rank, accum, n_global = per_rank_batch(cfg.actor.global_batch_size, world, cfg.actor.micro_batch_size,
                                                       n * T)
            assert per_rank == cfg.actor.global_batch_size // world and accum == 1
            assert (n * T) % per_rank == 0 and (n * T) // per_rank == 8  # 8 optimiser steps per epoch on every rank
        assert covered == [(r * B // world, B // world) for r in range(world)]
