"""2-rank NCCL parity worker (launched by tests/test_gpu_parity2.py::test_two_rank_nccl_update_matches_oracle with torchrun).

Each rank owns half of the environments, rolls out on its GPU, and runs the data-parallel update (seed+rank shuffles,
ONE NCCL all-reduce of the flat gradient buffer per optimiser step, metrics all-reduce).  Rank 0 gathers both rollout
batches and replays the update with the CPU oracle's data-parallel emulation (RunnerOracle.update_dp: per-rank shuffles,
gradients averaged over ranks - embodied_fsdp_actor_worker.py:511-523, FSDP/DDP gradient averaging) and compares."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle.runner_oracle import RunnerOracle
    from rlinf_b200.config import synthetic_ppo_config
    from rlinf_b200.runner import EmbodiedRunner

    graph = os.environ.get("RB200_DIST_GRAPH", "0") == "1"
    B, T, obs, act = 64, 16, 8, 2
    n = B * T
    cfg = synthetic_ppo_config(B=B, T=T, obs_dim=obs, action_dim=act, update_epoch=2, num_minibatches=2,
                               micro_batch_size=n // 2 // world // 2, world_size=1,
                               **{"actor.cuda_graph_update": graph, "actor.cuda_graph_multi_rank": graph})
    run = EmbodiedRunner(cfg)
    assert run.B == B // world
    out = {"rank": rank, "ok": True, "iters": []}
    p0 = {k: p.detach().cpu().clone() for k, p in run.actor.model.named_parameters()}
    flat0 = [torch.empty_like(run.actor.model.flat_params) for _ in range(world)]
    dist.all_gather(flat0, run.actor.model.flat_params)
    assert all(torch.equal(f, flat0[0]) for f in flat0), "ranks must start from identical parameters"
    orc = RunnerOracle(cfg, params=p0) if rank == 0 else None
    for it in range(3 if graph else 2):
        run.update_rollout_weights()
        run.rollout_phase()
        torch.cuda.synchronize()
        mine = {k: ({kk: vv.cpu() for kk, vv in v.items()} if isinstance(v, dict) else v.cpu())
                for k, v in run.buffer.as_batch().items()}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        m = run.update_phase()
        torch.cuda.synchronize()
        flats = [torch.empty_like(run.actor.model.flat_params) for _ in range(world)]
        dist.all_gather(flats, run.actor.model.flat_params)
        same = all(torch.equal(f, flats[0]) for f in flats)
        rec = {"replicas_identical": bool(same)}
        if rank == 0:
            om = orc.update_dp(gathered)
            worst = 0.0
            for name, p in run.actor.model.named_parameters():
                ref = orc.params[name].detach()
                worst = max(worst, float((p.cpu() - ref).abs().max()))
                if not torch.allclose(p.cpu(), ref, rtol=1e-4, atol=2e-5):
                    out["ok"] = False
                    rec.setdefault("bad_params", []).append(name)
            rec["max_param_err"] = worst
            for k, v in om.items():
                if k == "critic/value_clip_ratio" or k not in m:
                    continue
                if not np.isclose(m[k], v, rtol=2e-4, atol=1e-6, equal_nan=True):
                    out["ok"] = False
                    rec.setdefault("bad_metrics", []).append([k, m[k], v])
        if not same:
            out["ok"] = False
        out["iters"].append(rec)
    if rank == 0:
        with open(os.environ["RB200_DIST_OUT"], "w") as fh:
            json.dump(out, fh)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
