"""GPU parity of the SURVEY 8(f) rows: remaining ADV_REGISTRY / LOSS_REGISTRY entries (reference-generated goldens
`golden_next.npz`, `golden_r2.npz`), the fp64 masked normalisations (a24), and the Trajectory views of the device rollout
buffer against the reference's EmbodiedTrajectoryBuilder (a22 / f2)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_registries_hold_every_reference_entry():
    import rlinf_b200.algorithms as A

    assert sorted(A.ADV_REGISTRY) == ["gae", "grpo", "grpo_dynamic", "grpo_video", "opd", "raw", "reinpp"]
    assert sorted(A.LOSS_REGISTRY) == ["actor", "actor_critic", "decoupled_actor_critic", "opd"]
    assert sorted(A.LOSS_SCALE_REGISTRY) == ["agent_level", "group_level", "turn_level"]


def test_raw_and_reinpp_advantages_golden(golden):
    import rlinf_b200.algorithms as A

    g = golden("next")
    rewards, mask = _t(g["next_rewards"]), _t(g["next_mask"])
    G = int(g["next_group_size"][0])
    for name, kw in (("raw", dict(adv_type="raw", normalize_advantages=False)),
                     ("raw_norm", dict(adv_type="raw", normalize_advantages=True)),
                     ("reinpp", dict(adv_type="reinpp")),
                     ("reinpp_kl", dict(adv_type="reinpp", kl_beta=0.05, logprob=_t(g["next_lp"]),
                                        ref_logprob=_t(g["next_rlp"]), kl_penalty_type="k3"))):
        adv, ret = A.calculate_adv_and_returns(task_type="reasoning", rewards=rewards.clone(), loss_mask=mask,
                                               group_size=G, **kw)
        assert ret is None and adv.is_cuda and adv.shape == mask.shape
        torch.testing.assert_close(adv.cpu(), _t(g["next_adv_" + name]), rtol=RTOL, atol=1e-6, msg=name)
    with pytest.raises(IndexError):
        A.calculate_adv_and_returns(task_type="reasoning", adv_type="reinpp", rewards=rewards.clone(), loss_mask=mask,
                                    group_size=G, use_reinpp_baseline=True)


def test_grpo_video_and_dynamic_golden(golden):
    import rlinf_b200.algorithms as A

    g = golden("next")
    vr, vm = _t(g["vid_rewards"]), _t(g["vid_mask"])
    for mode in ("frame", "video"):
        adv, ret = A.get_adv_and_returns("grpo_video")(rewards=vr.clone(), loss_mask=vm, group_size=4, advantage_mode=mode)
        assert ret is None
        torch.testing.assert_close(adv.cpu(), _t(g["vid_adv_" + mode]), rtol=RTOL, atol=1e-6, msg=mode)
    with pytest.raises(ValueError):
        A.get_adv_and_returns("grpo_video")(rewards=vr, loss_mask=vm, group_size=4, advantage_mode="nope")
    g2 = golden("r2")
    G, Q, L, n = (int(v) for v in g2["dyn_cfg"])
    idx = [int(v) for v in g2["dyn_idx_to_traj"]]
    for mode in ("trajectory", "turn"):
        adv, ret = A.calculate_adv_and_returns(task_type="reasoning", adv_type="grpo_dynamic",
                                               rewards=_t(g2["dyn_rewards"]).clone(), loss_mask=_t(g2["dyn_mask"]),
                                               group_size=G, idx_to_traj=idx, num_sequence=n, advantage_mode=mode)
        assert ret is None
        torch.testing.assert_close(adv.cpu(), _t(g2["dyn_adv_" + mode]), rtol=RTOL, atol=1e-6, msg=mode)


def test_opd_advantages_and_loss_golden(golden):
    import rlinf_b200.algorithms as A

    g = golden("next")
    student, teacher, lmask = _t(g["opd_student"]), _t(g["opd_teacher"]), _t(g["opd_mask"])
    res = A.calculate_adv_and_returns(task_type="embodied", adv_type="opd", prev_logprobs=student,
                                      teacher_logprobs=teacher, loss_mask=lmask, normalize_advantages=False,
                                      num_action_chunks=2)
    assert list(res) == ["advantages"]
    assert torch.equal(res["advantages"].cpu(), _t(g["opd_adv"]))  # one fp32 subtraction per element
    n = lmask.shape[0] * lmask.shape[1]
    adv = res["advantages"].reshape(n, 2, -1)
    for name, mes in (("mean", None), ("ratio", 50)):
        lp = _t(g["opd_lp"]).cuda().requires_grad_(True)
        loss, metrics = A.get_policy_loss("opd")(logprobs=lp, advantages=adv, loss_mask=lmask.reshape(n, 2),
                                                 loss_mask_sum=_t(g["opd_msum"]), max_episode_steps=mes)
        loss.backward()
        torch.testing.assert_close(loss.detach().cpu(), _t(g[f"opd_{name}_loss"]), rtol=RTOL, atol=1e-7)
        torch.testing.assert_close(lp.grad.cpu(), _t(g[f"opd_{name}_dlp"]), rtol=RTOL, atol=1e-10)
        keys = [str(k) for k in g[f"opd_{name}_metric_keys"]]
        assert sorted(metrics) == keys
        np.testing.assert_allclose([float(metrics[k]) for k in keys], g[f"opd_{name}_metric_vals"], rtol=RTOL, atol=1e-7)


def test_decoupled_actor_critic_loss_golden(golden):
    """All four reference-generated cases through registry.policy_loss (task_type embodied -> fused route) and, for the
    same inputs, through the registered callable on preprocess_loss_inputs kwargs (the generic route)."""
    import rlinf_b200.algorithms as A
    from rlinf_b200.algorithms import losses as LS

    g = golden("next")
    for name in (str(c) for c in g["dec_cases"]):
        pre = f"dec_{name}_"
        bsz, C, Ad, use_ratio, dual, has_thr = (int(x) for x in g[pre + "cfg"])
        lpt, rt, prox = (str(x) for x in g[pre + "types"])
        has_mask = (pre + "mask") in g.files
        for route in ("fused", "generic"):
            new = _t(g[pre + "new"]).cuda().requires_grad_(True)
            val = _t(g[pre + "val"]).cuda().requires_grad_(True)
            kw = dict(task_type="embodied", loss_type="decoupled_actor_critic", logprob_type=lpt, reward_type=rt,
                      single_action_dim=Ad, logprobs=new, old_logprobs=_t(g[pre + "old"]), advantages=_t(g[pre + "adv"]),
                      returns=_t(g[pre + "ret"]), values=val, prev_values=_t(g[pre + "prev_v"]), clip_ratio_high=0.28,
                      clip_ratio_low=0.2, value_clip=0.2, huber_delta=1.5,
                      loss_mask=_t(g[pre + "mask"]) if has_mask else None,
                      loss_mask_sum=_t(g[pre + "mask_sum"]) if has_mask else None,
                      max_episode_steps=50 if use_ratio else None, critic_warmup=False,
                      proximal_logprobs=_t(g[pre + "proximal"]) if prox == "given" else None,
                      versions=_t(g[pre + "versions"]) if prox == "versions" else None,
                      current_version=5.0 if prox == "versions" else None,
                      behave_weight_threshold=float(g[pre + "thr"][0]) if has_thr > 0 else None)
            if dual:
                kw["clip_ratio_c"] = 3.0
            if route == "fused":
                loss, metrics = A.policy_loss(**kw)
            else:
                kw = {k: (v.cuda() if isinstance(v, torch.Tensor) and not v.is_cuda else v) for k, v in kw.items()}
                loss, metrics = A.get_policy_loss("decoupled_actor_critic")(**LS.preprocess_loss_inputs(**kw))
                metrics = LS.postprocess_loss_metric(metrics)
            loss.backward()
            tag = f"{name}/{route}"
            torch.testing.assert_close(loss.detach().cpu(), _t(g[pre + "loss"]), rtol=RTOL, atol=1e-7, msg=tag)
            dnew = new.grad.cpu() if new.grad is not None else torch.zeros_like(new).cpu()
            torch.testing.assert_close(dnew, _t(g[pre + "dnew"]), rtol=RTOL, atol=1e-9, msg=tag)
            torch.testing.assert_close(val.grad.cpu(), _t(g[pre + "dval"]), rtol=RTOL, atol=1e-9, msg=tag)
            keys = [str(k) for k in g[pre + "metric_keys"]]
            assert sorted(metrics) == keys, (tag, sorted(metrics), keys)
            for k, want in zip(keys, g[pre + "metric_vals"]):
                if k == "critic/value_clip_ratio":  # rounding-noise metric
                    continue
                np.testing.assert_allclose(float(metrics[k]), want, rtol=RTOL, atol=2e-7, err_msg=f"{tag} {k}")


def test_user_registered_loss_gets_preprocessed_inputs():
    """A callable registered by the USER under a built-in name has no fused marker: policy_loss hands it the kwargs of
    preprocess_loss_inputs (algorithms/utils.py:280-376), as the reference does."""
    import rlinf_b200.algorithms as A

    seen = {}
    old = A.LOSS_REGISTRY["actor"]
    try:
        @A.register_policy_loss("actor")
        def my_loss(**kw):
            seen.update({k: (tuple(v.shape) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()})
            return kw["logprobs"].sum(), {"x": torch.tensor(1.0)}

        n, C, Ad = 6, 2, 3
        lp = torch.zeros(n, C * Ad, device="cuda", requires_grad=True)
        loss, m = A.policy_loss(task_type="embodied", loss_type="actor", logprob_type="action_level",
                                reward_type="action_level", single_action_dim=Ad, logprobs=lp,
                                old_logprobs=torch.zeros(n, C * Ad, device="cuda"),
                                advantages=torch.zeros(n, C, device="cuda"), loss_mask=None, loss_mask_sum=None)
        assert seen["logprobs"] == (n, C) and seen["old_logprobs"] == (n, C) and seen["advantages"] == (n, C)
        assert m == {"x": 1.0}
    finally:
        A.LOSS_REGISTRY["actor"] = old


def test_masked_normalisations_golden(golden):
    from rlinf_b200.utils import masked_normalization, masked_stats, normalize_from_stats

    g = golden("r2")
    x, m = _t(g["mn_x"]).cuda(), _t(g["mn_mask"]).cuda()
    torch.testing.assert_close(masked_normalization(x).cpu(), _t(g["mn_plain"]), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(masked_normalization(x, m).cpu(), _t(g["mn_masked"]), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(masked_normalization(x, m, unbiased=True, eps=1e-6).cpu(), _t(g["mn_masked_unbiased"]),
                               rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(masked_stats(x, m).cpu().numpy(), g["mn_stats"], rtol=1e-12)
    np.testing.assert_allclose(masked_stats(x).cpu().numpy(), g["mn_stats_nomask"], rtol=1e-12)
    torch.testing.assert_close(normalize_from_stats(x, _t(g["mn_stats"])).cpu(), _t(g["mn_from_stats"]), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(normalize_from_stats(x, 2 * _t(g["mn_stats"])).cpu(), _t(g["mn_from_stats_x2"]),
                               rtol=1e-6, atol=1e-6)
    with pytest.raises(NotImplementedError):
        masked_normalization(x, m, dim=0)


def test_loss_scale_stages_on_device_tensors(golden):
    from rlinf_b200.algorithms import get_loss_scales

    g = golden("r2")
    idx = [int(v) for v in g["ls_idx_to_traj"]]
    batch = {"idx_to_traj": idx, "extra:idx_to_sub_traj": _t(g["ls_idx_to_sub"]), "response_mask": _t(g["ls_resp"]).cuda(),
             "advantages": _t(g["ls_adv"]).cuda(), "loss_scales": torch.ones(len(idx), device="cuda")}
    ctx = {"folding_scale": [], "data_parallel_world_size": 2, "actor_global_batch_size": 16}
    for name, fn in zip(("group_level", "agent_level", "turn_level"), get_loss_scales(["group_level", "agent_level", "turn_level"])):
        batch = fn(ctx, batch)
        torch.testing.assert_close(batch["advantages"].cpu(), _t(g[f"ls_after_{name}_adv"]), rtol=1e-6, atol=0)
        torch.testing.assert_close(batch["loss_scales"].cpu(), _t(g[f"ls_after_{name}_scales"]), rtol=1e-6, atol=0)


def test_rollout_buffer_trajectory_views_vs_reference_builder(golden):
    """a22 / f2: rows written into the device buffer in the env-worker's order equal what the reference's
    EmbodiedTrajectoryBuilder stacks (to_trajectory), chunks (to_splited_trajectories) and re-joins
    (convert_trajectories_to_batch) - per rollout epoch; the views share the buffer's storage (zero copy)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_r2 import coded_steps  # pure torch helper, no reference needed
    from rlinf_b200.rollout import RolloutBuffer, convert_trajectories_to_batch

    g = golden("r2")
    T, B, act, obs, E, split = (int(v) for v in g["traj_cfg"])
    steps = coded_steps(T, B, act, obs, E)
    keys = [str(k) for k in g["traj_batch_keys"]]
    for e in range(E):
        buf = RolloutBuffer(T, B, obs, act, 1)
        for t in range(T + 1):
            r = steps[e][t]
            if t < T:
                buf.states[t].copy_(r["states"])
                buf.actions[t].copy_(r["action"])
                buf.prev_logprobs[t].copy_(r["logp"])
                buf.rewards[t].copy_(r["reward"])  # reward of step t sits in row t (appended with the next step's output)
            buf.prev_values[t].copy_(r["value"])
            buf.dones[t].copy_(r["dones"])
            buf.truncations[t].copy_(r["trunc"])
            buf.terminations[t].copy_(r["term"])
        rows, rows1 = slice(e * T, (e + 1) * T), slice(e * (T + 1), (e + 1) * (T + 1))
        traj = buf.to_trajectory(max_episode_length=T)
        for k, sl in (("actions", rows), ("rewards", rows), ("prev_logprobs", rows), ("prev_values", rows1),
                      ("dones", rows1), ("terminations", rows1), ("truncations", rows1)):
            assert torch.equal(getattr(traj, k).cpu(), _t(g["traj_full_" + k])[sl]), (e, k)
        for k in ("states", "action"):
            assert torch.equal(traj.forward_inputs[k].cpu(), _t(g["traj_full_fi_" + k])[rows]), (e, k)
        parts = buf.to_splited_trajectories(split, max_episode_length=T)
        assert len(parts) == split
        for i, p in enumerate(parts):
            for k, sl in (("actions", rows), ("rewards", rows), ("dones", rows1), ("prev_values", rows1)):
                assert torch.equal(getattr(p, k).cpu(), _t(g[f"traj_part{i}_{k}"])[sl]), (e, i, k)
            assert torch.equal(p.forward_inputs["states"].cpu(), _t(g[f"traj_part{i}_fi_states"])[rows])
            assert p.actions.untyped_storage().data_ptr() == buf.actions.untyped_storage().data_ptr()  # a view
        batch = convert_trajectories_to_batch(parts)
        have = sorted(k for k in batch if isinstance(batch[k], torch.Tensor))
        assert set(have) <= set(keys) and {"actions", "rewards", "dones", "terminations", "truncations", "prev_logprobs",
                                           "prev_values"} <= set(have)
        for k in have:
            sl = rows1 if k in ("dones", "terminations", "truncations", "prev_values") else rows
            assert torch.equal(batch[k].cpu(), _t(g["traj_batch_" + k])[sl]), (e, k)
        assert batch["actions"].data_ptr() == buf.actions.data_ptr()  # adjacent views re-joined without a copy
        assert torch.equal(batch["forward_inputs"]["states"].cpu(), _t(g["traj_batch_fi_states"])[rows])
        # and the dict the actor consumes is the same thing
        ab = buf.as_batch()
        for k in ("rewards", "dones", "prev_values", "prev_logprobs"):
            assert torch.equal(ab[k].cpu(), batch[k].cpu()), k


def test_bucket_weight_sync_wire_format_round_trip(golden):
    """f1: actor -> rollout hand-off in BucketWeightSyncer's wire format: bucket layout + metadata equal the reference's
    (golden), payloads are views of the flat buffer, and a receiver policy ends up with identical parameters and the
    version; bf16 transport casts and comes back within bf16 rounding."""
    from rlinf_b200.policy import MLPPolicy
    from rlinf_b200.weight_sync import SYNCER_VERSION_KEY, TOTAL_BUCKETS_KEY, BucketWeightSync

    g = golden("r2")
    src, dst = MLPPolicy(128, 8, seed=1), MLPPolicy(128, 8, seed=2)
    assert [n for n, _ in src.named_parameters()] == [str(n) for n in g["bk_names"]]
    for tag, size, dt in (("200k", 200 * 1024, None), ("1k", 1024, None), ("bf16_300k", 300 * 1024, torch.bfloat16)):
        tx, rx = BucketWeightSync(src, size, dt), BucketWeightSync(dst, size, dt)
        buckets = list(tx.iter_buckets(version=7))
        layout = ["|".join(k for k in b if k not in (TOTAL_BUCKETS_KEY, SYNCER_VERSION_KEY)) for b in buckets]
        assert layout == [str(x) for x in g[f"bk_{tag}_layout"]], tag
        assert int(buckets[0][TOTAL_BUCKETS_KEY]) == int(g[f"bk_{tag}_total"]) == len(buckets)
        assert int(buckets[0][SYNCER_VERSION_KEY]) == 7 and buckets[0][SYNCER_VERSION_KEY].dtype == torch.int32
        if dt is None:  # zero-copy payload
            assert buckets[-1]["actor_mean.bias"].untyped_storage().data_ptr() == src.flat_params.untyped_storage().data_ptr()
        dst.flat_params.zero_()
        it = iter(buckets)
        assert rx.apply(lambda: next(it)) == 7
        if dt is None:
            assert torch.equal(dst.flat_params, src.flat_params)
        else:
            torch.testing.assert_close(dst.flat_params, src.flat_params, rtol=2 ** -8, atol=1e-8)
    # host-staged buckets (bucket_device = cpu in the reference) are accepted too; unknown keys are ignored (strict=False)
    tx, rx = BucketWeightSync(src, 1 << 30), BucketWeightSync(dst, 1 << 30)
    staged = [{k: v.cpu() for k, v in b.items()} for b in tx.iter_buckets(3)]
    staged[0]["some.other.buffer"] = torch.zeros(3)
    dst.flat_params.zero_()
    it = iter(staged)
    assert rx.apply(lambda: next(it)) == 3 and torch.equal(dst.flat_params, src.flat_params)
