"""Batched training driver: `run_mappo.py --option train --config <ini>` (run_mappo.py:86-340) with the per-env Python
loop replaced by E envs on the device.

    python -m marl_mass_b200.train --config examples/mass_td3_srew.ini --envs 4096 --iterations 30

Reads the reference's ini layout (sections MODEL_CONFIG / TRAIN_CONFIG / ENV_CONFIG, the keys run_mappo.py:113-171
reads), builds `MergeEnvBatched` + `BatchedMAPPORollout` (MAPPO.interact / train for the whole batch) and, every
`--eval-interval` iterations, runs the reference's evaluation protocol on `test_seeds` as one batch
(`evaluation.evaluation` = MAPPO.evaluation) and prints what the reference logs (run_mappo.py:295-332).
"""
import argparse
import configparser
import json
import time

import numpy as np


def load_ini(path):
    """-> (env_config dict for MergeEnvBatched, rollout kwargs, train dict).  Same keys and fallbacks as run_mappo.py."""
    c = configparser.ConfigParser()
    if not c.read(path):
        raise FileNotFoundError(path)
    E = "ENV_CONFIG"
    env_cfg = {
        "env_name": c.get(E, "env_name", fallback="merge-multi-agent-v0"),
        "seed": c.getint(E, "seed"), "simulation_frequency": c.getint(E, "simulation_frequency"),
        "duration": c.getint(E, "duration"), "policy_frequency": c.getint(E, "policy_frequency"),
        "COLLISION_REWARD": c.getint(E, "COLLISION_REWARD"), "HIGH_SPEED_REWARD": c.getint(E, "HIGH_SPEED_REWARD"),
        "HEADWAY_COST": c.getint(E, "HEADWAY_COST"), "HEADWAY_TIME": c.getfloat(E, "HEADWAY_TIME"),
        "MERGING_LANE_COST": c.getint(E, "MERGING_LANE_COST"), "traffic_density": c.getint(E, "traffic_density"),
        "action_masking": c.getboolean("MODEL_CONFIG", "action_masking", fallback=False),
        "safety_guarantee": c.get(E, "safety_guarantee"), "lateral_control": c.get(E, "lateral_control", fallback="steer"),
        "mixed_traffic": c.getboolean(E, "mixed_traffic", fallback=None),
        "traffic_type": c.get(E, "traffic_type", fallback="cav"), "agent_reward": c.get(E, "agent_reward", fallback="default"),
        "cbf_eta": c.getfloat(E, "cbf_eta", fallback=0.0),          # CBFType.GAMMA_B (run_mappo.py:138)
    }
    M, T = "MODEL_CONFIG", "TRAIN_CONFIG"
    rollout_kw = {
        "roll_out_n_steps": c.getint(M, "ROLL_OUT_N_STEPS"), "reward_gamma": c.getfloat(M, "reward_gamma"),
        "max_grad_norm": c.getfloat(M, "MAX_GRAD_NORM"), "reward_type": c.get(M, "reward_type"),
        "reward_scale": c.getfloat(T, "reward_scale"), "actor_lr": c.getfloat(T, "actor_lr"),
        "critic_lr": c.getfloat(T, "critic_lr"),
    }
    train = {"test_seeds": c.get(T, "test_seeds", fallback=",".join(str(i) for i in range(0, 600, 20))),
             "eval_episodes": c.getint(T, "EVAL_EPISODES", fallback=20), "torch_seed": c.getint(M, "torch_seed", fallback=0),
             # run_mappo.py:126,232: shared_network selects MAPPO_GI with its one ActorCriticNetwork
             "shared_network": c.getboolean(M, "shared_network", fallback=False),
             "critic_hidden_size": c.getint(M, "critic_hidden_size", fallback=128)}
    return env_cfg, rollout_kw, train


def train(config, n_envs=4096, iterations=30, device=0, eval_interval=10, minibatch=1 << 18, log=print):
    import torch
    from . import evaluation as ev
    from .env import DEFAULT_CONFIG, MergeEnvBatched
    from .rollout import BatchedMAPPOGIRollout, BatchedMAPPORollout
    import torch.distributed as dist
    from . import dist as mmd
    env_cfg, rollout_kw, tr = load_ini(config) if isinstance(config, str) else config
    rank, world, local_rank = mmd.rank_world()
    if world > 1:                 # torchrun: one process per GPU, n_envs envs each, gradients averaged
        device = local_rank
        torch.cuda.set_device(device)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    torch.manual_seed(tr["torch_seed"])
    if env_cfg.get("env_name", "merge-multi-agent-v1") != "merge-multi-agent-v1":
        # env ids merge-multi-agent-v0 / -v05 observe 5 x 5 (no heading column, n_s = 25) and -hdv-v1 has no agents: the
        # networks of this driver and the fused actor kernels are the reference's 30-input ones (Model_common.py:11-23
        # with n_s = 30).  Rejected before anything is built (a 25-column state would otherwise be fed to 30 inputs).
        raise ValueError("train: env id %r is not covered by the batched learner (n_s = 30 env id merge-multi-agent-v1 only)"
                         % (env_cfg.get("env_name"),))
    env = MergeEnvBatched(n_envs, dict(DEFAULT_CONFIG, **env_cfg), device=device)
    # rank-distinct spawn stream; its first observation goes to the rollout so that collect() does not re-spawn the
    # scenes with the (rank-independent) config seed
    obs0, _ = env.reset(seed=mmd.rank_seed(env_cfg["seed"], rank) if world > 1 else env_cfg["seed"])
    if tr.get("shared_network"):
        pol = BatchedMAPPOGIRollout(env, hidden_size=tr.get("critic_hidden_size", 128),
                                    seed=tr["torch_seed"] + 104729 * rank, obs=obs0, **rollout_kw)
    else:
        pol = BatchedMAPPORollout(env, seed=tr["torch_seed"] + 104729 * rank, obs=obs0, **rollout_kw)
    pol.sync_parameters()
    if rank != 0:
        log = lambda *_: None
    draws = {"n": 0}

    def act(obs, n_agents):      # MAPPO.action (mappo.py:231-236): a draw from the softmax, also at evaluation time
        draws["n"] += 1
        return pol.sample_actions(obs, n_agents, seed=tr["torch_seed"] + 7919, step=draws["n"])

    history = []
    for it in range(iterations):
        t0 = time.perf_counter()
        env.stats(reset=True)
        pol.collect()
        torch.cuda.synchronize(device)
        t1 = time.perf_counter()
        up = pol.update(minibatch=minibatch)
        torch.cuda.synchronize(device)
        t2 = time.perf_counter()
        s = mmd.all_reduce_stats(env.stats())
        rec = {"iteration": it, "n_gpus": world, "agent_steps": s["agent_steps"], "collect_s": t1 - t0, "update_s": t2 - t1,
               "agent_steps_per_s": s["agent_steps"] / (t1 - t0), "mean_step_reward": s["reward_sum"] / max(s["env_steps"], 1),
               "crashed_episode_frac": s["crashed_episodes"] / max(s["episodes"], 1),
               "average_speed": s["speed_sum"] / max(s["env_steps"], 1), **up}
        if rank == 0 and eval_interval and (it % eval_interval == 0 or it == iterations - 1):
            rewards, _, info = ev.evaluation(act, env_cfg, tr["test_seeds"], eval_episodes=tr["eval_episodes"], device=device)
            rec.update({"eval_reward": float(np.mean([np.sum(r) for r in rewards])),       # run_mappo.py:303-306
                        "eval_avg_speed": float(np.mean(info["avg_speeds"])), "eval_crashes": int(np.sum(info["crash_count"])),
                        "eval_min_headway": info["min_headway"], "eval_merge_percent": float(np.mean(info["merge_percents"]))})
        if world > 1 and it == iterations - 1:      # the replicas must still hold identical networks
            flat = torch.cat([q.detach().reshape(-1) for m in pol.networks() for q in m.parameters()])
            hi, lo = flat.clone(), flat.clone()
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            rec["replica_param_max_diff"] = float((hi - lo).abs().max())
        history.append(rec)
        log(json.dumps(rec))
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return history


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--iterations", type=int, default=30)
    ap.add_argument("--eval-interval", type=int, default=10)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    train(a.config, a.envs, a.iterations, a.device, a.eval_interval)


if __name__ == "__main__":
    main()
