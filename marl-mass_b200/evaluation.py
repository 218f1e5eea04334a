"""Batched evaluation protocol: `MAPPO.evaluation` (marl/mappo.py:255-361) with every
evaluation episode as one env of a single batch.

The reference runs `eval_episodes` episodes one after the other on `env_eval`, episode i from
`env.reset(is_training=False, testing_seeds=seeds[i], num_CAV=...)`, and returns
`(rewards, (vehicle_speed, vehicle_position), ext_info)`.  Here episode i is env i: the scenes are the reference's
own for those seeds (host replay of its MT19937 draws, `spawn.py`), all envs step together until the last one is
done, and each env's records stop at its own terminal step.  Same return structure, same keys.
"""
import time

import numpy as np

from ._lib import MAXV
from .env import MergeEnvBatched


def eval_num_cav(i, traffic_density):
    """The per-episode `num_CAV` argument of the training-time evaluation (mappo.py:280-286)."""
    return {1: (i + 1) % 3, 2: (i + 2) % 4, 3: (i + 4) % 6}[int(traffic_density)]


def evaluation(action_fn, config, test_seeds, eval_episodes=None, is_train=True, device=0, output_dir=None, tag="testing"):
    """action_fn(obs [E, 12, n_s] cuda f32, n_agents [E] cuda i32) -> actions [E, 12] (cuda, integer).

    Returns (rewards, (vehicle_speed, vehicle_position), ext_info) as `MAPPO.evaluation` does: rewards[i] is the list
    of global rewards of episode i; vehicle_speed[i] / vehicle_position[i] are arrays [steps_i, n_agents_i];
    ext_info has steps, avg_speeds, crash_count, step_time, min_headway, traffic_speeds, merge_percents.

    output_dir: where the reference records one video per episode (mappo.py:290-302, 321-327), this writes the same
    frames as PNG files `<tag>_episode_<i>/frame_<t>.png` (the first after reset, then one per policy step), rendered
    on the CPU from the device state (render.py); there is no video encoder in the image."""
    import torch
    seeds = [int(s) for s in (test_seeds.split(",") if isinstance(test_seeds, str) else test_seeds)]
    n_ep = len(seeds) if eval_episodes is None else int(eval_episodes)
    seeds = seeds[:n_ep]
    env = MergeEnvBatched(n_ep, config, device=device)
    td = env.config["traffic_density"]
    num_cav = [eval_num_cav(i, td) if is_train else 0 for i in range(n_ep)]
    env.reset_from_seeds(seeds, num_CAV=num_cav)
    v = env.buffers()
    n_agents = v["n_agents"].clone()
    n_host = n_agents.cpu().numpy()
    active = np.ones(n_ep, bool)
    rewards = [[] for _ in range(n_ep)]
    v_speed = [[] for _ in range(n_ep)]
    v_pos = [[] for _ in range(n_ep)]
    steps = np.zeros(n_ep, int)
    avg_speed = np.zeros(n_ep)
    traffic_speed = np.zeros(n_ep)
    merge_percent = np.full(n_ep, np.nan)
    crashed = np.zeros(n_ep, bool)
    min_headway = float("inf")
    t_total, n_steps = 0.0, 0

    def write_frames(st, which, t):
        import os
        from .render import render_scene, save_png
        for e in which:
            d = os.path.join(output_dir, "%s_episode_%d" % (tag, e))
            os.makedirs(d, exist_ok=True)
            save_png(os.path.join(d, "frame_%03d.png" % t), render_scene(st, int(e), env.config))
    if output_dir is not None:
        write_frames(env.get_state(), range(n_ep), 0)
    while active.any():
        t0 = time.process_time()
        a = action_fn(env.obs_view(), n_agents).to(torch.int8)
        env.step(a)
        torch.cuda.synchronize(device)
        t_total += time.process_time() - t0
        n_steps += 1
        st = env.get_state()
        out = {k: v[k].cpu().numpy() for k in ("reward", "done", "average_speed", "traffic_speed", "min_headway",
                                                "merge_percent")}
        if output_dir is not None:
            write_frames(st, np.nonzero(active)[0], n_steps)
        for e in np.nonzero(active)[0]:
            n = int(n_host[e])
            steps[e] += 1
            rewards[e].append(float(out["reward"][e]))
            avg_speed[e] += out["average_speed"][e]
            traffic_speed[e] += out["traffic_speed"][e]
            min_headway = min(min_headway, float(out["min_headway"][e]))
            v_speed[e].append(st["speed"][e, :n].copy())
            v_pos[e].append(st["x"][e, :n].copy())
            if out["done"][e]:
                active[e] = False
                merge_percent[e] = out["merge_percent"][e]
                crashed[e] = bool(st["crashed"][e, :n].any())
    env.close()
    ext_info = {"steps": [int(s) for s in steps], "avg_speeds": list(avg_speed / steps),
                "crash_count": [bool(c) for c in crashed], "step_time": [t_total / max(n_steps, 1)] * n_ep,
                "min_headway": min_headway, "traffic_speeds": list(traffic_speed / steps),
                "merge_percents": [float(m) for m in merge_percent]}
    return rewards, ([np.array(x) for x in v_speed], [np.array(x) for x in v_pos]), ext_info
