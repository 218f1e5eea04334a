#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) here.

    python oracle/refharness/gen_golden.py            # all cases
    python oracle/refharness/gen_golden.py mass_td3   # one case

Each file freezes, for a handful of episodes with i.i.d. uniform meta-actions:
  st_<field>[S, MAXV]   full vehicle state before every policy step and after the last one of
                        each episode (S = sum over episodes of T_ep + 1), see ref_loader.export_state
  st_{steps,time,n_veh,n_cav,n_merge}[S]
  ep_start[n_ep+1]      offsets into the S axis (episode j owns states ep_start[j]..ep_start[j+1]-1)
  act[T, MAXV] int8     meta-actions applied at step t (-1 padding)
  row_of_step[T]        index on the S axis of the pre-state of step t (post-state is row+1)
  obs[T, MAXV, n_s], reward[T], done[T], agents_rewards/regional_rewards/agents_dones[T, MAXV],
  average_speed/traffic_speed/min_headway/merge_percent[T]       -- MergeEnv.step return values
  avail_bits[T, MAXV]   per-agent _get_available_actions bitmask on the post-step state
  sh_<field>[T, 3, MAXV] per-sub-step shield record (ran, leader, front_adj, rear_adj,
                        constrain_adj, active, is_lc_safe, safe_acc, safe_steer, nom_acc, nom_steer)
  qp_{a,c_lead,c_adj,has_adj,lo,hi,u,active}[Q]                  -- every QP the reference posed
  config (json)

Test infrastructure; the committed fixtures are what travels to the GPU box.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_loader as rl  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden")

# name -> (config overrides, reset seeds, action-rng seed)
CASES = {
    # BASELINE configs[0] sibling on the LC env (test-configs_marl-cav-heading-unsafe.ini)
    "unsafe_td1": (dict(safety_guarantee="none", traffic_density=1, HEADWAY_TIME=1.2), [0, 7, 11], 100),
    "unsafe_td3": (dict(safety_guarantee="none", traffic_density=3, HEADWAY_TIME=1.2), [3, 4], 101),
    "unsafe_td2_mixed": (dict(safety_guarantee="none", traffic_density=2, traffic_type="mixed",
                              mixed_traffic=True), [5, 6, 21], 102),
    # BASELINE configs[1]: marl_cav-heading-t_headway-cbf-avs_cint.ini
    "hss_td3": (dict(safety_guarantee="cbf-avs_cint", traffic_density=3), [0, 20], 103),
    "hss_td3_mixed": (dict(safety_guarantee="cbf-avs_cint", traffic_density=3, traffic_type="mixed",
                           mixed_traffic=True), [40, 60], 104),
    # BASELINE configs[2]: marl_cav-heading-t_headway-cbf-cav.ini
    "mass_td1": (dict(safety_guarantee="cbf-cav", traffic_density=1, mixed_traffic=False), [0, 25, 50], 105),
    # BASELINE configs[3]: marl_cav-heading-t_headway-cbf-cav-td3-srew.ini
    "mass_td3_srew": (dict(safety_guarantee="cbf-cav", traffic_density=3, mixed_traffic=False,
                           agent_reward="srew", HIGH_SPEED_REWARD=4, HEADWAY_COST=1, MERGING_LANE_COST=8),
                      [0, 75], 106),
    # BASELINE configs[4] mixed variant: marl_cav-heading-t_headway-cbf-cav-mixed.ini
    "mass_td3_mixed": (dict(safety_guarantee="cbf-cav", traffic_density=3, mixed_traffic=True,
                            traffic_type="mixed"), [100, 125], 107),
    # marl_cav-heading-t_headway-cbf-cav-mixed-mrew.ini
    "mass_td2_mixed_mrew": (dict(safety_guarantee="cbf-cav", traffic_density=2, mixed_traffic=True,
                                 traffic_type="mixed", agent_reward="mrew", HIGH_SPEED_REWARD=4,
                                 HEADWAY_COST=1, MERGING_LANE_COST=8), [150, 175], 108),
    # BASELINE configs[0]: test-configs_marl-cav-unsafe.ini (env merge-multi-agent-v0: MDPVehicle, 5x5 observation)
    "v0_unsafe_td1": (dict(env_name="merge-multi-agent-v0", safety_guarantee="none", traffic_density=1,
                           HEADWAY_TIME=1.2, mixed_traffic=False), [0, 9, 13], 109),
    "v0_unsafe_td2_mixed": (dict(env_name="merge-multi-agent-v0", safety_guarantee="none", traffic_density=2,
                                 HEADWAY_TIME=1.2, mixed_traffic=True), [2, 8], 110),
    # lateral_control = steer_vel (configs_marl-cav-heading-unsafe-steer_vel.ini, marl_cav_heading-t_headway-cbf-av-steer_vel.ini,
    # configs_marl-cav-t_headway-unsafe-steer_vel.ini on env merge-multi-agent-v05)
    "steervel_unsafe_td2": (dict(safety_guarantee="none", traffic_density=2, HEADWAY_TIME=1.2,
                                 lateral_control="steer_vel"), [1, 2], 111),
    "steervel_hss_td3_mixed": (dict(safety_guarantee="cbf-av", traffic_density=3, traffic_type="mixed",
                                    mixed_traffic=True, lateral_control="steer_vel"), [3, 5], 112),
    "steervel_mass_td2": (dict(safety_guarantee="cbf-cav", traffic_density=2, mixed_traffic=False,
                               lateral_control="steer_vel"), [4, 6], 113),
    "v05_steervel_unsafe_td1": (dict(env_name="merge-multi-agent-v05", safety_guarantee="none", traffic_density=1,
                                     HEADWAY_TIME=1.2, lateral_control="steer_vel"), [7, 8], 114),
    # traffic_type = av (merge_env_v1.py:485-489): one shielded CAV among IDM / MOBIL vehicles
    # test-idm-td3.ini: env merge-multi-agent-hdv-v1 (MergeEnvLCHDV), traffic_type = hdv: IDM / MOBIL vehicles only
    "hdv_td3": (dict(env_name="merge-multi-agent-hdv-v1", safety_guarantee="cbf-cav", traffic_density=3,
                     traffic_type="hdv", mixed_traffic=True), [11, 12, 13], 116),
    "mass_td3_av": (dict(safety_guarantee="cbf-cav", traffic_density=3, traffic_type="av", mixed_traffic=True),
                    [30, 31, 32, 33, 34], 115),
    # exact ties: before every policy step all x positions and speeds are snapped to integers, so that vehicles of
    # different lanes share x / s, |ds| keys of the closest-vehicle sorts are equal on both sides of an ego, and equal
    # speeds keep the ties alive through the sub-steps.  Pins the reference's tie rules (stable sorts, "<=" scans).
    # Every step is its own one-step "episode" (pre-state, post-state) because the snap breaks the chain.
    "ties_mass_td3": (dict(safety_guarantee="cbf-cav", traffic_density=3, mixed_traffic=False), [40, 41], 116, "snap"),
    "ties_hss_td3_mixed": (dict(safety_guarantee="cbf-avs_cint", traffic_density=3, traffic_type="mixed",
                                mixed_traffic=True), [42, 43], 117, "snap"),
    # the same plus speeds on a 2.5 m/s grid (half-way between speed levels) and y snapped to a 0.5 m grid: vehicles exactly half-way between bc0 and bc1 during a lane change (the
    # closest-lane argmin ties and list order decides), lateral offsets exactly on the on_lane / is_lc margins
    # x on a 0.5 m grid: vehicles exactly on the after_end thresholds x = 217.5 / 317.5 / 417.5 (s > len - 2.5 is strict,
    # lane.py:95) as well as on the lane ends
    "ties_half_mass_td3_mixed": (dict(safety_guarantee="cbf-cav", traffic_density=3, traffic_type="mixed",
                                      mixed_traffic=True), [50, 51, 52], 120, "snaph"),
    # the all-HDV env with the snapping (IDM / MOBIL only; every vehicle observed)
    "ties_hdv_td3": (dict(env_name="merge-multi-agent-hdv-v1", safety_guarantee="cbf-cav", traffic_density=3,
                          traffic_type="hdv", mixed_traffic=True), [14, 15], 123, "snapy"),
    # the baseline supervisors (central_layer.py / decentralised_dmc.py, the 8 priority / dmc ini files): they only
    # REPLACE the meta-actions before _simulate (abstract.py:459-467).  The fixtures hold the policy's tuple (`act`), the
    # supervised tuple the env then executed (`new_act` = info["new_action"]) and the np.random.rand() draws the
    # supervisor consumed (`rand_draws`): groundwork for a supervisor kernel, and today a check that the v0 dynamics
    # under the supervised actions are the un-shielded dynamics
    "priority_v0_td3_mixed": (dict(env_name="merge-multi-agent-v0", safety_guarantee="priority", traffic_density=3,
                                   HEADWAY_TIME=1.2, mixed_traffic=True), [1, 2, 3, 4, 5, 6, 7], 121),
    "dmc_v0_td3_mixed": (dict(env_name="merge-multi-agent-v0", safety_guarantee="dmc", traffic_density=3,
                              HEADWAY_TIME=1.2, mixed_traffic=True), [1, 2, 3, 4, 5, 6, 7], 122),
    # the v0 env (MDPVehicle / IDMVehicle, no history, no shield) with the same snapping
    "ties_v0_unsafe_td2_mixed": (dict(env_name="merge-multi-agent-v0", safety_guarantee="none", traffic_density=2,
                                      HEADWAY_TIME=1.2), [47, 48], 119, "snapy"),
    "ties_y_mass_td3_mixed": (dict(safety_guarantee="cbf-cav", traffic_density=3, traffic_type="mixed",
                                   mixed_traffic=True), [44, 45, 46], 118, "snapy"),
}

SH_I = ("ran", "leader", "front_adj", "rear_adj", "constrain_adj", "active", "is_lc_safe")
SH_F = ("safe_acc", "safe_steer", "nom_acc", "nom_steer")


def run_case(name):
    overrides, seeds, aseed = CASES[name][:3]
    snap = len(CASES[name]) > 3 and CASES[name][3] in ("snap", "snapy", "snaph")
    snap_h = len(CASES[name]) > 3 and CASES[name][3] == "snaph"
    snap_y = len(CASES[name]) > 3 and CASES[name][3] == "snapy"
    hdv_env = overrides.get("env_name") == "merge-multi-agent-hdv-v1"
    env = rl.make_env(**overrides)
    rl.drain_shield_log()
    M = rl.MAXV
    n_s = env.n_s
    states, ep_start, acts, rows = [], [0], [], []
    outs = {k: [] for k in ("obs", "reward", "done", "agents_rewards", "regional_rewards", "agents_dones",
                            "average_speed", "traffic_speed", "min_headway", "merge_percent", "avail_bits")}
    sh = {k: [] for k in SH_I + SH_F}
    qps = []
    arng = np.random.RandomState(aseed)
    supervised = overrides.get("safety_guarantee") in ("priority", "dmc")
    new_acts, draws = [], []
    for seed in seeds:
        env.reset(is_training=False, testing_seeds=seed)
        rl.drain_shield_log()
        done = False
        while not done:
            if snap:
                for veh in env.road.vehicles:
                    veh.position[0] = float(np.round(veh.position[0] * 2) / 2 if snap_h else np.round(veh.position[0]))
                    veh.speed = float(np.round(veh.speed))
                    if snap_y:
                        veh.position[1] = float(np.round(veh.position[1] * 2) / 2)
                        # speeds on a 2.5 m/s grid: 12.5, 17.5, 22.5, 27.5 sit exactly between two speed levels, where
                        # FASTER / SLOWER round half to even (controller.py:302-305, 327-337)
                        veh.speed = float(np.round(veh.speed / 2.5) * 2.5)
                    # the newest history record IS the current state (log_step after every move,
                    # safe_controller.py:187-205, behavior.py:509-519): keep that invariant
                    if getattr(veh, "state_hist", None):
                        veh.state_hist[-1].update(veh.to_dict())
            st = rl.export_state(env)
            rows.append(len(states))
            states.append(st)
            n = int(st["n_cav"])
            a = arng.randint(0, 5, size=n)
            if supervised:      # log the global-generator draws of the supervisor (priority tie-breaks)
                log_rand, real_rand = [], np.random.rand

                def logging_rand(*shape):
                    v = real_rand(*shape)
                    log_rand.append(float(v))
                    return v
                np.random.rand = logging_rand
            try:
                obs, reward, done, info = env.step(tuple(int(x) for x in a))
            finally:
                if supervised:
                    np.random.rand = real_rand
            if supervised:
                na = np.full(M, -1, np.int8)
                na[:n] = np.asarray(info["new_action"], np.int64)
                new_acts.append(na)
                d = np.full(16, np.nan)
                d[:len(log_rand)] = log_rand
                draws.append(d)
            if hdv_env:     # every vehicle is observed and rewarded, nobody is controlled
                o = rl.step_outputs_hdv(env, obs, reward, done, info)
                n = int(st["n_veh"])
                a = np.full(n, -1)
            else:
                o = rl.step_outputs(env, obs, reward, done, info)
            ap = np.full(M, -1, np.int8)
            ap[:n] = a
            acts.append(ap)
            op = np.zeros((M, n_s))
            op[:n] = o["obs"]
            outs["obs"].append(op)
            for k in ("agents_rewards", "regional_rewards", "agents_dones"):
                p = np.zeros(M, o[k].dtype)
                p[:n] = o[k]
                outs[k].append(p)
            for k in ("reward", "done", "average_speed", "traffic_speed", "min_headway", "merge_percent"):
                outs[k].append(o[k])
            # per-agent _get_available_actions (abstract.py:219-240) on the post-step state, as a bitmask
            ab = np.zeros(M, np.int32)
            for i, cv in enumerate(env.controlled_vehicles):
                for act_id in env._get_available_actions(cv, env):
                    ab[i] |= 1 << int(act_id)
            outs["avail_bits"].append(ab)
            # shield records of this policy step, split by sub-step (order of execution per sub-step:
            # each CAV at most once, so a repeated vehicle id starts a new sub-step)
            log = rl.drain_shield_log()
            rec_i = {k: np.full((3, M), -1 if k in ("leader", "front_adj", "rear_adj") else 0, np.int32)
                     for k in SH_I}
            rec_f = {k: np.zeros((3, M)) for k in SH_F}
            # sub-step index: shield is skipped while len(state_hist) < 2, i.e. sub-steps 0,1 of the episode
            first_sub = 2 if int(st["steps"]) == 0 else 0
            sub, seen = first_sub, set()
            for e in log:
                if e["veh"] in seen:
                    sub, seen = sub + 1, set()
                seen.add(e["veh"])
                assert sub < 3, (name, seed, sub)
                v = e["veh"]
                rec_i["ran"][sub, v] = 1
                for k in ("leader", "front_adj", "rear_adj", "constrain_adj", "active", "is_lc_safe"):
                    rec_i[k][sub, v] = int(e[k])
                for k in SH_F:
                    rec_f[k][sub, v] = e[k]
                q = e["qp"]
                qps.append((q["a"], q["c_lead"], q["c_adj"] if q["c_adj"] is not None else 0.0,
                            float(q["c_adj"] is not None), q["lo"], q["hi"], q["u"], float(q["active"])))
            for k in SH_I:
                sh[k].append(rec_i[k])
            for k in SH_F:
                sh[k].append(rec_f[k])
            if snap:
                states.append(rl.export_state(env))
                ep_start.append(len(states))
        if not snap:
            states.append(rl.export_state(env))
            ep_start.append(len(states))
    data = {}
    for k in rl.F64_FIELDS + rl.I32_FIELDS:
        data["st_" + k] = np.stack([s[k] for s in states])
    for k in ("steps", "time", "n_veh", "n_cav", "n_merge"):
        data["st_" + k] = np.array([s[k] for s in states], np.int32)
    data["ep_start"] = np.array(ep_start, np.int32)
    data["act"] = np.stack(acts)
    data["row_of_step"] = np.array(rows, np.int32)
    if supervised:
        data["new_act"] = np.stack(new_acts)
        data["rand_draws"] = np.stack(draws)
    for k, v in outs.items():
        data[k] = np.stack(v) if np.ndim(v[0]) else np.array(v)
    for k in SH_I + SH_F:
        data["sh_" + k] = np.stack(sh[k])
    q = np.array(qps, np.float64).reshape(-1, 8)
    for j, k in enumerate(("a", "c_lead", "c_adj", "has_adj", "lo", "hi", "u", "active")):
        data["qp_" + k] = q[:, j]
    cfg = dict(rl.DEFAULT_ENV_CONFIG, **overrides)
    cfg["n_s"] = n_s
    cfg["seeds"] = seeds
    data["config"] = np.array(json.dumps(cfg))
    os.makedirs(OUT_DIR, exist_ok=True)
    path = os.path.join(OUT_DIR, name + ".npz")
    np.savez_compressed(path, **data)
    T = len(acts)
    crashed = int(np.sum(data["st_crashed"][np.array(ep_start[1:]) - 1].max(axis=1) > 0))
    print("%-22s episodes=%d steps=%d qps=%d crashed_eps=%d vetoes=%d active=%d size=%.0fKB" % (
        name, len(seeds), T, len(qps), crashed,
        int(np.sum((data["sh_ran"] == 1) & (data["sh_is_lc_safe"] == 0))),
        int(np.sum(data["sh_active"] > 0)), os.path.getsize(path) / 1024))


def check_geometry():
    """The constants our engines hard-code (SURVEY.md §8 'Fixed scenario constants')."""
    env = rl.make_env()
    env.reset()
    net = env.road.network
    expect = {("a", "b", 0): (0, 0, 320), ("b", "c", 0): (320, 0, 100), ("b", "c", 1): (320, 4, 100),
              ("c", "d", 0): (420, 0, 1000), ("j", "k", 0): (0, 10.5, 220), ("k", "b", 0): (220, 7.25, 100)}
    for l, (sx, sy, ln) in expect.items():
        lane = net.get_lane(l)
        assert lane.start[0] == sx and lane.start[1] == sy and lane.length == ln, (l, lane.start, lane.length)
        assert lane.direction[0] == 1.0 and lane.direction[1] == 0.0 and lane.heading == 0.0
    kb = net.get_lane(("k", "b", 0))
    assert kb.amplitude == 3.25 and kb.pulsation == 2 * np.pi / 200 and kb.phase == np.pi / 2
    ob = env.road.objects[0]
    assert ob.position[0] == 420.0 and ob.position[1] == 4.0 and len(env.road.objects) == 1


if __name__ == "__main__":
    check_geometry()
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        run_case(nm)
