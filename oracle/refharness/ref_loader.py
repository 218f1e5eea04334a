"""Drive the UNMODIFIED reference (/root/reference) head-less in this container.

TEST INFRASTRUCTURE ONLY.  Used by oracle/refharness/gen_golden.py to freeze golden
vectors under tests/golden/ and by oracle/refharness/time_reference.py.  Nothing in the
product path, `-m gpu` tests, smoke() or bench.py imports this module: /root/reference does
not exist on the GPU box.

What is substituted:
  * import stubs for gym / pygame / matplotlib (oracle/refharness/stubs) — no behaviour;
  * `cvxopt` -> closed-form stand-in (stubs/cvxopt; "parity unpinned" vs real cvxopt 1.2.7);
  * numpy-1.19 aliases np.int/np.float/np.bool/np.long and pandas-1.1 DataFrame.append.
Everything else — env, road, vehicles, controllers, IDM/MOBIL, shields — is the reference's
own code, executed as is.

Instrumentation is by wrapping, never by editing: `safe_controller.safety_layer` and
`decentral_layer.multi_agent_state` are wrapped to record per-sub-step shield outputs
(neighbour ids, QP active set, veto flag, shielded action).
"""
import os
import sys

import numpy as np
import pandas as pd

REFERENCE_ROOT = os.environ.get("MM_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))

LANES = [("a", "b", 0), ("b", "c", 0), ("b", "c", 1), ("c", "d", 0), ("j", "k", 0), ("k", "b", 0)]
LANE_ID = {l: i for i, l in enumerate(LANES)}
ACTION_ID = {"LANE_LEFT": 0, "IDLE": 1, "LANE_RIGHT": 2, "FASTER": 3, "SLOWER": 4}

KIND_NONE, KIND_CAV, KIND_HDV, KIND_MDP, KIND_IDM = 0, 1, 2, 3, 4
NB_NONE, NB_OBSTACLE = -1, -2

_loaded = {}


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "highway_env"))


def load():
    """Import the reference once; returns a namespace dict of the modules we touch."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    for alias, typ in (("int", int), ("float", float), ("bool", bool), ("long", int)):
        if alias not in np.__dict__:
            setattr(np, alias, typ)
    if not hasattr(pd.DataFrame, "append"):
        def _append(self, other, ignore_index=False, **kw):
            return pd.concat([self, other], ignore_index=ignore_index)
        pd.DataFrame.append = _append
    sys.path.insert(0, os.path.join(_HERE, "stubs"))
    sys.path.insert(1, REFERENCE_ROOT)
    import gym
    import cvxopt
    import highway_env  # noqa: F401  (registers the env ids)
    from highway_env.vehicle import safe_controller, behavior, controller
    from highway_env.vehicle.safety import decentral_layer, cbf
    from highway_env.road import objects

    _loaded.update(gym=gym, cvxopt=cvxopt, highway_env=highway_env, safe_controller=safe_controller,
                   behavior=behavior, controller=controller, decentral_layer=decentral_layer, cbf=cbf,
                   objects=objects, shield_log=[])
    _install_wrappers()
    return _loaded


def _veh_index(road, veh):
    for i, v in enumerate(road.vehicles):
        if v is veh:
            return i
    return NB_NONE


def _install_wrappers():
    dl = _loaded["decentral_layer"]
    sc = _loaded["safe_controller"]
    log = _loaded["shield_log"]
    orig_mas = dl.multi_agent_state
    orig_layer = dl.safety_layer
    state = {}

    def mas_wrapper(cbf, vehicle, road, perception_dist, is_ma_dynamics=False):
        # hist[-2] objects are returned by identity; rear-adjacent is a fresh dict -> match by position
        before = {id(v.state_hist[-2]): i for i, v in enumerate(road.vehicles)
                  if hasattr(v, "state_hist") and len(v.state_hist) >= 2}
        out = orig_mas(cbf=cbf, vehicle=vehicle, road=road, perception_dist=perception_dist,
                       is_ma_dynamics=is_ma_dynamics)
        s_ol, s_oa, s_oar = out[0], out[1], out[2]

        def ident(s):
            if s is None:
                return NB_NONE
            if id(s) in before:
                return before[id(s)]
            return NB_OBSTACLE

        rear = NB_NONE
        if s_oar is not None:
            for i, v in enumerate(road.vehicles):
                if v is not vehicle and v.position[0] == s_oar["x"] and v.position[1] == s_oar["y"]:
                    rear = i
                    break
        state["nb"] = (ident(s_ol), ident(s_oa), rear)
        state["constrain_adj"] = bool(cbf.constrain_adj)
        return out

    def layer_wrapper(safety_type, action, vehicle, dt, safe_dist="theadway", **kwargs):
        trace = _loaded["cvxopt"].solvers.trace
        n0 = len(trace)
        state.clear()
        safe_action, safe_diff, status = orig_layer(safety_type=safety_type, action=action, vehicle=vehicle,
                                                    dt=dt, safe_dist=safe_dist, **kwargs)
        qp = trace[n0] if len(trace) > n0 else None
        nb = state.get("nb", (NB_NONE, NB_NONE, NB_NONE))
        log.append(dict(
            veh=_veh_index(vehicle.road, vehicle), leader=nb[0], front_adj=nb[1], rear_adj=nb[2],
            constrain_adj=state.get("constrain_adj", False),
            active=(qp["active"] if qp else 0),
            qp=(dict(qp) if qp else None),
            is_lc_safe=bool(vehicle.is_lc_safe),
            safe_acc=float(safe_action["acceleration"]), safe_steer=float(safe_action["steering"]),
            nom_acc=float(action["acceleration"]), nom_steer=float(action["steering"]),
            is_safe=float(status.get("is_safe", 1.0)), is_invariant=float(status.get("is_invariant", 1.0)),
            min_headway=float(vehicle.min_headway),
        ))
        return safe_action, safe_diff, status

    dl.multi_agent_state = mas_wrapper
    dl.safety_layer = layer_wrapper
    sc.safety_layer = layer_wrapper


# --------------------------------------------------------------------------------------
# env construction (mirrors run_mappo.py:137-171: gym.make, then mutate env.config)
# --------------------------------------------------------------------------------------
DEFAULT_ENV_CONFIG = dict(
    env_name="merge-multi-agent-v1", seed=0, simulation_frequency=15, duration=20, policy_frequency=5,
    COLLISION_REWARD=200, HIGH_SPEED_REWARD=1, HEADWAY_COST=4, HEADWAY_TIME=0.5, MERGING_LANE_COST=4,
    traffic_density=1, safety_guarantee="none", lateral_control="steer", mixed_traffic=None,
    traffic_type="cav", agent_reward="default", cbf_eta=0.03125, action_masking=False,
)


def make_env(**overrides):
    """gym.make + config mutation exactly as run_mappo.py:137-171 does; takes effect at reset()."""
    ns = load()
    cfg = dict(DEFAULT_ENV_CONFIG, **overrides)
    ns["cbf"].CBFType.GAMMA_B = float(cfg["cbf_eta"])
    ns["cbf"].CBFType.TAU = float(cfg["HEADWAY_TIME"])
    env = ns["gym"].make(cfg["env_name"])
    for k in ("seed", "simulation_frequency", "duration", "policy_frequency", "COLLISION_REWARD",
              "HIGH_SPEED_REWARD", "HEADWAY_COST", "HEADWAY_TIME", "MERGING_LANE_COST", "traffic_density",
              "action_masking", "safety_guarantee", "lateral_control", "mixed_traffic", "traffic_type",
              "agent_reward"):
        env.config[k] = cfg[k]
    env.seed = cfg["seed"]
    return env


# --------------------------------------------------------------------------------------
# state export (teacher forcing: reference state -> SoA arrays our engines load)
# --------------------------------------------------------------------------------------
MAXV = 12
F64_FIELDS = ("x", "y", "heading", "speed", "target_speed", "gvx", "rec1_x", "rec1_vx", "rec2_x", "rec2_vx",
              "act_steer", "act_acc", "safe_steer", "safe_acc", "timer", "min_headway", "steering_angle")
I32_FIELDS = ("kind", "lane", "target_lane", "speed_index", "crashed", "hl_action", "hist_len", "fg_set",
              "is_collaborating", "is_lc_safe", "collaborate_adj")


def _kind(ns, v):
    if isinstance(v, ns["safe_controller"].MDPLCVehicle):
        return KIND_CAV
    if isinstance(v, ns["behavior"].IDMVehicleHist):
        return KIND_HDV
    if isinstance(v, ns["controller"].MDPVehicle):
        return KIND_MDP
    if isinstance(v, ns["behavior"].IDMVehicle):
        return KIND_IDM
    raise TypeError(type(v))


def export_state(env):
    """Snapshot every field the step path reads.  Returns dict of numpy arrays ([MAXV] per field)."""
    ns = load()
    vs = env.road.vehicles
    assert len(vs) <= MAXV
    out = {k: np.zeros(MAXV, np.float64) for k in F64_FIELDS}
    out.update({k: np.zeros(MAXV, np.int32) for k in I32_FIELDS})
    out["hl_action"][:] = -1
    for i, v in enumerate(vs):
        out["kind"][i] = _kind(ns, v)
        out["x"][i], out["y"][i] = v.position
        out["heading"][i] = v.heading
        out["speed"][i] = v.speed
        out["target_speed"][i] = v.target_speed
        out["lane"][i] = LANE_ID[tuple(v.lane_index)]
        out["target_lane"][i] = LANE_ID[tuple(v.target_lane_index)]
        out["speed_index"][i] = getattr(v, "speed_index", -1)
        out["crashed"][i] = int(v.crashed)
        out["act_steer"][i] = float(v.action["steering"])
        out["act_acc"][i] = float(v.action["acceleration"])
        out["timer"][i] = getattr(v, "timer", 0.0)
        hist = getattr(v, "state_hist", [])
        out["hist_len"][i] = min(len(hist), 2)
        if len(hist) >= 1:
            out["rec1_x"][i], out["rec1_vx"][i] = hist[-1]["x"], hist[-1]["vx"]
        if len(hist) >= 2:
            out["rec2_x"][i], out["rec2_vx"][i] = hist[-2]["x"], hist[-2]["vx"]
        if out["kind"][i] == KIND_CAV:
            out["hl_action"][i] = ACTION_ID.get(v.hl_action, -1) if v.hl_action is not None else -1
            out["safe_steer"][i] = float(v.safe_action["steering"])
            out["safe_acc"][i] = float(v.safe_action["acceleration"])
            out["fg_set"][i] = int(v.fg_params is not None)
            out["gvx"][i] = v.fg_params["g"]["vx"] if v.fg_params is not None else 0.0
            out["is_collaborating"][i] = int(v.is_collaborating)
            out["is_lc_safe"][i] = int(v.is_lc_safe)
            out["collaborate_adj"][i] = int(bool(v.collaborate_adj))
            out["min_headway"][i] = v.min_headway
            out["steering_angle"][i] = v.steering_angle
    out["n_veh"] = np.int32(len(vs))
    out["n_cav"] = np.int32(len(env.controlled_vehicles))
    out["n_merge"] = np.int32(env.n_merge)
    out["steps"] = np.int32(env.steps)
    out["time"] = np.int32(env.time)
    # controlled_vehicles are always the first n_cav entries of road.vehicles (merge_env_v1.py:327-343)
    for i, cv in enumerate(env.controlled_vehicles):
        assert cv is vs[i]
    return out


def step_outputs_hdv(env, obs, reward, done, info):
    """MergeEnvLCHDV.step (merge_env_v1.py:603-665): one observation row per vehicle, the reward is the mean of
    `_agent_reward` over ALL vehicles (`_reward`, 518-524); the per-vehicle terms are recomputed here with the
    reference's own `_agent_reward` so that they can be pinned too.  No regional rewards / per-agent dones."""
    n = len(env.road.vehicles)
    local = np.array([env._agent_reward(None, v) for v in env.road.vehicles], np.float64)
    assert abs(local.mean() - reward) < 1e-12
    return dict(
        obs=np.asarray(obs, np.float64).reshape(n, -1),
        reward=np.float64(reward), done=np.int32(bool(done)),
        agents_rewards=local, regional_rewards=np.zeros(n), agents_dones=np.zeros(n, np.int32),
        average_speed=np.float64(info["average_speed"]), traffic_speed=np.float64(info["traffic_speed"]),
        min_headway=np.float64(info["min_headway"]), merge_percent=np.float64(info.get("merge_percent", -1.0)),
        action_mask=np.zeros((n, 5), np.int32),
    )


def step_outputs(env, obs, reward, done, info):
    """Flatten what MergeEnv.step returns (merge_env_v1.py:126-166) into arrays."""
    n = len(env.controlled_vehicles)
    out = dict(
        obs=np.asarray(obs, np.float64).reshape(n, -1),
        reward=np.float64(reward), done=np.int32(bool(done)),
        agents_rewards=np.asarray(info["agents_rewards"], np.float64),
        regional_rewards=np.asarray(info["regional_rewards"], np.float64),
        agents_dones=np.asarray(info["agents_dones"], np.int32),
        average_speed=np.float64(info["average_speed"]),
        traffic_speed=np.float64(info["traffic_speed"]),
        min_headway=np.float64(info["min_headway"]),
        merge_percent=np.float64(info.get("merge_percent", -1.0)),
        action_mask=np.asarray(info["action_mask"], np.int32),
    )
    return out


def drain_shield_log():
    ns = load()
    log = list(ns["shield_log"])
    del ns["shield_log"][:]
    del ns["cvxopt"].solvers.trace[:]
    return log
