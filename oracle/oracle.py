"""ctypes binding of the CPU oracle (oracle/merge_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never by the product package.  State and outputs are dicts of numpy arrays laid out
env-major ([n_env, MAXV] per vehicle field), the same layout as tests/golden/*.npz.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmerge_oracle.so")

MAXV = 12
NS = 30
F64_FIELDS = ("x", "y", "heading", "speed", "target_speed", "gvx", "rec1_x", "rec1_vx", "rec2_x", "rec2_vx",
              "act_steer", "act_acc", "safe_steer", "safe_acc", "timer", "min_headway", "steering_angle")
I32_FIELDS = ("kind", "lane", "target_lane", "speed_index", "crashed", "hl_action", "hist_len", "fg_set",
              "is_collaborating", "is_lc_safe", "collaborate_adj")
ENV_FIELDS = ("n_veh", "n_cav", "n_merge", "steps", "time")
SH_I = ("ran", "leader", "front_adj", "rear_adj", "constrain_adj", "active", "is_lc_safe")
SH_F = ("safe_acc", "safe_steer", "nom_acc", "nom_steer", "lc_margin")

SHIELD = {"none": 0, "priority": 0, "dmc": 0, "cbf-hss": 1, "cbf-av": 1, "cbf-avs": 1, "cbf-avs_cint": 1,
          "cbf-mass": 2, "cbf-cav": 2}
REWARD = {"default": 0, "srew": 1, "mrew": 2}

_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)


class MoConfig(C.Structure):
    _fields_ = [("shield", C.c_int32), ("reward_kind", C.c_int32), ("duration_steps", C.c_int32),
                ("substeps", C.c_int32), ("dt", C.c_double), ("eta", C.c_double), ("tau", C.c_double),
                ("collision_reward", C.c_double), ("high_speed_reward", C.c_double),
                ("headway_cost", C.c_double), ("headway_time", C.c_double), ("merging_lane_cost", C.c_double),
                ("env_v0", C.c_int32), ("steer_vel", C.c_int32), ("env_hdv", C.c_int32)]


class MoState(C.Structure):
    _fields_ = [(k, _PD) for k in F64_FIELDS] + [(k, _PI) for k in I32_FIELDS] + [(k, _PI) for k in ENV_FIELDS]


class MoOut(C.Structure):
    _fields_ = [("obs", _PD), ("reward", _PD), ("done", _PI), ("agents_rewards", _PD), ("regional_rewards", _PD),
                ("agents_dones", _PI), ("average_speed", _PD), ("traffic_speed", _PD), ("min_headway", _PD),
                ("merge_percent", _PD)] + [("sh_" + k, _PI) for k in SH_I] + [("sh_" + k, _PD) for k in SH_F]


_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "merge_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.mo_step.argtypes = [C.POINTER(MoConfig), C.POINTER(MoState), C.POINTER(C.c_int8),
                                 C.POINTER(MoOut), C.c_int, C.c_int]
        _lib.mo_step.restype = None
        _lib.mo_observe.argtypes = [C.POINTER(MoState), _PD, C.c_int, C.c_int, C.c_int]
        _lib.mo_observe.restype = None
        _lib.mo_qp.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, _PI]
        _lib.mo_qp.restype = C.c_double
    return _lib


def make_config(cfg):
    """cfg: dict with the reference's ENV_CONFIG keys (run_mappo.py:142-171)."""
    sim, pol = int(cfg.get("simulation_frequency", 15)), int(cfg.get("policy_frequency", 5))
    v0 = cfg.get("env_name", "merge-multi-agent-v1") == "merge-multi-agent-v0"
    return MoConfig(
        env_hdv=int(cfg.get("env_name") == "merge-multi-agent-hdv-v1"), env_v0=int(v0), steer_vel=int(cfg.get("lateral_control", "steer") == "steer_vel" and not v0),
        shield=0 if v0 else SHIELD[cfg.get("safety_guarantee", "none")],
        reward_kind=0 if v0 else REWARD[cfg.get("agent_reward", "default")],
        duration_steps=int(cfg.get("duration", 20) * pol), substeps=sim // pol, dt=1 / sim,
        eta=float(cfg.get("cbf_eta", 0.0)), tau=float(cfg.get("HEADWAY_TIME", 1.2)),
        collision_reward=float(cfg.get("COLLISION_REWARD", 200)),
        high_speed_reward=float(cfg.get("HIGH_SPEED_REWARD", 1)), headway_cost=float(cfg.get("HEADWAY_COST", 4)),
        headway_time=float(cfg.get("HEADWAY_TIME", 1.2)), merging_lane_cost=float(cfg.get("MERGING_LANE_COST", 4)))


def empty_state(n_env):
    st = {k: np.zeros((n_env, MAXV), np.float64) for k in F64_FIELDS}
    st.update({k: np.zeros((n_env, MAXV), np.int32) for k in I32_FIELDS})
    st.update({k: np.zeros(n_env, np.int32) for k in ENV_FIELDS})
    st["hl_action"][:] = -1
    return st


def _c_state(st):
    for k in F64_FIELDS:
        assert st[k].dtype == np.float64 and st[k].flags.c_contiguous, k
    for k in I32_FIELDS + ENV_FIELDS:
        assert st[k].dtype == np.int32 and st[k].flags.c_contiguous, k
    return MoState(**{k: st[k].ctypes.data_as(_PD) for k in F64_FIELDS},
                   **{k: st[k].ctypes.data_as(_PI) for k in I32_FIELDS + ENV_FIELDS})


def empty_out(n_env):
    o = dict(obs=np.zeros((n_env, MAXV, NS)), reward=np.zeros(n_env), done=np.zeros(n_env, np.int32),
             agents_rewards=np.zeros((n_env, MAXV)), regional_rewards=np.zeros((n_env, MAXV)),
             agents_dones=np.zeros((n_env, MAXV), np.int32), average_speed=np.zeros(n_env),
             traffic_speed=np.zeros(n_env), min_headway=np.zeros(n_env), merge_percent=np.zeros(n_env))
    for k in SH_I:
        o["sh_" + k] = np.zeros((n_env, 3, MAXV), np.int32)
    for k in SH_F:
        o["sh_" + k] = np.zeros((n_env, 3, MAXV), np.float64)
    return o


def _c_out(o):
    kw = {}
    for name, typ in MoOut._fields_:
        kw[name] = o[name].ctypes.data_as(typ)
    return MoOut(**kw)


def step(cfg, st, actions, out=None, n_threads=1):
    """Advance every env in `st` (in place) by one policy step; returns the outputs dict."""
    n_env = st["n_veh"].shape[0]
    if not isinstance(cfg, MoConfig):
        cfg = make_config(cfg)
    actions = np.ascontiguousarray(actions, np.int8).reshape(n_env, MAXV)
    if out is None:
        out = empty_out(n_env)
    cs, co = _c_state(st), _c_out(out)
    lib().mo_step(C.byref(cfg), C.byref(cs), actions.ctypes.data_as(C.POINTER(C.c_int8)), C.byref(co),
                  n_env, int(n_threads))
    return out


def observe(st, steer_vel=False, env_hdv=False):
    n_env = st["n_veh"].shape[0]
    obs = np.zeros((n_env, MAXV, NS))
    cs = _c_state(st)
    lib().mo_observe(C.byref(cs), obs.ctypes.data_as(_PD), n_env, int(steer_vel), int(env_hdv))
    return obs


def qp(a, c_lead, c_adj, has_adj, lo, hi):
    act = C.c_int32(0)
    u = lib().mo_qp(a, c_lead, c_adj, int(has_adj), lo, hi, C.byref(act))
    return u, act.value


def state_from_golden(g, rows):
    """Rows of a golden file's S axis -> oracle state dict (copy)."""
    rows = np.atleast_1d(rows)
    st = {}
    for k in F64_FIELDS:
        st[k] = np.ascontiguousarray(g["st_" + k][rows], np.float64)
    for k in I32_FIELDS:
        st[k] = np.ascontiguousarray(g["st_" + k][rows], np.int32)
    for k in ENV_FIELDS:
        st[k] = np.ascontiguousarray(g["st_" + k][rows], np.int32)
    # the engines know two kinds: CAV (MDPLCVehicle, or MDPVehicle in the v0 env) and HDV (IDMVehicle[Hist])
    st["kind"] = np.where(st["kind"] == 3, 1, np.where(st["kind"] == 4, 2, st["kind"])).astype(np.int32)
    return st


# --------------------------------------------------------------------------------------------------
# caller-side restatements: numpy float64, pinned against tests/golden/mappo_caller.npz
# --------------------------------------------------------------------------------------------------
def mappo_discount(rewards, dones, final_value, gamma, cols_per_env=1):
    """marl/mappo.py:364-370 (`running_add = running_add * gamma + r[t]`) for columns [T, n], with the episode
    boundary handling of MAPPO.interact (102-158): a rollout segment that ended its episode starts from 0."""
    rewards = np.asarray(rewards, dtype=np.float64)
    T, n = rewards.shape
    out = np.zeros_like(rewards)
    run = np.zeros(n) if final_value is None else np.asarray(final_value, dtype=np.float64).copy()
    for t in range(T - 1, -1, -1):
        d = np.repeat(np.asarray(dones[t]) != 0, cols_per_env)
        run = np.where(d, 0.0, run)
        run = run * gamma + rewards[t]
        out[t] = run
    return out


def actor_log_probs(w, obs):
    """marl/single_agent/Model_common.py:5-23 with output_act = log_softmax: w = dict(fc1_weight, fc1_bias, ...)."""
    x = np.asarray(obs, dtype=np.float64)
    h = np.maximum(x @ w["fc1_weight"].T.astype(np.float64) + w["fc1_bias"], 0.0)
    h = np.maximum(h @ w["fc2_weight"].T.astype(np.float64) + w["fc2_bias"], 0.0)
    z = h @ w["fc3_weight"].T.astype(np.float64) + w["fc3_bias"]
    z = z - z.max(axis=1, keepdims=True)
    return z - np.log(np.exp(z).sum(axis=1, keepdims=True))
