"""Shared comparison helpers for the CPU (oracle vs golden) and GPU (CUDA vs oracle / golden) parity tests."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

F64_FIELDS = ("x", "y", "heading", "speed", "target_speed", "gvx", "rec1_x", "rec1_vx", "rec2_x", "rec2_vx",
              "act_steer", "act_acc", "safe_steer", "safe_acc", "timer", "min_headway", "steering_angle")
I32_FIELDS = ("kind", "lane", "target_lane", "speed_index", "crashed", "hl_action", "hist_len", "fg_set",
              "is_collaborating", "is_lc_safe")   # collaborate_adj is dead state
ENV_FIELDS = ("n_veh", "n_cav", "n_merge", "steps", "time")
SH_I = ("ran", "leader", "front_adj", "rear_adj", "constrain_adj", "active", "is_lc_safe")
SH_F = ("safe_acc", "safe_steer", "nom_acc", "nom_steer")
OUT_F = ("obs", "reward", "agents_rewards", "regional_rewards", "average_speed", "traffic_speed", "min_headway",
         "merge_percent")
OUT_I = ("done", "agents_dones")

# Lane-change veto boundary (SURVEY.md §7 "ill-posed discrete output"): when the adjacent barrier row is the
# active QP constraint the veto test evaluates a quantity that is 0 up to rounding; such solves are classified
# by |margin| and excluded from the bit-exact set.
LC_BOUNDARY_EPS = 1e-9


def load_golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    cfg = json.loads(str(g["config"]))
    return g, cfg


def golden_state(g, rows, all_fields):
    rows = np.atleast_1d(rows)
    st = {}
    for k in all_fields[0]:
        st[k] = np.ascontiguousarray(g["st_" + k][rows], np.float64)
    for k in all_fields[1]:
        st[k] = np.ascontiguousarray(g["st_" + k][rows], np.int32)
    for k in all_fields[2]:
        st[k] = np.ascontiguousarray(g["st_" + k][rows], np.int32)
    return st


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))


def used_mask(st):
    """[E, MAXV] mask of live vehicle slots."""
    n = st["n_veh"][:, None]
    return np.arange(st["x"].shape[1])[None, :] < n


# the v0 env's vehicles (MDPVehicle / IDMVehicle) have no history, shield or flag attributes to compare
V0_F64 = ("x", "y", "heading", "speed", "target_speed", "act_steer", "act_acc", "timer")
V0_I32 = ("kind", "lane", "target_lane", "speed_index", "crashed")


def obs25(obs):
    """Kinematics (v0) view of a KinematicLC observation: drop the heading column of each of the 5 rows."""
    obs = np.asarray(obs)
    return obs.reshape(obs.shape[:-1] + (5, 6))[..., :5].reshape(obs.shape[:-1] + (25,))


def compare_states(got, want, tol, what="", v0=False):
    """Discrete fields exact, continuous within tol (relative, floor 1); only live slots are compared."""
    m = used_mask(want)
    for k in ENV_FIELDS:
        assert np.array_equal(got[k], want[k]), "%s env field %s differs" % (what, k)
    for k in (V0_I32 if v0 else I32_FIELDS):
        bad = np.argwhere((got[k] != want[k]) & m)
        assert len(bad) == 0, "%s discrete field %s differs at %s" % (what, k, bad[:5].tolist())
    worst = 0.0
    for k in (V0_F64 if v0 else F64_FIELDS):
        if k == "rec1_x":
            continue
        err = rel_err(got[k], want[k]) * m
        worst = max(worst, float(err.max()))
        assert err.max() <= tol, "%s field %s rel err %.3e > %.1e" % (what, k, err.max(), tol)
    return worst


def near_tie_inside_step(orc, ocfg, pre, e, a):
    """True if, inside this policy step of env e, two vehicles end a sub-step within 4 ulp of each other in x.  The
    oracle (glibc) and the kernels (libdevice) agree on every +, *, / and sqrt bit for bit, but their sin / tan / atan /
    asin may differ in the last place; a vehicle that steers by 1e-7 rad can therefore end a sub-step one ulp apart on
    the two sides, and when another vehicle sits exactly there (the snapped grids make x + v dt coincide: 373.5 +
    17.5 / 15 == 374 + 10 / 15) the NEXT sub-step's front / rear classification sees a tie on one side and an order on
    the other.  Such a step is reported and its env dropped from the rest of the episode; anything else still fails."""
    one = {k: np.array(v[e:e + 1], copy=True) for k, v in pre.items()}
    full = ocfg.substeps
    try:
        for k in range(1, full):
            ocfg.substeps = k
            s1 = {kk: vv.copy() for kk, vv in one.items()}
            orc.step(ocfg, s1, a[e:e + 1], n_threads=1)
            n = int(s1["n_veh"][0])
            x = np.sort(s1["x"][0][:n])
            if n > 1 and (np.diff(x) <= 4 * np.spacing(np.abs(x[1:]))).any():
                return True
    finally:
        ocfg.substeps = full
    return False
