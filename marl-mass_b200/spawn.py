"""Host-side, seed-exact scene construction.

Mirrors the reference's reset path so that `reset(is_training=False, testing_seeds=s)` builds the SAME
scene the reference builds for seed `s`:
  AbstractEnv.reset           highway_env/envs/common/abstract.py:176-209  (np.random.seed(seed))
  MergeEnv._num_vehicles      highway_env/envs/merge_env_v1.py:180-211, 476-495
  MergeEnv._make_vehicles     highway_env/envs/merge_env_v1.py:265-364
  Vehicle/ControlledVehicle/MDPVehicle/IDMVehicle/MDPLCVehicle.__init__
                              vehicle/kinematics.py:36-53, controller.py:35-49,277-291, behavior.py:42-54,
                              safe_controller.py:27-61
The reference draws from the process-global legacy numpy generator in a fixed call order; a private
`np.random.RandomState(seed)` replays that stream exactly.  The result is an env-major state dict that
`MergeEnvBatched.set_state` uploads.  (The batched device-side spawn used for auto-reset follows the same law
with a counter-based generator instead — see csrc/merge_step.cu reset_kernel.)
"""
import numpy as np

from ._lib import ENV_FIELDS, F64_FIELDS, I32_FIELDS, MAXV

KIND_CAV, KIND_HDV = 1, 2
LANE_SX = (0.0, 320.0, 320.0, 420.0, 0.0, 220.0)
LANE_SY = (0.0, 0.0, 4.0, 0.0, 10.5, 7.25)
LANE_LEN = (320.0, 100.0, 100.0, 1000.0, 220.0, 100.0)
L_AB0, L_BC0, L_BC1, L_CD0, L_JK0, L_KB0 = range(6)


def _wrap_to_pi(x):
    return ((x + np.pi) % (2 * np.pi)) - np.pi


def closest_lane(x, y, heading=0.0):
    """RoadNetwork.get_closest_lane_index (road.py:51-65) on the fixed merge network."""
    best, bd = 0, None
    for l in range(6):
        s = x - LANE_SX[l]
        r = y - LANE_SY[l]
        h = 0.0
        if l == L_KB0:
            puls = 2 * np.pi / (2 * 100)
            r = r - 3.25 * np.sin(puls * s + np.pi / 2)
            h = np.arctan(3.25 * puls * np.cos(puls * s + np.pi / 2))
        d = abs(r) + max(s - LANE_LEN[l], 0) + max(0 - s, 0) + abs(_wrap_to_pi(heading - h))
        if bd is None or d < bd:
            best, bd = l, d
    return best


def num_vehicles(rs, traffic_density, traffic_type, num_CAV=0):
    lo_c, lo_h = {1: (1, 1), 2: (2, 2), 3: (4, 3)}[int(traffic_density)]
    if num_CAV == 0:
        num_CAV = rs.choice(np.arange(lo_c, lo_c + 3), 1)[0]
    num_HDV = rs.choice(np.arange(lo_h, lo_h + 3), 1)[0]
    if traffic_type == "cav":
        num_CAV, num_HDV = num_CAV + num_HDV, 0
    elif traffic_type == "av":          # merge_env_v1.py:485-489: one CAV, everybody else human-driven
        num_CAV, num_HDV = 1, num_CAV + num_HDV - 1
    elif traffic_type == "hdv":         # merge_env_v1.py:490-494: nobody is controlled
        num_CAV, num_HDV = 0, num_CAV + num_HDV
    elif traffic_type != "mixed":
        raise ValueError("traffic_type %r is not supported (cav | mixed | av | hdv)" % (traffic_type,))
    return int(num_CAV), int(num_HDV)


def spawn_scene(seed, traffic_density=1, traffic_type="cav", num_CAV=0, rng_out=None):
    """One scene: list of (kind, x, y, speed) in road.vehicles order plus n_merge.  rng_out (a list): receives the
    generator as the spawn left it - the reference keeps drawing from the same process-global stream afterwards (the
    baseline supervisors' np.random.rand() calls, central_layer.py:55 / idm_controller.py:60-79), and nothing else on
    the step path consumes it, so the adapter continues this replay step by step."""
    rs = np.random.RandomState(int(seed))
    if rng_out is not None:
        rng_out.append(rs)
    num_CAV, num_HDV = num_vehicles(rs, traffic_density, traffic_type, num_CAV)
    spawn_points_s = [10, 60, 110, 160, 210, 260]
    spawn_points_m = [5, 55, 105, 155, 205, 255]
    num_s_c = num_CAV // 2 if num_CAV != 1 else rs.choice(2)
    num_m_c = num_CAV - num_s_c
    sp_s_c = list(rs.choice(spawn_points_s, num_s_c, replace=False))
    sp_m_c = list(rs.choice(spawn_points_m, num_m_c, replace=False))
    for a in sp_s_c:
        spawn_points_s.remove(a)
    for b in sp_m_c:
        spawn_points_m.remove(b)
    num_s_h = num_HDV // 2 if num_HDV != 1 else rs.choice(2)
    num_m_h = num_HDV - num_s_h
    sp_s_h = list(rs.choice(spawn_points_s, num_s_h, replace=False))
    sp_m_h = list(rs.choice(spawn_points_m, num_m_h, replace=False))
    initial_speed = list(rs.rand(num_CAV + num_HDV) * 2 + 25)
    loc_noise = list(rs.rand(num_CAV + num_HDV) * 8 - 4)
    vehicles = []
    for pts, kind, y in ((sp_s_c, KIND_CAV, 0.0), (sp_m_c, KIND_CAV, 10.5), (sp_s_h, KIND_HDV, 0.0),
                         (sp_m_h, KIND_HDV, 10.5)):
        for pt in pts:
            x = float(pt + loc_noise.pop(0))
            vehicles.append((kind, x, y, float(initial_speed.pop(0))))
    return vehicles, int(num_m_c)


def empty_state(n_envs):
    st = {k: np.zeros((n_envs, MAXV), np.float64) for k in F64_FIELDS}
    st.update({k: np.zeros((n_envs, MAXV), np.int32) for k in I32_FIELDS})
    st.update({k: np.zeros(n_envs, np.int32) for k in ENV_FIELDS})
    st["hl_action"][:] = -1
    return st


def fill_scene(st, e, vehicles, n_merge):
    """Write one spawned scene into row `e` of an env-major state dict (constructor semantics)."""
    for k in F64_FIELDS + I32_FIELDS:
        st[k][e, :] = 0
    st["hl_action"][e, :] = -1
    n_cav = 0
    for i, (kind, x, y, speed) in enumerate(vehicles):
        st["kind"][e, i] = kind
        st["x"][e, i], st["y"][e, i], st["speed"][e, i] = x, y, speed
        lane = closest_lane(x, y, 0.0)
        st["lane"][e, i] = st["target_lane"][e, i] = lane
        if kind == KIND_CAV:
            n_cav += 1
            xi = (speed - 10.0) / (30.0 - 10.0)
            idx = int(np.clip(np.round(xi * 4), 0, 4))
            st["speed_index"][e, i] = idx
            st["target_speed"][e, i] = 10.0 + idx * (30.0 - 10.0) / 4
            st["min_headway"][e, i] = 180.0 / 40.0
        else:
            st["speed_index"][e, i] = -1
            st["target_speed"][e, i] = speed
            st["timer"][e, i] = (np.sum(np.array([x, y])) * np.pi) % 1.0
    st["n_veh"][e] = len(vehicles)
    st["n_cav"][e] = n_cav
    st["n_merge"][e] = n_merge
    st["steps"][e] = 0
    st["time"][e] = 0


def spawn_state(seeds, traffic_density=1, traffic_type="cav", num_CAV=0, rngs=None):
    """Env-major state dict holding the reference's scene for each seed (num_CAV: one value or one per seed).
    rngs (a list): receives one generator per scene, positioned after the spawn draws (see spawn_scene)."""
    seeds = list(seeds)
    n_cav = list(num_CAV) if hasattr(num_CAV, "__len__") else [num_CAV] * len(seeds)
    assert len(n_cav) == len(seeds)
    st = empty_state(len(seeds))
    for e, s in enumerate(seeds):
        vehicles, n_merge = spawn_scene(s, traffic_density, traffic_type, int(n_cav[e]), rng_out=rngs)
        fill_scene(st, e, vehicles, n_merge)
    return st
