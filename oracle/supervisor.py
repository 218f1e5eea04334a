"""CPU restatement of the reference's baseline supervisors `safety_supervisor` (priority) and `safety_layer_dmc` —
TEST INFRASTRUCTURE ONLY.

Follows (file:line in /root/reference):
  highway_env/vehicle/safety/central_layer.py:16-178      safety_supervisor (is_priority=True)
  highway_env/vehicle/safety/decentralised_dmc.py:16-67   _evaluate_vehicle_action
  highway_env/vehicle/safety/decentralised_dmc.py:70-198  safety_layer_dmc
  highway_env/envs/common/mdp_controller.py:19-126        mdp_controller (look-ahead step of a CAV)
  highway_env/envs/common/idm_controller.py:42-304        generate_actions / idm_controller (look-ahead step of an HDV)
  highway_env/envs/common/abstract.py:219-280, 614-618, 721-755   available actions, safety room, merging end, collision
  highway_env/road/road.py:67-109, 294-350                next_lane, surrounding_vehicles
  highway_env/road/lane.py:61-108, utils.py:31-121        lane predicates, not_zero, wrap_to_pi, rectangle tests

The supervisor runs once per policy step on env merge-multi-agent-v0, before `_simulate` (abstract.py:459-467), and only
replaces entries of the meta-action tuple.  It is a pure function of (scene, action tuple, the np.random.rand() draws it
consumes), which is how it is written here: small pure-Python loops over <= 11 vehicles x 18 look-ahead points.  Pinned
against the fixtures tests/golden/priority_v0_td3_mixed.npz and dmc_v0_td3_mixed.npz (oracle/refharness/gen_golden.py), see tests/test_oracle_golden.py.
No product code imports this module; the CUDA supervisors it is meant to check are not built yet.
"""
import math

L_AB0, L_BC0, L_BC1, L_CD0, L_JK0, L_KB0 = range(6)
LANE_SX = (0.0, 320.0, 320.0, 420.0, 0.0, 220.0)
LANE_SY = (0.0, 0.0, 4.0, 0.0, 10.5, 7.25)
LANE_LEN = (320.0, 100.0, 100.0, 1000.0, 220.0, 100.0)
FORBIDDEN = (False, False, True, False, True, True)
MAIN = (L_AB0, L_BC0, L_CD0)
AMPLITUDE, PULSATION, PHASE = 3.25, 2 * math.pi / (2 * 100.0), math.pi / 2
VLEN, VWID, LANE_WIDTH = 5.0, 2.0, 4.0
OBSTACLE = (420.0, 4.0)
N_POINTS = 18                     # simulation_frequency // policy_frequency * n_step = 3 * 6
DT = 1.0 / 15
KP_A = 1 / 0.6
KP_HEADING = 1 / 0.2
KP_LATERAL = 1 / 3 * KP_HEADING
PURSUIT_TAU = 0.5 * 0.2
MAX_STEER = math.pi / 3
IDLE, LANE_LEFT, LANE_RIGHT, FASTER, SLOWER = 1, 0, 2, 3, 4
KIND_CAV, KIND_HDV = (1, 3), (2, 4)          # kinds of the v1 / v0 fixtures


def not_zero(x, eps=1e-2):
    if abs(x) > eps:
        return x
    return eps if x > 0 else -eps


def wrap_to_pi(x):
    return ((x + math.pi) % (2 * math.pi)) - math.pi


def clip(x, lo, hi):
    return lo if x < lo else (hi if x > hi else x)


def lane_coords(lane, x, y):
    s = x - LANE_SX[lane]
    r = y - LANE_SY[lane]
    if lane == L_KB0:
        r -= AMPLITUDE * math.sin(PULSATION * s + PHASE)
    return s, r


def lane_heading(lane, s):
    if lane == L_KB0:
        return math.atan(AMPLITUDE * PULSATION * math.cos(PULSATION * s + PHASE))
    return 0.0


def lane_distance(lane, x, y):
    s, r = lane_coords(lane, x, y)
    return abs(r) + max(s - LANE_LEN[lane], 0) + max(0 - s, 0)


def after_end(lane, x, y):
    return lane_coords(lane, x, y)[0] > LANE_LEN[lane] - VLEN / 2


def is_reachable_from(lane, x, y):
    if FORBIDDEN[lane]:
        return False
    s, r = lane_coords(lane, x, y)
    return abs(r) <= 2 * LANE_WIDTH and 0 <= s < LANE_LEN[lane] + VLEN


def on_lane(lane, x, y, margin):
    s, r = lane_coords(lane, x, y)
    return abs(r) <= LANE_WIDTH / 2 + margin and -VLEN <= s < LANE_LEN[lane] + VLEN


def next_lane(lane, x, y):
    if lane in (L_AB0, L_KB0):         # 1 lane -> 2 lanes: the closer of bc0 / bc1, the first on a tie
        return L_BC0 if lane_distance(L_BC0, x, y) <= lane_distance(L_BC1, x, y) else L_BC1
    if lane in (L_BC0, L_BC1):
        return L_CD0
    if lane == L_JK0:
        return L_KB0
    return lane                        # cd0: end of the graph


def side_lanes(lane):
    return {L_BC0: (L_BC1,), L_BC1: (L_BC0,)}.get(lane, ())


class Veh(object):
    __slots__ = ("cav", "x", "y", "heading", "speed", "target_speed", "lane", "target_lane", "speed_index", "crashed",
                 "steer", "acc", "traj")

    def copy(self):
        v = Veh()
        for k in self.__slots__:
            setattr(v, k, getattr(self, k))
        v.traj = list(self.traj)
        return v


def vehicles_of(state, e):
    out = []
    for i in range(int(state["n_veh"][e])):
        v = Veh()
        v.cav = int(state["kind"][e, i]) in KIND_CAV
        v.x, v.y = float(state["x"][e, i]), float(state["y"][e, i])
        v.heading, v.speed = float(state["heading"][e, i]), float(state["speed"][e, i])
        v.target_speed = float(state["target_speed"][e, i])
        v.lane, v.target_lane = int(state["lane"][e, i]), int(state["target_lane"][e, i])
        v.speed_index = int(state["speed_index"][e, i])
        v.crashed = bool(state["crashed"][e, i])
        v.steer = v.acc = 0.0
        v.traj = []
        out.append(v)
    return out


def steering_control(target_lane, v):
    s, r = lane_coords(target_lane, v.x, v.y)
    future_heading = lane_heading(target_lane, s + v.speed * PURSUIT_TAU)
    lateral_speed_command = -KP_LATERAL * r
    heading_command = math.asin(clip(lateral_speed_command / not_zero(v.speed), -1, 1))
    heading_ref = future_heading + clip(heading_command, -math.pi / 4, math.pi / 4)
    heading_rate_command = KP_HEADING * wrap_to_pi(heading_ref - v.heading)
    steering = math.asin(clip(VLEN / 2 / not_zero(v.speed) * heading_rate_command, -1, 1))
    return clip(steering, -MAX_STEER, MAX_STEER)


def follow_road(v):
    if after_end(v.target_lane, v.x, v.y):
        v.target_lane = next_lane(v.target_lane, v.x, v.y)


def clip_actions(steer, acc, speed, crashed):
    if crashed:
        steer, acc = 0.0, -1.0 * speed
    if speed > 40:
        acc = min(acc, 1.0 * (40 - speed))
    elif speed < -40:
        acc = max(acc, 1.0 * (40 - speed))
    return steer, acc


def bicycle(v, steer, acc):
    beta = math.atan(1 / 2 * math.tan(steer))
    vx, vy = v.speed * math.cos(v.heading + beta), v.speed * math.sin(v.heading + beta)
    v.x += vx * DT
    v.y += vy * DT
    v.heading += v.speed * math.sin(beta) / (VLEN / 2) * DT
    v.speed += acc * DT
    v.traj.append((v.x, v.y, v.heading, v.speed))


def mdp_controller(v, action):
    """One look-ahead sub-step of a CAV; the meta-action is re-applied on EVERY call (FASTER keeps adding 5 m/s to the
    target speed), `lane` is never updated, the speed is not floored at 0 (mdp_controller.py:19-66)."""
    follow_road(v)
    if action == FASTER:
        v.target_speed += 5
    elif action == SLOWER:
        v.target_speed -= 5
    elif action in (LANE_RIGHT, LANE_LEFT):
        road_lanes = {L_BC0: (L_BC0, L_BC1), L_BC1: (L_BC0, L_BC1)}.get(v.target_lane, (v.target_lane,))
        idx = road_lanes.index(v.target_lane) + (1 if action == LANE_RIGHT else -1)
        cand = road_lanes[int(clip(idx, 0, len(road_lanes) - 1))]
        if is_reachable_from(cand, v.x, v.y):
            v.target_lane = cand
    steer = clip(steering_control(v.target_lane, v), -MAX_STEER, MAX_STEER)
    acc = KP_A * (v.target_speed - v.speed)
    v.steer, v.acc = steer, acc
    steer, acc = clip_actions(steer, acc, v.speed, v.crashed)
    bicycle(v, steer, acc)


def neighbour_vehicles(v, vehicles):
    """idm_controller.py:243-271 on the vehicle's own lane; candidates = vehicles, then the obstacle."""
    lane = v.lane
    s = lane_coords(lane, v.x, v.y)[0]
    s_front = s_rear = None
    front = rear = None
    for o in list(vehicles) + [None]:
        if o is v:
            continue
        ox, oy = (OBSTACLE if o is None else (o.x, o.y))
        if not on_lane(lane, ox, oy, 1):
            continue
        s_v = lane_coords(lane, ox, oy)[0]
        if s <= s_v and (s_front is None or s_v <= s_front):
            s_front, front = s_v, (o if o is not None else "obstacle")
        if s_v < s and (s_rear is None or s_v > s_rear):
            s_rear, rear = s_v, (o if o is not None else "obstacle")
    return front, rear


def idm_acceleration(ego, front):
    """idm_controller.py:198-226 (ego is never a road object here)."""
    target = not_zero(ego.target_speed)
    acc = 3.0 * (1 - math.pow(max(ego.speed, 0) / target, 4.0))
    if front is not None:
        fx, fy, fh, fs = (OBSTACLE[0], OBSTACLE[1], 0.0, 0.0) if front == "obstacle" else (front.x, front.y, front.heading,
                                                                                         front.speed)
        d = lane_coords(ego.lane, fx, fy)[0] - lane_coords(ego.lane, ego.x, ego.y)[0]
        ch, sh = math.cos(ego.heading), math.sin(ego.heading)
        dv = (ego.speed * ch - fs * math.cos(fh)) * ch + (ego.speed * sh - fs * math.sin(fh)) * sh
        d_star = 10.0 + ego.speed * 1.5 + ego.speed * dv / (2 * math.sqrt(15.0))
        acc -= 3.0 * math.pow(d_star / not_zero(d), 2)
    return acc


def generate_actions(v, vehicles, draw_steer, draw_acc):
    """idm_controller.py:60-79.  `mobil` compares the ego's IDM acceleration behind the SAME preceding vehicle twice
    (both neighbour queries are on the own lane), so its gain is 0 < 0.1 and no lane change is ever started; an ongoing
    one steers back to the current lane (change_lane_policy returns `lane_index`)."""
    front, _ = neighbour_vehicles(v, vehicles)
    follow_road(v)
    steer_lane = v.target_lane if v.lane == v.target_lane else v.lane
    steer = clip(steering_control(steer_lane, v) * (draw_steer * 0.1 + 0.95), -MAX_STEER, MAX_STEER)
    acc = clip(idm_acceleration(v, front) * (draw_acc * 0.1 + 0.95), -6.0, 6.0)
    v.steer, v.acc = steer, acc


def idm_controller(v):
    if v.crashed:
        v.traj.append((v.x, v.y, v.heading, v.speed))
        return
    v.steer, v.acc = clip_actions(v.steer, v.acc, v.speed, v.crashed)
    bicycle(v, v.steer, v.acc)


def surrounding_vehicles(v, vehicles, lane=None):
    """road.py:294-350: front / rear by world x inside the hard-coded lane groups."""
    lane = v.lane if lane is None else lane
    groups = {L_AB0: (L_AB0, L_BC0), L_BC0: (L_AB0, L_BC0, L_CD0), L_CD0: (L_BC0, L_CD0),
              L_JK0: (L_JK0, L_KB0), L_KB0: (L_JK0, L_KB0, L_BC1), L_BC1: (L_KB0, L_BC1)}[lane]
    s = v.x
    s_front = s_rear = None
    front = rear = None
    for o in vehicles:
        if o is v or o.lane not in groups:
            continue
        s_v = o.x
        if s <= s_v and (s_front is None or s_v <= s_front):
            s_front, front = s_v, o
        if s_v < s and (s_rear is None or s_v > s_rear):
            s_rear, rear = s_v, o
    return front, rear


def point_in_rotated_rectangle(px, py, cx, cy, length, width, angle):
    c, s = math.cos(angle), math.sin(angle)
    dx, dy = px - cx, py - cy
    rx, ry = c * dx - s * dy, s * dx + c * dy
    return -length / 2 <= rx <= length / 2 and -width / 2 <= ry <= width / 2


def has_corner_inside(r1, r2):
    (c1x, c1y, l1, w1, a1), (c2x, c2y, l2, w2, a2) = r1, r2
    c, s = math.cos(a1), math.sin(a1)
    for px, py in ((0, 0), (-l1 / 2, 0), (l1 / 2, 0), (0, -w1 / 2), (0, w1 / 2), (-l1 / 2, -w1 / 2), (-l1 / 2, w1 / 2),
                   (l1 / 2, -w1 / 2), (l1 / 2, w1 / 2)):
        if point_in_rotated_rectangle(c1x + c * px - s * py, c1y + s * px + c * py, c2x, c2y, l2, w2, a2):
            return True
    return False


def is_colliding(v, ox, oy, oh, olen, owid):
    if math.hypot(ox - v.x, oy - v.y) > VLEN:
        return False
    r1 = (v.x, v.y, 0.9 * VLEN, 0.9 * VWID, v.heading)
    r2 = (ox, oy, 0.9 * olen, 0.9 * owid, oh)
    return has_corner_inside(r1, r2) or has_corner_inside(r2, r1)


def available_actions(v):
    """abstract.py:219-240 on the ORIGINAL vehicle."""
    acts = [IDLE]
    for l in side_lanes(v.lane):
        if l < v.lane and is_reachable_from(l, v.x, v.y):
            acts.append(LANE_LEFT)
        if l > v.lane and is_reachable_from(l, v.x, v.y):
            acts.append(LANE_RIGHT)
    if v.speed_index < 4:
        acts.append(FASTER)
    if v.speed_index > 0:
        acts.append(SLOWER)
    return acts


def check_safety_room(v, action, surrounding, time_steps):
    """abstract.py:242-280.  `v` keeps being stepped across calls; only its first trajectory points are read."""
    rooms = []
    for t in range(time_steps + 1):
        mdp_controller(v, action)
        room = (420.0 - v.x) if v.lane == L_BC1 else 100.0
        if action in (LANE_LEFT, LANE_RIGHT):
            for o in surrounding:
                if o is not None and abs(o.traj[t][0] - v.traj[t][0]) <= room:
                    room = abs(o.traj[t][0] - v.traj[t][0])
        else:
            o = surrounding[0] if v.lane in MAIN else surrounding[2]
            if o is not None and (o.traj[t][0] - v.traj[t][0]) <= room:
                room = o.traj[t][0] - v.traj[t][0]
        rooms.append(room)
    return min(rooms)


def headway_distance(v, vehicles):
    """abstract.py:620-635."""
    headway = 60
    nxt = next_lane(v.lane, v.x, v.y)
    for o in vehicles:
        if o.lane == v.lane and o.x > v.x:
            headway = min(headway, o.x - v.x)
        if v.lane != L_BC1 and o.lane == nxt and o.x > v.x:
            headway = min(headway, o.x - v.x)
    return headway


def priority_order(vehicles, cavs, draws, headway_time):
    """central_layer.py:33-63 / decentralised_dmc.py:88-117: smaller number first (PriorityQueue)."""
    keyed = []
    for n, i in enumerate(cavs):
        v = vehicles[i]
        p = 0.0
        if v.lane == L_BC1:
            p = -0.5
            p -= (100.0 - (420.0 - v.x)) / 100.0
        p += 0.5 * math.log(headway_distance(v, vehicles) / (headway_time * v.speed)) if v.speed > 0 else 0
        p += float(draws[n]) * 0.001
        keyed.append((p, n, i))
    return [i for _, _, i in sorted(keyed)]


def dmc_supervisor(state, e, actions, draws, headway_time=1.2):
    """-> the supervised action list of env `e` of an env-major state dict (decentralised_dmc.py:70-198).
    `draws`: the np.random.rand() values in the order the reference consumes them (one per CAV for the priority
    numbers, then two per HDV in list order)."""
    original = vehicles_of(state, e)
    copies = [v.copy() for v in original]
    cavs = [i for i, v in enumerate(copies) if v.cav]
    actions = [int(a) for a in actions[:len(cavs)]]
    k = len(cavs)                                  # priority draws
    order = priority_order(copies, cavs, draws, headway_time)
    for i in order:                                # step 1: every CAV 18 points ahead with its own action
        for _ in range(N_POINTS):
            mdp_controller(copies[i], actions[i])
    for v in copies:                               # step 1a: every HDV, one IDM decision then open loop
        if v.cav:
            continue
        generate_actions(v, copies, float(draws[k]), float(draws[k + 1]))
        k += 2
        for _ in range(N_POINTS):
            idm_controller(v)
    result = list(actions)
    for i in order:                                # step 2, in priority order: a CAV that crashed in its look-ahead
        # is left standing at the crash point for the neighbour queries of the CAVs after it
        v = copies[i]
        if v.lane in MAIN:
            v_fl, v_rl = surrounding_vehicles(v, copies)
            if side_lanes(v.lane):
                v_fr, v_rr = surrounding_vehicles(v, copies, side_lanes(v.lane)[0])
            elif v.lane == L_AB0 and v.x > 220.0:
                v_fr, v_rr = surrounding_vehicles(v, copies, L_KB0)
            else:
                v_fr, v_rr = None, None
        else:
            v_fr, v_rr = surrounding_vehicles(v, copies)
            if side_lanes(v.lane):
                v_fl, v_rl = surrounding_vehicles(v, copies, side_lanes(v.lane)[0])
            elif v.lane == L_KB0:
                v_fl, v_rl = surrounding_vehicles(v, copies, L_AB0)
            else:
                v_fl, v_rl = None, None
        neighbours = [v_fl, v_rr, v_fr, v_rl]
        acts = available_actions(original[i])
        v.crashed = False
        for t in range(N_POINTS):
            v.x, v.y, v.heading = v.traj[t][0], v.traj[t][1], v.traj[t][2]
            for o in neighbours:
                if o is None or v.crashed or o is v:
                    continue
                ox, oy, oh, osp = o.traj[t]
                if is_colliding(v, ox, oy, oh, VLEN, VWID):
                    v.speed = min((v.speed, osp), key=abs)
                    v.crashed = o.crashed = True
            if not v.crashed and is_colliding(v, OBSTACLE[0], OBSTACLE[1], 0.0, 2.0, 2.0):
                v.speed = min((v.speed, 0), key=abs)
                v.crashed = True
            if v.crashed:
                rooms = []
                for a in acts:
                    c = original[i].copy()
                    rooms.append(sum(check_safety_room(c, a, neighbours, tt) for tt in range(N_POINTS)))
                result[i] = acts[rooms.index(max(rooms))]
                break
    return result


def _neighbour_sets(v, road):
    """central_layer.py:84-110: (v_fl, v_rl, v_fr, v_rr) of an ego on the (partly propagated) copy of the scene."""
    if v.lane in MAIN:
        v_fl, v_rl = surrounding_vehicles(v, road)
        if side_lanes(v.lane):
            v_fr, v_rr = surrounding_vehicles(v, road, side_lanes(v.lane)[0])
        elif v.lane == L_AB0 and v.x > 220.0:
            v_fr, v_rr = surrounding_vehicles(v, road, L_KB0)
        else:
            v_fr, v_rr = None, None
    else:
        v_fr, v_rr = surrounding_vehicles(v, road)
        if side_lanes(v.lane):
            v_fl, v_rl = surrounding_vehicles(v, road, side_lanes(v.lane)[0])
        elif v.lane == L_KB0:
            v_fl, v_rl = surrounding_vehicles(v, road, L_AB0)
        else:
            v_fl, v_rl = None, None
    return v_fl, v_rl, v_fr, v_rr


def priority_supervisor(state, e, actions, draws, headway_time=1.2):
    """-> the supervised action list of env `e` (central_layer.py:16-178, `safety_supervisor` with is_priority=True).
    CAVs are handled one after the other in priority order; each rolls itself and its four surrounding vehicles 18
    points ahead ON THE SHARED COPY of the scene (vehicles an earlier ego already propagated are not moved again, an
    ego that was propagated as somebody's neighbour restarts from its real state), and at the first predicted collision
    takes the available action with the largest safety room - for the rest of its own look-ahead as well.
    `draws`: np.random.rand() values in consumption order (one per CAV, then two per IDM decision as they occur)."""
    original = vehicles_of(state, e)
    road = [v.copy() for v in original]
    cavs = [i for i, v in enumerate(road) if v.cav]
    actions = [int(a) for a in actions[:len(cavs)]]
    k = len(cavs)
    order = priority_order(road, cavs, draws, headway_time)
    captured = {i: road[i] for i in cavs}          # the objects the queue holds
    for turn, i in enumerate(order):
        first_change = True
        if len(captured[i].traj) == N_POINTS:      # moved as a neighbour before: start again from the real state
            road[i] = original[i].copy()
        v = road[i]
        acts = available_actions(v)
        v_fl, v_rl, v_fr, v_rr = _neighbour_sets(v, road)
        for t in range(N_POINTS):
            for o in (v_fl, v_fr, v, v_rl, v_rr):
                if o is None:
                    continue
                if len(o.traj) == N_POINTS and turn != 0 and o is not v:
                    continue
                if not o.cav:
                    if t == 0:
                        generate_actions(o, road, float(draws[k]), float(draws[k + 1]))
                        k += 2
                    idm_controller(o)
                elif o is not v:
                    # `actions[v.id]` (central_layer.py:135) with v.id == 0 for every vehicle: ControlledVehicle sets
                    # id = 0 (controller.py:49) and the per-vehicle assignment in reset is commented out
                    # (abstract.py:196-198), so every neighbouring CAV is rolled forward with the FIRST CAV's action
                    mdp_controller(o, actions[0])
                else:
                    mdp_controller(o, actions[i])
            for o in (v_fl, v_rl, v_fr, v_rr):
                if o is None or v.crashed or o is v:
                    continue
                ox, oy, oh, osp = o.traj[t]
                if is_colliding(v, ox, oy, oh, VLEN, VWID):
                    v.speed = min((v.speed, osp), key=abs)
                    v.crashed = o.crashed = True
            if not v.crashed and is_colliding(v, OBSTACLE[0], OBSTACLE[1], 0.0, 2.0, 2.0):
                v.speed = min((v.speed, 0), key=abs)
                v.crashed = True
            if v.crashed:
                rooms, updated = [], []
                for a in acts:
                    c = original[i].copy()
                    rooms.append(check_safety_room(c, a, [v_fl, v_rl, v_fr, v_rr], t))
                    updated.append(c)
                best = rooms.index(max(rooms))
                v = road[i] = updated[best]
                if first_change:
                    first_change = False
                    actions[i] = acts[best]
                for o in (v_fl, v_rl, v_fr, v_rr):
                    if o is not None and o.crashed:
                        o.crashed = False
    return actions
