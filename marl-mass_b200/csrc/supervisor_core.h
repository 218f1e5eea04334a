// supervisor_core.h — the baseline supervisors `priority` / `dmc` as host+device code (mm_supervise; called by mm_step when
// mm_config.supervisor is set).
//
// What it is: the look-ahead action filters the reference runs once per policy step on env merge-multi-agent-v0
// (highway_env/vehicle/safety/central_layer.py:16-178 `safety_supervisor`, decentralised_dmc.py:16-198
// `safety_layer_dmc`, with envs/common/mdp_controller.py, idm_controller.py and abstract.py:219-280, 614-635, 721-755),
// written as plain functions over one scene so that the SAME source compiles for the host (g++, the CPU parity test
// tests/test_host_cpu.py::test_supervisor_core_*) and for sm_100a (MM_HD = __host__ __device__ under nvcc).
// The logic is pinned on the CPU against the reference fixtures (every step of priority_v0_td3_mixed / dmc_v0_td3_mixed,
// through the host build of this header) and on a B200 through the kernel of supervisor.cu.
//
// Scene = up to 12 vehicles (CAVs first).  A supervisor is a pure function
//     (scene, meta-action tuple, np.random.rand() draws in consumption order) -> supervised tuple.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define MM_HD __host__ __device__
#define MM_HD_CALL __host__ __device__ __noinline__   // big helpers stay calls on the device: inlined everywhere, ptxas
                                                      // needs ~9 minutes for the two supervisors
#else
#define MM_HD
#define MM_HD_CALL
#endif

namespace mmsup {

constexpr int MAXV = 12, NPTS = 18;   // n_points = simulation_frequency // policy_frequency * n_step = 3 * 6
constexpr int L_AB0 = 0, L_BC0 = 1, L_BC1 = 2, L_CD0 = 3, L_JK0 = 4, L_KB0 = 5;
constexpr int A_LANE_LEFT = 0, A_IDLE = 1, A_LANE_RIGHT = 2, A_FASTER = 3, A_SLOWER = 4;
constexpr double PI = 3.141592653589793;
constexpr double VLEN = 5.0, VWID = 2.0, LANE_WIDTH = 4.0, OBST_X = 420.0, OBST_Y = 4.0;
constexpr double DT = 1.0 / 15;
constexpr double KP_A = 1 / 0.6, KP_HEADING = 1 / 0.2, KP_LATERAL = 1.0 / 3 * (1 / 0.2), PURSUIT_TAU = 0.5 * 0.2;
constexpr double MAX_STEER = PI / 3;

struct Veh {
    double x, y, heading, speed, target_speed, steer, acc;
    int lane, target_lane, speed_index;
    int cav, crashed;
    int n_traj;                                   // points appended so far; only the first NPTS are kept
    double tx[NPTS], ty[NPTS], th[NPTS], tv[NPTS];
};

MM_HD inline double lane_sx(int l) { return l == L_BC0 || l == L_BC1 ? 320.0 : l == L_CD0 ? 420.0 : l == L_KB0 ? 220.0 : 0.0; }
MM_HD inline double lane_sy(int l) { return l == L_BC1 ? 4.0 : l == L_JK0 ? 10.5 : l == L_KB0 ? 7.25 : 0.0; }
MM_HD inline double lane_len(int l) { return l == L_AB0 ? 320.0 : l == L_CD0 ? 1000.0 : l == L_JK0 ? 220.0 : 100.0; }
MM_HD inline bool lane_forbidden(int l) { return l == L_BC1 || l == L_JK0 || l == L_KB0; }
MM_HD inline bool lane_main(int l) { return l == L_AB0 || l == L_BC0 || l == L_CD0; }
MM_HD inline double not_zero(double x) { return fabs(x) > 1e-2 ? x : (x > 0 ? 1e-2 : -1e-2); }
MM_HD inline double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
MM_HD inline double wrap_to_pi(double x) {        // ((x + pi) % (2 pi)) - pi with Python's sign convention
    double m = fmod(x + PI, 2 * PI);
    if (m < 0) m += 2 * PI;
    return m - PI;
}
MM_HD inline double lane_s(int l, double x) { return x - lane_sx(l); }
MM_HD inline double lane_r(int l, double x, double y) {
    double r = y - lane_sy(l);
    if (l == L_KB0) r -= 3.25 * sin(PI / 100.0 * lane_s(l, x) + PI / 2);
    return r;
}
MM_HD inline double lane_heading(int l, double s) {
    return l == L_KB0 ? atan(3.25 * (PI / 100.0) * cos(PI / 100.0 * s + PI / 2)) : 0.0;
}
MM_HD inline double lane_distance(int l, double x, double y) {
    double s = lane_s(l, x), r = lane_r(l, x, y);
    return fabs(r) + fmax(s - lane_len(l), 0.0) + fmax(0.0 - s, 0.0);
}
MM_HD inline bool after_end(int l, double x) { return lane_s(l, x) > lane_len(l) - VLEN / 2; }
MM_HD inline bool is_reachable_from(int l, double x, double y) {
    if (lane_forbidden(l)) return false;
    double s = lane_s(l, x);
    return fabs(lane_r(l, x, y)) <= 2 * LANE_WIDTH && 0 <= s && s < lane_len(l) + VLEN;
}
MM_HD inline bool on_lane(int l, double x, double y, double margin) {
    double s = lane_s(l, x);
    return fabs(lane_r(l, x, y)) <= LANE_WIDTH / 2 + margin && -VLEN <= s && s < lane_len(l) + VLEN;
}
MM_HD inline int next_lane(int l, double x, double y) {
    if (l == L_AB0 || l == L_KB0) return lane_distance(L_BC0, x, y) <= lane_distance(L_BC1, x, y) ? L_BC0 : L_BC1;
    if (l == L_BC0 || l == L_BC1) return L_CD0;
    if (l == L_JK0) return L_KB0;
    return l;
}
MM_HD inline int side_lane(int l) { return l == L_BC0 ? L_BC1 : (l == L_BC1 ? L_BC0 : -1); }

MM_HD inline double steering_control(int target_lane, const Veh &v) {
    double s = lane_s(target_lane, v.x), r = lane_r(target_lane, v.x, v.y);
    double future_heading = lane_heading(target_lane, s + v.speed * PURSUIT_TAU);
    double heading_command = asin(clipd(-KP_LATERAL * r / not_zero(v.speed), -1, 1));
    double heading_ref = future_heading + clipd(heading_command, -PI / 4, PI / 4);
    double rate = KP_HEADING * wrap_to_pi(heading_ref - v.heading);
    double steering = asin(clipd(VLEN / 2 / not_zero(v.speed) * rate, -1, 1));
    return clipd(steering, -MAX_STEER, MAX_STEER);
}
MM_HD inline void follow_road(Veh &v) {
    if (after_end(v.target_lane, v.x)) v.target_lane = next_lane(v.target_lane, v.x, v.y);
}
MM_HD inline void clip_actions(double &steer, double &acc, double speed, bool crashed) {
    if (crashed) { steer = 0.0; acc = -1.0 * speed; }
    if (speed > 40) acc = fmin(acc, 1.0 * (40 - speed));
    else if (speed < -40) acc = fmax(acc, 1.0 * (40 - speed));
}
MM_HD inline void push_traj(Veh &v) {
    if (v.n_traj < NPTS) { v.tx[v.n_traj] = v.x; v.ty[v.n_traj] = v.y; v.th[v.n_traj] = v.heading; v.tv[v.n_traj] = v.speed; }
    ++v.n_traj;
}
MM_HD inline void bicycle(Veh &v, double steer, double acc) {
    double beta = atan(1.0 / 2 * tan(steer));
    double vx = v.speed * cos(v.heading + beta), vy = v.speed * sin(v.heading + beta);
    v.x += vx * DT;
    v.y += vy * DT;
    v.heading += v.speed * sin(beta) / (VLEN / 2) * DT;
    v.speed += acc * DT;
    push_traj(v);
}
// mdp_controller.py:19-66: the meta-action is re-applied on every look-ahead point, `lane` is never updated
MM_HD_CALL inline void mdp_controller(Veh &v, int action) {
    follow_road(v);
    if (action == A_FASTER) v.target_speed += 5;
    else if (action == A_SLOWER) v.target_speed -= 5;
    else if (action == A_LANE_RIGHT || action == A_LANE_LEFT) {
        int cand = v.target_lane;
        if (v.target_lane == L_BC0 && action == A_LANE_RIGHT) cand = L_BC1;
        if (v.target_lane == L_BC1 && action == A_LANE_LEFT) cand = L_BC0;
        if (is_reachable_from(cand, v.x, v.y)) v.target_lane = cand;
    }
    double steer = clipd(steering_control(v.target_lane, v), -MAX_STEER, MAX_STEER);
    double acc = KP_A * (v.target_speed - v.speed);
    v.steer = steer; v.acc = acc;
    clip_actions(steer, acc, v.speed, v.crashed != 0);
    bicycle(v, steer, acc);
}

constexpr int OBSTACLE_ID = -2;
// idm_controller.py:243-271: front vehicle on the ego's own lane; candidates = vehicles in list order, then the obstacle
MM_HD inline int front_on_own_lane(const Veh *road, int n, int self) {
    const Veh &v = road[self];
    double s = lane_s(v.lane, v.x), s_front = 0;
    int front = -1;
    for (int j = 0; j <= n; ++j) {
        if (j == self) continue;
        double ox = j == n ? OBST_X : road[j].x, oy = j == n ? OBST_Y : road[j].y;
        if (!on_lane(v.lane, ox, oy, 1)) continue;
        double s_v = lane_s(v.lane, ox);
        if (s <= s_v && (front == -1 || s_v <= s_front)) { s_front = s_v; front = j == n ? OBSTACLE_ID : j; }
    }
    return front;
}
MM_HD inline double idm_acceleration(const Veh *road, int self, int front) {
    const Veh &e = road[self];
    double acc = 3.0 * (1 - pow(fmax(e.speed, 0.0) / not_zero(e.target_speed), 4.0));
    if (front != -1) {
        double fx = OBST_X, fh = 0.0, fs = 0.0;
        if (front != OBSTACLE_ID) { fx = road[front].x; fh = road[front].heading; fs = road[front].speed; }
        double d = lane_s(e.lane, fx) - lane_s(e.lane, e.x);
        double ch = cos(e.heading), sh = sin(e.heading);
        double dv = (e.speed * ch - fs * cos(fh)) * ch + (e.speed * sh - fs * sin(fh)) * sh;
        double d_star = 10.0 + e.speed * 1.5 + e.speed * dv / (2 * sqrt(15.0));
        acc -= 3.0 * pow(d_star / not_zero(d), 2.0);
    }
    return acc;
}
// idm_controller.py:60-79 (mobil's gain is identically 0 in the look-ahead: no lane change is ever started there)
MM_HD_CALL inline void generate_actions(Veh *road, int n, int self, double draw_steer, double draw_acc) {
    Veh &v = road[self];
    int front = front_on_own_lane(road, n, self);
    follow_road(v);
    int steer_lane = v.lane == v.target_lane ? v.target_lane : v.lane;
    v.steer = clipd(steering_control(steer_lane, v) * (draw_steer * 0.1 + 0.95), -MAX_STEER, MAX_STEER);
    v.acc = clipd(idm_acceleration(road, self, front) * (draw_acc * 0.1 + 0.95), -6.0, 6.0);
}
MM_HD_CALL inline void idm_controller(Veh &v) {
    if (v.crashed) { push_traj(v); return; }
    clip_actions(v.steer, v.acc, v.speed, false);
    bicycle(v, v.steer, v.acc);
}

MM_HD inline bool in_group(int query, int lane) {       // road.py:315-342
    switch (query) {
    case L_AB0: return lane == L_AB0 || lane == L_BC0;
    case L_BC0: return lane == L_AB0 || lane == L_BC0 || lane == L_CD0;
    case L_CD0: return lane == L_BC0 || lane == L_CD0;
    case L_JK0: return lane == L_JK0 || lane == L_KB0;
    case L_KB0: return lane == L_JK0 || lane == L_KB0 || lane == L_BC1;
    default: return lane == L_KB0 || lane == L_BC1;
    }
}
MM_HD inline void surrounding(const Veh *road, int n, int self, int lane, int &front, int &rear) {
    double s = road[self].x, s_front = 0, s_rear = 0;
    front = rear = -1;
    for (int j = 0; j < n; ++j) {
        if (j == self || !in_group(lane, road[j].lane)) continue;
        double s_v = road[j].x;
        if (s <= s_v && (front < 0 || s_v <= s_front)) { s_front = s_v; front = j; }
        if (s_v < s && (rear < 0 || s_v > s_rear)) { s_rear = s_v; rear = j; }
    }
}
// (v_fl, v_rl, v_fr, v_rr), central_layer.py:84-110 / decentralised_dmc.py:139-168
MM_HD_CALL inline void neighbour_sets(const Veh *road, int n, int self, int nb[4]) {
    const Veh &v = road[self];
    int fl = -1, rl = -1, fr = -1, rr = -1;
    if (lane_main(v.lane)) {
        surrounding(road, n, self, v.lane, fl, rl);
        if (side_lane(v.lane) >= 0) surrounding(road, n, self, side_lane(v.lane), fr, rr);
        else if (v.lane == L_AB0 && v.x > 220.0) surrounding(road, n, self, L_KB0, fr, rr);
    } else {
        surrounding(road, n, self, v.lane, fr, rr);
        if (side_lane(v.lane) >= 0) surrounding(road, n, self, side_lane(v.lane), fl, rl);
        else if (v.lane == L_KB0) surrounding(road, n, self, L_AB0, fl, rl);
    }
    nb[0] = fl; nb[1] = rl; nb[2] = fr; nb[3] = rr;
}

MM_HD inline bool point_in_rotated_rectangle(double px, double py, double cx, double cy, double l, double w, double a) {
    double c = cos(a), s = sin(a), dx = px - cx, dy = py - cy;
    double rx = c * dx - s * dy, ry = s * dx + c * dy;
    return -l / 2 <= rx && rx <= l / 2 && -w / 2 <= ry && ry <= w / 2;
}
MM_HD inline bool has_corner_inside(double c1x, double c1y, double l1, double w1, double a1, double c2x, double c2y, double l2,
                                    double w2, double a2) {
    const double px[9] = {0, -l1 / 2, l1 / 2, 0, 0, -l1 / 2, -l1 / 2, l1 / 2, l1 / 2};
    const double py[9] = {0, 0, 0, -w1 / 2, w1 / 2, -w1 / 2, w1 / 2, -w1 / 2, w1 / 2};
    double c = cos(a1), s = sin(a1);
    for (int k = 0; k < 9; ++k)
        if (point_in_rotated_rectangle(c1x + c * px[k] - s * py[k], c1y + s * px[k] + c * py[k], c2x, c2y, l2, w2, a2)) return true;
    return false;
}
MM_HD_CALL inline bool is_colliding(const Veh &v, double ox, double oy, double oh, double olen, double owid) {
    if (hypot(ox - v.x, oy - v.y) > VLEN) return false;
    return has_corner_inside(v.x, v.y, 0.9 * VLEN, 0.9 * VWID, v.heading, ox, oy, 0.9 * olen, 0.9 * owid, oh) ||
           has_corner_inside(ox, oy, 0.9 * olen, 0.9 * owid, oh, v.x, v.y, 0.9 * VLEN, 0.9 * VWID, v.heading);
}
MM_HD inline double min_abs(double a, double b) { return fabs(b) < fabs(a) ? b : a; }   // min([a, b], key=abs)

// abstract.py:219-240 -> bit mask over the 5 actions, in the list order IDLE, LANE_LEFT, LANE_RIGHT, FASTER, SLOWER
MM_HD inline int available_actions(const Veh &v, int acts[5]) {
    int n = 0;
    acts[n++] = A_IDLE;
    int sl = side_lane(v.lane);
    if (sl >= 0) {
        if (sl < v.lane && is_reachable_from(sl, v.x, v.y)) acts[n++] = A_LANE_LEFT;
        if (sl > v.lane && is_reachable_from(sl, v.x, v.y)) acts[n++] = A_LANE_RIGHT;
    }
    if (v.speed_index < 4) acts[n++] = A_FASTER;
    if (v.speed_index > 0) acts[n++] = A_SLOWER;
    return n;
}
// abstract.py:242-280: `c` keeps being stepped across calls; only its first NPTS trajectory points are ever read.
// nb = (v_fl, *, v_fr, *) in positions 0 and 2 for both supervisors
MM_HD_CALL inline double check_safety_room(Veh &c, int action, const Veh *road, const int nb[4], int time_steps) {
    double best = 0;
    for (int t = 0; t <= time_steps; ++t) {
        mdp_controller(c, action);
        double room = c.lane == L_BC1 ? 420.0 - c.x : 100.0;
        if (action == A_LANE_LEFT || action == A_LANE_RIGHT) {
            for (int q = 0; q < 4; ++q)
                if (nb[q] >= 0 && fabs(road[nb[q]].tx[t] - c.tx[t]) <= room) room = fabs(road[nb[q]].tx[t] - c.tx[t]);
        } else {
            int o = lane_main(c.lane) ? nb[0] : nb[2];
            if (o >= 0 && road[o].tx[t] - c.tx[t] <= room) room = road[o].tx[t] - c.tx[t];
        }
        if (t == 0 || room < best) best = room;
    }
    return best;
}

// The device's two-pass mode of the dmc supervisor (supervisor.cu): a predicted collision is RECORDED - the ego, its
// available actions, and the x-trajectories of its four neighbours, which is all check_safety_room reads of `road` - and
// evaluated later, one (task, action) pair per lane.  Nothing after the evaluation depends on its result except the
// ego's entry of the action tuple, so the scan of the scene goes on without it.
struct DmcTask {
    int env, cav, n_acts, acts[5], nb[4];
    double tx[4][NPTS];
};
struct DmcTaskSink {
    DmcTask *tasks;
    int *count;          // atomically incremented; may run past `capacity` (those collisions are evaluated in place)
    int capacity, env;
};
// check_safety_room on a recorded task: tx[q][t] == road[nb[q]].tx[t]
MM_HD_CALL inline double check_safety_room_tx(Veh &c, int action, const double (*tx)[NPTS], const int nb[4], int time_steps) {
    double best = 0;
    for (int t = 0; t <= time_steps; ++t) {
        mdp_controller(c, action);
        double room = c.lane == L_BC1 ? 420.0 - c.x : 100.0;
        if (action == A_LANE_LEFT || action == A_LANE_RIGHT) {
            for (int q = 0; q < 4; ++q)
                if (nb[q] >= 0 && fabs(tx[q][t] - c.tx[t]) <= room) room = fabs(tx[q][t] - c.tx[t]);
        } else {
            const int q = lane_main(c.lane) ? 0 : 2;
            if (nb[q] >= 0 && tx[q][t] - c.tx[t] <= room) room = tx[q][t] - c.tx[t];
        }
        if (t == 0 || room < best) best = room;
    }
    return best;
}

// abstract.py:620-635
MM_HD inline double headway_distance(const Veh *road, int n, int self) {
    const Veh &v = road[self];
    double headway = 60;
    int nxt = next_lane(v.lane, v.x, v.y);
    for (int j = 0; j < n; ++j) {
        const Veh &o = road[j];
        if (o.lane == v.lane && o.x > v.x) headway = fmin(headway, o.x - v.x);
        if (v.lane != L_BC1 && o.lane == nxt && o.x > v.x) headway = fmin(headway, o.x - v.x);
    }
    return headway;
}
// central_layer.py:33-63 / decentralised_dmc.py:88-117: CAV indices by ascending priority number (PriorityQueue)
MM_HD_CALL inline void priority_order(const Veh *road, int n, int n_cav, const double *draws, double headway_time, int order[MAXV]) {
    double key[MAXV];
    for (int i = 0; i < n_cav; ++i) {
        const Veh &v = road[i];
        double p = 0.0;
        if (v.lane == L_BC1) { p = -0.5; p -= (100.0 - (420.0 - v.x)) / 100.0; }
        if (v.speed > 0) p += 0.5 * log(headway_distance(road, n, i) / (headway_time * v.speed));
        p += draws[i] * 0.001;
        key[i] = p;
        int q = i;
        while (q > 0 && key[order[q - 1]] > p) { order[q] = order[q - 1]; --q; }   // stable insertion
        order[q] = i;
    }
}

// the collision test of one look-ahead point (abstract.py:721-755 for the four neighbours, then the obstacle)
MM_HD_CALL inline void collide_at(Veh *road, int self, const int nb[4], int t) {
    Veh &v = road[self];
    for (int q = 0; q < 4; ++q) {
        int o = nb[q];
        if (o < 0 || v.crashed || o == self) continue;
        if (is_colliding(v, road[o].tx[t], road[o].ty[t], road[o].th[t], VLEN, VWID)) {
            v.speed = min_abs(v.speed, road[o].tv[t]);
            v.crashed = road[o].crashed = 1;
        }
    }
    if (!v.crashed && is_colliding(v, OBST_X, OBST_Y, 0.0, 2.0, 2.0)) { v.speed = min_abs(v.speed, 0.0); v.crashed = 1; }
}

// decentralised_dmc.py:70-198.  `road`: working copy (modified); `orig`: the scene; actions[n_cav] in/out.
MM_HD inline void dmc_supervisor(Veh *road, const Veh *orig, int n, int n_cav, int *actions, const double *draws,
                                 double headway_time, int *n_used = nullptr, DmcTaskSink *sink = nullptr) {
    (void)sink;
    int order[MAXV];
    priority_order(road, n, n_cav, draws, headway_time, order);
    int k = n_cav;
    for (int q = 0; q < n_cav; ++q)
        for (int t = 0; t < NPTS; ++t) mdp_controller(road[order[q]], actions[order[q]]);
    for (int j = n_cav; j < n; ++j) {
        generate_actions(road, n, j, draws[k], draws[k + 1]);
        k += 2;
        for (int t = 0; t < NPTS; ++t) idm_controller(road[j]);
    }
    if (n_used) *n_used = k;      // np.random.rand() calls of this supervisor run: one per CAV, two per IDM vehicle
    for (int q = 0; q < n_cav; ++q) {
        const int i = order[q];
        Veh &v = road[i];
        int fl_rl_fr_rr[4], nb[4], acts[5];
        neighbour_sets(road, n, i, fl_rl_fr_rr);
        nb[0] = fl_rl_fr_rr[0]; nb[1] = fl_rl_fr_rr[3]; nb[2] = fl_rl_fr_rr[2]; nb[3] = fl_rl_fr_rr[1];   // [fl, rr, fr, rl]
        const int n_acts = available_actions(orig[i], acts);
        v.crashed = 0;
        for (int t = 0; t < NPTS; ++t) {
            v.x = v.tx[t]; v.y = v.ty[t]; v.heading = v.th[t];
            collide_at(road, i, nb, t);
            if (v.crashed) {
#if defined(__CUDA_ARCH__)
                if (sink) {
                    const int slot = atomicAdd(sink->count, 1);
                    if (slot < sink->capacity) {
                        DmcTask &task = sink->tasks[slot];
                        task.env = sink->env; task.cav = i; task.n_acts = n_acts;
                        for (int a = 0; a < 5; ++a) task.acts[a] = a < n_acts ? acts[a] : A_IDLE;
                        for (int q = 0; q < 4; ++q) {
                            task.nb[q] = nb[q];
                            if (nb[q] >= 0)
                                for (int tt = 0; tt < NPTS; ++tt) task.tx[q][tt] = road[nb[q]].tx[tt];
                        }
                        break;
                    }
                }
#endif
                double best_room = 0;
                int best = 0;
                for (int a = 0; a < n_acts; ++a) {
                    Veh c = orig[i];
                    c.n_traj = 0;
                    double room = 0;
                    for (int tt = 0; tt < NPTS; ++tt) room += check_safety_room(c, acts[a], road, nb, tt);
                    if (a == 0 || room > best_room) { best_room = room; best = a; }
                }
                actions[i] = acts[best];
                break;
            }
        }
    }
}

// central_layer.py:16-178 (is_priority = True)
MM_HD inline void priority_supervisor(Veh *road, const Veh *orig, int n, int n_cav, int *actions, const double *draws,
                                      double headway_time, int *n_used = nullptr) {
    int order[MAXV];
    priority_order(road, n, n_cav, draws, headway_time, order);
    int k = n_cav;
    for (int turn = 0; turn < n_cav; ++turn) {
        const int i = order[turn];
        bool first_change = true;
        if (road[i].n_traj == NPTS) { road[i] = orig[i]; road[i].n_traj = 0; }   // moved as a neighbour before: restart
        int nb[4], acts[5];
        const int n_acts = available_actions(road[i], acts);
        neighbour_sets(road, n, i, nb);                                            // [fl, rl, fr, rr]
        const int step_order[5] = {nb[0], nb[2], i, nb[1], nb[3]};                 // v_fl, v_fr, ego, v_rl, v_rr
        for (int t = 0; t < NPTS; ++t) {
            for (int q = 0; q < 5; ++q) {
                const int o = step_order[q];
                if (o < 0) continue;
                if (road[o].n_traj == NPTS && turn != 0 && o != i) continue;
                if (!road[o].cav) {
                    if (t == 0) { generate_actions(road, n, o, draws[k], draws[k + 1]); k += 2; }
                    idm_controller(road[o]);
                } else {
                    // actions[v.id] with id == 0 for every vehicle (controller.py:49): the FIRST CAV's action
                    mdp_controller(road[o], o == i ? actions[i] : actions[0]);
                }
            }
            collide_at(road, i, nb, t);
            if (road[i].crashed) {
                double best_room = 0;
                int best = 0;
                Veh best_c = orig[i];
                for (int a = 0; a < n_acts; ++a) {
                    Veh c = orig[i];
                    c.n_traj = 0;
                    double room = check_safety_room(c, acts[a], road, nb, t);
                    if (a == 0 || room > best_room) { best_room = room; best = a; best_c = c; }
                }
                road[i] = best_c;
                if (first_change) { first_change = false; actions[i] = acts[best]; }
                for (int q = 0; q < 4; ++q)
                    if (nb[q] >= 0 && road[nb[q]].crashed) road[nb[q]].crashed = 0;
            }
        }
    }
    if (n_used) *n_used = k;      // one per CAV, then two per IDM decision the look-aheads made
}

}  // namespace mmsup
