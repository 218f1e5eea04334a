/* merge_oracle.c — CPU float64 restatement (see merge_oracle.h).  TEST INFRASTRUCTURE ONLY.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Expression order mirrors the Python source so that IEEE-754 double results agree to the last
 * bit wherever libm agrees; build with -ffp-contract=off (Python never fuses multiply-add).
 */
#include "merge_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MAXV MO_MAXV
#define PI 3.141592653589793

enum { L_AB0 = 0, L_BC0 = 1, L_BC1 = 2, L_CD0 = 3, L_JK0 = 4, L_KB0 = 5, N_LANES = 6 };
enum { A_LANE_LEFT = 0, A_IDLE = 1, A_LANE_RIGHT = 2, A_FASTER = 3, A_SLOWER = 4 };

/* merge_env_v1.py:222-248 with ends=[220,100,100,1000] (abstract.py:78); graph insertion order */
static const double LANE_SX[N_LANES] = {0.0, 320.0, 320.0, 420.0, 0.0, 220.0};
static const double LANE_SY[N_LANES] = {0.0, 0.0, 4.0, 0.0, 10.5, 7.25};
static const double LANE_LEN[N_LANES] = {320.0, 100.0, 100.0, 1000.0, 220.0, 100.0};
static const int LANE_FORBIDDEN[N_LANES] = {0, 0, 1, 0, 1, 1};
#define SINE_AMPLITUDE 3.25
#define OBST_X 420.0
#define OBST_Y 4.0

#define VEH_LENGTH 5.0
#define VEH_WIDTH 2.0
#define LANE_WIDTH 4.0
#define MAX_SPEED 40.0
#define PERCEPTION 180.0

typedef struct {
    double x, y, heading, speed, target_speed, gvx, rec1_x, rec1_vx, rec2_x, rec2_vx;
    double act_steer, act_acc, safe_steer, safe_acc, timer, min_headway, steering_angle;
    int kind, lane, target_lane, speed_index, crashed, hl_action, hist_len, fg_set;
    int is_collaborating, is_lc_safe, collaborate_adj;
} veh_t;

typedef struct {
    veh_t v[MAXV];
    int n_veh, n_cav, n_merge, steps, time;
} env_t;

/* ---------------------------------------------------------------- utils.py */

/* utils.py:31-37 */
static double not_zero(double x) {
    const double eps = 1e-2;
    if (fabs(x) > eps) return x;
    else if (x > 0) return eps;
    else return -eps;
}

/* Python float % (floored modulo), used by utils.py:40-41 and behavior.py:54 */
static double pymod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0) {
        if ((b < 0) != (r < 0)) r += b;
    } else {
        r = copysign(0.0, b);
    }
    return r;
}

/* utils.py:40-41 */
static double wrap_to_pi(double x) { return pymod(x + PI, 2 * PI) - PI; }

static double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* utils.py:16-18 */
static double lmap(double v, double x0, double x1, double y0, double y1) {
    return y0 + (v - x0) * (y1 - y0) / (x1 - x0);
}

/* ---------------------------------------------------------------- lane.py */

static double sine_pulsation(void) { return 2 * PI / (2 * 100.0); } /* merge_env_v1.py:240 */
static double sine_phase(void) { return PI / 2; }

/* lane.py:164-168, 208-210: direction=(1,0), direction_lateral=(-0,1) for every lane */
static void lane_local(int lane, double px, double py, double *s, double *r) {
    double dx = px - LANE_SX[lane], dy = py - LANE_SY[lane];
    double lon = dx * 1.0 + dy * 0.0;
    double lat = dx * -0.0 + dy * 1.0;
    if (lane == L_KB0) lat = lat - SINE_AMPLITUDE * sin(sine_pulsation() * lon + sine_phase());
    *s = lon;
    *r = lat;
}

/* lane.py:158-159, 204-206 */
static double lane_heading_at(int lane, double s) {
    if (lane == L_KB0)
        return 0.0 + atan(SINE_AMPLITUDE * sine_pulsation() * cos(sine_pulsation() * s + sine_phase()));
    return 0.0;
}

/* lane.py:61-76 */
static int lane_on_lane(int lane, double px, double py, double margin) {
    double s, r;
    lane_local(lane, px, py, &s, &r);
    return fabs(r) <= LANE_WIDTH / 2 + margin && -VEH_LENGTH <= s && s < LANE_LEN[lane] + VEH_LENGTH;
}

/* lane.py:78-90 */
static int lane_is_reachable_from(int lane, double px, double py) {
    double s, r;
    if (LANE_FORBIDDEN[lane]) return 0;
    lane_local(lane, px, py, &s, &r);
    return fabs(r) <= 2 * LANE_WIDTH && 0 <= s && s < LANE_LEN[lane] + VEH_LENGTH;
}

/* lane.py:92-95 */
static int lane_after_end(int lane, double px, double py) {
    double s, r;
    lane_local(lane, px, py, &s, &r);
    return s > LANE_LEN[lane] - VEH_LENGTH / 2;
}

/* lane.py:97-100 */
static double lane_distance(int lane, double px, double py) {
    double s, r;
    lane_local(lane, px, py, &s, &r);
    return fabs(r) + fmax(s - LANE_LEN[lane], 0) + fmax(0 - s, 0);
}

/* lane.py:102-108 */
static double lane_distance_with_heading(int lane, double px, double py, double heading) {
    double s, r;
    lane_local(lane, px, py, &s, &r);
    double angle = fabs(wrap_to_pi(heading - lane_heading_at(lane, s)));
    return fabs(r) + fmax(s - LANE_LEN[lane], 0) + fmax(0 - s, 0) + 1.0 * angle;
}

/* ---------------------------------------------------------------- road.py */

/* road.py:51-65: np.argmin -> first minimum in graph insertion order */
static int closest_lane_index(double px, double py, double heading) {
    int best = 0;
    double bd = lane_distance_with_heading(0, px, py, heading);
    for (int l = 1; l < N_LANES; ++l) {
        double d = lane_distance_with_heading(l, px, py, heading);
        if (d < bd) { bd = d; best = l; }
    }
    return best;
}

/* road.py:67-109 with route=None; every node has one successor so randint(1)==0 */
static int next_lane(int lane, double px, double py) {
    switch (lane) {
    case L_AB0:
    case L_KB0: /* 1 lane -> 2 lanes: min(range(2), key=distance) keeps the first minimum */
        return lane_distance(L_BC0, px, py) <= lane_distance(L_BC1, px, py) ? L_BC0 : L_BC1;
    case L_BC0:
    case L_BC1:
        return L_CD0;
    case L_JK0:
        return L_KB0; /* same lane count -> same id */
    default:
        return L_CD0; /* graph["d"] KeyError -> current index (road.py:96-98) */
    }
}

/* road id (from,to) of a lane and its index on that road */
static int lane_road(int lane) { return lane == L_BC1 ? L_BC0 : lane; }
static int lane_id_on_road(int lane) { return lane == L_BC1 ? 1 : 0; }

/* kinematics.py:161-173 */
static double lane_distance_to(const veh_t *self, double ox, double oy) {
    double s0, r0, s1, r1;
    lane_local(self->lane, ox, oy, &s1, &r1);
    lane_local(self->lane, self->x, self->y, &s0, &r0);
    return s1 - s0;
}

/* ---------------------------------------------------------------- controller.py */

#define TAU_DS 0.2
#define PURSUIT_TAU (0.5 * TAU_DS)
#define KP_A (1 / 0.6)
#define KP_HEADING (1 / TAU_DS)
#define KP_LATERAL (1.0 / 3 * KP_HEADING)
#define MAX_STEERING_ANGLE (PI / 3)

/* controller.py:146-187 (safe_controller.py:84-98 is the identity for lateral_ctrl == "steer") */
static double steering_control(const veh_t *v, int target_lane) {
    double s, r;
    lane_local(target_lane, v->x, v->y, &s, &r);
    double lane_next = s + v->speed * PURSUIT_TAU;
    double future_heading = lane_heading_at(target_lane, lane_next);
    double lateral_speed_command = -KP_LATERAL * r;
    double heading_command = asin(clipd(lateral_speed_command / not_zero(v->speed), -1, 1));
    double heading_ref = future_heading + clipd(heading_command, -PI / 4, PI / 4);
    double heading_rate_command = KP_HEADING * wrap_to_pi(heading_ref - v->heading);
    double steering = asin(clipd(VEH_LENGTH / 2 / not_zero(v->speed) * heading_rate_command, -1, 1));
    return clipd(steering, -MAX_STEERING_ANGLE, MAX_STEERING_ANGLE);
}

/* safe_controller.py:84-98: MDPLCVehicle.steering_control — a steering VELOCITY in steer_vel mode */
static double lc_steering_control(const veh_t *v, int target_lane, int steer_vel) {
    double steering_ref = steering_control(v, target_lane);
    if (steer_vel) {
        steering_ref = steering_ref * 0.125;            /* STEER_TARGET_RF */
        return 20 * (steering_ref - v->steering_angle); /* KP_STEER */
    }
    return steering_ref;
}

/* controller.py:136-144 */
static void follow_road(veh_t *v) {
    if (lane_after_end(v->target_lane, v->x, v->y)) v->target_lane = next_lane(v->target_lane, v->x, v->y);
}

/* controller.py:327-337; np.round is half-to-even == rint() in the default rounding mode */
static int speed_to_index(double speed) {
    double x = (speed - 10.0) / (30.0 - 10.0);
    return (int)clipd(rint(x * (5 - 1)), 0, 5 - 1);
}

/* controller.py:90-134 called with a meta action (or -1 for None), after MDPVehicle.act (293-311) */
static void cav_act(veh_t *v, int action, int steer_vel) {
    if (action == A_FASTER || action == A_SLOWER) {
        int idx = speed_to_index(v->speed) + (action == A_FASTER ? 1 : -1);
        idx = idx < 0 ? 0 : (idx > 4 ? 4 : idx);
        v->speed_index = idx;
        v->target_speed = 10.0 + idx * (30.0 - 10.0) / (5 - 1);
        action = -1;
    }
    follow_road(v);
    /* LANE_RIGHT/LANE_LEFT: candidate = clip(id +- 1) on the target lane's road; the only candidate that
       is a different, non-forbidden lane in this network is bc0 from bc1 */
    if (action == A_LANE_LEFT) {
        if (v->target_lane == L_BC1) {
            if (lane_is_reachable_from(L_BC0, v->x, v->y)) v->target_lane = L_BC0;
        }
    }
    double steering = lc_steering_control(v, v->target_lane, steer_vel);
    double acc = KP_A * (v->target_speed - v->speed);
    v->act_steer = clipd(steering, -MAX_STEERING_ANGLE, MAX_STEERING_ANGLE);
    v->act_acc = acc;
}

/* controller.py:257-267 */
static void get_corner(const veh_t *v, int left, double *cx, double *cy) {
    const double corner_len = sqrt((VEH_WIDTH / 2) * (VEH_WIDTH / 2) + (VEH_LENGTH / 2) * (VEH_LENGTH / 2)) + 0.0075;
    const double corner_alpha = atan(VEH_WIDTH / VEH_LENGTH);
    *cx = v->x + (corner_len * cos(corner_alpha + v->heading));
    if (left)
        *cy = v->y - (corner_len * sin(corner_alpha + v->heading)) + 0.01;
    else
        *cy = v->y - (corner_len * sin(-corner_alpha + v->heading)) + 0.01;
}

/* ---------------------------------------------------------------- behavior.py (IDM / MOBIL) */

/* index MAXV denotes the obstacle (road.objects[0]) */
#define OBST MAXV

static void ent_pos(const env_t *e, int i, double *x, double *y) {
    if (i == OBST) { *x = OBST_X; *y = OBST_Y; } else { *x = e->v[i].x; *y = e->v[i].y; }
}

/* road.py:352-381 */
static void neighbour_vehicles(const env_t *e, int self, int lane, int *front, int *rear) {
    double s, r, s_front = 0, s_rear = 0;
    lane_local(lane, e->v[self].x, e->v[self].y, &s, &r);
    *front = -1;
    *rear = -1;
    for (int j = 0; j <= e->n_veh; ++j) {
        int id = (j == e->n_veh) ? OBST : j;
        if (id == self) continue;
        double px, py, s_v, lat_v;
        ent_pos(e, id, &px, &py);
        lane_local(lane, px, py, &s_v, &lat_v);
        if (!(fabs(lat_v) <= LANE_WIDTH / 2 + 1 && -VEH_LENGTH <= s_v && s_v < LANE_LEN[lane] + VEH_LENGTH)) continue;
        if (s <= s_v && (*front < 0 || s_v <= s_front)) { s_front = s_v; *front = id; }
        if (s_v < s && (*rear < 0 || s_v > s_rear)) { s_rear = s_v; *rear = id; }
    }
}

/* behavior.py:141-156 */
static double desired_gap(const env_t *e, int ego, int front) {
    const veh_t *a = &e->v[ego];
    double fvx = 0, fvy = 0;
    if (front != OBST) {
        const veh_t *b = &e->v[front];
        fvx = b->speed * cos(b->heading);
        fvy = b->speed * sin(b->heading);
    }
    double dirx = cos(a->heading), diry = sin(a->heading);
    double evx = a->speed * dirx, evy = a->speed * diry;
    double dv = (evx - fvx) * dirx + (evy - fvy) * diry;
    double ab = -3.0 * -5.0;
    return 10.0 + a->speed * 1.5 + a->speed * dv / (2 * sqrt(ab));
}

/* behavior.py:111-139; ego/front are entity ids, -1 = None, OBST = obstacle */
static double idm_acceleration(const env_t *e, int ego, int front) {
    if (ego < 0 || ego == OBST) return 0;
    const veh_t *a = &e->v[ego];
    double ego_target_speed = not_zero(a->target_speed);
    double acc = 3.0 * (1 - pow(fmax(a->speed, 0) / ego_target_speed, 4.0));
    if (front >= 0) {
        double fx, fy;
        ent_pos(e, front, &fx, &fy);
        double d = lane_distance_to(a, fx, fy);
        double q = desired_gap(e, ego, front) / not_zero(d);
        acc -= 3.0 * (q * q);
    }
    return acc;
}

/* behavior.py:225-266 with route=None, POLITENESS=0 */
static int mobil(const env_t *e, int self, int lane) {
    int new_preceding, new_following, old_preceding, old_following;
    neighbour_vehicles(e, self, lane, &new_preceding, &new_following);
    double new_following_a = idm_acceleration(e, new_following, new_preceding);
    double new_following_pred_a = idm_acceleration(e, new_following, self);
    if (new_following_pred_a < -9.0) return 0;
    neighbour_vehicles(e, self, e->v[self].lane, &old_preceding, &old_following);
    double self_pred_a = idm_acceleration(e, self, new_preceding);
    double self_a = idm_acceleration(e, self, old_preceding);
    double old_following_a = idm_acceleration(e, old_following, self);
    double old_following_pred_a = idm_acceleration(e, old_following, old_preceding);
    double jerk = self_pred_a - self_a + 0. * (new_following_pred_a - new_following_a + old_following_pred_a - old_following_a);
    if (jerk < 0.1) return 0;
    return 1;
}

/* behavior.py:186-223 */
static void change_lane_policy(env_t *e, int self) {
    veh_t *v = &e->v[self];
    if (v->lane != v->target_lane) {
        if (lane_road(v->lane) == lane_road(v->target_lane)) {
            for (int j = 0; j < e->n_veh; ++j) {
                const veh_t *o = &e->v[j];
                if (j != self && o->lane != v->target_lane && o->target_lane == v->target_lane) {
                    double d = lane_distance_to(v, o->x, o->y);
                    double d_star = desired_gap(e, self, j);
                    if (0 < d && d < d_star) { v->target_lane = v->lane; break; }
                }
            }
        }
        return;
    }
    if (!(1.0 < v->timer)) return; /* utils.do_every */
    v->timer = 0;
    /* side_lanes: bc0 -> [bc1], bc1 -> [bc0]; bc1 is forbidden so only bc1 -> bc0 can be reachable */
    int side = v->lane == L_BC0 ? L_BC1 : (v->lane == L_BC1 ? L_BC0 : -1);
    if (side < 0) return;
    if (!lane_is_reachable_from(side, v->x, v->y)) return;
    if (mobil(e, self, side)) v->target_lane = side;
}

/* behavior.py:74-100 */
static void hdv_act(env_t *e, int self) {
    veh_t *v = &e->v[self];
    if (v->crashed) return;
    int front, rear;
    neighbour_vehicles(e, self, v->lane, &front, &rear);
    follow_road(v);
    change_lane_policy(e, self);
    double steering = steering_control(v, v->target_lane);
    steering = clipd(steering, -MAX_STEERING_ANGLE, MAX_STEERING_ANGLE);
    double acc = idm_acceleration(e, self, front);
    acc = clipd(acc, -6.0, 6.0);
    v->act_steer = steering;
    v->act_acc = acc;
}

/* ---------------------------------------------------------------- kinematics.py */

/* kinematics.py:143-152 */
static void clip_actions(veh_t *v) {
    if (v->crashed) {
        v->act_steer = 0;
        v->act_acc = -1.0 * v->speed;
    }
    if (v->speed > MAX_SPEED) v->act_acc = fmin(v->act_acc, 1.0 * (MAX_SPEED - v->speed));
    else if (v->speed < -MAX_SPEED) v->act_acc = fmax(v->act_acc, 1.0 * (MAX_SPEED - v->speed));
}

/* kinematics.py:133-141 / safe_controller.py:151-172: bicycle model with (steering, acceleration) */
static void integrate(veh_t *v, double steering, double acceleration, double dt, double *beta_out) {
    double beta = atan(1.0 / 2 * tan(steering));
    double vx = v->speed * cos(v->heading + beta);
    double vy = v->speed * sin(v->heading + beta);
    v->x += vx * dt;
    v->y += vy * dt;
    v->heading += v->speed * sin(beta) / (VEH_LENGTH / 2) * dt;
    v->speed += acceleration * dt;
    v->speed = fmax(0, v->speed);
    *beta_out = beta;
}

/* utils.py:55-70 — note the rotation uses +angle (as written in the reference) */
static int point_in_rotated_rectangle(double px, double py, double cx, double cy, double length, double width, double angle) {
    double c = cos(angle), s = sin(angle);
    double dx = px - cx, dy = py - cy;
    double rx = c * dx + -s * dy;
    double ry = s * dx + c * dy;
    return -length / 2 <= rx && rx <= length / 2 && -width / 2 <= ry && ry <= width / 2;
}

/* utils.py:102-121 */
static int has_corner_inside(double c1x, double c1y, double l1, double w1, double a1,
                             double c2x, double c2y, double l2, double w2, double a2) {
    const double lx = l1 / 2, wy = w1 / 2;
    const double pts[9][2] = {{0, 0}, {-lx, 0}, {lx, 0}, {0, -wy}, {0, wy}, {-lx, -wy}, {-lx, wy}, {lx, -wy}, {lx, wy}};
    double c = cos(a1), s = sin(a1);
    for (int k = 0; k < 9; ++k) {
        double rx = c * pts[k][0] + -s * pts[k][1];
        double ry = s * pts[k][0] + c * pts[k][1];
        if (point_in_rotated_rectangle(c1x + rx, c1y + ry, c2x, c2y, l2, w2, a2)) return 1;
    }
    return 0;
}

/* kinematics.py:202-209 */
static int is_colliding(const veh_t *a, double ox, double oy, double oheading, double olen, double owid) {
    double dx = ox - a->x, dy = oy - a->y;
    if (sqrt(dx * dx + dy * dy) > VEH_LENGTH) return 0;
    return has_corner_inside(a->x, a->y, 0.9 * VEH_LENGTH, 0.9 * VEH_WIDTH, a->heading, ox, oy, 0.9 * olen, 0.9 * owid, oheading)
        || has_corner_inside(ox, oy, 0.9 * olen, 0.9 * owid, oheading, a->x, a->y, 0.9 * VEH_LENGTH, 0.9 * VEH_WIDTH, a->heading);
}

/* road.py:288-292 + kinematics.py:175-200 */
static void collision_loop(env_t *e) {
    for (int i = 0; i < e->n_veh; ++i) {
        veh_t *a = &e->v[i];
        for (int j = 0; j < e->n_veh; ++j) {
            if (a->crashed || j == i) continue;
            veh_t *b = &e->v[j];
            if (is_colliding(a, b->x, b->y, b->heading, VEH_LENGTH, VEH_WIDTH)) {
                double m = fabs(a->speed) <= fabs(b->speed) ? a->speed : b->speed;
                a->speed = b->speed = m;
                a->crashed = b->crashed = 1;
            }
        }
        if (!a->crashed && is_colliding(a, OBST_X, OBST_Y, 0.0, 2.0, 2.0)) {
            a->speed = fabs(a->speed) <= 0 ? a->speed : 0;
            a->crashed = 1;
        }
    }
}

/* ---------------------------------------------------------------- shields: cbf.py + decentral_layer.py */

double mo_qp(double a, double c_lead, double c_adj, int has_adj, double lo, double hi, int32_t *active_out) {
    /* closed-form minimiser of cbf.py:110-135's QP (P=diag(1,1,1e18), rows cbf.py:288-322/374-422);
       mirrors oracle/refharness/stubs/cvxopt/solvers.py::solve_clamp */
    int active = 0;
    double u = 0.0;
    if (u > hi) { u = hi; active = MO_ACT_UPPER; }
    if (u < lo) { u = lo; active = MO_ACT_LOWER; }
    if (a > 0.0) {
        double lim = c_lead / a;
        int row = MO_ACT_LEAD;
        if (has_adj && c_adj / a < lim) { lim = c_adj / a; row = MO_ACT_ADJ; }
        if (u > lim) {
            if (lim >= lo) { u = lim; active = row; }
            else { u = lo; active = row | MO_ACT_LOWER | MO_ACT_SLACK; }
        }
    } else if (a < 0.0) {
        double lim = c_lead / a;
        int row = MO_ACT_LEAD;
        if (has_adj && c_adj / a > lim) { lim = c_adj / a; row = MO_ACT_ADJ; }
        if (u < lim) {
            if (lim <= hi) { u = lim; active = row; }
            else { u = hi; active = row | MO_ACT_UPPER | MO_ACT_SLACK; }
        }
    } else {
        double c = has_adj ? fmin(c_lead, c_adj) : c_lead;
        if (c < 0.0) active |= MO_ACT_SLACK | MO_ACT_LEAD;
    }
    if (active_out) *active_out = active;
    return u;
}

/* decentral_layer.py:15-20 */
static int is_same_lane(const veh_t *v, int lane2) {
    int nl = next_lane(v->lane, v->x, v->y);
    return v->lane == lane2 || lane2 == nl;
}

/* decentral_layer.py:23-39 */
static int is_adj_lane(const veh_t *v, int lane2) {
    int l1 = v->lane;
    int nl = next_lane(l1, v->x, v->y);
    if (lane_road(l1) == lane_road(lane2) && abs(lane_id_on_road(l1) - lane_id_on_road(lane2)) == 1)
        return lane_id_on_road(l1) - lane_id_on_road(lane2);
    else if (lane_road(nl) == lane_road(lane2) && abs(lane_id_on_road(nl) - lane_id_on_road(lane2)) == 1)
        return lane_id_on_road(nl) - lane_id_on_road(lane2);
    return 0;
}

/* decentral_layer.py:46-57 */
static int is_approaching_same_lane(const veh_t *ve, const veh_t *vl) {
    if (lane_distance_to(ve, vl->x, vl->y) < 0) return 0;
    double y_dist = vl->y - ve->y;
    int dist_cond = fabs(y_dist) <= 3.5;
    int heading_cond = y_dist < 0 ? (vl->heading > 0.037) : (vl->heading < -0.037);
    return dist_cond && heading_cond;
}

/* road.py:257-267 — stable sort by |lane_distance_to|, first `count` */
static int close_vehicles_to(const env_t *e, int self, int count, int *out) {
    const veh_t *a = &e->v[self];
    double key[MAXV];
    int n = 0;
    for (int j = 0; j < e->n_veh; ++j) {
        if (j == self) continue;
        const veh_t *b = &e->v[j];
        double dx = b->x - a->x, dy = b->y - a->y;
        if (!(sqrt(dx * dx + dy * dy) < PERCEPTION)) continue;
        double k = fabs(lane_distance_to(a, b->x, b->y));
        int p = n;
        while (p > 0 && key[p - 1] > k) { key[p] = key[p - 1]; out[p] = out[p - 1]; --p; } /* stable insertion */
        key[p] = k;
        out[p] = j;
        ++n;
    }
    return n < count ? n : count;
}

typedef struct {
    int ran, leader, front_adj, rear_adj, constrain_adj, active, is_lc_safe;
    double safe_acc, safe_steer, nom_acc, nom_steer, lc_margin;
} shield_rec;

/* decentral_layer.py:767-817 -> 290-518 (HSS) / 521-764 (MASS) with multi_agent_state 85-257 and
   cbf.py CBF_AV/CBF_CAV.  Writes v->safe_*, may overwrite v->target_lane.  SURVEY.md Appendix B. */
static void shield(const mo_config *cfg, env_t *e, int self, shield_rec *rec) {
    veh_t *v = &e->v[self];
    const double dt = cfg->dt, eta = cfg->eta, tau = cfg->tau;
    const int mass = cfg->shield == MO_SHIELD_MASS;
    const double acc_lo = -12.5, acc_hi = 6.0; /* CBF_AV.ACCELERATION_RANGE == MDPLCVehicle MIN/MAX_ACC */

    /* safety_layer: velocity box */
    double v_min = v->speed + acc_lo * dt;
    if (mass) v_min = fmax(0, v_min);
    double v_max = v->speed + acc_hi * dt;

    double ex = v->x;
    double evx_raw = v->speed * cos(v->heading);
    double evx = evx_raw > 1 ? evx_raw : 1;

    /* virtual stopped vehicles beyond perception */
    double x_ol = ex + PERCEPTION + 1, x_oa = ex + PERCEPTION + 1, x_oar = ex - PERCEPTION - 1;
    int has_ol = 0, has_oa = 0, has_oar = 0;
    double vx_ol = 0, vx_oa = 0, vx_oar = 0;
    double a_ol = 0, a_oa = 0, g_ol = 0, g_oa = 0; /* MASS: defaults of multi_agent_state (a 0, gp.vx 0) */
    int constrain_adj = 0;
    int id_ol = MO_NB_NONE, id_oa = MO_NB_NONE, id_oar = MO_NB_NONE;

    int nb[MAXV];
    int n_nb = close_vehicles_to(e, self, 5, nb);
    for (int k = 0; k < n_nb; ++k) {
        veh_t *o = &e->v[nb[k]];
        int v_a = is_adj_lane(v, o->lane);
        int a_v = is_adj_lane(o, v->lane);
        int approaching = is_approaching_same_lane(v, o);
        double d = lane_distance_to(v, o->x, o->y);
        int o_is_cav = o->kind == MO_KIND_CAV;
        if (!approaching && (v_a || a_v)) {
            if (!has_oar && d < 0) {
                has_oar = 1;
                x_oar = o->x;
                vx_oar = o->speed * cos(o->heading);
                id_oar = nb[k];
            } else if (!has_oa && d >= 0) {
                has_oa = 1;
                x_oa = o->rec2_x;
                vx_oa = o->rec2_vx;
                id_oa = nb[k];
                if (mass) {
                    a_oa = o_is_cav ? o->safe_acc : acc_lo;
                    g_oa = o_is_cav ? o->gvx : 1;
                    double cx, cy;
                    int have_corner = 0;
                    if (v_a == -1 || a_v == 1) { get_corner(o, 1, &cx, &cy); have_corner = 1; }
                    else if (v_a == 1 || a_v == -1) { get_corner(o, 0, &cx, &cy); have_corner = 1; }
                    if (have_corner) constrain_adj = !lane_on_lane(o->lane, cx, cy, 0);
                }
            }
        } else if (!o_is_cav && v->lane == L_AB0 && o->lane == L_KB0 && d >= 0) {
            /* on-ramp HDV twin: the record is mutated in place (decentral_layer.py:175-184) */
            o->rec2_x = o->rec2_x + 0.5 * evx_raw;
            has_oa = 1;
            x_oa = o->rec2_x;
            vx_oa = o->rec2_vx;
            id_oa = nb[k];
            constrain_adj = 1;
            a_oa = acc_lo;
            g_oa = 1;
        } else if (!has_ol && (is_same_lane(v, o->lane) || approaching) && d > 0) {
            has_ol = 1;
            x_ol = o->rec2_x;
            vx_ol = o->rec2_vx;
            id_ol = nb[k];
            if (mass) {
                a_ol = o_is_cav ? o->safe_acc : acc_lo;
                g_ol = o_is_cav ? o->gvx : 1;
            }
        }
    }
    /* obstacle (decentral_layer.py:213-246) */
    if (!(v->x > OBST_X)) {
        if ((!has_ol || OBST_X <= x_ol) && fabs(OBST_Y - v->y) <= 2) {
            has_ol = 1; x_ol = OBST_X; vx_ol = 0; id_ol = MO_NB_OBSTACLE;
            if (mass) { a_ol = 0; g_ol = 0; }
        }
        if ((!has_oa || OBST_X <= x_oa) && 2 < fabs(OBST_Y - v->y) && fabs(OBST_Y - v->y) <= 4) {
            has_oa = 1; x_oa = OBST_X; vx_oa = 0; id_oa = MO_NB_OBSTACLE;
            if (mass) { a_oa = 0; g_oa = 0; constrain_adj = 0; }
        }
    }
    if (!mass) { g_ol = 1; g_oa = 1; } /* HSS: g = diag(ge*dt, dt, dt, ...) */

    /* safe distances (decentral_layer.py:448-468) */
    double sv_oar = has_oar ? vx_oar : 0;
    sv_oar = sv_oar + acc_hi * dt;
    sv_oar = sv_oar > 1 ? sv_oar : 1;
    double buffer = (acc_hi + 0.1) * dt * tau;
    double sd_l = evx * tau + VEH_LENGTH + buffer;
    double sd_a = sd_l;
    double sd_r = sv_oar * tau + VEH_LENGTH + buffer;
    v->min_headway = (x_ol - ex - VEH_LENGTH) / evx;

    /* one-step velocity predictions (simplified_control, decentral_layer.py:60-77) */
    double v_ll = fmax(0, evx + v->act_acc * dt);
    double v_ol = has_ol ? fmax(0, vx_ol + (mass ? a_ol : acc_lo) * dt) : 0;
    double v_oa = has_oa ? fmax(0, vx_oa + (mass ? a_oa : acc_lo) * dt) : 0;
    double v_oar = has_oar ? fmax(0, vx_oar + acc_hi * dt) : 0;

    /* define_pq (cbf.py:262-283, 365-372) */
    double q_lon = -VEH_LENGTH - sd_l;
    double q_lona = -VEH_LENGTH - sd_a;
    if (mass && constrain_adj) q_lona = -VEH_LENGTH - sd_a - 2.0134;
    double q_lonr = -VEH_LENGTH - sd_r;

    /* G, h (cbf.py:288-322, 374-422); g is diagonal so every dot() has at most two non-zero terms */
    double ge_dt = v->gvx * dt, gol_dt = g_ol * dt, goa_dt = g_oa * dt, gr_dt = 1 * dt;
    double a = ge_dt;
    double dl = -ex + x_ol;            /* dot(p_lon, x)  */
    double da = -ex + x_oa;            /* dot(p_lona, x) */
    double dr = ex + -x_oar;           /* dot(p_lonr, x) */
    double c_lead = dl + (eta - 1) * dl + eta * q_lon + (-(ge_dt * v_ll) + gol_dt * v_ol);
    double hi = v_max - v_ll;
    double lo = -(-v_min + v_ll);
    int has_adj = mass && constrain_adj;
    double c_adj = 0;
    if (has_adj) c_adj = da + (eta - 1) * da + eta * q_lona + (-(ge_dt * v_ll) + goa_dt * v_oa);

    int32_t active = 0;
    double u = mo_qp(a, c_lead, c_adj, has_adj, lo, hi, &active);
    double v_safe = v_ll + u;

    double steer = v->act_steer;
    v->is_collaborating = constrain_adj;
    v->is_lc_safe = 1;

    /* is_lc_allowed (cbf.py:324-339) */
    double hls_a = da + q_lona;
    double hlds_a = da + (-ge_dt * v_safe + goa_dt * v_oa) + q_lona;
    double hls_r = dr + q_lonr;
    double hlds_r = dr + (ge_dt * v_safe + -gr_dt * v_oar) + q_lonr;
    double cond_a = hlds_a + (eta - 1) * hls_a;
    double cond_r = hlds_r + (eta - 1) * hls_r;
    /* Tie rule (DESIGN.md "Veto tie rule"): when the adjacent row is the binding QP constraint (and feasible),
       cond_a is 0 in exact arithmetic and +-1e-15 rounding noise in float64; it is taken
       as satisfied, which is also what an interior-point solve (real cvxopt) yields. */
    int adj_inv_ok = (active == MO_ACT_ADJ) ? 1 : (cond_a >= 0);
    int allowed = (hls_a >= 0 && adj_inv_ok) && (hls_r >= 0 && cond_r >= 0);
    rec->lc_margin = fmin(fmin(fabs(hls_a), fabs(cond_a)), fmin(fabs(hls_r), fabs(cond_r)));

    if (!mass) {
        if (!allowed) {
            v->target_lane = v->lane;
            steer = lc_steering_control(v, v->target_lane, cfg->steer_vel);
            v->is_lc_safe = 0;
        }
    } else {
        double cx, cy;
        int can_abort = 1;
        get_corner(v, 1, &cx, &cy);
        can_abort = can_abort && lane_on_lane(v->lane, cx, cy, 0);
        get_corner(v, 0, &cx, &cy);
        can_abort = can_abort && lane_on_lane(v->lane, cx, cy, 0);
        if (can_abort && !allowed) {
            v->target_lane = v->lane;
            steer = lc_steering_control(v, v->target_lane, cfg->steer_vel);
            v->is_lc_safe = 0;
        } else if ((v->hl_action == A_LANE_RIGHT || v->hl_action == A_LANE_LEFT) && v->speed < 1.6667) {
            v_safe = v_ll;
        }
        /* can_collaborate_adj (cbf.py:424-430) sees the QP's v_safe (u_safe_ma is a copy made before the
           bypass above); carried but never read on a live path (Appendix B.14) */
        v->collaborate_adj = cond_a >= -1e-6;
    }

    v->safe_acc = (v_safe - evx) / dt; /* derived_acceleration */
    v->safe_steer = steer;

    rec->ran = 1;
    rec->leader = id_ol;
    rec->front_adj = id_oa;
    rec->rear_adj = id_oar;
    rec->constrain_adj = constrain_adj;
    rec->active = active;
    rec->is_lc_safe = v->is_lc_safe;
    rec->safe_acc = v->safe_acc;
    rec->safe_steer = v->safe_steer;
    rec->nom_acc = v->act_acc;
    rec->nom_steer = v->act_steer;
}

/* ---------------------------------------------------------------- road.act / road.step */

/* road.py:277/286: sorted(vehicles, key=x, reverse=True) — stable, so ties keep list order */
static void order_by_x_desc(const env_t *e, int *ord) {
    for (int i = 0; i < e->n_veh; ++i) {
        int p = i;
        while (p > 0 && e->v[ord[p - 1]].x < e->v[i].x) { ord[p] = ord[p - 1]; --p; }
        ord[p] = i;
    }
}

/* safe_controller.py:106-185 */
static void cav_step(const mo_config *cfg, env_t *e, int self, shield_rec *rec) {
    veh_t *v = &e->v[self];
    clip_actions(v);
    if (!cfg->env_v0) v->act_acc = clipd(v->act_acc, -12.5, 6.0); /* safe_controller.py:100-104; MDPVehicle (v0) steps with kinematics.py:122-141 */
    if (cfg->env_v0 || cfg->shield == MO_SHIELD_NONE || !v->fg_set || v->hist_len < 2) { /* safe_controller.py:229-239 */
        v->safe_acc = v->act_acc;
        v->safe_steer = v->act_steer;
    } else {
        shield(cfg, e, self, rec);
    }
    double beta;
    if (cfg->steer_vel && !cfg->env_v0) {
        /* safe_controller.py:124-150: the steering angle is a state driven by the steering velocity; note that
           the heading increment is NOT multiplied by dt (as written in the reference) */
        beta = atan(1.0 / 2 * tan(v->steering_angle));
        double vx = v->speed * cos(v->heading + beta);
        double vy = v->speed * sin(v->heading + beta);
        v->x += vx * cfg->dt;
        v->y += vy * cfg->dt;
        double d_heading = v->speed * sin(beta) / (VEH_LENGTH / 2);
        v->heading += d_heading;
        v->speed += v->safe_acc * cfg->dt;
        v->steering_angle += v->safe_steer * cfg->dt;
        v->speed = fmax(0, v->speed);
    } else {
        integrate(v, v->safe_steer, v->safe_acc, cfg->dt, &beta);
    }
    v->gvx = cos(v->heading + beta);
    v->fg_set = 1;
    v->lane = closest_lane_index(v->x, v->y, v->heading);
    /* log_step (safe_controller.py:187-205): state_hist.append(to_dict()) */
    v->rec2_x = v->rec1_x;
    v->rec2_vx = v->rec1_vx;
    v->rec1_x = v->x;
    v->rec1_vx = v->speed * cos(v->heading);
    if (v->hist_len < 2) v->hist_len++;
}

/* behavior.py:504-522 (IDMVehicleHist.step) -> behavior.py:102-109 -> kinematics.py:122-141 */
static void hdv_step(const mo_config *cfg, env_t *e, int self) {
    veh_t *v = &e->v[self];
    v->timer += cfg->dt;
    clip_actions(v);
    double beta;
    integrate(v, v->act_steer, v->act_acc, cfg->dt, &beta);
    v->lane = closest_lane_index(v->x, v->y, v->heading);
    v->rec2_x = v->rec1_x;
    v->rec2_vx = v->rec1_vx;
    v->rec1_x = v->x;
    v->rec1_vx = v->speed * cos(v->heading);
    if (v->hist_len < 2) v->hist_len++;
}

/* vehicles that are observed / rewarded: the controlled ones, or all of them in MergeEnvLCHDV (observation.py:430-442,
 * merge_env_v1.py:518-524) */
static int n_out(const mo_config *cfg, const env_t *e) { return cfg->env_hdv ? e->n_veh : e->n_cav; }

/* merge_env_v1.py:168-172; MergeEnvLCHDV: 670-673 (any vehicle crashed, no x < 0 clause) */
static int is_terminal(const mo_config *cfg, const env_t *e) {
    for (int i = 0; i < n_out(cfg, e); ++i)
        if (e->v[i].crashed) return 1;
    if (e->steps >= cfg->duration_steps) return 1;
    if (cfg->env_hdv) return 0;
    for (int i = 0; i < e->n_cav; ++i)
        if (e->v[i].x < 0) return 1;
    return 0;
}

/* ---------------------------------------------------------------- observation / reward */

/* observation.py:241-273 + 181-193, absolute=False, normalize=True, clip=False, "steer" mode */
static void observe_agent(const env_t *e, int self, int steer_vel, double *obs /* [MO_NS] */) {
    const veh_t *a = &e->v[self];
    memset(obs, 0, sizeof(double) * MO_NS);
    double evx = a->speed * cos(a->heading), evy = a->speed * sin(a->heading);
    double rows[MO_OBS_ROWS][MO_OBS_FEATS];
    int n_rows = 1;
    rows[0][0] = 1; rows[0][1] = a->x; rows[0][2] = a->y; rows[0][3] = evx; rows[0][4] = evy; rows[0][5] = a->heading;
    int nb[MAXV];
    int n_nb = close_vehicles_to(e, self, MO_OBS_ROWS - 1, nb);
    for (int k = 0; k < n_nb; ++k) {
        const veh_t *o = &e->v[nb[k]];
        double *r = rows[n_rows++];
        r[0] = 1;
        r[1] = o->x - a->x;
        r[2] = o->y - a->y;
        r[3] = o->speed * cos(o->heading) - evx;
        r[4] = o->speed * sin(o->heading) - evy;
        r[5] = o->heading;
        if (steer_vel && o->kind == MO_KIND_CAV) r[5] = o->heading - a->heading; /* safe_controller.py:75-81 */
    }
    for (int k = 0; k < n_rows; ++k) {
        double *o = obs + k * MO_OBS_FEATS;
        o[0] = rows[k][0];
        o[1] = lmap(rows[k][1], -5.0 * 30, 5.0 * 30, -1, 1);
        o[2] = lmap(rows[k][2], -12, 12, -1, 1);
        o[3] = lmap(rows[k][3], -1.5 * 30, 1.5 * 30, -1, 1);
        o[4] = lmap(rows[k][4], -1.5 * 30, 1.5 * 30, -1, 1);
        o[5] = lmap(rows[k][5], -PI / 2, PI / 2, -1, 1);
    }
}

/* abstract.py:620-635 */
static double headway_distance(const env_t *e, int self) {
    const veh_t *a = &e->v[self];
    double hd = 60;
    int nl = next_lane(a->lane, a->x, a->y);
    for (int j = 0; j < e->n_veh; ++j) {
        const veh_t *o = &e->v[j];
        if (o->lane == a->lane && o->x > a->x) {
            double d = o->x - a->x;
            if (d < hd) hd = d;
        }
        if (a->lane != L_BC1 && o->lane == nl && o->x > a->x) {
            double d = o->x - a->x;
            if (d < hd) hd = d;
        }
    }
    return hd;
}

/* merge_env_v1.py:64-89 (default) and 439-474 (srew / mrew; only for MDPLCVehicle) */
static double agent_reward(const mo_config *cfg, const env_t *e, int self) {
    const veh_t *a = &e->v[self];
    int special = (cfg->reward_kind == MO_REW_SREW || cfg->reward_kind == MO_REW_MREW) && a->kind == MO_KIND_CAV;
    int is_mrew = cfg->reward_kind == MO_REW_MREW;
    double r1 = 30.0;
    if (special && a->is_collaborating && is_mrew) r1 = 10.0 + (30.0 - 10.0) / 2;
    double scaled_speed = lmap(a->speed, 10.0, r1, 0, 1);
    double merging_lane_cost = 0;
    if (a->lane == L_BC1 && (!special || !is_mrew || a->is_lc_safe)) {
        double d = a->x - 420.0;
        merging_lane_cost = -exp(-(d * d) / (10 * 100.0));
    }
    double hd = headway_distance(e, self);
    double headway_cost = 0;
    if (a->speed > 0) {
        headway_cost = log(hd / (cfg->headway_time * a->speed));
        if (special) headway_cost = -1 * headway_cost;
    }
    return cfg->collision_reward * (-1 * a->crashed) + (cfg->high_speed_reward * clipd(scaled_speed, 0, 1))
         + cfg->merging_lane_cost * merging_lane_cost + cfg->headway_cost * (headway_cost < 0 ? headway_cost : 0);
}

/* road.py:294-350 */
static void surrounding_vehicles(const env_t *e, int self, int lane, int *front, int *rear) {
    static const int GROUP[N_LANES][N_LANES] = {
        /* query ab0 */ {1, 1, 0, 0, 0, 0},
        /* query bc0 */ {1, 1, 0, 1, 0, 0},
        /* query bc1 */ {0, 0, 1, 0, 0, 1},
        /* query cd0 */ {0, 1, 0, 1, 0, 0},
        /* query jk0 */ {0, 0, 0, 0, 1, 1},
        /* query kb0 */ {0, 0, 1, 0, 1, 1},
    };
    double s = e->v[self].x, s_front = 0, s_rear = 0;
    *front = -1;
    *rear = -1;
    for (int j = 0; j < e->n_veh; ++j) {
        if (j == self || !GROUP[lane][e->v[j].lane]) continue;
        double s_v = e->v[j].x;
        if (s <= s_v && (*front < 0 || s_v <= s_front)) { s_front = s_v; *front = j; }
        if (s_v < s && (*rear < 0 || s_v > s_rear)) { s_rear = s_v; *rear = j; }
    }
}

/* merge_env_v1.py:91-124 */
static double regional_reward(const env_t *e, int self, const double *local) {
    const veh_t *a = &e->v[self];
    int fl = -1, rl = -1, fr = -1, rr = -1;
    if (a->lane == L_AB0 || a->lane == L_BC0 || a->lane == L_CD0) {
        surrounding_vehicles(e, self, a->lane, &fl, &rl);
        if (a->lane == L_BC0) surrounding_vehicles(e, self, L_BC1, &fr, &rr);
        else if (a->lane == L_AB0 && a->x > 220) surrounding_vehicles(e, self, L_KB0, &fr, &rr);
    } else {
        surrounding_vehicles(e, self, a->lane, &fr, &rr);
        if (a->lane == L_BC1) surrounding_vehicles(e, self, L_BC0, &fl, &rl);
        else if (a->lane == L_KB0) surrounding_vehicles(e, self, L_AB0, &fl, &rl);
    }
    int order[5] = {fl, fr, self, rl, rr};
    double sum = 0;
    int cnt = 0;
    for (int k = 0; k < 5; ++k) {
        int j = order[k];
        if (j >= 0 && e->v[j].kind == MO_KIND_CAV) { sum = sum + local[j]; ++cnt; }
    }
    return sum / cnt;
}

/* merge_env_v1.py:373-386 */
static double min_time_headway(const mo_config *cfg, const env_t *e) {   /* hdv env: 587-601, over every vehicle */
    double mh = INFINITY;
    for (int i = 0; i < n_out(cfg, e); ++i) {
        const veh_t *a = &e->v[i];
        double hd = headway_distance(e, i);
        if (fabs(OBST_Y - a->y) <= 2 && OBST_X > a->x) {
            double d = OBST_X - a->x;
            if (d < hd) hd = d;
        }
        hd = hd - VEH_LENGTH;
        double vx = a->speed * cos(a->heading);
        mh = fmin(mh, hd / (vx > 1 ? vx : 1));
    }
    return mh;
}

/* ---------------------------------------------------------------- SoA <-> AoS */

static void load_env(const mo_state *s, int ei, env_t *e) {
    memset(e, 0, sizeof(*e));
    e->n_veh = s->n_veh[ei]; e->n_cav = s->n_cav[ei]; e->n_merge = s->n_merge[ei];
    e->steps = s->steps[ei]; e->time = s->time[ei];
    for (int i = 0; i < MAXV; ++i) {
        int k = ei * MAXV + i;
        veh_t *v = &e->v[i];
        v->x = s->x[k]; v->y = s->y[k]; v->heading = s->heading[k]; v->speed = s->speed[k];
        v->target_speed = s->target_speed[k]; v->gvx = s->gvx[k];
        v->rec1_x = s->rec1_x[k]; v->rec1_vx = s->rec1_vx[k]; v->rec2_x = s->rec2_x[k]; v->rec2_vx = s->rec2_vx[k];
        v->act_steer = s->act_steer[k]; v->act_acc = s->act_acc[k];
        v->safe_steer = s->safe_steer[k]; v->safe_acc = s->safe_acc[k];
        v->timer = s->timer[k]; v->min_headway = s->min_headway[k]; v->steering_angle = s->steering_angle[k];
        v->kind = s->kind[k]; v->lane = s->lane[k]; v->target_lane = s->target_lane[k];
        v->speed_index = s->speed_index[k]; v->crashed = s->crashed[k]; v->hl_action = s->hl_action[k];
        v->hist_len = s->hist_len[k]; v->fg_set = s->fg_set[k];
        v->is_collaborating = s->is_collaborating[k]; v->is_lc_safe = s->is_lc_safe[k];
        v->collaborate_adj = s->collaborate_adj[k];
    }
}

static void store_env(const mo_state *s, int ei, const env_t *e) {
    s->n_veh[ei] = e->n_veh; s->n_cav[ei] = e->n_cav; s->n_merge[ei] = e->n_merge;
    s->steps[ei] = e->steps; s->time[ei] = e->time;
    for (int i = 0; i < MAXV; ++i) {
        int k = ei * MAXV + i;
        const veh_t *v = &e->v[i];
        s->x[k] = v->x; s->y[k] = v->y; s->heading[k] = v->heading; s->speed[k] = v->speed;
        s->target_speed[k] = v->target_speed; s->gvx[k] = v->gvx;
        s->rec1_x[k] = v->rec1_x; s->rec1_vx[k] = v->rec1_vx; s->rec2_x[k] = v->rec2_x; s->rec2_vx[k] = v->rec2_vx;
        s->act_steer[k] = v->act_steer; s->act_acc[k] = v->act_acc;
        s->safe_steer[k] = v->safe_steer; s->safe_acc[k] = v->safe_acc;
        s->timer[k] = v->timer; s->min_headway[k] = v->min_headway; s->steering_angle[k] = v->steering_angle;
        s->kind[k] = v->kind; s->lane[k] = v->lane; s->target_lane[k] = v->target_lane;
        s->speed_index[k] = v->speed_index; s->crashed[k] = v->crashed; s->hl_action[k] = v->hl_action;
        s->hist_len[k] = v->hist_len; s->fg_set[k] = v->fg_set;
        s->is_collaborating[k] = v->is_collaborating; s->is_lc_safe[k] = v->is_lc_safe;
        s->collaborate_adj[k] = v->collaborate_adj;
    }
}

/* ---------------------------------------------------------------- one policy step of one env */

static void step_env(const mo_config *cfg, env_t *e, const int8_t *act, const mo_out *out, int ei) {
    shield_rec recs[3][MAXV];
    memset(recs, 0, sizeof(recs));
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < MAXV; ++i) recs[k][i].leader = recs[k][i].front_adj = recs[k][i].rear_adj = MO_NB_NONE;

    /* abstract.py:443-467 */
    e->steps += 1;
    /* abstract.py:512-532 */
    for (int k = 0; k < cfg->substeps; ++k) {
        if (e->time % cfg->substeps == 0) {
            /* action.py:226-231 -> safe_controller.py:63-66 -> controller.py:293-311 */
            for (int i = 0; i < e->n_cav; ++i) {
                e->v[i].hl_action = act[i];
                cav_act(&e->v[i], act[i], cfg->steer_vel && !cfg->env_v0);
            }
        }
        int ord[MAXV];
        order_by_x_desc(e, ord);
        for (int p = 0; p < e->n_veh; ++p) { /* road.py:269-278 */
            int i = ord[p];
            if (e->v[i].kind == MO_KIND_CAV) cav_act(&e->v[i], -1, cfg->steer_vel && !cfg->env_v0);
            else hdv_act(e, i);
        }
        order_by_x_desc(e, ord); /* positions are unchanged by act(): same order (road.py:286) */
        for (int p = 0; p < e->n_veh; ++p) {
            int i = ord[p];
            if (e->v[i].kind == MO_KIND_CAV) cav_step(cfg, e, i, &recs[k < 3 ? k : 2][i]);
            else hdv_step(cfg, e, i);
        }
        collision_loop(e);
        e->time += 1;
        if (is_terminal(cfg, e)) break;
    }

    /* abstract.py:469-498 + merge_env_v1.py:126-166 */
    double *obs = out->obs + (size_t)ei * MAXV * MO_NS;
    memset(obs, 0, sizeof(double) * MAXV * MO_NS);
    double local[MAXV];
    memset(local, 0, sizeof(local));
    double rsum = 0, ssum = 0, tsum = 0;
    const int no = n_out(cfg, e);
    for (int i = 0; i < no; ++i) {
        observe_agent(e, i, cfg->steer_vel && !cfg->env_v0, obs + i * MO_NS);
        local[i] = agent_reward(cfg, e, i);
        rsum += local[i];
        ssum += e->v[i].speed;
    }
    for (int i = 0; i < e->n_veh; ++i) tsum += e->v[i].speed;
    int done = is_terminal(cfg, e);
    out->reward[ei] = rsum / no;
    out->done[ei] = done;
    out->average_speed[ei] = ssum / no;
    out->traffic_speed[ei] = tsum / e->n_veh;
    out->min_headway[ei] = min_time_headway(cfg, e);
    for (int i = 0; i < MAXV; ++i) {
        int k = ei * MAXV + i;
        out->agents_rewards[k] = 0; out->regional_rewards[k] = 0; out->agents_dones[k] = 0;
    }
    for (int i = 0; i < no; ++i) {
        int k = ei * MAXV + i;
        out->agents_rewards[k] = local[i];
        if (cfg->env_hdv) continue;     /* MergeEnvLCHDV.step returns neither regional rewards nor per-agent dones */
        out->regional_rewards[k] = regional_reward(e, i, local);
        out->agents_dones[k] = e->v[i].crashed || e->steps >= cfg->duration_steps || e->v[i].x < 0;
    }
    double mp = -1.0;
    if (done) {
        int n_rem = 0;
        for (int i = 0; i < no; ++i)
            if (e->v[i].lane == L_BC1 || e->v[i].lane == L_KB0 || e->v[i].lane == L_JK0) n_rem++;
        mp = e->n_merge > 0 ? (double)(e->n_merge - n_rem) / e->n_merge * 100 : 100.0;
    }
    out->merge_percent[ei] = mp;
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < MAXV; ++i) {
            size_t idx = ((size_t)ei * 3 + k) * MAXV + i;
            const shield_rec *r = &recs[k][i];
            out->sh_ran[idx] = r->ran; out->sh_leader[idx] = r->leader; out->sh_front_adj[idx] = r->front_adj;
            out->sh_rear_adj[idx] = r->rear_adj; out->sh_constrain_adj[idx] = r->constrain_adj;
            out->sh_active[idx] = r->active; out->sh_is_lc_safe[idx] = r->is_lc_safe;
            out->sh_safe_acc[idx] = r->safe_acc; out->sh_safe_steer[idx] = r->safe_steer;
            out->sh_nom_acc[idx] = r->nom_acc; out->sh_nom_steer[idx] = r->nom_steer;
            out->sh_lc_margin[idx] = r->lc_margin;
        }
}

typedef struct {
    const mo_config *cfg; const mo_state *st; const int8_t *actions; const mo_out *out;
    int begin, end;
} step_job;

static void *step_worker(void *arg) {
    step_job *j = (step_job *)arg;
    for (int ei = j->begin; ei < j->end; ++ei) {
        env_t e;
        load_env(j->st, ei, &e);
        step_env(j->cfg, &e, j->actions + (size_t)ei * MAXV, j->out, ei);
        store_env(j->st, ei, &e);
    }
    return NULL;
}

void mo_step(const mo_config *cfg, const mo_state *st, const int8_t *actions, const mo_out *out, int n_env, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_env) n_threads = n_env > 0 ? n_env : 1;
    if (n_threads > 256) n_threads = 256;
    step_job jobs[256];
    pthread_t tids[256];
    int chunk = (n_env + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        int b = t * chunk, en = b + chunk;
        if (b > n_env) b = n_env;
        if (en > n_env) en = n_env;
        jobs[t] = (step_job){cfg, st, actions, out, b, en};
    }
    for (int t = 1; t < n_threads; ++t) pthread_create(&tids[t], NULL, step_worker, &jobs[t]);
    step_worker(&jobs[0]);
    for (int t = 1; t < n_threads; ++t) pthread_join(tids[t], NULL);
}

void mo_observe(const mo_state *st, double *obs, int n_env, int steer_vel, int env_hdv) {
    for (int ei = 0; ei < n_env; ++ei) {
        env_t e;
        load_env(st, ei, &e);
        double *o = obs + (size_t)ei * MAXV * MO_NS;
        memset(o, 0, sizeof(double) * MAXV * MO_NS);
        for (int i = 0; i < (env_hdv ? e.n_veh : e.n_cav); ++i) observe_agent(&e, i, steer_vel, o + i * MO_NS);
    }
}
