// merge_coop.cu — the warp-cooperative build of the policy-step physics (sm_100a): half a warp per environment, one lane
// per vehicle.  Same semantics as step_kernel in merge_step.cu (reference map there); a different mapping.
//
// Why a second mapping.  step_kernel gives every env one thread and walks its vehicles front to back: 33 serial vehicle
// moves per policy step, ~130 k dependent warp-instructions per warp.  That is the right shape when there are enough envs
// to fill the machine with such threads (>= ~50 k), and a poor one for small batches: 4 096 envs are 128 warps on 592
// warp schedulers, each crawling through its chain at one instruction per ~7 cycles.  Here the vehicles of an env are
// processed side by side:
//
//   * CTA = 256 threads = 16 envs; lanes 0..15 of a half-warp are the vehicle slots of one env (an env has <= 11).
//   * Everything a vehicle does on its own runs once, in parallel over the lanes: the meta-action and the steering /
//     speed control laws (controller.py:90-197), the bicycle step and the closest-lane search (kinematics.py:122-152).
//     A vehicle's new position does not depend on its acceleration, only on its steering command, so the moves of a
//     sub-step can be computed before any shield has run ("nominal move").
//   * The reference runs the shields in x-descending order, each follower seeing the vehicles ahead of it already moved
//     (road.py:286).  A lane reproduces exactly that view: for a vehicle of lower rank it reads the nominal-move arrays,
//     for one of higher rank the sub-step-start state.  With the views in place the neighbour classification of all
//     vehicles (multi_agent_state, decentral_layer.py:85-257) runs in parallel; ranks come from counting (a lane's rank =
//     number of vehicles with larger x, ties by slot: the stable sort), the five nearest from an unrolled selection over
//     register-resident keys (the reference's sorted()[:count], road.py:257-267).
//   * What remains serial in MASS is the chain "my QP needs my leader's shielded acceleration of this sub-step"
//     (decentral_layer.py:133-135, 209-211).  It is evaluated as a fixed-point iteration: every lane solves its
//     closed-form QP with the current accelerations of its leader / front-adjacent vehicle until no value of the env
//     changes (ballot).  The dependencies point to lower ranks only, so after k rounds every vehicle whose chain is <= k
//     long holds its final value, bit for bit what the sequential order gives; platoons are short, 2-4 rounds.
//   * The one thing that can invalidate a nominal move is a lane-change veto that re-steers a vehicle in the middle of a
//     lane change (cbf.py:324-339, decentral_layer.py:501-506 / 728-744; 0.8 % of the shield calls).  The lowest such
//     rank recomputes its move, and the lanes behind it classify and solve again; everything up to that rank is final.
//   * Shared memory holds the state all lanes read (current planes as in step_kernel + the nominal-move / record arrays);
//     a half-warp synchronises with __syncwarp only.  Cold fields live in registers / shared memory for the whole policy
//     step and are written to HBM once (step_kernel rewrites them every sub-step).
//
// Scope: all-CAV envs of env id merge-multi-agent-v1 with lateral_control "steer" (the BASELINE configs[1..4] scenes),
// shields none / HSS / MASS; launch_step falls back to step_kernel for everything else.
#include <cuda_runtime.h>
#include <math_constants.h>
#include "mm_internal.h"

#define MM_KNS mmc
#define MM_NHOT 6
#define MM_PW 17            // 16 env columns per CTA; the odd stride spreads the 11 slots of one column over distinct banks
#define MM_SPEC 1
#define MM_SPEC_SHIELD 2    // only is_cav() / the env-kind reads of the shared device code are folded; the shield kind is a template parameter here
// the steering law and the trigonometry inline, as in the thread-per-env builds (HSS 4 096 envs: no change; MASS: -1.7 %)
#define MM_STEER_FN __forceinline__
#define MM_TRIG_FN __forceinline__
#define MM_TRIG1_FN __forceinline__
#include "mm_device.cuh"

namespace mmc {

constexpr int CENVS = 16;                 // envs per CTA
constexpr int CLANES = 16;                // lanes per env (vehicle slots; 11 used)
constexpr int CTHREADS = CENVS * CLANES;  // 256
constexpr int NARR = 12;                  // f64 arrays [CENVS][CLANES] next to the planes
constexpr size_t COOP_SMEM = (size_t)PLANES_F64 * sizeof(double) + (size_t)NARR * CENVS * CLANES * sizeof(double) +
                             (size_t)CENVS * CLANES * sizeof(uint32_t) + (size_t)CENVS * CLANES;

#ifndef MM_COOP_MIN_BLOCKS
#define MM_COOP_MIN_BLOCKS 2
#endif

struct Geom {          // a nominal move: everything of vehicle_step that depends on the steering command only
    double nx, ny, nh, ncos, nsin, ngvx, sb, cb;
    int lane;
};

// kinematics.py:133-140 / safe_controller.py:151-172 with the trigonometry of vehicle_step (mm_device.cuh): same
// expressions, same order
__device__ __forceinline__ Geom move_geom(double x, double y, double h, double v, double ch, double sh, double steer, double dt) {
    Geom g;
    double t = 1.0 / 2 * m_tan(steer);
    g.cb = 1.0 / sqrt(1.0 + t * t);
    g.sb = t * g.cb;
    double c_hb = ch * g.cb - sh * g.sb, s_hb = sh * g.cb + ch * g.sb;
    g.nx = x + v * c_hb * dt;
    g.ny = y + v * s_hb * dt;
    g.nh = h + div_nz(v * g.sb, VLEN / 2) * dt;
    double2 scn = m_sincos(g.nh);
    g.ncos = scn.y;
    g.nsin = scn.x;
    g.ngvx = scn.y * g.cb - scn.x * g.sb;
    g.lane = closest_lane(g.nx, g.ny, g.nh);
    return g;
}

template <int SHIELD, bool DIAG>
__global__ void __launch_bounds__(CTHREADS, MM_COOP_MIN_BLOCKS) coop_step_kernel(const __grid_constant__ StepParams p) {
    const int lane16 = threadIdx.x & 15, ce = threadIdx.x >> 4;
    const unsigned hmask = 0xFFFFu << (threadIdx.x & 16);          // the 16 lanes of this env inside the warp
    const int local = blockIdx.x * CENVS + ce;
    const bool valid = local < p.env_count;
    const size_t e = (size_t)p.env_offset + (valid ? local : 0);
    const int i = lane16;                                           // own vehicle slot
    const int col = ce * CLANES;                                    // base of this env's row in the [CENVS][CLANES] arrays

    double *A = sm_planes + PLANES_F64;
    double *NX = A, *NY = A + 256, *NH = A + 512, *NCH = A + 768, *NSH = A + 1024, *NGV = A + 1280;   // nominal move
    double *R2X = A + 1536, *R2VX = A + 1792, *R1VX = A + 2048, *SACC = A + 2304, *GVXO = A + 2560;   // records at sub-step start
    double *ACUR = A + 2816;                                        // this sub-step's (shielded) accelerations
    uint32_t *NFL = reinterpret_cast<uint32_t *>(A + NARR * 256);   // flags after the nominal move (lane updated)
    uint8_t *RANK = reinterpret_cast<uint8_t *>(NFL + 256);

    Env ev;
    ev.tid = ce;
    ev.g = p.st.f64 + f64_index(e, 0, 0);
    ev.live = 0;
    ev.pos = 0;
    const uint32_t ei = valid ? p.st.einfo[e] : 0u;
    ev.n_veh = (ei >> EI_NVEH_SHIFT) & EI_4BIT;
    ev.n_cav = (ei >> EI_NCAV_SHIFT) & EI_4BIT;
    int steps = (ei >> EI_STEPS_SHIFT) & EI_STEPS_MASK, time = (ei >> EI_TIME_SHIFT) & EI_TIME_MASK;
    const int n = ev.n_veh;
    const bool has_v = valid && i < n;
    const double dt = p.cfg.dt, eta = p.cfg.eta, tau = p.cfg.tau;
    constexpr bool mass = SHIELD == MM_SHIELD_MASS;

    // state of the own vehicle: hot fields go to the planes every lane reads, cold ones stay in registers
    double gvx = 0, rec1vx = 0, rec2x = 0, rec2vx = 0, act_steer = 0, act_acc = 0, safe_steer = 0, safe_acc = 0, minhw = 0;
    int action = A_IDLE;
    if (has_v) {
#pragma unroll
        for (int f = 0; f < N_HOT; ++f) SMF(f, i) = GF(f, i);
        FL(i) = p.st.flags[flags_index(e, i)];
        gvx = GF(F_GVX, i); rec1vx = GF(F_REC1VX, i); rec2x = GF(F_REC2X, i); rec2vx = GF(F_REC2VX, i);
        act_steer = GF(F_ACT_STEER, i); act_acc = GF(F_ACT_ACC, i);
        safe_steer = GF(F_SAFE_STEER, i); safe_acc = GF(F_SAFE_ACC, i); minhw = GF(F_MINHW, i);
        const int a = (int)p.actions[e * MAXV + i];
        action = (a >= 0 && a <= 4) ? a : A_IDLE;
    }
    if (DIAG && valid && i < MAXV) {   // clear this policy step's record rows of the slot (occupied or not)
        const size_t plane = (size_t)p.n_envs * 3 * MAXV;
        for (int sub = 0; sub < 3; ++sub) {
            const size_t idx = (e * 3 + sub) * MAXV + i;
            int32_t *si = p.out.sh_i;
            si[idx] = 0; si[plane + idx] = MM_NB_NONE; si[2 * plane + idx] = MM_NB_NONE; si[3 * plane + idx] = MM_NB_NONE;
            si[4 * plane + idx] = 0; si[5 * plane + idx] = 0; si[6 * plane + idx] = 0; si[7 * plane + idx] = 0;
            si[8 * plane + idx] = -1; si[9 * plane + idx] = 0;
            for (int q = 0; q < 10; ++q) p.out.sh_f[q * plane + idx] = 0.0;
        }
    }
    if (valid) steps = min(steps + 1, (int)EI_STEPS_MASK);   // abstract.py:457
    bool running = valid;
    uint32_t n_solves = 0, n_active = 0, n_veto = 0;
    __syncwarp();

#pragma unroll 1
    for (int sub = 0; sub < p.cfg.substeps; ++sub) {   // abstract.py:514-531
        const bool act_now = running && has_v;
        const bool apply_meta = time % p.cfg.substeps == 0;   // abstract.py:516-519
        // ---- own controls (MDPLCVehicle.act, controller.py:293-311 + 90-134) and the nominal move, all lanes at once
        double x = 0, y = 0, h = 0, v = 0, ch = 1, sh = 0, steer = 0, acc = 0;
        uint32_t f = 0;
        int rank = 0;
        Geom g{};
        if (act_now) {
            x = X(i); y = Y(i); h = H(i); v = V(i); ch = CH(i); sh = SH(i);
            // road.py:277,286: processing order = stable sort by x, descending
            for (int j = 0; j < n; ++j) {
                const double xj = X(j);
                rank += (xj > x || (xj == x && j < i)) ? 1 : 0;
            }
            RANK[col + i] = (uint8_t)rank;
            cav_act(ev, i, apply_meta ? action : A_NONE, false, steer, acc);
            f = FL(i);
            // clip_actions (kinematics.py:122-131, safe_controller.py:106-122)
            if (f & FL_CRASHED) { steer = 0.0; acc = -1.0 * v; }
            if (v > 40.0) acc = fmin(acc, 1.0 * (40.0 - v));
            else if (v < -40.0) acc = fmax(acc, 1.0 * (40.0 - v));
            acc = clipd(acc, ACC_LO, ACC_HI);
            act_steer = steer;
            act_acc = acc;
            g = move_geom(x, y, h, v, ch, sh, steer, dt);
            NX[col + i] = g.nx; NY[col + i] = g.ny; NH[col + i] = g.nh; NCH[col + i] = g.ncos; NSH[col + i] = g.nsin;
            NGV[col + i] = g.ngvx;
            NFL[col + i] = fl_set(f, FL_LANE_SHIFT, FL_3BIT, (uint32_t)g.lane);
            R2X[col + i] = rec2x; R2VX[col + i] = rec2vx; R1VX[col + i] = rec1vx; SACC[col + i] = safe_acc; GVXO[col + i] = gvx;
            ACUR[col + i] = acc;      // an unshielded CAV's safe action is its clipped nominal one
        }
        __syncwarp();

        // ---- shields (get_safe_action gate, safe_controller.py:229-239)
        const bool shielded = SHIELD != MM_SHIELD_NONE && act_now && (f & FL_FG) && fl_hist(f) >= 2;
        double out_acc = acc, out_steer = steer;
        bool veto = false, lc_safe = true, constrain_adj = false, cadj_flag = false;
        int id_ol = MM_NB_NONE, id_oa = MM_NB_NONE, id_oar = MM_NB_NONE, active = 0;
        double lc_margin = 0.0;
        if (SHIELD != MM_SHIELD_NONE) {
            const int elane = fl_lane(f);
            const double ex = x, ey = y, eh = h, espeed = v;
            int start = 0;
            bool resteered = false;
#pragma unroll 1
            for (;;) {
                const bool redo = shielded && rank >= start;
                // quantities of the QP that do not depend on this sub-step's accelerations of the others
                double evx = 1, v_ll = 0, ge_dt = 0, dl = 0, da = 0, dr = 0, q_lon = 0, q_lona = 0, q_lonr = 0, hi = 0, lo = 0;
                double vx_ol = 0, vx_oa = 0, g_ol = 0, g_oa = 0, v_oar = 0, a_ol_fix = 0, a_oa_fix = 0;
                bool has_ol = false, has_oa = false, has_adj = false, ol_iter = false, oa_iter = false;
                int src_ol = 0, src_oa = 0;
                if (redo) {
                    double v_min = espeed + ACC_LO * dt;
                    if (mass) v_min = fmax(0.0, v_min);
                    const double v_max = espeed + ACC_HI * dt;
                    const double evx_raw = (f & FL_CRASHED) ? espeed * ch : rec1vx;
                    evx = evx_raw > 1 ? evx_raw : 1;
                    const double es = lane_s(elane, ex);
                    // ---- close_vehicles_to(count = 5) over this lane's view of the road (road.py:257-267)
                    double key[SMV];
#pragma unroll
                    for (int j = 0; j < SMV; ++j) {
                        key[j] = CUDART_INF;
                        if (j < n && j != i) {
                            const bool mv = RANK[col + j] < rank;
                            const double ox = mv ? NX[col + j] : X(j), oy = mv ? NY[col + j] : Y(j);
                            const double dx = ox - ex, dy = oy - ey;
                            if (dx * dx + dy * dy < PERCEPTION_SQ_LT) key[j] = fabs(lane_s(elane, ox) - es);
                        }
                    }
                    const int e_next = next_lane(elane, ex, ey);
                    const bool elane_bc = (elane == L_BC0) | (elane == L_BC1);
                    const int e_eff = elane_bc ? elane : e_next;
                    const bool e_eff_bc = (e_eff == L_BC0) | (e_eff == L_BC1);
                    bool oa_left = false;
                    id_ol = MM_NB_NONE; id_oa = MM_NB_NONE; id_oar = MM_NB_NONE;
                    uint32_t taken = 0;
#pragma unroll 1
                    for (int k = 0; k < 5; ++k) {
                        // k-th smallest (key, slot): sorted() is stable, equal keys keep slot order - the first minimum
                        // among the slots not taken yet
                        double best = CUDART_INF;
                        int o = -1;
#pragma unroll
                        for (int j = 0; j < SMV; ++j)
                            if (!((taken >> j) & 1u) && key[j] < best) { best = key[j]; o = j; }
                        if (o < 0) break;
                        taken |= 1u << o;
                        // multi_agent_state (decentral_layer.py:85-257), one candidate
                        const bool mv = RANK[col + o] < rank;
                        const double ox = mv ? NX[col + o] : X(o);
                        const double d = lane_s(elane, ox) - es;
                        if (d < 0 ? id_oar != MM_NB_NONE : (id_oa != MM_NB_NONE && id_ol != MM_NB_NONE)) continue;
                        const uint32_t fo = mv ? NFL[col + o] : FL(o);
                        const int olane = fl_lane(fo);
                        const double oy = mv ? NY[col + o] : Y(o);
                        const bool olane_bc = (olane == L_BC0) | (olane == L_BC1);
                        const int v_a = (e_eff_bc & olane_bc & (e_eff != olane)) ? (e_eff == L_BC1 ? 1 : -1) : 0;
                        int a_v = 0;
                        if (elane_bc) {
                            int o_eff = olane;
                            if (olane == L_AB0 || olane == L_KB0) o_eff = next_lane(olane, ox, oy);
                            const bool o_eff_bc = (o_eff == L_BC0) | (o_eff == L_BC1);
                            a_v = (o_eff_bc & (o_eff != elane)) ? (o_eff == L_BC1 ? 1 : -1) : 0;
                        }
                        bool approaching = false;     // is_approaching_same_lane (decentral_layer.py:46-57)
                        if (!(d < 0)) {
                            const double y_dist = oy - ey, oh = mv ? NH[col + o] : H(o);
                            const bool hc = y_dist < 0 ? (oh > 0.037) : (oh < -0.037);
                            approaching = fabs(y_dist) <= 3.5 && hc;
                        }
                        if (!approaching && (v_a != 0 || a_v != 0)) {
                            if (id_oar == MM_NB_NONE && d < 0) {
                                id_oar = o;
                            } else if (id_oa == MM_NB_NONE && d >= 0) {
                                id_oa = o;
                                oa_left = (v_a == -1 || a_v == 1);
                            }
                        } else if (id_ol == MM_NB_NONE && d > 0) {
                            if ((elane == olane) || (olane == e_next) || approaching) id_ol = o;
                        }
                    }
                    // ---- the records of the three roles through the same view
                    const bool has_ol0 = id_ol != MM_NB_NONE, has_oa0 = id_oa != MM_NB_NONE, has_oar0 = id_oar != MM_NB_NONE;
                    double x_ol = ex + PERCEPTION + 1, x_oa = ex + PERCEPTION + 1, x_oar = ex - PERCEPTION - 1, vx_oar = 0;
                    constrain_adj = false;
                    if (has_oar0) {
                        const int jr = id_oar;
                        const bool mv = RANK[col + jr] < rank;
                        const uint32_t fo = FL(jr);
                        if (mv) {      // cannot happen while vehicles drive forwards; kept for completeness of the view
                            const double nv = fmax(0.0, V(jr) + ACUR[col + jr] * dt);
                            x_oar = NX[col + jr];
                            vx_oar = nv * NCH[col + jr];
                        } else {
                            x_oar = X(jr);
                            vx_oar = (fo & FL_CRASHED) ? V(jr) * CH(jr) : R1VX[col + jr];
                        }
                    }
                    if (has_ol0) {
                        const int jl = id_ol;
                        const bool mv = RANK[col + jl] < rank;
                        // a vehicle that has stepped: its record before that step = its sub-step-start state
                        x_ol = mv ? X(jl) : R2X[col + jl];
                        vx_ol = mv ? R1VX[col + jl] : R2VX[col + jl];
                        if (mass) {
                            g_ol = mv ? NGV[col + jl] : GVXO[col + jl];
                            ol_iter = mv; src_ol = jl;
                            a_ol_fix = SACC[col + jl];
                        }
                    }
                    if (has_oa0) {
                        const int ja = id_oa;
                        const bool mv = RANK[col + ja] < rank;
                        x_oa = mv ? X(ja) : R2X[col + ja];
                        vx_oa = mv ? R1VX[col + ja] : R2VX[col + ja];
                        if (mass) {
                            g_oa = mv ? NGV[col + ja] : GVXO[col + ja];
                            oa_iter = mv; src_oa = ja;
                            a_oa_fix = SACC[col + ja];
                            double cx, cy;
                            get_corner(mv ? NX[col + ja] : X(ja), mv ? NY[col + ja] : Y(ja), mv ? NCH[col + ja] : CH(ja),
                                       mv ? NSH[col + ja] : SH(ja), oa_left, cx, cy);
                            constrain_adj = !on_lane(fl_lane(mv ? NFL[col + ja] : FL(ja)), cx, cy, 0.0);
                        }
                    }
                    has_ol = has_ol0; has_oa = has_oa0;
                    bool has_oar = has_oar0;
                    // the obstacle can take over either role (decentral_layer.py:213-246)
                    if (!(ex > OBST_X)) {
                        const double ady = fabs(OBST_Y - ey);
                        if ((!has_ol || OBST_X <= x_ol) && ady <= 2) {
                            has_ol = true; id_ol = MM_NB_OBSTACLE; x_ol = OBST_X; vx_ol = 0;
                            if (mass) { ol_iter = false; a_ol_fix = 0; g_ol = 0; }
                        }
                        if ((!has_oa || OBST_X <= x_oa) && 2 < ady && ady <= 4) {
                            has_oa = true; id_oa = MM_NB_OBSTACLE; x_oa = OBST_X; vx_oa = 0;
                            if (mass) { oa_iter = false; a_oa_fix = 0; g_oa = 0; constrain_adj = false; }
                        }
                    }
                    if (!mass) { g_ol = 1.0; g_oa = 1.0; }
                    // safe distances and headway (decentral_layer.py:448-468)
                    double sv_oar = (has_oar ? vx_oar : 0.0) + ACC_HI * dt;
                    sv_oar = sv_oar > 1 ? sv_oar : 1;
                    const double buffer = (ACC_HI + 0.1) * dt * tau;
                    const double sd_l = evx * tau + VLEN + buffer;
                    const double sd_r = sv_oar * tau + VLEN + buffer;
                    minhw = (x_ol - ex - VLEN) / evx;
                    v_ll = fmax(0.0, evx + act_acc * dt);
                    v_oar = has_oar ? fmax(0.0, vx_oar + ACC_HI * dt) : 0.0;
                    q_lon = -VLEN - sd_l;
                    q_lona = -VLEN - sd_l;
                    has_adj = mass && constrain_adj;
                    if (has_adj) q_lona = -VLEN - sd_l - 2.0134;
                    q_lonr = -VLEN - sd_r;
                    ge_dt = gvx * dt;
                    dl = -ex + x_ol; da = -ex + x_oa; dr = ex + -x_oar;
                    hi = v_max - v_ll;
                    lo = -(-v_min + v_ll);
                }
                __syncwarp(hmask);
                // ---- the QPs: fixed-point rounds over the accelerations of this sub-step (one round in HSS)
                bool allowed = true;
                double v_safe = 0, cond_a = 0, hls_a = 0, hls_r = 0, cond_r = 0;
                bool can_abort = false, can_abort_known = false;
                int rounds = 0;
#pragma unroll 1
                for (;;) {
                    bool changed = false;
                    if (redo) {
                        const double a_ol = mass ? (ol_iter ? ACUR[col + src_ol] : a_ol_fix) : ACC_LO;
                        const double a_oa = mass ? (oa_iter ? ACUR[col + src_oa] : a_oa_fix) : ACC_LO;
                        const double gol_dt = g_ol * dt, goa_dt = g_oa * dt, gr_dt = 1 * dt;
                        // one-step predictions (decentral_layer.py:60-77); a missing role predicts 0
                        const double v_ol = has_ol ? fmax(0.0, vx_ol + a_ol * dt) : 0.0;
                        const double v_oa = has_oa ? fmax(0.0, vx_oa + a_oa * dt) : 0.0;
                        const double c_lead = dl + (eta - 1) * dl + eta * q_lon + (-(ge_dt * v_ll) + gol_dt * v_ol);
                        double c_adj = 0.0;
                        if (has_adj) c_adj = da + (eta - 1) * da + eta * q_lona + (-(ge_dt * v_ll) + goa_dt * v_oa);
                        const double u = solve_cbf_qp(ge_dt, c_lead, c_adj, has_adj, lo, hi, active);
                        v_safe = v_ll + u;
                        // lane-change veto (cbf.py:324-339)
                        hls_a = da + q_lona;
                        const double hlds_a = da + (-ge_dt * v_safe + goa_dt * v_oa) + q_lona;
                        hls_r = dr + q_lonr;
                        const double hlds_r = dr + (ge_dt * v_safe + -gr_dt * v_oar) + q_lonr;
                        cond_a = hlds_a + (eta - 1) * hls_a;
                        cond_r = hlds_r + (eta - 1) * hls_r;
                        const bool adj_inv_ok = (active == MM_ACT_ADJ) ? true : (cond_a >= 0);
                        allowed = (hls_a >= 0 && adj_inv_ok) && (hls_r >= 0 && cond_r >= 0);
                        if (!mass) {
                            veto = !allowed;
                        } else {
                            veto = false;
                            if (!allowed) {
                                if (!can_abort_known) {     // can_abort_lc (decentral_layer.py:728-736): own corners on the own lane
                                    double cx, cy;
                                    get_corner(ex, ey, ch, sh, true, cx, cy);
                                    can_abort = on_lane(elane, cx, cy, 0.0);
                                    if (can_abort) {
                                        get_corner(ex, ey, ch, sh, false, cx, cy);
                                        can_abort = on_lane(elane, cx, cy, 0.0);
                                    }
                                    can_abort_known = true;
                                }
                                veto = can_abort;
                            }
                            const int hl = fl_hl(f);
                            if (!veto && (hl == A_LANE_RIGHT || hl == A_LANE_LEFT) && espeed < 1.6667) v_safe = v_ll;
                        }
                        out_acc = (v_safe - evx) / dt;
                        changed = out_acc != ACUR[col + i];
                    }
                    if (!mass) { if (redo) ACUR[col + i] = out_acc; break; }
                    // publish after every lane has read the round's inputs
                    __syncwarp(hmask);
                    if (redo && changed) ACUR[col + i] = out_acc;
                    __syncwarp(hmask);
                    if (!(__ballot_sync(hmask, changed) & hmask) || ++rounds > CLANES) break;   // a chain is at most n long
                }
                if (redo) {
                    cadj_flag = cond_a >= -1e-6;
                    lc_safe = !veto;
                    if (DIAG) lc_margin = fmin(fmin(fabs(hls_a), fabs(cond_a)), fmin(fabs(hls_r), fabs(cond_r)));
                }
                // ---- a veto that changes the steering command invalidates the nominal move of that vehicle: the lowest such
                // rank re-steers (cbf.py:501-506, decentral_layer.py:737-744), everything behind it is evaluated again
                const bool same_cmd = fl_tlane(f) == elane && !(f & FL_CRASHED);
                const bool need = shielded && rank >= start && veto && !same_cmd && !resteered;
                const unsigned need_mask = __ballot_sync(hmask, need) & hmask;
                if (!need_mask) break;
                // lowest rank among the lanes that need it
                int first = 99;
                for (unsigned m = need_mask; m; m &= m - 1) {
                    const int l = (__ffs(m) - 1) & 15;
                    const int r = RANK[col + l];
                    first = r < first ? r : first;
                }
                if (need && rank == first) {
                    out_steer = steering_control(ex, ey, eh, espeed, elane);
                    g = move_geom(x, y, h, v, ch, sh, out_steer, dt);
                    NX[col + i] = g.nx; NY[col + i] = g.ny; NH[col + i] = g.nh; NCH[col + i] = g.ncos; NSH[col + i] = g.nsin;
                    NGV[col + i] = g.ngvx;
                    NFL[col + i] = fl_set(f, FL_LANE_SHIFT, FL_3BIT, (uint32_t)g.lane);
                    resteered = true;
                }
                start = first + 1;
                __syncwarp(hmask);
            }
            if (shielded) {
                // flags the shield leaves on the vehicle (decentral_layer.py:501-506, 737-764)
                if (mass) f = cadj_flag ? (f | FL_CADJ) : (f & ~FL_CADJ);
                if (veto) {
                    f = fl_set(f, FL_TLANE_SHIFT, FL_3BIT, (uint32_t)fl_lane(f));
                }
                f = constrain_adj ? (f | FL_COLLAB) : (f & ~FL_COLLAB);
                f = lc_safe ? (f | FL_LCSAFE) : (f & ~FL_LCSAFE);
                n_solves += 1; n_active += active != 0; n_veto += !lc_safe;
            }
        }
        // ---- commit the moves (kinematics.py:133-152, log_step) and publish the new state
        if (act_now) {
            const double acc_f = shielded ? out_acc : acc, steer_f = shielded ? out_steer : steer;
            safe_steer = steer_f;
            safe_acc = acc_f;
            const double nv = fmax(0.0, v + acc_f * dt);
            gvx = g.ngvx;
            f |= FL_FG;
            f = fl_set(f, FL_LANE_SHIFT, FL_3BIT, (uint32_t)g.lane);
            rec2x = x;
            rec2vx = rec1vx;
            rec1vx = nv * g.ncos;
            const int hist = fl_hist(f);
            if (hist < 2) f = fl_set(f, FL_HIST_SHIFT, 3u, (uint32_t)(hist + 1));
            if (DIAG) {
                const size_t plane = (size_t)p.n_envs * 3 * MAXV, idx = (e * 3 + (sub < 3 ? sub : 2)) * MAXV + i;
                int32_t *si = p.out.sh_i;
                double *sf = p.out.sh_f;
                if (shielded) {
                    si[idx] = 1; si[plane + idx] = id_ol; si[2 * plane + idx] = id_oa; si[3 * plane + idx] = id_oar;
                    si[4 * plane + idx] = constrain_adj; si[5 * plane + idx] = active; si[6 * plane + idx] = lc_safe;
                    sf[idx] = acc_f; sf[plane + idx] = steer_f; sf[2 * plane + idx] = acc; sf[3 * plane + idx] = steer;
                    sf[4 * plane + idx] = lc_margin;
                } else {
                    sf[idx] = acc_f; sf[plane + idx] = steer_f; sf[2 * plane + idx] = acc_f; sf[3 * plane + idx] = steer_f;
                }
                const int hl = fl_hl(f);
                si[7 * plane + idx] = 1; si[8 * plane + idx] = hl == A_NONE ? -1 : hl; si[9 * plane + idx] = g.lane;
                sf[5 * plane + idx] = g.nx; sf[6 * plane + idx] = g.ny; sf[7 * plane + idx] = g.nh; sf[8 * plane + idx] = nv;
                sf[9 * plane + idx] = minhw;
            }
        }
        __syncwarp();      // every lane has finished reading the sub-step-start planes
        if (act_now) {
            X(i) = g.nx; Y(i) = g.ny; H(i) = g.nh; V(i) = fmax(0.0, v + (shielded ? out_acc : acc) * dt);
            CH(i) = g.ncos; SH(i) = g.nsin;
            FL(i) = f;
        }
        __syncwarp();
        // ---- collisions (road.py:288-292): a lane looks for partners of its own vehicle within LENGTH (or the obstacle);
        // the rare env that has any runs the reference's ordered pass on its first lane
        bool close = false;
        if (act_now) {
            const double ax = X(i), ay = Y(i);
            {
                const double dx = OBST_X - ax, dy = OBST_Y - ay;
                close = !(dx * dx + dy * dy > VLEN_SQ_GT);
            }
            for (int j = 0; j < n; ++j) {
                if (j == i) continue;
                const double dx = X(j) - ax, dy = Y(j) - ay;
                close = close || !(dx * dx + dy * dy > VLEN_SQ_GT);
            }
        }
        const unsigned close_mask = (__ballot_sync(hmask, close) & hmask) >> (threadIdx.x & 16);
        if (close_mask && running && i == 0) collision_pass(ev, close_mask);
        __syncwarp();
        // _is_terminal (merge_env_v1.py:168-172): a lane looks at its own vehicle, a ballot gathers the env's answer
        const bool out_i = running && has_v && i < ev.n_cav && ((FL(i) & FL_CRASHED) || X(i) < 0);
        const bool any_out = (__ballot_sync(hmask, out_i) & hmask) != 0;
        if (running) {
            time = min(time + 1, (int)EI_TIME_MASK);
            if (steps >= p.cfg.duration_steps || any_out) running = false;   // abstract.py:530
        }
        __syncwarp();
    }

    // ---- write the state back: every field once per policy step
    if (has_v) {
#pragma unroll
        for (int fld = 0; fld < N_HOT; ++fld) GF(fld, i) = SMF(fld, i);
        p.st.flags[flags_index(e, i)] = FL(i);
        GF(F_GVX, i) = gvx; GF(F_REC1VX, i) = rec1vx; GF(F_REC2X, i) = rec2x; GF(F_REC2VX, i) = rec2vx;
        GF(F_ACT_STEER, i) = act_steer; GF(F_ACT_ACC, i) = act_acc;
        GF(F_SAFE_STEER, i) = safe_steer; GF(F_SAFE_ACC, i) = safe_acc; GF(F_MINHW, i) = minhw;
    }
    if (valid && i == 0)
        p.st.einfo[e] = (ei & 0xfffu) | ((uint32_t)steps << EI_STEPS_SHIFT) | ((uint32_t)time << EI_TIME_SHIFT);
    // shield counters: the 32 envs of a statistics row are spread over two CTAs here, hence atomics (integers: exact in any order)
    n_solves = __reduce_add_sync(0xffffffffu, n_solves);
    n_active = __reduce_add_sync(0xffffffffu, n_active);
    n_veto = __reduce_add_sync(0xffffffffu, n_veto);
    if ((threadIdx.x & 31) == 0 && n_solves) {
        double *row = p.out.stats + (((size_t)p.env_offset + (size_t)blockIdx.x * CENVS) >> 5) * N_STATS;
        atomicAdd(row + ST_SOLVES, (double)n_solves);
        atomicAdd(row + ST_ACTIVE, (double)n_active);
        atomicAdd(row + ST_VETOES, (double)n_veto);
    }
}

template <int SHIELD>
static void launch_coop_t(const StepParams &p, bool diag, void *stream) {
    static bool ready[MM_MAX_DEVICES] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MM_MAX_DEVICES) return;
    if (!ready[dev]) {
        cudaError_t ce = cudaFuncSetAttribute(coop_step_kernel<SHIELD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COOP_SMEM);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(coop_step_kernel<SHIELD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COOP_SMEM);
        if (ce != cudaSuccess) return;
        ready[dev] = true;
    }
    const int grid = (p.env_count + CENVS - 1) / CENVS;
    if (grid <= 0) return;
    if (diag) coop_step_kernel<SHIELD, true><<<grid, CTHREADS, COOP_SMEM, (cudaStream_t)stream>>>(p);
    else coop_step_kernel<SHIELD, false><<<grid, CTHREADS, COOP_SMEM, (cudaStream_t)stream>>>(p);
}

}  // namespace mmc

namespace mm {
void launch_step_coop(const StepParams &p, bool diag, void *stream) {
    if (p.cfg.shield == MM_SHIELD_MASS) mmc::launch_coop_t<MM_SHIELD_MASS>(p, diag, stream);
    else if (p.cfg.shield == MM_SHIELD_HSS) mmc::launch_coop_t<MM_SHIELD_HSS>(p, diag, stream);
    else mmc::launch_coop_t<MM_SHIELD_NONE>(p, diag, stream);
}
}  // namespace mm
