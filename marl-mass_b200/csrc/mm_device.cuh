// mm_device.cuh — device functions of the merge environment and its shields, shared by the kernels that include it
// (merge_step.cu: physics step kernel; merge_step_occ4.cu / merge_step_spec.cu: its other builds; merge_outputs.cu:
// observation / reward / info kernel).  The includer defines
//   MM_KNS   namespace of this copy (each translation unit gets its own, all `using namespace mm`)
//   MM_NHOT  f64 fields staged in shared memory (6: x y heading speed cos sin; 4: without cos / sin)
//   MM_PW    plane width = env columns per CTA (threads of the step kernel; 32 in the outputs kernel)
// Reference map: see the head of merge_step.cu.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "mm_internal.h"

// MM_SPEC (merge_step.cu, "Other builds"): 0 = every configuration; 1 = all-CAV envs of the plain LC env with the shield
// kind MM_SPEC_SHIELD, the configuration reads below folded to constants
#ifndef MM_SPEC
#define MM_SPEC 0
#endif
#if MM_SPEC
#define CFG_SHIELD(cfg) (MM_SPEC_SHIELD)
#define CFG_V0(cfg) false
#define CFG_STEER_VEL(cfg) false
#define CFG_ENV_HDV(cfg) false
#else
#define CFG_SHIELD(cfg) ((cfg).shield)
#define CFG_V0(cfg) ((cfg).env_v0 != 0)
#define CFG_STEER_VEL(cfg) ((cfg).steer_vel != 0)
#define CFG_ENV_HDV(cfg) ((cfg).env_hdv != 0)
#endif

namespace MM_KNS {
using namespace mm;

constexpr int N_HOT = MM_NHOT;       // fields F_X .. staged in shared memory during a step (6: including cos / sin heading)
constexpr int BLOCK = MM_PW;   // env columns per CTA
// build-time experiment knobs (profiles/README.md)
#ifndef MM_INL_A
#define MM_INL_A __forceinline__   // closest_lane: one hot call site
#endif
#ifndef MM_INL_B
#define MM_INL_B __forceinline__   // cav_act: measured -2 % inlined (profiles/README.md)
#endif
#ifndef MM_MIN_BLOCKS
#define MM_MIN_BLOCKS 3   // CTAs per SM the step kernel is compiled for (register budget 65536 / (128 * n))
#endif
constexpr double PI = 3.141592653589793;
constexpr double TWO_PI = 2 * PI;

enum { L_AB0 = 0, L_BC0 = 1, L_BC1 = 2, L_CD0 = 3, L_JK0 = 4, L_KB0 = 5, N_LANES = 6 };
enum { A_LANE_LEFT = 0, A_IDLE = 1, A_LANE_RIGHT = 2, A_FASTER = 3, A_SLOWER = 4, A_NONE = 7 };
constexpr int OBST = MAXV;  // entity id of the single obstacle at (420, 4)

__constant__ double c_lane_sx[N_LANES] = {0.0, 320.0, 320.0, 420.0, 0.0, 220.0};
__constant__ double c_lane_sy[N_LANES] = {0.0, 0.0, 4.0, 0.0, 10.5, 7.25};
__constant__ double c_lane_len[N_LANES] = {320.0, 100.0, 100.0, 1000.0, 220.0, 100.0};

constexpr double OBST_X = 420.0, OBST_Y = 4.0;
constexpr double VLEN = 5.0, VWID = 2.0, LWIDTH = 4.0;
constexpr double SINE_AMP = 3.25;
constexpr double SINE_PULS = 2 * PI / (2 * 100.0);
constexpr double SINE_PHASE = PI / 2;
constexpr double PERCEPTION = 180.0;
constexpr double KP_A = 1 / 0.6;
constexpr double KP_HEADING = 1 / 0.2;
constexpr double KP_LATERAL = 1.0 / 3 * KP_HEADING;
constexpr double PURSUIT_TAU = 0.5 * 0.2;
constexpr double MAX_STEER = PI / 3;
constexpr double ACC_LO = -12.5, ACC_HI = 6.0;
// sqrt is correctly rounded and monotone, so the reference's norm tests have exact squared-distance forms:
//   sqrt(q) < 180  <=>  q < 0x1.fa3ffffffffffp+14   (the smallest double whose sqrt rounds to 180.0)
//   sqrt(q) > 5    <=>  q > 0x1.9000000000001p+4    (the largest double whose sqrt rounds to 5.0)
constexpr double PERCEPTION_SQ_LT = 0x1.fa3ffffffffffp+14;
constexpr double VLEN_SQ_GT = 0x1.9000000000001p+4;

// ------------------------------------------------------------------------------------------------
// per-thread view of one environment
// ------------------------------------------------------------------------------------------------
// The staged planes are addressed through the dynamic shared-memory symbol itself so that every function, inlined
// or not, knows the address space (LDS/STS instead of generic LD/ST) and the per-thread view is a single index.
#ifndef MM_TMA
#define MM_TMA 1   // 1: hot planes move between HBM and shared memory as cp.async.bulk transactions; 0: per-thread loads
#endif
// Staged slots: an env never has more than 11 vehicles.  6 f64 planes x 11 slots + the flags = 72 KB per CTA, three
// CTAs per SM.
constexpr int SMV = 11;
constexpr int PLANES_F64 = N_HOT * SMV * BLOCK + (SMV + 1) * BLOCK / 2;   // doubles: hot f64 planes + the u32 flags plane
extern __shared__ __align__(16) double sm_planes[];   // [N_HOT][SMV][BLOCK] f64 + [SMV+1][BLOCK] u32
// MM_TPE (thread-per-env kernels: one CTA = one tile, thread t = env column t): the column index, the tile base and the
// shield's three configuration doubles are read from threadIdx / static shared memory instead of through `Env &` and
// `const mm_config &`.  Both live in the kernel's stack frame / parameter space; out-of-line functions reached them with
// local-memory and generic loads that miss the L1 (28 KB next to 221 KB of shared memory) - ncu: ~4 % of the step
// kernel's stall samples behind five such loads per vehicle step.
#ifdef MM_TPE
__shared__ double *s_tile_g;
__shared__ double s_cfgd[3];                  // dt, eta, tau
#define EV_TID ((int)threadIdx.x)
#define EV_G (s_tile_g + threadIdx.x)
#define CFG_DT(cfg) (s_cfgd[0])
#define CFG_ETA(cfg) (s_cfgd[1])
#define CFG_TAU(cfg) (s_cfgd[2])
#else
#define EV_TID (ev.tid)
#define EV_G (ev.g)
#define CFG_DT(cfg) ((cfg).dt)
#define CFG_ETA(cfg) ((cfg).eta)
#define CFG_TAU(cfg) ((cfg).tau)
#endif
struct Env {
    int tid;                    // threadIdx.x: column of this env inside the CTA's planes
    double *g;                  // this env's column of its tile; element (f, i) at [(f*MAXV+i)*TILE]
    int n_veh, n_cav;
    uint64_t live;              // slot ids by current x, descending, 4 bits each (kept sorted after every move)
    uint64_t pos;               // inverse: nibble i = position of slot i in `live`
};
__device__ __forceinline__ int nib(uint64_t w, int k) { return (int)((w >> (4 * k)) & 15ull); }

#define SMF(f, i) (sm_planes[((f) * SMV + (i)) * BLOCK + EV_TID])
#define X(i) SMF(F_X, i)
#define Y(i) SMF(F_Y, i)
#define H(i) SMF(F_H, i)
#define V(i) SMF(F_V, i)
#if MM_NHOT >= 6
#define CH(i) SMF(F_COSH, i)
#define SH(i) SMF(F_SINH, i)
#else
#define CH(i) GF(F_COSH, i)
#define SH(i) GF(F_SINH, i)
#endif
#define FL(i) (reinterpret_cast<uint32_t *>(sm_planes + N_HOT * SMV * BLOCK)[(i) * BLOCK + EV_TID])
#define GF(f, i) (*tile_ptr(EV_G, (f), (i)))

__device__ __forceinline__ double *tile_ptr(double *col, int f, int i) {
    double *q = col + ((f) * MAXV + (i)) * TILE;
    __builtin_assume(__isGlobal(q));   // lets out-of-line functions emit LDG/STG instead of generic accesses
    return q;
}
__device__ __forceinline__ int fl_kind(uint32_t f) { return f & FL_KIND_MASK; }
__device__ __forceinline__ bool is_cav(uint32_t f) { return MM_SPEC ? true : fl_kind(f) == MM_KIND_CAV; }
__device__ __forceinline__ int fl_lane(uint32_t f) { return (f >> FL_LANE_SHIFT) & FL_3BIT; }
__device__ __forceinline__ int fl_tlane(uint32_t f) { return (f >> FL_TLANE_SHIFT) & FL_3BIT; }
__device__ __forceinline__ int fl_hl(uint32_t f) { return (f >> FL_HL_SHIFT) & FL_3BIT; }
__device__ __forceinline__ int fl_hist(uint32_t f) { return (f >> FL_HIST_SHIFT) & 3u; }
__device__ __forceinline__ uint32_t fl_set(uint32_t f, uint32_t shift, uint32_t mask, uint32_t v) {
    return (f & ~(mask << shift)) | (v << shift);
}


// ------------------------------------------------------------------------------------------------
// double-precision libm entry points, one copy each.  Inlining them at every call site made the step kernel
// 233 KB of SASS and instruction-cache misses its top stall (profiles/r1_v1_*); as out-of-line functions
// the whole kernel is a fraction of that.  Same libdevice code, same bits.
// ------------------------------------------------------------------------------------------------
#ifndef MM_HDV_FN
#define MM_HDV_FN __noinline__   // the IDM / MOBIL helpers (generic builds only)
#endif
#ifndef MM_TRIG_FN
#define MM_TRIG_FN __noinline__
#endif
#ifndef MM_TRIG1_FN
#define MM_TRIG1_FN __noinline__
#endif
__device__ MM_TRIG_FN double m_sin(double x) { return sin(x); }
__device__ MM_TRIG_FN double m_cos(double x) { return cos(x); }
__device__ MM_TRIG_FN double2 m_sincos(double x) { double2 r; sincos(x, &r.x, &r.y); return r; }
__device__ MM_TRIG1_FN double m_tan(double x) { return tan(x); }
__device__ MM_TRIG1_FN double m_atan(double x) { return atan(x); }
__device__ MM_TRIG1_FN double m_asin(double x) { return asin(x); }
__device__ __noinline__ double m_exp(double x) { return exp(x); }
__device__ __noinline__ double m_log(double x) { return log(x); }
__device__ __noinline__ double m_pow(double x, double y) { return pow(x, y); }
__device__ __noinline__ double m_fmod(double x, double y) { return fmod(x, y); }

// ------------------------------------------------------------------------------------------------
// scalar helpers (utils.py:16-41)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double not_zero(double x) {
    if (fabs(x) > 1e-2) return x;
    return x > 0 ? 1e-2 : -1e-2;
}
__device__ __forceinline__ double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
// x / y for a finite y != 0 where x is often exactly zero (a vehicle on its lane's centre line, a zero steering angle).
// CUDA's f64 division checks its fast result against the normal range and otherwise calls a ~90-instruction slow path;
// a zero quotient fails that check, and with it every warp that held one such lane took the call (ncu: 4.1 % of the step
// kernel's instructions).  +-0 / y is +-0 with the product's sign, which x * y delivers; the division itself always sees
// a non-zero numerator.  Same bits as the plain division.
__device__ __forceinline__ double div_nz(double x, double y) {
    const bool z = x == 0.0;
    double num = z ? 1.0 : x;
    asm("" : "+d"(num));            // opaque: otherwise the compiler folds the two selects back into a plain x / y
    const double q = num / y;
    return z ? x * y : q;
}
// Python's floored float modulo by a positive modulus
__device__ __forceinline__ double pymod_pos(double a, double b) {
    if (a >= 0 && a < b) return a;  // fmod(a, b) == a exactly when 0 <= a < b: skip the (iterative) fmod
    double r = m_fmod(a, b);
    if (r < 0) r += b;
    return r;
}
__device__ __forceinline__ double wrap_to_pi(double x) { return pymod_pos(x + PI, TWO_PI) - PI; }
__device__ __forceinline__ double lmap(double v, double x0, double x1, double y0, double y1) {
    return y0 + (v - x0) * (y1 - y0) / (x1 - x0);
}

// ------------------------------------------------------------------------------------------------
// lane algebra: every lane is x-aligned, so s = x - start.x and r = y - start.y (- sine offset on kb0)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double lane_s(int lane, double px) { return px - c_lane_sx[lane]; }
__device__ __forceinline__ double lane_r(int lane, double s, double py) {
    double r = py - c_lane_sy[lane];
    if (lane == L_KB0) r = r - SINE_AMP * m_sin(SINE_PULS * s + SINE_PHASE);
    return r;
}
__device__ __forceinline__ double lane_heading_at(int lane, double s) {
    if (lane == L_KB0) return m_atan(SINE_AMP * SINE_PULS * m_cos(SINE_PULS * s + SINE_PHASE));
    return 0.0;
}
__device__ __forceinline__ bool on_lane(int lane, double px, double py, double margin) {
    double s = lane_s(lane, px);
    double r = lane_r(lane, s, py);
    return fabs(r) <= LWIDTH / 2 + margin && -VLEN <= s && s < c_lane_len[lane] + VLEN;
}
// road.py:67-109 with route=None (every node has a single successor).  From ab0 / kb0 the next road has two
// lanes and the closer one (lane.py:97-100 distance, first minimum) is taken; bc0 and bc1 share start.x and
// length, so the two distances differ only in |r|.
__device__ __forceinline__ int next_lane(int lane, double px, double py) {
    if (lane == L_AB0 || lane == L_KB0) {
        double s = px - 320.0;
        double over = fmax(s - 100.0, 0.0), under = fmax(0.0 - s, 0.0);
        return fabs(py - 0.0) + over + under <= fabs(py - 4.0) + over + under ? L_BC0 : L_BC1;
    }
    if (lane == L_JK0) return L_KB0;
    return L_CD0;
}
__device__ __forceinline__ int lane_road(int lane) { return lane == L_BC1 ? L_BC0 : lane; }
__device__ __forceinline__ int lane_rid(int lane) { return lane == L_BC1 ? 1 : 0; }

// road.py:51-65 + lane.py:102-108: first minimum over [ab0, bc0, bc1, cd0, jk0, kb0].
// The five straight lanes share heading_at == 0, so one wrap_to_pi serves them; kb0 (last in argmin order)
// is evaluated only if its heading-free lower bound can still beat the incumbent.
__device__ MM_INL_A int closest_lane(double px, double py, double heading) {
    double ang0 = fabs(wrap_to_pi(heading - 0.0));
    int best = 0;
    double bd = CUDART_INF;
#pragma unroll
    for (int l = 0; l < 5; ++l) {
        double s = lane_s(l, px);
        double r = py - c_lane_sy[l];
        double d = fabs(r) + fmax(s - c_lane_len[l], 0.0) + fmax(0.0 - s, 0.0) + 1.0 * ang0;
        if (d < bd) { bd = d; best = l; }
    }
    double s = lane_s(L_KB0, px);
    double along = fmax(s - c_lane_len[L_KB0], 0.0) + fmax(0.0 - s, 0.0);
    // |r| >= 0 and angle >= 0 and fp addition is monotone, so d >= along; moreover |r| >= |y - 7.25| - 3.25 (the
    // sine offset is at most the amplitude), which rules kb0 out for main-road vehicles driving alongside the ramp
    // without evaluating the sine lane (1e-9 m of slack against the ~1e-15 rounding of the exact expression)
    if (along < bd && along + (fabs(py - c_lane_sy[L_KB0]) - SINE_AMP) - 1e-9 < bd) {
        // lane_r and lane_heading_at of the sine lane at the same s: one sincos serves both
        double2 sc = m_sincos(SINE_PULS * s + SINE_PHASE);
        double r = (py - c_lane_sy[L_KB0]) - SINE_AMP * sc.x;
        double ang = fabs(wrap_to_pi(heading - m_atan(SINE_AMP * SINE_PULS * sc.y)));
        double d = fabs(r) + fmax(s - c_lane_len[L_KB0], 0.0) + fmax(0.0 - s, 0.0) + 1.0 * ang;
        if (d < bd) best = L_KB0;
    }
    return best;
}

// controller.py:146-187
#ifndef MM_STEER_FN
#define MM_STEER_FN __noinline__
#endif
__device__ MM_STEER_FN double steering_control(double px, double py, double heading, double speed, int tlane) {
    double s = lane_s(tlane, px);
    double r = lane_r(tlane, s, py);
    double future_heading = lane_heading_at(tlane, s + speed * PURSUIT_TAU);
    double lat_cmd = -KP_LATERAL * r;
    double nz = not_zero(speed);
    double heading_cmd = m_asin(clipd(div_nz(lat_cmd, nz), -1.0, 1.0));
    double heading_ref = future_heading + clipd(heading_cmd, -PI / 4, PI / 4);
    double rate_cmd = KP_HEADING * wrap_to_pi(heading_ref - heading);
    double steering = m_asin(clipd(VLEN / 2 / nz * rate_cmd, -1.0, 1.0));
    return clipd(steering, -MAX_STEER, MAX_STEER);
}

__device__ __forceinline__ int speed_to_index(double speed) {
    double x = (speed - 10.0) / (30.0 - 10.0);
    return (int)clipd(rint(x * 4), 0.0, 4.0);
}

// controller.py:136-144
__device__ __forceinline__ int follow_road(int tlane, double px, double py) {
    double s = lane_s(tlane, px);
    if (s > c_lane_len[tlane] - VLEN / 2) return next_lane(tlane, px, py);
    return tlane;
}

// MDPLCVehicle.act -> MDPVehicle.act -> ControlledVehicle.act (safe_controller.py:63-66, controller.py:293-311, 90-134)
__device__ MM_INL_B void cav_act(Env &ev, int i, int action, bool steer_vel, double &steer, double &acc) {
    uint32_t f = FL(i);
    double px = X(i), py = Y(i), speed = V(i);
    if (action != A_NONE) f = fl_set(f, FL_HL_SHIFT, FL_3BIT, (uint32_t)action);
    if (action == A_FASTER || action == A_SLOWER) {
        int idx = speed_to_index(speed) + (action == A_FASTER ? 1 : -1);
        idx = min(max(idx, 0), 4);
        f = fl_set(f, FL_SIDX_SHIFT, FL_3BIT, (uint32_t)idx);
        GF(F_TSPEED, i) = 10.0 + idx * (30.0 - 10.0) / 4;
    }
    int tl = follow_road(fl_tlane(f), px, py);
    // the only reachable, non-forbidden side lane of the network is bc0 seen from bc1 (LANE_LEFT)
    if (action == A_LANE_LEFT && tl == L_BC1) {
        double s = lane_s(L_BC0, px), r = py - c_lane_sy[L_BC0];
        if (fabs(r) <= 2 * LWIDTH && 0 <= s && s < c_lane_len[L_BC0] + VLEN) tl = L_BC0;
    }
    f = fl_set(f, FL_TLANE_SHIFT, FL_3BIT, (uint32_t)tl);
    FL(i) = f;
    steer = steering_control(px, py, H(i), speed, tl);
    // MDPLCVehicle.steering_control in steer_vel mode: a steering velocity towards 1/8 of the reference angle
    // (safe_controller.py:93-96), clipped like any steering command in ControlledVehicle.act (controller.py:131-133)
    if (steer_vel) steer = clipd(20 * (steer * 0.125 - GF(F_STEERANG, i)), -MAX_STEER, MAX_STEER);
    // a CAV's target_speed is always index_to_speed(speed_index) (controller.py:281-283, 302-307): same double as the
    // stored field, without the L2 round trip
    acc = KP_A * ((10.0 + (int)((f >> FL_SIDX_SHIFT) & FL_3BIT) * (30.0 - 10.0) / 4) - speed);
}

// ------------------------------------------------------------------------------------------------
// HDV: IDM + MOBIL (behavior.py)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ent_pos(const Env &ev, int id, double &px, double &py) {
    if (id == OBST) { px = OBST_X; py = OBST_Y; } else { px = X(id); py = Y(id); }
}

// road.py:352-381 (candidates: vehicles in list order, then the obstacle): the exhaustive scan, kept for the tie cases
__device__ MM_HDV_FN void neighbour_vehicles_scan(const Env &ev, int self, int lane, int &front, int &rear) {
    double s = lane_s(lane, X(self)), s_front = 0, s_rear = 0;
    front = -1;
    rear = -1;
    for (int j = 0; j <= ev.n_veh; ++j) {
        int id = j == ev.n_veh ? OBST : j;
        if (id == self) continue;
        double px, py;
        ent_pos(ev, id, px, py);
        double s_v = lane_s(lane, px);
        if (!(-VLEN <= s_v && s_v < c_lane_len[lane] + VLEN)) continue;
        double lat = lane_r(lane, s_v, py);
        if (!(fabs(lat) <= LWIDTH / 2 + 1)) continue;
        if (s <= s_v && (front < 0 || s_v <= s_front)) { s_front = s_v; front = id; }
        if (s_v < s && (rear < 0 || s_v > s_rear)) { s_rear = s_v; rear = id; }
    }
}

// The same through the x-sorted order: lane_s is monotone in x, so the front vehicle is the first one ahead of the ego
// that lies on the lane (margin 1 m), the rear vehicle the first one behind.  The obstacle at (420, 4) is on bc1 only
// (s = 100, r = 0; every other lane fails on_lane), where it competes by its s like a vehicle and, being last in the
// candidate list, wins ties for the front place.  Equal longitudinal positions between vehicles are where list order
// decides: any such equality on the way defers to the scan.
__device__ MM_HDV_FN void neighbour_vehicles(const Env &ev, int self, int lane, int &front, int &rear) {
    const uint64_t live = ev.live;
    const int ps = nib(ev.pos, self);
    const double s = lane_s(lane, X(self)), hi = c_lane_len[lane] + VLEN;
    front = -1;
    rear = -1;
    bool tie = false;
    double s_front = 0, s_rear = 0, prev = s;
    for (int p = ps - 1; p >= 0; --p) {               // ahead: s_v grows
        const int j = nib(live, p);
        const double s_v = lane_s(lane, X(j));
        tie |= s_v == prev;
        prev = s_v;
        if (front >= 0) break;                        // one look past the hit, for the tie test only
        if (!(s_v < hi)) break;
        if (!(-VLEN <= s_v)) continue;
        if (!(fabs(lane_r(lane, s_v, Y(j))) <= LWIDTH / 2 + 1)) continue;
        s_front = s_v; front = j;
    }
    prev = s;
    for (int p = ps + 1; p < ev.n_veh; ++p) {         // behind: s_v shrinks
        const int j = nib(live, p);
        const double s_v = lane_s(lane, X(j));
        tie |= s_v == prev;
        prev = s_v;
        if (rear >= 0) break;
        if (!(-VLEN <= s_v)) break;
        if (!(s_v < hi)) continue;
        if (!(fabs(lane_r(lane, s_v, Y(j))) <= LWIDTH / 2 + 1)) continue;
        s_rear = s_v; rear = j;
    }
    if (tie) { neighbour_vehicles_scan(ev, self, lane, front, rear); return; }
    if (lane == L_BC1) {
        const double s_o = OBST_X - 320.0;            // lane_s(bc1, 420)
        if (s <= s_o && (front < 0 || s_o <= s_front)) front = OBST;
        if (s_o < s && (rear < 0 || s_o > s_rear)) rear = OBST;
    }
}

// behavior.py:141-156
__device__ MM_HDV_FN double desired_gap(const Env &ev, int ego, int front) {
    double fvx = 0, fvy = 0;
    if (front != OBST) {
        fvx = V(front) * CH(front);
        fvy = V(front) * SH(front);
    }
    double ec = CH(ego), es = SH(ego), speed = V(ego);
    double dv = (speed * ec - fvx) * ec + (speed * es - fvy) * es;
    return 10.0 + speed * 1.5 + speed * dv / (2 * sqrt(15.0));
}

// behavior.py:111-139 (ego/front: -1 None, OBST obstacle)
__device__ MM_HDV_FN double idm_acc(const Env &ev, int ego, int front) {
    if (ego < 0 || ego == OBST) return 0.0;
    double acc = 3.0 * (1 - m_pow(fmax(V(ego), 0.0) / not_zero(GF(F_TSPEED, ego)), 4.0));
    if (front >= 0) {
        double fx, fy;
        ent_pos(ev, front, fx, fy);
        int el = fl_lane(FL(ego));
        double d = lane_s(el, fx) - lane_s(el, X(ego));
        double q = desired_gap(ev, ego, front) / not_zero(d);
        acc -= 3.0 * (q * q);
    }
    return acc;
}

// behavior.py:186-266 (route None, POLITENESS 0: the follower terms enter the jerk with weight 0.0)
__device__ MM_HDV_FN int hdv_change_lane(Env &ev, int self, uint32_t f, int tl) {
    int lane = fl_lane(f);
    if (lane != tl) {
        if (lane_road(lane) == lane_road(tl)) {
            for (int j = 0; j < ev.n_veh; ++j) {
                uint32_t fj = FL(j);
                if (j != self && fl_lane(fj) != tl && fl_tlane(fj) == tl) {
                    double d = lane_s(lane, X(j)) - lane_s(lane, X(self));
                    if (0 < d && d < desired_gap(ev, self, j)) return lane;
                }
            }
        }
        return tl;
    }
    double timer = GF(F_TIMER, self);
    if (!(1.0 < timer)) return tl;
    GF(F_TIMER, self) = 0.0;
    if (lane != L_BC1) return tl;  // bc0's side lane bc1 is forbidden; nothing else has side lanes
    {
        double s = lane_s(L_BC0, X(self)), r = Y(self) - c_lane_sy[L_BC0];
        if (!(fabs(r) <= 2 * LWIDTH && 0 <= s && s < c_lane_len[L_BC0] + VLEN)) return tl;
    }
    int new_prec, new_foll, old_prec, old_foll;
    neighbour_vehicles(ev, self, L_BC0, new_prec, new_foll);
    double nf_a = idm_acc(ev, new_foll, new_prec);
    double nf_pred = idm_acc(ev, new_foll, self);
    if (nf_pred < -9.0) return tl;
    neighbour_vehicles(ev, self, lane, old_prec, old_foll);
    double self_pred = idm_acc(ev, self, new_prec);
    double self_a = idm_acc(ev, self, old_prec);
    double of_a = idm_acc(ev, old_foll, self);
    double of_pred = idm_acc(ev, old_foll, old_prec);
    double jerk = self_pred - self_a + 0. * (nf_pred - nf_a + of_pred - of_a);
    if (jerk < 0.1) return tl;
    return L_BC0;
}

// behavior.py:74-100
__device__ MM_HDV_FN void hdv_act(Env &ev, int i) {
    uint32_t f = FL(i);
    if (f & FL_CRASHED) return;
    int front, rear;
    neighbour_vehicles(ev, i, fl_lane(f), front, rear);
    int tl = follow_road(fl_tlane(f), X(i), Y(i));
    tl = hdv_change_lane(ev, i, f, tl);
    FL(i) = fl_set(f, FL_TLANE_SHIFT, FL_3BIT, (uint32_t)tl);
    GF(F_ACT_STEER, i) = steering_control(X(i), Y(i), H(i), V(i), tl);
    GF(F_ACT_ACC, i) = clipd(idm_acc(ev, i, front), -6.0, 6.0);
}

// ------------------------------------------------------------------------------------------------
// shields
// ------------------------------------------------------------------------------------------------
// Closed-form minimiser of the reference's QP (cbf.py:110-135 with rows cbf.py:288-322 / 374-422):
//   min 1/2 (u0^2 + u1^2 + 1e18 s^2)  s.t.  a u0 - s <= c_lead, [a u0 - s <= c_adj,]  lo <= u0 <= hi.
__device__ __forceinline__ double solve_cbf_qp(double a, double c_lead, double c_adj, bool has_adj, double lo,
                                               double hi, int &active) {
    double u = 0.0;
    active = 0;
    if (u > hi) { u = hi; active = MM_ACT_UPPER; }
    if (u < lo) { u = lo; active = MM_ACT_LOWER; }
    if (a > 0.0) {
        double lim = c_lead / a;
        int row = MM_ACT_LEAD;
        if (has_adj) {
            double la = c_adj / a;
            if (la < lim) { lim = la; row = MM_ACT_ADJ; }
        }
        if (u > lim) {
            if (lim >= lo) { u = lim; active = row; }
            else { u = lo; active = row | MM_ACT_LOWER | MM_ACT_SLACK; }
        }
    } else if (a < 0.0) {
        double lim = c_lead / a;
        int row = MM_ACT_LEAD;
        if (has_adj) {
            double la = c_adj / a;
            if (la > lim) { lim = la; row = MM_ACT_ADJ; }
        }
        if (u < lim) {
            if (lim <= hi) { u = lim; active = row; }
            else { u = hi; active = row | MM_ACT_UPPER | MM_ACT_SLACK; }
        }
    } else {
        double c = has_adj ? fmin(c_lead, c_adj) : c_lead;
        if (c < 0.0) active |= MM_ACT_SLACK | MM_ACT_LEAD;
    }
    return u;
}

// controller.py:257-267; left: dir == "L".  cos/sin(+-alpha + heading) by angle addition from the cached cos/sin of
// the heading (alpha = atan(WIDTH / LENGTH) = atan(0.4): cos = 5/sqrt(29), sin = 2/sqrt(29)).
__device__ __forceinline__ void get_corner(double px, double py, double ch, double sh, bool left, double &cx, double &cy) {
    const double corner_len = 2.700082417423684;   // sqrt(1^2 + 2.5^2) + 0.0075
    const double ca = 0.9284766908852593, sa = 0.3713906763541037;
    double sin_pa = sa * ch + ca * sh;             // sin(alpha + heading)
    double sin_ma = ca * sh - sa * ch;             // sin(-alpha + heading)
    cx = px + (corner_len * (ca * ch - sa * sh));
    cy = py - (corner_len * (left ? sin_pa : sin_ma)) + 0.01;
}

struct ShieldRec {
    int leader, front_adj, rear_adj, constrain_adj, active, is_lc_safe;
    double lc_margin, min_headway;
};

// road.py:257-267: the `count` nearest (by |longitudinal offset in the ego lane|, stable) among vehicles
// closer than 180 m, as packed ids (4 bits each, nearest first) and their number (<= K).
//
// Every lane is x-aligned, so the key |lane_s(el, ox) - es| never decreases along either side of the ego in the
// x-sorted order `ev.live` (floating-point subtraction is monotone): the K nearest come out of a two-sided walk from
// the ego's position, a merge by key, instead of a scan of all vehicles with a top-K insertion.  sorted() is stable,
// i.e. equal keys are ordered by slot id; two equal keys next to each other in the merge (or at the K-th boundary)
// are the only case the walk cannot order, and it then defers to the scan below.
template <int K>
__device__ __noinline__ int close_vehicles_scan(const Env &ev, int self, uint32_t &packed_ids) {
    const double ex = X(self), ey = Y(self);
    const int el = fl_lane(FL(self));
    const double es = lane_s(el, ex);
    double last_key = -1.0;
    int last_id = -1, n = 0;
    uint32_t ids = 0;
    for (int k = 0; k < K; ++k) {   // k-th smallest (key, slot id) by selection: rare path, no scratch memory
        double best_key = CUDART_INF;
        int best = -1;
        for (int j = 0; j < ev.n_veh; ++j) {
            if (j == self) continue;
            double ox = X(j), dx = ox - ex, dy = Y(j) - ey;
            if (!(dx * dx + dy * dy < PERCEPTION_SQ_LT)) continue;  // np.linalg.norm(...) < 180
            double key = fabs(lane_s(el, ox) - es);
            if (key < last_key || (key == last_key && j <= last_id)) continue;  // already emitted
            if (key < best_key) { best_key = key; best = j; }
        }
        if (best < 0) break;
        ids |= (uint32_t)best << (4 * k);
        last_key = best_key; last_id = best;
        ++n;
    }
    packed_ids = ids;
    return n;
}

template <int K>
__device__ __forceinline__ int close_vehicles(const Env &ev, int self, uint32_t &packed_ids) {
    const double ex = X(self), ey = Y(self);
    const int el = fl_lane(FL(self));
    const double es = lane_s(el, ex);
    const uint64_t live = ev.live;
    const int n_veh = ev.n_veh;
    // cursor of each side: next position to look at (front: towards larger x); -1 / n_veh = exhausted
    int pf = nib(ev.pos, self) - 1, pr = pf + 2;
    double kf = CUDART_INF, kr = CUDART_INF;
    int idf = 0, idr = 0;
    // next vehicle of one side inside the perception radius: its key and id (key = inf when the side is exhausted)
    auto advance = [&](int pc, const int dir, double &key, int &id) -> int {
        key = CUDART_INF;
        while (pc >= 0 && pc < n_veh) {
            int j = nib(live, pc);
            pc += dir;
            double ox = X(j), dx = ox - ex, dy = Y(j) - ey, dx2 = dx * dx;
            if (!(dx2 < PERCEPTION_SQ_LT)) return -1;                // farther ones on this side fail as well
            if (!(dx2 + dy * dy < PERCEPTION_SQ_LT)) continue;       // np.linalg.norm(...) < 180
            key = fabs(lane_s(el, ox) - es);
            id = j;
            break;
        }
        return pc;
    };
    pf = advance(pf, -1, kf, idf);
    pr = advance(pr, +1, kr, idr);
    int n = 0;
    uint32_t ids = 0;
    double last = -1.0;
    bool tie = false;
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        const bool front = kf <= kr;
        const double key = front ? kf : kr;
        if (key == CUDART_INF) break;   // both sides exhausted
        tie |= key == last;
        last = key;
        ids |= (uint32_t)(front ? idf : idr) << (4 * k);
        ++n;
        // refill the side that was consumed; one copy of the loop serves both sides (lanes stay converged)
        double nk;
        int nid = 0;
        int pc = advance(front ? pf : pr, front ? -1 : +1, nk, nid);
        if (front) { pf = pc; kf = nk; idf = nid; } else { pr = pc; kr = nk; idr = nid; }
    }
    tie |= n == K && fmin(kf, kr) == last;
    if (tie) return close_vehicles_scan<K>(ev, self, packed_ids);
    packed_ids = ids;
    return n;
}

// safety_layer -> safe_action_hss / safe_action_mass (decentral_layer.py:767-817, 290-518, 521-764) with
// multi_agent_state (85-257) and CBF_AV / CBF_CAV (cbf.py:197-430).  Returns the shielded (steer, acc).
// `act_steer`, `act_acc`: the clipped nominal action; `rec1vx`, `ge`: the ego's last logged vx and fg g.vx (the caller
// has them in registers already).  WITH_MARGIN: also report the veto margin (diagnostic builds only).
// QUERY: evaluate only (mm_shield_query) - nothing outside the caller's shared-memory copy of the scene is written: the
// min_headway field and the in-place shift of an on-ramp HDV's record stay local.
#ifndef MM_SHIELD_FN
#define MM_SHIELD_FN __noinline__
#endif
template <bool WITH_MARGIN, bool QUERY = false>
__device__ MM_SHIELD_FN void shield(Env &ev, const mm_config &cfg, int self, double act_steer, double act_acc, double rec1vx,
                                    double ge, double &out_steer, double &out_acc, ShieldRec &rec) {
    const double dt = CFG_DT(cfg), eta = CFG_ETA(cfg), tau = CFG_TAU(cfg);
    const bool mass = CFG_SHIELD(cfg) == MM_SHIELD_MASS;
    // read the tile base now, so that the record fetch below does not start with a load in front of its address
    // arithmetic (ncu: 1.4 % of the kernel's stall samples at that one wait)
    double *const gbase = EV_G;
#define GQ(f, i) (*tile_ptr(gbase, (f), (i)))
    uint32_t f = FL(self);
    const int elane = fl_lane(f);
    const double ex = X(self), ey = Y(self), eh = H(self), espeed = V(self);

    double v_min = espeed + ACC_LO * dt;
    if (mass) v_min = fmax(0.0, v_min);
    double v_max = espeed + ACC_HI * dt;
    // to_dict()["vx"] = speed * cos(heading): for a vehicle that has not crashed this is bit-for-bit the value its
    // last log_step recorded (same operands), so the record is reused instead of a cosine
    double evx_raw = (f & FL_CRASHED) ? espeed * CH(self) : rec1vx;
    double evx = evx_raw > 1 ? evx_raw : 1;
    double es = lane_s(elane, ex);

    // multi_agent_state (decentral_layer.py:85-257) in two passes: the walk over the <= 5 nearest vehicles only decides
    // WHO is the rear-adjacent / front-adjacent / leading vehicle (shared-memory reads and integer logic); the records
    // of those three are then fetched together, all lanes converged, with one exposed L2 latency instead of one per role
    // and per loop iteration.  Nothing the fetch reads changes in between.
    int id_ol = MM_NB_NONE, id_oa = MM_NB_NONE, id_oar = MM_NB_NONE;
    bool oa_left = false, oa_onramp = false;
    double x_onramp = 0, vx_onramp = 0;

    uint32_t nb_ids;
    int n_nb = close_vehicles<5>(ev, self, nb_ids);
    const int e_next = next_lane(elane, ex, ey);
    // is_adj_lane(ego, other) (decentral_layer.py:23-39) can only be non-zero through the two-lane b->c road: the ego
    // side is its own lane when that is bc0/bc1, else its next lane
    const bool elane_bc = (elane == L_BC0) | (elane == L_BC1);
    const int e_eff = elane_bc ? elane : e_next;
    const bool e_eff_bc = (e_eff == L_BC0) | (e_eff == L_BC1);
    const bool all_cav = MM_SPEC ? true : ev.n_cav == ev.n_veh;
#pragma unroll 1
    for (int k = 0; k < n_nb; ++k) {
        int o = (int)((nb_ids >> (4 * k)) & 15u);
        double ox = X(o);
        double d = lane_s(elane, ox) - es;
        // A vehicle behind can only become the rear-adjacent one, a CAV ahead only front-adjacent or leader: skip
        // the classification when those roles are taken (HDVs ahead always run it: the on-ramp branch has side effects)
        if (d < 0 ? id_oar != MM_NB_NONE : (all_cav && id_oa != MM_NB_NONE && id_ol != MM_NB_NONE)) continue;
        uint32_t fo = FL(o);
        int olane = fl_lane(fo);
        double oy = Y(o);
        const bool olane_bc = (olane == L_BC0) | (olane == L_BC1);
        int v_a = (e_eff_bc & olane_bc & (e_eff != olane)) ? (e_eff == L_BC1 ? 1 : -1) : 0;
        int a_v = 0;
        if (elane_bc) {   // the other side: its lane if bc0/bc1, else its next lane (only ab0 / kb0 lead into b->c)
            int o_eff = olane;
            if (olane == L_AB0 || olane == L_KB0) o_eff = next_lane(olane, ox, oy);
            bool o_eff_bc = (o_eff == L_BC0) | (o_eff == L_BC1);
            a_v = (o_eff_bc & (o_eff != elane)) ? (o_eff == L_BC1 ? 1 : -1) : 0;
        }
        // is_approaching_same_lane (decentral_layer.py:46-57)
        bool approaching = false;
        if (!(d < 0)) {
            double y_dist = oy - ey, oh = H(o);
            bool hc = y_dist < 0 ? (oh > 0.037) : (oh < -0.037);
            approaching = fabs(y_dist) <= 3.5 && hc;
        }
        if (!approaching && (v_a != 0 || a_v != 0)) {
            if (id_oar == MM_NB_NONE && d < 0) {
                id_oar = o;                                   // rear-adjacent
            } else if (id_oa == MM_NB_NONE && d >= 0) {
                id_oa = o;                                    // front-adjacent
                oa_left = (v_a == -1 || a_v == 1);
            }
        } else if (!is_cav(fo) && elane == L_AB0 && olane == L_KB0 && d >= 0) {
            // on-ramp HDV: its record is shifted IN PLACE by half a second of ego speed (decentral_layer.py:175-184);
            // it takes the front-adjacent role whoever held it
            x_onramp = GF(F_REC2X, o) + 0.5 * evx_raw;
            if (!QUERY) GF(F_REC2X, o) = x_onramp;
            vx_onramp = GF(F_REC2VX, o);
            id_oa = o;
            oa_onramp = true;
        } else if (id_ol == MM_NB_NONE && d > 0) {
            if ((elane == olane) || (olane == e_next) || approaching) id_ol = o;   // leader
        }
    }
    const bool has_ol0 = id_ol != MM_NB_NONE, has_oa0 = id_oa != MM_NB_NONE, has_oar0 = id_oar != MM_NB_NONE;
    double x_ol = ex + PERCEPTION + 1, x_oa = ex + PERCEPTION + 1, x_oar = ex - PERCEPTION - 1;
    double vx_ol = 0, vx_oa = 0, vx_oar = 0, a_ol = 0, a_oa = 0, g_ol = 0, g_oa = 0;
    bool constrain_adj = false;
    {
        // the three fetches, issued back to back (an absent role reads the ego's own column and is discarded)
        const int jl = has_ol0 ? id_ol : self, ja = has_oa0 ? id_oa : self, jr = has_oar0 ? id_oar : self;
        const double l_x = GQ(F_REC2X, jl), l_vx = GQ(F_REC2VX, jl);            // leader: record before its last step
        const double a_x = GQ(F_REC2X, ja), a_vx = GQ(F_REC2VX, ja);            // front-adjacent: same
        const double r_vx = GQ(F_REC1VX, jr);                                    // rear-adjacent: current state
        double l_a = 0, l_g = 0, a_a = 0, a_g = 0, a_ch = 0, a_sh = 0;
        if (mass) {
            l_a = GQ(F_SAFE_ACC, jl); l_g = GQ(F_GVX, jl);
            a_a = GQ(F_SAFE_ACC, ja); a_g = GQ(F_GVX, ja);
            a_ch = CH(ja); a_sh = SH(ja);
        }
        if (has_oar0) {
            uint32_t fo = FL(jr);
            x_oar = X(jr);
            vx_oar = (fo & FL_CRASHED) ? V(jr) * CH(jr) : r_vx;
        }
        if (has_ol0) {
            x_ol = l_x; vx_ol = l_vx;
            if (mass) {
                bool o_cav = is_cav(FL(jl));
                a_ol = o_cav ? l_a : ACC_LO;
                g_ol = o_cav ? l_g : 1.0;
            }
        }
        if (oa_onramp) {
            x_oa = x_onramp; vx_oa = vx_onramp;
            constrain_adj = true;
            a_oa = ACC_LO;
            g_oa = 1.0;
        } else if (has_oa0) {
            x_oa = a_x; vx_oa = a_vx;
            if (mass) {
                uint32_t fo = FL(ja);
                bool o_cav = is_cav(fo);
                a_oa = o_cav ? a_a : ACC_LO;
                g_oa = o_cav ? a_g : 1.0;
                double cx, cy;
                get_corner(X(ja), Y(ja), a_ch, a_sh, oa_left, cx, cy);
                constrain_adj = !on_lane(fl_lane(fo), cx, cy, 0.0);
            }
        }
    }
    bool has_ol = has_ol0, has_oa = has_oa0, has_oar = has_oar0;
    // the obstacle can take over either role (decentral_layer.py:213-246)
    if (!(ex > OBST_X)) {
        double ady = fabs(OBST_Y - ey);
        if ((!has_ol || OBST_X <= x_ol) && ady <= 2) {
            has_ol = true; id_ol = MM_NB_OBSTACLE; x_ol = OBST_X; vx_ol = 0;
            if (mass) { a_ol = 0; g_ol = 0; }
        }
        if ((!has_oa || OBST_X <= x_oa) && 2 < ady && ady <= 4) {
            has_oa = true; id_oa = MM_NB_OBSTACLE; x_oa = OBST_X; vx_oa = 0;
            if (mass) { a_oa = 0; g_oa = 0; constrain_adj = false; }
        }
    }
    if (!mass) { g_ol = 1.0; g_oa = 1.0; }

    // safe distances and headway (decentral_layer.py:448-468)
    double sv_oar = (has_oar ? vx_oar : 0.0) + ACC_HI * dt;
    sv_oar = sv_oar > 1 ? sv_oar : 1;
    double buffer = (ACC_HI + 0.1) * dt * tau;
    double sd_l = evx * tau + VLEN + buffer;
    double sd_r = sv_oar * tau + VLEN + buffer;
    rec.min_headway = (x_ol - ex - VLEN) / evx;
    if (!QUERY) GF(F_MINHW, self) = rec.min_headway;

    // one-step predictions (decentral_layer.py:60-77)
    double v_ll = fmax(0.0, evx + act_acc * dt);
    double v_ol = has_ol ? fmax(0.0, vx_ol + (mass ? a_ol : ACC_LO) * dt) : 0.0;
    double v_oa = has_oa ? fmax(0.0, vx_oa + (mass ? a_oa : ACC_LO) * dt) : 0.0;
    double v_oar = has_oar ? fmax(0.0, vx_oar + ACC_HI * dt) : 0.0;

    double q_lon = -VLEN - sd_l;
    double q_lona = -VLEN - sd_l;
    bool has_adj = mass && constrain_adj;
    if (has_adj) q_lona = -VLEN - sd_l - 2.0134;
    double q_lonr = -VLEN - sd_r;

    double ge_dt = ge * dt, gol_dt = g_ol * dt, goa_dt = g_oa * dt, gr_dt = 1 * dt;
    double dl = -ex + x_ol, da = -ex + x_oa, dr = ex + -x_oar;
    double c_lead = dl + (eta - 1) * dl + eta * q_lon + (-(ge_dt * v_ll) + gol_dt * v_ol);
    double c_adj = 0.0;
    if (has_adj) c_adj = da + (eta - 1) * da + eta * q_lona + (-(ge_dt * v_ll) + goa_dt * v_oa);
    double hi = v_max - v_ll;
    double lo = -(-v_min + v_ll);
    int active;
    double u = solve_cbf_qp(ge_dt, c_lead, c_adj, has_adj, lo, hi, active);
    double v_safe = v_ll + u;

    // lane-change veto (cbf.py:324-339)
    double hls_a = da + q_lona;
    double hlds_a = da + (-ge_dt * v_safe + goa_dt * v_oa) + q_lona;
    double hls_r = dr + q_lonr;
    double hlds_r = dr + (ge_dt * v_safe + -gr_dt * v_oar) + q_lonr;
    double cond_a = hlds_a + (eta - 1) * hls_a;
    double cond_r = hlds_r + (eta - 1) * hls_r;
    // Tie rule: when the adjacent row is the binding (feasible) QP constraint cond_a is exactly 0 in exact
    // arithmetic and rounding noise in float64; it counts as satisfied (as an interior-point solve would give).
    bool adj_inv_ok = (active == MM_ACT_ADJ) ? true : (cond_a >= 0);
    bool allowed = (hls_a >= 0 && adj_inv_ok) && (hls_r >= 0 && cond_r >= 0);
    if (WITH_MARGIN) rec.lc_margin = fmin(fmin(fabs(hls_a), fabs(cond_a)), fmin(fabs(hls_r), fabs(cond_r)));

    double steer = act_steer;
    bool lc_safe = true;
    bool veto;
    if (!mass) {
        veto = !allowed;
    } else {
        // can_abort_lc (decentral_layer.py:728-736) only matters when the lane change is not allowed
        veto = false;
        if (!allowed) {
            double cx, cy;
            const double ech = CH(self), esh = SH(self);
            get_corner(ex, ey, ech, esh, true, cx, cy);
            bool can_abort = on_lane(elane, cx, cy, 0.0);
            if (can_abort) {
                get_corner(ex, ey, ech, esh, false, cx, cy);
                can_abort = on_lane(elane, cx, cy, 0.0);
            }
            veto = can_abort;
        }
        int hl = fl_hl(f);
        if (!veto && (hl == A_LANE_RIGHT || hl == A_LANE_LEFT) && espeed < 1.6667) v_safe = v_ll;
        f = cond_a >= -1e-6 ? (f | FL_CADJ) : (f & ~FL_CADJ);
    }
    if (veto) {
        // Re-steering towards the own lane is the nominal command itself when no lane change was under way (same
        // lane, same state -> same expression); only a vehicle that was actually changing lanes pays for a second
        // steering law.  (A crashed vehicle's nominal steering was zeroed by clip_actions, so it always recomputes.)
        bool same_cmd = fl_tlane(f) == elane && !(f & FL_CRASHED) && !CFG_STEER_VEL(cfg);
        f = fl_set(f, FL_TLANE_SHIFT, FL_3BIT, (uint32_t)elane);
        steer = same_cmd ? act_steer : steering_control(ex, ey, eh, espeed, elane);
        if (CFG_STEER_VEL(cfg)) steer = 20 * (steer * 0.125 - GF(F_STEERANG, self));   // not clipped on this path
        lc_safe = false;
    }
    f = constrain_adj ? (f | FL_COLLAB) : (f & ~FL_COLLAB);
    f = lc_safe ? (f | FL_LCSAFE) : (f & ~FL_LCSAFE);
    FL(self) = f;

    out_acc = (v_safe - evx) / dt;
    out_steer = steer;
    rec.leader = id_ol; rec.front_adj = id_oa; rec.rear_adj = id_oar;
    rec.constrain_adj = constrain_adj; rec.active = active; rec.is_lc_safe = lc_safe;
}

// ------------------------------------------------------------------------------------------------
// x-sorted order (ev.live / ev.pos)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t invert_order(uint64_t ord, int n) {
    uint64_t pos = 0;
    for (int p = 0; p < n; ++p) pos |= (uint64_t)p << (4 * nib(ord, p));
    return pos;
}
// Slot i has just moved to x = nx: restore the descending order by neighbouring swaps (a vehicle advances by at most
// 2.7 m per sub-step, so this is almost always zero swaps) and keep the inverse in step.
__device__ __forceinline__ void reorder_after_move(Env &ev, int i, double nx) {
    uint64_t live = ev.live, pos = ev.pos;
    int p = nib(pos, i);
    while (p > 0) {
        int a = nib(live, p - 1);
        if (!(X(a) < nx)) break;
        live = (live & ~(0xffull << (4 * (p - 1)))) | ((uint64_t)i << (4 * (p - 1))) | ((uint64_t)a << (4 * p));
        pos = (pos & ~(15ull << (4 * a))) | ((uint64_t)p << (4 * a));
        --p;
    }
    while (p < ev.n_veh - 1) {
        int b = nib(live, p + 1);
        if (!(X(b) > nx)) break;
        live = (live & ~(0xffull << (4 * p))) | ((uint64_t)b << (4 * p)) | ((uint64_t)i << (4 * (p + 1)));
        pos = (pos & ~(15ull << (4 * b))) | ((uint64_t)p << (4 * b));
        ++p;
    }
    ev.live = live;
    ev.pos = (pos & ~(15ull << (4 * i))) | ((uint64_t)p << (4 * i));
}

// ------------------------------------------------------------------------------------------------
// integration (kinematics.py:122-152, safe_controller.py:100-185, behavior.py:102-109,504-522)
// ------------------------------------------------------------------------------------------------
// `steer`, `acc`: the low-level action act() produced for this sub-step.
// Trigonometry: with t = tan(delta)/2 the slip angle beta = atan(t) has cos = 1/sqrt(1+t^2), sin = t/sqrt(1+t^2),
// and cos/sin(heading + beta) follow by angle addition from the cached cos/sin(heading); one sincos of the new
// heading refreshes the cache and gives g.vx = cos(heading' + beta) and the logged vx.  That is 2 libm calls per
// move instead of 6 (tan, atan, sincos, sin, cos, cos), each result within a few ulp of the reference's chain.
template <bool DIAG>
__device__ void vehicle_step(Env &ev, const StepParams &p, int i, int sub, size_t e_glob, uint32_t &shield_counts,
                             double steer, double acc, const double rec1vx, const double ge) {
    uint32_t f = FL(i);
    const bool cav = is_cav(f);
    const double dt = p.cfg.dt;
    const double speed = V(i), heading = H(i);
    const double ch = CH(i), sh = SH(i);
    const bool shielded = cav && CFG_SHIELD(p.cfg) != MM_SHIELD_NONE && !CFG_V0(p.cfg) && (f & FL_FG) && fl_hist(f) >= 2;
    if (!cav) GF(F_TIMER, i) = GF(F_TIMER, i) + dt;
    // clip_actions
    if (f & FL_CRASHED) { steer = 0.0; acc = -1.0 * speed; }
    if (speed > 40.0) acc = fmin(acc, 1.0 * (40.0 - speed));
    else if (speed < -40.0) acc = fmax(acc, 1.0 * (40.0 - speed));
    if (cav && !CFG_V0(p.cfg)) acc = clipd(acc, ACC_LO, ACC_HI);   // MDPLCVehicle.clip_actions; MDPVehicle (v0) has none
    GF(F_ACT_STEER, i) = steer;
    GF(F_ACT_ACC, i) = acc;
    if (cav) {
        // get_safe_action gate (safe_controller.py:229-239)
        if (shielded) {
            ShieldRec rec;
            double nom_steer = steer, nom_acc = acc;
            shield<DIAG>(ev, p.cfg, i, nom_steer, nom_acc, rec1vx, ge, steer, acc, rec);
            f = FL(i);
            // solves / active / vetoes of this policy step: three 10-bit counters in one register (<= 33 each)
            shield_counts += 1u | ((uint32_t)(rec.active != 0) << 10) | ((uint32_t)(!rec.is_lc_safe) << 20);
            if (DIAG) {
                size_t plane = (size_t)p.n_envs * 3 * MAXV;
                size_t idx = (e_glob * 3 + sub) * MAXV + i;
                int32_t *si = p.out.sh_i;
                si[idx] = 1; si[plane + idx] = rec.leader; si[2 * plane + idx] = rec.front_adj;
                si[3 * plane + idx] = rec.rear_adj; si[4 * plane + idx] = rec.constrain_adj;
                si[5 * plane + idx] = rec.active; si[6 * plane + idx] = rec.is_lc_safe;
                double *sf = p.out.sh_f;
                sf[idx] = acc; sf[plane + idx] = steer; sf[2 * plane + idx] = nom_acc;
                sf[3 * plane + idx] = nom_steer; sf[4 * plane + idx] = rec.lc_margin;
            }
        }
        GF(F_SAFE_STEER, i) = steer;
        GF(F_SAFE_ACC, i) = acc;
    }
    // modified bicycle model (kinematics.py:133-140, safe_controller.py:151-172)
    // steer_vel (safe_controller.py:124-150): the wheel angle is a state, the command is its rate, and the heading
    // increment is not multiplied by dt (as written in the reference)
    const bool sv = cav && CFG_STEER_VEL(p.cfg) && !CFG_V0(p.cfg);
    double wheel = steer;
    if (sv) {
        wheel = GF(F_STEERANG, i);
        GF(F_STEERANG, i) = wheel + steer * dt;
    }
    double t = 1.0 / 2 * m_tan(wheel);
    double cb = 1.0 / sqrt(1.0 + t * t), sb = t * cb;           // cos / sin of the slip angle
    double c_hb = ch * cb - sh * sb, s_hb = sh * cb + ch * sb;  // cos / sin (heading + beta)
    double nx = X(i) + speed * c_hb * dt;
    double ny = Y(i) + speed * s_hb * dt;
    double nh = sv ? heading + div_nz(speed * sb, VLEN / 2) : heading + div_nz(speed * sb, VLEN / 2) * dt;
    double nv = fmax(0.0, speed + acc * dt);
    double2 scn = m_sincos(nh);
    if (cav) {
        GF(F_GVX, i) = scn.y * cb - scn.x * sb;                 // cos(heading' + beta)
        f |= FL_FG;
    }
    // on_state_update + log_step
    int lane = closest_lane(nx, ny, nh);
    f = fl_set(f, FL_LANE_SHIFT, FL_3BIT, (uint32_t)lane);
    GF(F_REC2X, i) = X(i);
    GF(F_REC2VX, i) = rec1vx;
    GF(F_REC1VX, i) = nv * scn.y;
    CH(i) = scn.y;
    SH(i) = scn.x;
    int hist = fl_hist(f);
    if (hist < 2) f = fl_set(f, FL_HIST_SHIFT, 3u, (uint32_t)(hist + 1));
    X(i) = nx; Y(i) = ny; H(i) = nh; V(i) = nv;
    FL(i) = f;
    reorder_after_move(ev, i, nx);
    if (DIAG) {
        // control profile of this sub-step (safe_controller.py:187-227 log_step: the state after the move, the action
        // that produced it, the high-level action in force), for every vehicle, shielded or not
        const size_t plane = (size_t)p.n_envs * 3 * MAXV, idx = (e_glob * 3 + sub) * MAXV + i;
        int32_t *si = p.out.sh_i;
        double *sf = p.out.sh_f;
        const int hl = fl_hl(f);
        si[7 * plane + idx] = 1; si[8 * plane + idx] = hl == A_NONE ? -1 : hl; si[9 * plane + idx] = lane;
        if (!si[idx]) {   // the shield did not run: the applied action is the (clipped) nominal one
            sf[idx] = acc; sf[plane + idx] = steer; sf[2 * plane + idx] = acc; sf[3 * plane + idx] = steer;
        }
        sf[5 * plane + idx] = nx; sf[6 * plane + idx] = ny; sf[7 * plane + idx] = nh; sf[8 * plane + idx] = nv;
        sf[9 * plane + idx] = GF(F_MINHW, i);
    }
}

// ------------------------------------------------------------------------------------------------
// collisions (road.py:288-292, kinematics.py:175-209, utils.py:55-121)
// ------------------------------------------------------------------------------------------------
// does rect1 (centre c1, half sizes lx/wy, heading cos/sin co1/s1) have one of its 9 sample points inside rect2?
// NB the reference rotates (p - c2) by +a2, not -a2; reproduced as is.
__device__ __noinline__ bool has_corner_inside(double c1x, double c1y, double lx, double wy, double co1, double s1,
                                               double c2x, double c2y, double l2, double w2, double co2, double s2) {
    // sample points in the reference's order: centre, -l, +l, -w, +w, -l-w, -l+w, +l-w, +l+w (utils.py:115-117);
    // sign of the l / w component of point k packed two bits each (0: zero, 1: plus, 2: minus)
    const uint32_t lsel = 0x16818u, wsel = 0x19980u;
#pragma unroll 1
    for (int k = 0; k < 9; ++k) {
        uint32_t lc = (lsel >> (2 * k)) & 3u, wc = (wsel >> (2 * k)) & 3u;
        double pxk = lc == 0 ? 0.0 : (lc == 1 ? lx : -lx);
        double pyk = wc == 0 ? 0.0 : (wc == 1 ? wy : -wy);
        double rx = co1 * pxk + -s1 * pyk;
        double ry = s1 * pxk + co1 * pyk;
        double dx = (c1x + rx) - c2x, dy = (c1y + ry) - c2y;
        double ux = co2 * dx + -s2 * dy;
        double uy = s2 * dx + co2 * dy;
        if (-l2 / 2 <= ux && ux <= l2 / 2 && -w2 / 2 <= uy && uy <= w2 / 2) return true;
    }
    return false;
}

__device__ __noinline__ bool rects_intersect(double ax, double ay, double aco, double asn, double bx, double by,
                                             double bco, double bsn, double blen, double bwid) {
    return has_corner_inside(ax, ay, 0.9 * VLEN / 2, 0.9 * VWID / 2, aco, asn, bx, by, 0.9 * blen, 0.9 * bwid, bco, bsn) ||
           has_corner_inside(bx, by, 0.9 * blen / 2, 0.9 * bwid / 2, bco, bsn, ax, ay, 0.9 * VLEN, 0.9 * VWID, aco, asn);
}

// Conservative pre-test for has_corner_inside(rect1 -> rect2): with u = R(a2)(c1 + R(a1)q - c2) and q ranging over
// rect1's sample points (|qx| <= lx1, |qy| <= wy1), |ux| >= |cos a2| |dx| - |sin a2| |dy| and the analogous bound for
// uy hold with |dx| >= |Dx| - (lx1 + |sin a1| wy1), |dy| <= |Dy| + lx1 + wy1, ...; if either bound clears the half
// size of rect2 by more than 1e-6 m no sample point can pass the inside test (the reference's own rounding error
// there is ~1e-15), so the exact 9-point test is skipped.  Side-by-side vehicles on bc0/bc1 and queues behind
// the obstacle are the common case this removes.
__device__ __forceinline__ bool may_have_corner_inside(double adx, double ady, double lx1, double wy1, double as1,
                                                       double l2, double w2, double ac2, double as2) {
    double reach = lx1 + wy1;
    double uy_min = ac2 * fmax(0.0, ady - (as1 * lx1 + wy1)) - as2 * (adx + reach);
    if (uy_min > 0.5 * w2 + 1e-6) return false;
    double ux_min = ac2 * fmax(0.0, adx - (lx1 + as1 * wy1)) - as2 * (ady + reach);
    if (ux_min > 0.5 * l2 + 1e-6) return false;
    return true;
}
__device__ __forceinline__ bool may_intersect(double adx, double ady, double aco, double asn, double bco, double bsn,
                                              double blen, double bwid) {
    aco = fabs(aco); asn = fabs(asn); bco = fabs(bco); bsn = fabs(bsn);
    return may_have_corner_inside(adx, ady, 0.9 * VLEN / 2, 0.9 * VWID / 2, asn, 0.9 * blen, 0.9 * bwid, bco, bsn) ||
           may_have_corner_inside(adx, ady, 0.9 * blen / 2, 0.9 * bwid / 2, bsn, 0.9 * VLEN, 0.9 * VWID, aco, asn);
}

// Vehicles with a partner (or the obstacle) within LENGTH, centre to centre, as a bit mask over the slots: the pass below
// cannot change anything for the others.  In the x-sorted order a vehicle's partners within LENGTH are its immediate
// successors.
__device__ __forceinline__ uint32_t close_pair_mask(const Env &ev) {
    const uint64_t live = ev.live;
    uint32_t mask = 0;
    for (int p = 0; p < ev.n_veh; ++p) {
        const int a = nib(live, p);
        const double ax = X(a), ay = Y(a);
        {
            double dx = OBST_X - ax, dy = OBST_Y - ay;
            if (!(dx * dx + dy * dy > VLEN_SQ_GT)) mask |= 1u << a;
        }
        for (int q = p + 1; q < ev.n_veh; ++q) {
            const int b = nib(live, q);
            double dx = X(b) - ax, dx2 = dx * dx;
            if (dx2 > VLEN_SQ_GT) break;             // x only grows apart from here on
            double dy = Y(b) - ay;
            if (!(dx2 + dy * dy > VLEN_SQ_GT)) mask |= (1u << a) | (1u << b);
        }
    }
    return mask;
}

// road.py:288-292 restricted to the vehicles of `mask` (slot order is the reference's list order)
__device__ __noinline__ void collision_pass(Env &ev, uint32_t mask) {
    for (uint32_t mi = mask; mi; mi &= mi - 1) {
        const int i = __ffs(mi) - 1;
        double ax = X(i), ay = Y(i);
        for (uint32_t mj = mask; mj; mj &= mj - 1) {
            const int j = __ffs(mj) - 1;
            if (j == i) continue;
            if (FL(i) & FL_CRASHED) break;
            // The pair {j < i} already had its turn as (j, i) unless j was crashed by then (check_collision returns
            // early for a crashed caller); an uncrashed j means that test ran and was negative, and the geometry
            // has not changed since, so only crashed lower-index partners need the test from this side.
            if (j < i && !(FL(j) & FL_CRASHED)) continue;
            double dx = X(j) - ax, dy = Y(j) - ay;
            if (dx * dx + dy * dy > VLEN_SQ_GT) continue;  // np.linalg.norm(...) > LENGTH
            double aco = CH(i), asn = SH(i), bco = CH(j), bsn = SH(j);
            if (!may_intersect(fabs(dx), fabs(dy), aco, asn, bco, bsn, VLEN, VWID)) continue;
            if (rects_intersect(ax, ay, aco, asn, X(j), Y(j), bco, bsn, VLEN, VWID)) {
                double va = V(i), vb = V(j);
                double m = fabs(va) <= fabs(vb) ? va : vb;
                V(i) = m; V(j) = m;
                FL(i) |= FL_CRASHED; FL(j) |= FL_CRASHED;
            }
        }
        if (!(FL(i) & FL_CRASHED)) {
            double dx = OBST_X - ax, dy = OBST_Y - ay;
            if (!(dx * dx + dy * dy > VLEN_SQ_GT)) {
                double aco = CH(i), asn = SH(i);
                if (may_intersect(fabs(dx), fabs(dy), aco, asn, 1.0, 0.0, 2.0, 2.0) &&
                    rects_intersect(ax, ay, aco, asn, OBST_X, OBST_Y, 1.0, 0.0, 2.0, 2.0)) {
                    double va = V(i);
                    V(i) = fabs(va) <= 0 ? va : 0.0;
                    FL(i) |= FL_CRASHED;
                }
            }
        }
    }
}

// vehicles that are observed / rewarded: the controlled ones, or all of them in MergeEnvLCHDV (observation.py:430-442,
// merge_env_v1.py:518-524)
__device__ __forceinline__ int n_observed(const Env &ev, const mm_config &cfg) { return CFG_ENV_HDV(cfg) ? ev.n_veh : ev.n_cav; }

// merge_env_v1.py:168-172; MergeEnvLCHDV 670-673: any vehicle crashed, no x < 0 clause
__device__ __forceinline__ bool is_terminal(const Env &ev, int steps, const mm_config &cfg) {
    bool t = steps >= cfg.duration_steps;
    const int n = n_observed(ev, cfg);
    for (int i = 0; i < n; ++i) t = t || (FL(i) & FL_CRASHED) || (!CFG_ENV_HDV(cfg) && X(i) < 0);
    return t;
}

// ------------------------------------------------------------------------------------------------
// observation, rewards, info
// ------------------------------------------------------------------------------------------------
// observation.py:241-273 + normalize_obs 181-193: ego row absolute, 4 nearest rows relative, no clipping.
// The rows are float32 outputs, so lmap's divisions by the constant ranges are multiplications by the
// reciprocals here (a <= 1-ulp float64 difference, invisible after rounding to float32).
// observation / reward helpers: one call site each in the outputs kernel, which inlines them so that the per-thread Env
// stays in registers (a reference to it passed to an out-of-line function lives in local memory)
#ifndef MM_OUT_FN
#define MM_OUT_FN __noinline__
#endif
#ifndef MM_OBS_F32
#define MM_OBS_F32 0   // experiment knob: 1 = normalise in float32 after the float64 differences (profiles/README.md)
#endif
// Returns the slots of the observed neighbours, 4 bits each in row order, 0xF = no vehicle in that row.
__device__ MM_OUT_FN uint32_t observe_agent(const Env &ev, int self, bool steer_vel, float2 *dst) {
    // Vehicle.velocity (kinematics.py:215-217) = speed * [cos, sin](heading)
    double ex = X(self), ey = Y(self), evx = V(self) * CH(self), evy = V(self) * SH(self);
    uint32_t nb_ids;
    int n_nb = close_vehicles<4>(ev, self, nb_ids);
    // `dst` is the agent's 120-byte row in the CTA's shared-memory staging block (8-byte aligned); the block leaves for
    // HBM as coalesced streaming stores once every row of the CTA is built
#if MM_OBS_F32
    // The differences (the only place where magnitudes cancel) are taken in float64; the affine maps onto [-1, 1] run in
    // float32: <= 3 float32 roundings (~2e-7) against the float64 evaluation, inside the 2e-6 bar of the parity tests
    const float KX = (float)(2.0 / 300.0), KY = (float)(2.0 / 24.0), KV = (float)(2.0 / 90.0), KH = (float)(2.0 / PI);
    const float HPI = (float)(PI / 2);
    dst[0] = make_float2(1.0f, ((float)ex + 150.0f) * KX - 1.0f);
    dst[1] = make_float2(((float)ey + 12.0f) * KY - 1.0f, ((float)evx + 45.0f) * KV - 1.0f);
    dst[2] = make_float2(((float)evy + 45.0f) * KV - 1.0f, ((float)H(self) + HPI) * KH - 1.0f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float2 a = make_float2(0.f, 0.f), b = a, c = a;
        if (k < n_nb) {
            int o = (int)((nb_ids >> (4 * k)) & 15u);
            a = make_float2(1.0f, ((float)(X(o) - ex) + 150.0f) * KX - 1.0f);
            const double ov = V(o);
            b = make_float2(((float)(Y(o) - ey) + 12.0f) * KY - 1.0f, ((float)(ov * CH(o) - evx) + 45.0f) * KV - 1.0f);
            double oh = H(o);
            if (steer_vel && o < ev.n_cav) oh = oh - H(self);   // MDPLCVehicle.to_dict(origin) (safe_controller.py:75-81)
            c = make_float2(((float)(ov * SH(o) - evy) + 45.0f) * KV - 1.0f, ((float)oh + HPI) * KH - 1.0f);
        }
        dst[3 * (k + 1)] = a;
        dst[3 * (k + 1) + 1] = b;
        dst[3 * (k + 1) + 2] = c;
    }
#else
    const double KX = 2.0 / 300.0, KY = 2.0 / 24.0, KV = 2.0 / 90.0, KH = 2.0 / PI;
    dst[0] = make_float2(1.0f, (float)((ex + 150.0) * KX - 1.0));
    dst[1] = make_float2((float)((ey + 12.0) * KY - 1.0), (float)((evx + 45.0) * KV - 1.0));
    dst[2] = make_float2((float)((evy + 45.0) * KV - 1.0), (float)((H(self) + PI / 2) * KH - 1.0));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float2 a = make_float2(0.f, 0.f), b = a, c = a;
        if (k < n_nb) {
            int o = (int)((nb_ids >> (4 * k)) & 15u);
            a = make_float2(1.0f, (float)(((X(o) - ex) + 150.0) * KX - 1.0));
            const double ov = V(o);
            b = make_float2((float)(((Y(o) - ey) + 12.0) * KY - 1.0), (float)(((ov * CH(o) - evx) + 45.0) * KV - 1.0));
            double oh = H(o);
            if (steer_vel && o < ev.n_cav) oh = oh - H(self);   // MDPLCVehicle.to_dict(origin) (safe_controller.py:75-81)
            c = make_float2((float)(((ov * SH(o) - evy) + 45.0) * KV - 1.0), (float)((oh + PI / 2) * KH - 1.0));
        }
        dst[3 * (k + 1)] = a;
        dst[3 * (k + 1) + 1] = b;
        dst[3 * (k + 1) + 2] = c;
    }
#endif
    return (nb_ids | (0xFFFFu << (4 * n_nb))) & 0xFFFFu;
}

// abstract.py:620-635: the smallest x-gap to a vehicle strictly ahead in the same lane or (unless on bc1) in the next
// lane, default 60.  Walking ahead in the x-sorted order the first match is the minimum, and nothing 60 m ahead matters.
__device__ MM_OUT_FN double headway_distance(const Env &ev, int self) {
    const double ex = X(self);
    const int lane = fl_lane(FL(self));
    const int nl = next_lane(lane, ex, Y(self));
    const bool use_next = lane != L_BC1;
    const uint64_t live = ev.live;
    for (int p = nib(ev.pos, self) - 1; p >= 0; --p) {
        const int j = nib(live, p);
        const double xj = X(j), d = xj - ex;
        if (!(xj > ex)) continue;          // same x: not ahead
        if (!(d < 60)) break;
        const int lj = fl_lane(FL(j));
        if (lj == lane || (use_next && lj == nl)) return d;
    }
    return 60;
}

// merge_env_v1.py:64-89 and 439-474
__device__ MM_OUT_FN double agent_reward(const Env &ev, const mm_config &cfg, int self, double hd) {
    uint32_t f = FL(self);
    bool special = cfg.reward_kind != MM_REW_DEFAULT && fl_kind(f) == MM_KIND_CAV;
    bool mrew = cfg.reward_kind == MM_REW_MREW;
    double speed = V(self);
    double r1 = 30.0;
    if (special && mrew && (f & FL_COLLAB)) r1 = 10.0 + (30.0 - 10.0) / 2;
    double scaled = lmap(speed, 10.0, r1, 0, 1);
    double merging = 0.0;
    if (fl_lane(f) == L_BC1 && (!special || !mrew || (f & FL_LCSAFE))) {
        double d = X(self) - 420.0;
        merging = -m_exp(-(d * d) / (10 * 100.0));
    }
    double hc = 0.0;
    if (speed > 0) {
        hc = m_log(hd / (cfg.headway_time * speed));
        if (special) hc = -1 * hc;
    }
    double crashed = (f & FL_CRASHED) ? 1.0 : 0.0;
    return cfg.collision_reward * (-1 * crashed) + (cfg.high_speed_reward * clipd(scaled, 0.0, 1.0)) +
           cfg.merging_lane_cost * merging + cfg.headway_cost * (hc < 0 ? hc : 0.0);
}

// road.py:294-350: nearest vehicle ahead / behind by world x among the lanes visible from a query lane
// (bit l of a mask = lane l is visible).  Two queries share one pass over the vehicles.
__device__ __forceinline__ uint32_t visible_lanes(int qlane) {
    // ab0:{ab0,bc0} bc0:{ab0,bc0,cd0} bc1:{kb0,bc1} cd0:{bc0,cd0} jk0:{jk0,kb0} kb0:{jk0,kb0,bc1}
    return (0x34300A240B03ull >> (8 * qlane)) & 0xffu;
}
__device__ __noinline__ void surrounding2_scan(const Env &ev, int self, uint32_t m1, uint32_t m2, int &f1, int &r1, int &f2,
                                               int &r2) {
    double s = X(self), sf1 = 0, sr1 = 0, sf2 = 0, sr2 = 0;
    f1 = r1 = f2 = r2 = -1;
    for (int j = 0; j < ev.n_veh; ++j) {
        if (j == self) continue;
        uint32_t bit = 1u << fl_lane(FL(j));
        double s_v = X(j);
        bool ahead = s <= s_v, behind = s_v < s;
        if (m1 & bit) {
            if (ahead && (f1 < 0 || s_v <= sf1)) { sf1 = s_v; f1 = j; }
            if (behind && (r1 < 0 || s_v > sr1)) { sr1 = s_v; r1 = j; }
        }
        if (m2 & bit) {
            if (ahead && (f2 < 0 || s_v <= sf2)) { sf2 = s_v; f2 = j; }
            if (behind && (r2 < 0 || s_v > sr2)) { sr2 = s_v; r2 = j; }
        }
    }
}

// The same through the x-sorted order: the nearest visible vehicle ahead / behind is the first one met walking away from
// the ego.  Equal x values (with the ego or between two candidates) are where the reference's tie rules bite (ahead
// includes equality and prefers the later slot, behind prefers the earlier): any equality on the way defers to the scan.
__device__ __forceinline__ void surrounding2(const Env &ev, int self, uint32_t m1, uint32_t m2, int &f1, int &r1, int &f2,
                                             int &r2) {
    const uint64_t live = ev.live;
    const int ps = nib(ev.pos, self);
    const bool need2 = m2 != 0;
    f1 = r1 = f2 = r2 = -1;
    bool tie = false;
    double prev = X(self);
    int p = ps - 1;
    for (; p >= 0 && (f1 < 0 || (need2 && f2 < 0)); --p) {
        const int j = nib(live, p);
        const double xj = X(j);
        tie |= xj == prev;
        prev = xj;
        const uint32_t bit = 1u << fl_lane(FL(j));
        if (f1 < 0 && (m1 & bit)) f1 = j;
        if (f2 < 0 && (m2 & bit)) f2 = j;
    }
    if (p >= 0) tie |= X(nib(live, p)) == prev;
    prev = X(self);
    p = ps + 1;
    for (; p < ev.n_veh && (r1 < 0 || (need2 && r2 < 0)); ++p) {
        const int j = nib(live, p);
        const double xj = X(j);
        tie |= xj == prev;
        prev = xj;
        const uint32_t bit = 1u << fl_lane(FL(j));
        if (r1 < 0 && (m1 & bit)) r1 = j;
        if (r2 < 0 && (m2 & bit)) r2 = j;
    }
    if (p < ev.n_veh) tie |= X(nib(live, p)) == prev;
    if (tie) surrounding2_scan(ev, self, m1, m2, f1, r1, f2, r2);
}

// ------------------------------------------------------------------------------------------------
// x-descending stable order packed as 4-bit slot ids (road.py:277,286)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t order_by_x_desc(const Env &ev) {
    uint64_t ord = 0;
    for (int i = 0; i < ev.n_veh; ++i) {
        double xi = X(i);
        int p = i;
        while (p > 0 && X((int)((ord >> (4 * (p - 1))) & 15u)) < xi) --p;
        uint64_t low = ord & ((1ull << (4 * p)) - 1ull);
        uint64_t high = (ord >> (4 * p)) << (4 * (p + 1));
        ord = low | ((uint64_t)i << (4 * p)) | high;
    }
    return ord;
}

}  // namespace MM_KNS
