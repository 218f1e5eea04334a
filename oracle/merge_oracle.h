/* merge_oracle.h — CPU float64 restatement of the MARL-MASS merge env step + HSS/MASS shields.
 *
 * TEST INFRASTRUCTURE.  This is the parity oracle, not the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * It is pinned against the tests/golden fixtures, which were produced by running the unmodified
 * reference (oracle/refharness/gen_golden.py).  PARITY UNPINNED for one thing only: the
 * reference's QP goes through cvxopt 1.2.7 (absent here); both the fixtures and this file use
 * the closed-form minimiser of the same QP.
 *
 * Layout: every per-vehicle array is env-major [n_env][MO_MAXV]; slot order is the reference's
 * road.vehicles list order (CAVs first), which is also the vehicle id.
 */
#ifndef MERGE_ORACLE_H
#define MERGE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MO_MAXV 12
#define MO_OBS_ROWS 5
#define MO_OBS_FEATS 6
#define MO_NS (MO_OBS_ROWS * MO_OBS_FEATS)

enum { MO_KIND_NONE = 0, MO_KIND_CAV = 1, MO_KIND_HDV = 2 };
enum { MO_SHIELD_NONE = 0, MO_SHIELD_HSS = 1, MO_SHIELD_MASS = 2 };
enum { MO_REW_DEFAULT = 0, MO_REW_SREW = 1, MO_REW_MREW = 2 };
enum { MO_NB_NONE = -1, MO_NB_OBSTACLE = -2 };
/* QP active-set code (ours; the reference discards multipliers) */
enum { MO_ACT_LEAD = 1, MO_ACT_UPPER = 2, MO_ACT_LOWER = 4, MO_ACT_ADJ = 8, MO_ACT_SLACK = 16 };

typedef struct {
    int32_t shield;          /* MO_SHIELD_*  (safety_guarantee: none | cbf-avs_cint/hss/av/avs | cbf-cav/mass) */
    int32_t reward_kind;     /* MO_REW_*     (agent_reward) */
    int32_t duration_steps;  /* duration * policy_frequency = 100 */
    int32_t substeps;        /* simulation_frequency // policy_frequency = 3 */
    double dt;               /* 1 / simulation_frequency */
    double eta;              /* CBFType.GAMMA_B (cbf_eta) */
    double tau;              /* CBFType.TAU */
    double collision_reward, high_speed_reward, headway_cost, headway_time, merging_lane_cost;
    int32_t env_v0;          /* 1: merge-multi-agent-v0 (MDPVehicle: no [-12.5, 6] acceleration clip, never shielded) */
    int32_t steer_vel;       /* 1: lateral_control = steer_vel (safe_controller.py:84-98, 124-150) */
    int32_t env_hdv;         /* 1: merge-multi-agent-hdv-v1 (MergeEnvLCHDV, merge_env_v1.py:552-674): no controlled vehicles;
                                every vehicle is observed and rewarded, any crash ends the episode */
} mo_config;

typedef struct {
    double *x, *y, *heading, *speed, *target_speed, *gvx, *rec1_x, *rec1_vx, *rec2_x, *rec2_vx,
           *act_steer, *act_acc, *safe_steer, *safe_acc, *timer, *min_headway, *steering_angle;
    int32_t *kind, *lane, *target_lane, *speed_index, *crashed, *hl_action, *hist_len, *fg_set,
            *is_collaborating, *is_lc_safe, *collaborate_adj;
    int32_t *n_veh, *n_cav, *n_merge, *steps, *time;   /* [n_env] */
} mo_state;

typedef struct {
    double *obs;               /* [n_env][MO_MAXV][MO_NS] */
    double *reward;            /* [n_env] global reward (mean of local) */
    int32_t *done;             /* [n_env] */
    double *agents_rewards, *regional_rewards;   /* [n_env][MO_MAXV] */
    int32_t *agents_dones;     /* [n_env][MO_MAXV] */
    double *average_speed, *traffic_speed, *min_headway, *merge_percent;   /* [n_env] */
    /* per-sub-step shield record, [n_env][3][MO_MAXV] */
    int32_t *sh_ran, *sh_leader, *sh_front_adj, *sh_rear_adj, *sh_constrain_adj, *sh_active, *sh_is_lc_safe;
    double *sh_safe_acc, *sh_safe_steer, *sh_nom_acc, *sh_nom_steer, *sh_lc_margin;
} mo_out;

/* One policy step (AbstractEnv.step + MergeEnv.step) for envs [0, n_env); actions [n_env][MO_MAXV] int8. */
void mo_step(const mo_config *cfg, const mo_state *st, const int8_t *actions, const mo_out *out,
             int n_env, int n_threads);

/* Observation only (reset() returns it): obs [n_env][MO_MAXV][MO_NS]. */
void mo_observe(const mo_state *st, double *obs, int n_env, int steer_vel, int env_hdv);

/* The QP alone: closed-form minimiser; returns u, writes the active-set code. */
double mo_qp(double a, double c_lead, double c_adj, int has_adj, double lo, double hi, int32_t *active);

#ifdef __cplusplus
}
#endif
#endif
