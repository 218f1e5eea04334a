/* marl_mass_b200.h — C ABI of the B200-native batched merge environment + HSS/MASS CBF shields.
 *
 * The reference (hkbharath/MARL-MASS) is pure Python and has no FFI; its boundary for this path is two
 * Python call surfaces.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference tree).  The Python host (marl-mass_b200/env.py) binds these with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: every function returns 0 on success and a negative mm_status on failure, never throws;
 * mm_last_error() gives the text.  A handle owns all of its device memory.  Kernels are enqueued on the
 * caller's stream (a cudaStream_t passed as void*; NULL = default stream) with no implicit synchronisation,
 * except the *_host entry points, which return after the results are in the host buffers.
 * One host thread per handle.
 *
 * Host-side state/diagnostic arrays are env-major, [n_envs][MM_MAXV] per vehicle field, slot = index in the
 * reference's road.vehicles (CAVs first, which is also controlled_vehicles order and the vehicle id).
 */
#ifndef MARL_MASS_B200_H
#define MARL_MASS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM_MAXV 12          /* vehicle slots per env (reference max is 11: td3 = 6 CAV + 5 HDV) */
#define MM_OBS_ROWS 5       /* KinematicLCObservation vehicles_count (observation.py:132) */
#define MM_OBS_FEATS 6      /* presence,x,y,vx,vy,heading (observation.py:232) */
#define MM_NS 30            /* MergeEnvLCMARL.n_s (merge_env_v1.py:413) */
#define MM_NA 5             /* n_a: LANE_LEFT, IDLE, LANE_RIGHT, FASTER, SLOWER (action.py:141-147) */
#define MM_VEH_F32 4        /* floats per vehicle of the packed host path: x, y, heading, speed */

typedef enum { MM_OK = 0, MM_ERR_ARG = -1, MM_ERR_CUDA = -2, MM_ERR_STATE = -3 } mm_status;

enum { MM_KIND_NONE = 0, MM_KIND_CAV = 1, MM_KIND_HDV = 2 };
enum { MM_SHIELD_NONE = 0, MM_SHIELD_HSS = 1, MM_SHIELD_MASS = 2 };
enum { MM_REW_DEFAULT = 0, MM_REW_SREW = 1, MM_REW_MREW = 2 };
enum { MM_TRAFFIC_CAV = 0, MM_TRAFFIC_MIXED = 1, MM_TRAFFIC_AV = 2 /* one CAV, the rest HDVs (merge_env_v1.py:485-489) */,
       MM_TRAFFIC_HDV = 3 /* HDVs only (490-494); with env_hdv */ };
enum { MM_SUPERVISOR_NONE = 0, MM_SUPERVISOR_PRIORITY = 1, MM_SUPERVISOR_DMC = 2 };
enum { MM_NB_NONE = -1, MM_NB_OBSTACLE = -2 };
/* QP active-set code reported per shield solve */
enum { MM_ACT_LEAD = 1, MM_ACT_UPPER = 2, MM_ACT_LOWER = 4, MM_ACT_ADJ = 8, MM_ACT_SLACK = 16 };

/* Replaces: env.config[...] keys set by run_mappo.py:145-171 and the class globals CBFType.GAMMA_B / TAU
 * (run_mappo.py:137-139).  As in the reference, a changed config is picked up by the next reset. */
/* ABI guard: MM_ABI_VERSION changes whenever a struct of this header changes layout or an entry point changes its
 * signature; mm_abi_version() returns the value the library was built with, and mm_create / mm_set_config reject a
 * config whose struct_size field is not sizeof(mm_config) (a binding built against another revision of the header
 * fails loudly instead of handing the library garbage). */
#define MM_ABI_VERSION 8
int mm_abi_version(void);

typedef struct {
    int32_t struct_size;     /* = sizeof(mm_config), set by the caller (ABI guard) */
    int32_t shield;          /* safety_guarantee: none -> NONE; cbf-avs_cint|hss|av|avs -> HSS; cbf-cav|mass -> MASS */
    int32_t reward_kind;     /* agent_reward: default | srew | mrew */
    int32_t traffic_density; /* 1..3 (merge_env_v1.py:180-211) */
    int32_t traffic_type;    /* cav (all vehicles are CAVs) | mixed (HDVs are IDMVehicleHist) */
    int32_t duration_steps;  /* duration * policy_frequency (100) */
    int32_t substeps;        /* simulation_frequency // policy_frequency (3) */
    double dt;               /* 1 / simulation_frequency */
    double eta;              /* cbf_eta -> CBFType.GAMMA_B */
    double tau;              /* HEADWAY_TIME -> CBFType.TAU */
    double collision_reward, high_speed_reward, headway_cost, headway_time, merging_lane_cost;
    int32_t env_v0;          /* 1: env id merge-multi-agent-v0 (MergeEnvMARL, merge_env_v1.py:389-408): MDPVehicle CAVs
                                (no [-12.5, 6] acceleration clip, never shielded); obs columns 0..4 of each row */
    int32_t steer_vel;       /* 1: lateral_control = steer_vel (safe_controller.py:84-98,124-150): steering angle is a
                                state driven by a steering-velocity command; neighbour CAV headings are observed
                                relative to the ego */
    int32_t couple_counts;   /* device spawn only (no reference counterpart; default 0).  1: the vehicle COUNTS of an
                                episode (merge_env_v1.py:180-211) are drawn once per 128-env tile instead of once per
                                env; every env still follows the reference's law, but the envs of a tile share
                                (n_CAV, n_HDV), which lets a CTA skip the vehicle ranks none of its envs has */
    int32_t env_hdv;         /* 1: env id merge-multi-agent-hdv-v1 (MergeEnvLCHDV, merge_env_v1.py:552-674; traffic_type
                                hdv): no controlled vehicles - every vehicle is observed and rewarded (its row in obs /
                                agents_rewards), any crash ends the episode, no regional rewards / per-agent dones */
    int32_t supervisor;      /* safety_guarantee: priority -> PRIORITY, dmc -> DMC, else NONE.  The look-ahead baselines
                                (vehicle/safety/central_layer.py:16-178, decentralised_dmc.py:70-198) that AbstractEnv.step
                                runs on the action tuple before _simulate (abstract.py:459-464); shield must be NONE */
} mm_config;

typedef struct mm_env mm_env;

/* Host mirror of the full per-env state, for teacher forcing and checkpointing.
 * Replaces: nothing in the reference (it never checkpoints env state); fields follow
 * Vehicle / ControlledVehicle / MDPLCVehicle / IDMVehicleHist attributes.
 * rec1_* / rec2_* are state_hist[-1] / state_hist[-2] (x and vx of the record).  The reference logs a record right
 * after every move (safe_controller.py:187-205, behavior.py:509-519), so between policy steps rec1_x == x always:
 * mm_set_state does not store rec1_x (it is not read back from the caller's buffer), mm_get_state returns x for it. */
typedef struct {
    double *x, *y, *heading, *speed, *target_speed, *gvx, *rec1_x, *rec1_vx, *rec2_x, *rec2_vx,
           *act_steer, *act_acc, *safe_steer, *safe_acc, *timer, *min_headway,
           *steering_angle;                                                        /* [n_envs][MM_MAXV] */
    int32_t *kind, *lane, *target_lane, *speed_index, *crashed, *hl_action, *hist_len, *fg_set,
            *is_collaborating, *is_lc_safe, *collaborate_adj;                        /* [n_envs][MM_MAXV] */
    int32_t *n_veh, *n_cav, *n_merge, *steps, *time;                                 /* [n_envs] */
} mm_state_host;

/* Device pointers of the step outputs (valid until mm_destroy; contents valid after the step's stream work).
 * Replaces the return values of MergeEnv.step (merge_env_v1.py:126-166) and AbstractEnv.step's info dict
 * (abstract.py:489-498). */
typedef struct {
    float *obs;               /* [n_envs][MM_MAXV][MM_NS]; rows >= n_agents are zero */
    float *reward;            /* [n_envs] mean of local rewards (merge_env_v1.py:517-524) */
    uint8_t *done;            /* [n_envs] _is_terminal (merge_env_v1.py:168-172) */
    float *agents_rewards;    /* [n_envs][MM_MAXV] info["agents_rewards"] */
    float *regional_rewards;  /* [n_envs][MM_MAXV] info["regional_rewards"] */
    uint8_t *agents_dones;    /* [n_envs][MM_MAXV] info["agents_dones"] */
    float *average_speed;     /* [n_envs] info["average_speed"] */
    float *traffic_speed;     /* [n_envs] info["traffic_speed"] */
    float *min_headway;       /* [n_envs] info["min_headway"] */
    float *merge_percent;     /* [n_envs] info["merge_percent"] at done, else -1 */
    int32_t *n_agents;        /* [n_envs] len(env.controlled_vehicles) */
    int8_t *actions;          /* [n_envs][MM_MAXV] device-side action buffer read by mm_step(actions=NULL) */
    uint8_t *action_mask;     /* [n_envs][MM_MAXV] bit a = meta-action a available (_get_available_actions,
                                 abstract.py:219-240), per agent; info["action_mask"] when action_masking is on */
    int8_t *new_actions;      /* [n_envs][MM_MAXV] info["new_action"]: the tuple _simulate executed when a baseline
                                 supervisor is configured (abstract.py:458-467); not written otherwise */
} mm_buffers;

/* Host arrays receiving the per-sub-step shield record of the last mm_step, [n_envs][3][MM_MAXV].
 * Replaces: safe_status/safe_diff/safe_action logged by MDPLCVehicle.log_step (safe_controller.py:187-227). */
typedef struct {
    int32_t *ran, *leader, *front_adj, *rear_adj, *constrain_adj, *active, *is_lc_safe;
    /* control profile (safe_controller.py:187-227 log_step), every vehicle that moved in the sub-step: */
    int32_t *moved, *hl_action /* -1 none */, *lane;
    double *safe_acc, *safe_steer, *nom_acc, *nom_steer, *lc_margin;   /* applied and nominal action (all vehicles) */
    double *x, *y, *heading, *speed, *min_headway;                     /* state after the sub-step's move */
} mm_shield_diag_host;

/* Episode statistics accumulated on the device since the last mm_stats(reset=1).
 * Sums are what ranks all-reduce(SUM); min_headway all-reduces with MIN. */
typedef struct {
    double agent_steps, env_steps, episodes, crashed_episodes, reward_sum, speed_sum, merge_percent_sum,
           shield_solves, shield_active, lane_change_vetoes;
    double min_headway;
} mm_stats_t;

/* gym.make(env_id) + config mutation (run_mappo.py:143-171).  record_diag != 0 keeps the per-sub-step shield
 * record (test / control-profile use; costs extra stores). */
int mm_create(const mm_config *cfg, int n_envs, int device, int record_diag, mm_env **out);
int mm_destroy(mm_env *env);
int mm_set_config(mm_env *env, const mm_config *cfg);
int mm_num_envs(const mm_env *env);

/* AbstractEnv.reset (abstract.py:176-209) for every env whose mask byte is non-zero (mask == NULL: all).
 * Spawn follows merge_env_v1.py:180-211,265-364 with a counter-based RNG keyed by (seed, env index, episode
 * counter): the same law as the reference, not the same MT19937 stream (use mm_set_state for exact scenes).
 * num_cav > 0 pins the CAV count (reset(num_CAV=...)).  mask is a DEVICE pointer.  Writes the first obs. */
int mm_reset(mm_env *env, uint64_t seed, const uint8_t *mask_dev, int num_cav, void *stream);

/* MergeEnv.step (merge_env_v1.py:126-166): actions_dev is [n_envs][MM_MAXV] int8 on the device
 * (NULL: use mm_buffers.actions).  auto_reset != 0 re-spawns finished envs after their outputs are written
 * (the obs of a finished env is then the first obs of its next episode, as MAPPO.interact does, mappo.py:133-135). */
int mm_step(mm_env *env, const int8_t *actions_dev, int auto_reset, void *stream);

/* Same step through HOST buffers (pinned or pageable): copies actions in, steps, copies obs/reward/done out,
 * chunked over internal streams so copies overlap compute.  Any output pointer may be NULL. */
int mm_step_host(mm_env *env, const int8_t *actions, int auto_reset, float *obs, float *reward, uint8_t *done,
                 float *regional_rewards, int32_t *n_agents);

/* mm_step_host with the observations as the reference returns them (obs ndarray [A, n_s] per env,
 * merge_env_v1.py:126-166): only the rows of the agents that exist.  Env e's rows are
 * obs_rows[row_offset[e] * 30 .. (row_offset[e] + n_agents[e]) * 30); offsets are absolute, increasing, and dense inside
 * every chunk of the call (half a wave of the step kernel, 28 416 envs on a 148-SM part; chunk c starts at row chunk_first_env * MM_MAXV; row_offset[n_envs] = n_envs * MM_MAXV).
 * obs_rows has room for n_envs * MM_MAXV rows of 30 f32 (pinned memory for full PCIe speed); only the packed rows are
 * transferred: a quarter fewer bytes at hard density.  row_offset [n_envs + 1] int64 host; the other outputs as in
 * mm_step_host (nullable). */
int mm_step_host_ragged(mm_env *env, const int8_t *actions, int auto_reset, float *obs_rows, int64_t *row_offset,
                        float *reward, uint8_t *done, float *regional_rewards, int32_t *n_agents);
/* The same step with the observation in PACKED form: instead of the [A, 30] rows, what they are a function of.
 *   veh     [sum n_veh][MM_VEH_F32] f32   x, y, heading, speed of every vehicle, envs in order, slots in order (CAVs first)
 *   nbr     [sum n_agents] u16   per agent: the slots of the (up to) 4 other vehicles its observation shows, 4 bits each in
 *                                row order (close_vehicles_to(count = 4), road.py:257-267), 0xF = that row is empty
 *   n_veh, n_agents [n_envs] u8  the counts; all offsets are their running sums (dense from the first env to the last)
 *   reward [n_envs] f32, done [n_envs] u8, regional_rewards [n_envs][MM_MAXV] f32 as in mm_step_host (nullable)
 * About 215 bytes per env-step at hard density against 1.1 KB for the packed rows of mm_step_host_ragged: the host path
 * is bound by PCIe and host memory, not by the kernels.  Row i of env e is rebuilt by mm_expand_obs_rows:
 * ego columns from veh[i] (vx, vy = speed * cos / sin(heading), kinematics.py:215-217), row k + 1 from veh[slot k of nbr]
 * - veh[i] (observation.py:241-273, 181-193); inputs are float32, so the rebuilt rows agree with mm_step_host's to ~2e-7 (float32 rounding of positions up to 450 m mapped onto
 * [-1, 1]), inside the 1e-4 tolerance of the path.  With auto_reset the packed state of a finished env is that of its next
 * episode (like obs), reward / done / regional_rewards describe the finished step.  veh needs room for n_envs * 11 rows,
 * nbr for n_envs * MM_MAXV entries (pinned memory for full PCIe speed). */
typedef struct {
    float *veh;
    uint16_t *nbr;
    uint8_t *n_veh, *n_agents;
    float *reward;
    uint8_t *done;
    float *regional_rewards;
} mm_packed_host;
int mm_step_host_packed(mm_env *env, const int8_t *actions, int auto_reset, const mm_packed_host *out);
/* Pure host function (no device): obs_rows [sum n_agents][30] f32 and row_offset [n_envs + 1] from a packed step, on
 * n_threads host threads.  steer_vel: the config's lateral_control == "steer_vel" (neighbour CAV headings relative). */
int mm_expand_obs_rows(const mm_packed_host *in, int n_envs, int steer_vel, float *obs_rows, int64_t *row_offset, int n_threads);
int mm_buffers_get(mm_env *env, mm_buffers *out);
int mm_get_state(mm_env *env, mm_state_host *dst);        /* synchronous */
int mm_set_state(mm_env *env, const mm_state_host *src);  /* synchronous; also refreshes obs / n_agents */
int mm_get_shield_diag(mm_env *env, mm_shield_diag_host *dst);   /* requires record_diag */
int mm_stats(mm_env *env, mm_stats_t *out, int reset);    /* synchronous */

/* The shield alone: safety_layer(safety_type, action, vehicle, dt, ...) (highway_env/vehicle/safety/decentral_layer.py:
 * 767-817, as MDPLCVehicle.step calls it through get_safe_action, safe_controller.py:229-250) for every CAV of the current
 * scenes, against the scenes as they are - the view the front-most vehicle of a sub-step has; nothing is stepped and no
 * state is written (the reference's side effects on the vehicle - is_collaborating, is_lc_safe, target_lane_index,
 * min_headway, the on-ramp HDV record shift - are returned / kept local instead).  Inputs: the clipped low-level action
 * per CAV, nom_steer / nom_acc [n_envs][MM_MAXV] f64 DEVICE.  Outputs [n_envs][MM_MAXV], DEVICE: the shielded action,
 * the time headway the shield computed, and the status the reference logs: whether the shield ran (its gate: fg_params
 * set and two history records), the neighbour classification of multi_agent_state (slot ids, MM_NB_NONE, MM_NB_OBSTACLE),
 * constrain_adj, the QP active set (MM_ACT_*), is_lc_safe. */
typedef struct {
    double *safe_steer, *safe_acc, *min_headway;
    int32_t *ran, *leader, *front_adj, *rear_adj, *constrain_adj, *active, *is_lc_safe;
} mm_shield_query_out;
int mm_shield_query(mm_env *env, const double *nom_steer, const double *nom_acc, const mm_shield_query_out *out, void *stream);

/* The CBF-QP alone (cbf.py:110-161 through cvxopt.solvers.qp): n independent solves on DEVICE arrays;
 * u and active may alias nothing else.  Used for the solves/s microbenchmark and the QP known-answer tests. */
int mm_shield_qp(const double *a, const double *c_lead, const double *c_adj, const uint8_t *has_adj,
                 const double *lo, const double *hi, int64_t n, double *u, uint8_t *active, void *stream);

/* Caller side of the path (BASELINE configs[3]: policy + env on the device).
 *
 * mm_actor_sample: the shared MAPPO actor (marl/single_agent/Model_common.py:5-23: Linear 30-128, ReLU, Linear 128-128,
 * ReLU, Linear 128-5, log-softmax) evaluated on n_rows observation rows [n_rows][30] f32 and, fused with it, the
 * exploration draw of marl/mappo.py:209-228 (np.random.choice(p = softmax): inverse CDF of one uniform per row;
 * here Philox4x32-10 keyed by (seed, step, row)).  Weights are torch.nn.Linear parameters as they lie in memory
 * (weight [out][in], bias [out]), TF32 tensor-core math with fp32 accumulation: the two 128-wide layers run as
 * tcgen05.mma kind::tf32 with the accumulators in TMEM (one persistent CTA per SM, 128 rows per tile), the 128 -> 5
 * output layer, the log-softmax and the draw in the thread that owns the row.  n_agents (nullable) [n_rows / 12]:
 * rows whose slot index (row % 12) is >= n_agents[row / 12] get action 1 (IDLE).  action_mask (nullable) [n_rows] u8,
 * bit k = action k available (the env's action_mask buffer): the invalid-action masking of the MAPPO_GI actor
 * (marl/single_agent/Model_gi.py:63-66, logits of unavailable actions = -1e8 before the log-softmax).  logp_all [n_rows][5] and
 * logp_sel [n_rows] are optional outputs (log-probabilities of all actions / of the drawn one).  All pointers are
 * DEVICE pointers; enqueued on `stream`, no synchronisation. */
int mm_actor_sample(const float *obs, const int32_t *n_agents, int64_t n_rows, const float *w1, const float *b1,
                    const float *w2, const float *b2, const float *w3, const float *b3, uint64_t seed, uint64_t step,
                    const uint8_t *action_mask, int8_t *actions, float *logp_all, float *logp_sel, void *stream);

/* The same fused forward + draw for a 30 - h1 - 128 - 5 network, fp16 tensor-core operands with fp32 accumulation:
 * h1 = 128 is the MAPPO actor above; h1 = 160 is the shared-trunk network of MAPPO_GI with state_split
 * (marl/single_agent/Model_gi.py:137-220): its three first-layer blocks fc11 / fc12 / fc13 over fixed column lists are
 * one 30 -> 160 layer whose weight w1 [160][30] is zero outside the blocks (the caller scatters them; b1 = the three
 * biases concatenated), w2 = fc2.weight [128][160], w3 / b3 = actor_linear.  value_w [128] / value_b [1] (nullable) =
 * critic_linear: values [n_rows] receives V(s) from the same hidden activations (the bootstrap value of
 * mappo_gi.py:396-404).  Rollout-buffer appends of MAPPO.interact (mappo.py:117-131), fused: obs_copy (nullable)
 * [n_rows][30] receives the rows the actions were drawn from (point it at slot t of the state buffer), live_out (nullable)
 * [n_rows] u8 = 1 where the row belongs to an agent that exists (needs n_agents), and `actions` can be slot t of the
 * action buffer.  Everything else as in mm_actor_sample. */
int mm_actor_sample_mlp(const float *obs, const int32_t *n_agents, int64_t n_rows, int h1, const float *w1, const float *b1,
                        const float *w2, const float *b2, const float *w3, const float *b3, const float *value_w,
                        const float *value_b, uint64_t seed, uint64_t step, const uint8_t *action_mask, int8_t *actions,
                        float *logp_all, float *logp_sel, float *values, float *obs_copy, uint8_t *live_out, void *stream);

/* Implementation switch of mm_actor_sample (process-wide): 0 = tcgen05 with fp16 operands, one warpgroup per 128-row
 * tile and the weights once per SM (default, the kernel of mm_actor_sample_mlp); 1 = the warp-level mma.sync TF32 kernel
 * kept as an independent cross-check; 2 = the first tcgen05 kernel (TF32 operands, one CTA per SM); 3 = tcgen05 with
 * fp16 operands, two CTAs per SM (also reachable through mm_actor_sample_mlp while selected). */
int mm_set_actor_impl(int impl);

/* mm_discounted_returns: MAPPO._discount_reward (marl/mappo.py:364-370) for every (env, agent) column of a rollout at
 * once: out[t][c] = rewards[t][c] + gamma * out[t+1][c], restarted after a step with dones[t][c / cols_per_env] != 0,
 * seeded after the last step with final_value[c] (nullable: 0).  rewards / out [T][n_cols] f32, dones
 * [T][n_cols / cols_per_env] u8, DEVICE pointers. */
int mm_discounted_returns(const float *rewards, const uint8_t *dones, const float *final_value, float gamma, int T,
                          int64_t n_cols, int cols_per_env, float *out, void *stream);

/* The step kernel (the physics of a policy step) exists in several builds of one source: generic for 3 CTAs per SM
 * (6 staged fields, 168 registers; the faster one per unit of work), generic for 4 CTAs per SM (4 staged fields, 128
 * registers), and two builds specialised at compile time for all-CAV envs of env id merge-multi-agent-v1 with
 * lateral_control "steer" under the MASS / the HSS shield (no IDM / MOBIL code, configuration reads folded to constants).
 * A fifth build maps half a warp to an env and a lane to a vehicle (the warp-cooperative build, merge_coop.cu) for
 * batches too small to fill the machine with one thread per env.
 * 0 (default): automatic - for all-CAV envs of the plain LC env the warp-cooperative build up to 12 288 envs and a
 * specialised build above (when none of the handle's envs can hold an HDV: spawned under traffic_type cav, or checked by
 * mm_set_state); the 4-CTA build for small grids whose wave structure
 * favours it (e.g. 65 536 envs = 512 CTAs on 148 SMs: one wave instead of a full and an almost empty one, -27 % step
 * time).  3 / 4: force a generic build; 5: automatic among the generic builds only; 6: 4 CTAs per SM forced, specialised
 * builds allowed; 7: the warp-cooperative build forced where it applies; 8: automatic among the one-thread-per-env builds
 * (process-wide; tests and A/B timing). */
int mm_set_step_variant(int variant);
enum { MM_BUILD_GENERIC3 = 3, MM_BUILD_GENERIC4 = 4, MM_BUILD_SPEC_HSS = 31, MM_BUILD_SPEC_MASS = 32,
       MM_BUILD_SPEC_HSS4 = 41, MM_BUILD_SPEC_MASS4 = 42 /* specialised and built for 4 CTAs per SM */,
       MM_BUILD_COOP = 50 /* + shield kind: the warp-cooperative build (half a warp per env) */ };
/* which build the handle's last mm_step / mm_step_host* launched (0 before the first step) */
int mm_step_build(const mm_env *env);

/* The baseline supervisors of safety_guarantee = priority | dmc (highway_env/vehicle/safety/central_layer.py:16-178,
 * decentralised_dmc.py:70-198; called from AbstractEnv.step before _simulate, abstract.py:459-467).  With
 * mm_config.supervisor set, mm_step / mm_step_host* run them on every policy step: the supervised tuples land in
 * mm_buffers.new_actions and are what the step executes.  The reference draws np.random.rand() numbers for them (one per
 * CAV for the priority tie-break, then two per IDM decision of the look-ahead); the batched path draws from
 * Philox4x32-10 keyed (seed, env, episode, policy step).  For teacher forcing and for the seed-exact single-env adapter
 * the caller can supply the draws: mm_set_supervisor_draws(draws_dev [n_envs][MM_SUPERVISOR_DRAWS] f64 DEVICE, consumed in
 * order; NULL returns to Philox) applies to every following step until changed, and mm_supervisor_draws_used copies to
 * the host how many draws each env consumed in the last step (the adapter advances its MT19937 replay by that).
 *
 * mm_supervise is the supervisor alone: rewrites actions [n_envs][MM_MAXV] int8 DEVICE in place for the current scenes;
 * kind 0 = priority, 1 = dmc; draws_dev as above (NULL: Philox); n_used_dev (nullable) [n_envs] int32 DEVICE.  It returns
 * the reference's tuples on every step of the reference fixtures (tests/test_zz_supervisor_gpu.py). */
#define MM_SUPERVISOR_DRAWS 32
int mm_supervise(mm_env *env, int kind, int8_t *actions_dev, const double *draws_dev, int32_t *n_used_dev, void *stream);
int mm_set_supervisor_draws(mm_env *env, const double *draws_dev);
int mm_supervisor_draws_used(mm_env *env, int32_t *n_used_host);

/* Launch bookkeeping for bench.py ("gpu_launches") */
int64_t mm_kernel_launches(const mm_env *env);
const char *mm_last_error(void);
const char *mm_version(void);

#ifdef __cplusplus
}
#endif
#endif
