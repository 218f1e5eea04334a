// supervisor.cu — the baseline supervisors `priority` / `dmc` (central_layer.py, decentralised_dmc.py) on the device:
// one thread per env loads its scene from the tiled state, runs supervisor_core.h and rewrites the env's action tuple
// (abstract.py:459-464: `self.new_action = safety_supervisor(env, actions)` before _simulate).  The logic is the
// host-verified source (tests/test_host_cpu.py::test_supervisor_core_*); on a B200 the kernel returns the reference's
// tuples on every step of both fixtures (tests/test_zz_supervisor_gpu.py).
//
// Draws.  The reference takes np.random.rand() numbers from the process-global MT19937 stream: one per CAV for the
// priority tie-break, then two per IDM decision of the look-ahead.  `draws` != null: the caller supplies them in
// consumption order (teacher forcing; the single-env adapter replays the MT19937 stream and advances it by the count this
// kernel reports).  `draws` == null: Philox4x32-10 keyed (seed, env, episode, policy step) - the batched mode.
#include <cuda_runtime.h>
#include <cstdlib>

#include "mm_internal.h"
#include "mm_philox.cuh"
#include "supervisor_core.h"

namespace mm {

__device__ inline mmsup::Veh load_vehicle(const DevState &st, size_t e, int i, int n_cav) {
    mmsup::Veh v;
    const uint32_t f = st.flags[flags_index(e, i)];
    v.x = st.f64[f64_index(e, F_X, i)];
    v.y = st.f64[f64_index(e, F_Y, i)];
    v.heading = st.f64[f64_index(e, F_H, i)];
    v.speed = st.f64[f64_index(e, F_V, i)];
    v.target_speed = st.f64[f64_index(e, F_TSPEED, i)];
    v.steer = v.acc = 0.0;
    v.lane = (int)((f >> FL_LANE_SHIFT) & FL_3BIT);
    v.target_lane = (int)((f >> FL_TLANE_SHIFT) & FL_3BIT);
    v.speed_index = (int)((f >> FL_SIDX_SHIFT) & FL_3BIT);
    v.cav = i < n_cav;
    v.crashed = (f & FL_CRASHED) != 0;
    v.n_traj = 0;
    return v;
}

// `tasks` != null (dmc only): predicted collisions are recorded for dmc_tasks_kernel instead of being evaluated here
__global__ void __launch_bounds__(64) supervisor_kernel(DevState st, int env_offset, int env_count, int kind, int8_t *actions,
                                                        const double *draws, int draws_per_env, double headway_time,
                                                        uint64_t seed, int32_t *n_used_out, mmsup::DmcTask *tasks,
                                                        int *task_count, int task_capacity) {
    const int local = blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= env_count) return;
    const size_t e = (size_t)env_offset + local;
    const uint32_t ei = st.einfo[e];
    const int n = (ei >> EI_NVEH_SHIFT) & EI_4BIT, n_cav = (ei >> EI_NCAV_SHIFT) & EI_4BIT;
    if (n_used_out) n_used_out[e] = 0;
    if (n_cav == 0) return;
    mmsup::Veh orig[mmsup::MAXV], road[mmsup::MAXV];
    for (int i = 0; i < n; ++i) {
        orig[i] = load_vehicle(st, e, i, n_cav);
        road[i] = orig[i];
    }
    int act[mmsup::MAXV];
    for (int i = 0; i < n_cav; ++i) act[i] = actions[e * MAXV + i];
    double own[MM_SUPERVISOR_DRAWS];
    const double *d = draws ? draws + e * draws_per_env : own;
    if (!draws) {
        const uint32_t step = (ei >> EI_STEPS_SHIFT) & EI_STEPS_MASK;
        Philox rng(seed ^ 0x7375'7065'7276'6973ull, (uint64_t)e, st.episode[e], (step + 1u) << 8);
        for (int k = 0; k < MM_SUPERVISOR_DRAWS; ++k) own[k] = rng.uniform();
    }
    int used = 0;
    if (kind == 0) {
        mmsup::priority_supervisor(road, orig, n, n_cav, act, d, headway_time, &used);
    } else {
        mmsup::DmcTaskSink sink{tasks, task_count, task_capacity, local};
        mmsup::dmc_supervisor(road, orig, n, n_cav, act, d, headway_time, &used, tasks ? &sink : nullptr);
    }
    for (int i = 0; i < n_cav; ++i) actions[e * MAXV + i] = (int8_t)act[i];
    if (n_used_out) n_used_out[e] = used;
}

// Pass 2 of dmc.  ncu on the one-pass kernel (profiles/r2_final_supervisor_two_pass_experiments.txt): 4.65 of 32 lanes active
// - a predicted collision makes its env evaluate every available action over the whole horizon (5 x 171 controller
// steps, decentralised_dmc.py:170-198) while the other 31 envs of the warp wait, and the kernel lasts as long as its
// slowest env.  Here every (collision, action) pair is a lane: eight lanes per task (five in use), the same 171-step
// loop in all of them, and the reference's selection rule (first strict maximum in list order) replayed over the
// group's results.  The ego is re-read from the state (the supervisor never modifies it), the neighbours'
// x-trajectories come from the task record.
__global__ void __launch_bounds__(128) dmc_tasks_kernel(DevState st, int env_offset, const mmsup::DmcTask *tasks,
                                                        const int *task_count, int task_capacity, int8_t *actions) {
    const int n_tasks = min(*task_count, task_capacity);
    const int lane = threadIdx.x & 31, a = lane & 7;
    const unsigned group_mask = 0xffu << (lane & 24);
    const int n_groups = gridDim.x * (blockDim.x >> 3);
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; t < n_tasks; t += n_groups) {
        const mmsup::DmcTask &task = tasks[t];
        const size_t e = (size_t)env_offset + task.env;
        const int i = task.cav, n_acts = task.n_acts;
        const int n_cav = (st.einfo[e] >> EI_NCAV_SHIFT) & EI_4BIT;
        double room = 0;
        if (a < n_acts) {
            mmsup::Veh c = load_vehicle(st, e, i, n_cav);
            for (int tt = 0; tt < mmsup::NPTS; ++tt) room += mmsup::check_safety_room_tx(c, task.acts[a], task.tx, task.nb, tt);
        }
        double best_room = 0;
        int best = 0;
        for (int k = 0; k < n_acts; ++k) {
            const double rk = __shfl_sync(group_mask, room, (lane & 24) + k);
            if (k == 0 || rk > best_room) { best_room = rk; best = k; }
        }
        if (a == 0) actions[e * MAXV + i] = (int8_t)task.acts[best];
    }
}

// `tasks` (task_capacity records) and `task_count` (one int) are scratch private to this call (capi.cu hands disjoint
// pieces to the chunks of a host-path step, which run on different streams); null -> one pass
void launch_supervisor(const DevState &st, int env_offset, int env_count, int kind, int8_t *actions, const double *draws,
                       int draws_per_env, double headway_time, uint64_t seed, int32_t *n_used_out, void *tasks,
                       int *task_count, int task_capacity, void *stream) {
    // ~17 KB of local memory per thread (12 vehicles x 18 trajectory points, scene + working copy): small CTAs;
    // cudaLimitStackSize is raised by the caller (capi.cu)
    const int block = 64, grid = (env_count + block - 1) / block;
    if (grid <= 0) return;
    cudaStream_t s = (cudaStream_t)stream;
    mmsup::DmcTask *tk = kind == 1 ? static_cast<mmsup::DmcTask *>(tasks) : nullptr;
    // test knob: MM_SUP_TASK_CAP=<n> shrinks the task list, so that collisions past the n-th are evaluated in place
    static const int cap_override = [] { const char *v = getenv("MM_SUP_TASK_CAP"); return v ? atoi(v) : -1; }();
    if (cap_override >= 0 && cap_override < task_capacity) task_capacity = cap_override;
    if (task_capacity <= 0) tk = nullptr;
    if (tk) cudaMemsetAsync(task_count, 0, sizeof(int), s);
    supervisor_kernel<<<grid, block, 0, s>>>(st, env_offset, env_count, kind, actions, draws, draws_per_env, headway_time, seed,
                                             n_used_out, tk, task_count, task_capacity);
    if (tk) {
        int sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int groups = (task_capacity + 15) / 16;          // 16 eight-lane groups per 128-thread CTA
        const int grid2 = groups < sms * 8 ? groups : sms * 8;
        dmc_tasks_kernel<<<grid2 > 0 ? grid2 : 1, 128, 0, s>>>(st, env_offset, tk, task_count, task_capacity, actions);
    }
}

size_t supervisor_task_bytes() { return sizeof(mmsup::DmcTask); }

}  // namespace mm
