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

#include "mm_internal.h"
#include "mm_philox.cuh"
#include "supervisor_core.h"

namespace mm {

__global__ void __launch_bounds__(64) supervisor_kernel(DevState st, int env_offset, int env_count, int kind, int8_t *actions,
                                                        const double *draws, int draws_per_env, double headway_time,
                                                        uint64_t seed, int32_t *n_used_out) {
    const int local = blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= env_count) return;
    const size_t e = (size_t)env_offset + local;
    const uint32_t ei = st.einfo[e];
    const int n = (ei >> EI_NVEH_SHIFT) & EI_4BIT, n_cav = (ei >> EI_NCAV_SHIFT) & EI_4BIT;
    if (n_used_out) n_used_out[e] = 0;
    if (n_cav == 0) return;
    mmsup::Veh orig[mmsup::MAXV], road[mmsup::MAXV];
    for (int i = 0; i < n; ++i) {
        mmsup::Veh &v = orig[i];
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
        road[i] = v;
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
    if (kind == 0) mmsup::priority_supervisor(road, orig, n, n_cav, act, d, headway_time, &used);
    else mmsup::dmc_supervisor(road, orig, n, n_cav, act, d, headway_time, &used);
    for (int i = 0; i < n_cav; ++i) actions[e * MAXV + i] = (int8_t)act[i];
    if (n_used_out) n_used_out[e] = used;
}

void launch_supervisor(const DevState &st, int env_offset, int env_count, int kind, int8_t *actions, const double *draws,
                       int draws_per_env, double headway_time, uint64_t seed, int32_t *n_used_out, void *stream) {
    // ~17 KB of local memory per thread (12 vehicles x 18 trajectory points, scene + working copy): small CTAs;
    // cudaLimitStackSize is raised by the caller (capi.cu)
    const int block = 64, grid = (env_count + block - 1) / block;
    if (grid <= 0) return;
    supervisor_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(st, env_offset, env_count, kind, actions, draws, draws_per_env,
                                                               headway_time, seed, n_used_out);
}

}  // namespace mm
