// supervisor.cu — the baseline supervisors `priority` / `dmc` (central_layer.py, decentralised_dmc.py) on the device:
// one thread per env loads its scene from the tiled state, runs supervisor_core.h and rewrites the env's action tuple.
// The logic is the host-verified source (tests/test_host_cpu.py::test_supervisor_core_*); on a B200 the kernel returns the
// reference's tuples on every step of both fixtures (tests/test_zz_supervisor_gpu.py).  Nothing in the step path calls
// it yet and the config layer still rejects the two values.
#include <cuda_runtime.h>

#include "mm_internal.h"
#include "supervisor_core.h"

namespace mm {

__global__ void __launch_bounds__(64) supervisor_kernel(DevState st, int n_envs, int kind, int8_t *actions, const double *draws,
                                                        int draws_per_env, double headway_time) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_envs) return;
    const uint32_t ei = st.einfo[e];
    const int n = (ei >> EI_NVEH_SHIFT) & EI_4BIT, n_cav = (ei >> EI_NCAV_SHIFT) & EI_4BIT;
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
    for (int i = 0; i < n_cav; ++i) act[i] = actions[(size_t)e * MAXV + i];
    const double *d = draws + (size_t)e * draws_per_env;
    if (kind == 0) mmsup::priority_supervisor(road, orig, n, n_cav, act, d, headway_time);
    else mmsup::dmc_supervisor(road, orig, n, n_cav, act, d, headway_time);
    for (int i = 0; i < n_cav; ++i) actions[(size_t)e * MAXV + i] = (int8_t)act[i];
}

void launch_supervisor(const DevState &st, int n_envs, int kind, int8_t *actions, const double *draws, int draws_per_env,
                       double headway_time, void *stream) {
    // ~17 KB of local memory per thread (12 vehicles x 18 trajectory points, scene + working copy): small CTAs;
    // cudaLimitStackSize is raised by the caller (capi.cu)
    const int block = 64, grid = (n_envs + block - 1) / block;
    supervisor_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(st, n_envs, kind, actions, draws, draws_per_env, headway_time);
}

}  // namespace mm
