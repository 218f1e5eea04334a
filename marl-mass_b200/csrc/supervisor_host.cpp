// Host harness of supervisor_core.h for the CPU parity test (tests/test_host_cpu.py): one scene in, supervised tuple out.
// g++ -O2 -ffp-contract=off -shared -fPIC supervisor_host.cpp -o libsupervisor_host.so   (not part of the product library)
#include "supervisor_core.h"

extern "C" int mm_supervisor_host(int kind /* 0 priority, 1 dmc */, int n, int n_cav, const double *x, const double *y,
                                  const double *heading, const double *speed, const double *target_speed, const int *lane,
                                  const int *target_lane, const int *speed_index, const int *crashed, int *actions,
                                  const double *draws, double headway_time, int *n_draws_used) {
    using namespace mmsup;
    if (n < 1 || n > MAXV || n_cav < 0 || n_cav > n) return 1;
    Veh orig[MAXV], road[MAXV];
    for (int i = 0; i < n; ++i) {
        Veh &v = orig[i];
        v.x = x[i]; v.y = y[i]; v.heading = heading[i]; v.speed = speed[i]; v.target_speed = target_speed[i];
        v.steer = v.acc = 0.0;
        v.lane = lane[i]; v.target_lane = target_lane[i]; v.speed_index = speed_index[i];
        v.cav = i < n_cav; v.crashed = crashed[i] != 0; v.n_traj = 0;
        road[i] = v;
    }
    if (kind == 0) priority_supervisor(road, orig, n, n_cav, actions, draws, headway_time, n_draws_used);
    else dmc_supervisor(road, orig, n, n_cav, actions, draws, headway_time, n_draws_used);
    return 0;
}
