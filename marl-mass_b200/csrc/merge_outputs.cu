// merge_outputs.cu — observation / reward / info kernel of the batched merge environment (sm_100a).
//
// What MergeEnv.step returns after _simulate (envs/merge_env_v1.py:126-166, abstract.py:469-498): the KinematicLC
// observation rows, local / regional / global rewards, terminal flags, info scalars, action masks and the episode
// statistics.  Unlike the physics of a policy step (merge_step.cu: a serial chain of 33 vehicle moves per env, one thread
// per env) all of this is independent per AGENT, so it runs as its own kernel with one thread per (env, vehicle slot):
//
//   CTA = 32 envs x 11 slots = 352 threads; warp w owns slot w of the 32 envs, lane c owns env column c.
//   * every thread loads its own vehicle's hot fields (x y heading speed cos sin) and flags into shared-memory planes
//     [field][slot][32] - a warp's load is one fully used 256-byte run of the tile;
//   * the x-descending order every neighbour walk starts from is built by counting: a thread's rank is the number of
//     vehicles ahead of its own (ties by slot id, i.e. the reference's stable sort, road.py:277), published with one
//     shared-memory atomicOr per thread;
//   * observation rows are staged in shared memory ([32][362] floats, padded so that 64-bit stores of a half-warp hit
//     16 distinct bank pairs) and leave for HBM as coalesced streaming stores: only the rows of the agents that exist
//     on the step path (absent rows stay zero from the (re)spawn);
//   * local rewards go through shared memory to the regional-reward pass; the per-env sums are taken sequentially in
//     slot order by the env's slot-0 thread (same rounding order as the reference's Python sum);
//   * per-env scalars and the statistics row of the CTA's 32 envs are spread over four warps (reward / speed / traffic
//     speed / headway), each folding its column with warp shuffles: no atomics, no single slow warp.
//
// The same kernel with WITH_REWARDS = false writes the first observation after a (re)spawn or mm_set_state.
#include <cuda_runtime.h>
#include <math_constants.h>
#include "mm_internal.h"

#define MM_KNS mmo
#define MM_NHOT 6
#define MM_PW 32
#define MM_OUT_FN __forceinline__
#include "mm_device.cuh"

namespace mmo {

constexpr int OENVS = 32;                  // env columns per CTA
constexpr int OWARPS = SMV;                // one warp per staged vehicle slot
constexpr int OTHREADS = OENVS * OWARPS;   // 352
constexpr int OBS_STRIDE = MAXV * NS + 2;  // floats per env in the staging block (362: conflict-free 64-bit stores)
constexpr size_t OUT_SMEM = (size_t)PLANES_F64 * sizeof(double)            // hot planes + flags
                            + (size_t)OENVS * OBS_STRIDE * sizeof(float)    // observation rows
                            + (size_t)2 * SMV * OENVS * sizeof(double)      // local rewards, headway terms
                            + (size_t)2 * OENVS * sizeof(unsigned long long)   // x-order and its inverse
                            + (size_t)2 * OENVS * MAXV * sizeof(float)      // agents / regional rewards
                            + (size_t)2 * OENVS * MAXV                      // agents_dones, action masks
                            + (size_t)OENVS * sizeof(int);                  // rows to write per env (-1: env not written)
// 74 880 bytes: three CTAs per SM fit the 227 KB of shared memory

#ifndef MM_OUT_MIN_BLOCKS
#define MM_OUT_MIN_BLOCKS 3   // measured: 2 -> 3 CTAs per SM (56 registers) -2.8 % per policy step at 2^20 envs
#endif

template <bool WITH_REWARDS>
__global__ void __launch_bounds__(OTHREADS, MM_OUT_MIN_BLOCKS) outputs_kernel(const __grid_constant__ StepParams p) {
    const int c = threadIdx.x & 31, i = threadIdx.x >> 5;
    const int local = blockIdx.x * OENVS + c;
    bool valid = local < p.env_count;
    const size_t e = (size_t)p.env_offset + (valid ? local : 0);
    if (p.obs_mask) {   // refresh after a masked (re)spawn: only the envs whose byte is set; most CTAs have none
        valid = valid && p.obs_mask[e] != 0;
        if (!__syncthreads_or(valid)) return;
    }
    float *s_obs = reinterpret_cast<float *>(sm_planes + PLANES_F64);
    double *s_local = reinterpret_cast<double *>(s_obs + OENVS * OBS_STRIDE);
    double *s_hw = s_local + SMV * OENVS;
    unsigned long long *s_live = reinterpret_cast<unsigned long long *>(s_hw + SMV * OENVS), *s_pos = s_live + OENVS;
    float *s_ar = reinterpret_cast<float *>(s_pos + OENVS), *s_rr = s_ar + OENVS * MAXV;
    uint8_t *s_ad = reinterpret_cast<uint8_t *>(s_rr + OENVS * MAXV), *s_am = s_ad + OENVS * MAXV;
    int *s_rows = reinterpret_cast<int *>(s_am + OENVS * MAXV);

    Env ev;
    ev.tid = c;
    ev.g = p.st.f64 + f64_index(e, 0, 0);
    const uint32_t ei = valid ? p.st.einfo[e] : 0u;
    ev.n_veh = (ei >> EI_NVEH_SHIFT) & EI_4BIT;
    ev.n_cav = (ei >> EI_NCAV_SHIFT) & EI_4BIT;
    const int n_merge = (ei >> EI_NMERGE_SHIFT) & EI_4BIT;
    const int steps = (ei >> EI_STEPS_SHIFT) & EI_STEPS_MASK;
    const int n_obs = n_observed(ev, p.cfg);      // the controlled vehicles; every vehicle in the all-HDV env
    const bool has_v = valid && i < ev.n_veh;
    if (i == 0) {
        s_live[c] = 0ull;
        s_pos[c] = 0ull;
        s_rows[c] = valid ? (WITH_REWARDS ? n_obs : MAXV) : -1;
    }
    if (has_v) {
#pragma unroll
        for (int f = 0; f < N_HOT; ++f) SMF(f, i) = __ldcs(tile_ptr(ev.g, f, i));
        FL(i) = __ldcs(p.st.flags + flags_index(e, i));
    }
    __syncthreads();
    if (has_v) {
        // road.py:277: stable sort by x, descending - rank = vehicles ahead, equal x ordered by slot id
        const double xi = X(i);
        int rank = 0;
        for (int j = 0; j < ev.n_veh; ++j) {
            const double xj = X(j);
            rank += (xj > xi || (xj == xi && j < i)) ? 1 : 0;
        }
        atomicOr(&s_live[c], (unsigned long long)i << (4 * rank));
        atomicOr(&s_pos[c], (unsigned long long)rank << (4 * i));
    }
    __syncthreads();
    ev.live = s_live[c];
    ev.pos = s_pos[c];

    const bool sv = p.cfg.steer_vel && !p.cfg.env_v0;
    const bool is_agent = valid && i < n_obs;
    {
        float2 *row = reinterpret_cast<float2 *>(s_obs + c * OBS_STRIDE + i * NS);
        uint32_t nbw = 0xFFFFu;
        if (is_agent) {
            nbw = observe_agent(ev, i, sv, row);
        } else if (!WITH_REWARDS) {
#pragma unroll
            for (int q = 0; q < NS / 2; ++q) row[q] = make_float2(0.f, 0.f);
        }
        if (!WITH_REWARDS && i == SMV - 1) {   // slot 11 is never occupied
#pragma unroll
            for (int q = 0; q < NS / 2; ++q) row[NS / 2 + q] = make_float2(0.f, 0.f);
        }
        if (p.out.veh != nullptr && valid) {
            // packed-state outputs (mm_step_host_packed): what the observation rows are a function of
            if (has_v)      // one aligned 16-byte store per vehicle
                __stcs(reinterpret_cast<float4 *>(p.out.veh) + (e * MAXV + i),
                       make_float4((float)X(i), (float)Y(i), (float)H(i), (float)V(i)));
            p.out.nbr[e * MAXV + i] = (uint16_t)nbw;
            if (i == SMV - 1) p.out.nbr[e * MAXV + MAXV - 1] = 0xFFFFu;
        }
    }
    {
        // _get_available_actions (abstract.py:219-240): IDLE always; LANE_LEFT only from bc1 when bc0 is reachable (the
        // one non-forbidden side lane of the network); FASTER / SLOWER by the speed index
        uint32_t bits = 0;
        if (valid && i < ev.n_cav) {
            const uint32_t f = FL(i);
            bits = 1u << A_IDLE;
            if (fl_lane(f) == L_BC1) {
                double s = lane_s(L_BC0, X(i)), r = Y(i) - c_lane_sy[L_BC0];
                if (fabs(r) <= 2 * LWIDTH && 0 <= s && s < c_lane_len[L_BC0] + VLEN) bits |= 1u << A_LANE_LEFT;
            }
            const int sidx = (int)((f >> FL_SIDX_SHIFT) & FL_3BIT);
            if (sidx < 4) bits |= 1u << A_FASTER;
            if (sidx > 0) bits |= 1u << A_SLOWER;
        }
        s_am[c * MAXV + i] = (uint8_t)bits;
        if (i == SMV - 1) s_am[c * MAXV + MAXV - 1] = 0;
    }

    double loc = 0.0;
    if (WITH_REWARDS) {
        if (is_agent) {
            double hd = headway_distance(ev, i);
            loc = agent_reward(ev, p.cfg, i, hd);
            s_local[i * OENVS + c] = loc;
            // merge_env_v1.py:373-386: time headway to the vehicle ahead or the obstacle
            const double ex = X(i);
            if (fabs(OBST_Y - Y(i)) <= 2 && OBST_X > ex) {
                double d = OBST_X - ex;
                if (d < hd) hd = d;
            }
            hd = hd - VLEN;
            const double vxi = V(i) * CH(i);
            s_hw[i * OENVS + c] = hd / (vxi > 1 ? vxi : 1);
        }
        __syncthreads();
        float lr = 0.f, rr = 0.f;
        uint8_t ad = 0;
        if (is_agent && p.cfg.env_hdv) {
            // MergeEnvLCHDV.step (merge_env_v1.py:603-665): no regional rewards / per-agent dones; the per-vehicle
            // reward terms are kept for inspection
            lr = (float)loc;
        } else if (is_agent) {
            // regional reward (merge_env_v1.py:91-124): own-lane group plus, where one exists, the group across
            const int lane = fl_lane(FL(i));
            const bool on_main = lane == L_AB0 || lane == L_BC0 || lane == L_CD0;
            int across = -1;
            if (lane == L_BC0) across = L_BC1;
            else if (lane == L_AB0 && X(i) > 220) across = L_KB0;
            else if (lane == L_BC1) across = L_BC0;
            else if (lane == L_KB0) across = L_AB0;
            int fo, ro, fa, ra;
            surrounding2(ev, i, visible_lanes(lane), across >= 0 ? visible_lanes(across) : 0u, fo, ro, fa, ra);
            const int fl_ = on_main ? fo : fa, rl = on_main ? ro : ra, fr = on_main ? fa : fo, rrr = on_main ? ra : ro;
            const int cand[5] = {fl_, fr, i, rl, rrr};
            double sum = 0;
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int j = cand[k];
                if (j >= 0 && j < ev.n_cav) { sum = sum + s_local[j * OENVS + c]; ++cnt; }
            }
            lr = (float)loc;
            rr = (float)(sum / cnt);
            ad = ((FL(i) & FL_CRASHED) || steps >= p.cfg.duration_steps || X(i) < 0) ? 1 : 0;
        }
        s_ar[c * MAXV + i] = lr;
        s_rr[c * MAXV + i] = rr;
        s_ad[c * MAXV + i] = ad;
        if (i == SMV - 1) {
            s_ar[c * MAXV + MAXV - 1] = 0.f;
            s_rr[c * MAXV + MAXV - 1] = 0.f;
            s_ad[c * MAXV + MAXV - 1] = 0;
        }
        // per-env scalars (merge_env_v1.py:126-166, 517-524; abstract.py:489-498) and the statistics row of the CTA's 32
        // envs (one row of partial sums per 32 envs, no atomics; mm_stats() folds the rows).  The sums run in slot order
        // (the rounding order of the reference's Python sums); the four quantities go to four different warps so that no
        // single warp holds the CTA back at the final barrier.
        double *stat_row = p.out.stats + (((size_t)p.env_offset + (size_t)blockIdx.x * OENVS) >> 5) * N_STATS;
        const unsigned full = 0xffffffffu;
        auto fold_sum = [&](double x) {
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(full, x, off);
            return x;
        };
        if (i == 0) {          // reward, done, merge percent; episode counters
            double rsum = 0;
            bool any_crash = false;
            int n_rem = 0;
            for (int k = 0; k < n_obs; ++k) {
                rsum += s_local[k * OENVS + c];
                const uint32_t fk = FL(k);
                any_crash = any_crash || (fk & FL_CRASHED);
                const int lane = fl_lane(fk);
                if (lane == L_BC1 || lane == L_KB0 || lane == L_JK0) ++n_rem;
            }
            double reward = 0, mp = -1.0;
            bool done = false;
            if (valid) {
                done = is_terminal(ev, steps, p.cfg);
                reward = rsum / n_obs;
                if (done) mp = n_merge > 0 ? (double)(n_merge - n_rem) / n_merge * 100 : 100.0;
                p.out.reward[e] = (float)reward;
                p.out.done[e] = done ? 1 : 0;
                p.out.merge_percent[e] = (float)mp;
                p.out.n_agents[e] = n_obs;
            }
            const int agents = __reduce_add_sync(full, valid ? n_obs : 0), envs = __reduce_add_sync(full, valid ? 1 : 0);
            const int episodes = __reduce_add_sync(full, done ? 1 : 0), crashed = __reduce_add_sync(full, (done && any_crash) ? 1 : 0);
            const double rew = fold_sum(reward), mrg = fold_sum(done ? mp : 0.0);
            if (c == 0) {
                stat_row[ST_AGENT_STEPS] += (double)agents; stat_row[ST_ENV_STEPS] += (double)envs;
                stat_row[ST_EPISODES] += (double)episodes; stat_row[ST_CRASHED] += (double)crashed;
                stat_row[ST_REWARD] += rew; stat_row[ST_MERGE] += mrg;
            }
        } else if (i == 1) {   // average speed of the observed vehicles
            double ssum = 0;
            for (int k = 0; k < n_obs; ++k) ssum += V(k);
            const double avg = valid ? ssum / n_obs : 0.0;
            if (valid) p.out.average_speed[e] = (float)avg;
            const double spd = fold_sum(avg);
            if (c == 0) stat_row[ST_SPEED] += spd;
        } else if (i == 2) {   // traffic speed: every vehicle
            double tsum = 0;
            for (int k = 0; k < ev.n_veh; ++k) tsum += V(k);
            if (valid) p.out.traffic_speed[e] = (float)(tsum / ev.n_veh);
        } else if (i == 3) {   // smallest time headway
            double minhw = CUDART_INF;
            for (int k = 0; k < n_obs; ++k) minhw = fmin(minhw, s_hw[k * OENVS + c]);
            if (valid) p.out.min_headway[e] = (float)minhw;
            double x = valid ? minhw : CUDART_INF;
            for (int off = 16; off > 0; off >>= 1) x = fmin(x, __shfl_down_sync(full, x, off));
            if (c == 0) stat_row[ST_MINHW] = fmin(stat_row[ST_MINHW], x);
        }
    } else if (i == 0 && valid) {
        p.out.n_agents[e] = n_obs;
    }
    __syncthreads();

    // coalesced write-out of the CTA's blocks.  Observation rows: warp w takes envs w, w + 11, w + 22.
    const size_t e0 = (size_t)p.env_offset + (size_t)blockIdx.x * OENVS;
    for (int cc = i; cc < OENVS; cc += OWARPS) {
        const int rows = s_rows[cc];
        if (rows < 0) continue;
        const float2 *src = reinterpret_cast<const float2 *>(s_obs + cc * OBS_STRIDE);
        float2 *dst = reinterpret_cast<float2 *>(p.out.obs + (e0 + cc) * (size_t)(MAXV * NS));
        for (int q = c; q < rows * (NS / 2); q += 32) __stcs(dst + q, src[q]);
    }
    for (int k = threadIdx.x; k < OENVS * MAXV; k += OTHREADS) {
        if (s_rows[k / MAXV] < 0) continue;
        const size_t g = e0 * MAXV + k;
        p.out.action_mask[g] = s_am[k];
        if (WITH_REWARDS) {
            p.out.agents_rewards[g] = s_ar[k];
            p.out.regional_rewards[g] = s_rr[k];
            p.out.agents_dones[g] = s_ad[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// packed-state rows for the host path (mm_step_host_packed)
// ------------------------------------------------------------------------------------------------
// Pass 1: exclusive scans of n_veh and n_agents over the chunk (one CTA), chained to the totals of the previous chunk.
// Pass 2: a warp moves one env's vehicle rows and neighbour words to their packed place.  Offsets are absolute and
// dense over the whole batch, so the host derives them from the two count arrays alone.
__global__ void __launch_bounds__(1024) packed_offsets_kernel(const uint32_t *__restrict__ einfo, const int32_t *__restrict__ n_agents,
                                                              int count, const int64_t *__restrict__ base_in,
                                                              int64_t *__restrict__ base_out, int64_t *__restrict__ base_out_host,
                                                              int32_t *__restrict__ voff, int32_t *__restrict__ aoff,
                                                              uint8_t *__restrict__ n_veh_u8, uint8_t *__restrict__ n_agents_u8) {
    // One CTA, 32 warps; warp w owns the contiguous segment [w * seg, (w + 1) * seg) and walks it 32 envs at a time with
    // coalesced loads and a warp scan.  Pass 1 sums the segments, the 32 sums are scanned, pass 2 writes the offsets.
    // (The first version gave every thread a private run of 54 envs: strided loads and stores, 192 us per chunk, and the
    // chunks' scans are chained - 3.6 ms of a 2^20-env step.)
    __shared__ int wv[32], wa[32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg = ((count + 31) / 32 + 31) / 32 * 32;
    const int lo = min(warp * seg, count), hi = min(lo + seg, count);
    int sv = 0, sa = 0;
    for (int e = lo + lane; e < hi; e += 32) {
        sv += (int)((einfo[e] >> EI_NVEH_SHIFT) & EI_4BIT);
        sa += n_agents[e];
    }
    sv = __reduce_add_sync(full, sv);
    sa = __reduce_add_sync(full, sa);
    if (lane == 0) { wv[warp] = sv; wa[warp] = sa; }
    __syncthreads();
    if (warp == 0) {
        const int a = wv[lane], b = wa[lane];
        int xa = a, xb = b;
        for (int off = 1; off < 32; off <<= 1) {
            const int u = __shfl_up_sync(full, xa, off), w = __shfl_up_sync(full, xb, off);
            if (lane >= off) { xa += u; xb += w; }
        }
        wv[lane] = xa - a;
        wa[lane] = xb - b;
        if (lane == 31) {     // totals of this chunk, chained; the host's copy goes straight to mapped pinned memory
            const int64_t tv = base_in[0] + xa, ta = base_in[1] + xb;
            base_out[0] = tv;
            base_out[1] = ta;
            base_out_host[0] = tv;
            base_out_host[1] = ta;
            __threadfence_system();
        }
    }
    __syncthreads();
    int rv = wv[warp], ra = wa[warp];
    for (int e0 = lo; e0 < hi; e0 += 32) {
        const int e = e0 + lane;
        int nv = 0, na = 0;
        if (e < hi) {
            nv = (int)((einfo[e] >> EI_NVEH_SHIFT) & EI_4BIT);
            na = n_agents[e];
        }
        int iv = nv, ia = na;
        for (int off = 1; off < 32; off <<= 1) {
            const int u = __shfl_up_sync(full, iv, off), w = __shfl_up_sync(full, ia, off);
            if (lane >= off) { iv += u; ia += w; }
        }
        if (e < hi) {
            voff[e] = rv + iv - nv;
            aoff[e] = ra + ia - na;
            n_veh_u8[e] = (uint8_t)nv;
            n_agents_u8[e] = (uint8_t)na;
        }
        rv += __shfl_sync(full, iv, 31);
        ra += __shfl_sync(full, ia, 31);
    }
}

__global__ void __launch_bounds__(256) packed_copy_kernel(const float *__restrict__ veh, const uint16_t *__restrict__ nbr,
                                                          const uint8_t *__restrict__ n_veh_u8, const uint8_t *__restrict__ n_agents_u8,
                                                          const int32_t *__restrict__ voff, const int32_t *__restrict__ aoff, int count,
                                                          const int64_t *__restrict__ base_in, float *__restrict__ veh_packed,
                                                          uint16_t *__restrict__ nbr_packed) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int64_t bv = base_in[0], ba = base_in[1];
    for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < count; e += warps) {
        const int nf = (int)n_veh_u8[e] * MM_VEH_F32;
        const float *src = veh + (size_t)e * MAXV * MM_VEH_F32;
        float *dst = veh_packed + (bv + voff[e]) * MM_VEH_F32;
        for (int k = lane; k < nf; k += 32) dst[k] = __ldcs(src + k);
        const int na = (int)n_agents_u8[e];
        if (lane < na) nbr_packed[ba + aoff[e] + lane] = nbr[(size_t)e * MAXV + lane];
    }
}

void launch_packed_pack_impl(const uint32_t *einfo, const int32_t *n_agents, const float *veh, const uint16_t *nbr, int count,
                             const int64_t *base_in, int64_t *base_out, int64_t *base_out_host, int32_t *voff, int32_t *aoff,
                             float *veh_packed, uint16_t *nbr_packed, uint8_t *n_veh_u8, uint8_t *n_agents_u8, void *stream) {
    if (count <= 0) return;
    packed_offsets_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(einfo, n_agents, count, base_in, base_out, base_out_host, voff, aoff,
                                                                n_veh_u8, n_agents_u8);
    packed_copy_kernel<<<(count + 7) / 8, 256, 0, (cudaStream_t)stream>>>(veh, nbr, n_veh_u8, n_agents_u8, voff, aoff, count, base_in,
                                                                        veh_packed, nbr_packed);
}

void launch_outputs_impl(const StepParams &p, bool with_rewards, void *stream) {
    static bool ready[MM_MAX_DEVICES] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MM_MAX_DEVICES) return;
    if (!ready[dev]) {
        cudaError_t ce = cudaFuncSetAttribute(outputs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OUT_SMEM);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(outputs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OUT_SMEM);
        if (ce != cudaSuccess) return;   // stays pending for the caller's cudaGetLastError
        ready[dev] = true;
    }
    const int grid = (p.env_count + OENVS - 1) / OENVS;
    if (grid <= 0) return;
    if (with_rewards) outputs_kernel<true><<<grid, OTHREADS, OUT_SMEM, (cudaStream_t)stream>>>(p);
    else outputs_kernel<false><<<grid, OTHREADS, OUT_SMEM, (cudaStream_t)stream>>>(p);
}

}  // namespace mmo

namespace mm {
void launch_outputs(const StepParams &p, bool with_rewards, void *stream) { mmo::launch_outputs_impl(p, with_rewards, stream); }
void launch_packed_pack(const uint32_t *einfo, const int32_t *n_agents, const float *veh, const uint16_t *nbr, int count,
                        const int64_t *base_in, int64_t *base_out, int64_t *base_out_host, int32_t *voff, int32_t *aoff,
                        float *veh_packed, uint16_t *nbr_packed, uint8_t *n_veh_u8, uint8_t *n_agents_u8, void *stream) {
    mmo::launch_packed_pack_impl(einfo, n_agents, veh, nbr, count, base_in, base_out, base_out_host, voff, aoff, veh_packed, nbr_packed,
                                 n_veh_u8, n_agents_u8, stream);
}
}  // namespace mm
