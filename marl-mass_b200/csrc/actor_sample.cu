// actor_sample.cu — the MAPPO actor (30-128-128-5, ReLU, log-softmax) and its exploration draw, fused in one kernel.
//
// Reference: marl/single_agent/Model_common.py:5-23 (ActorNetwork), marl/mappo.py:209-228 (_softmax_action +
// exploration_action: np.random.choice(p=softmax) = inverse CDF of one uniform).  The reference evaluates the shared
// actor once per agent per step on the CPU; the batched rollout (BASELINE configs[3]) evaluates it for every
// (env, agent) row of the observation buffer the step kernel has just written, 786 432 rows at 65 536 envs.  As three
// torch Linear layers that is 1.7 ms per step against 0.9 ms for the env step itself (profiles/README.md): the hidden
// activations (2 x 400 MB) go through HBM four times.  Here a warp carries 16 rows through all three layers in
// registers and only the observations (120 B / row) and the actions (1 B / row) touch HBM.
//
// Mapping: mma.sync.m16n8k8 TF32 (fp32 accumulate) per warp.  The accumulator fragment of layer L (thread holds rows
// g, g+8 and columns 2t, 2t+1 of each 8-wide tile) is reused directly as the A fragment of layer L+1 (rows g, g+8,
// columns t, t+4) by permuting the K index of that layer: A column t <-> hidden unit 8kk+2t, A column t+4 <-> 8kk+2t+1,
// and the weights are laid out in shared memory in fragment order with the same permutation, so no shuffle or
// shared-memory round trip separates the layers.  This op is ~33 GFLOP per step; legacy warp-level MMA is far from the
// tcgen05 peak but already makes the op a small fraction of the env step, so the simpler pipeline was kept.
#include <cuda_runtime.h>
#include <stdint.h>
#include "mm_internal.h"

namespace mm {

constexpr int AC_IN = 30, AC_HID = 128, AC_OUT = 5;
constexpr int AC_KT1 = 4;              // k-tiles of layer 1 (30 inputs padded to 32)
constexpr int AC_NT = AC_HID / 8;      // 16 n-tiles of the hidden layers = k-tiles of the next layer
constexpr int AC_THREADS = 256;
constexpr int AC_GRP = 8;              // layer-2 n-tiles accumulated together (independent MMA chains per warp)
// shared memory (32-bit words): fragment-ordered TF32 weights, then the fp32 biases
constexpr int AC_W1F = 0;
constexpr int AC_W2F = AC_W1F + AC_KT1 * AC_NT * 64;
constexpr int AC_W3F = AC_W2F + AC_NT * AC_NT * 64;
constexpr int AC_B1 = AC_W3F + AC_NT * 64;
constexpr int AC_B2 = AC_B1 + AC_HID;
constexpr int AC_B3 = AC_B2 + AC_HID;
constexpr int AC_SMEM_WORDS = AC_B3 + 8;

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Philox4x32-10, one counter block per row: (seed, step) is the key, the row index the counter
__device__ __forceinline__ uint32_t philox_row(uint64_t seed, uint64_t step, uint64_t row) {
    uint32_t k0 = (uint32_t)seed ^ (uint32_t)(step * 0x9E3779B97F4A7C15ull), k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 7);
    uint32_t c0 = (uint32_t)row, c1 = (uint32_t)(row >> 32), c2 = (uint32_t)step, c3 = 0x6d6d6173u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

__global__ void __launch_bounds__(AC_THREADS, 1)
actor_sample_kernel(const float *__restrict__ obs, const int32_t *__restrict__ n_agents, int64_t n_rows,
                    const float *__restrict__ w1, const float *__restrict__ b1, const float *__restrict__ w2,
                    const float *__restrict__ b2, const float *__restrict__ w3, const float *__restrict__ b3,
                    uint64_t seed, uint64_t step, int8_t *__restrict__ actions, float *__restrict__ logp_all,
                    float *__restrict__ logp_sel) {
    extern __shared__ __align__(16) uint32_t ac_sm[];
    float *sm_f = reinterpret_cast<float *>(ac_sm);
    const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, t = lane & 3;

    // weights into fragment order (torch Linear weight is [out][in]; B(k, n) = weight[n][k])
    for (int idx = tid; idx < AC_KT1 * AC_NT * 64; idx += AC_THREADS) {
        int r = idx & 1, ln = (idx >> 1) & 31, tile = idx >> 6, j = tile % AC_NT, kk = tile / AC_NT;
        int k = 8 * kk + (ln & 3) + 4 * r, n = 8 * j + (ln >> 2);
        ac_sm[AC_W1F + idx] = to_tf32(k < AC_IN ? w1[n * AC_IN + k] : 0.f);
    }
    for (int idx = tid; idx < AC_NT * AC_NT * 64; idx += AC_THREADS) {
        int r = idx & 1, ln = (idx >> 1) & 31, tile = idx >> 6, j = tile % AC_NT, kk = tile / AC_NT;
        int k = 8 * kk + 2 * (ln & 3) + r, n = 8 * j + (ln >> 2);
        ac_sm[AC_W2F + idx] = to_tf32(w2[n * AC_HID + k]);
    }
    for (int idx = tid; idx < AC_NT * 64; idx += AC_THREADS) {
        int r = idx & 1, ln = (idx >> 1) & 31, kk = idx >> 6;
        int k = 8 * kk + 2 * (ln & 3) + r, n = ln >> 2;
        ac_sm[AC_W3F + idx] = to_tf32(n < AC_OUT ? w3[n * AC_HID + k] : 0.f);
    }
    for (int idx = tid; idx < AC_HID; idx += AC_THREADS) {
        sm_f[AC_B1 + idx] = b1[idx];
        sm_f[AC_B2 + idx] = b2[idx];
    }
    if (tid < 8) sm_f[AC_B3 + tid] = tid < AC_OUT ? b3[tid] : 0.f;
    __syncthreads();

    const int64_t n_tiles = (n_rows + 15) / 16;
    const int64_t warp0 = (int64_t)blockIdx.x * (AC_THREADS / 32) + (tid >> 5);
    const int64_t warp_stride = (int64_t)gridDim.x * (AC_THREADS / 32);
    const uint2 *w1f = reinterpret_cast<const uint2 *>(ac_sm + AC_W1F) + lane;
    const uint2 *w2f = reinterpret_cast<const uint2 *>(ac_sm + AC_W2F) + lane;
    const uint2 *w3f = reinterpret_cast<const uint2 *>(ac_sm + AC_W3F) + lane;

    for (int64_t tile = warp0; tile < n_tiles; tile += warp_stride) {
        const int64_t r_lo = tile * 16 + g, r_hi = r_lo + 8;
        const bool ok_lo = r_lo < n_rows, ok_hi = r_hi < n_rows;
        const float *o_lo = obs + r_lo * AC_IN, *o_hi = obs + r_hi * AC_IN;

        // layer 1: [16 x 32] x [32 x 128]
        uint32_t a1[AC_KT1][4];
#pragma unroll
        for (int kk = 0; kk < AC_KT1; ++kk) {
            const int c0 = 8 * kk + t, c1 = c0 + 4;
            a1[kk][0] = to_tf32(ok_lo && c0 < AC_IN ? o_lo[c0] : 0.f);
            a1[kk][1] = to_tf32(ok_hi && c0 < AC_IN ? o_hi[c0] : 0.f);
            a1[kk][2] = to_tf32(ok_lo && c1 < AC_IN ? o_lo[c1] : 0.f);
            a1[kk][3] = to_tf32(ok_hi && c1 < AC_IN ? o_hi[c1] : 0.f);
        }
        uint32_t h1[AC_NT][4];   // relu(layer 1) as the A fragments of layer 2
#pragma unroll
        for (int j = 0; j < AC_NT; ++j) {
            float acc[4];
            acc[0] = acc[2] = sm_f[AC_B1 + 8 * j + 2 * t];
            acc[1] = acc[3] = sm_f[AC_B1 + 8 * j + 2 * t + 1];
#pragma unroll
            for (int kk = 0; kk < AC_KT1; ++kk) {
                uint2 b = w1f[(kk * AC_NT + j) * 32];
                mma_tf32(acc, a1[kk], b.x, b.y);
            }
            // accumulator (rows g, g+8; cols 2t, 2t+1) -> A fragment (rows g, g+8; k-columns t, t+4) of k-tile j
            h1[j][0] = to_tf32(fmaxf(acc[0], 0.f));
            h1[j][1] = to_tf32(fmaxf(acc[2], 0.f));
            h1[j][2] = to_tf32(fmaxf(acc[1], 0.f));
            h1[j][3] = to_tf32(fmaxf(acc[3], 0.f));
        }
        // layers 2 and 3: each group of 8 hidden n-tiles of layer 2 is consumed by layer 3 as soon as it is complete
        float lg[4];
        lg[0] = lg[2] = sm_f[AC_B3 + 2 * t];
        lg[1] = lg[3] = sm_f[AC_B3 + 2 * t + 1];
#pragma unroll
        for (int jb = 0; jb < AC_NT; jb += AC_GRP) {
            float acc[AC_GRP][4];
#pragma unroll
            for (int q = 0; q < AC_GRP; ++q) {
                acc[q][0] = acc[q][2] = sm_f[AC_B2 + 8 * (jb + q) + 2 * t];
                acc[q][1] = acc[q][3] = sm_f[AC_B2 + 8 * (jb + q) + 2 * t + 1];
            }
#pragma unroll
            for (int kk = 0; kk < AC_NT; ++kk) {
#pragma unroll
                for (int q = 0; q < AC_GRP; ++q) {
                    uint2 b = w2f[(kk * AC_NT + jb + q) * 32];
                    mma_tf32(acc[q], h1[kk], b.x, b.y);
                }
            }
#pragma unroll
            for (int q = 0; q < AC_GRP; ++q) {
                uint32_t h2[4] = {to_tf32(fmaxf(acc[q][0], 0.f)), to_tf32(fmaxf(acc[q][2], 0.f)),
                                  to_tf32(fmaxf(acc[q][1], 0.f)), to_tf32(fmaxf(acc[q][3], 0.f))};
                uint2 b = w3f[(jb + q) * 32];
                mma_tf32(lg, h2, b.x, b.y);
            }
        }

        // logits of row g: columns (2t, 2t+1) in lg[0..1] of the quad's threads; row g+8 in lg[2..3]
        const unsigned full = 0xffffffffu;
        const int q0 = lane & ~3;
        float l_lo[AC_OUT], l_hi[AC_OUT];
#pragma unroll
        for (int k = 0; k < AC_OUT; ++k) {
            const int src = q0 + (k >> 1);
            float lo_e = __shfl_sync(full, lg[0], src), lo_o = __shfl_sync(full, lg[1], src);
            float hi_e = __shfl_sync(full, lg[2], src), hi_o = __shfl_sync(full, lg[3], src);
            l_lo[k] = (k & 1) ? lo_o : lo_e;
            l_hi[k] = (k & 1) ? hi_o : hi_e;
        }
        // thread t = 0 finishes row g, t = 1 row g+8 (the other two threads of the quad idle here)
        if (t < 2) {
            const int64_t row = t == 0 ? r_lo : r_hi;
            if (row < n_rows) {
                float l[AC_OUT];
#pragma unroll
                for (int k = 0; k < AC_OUT; ++k) l[k] = t == 0 ? l_lo[k] : l_hi[k];
                float m = l[0];
#pragma unroll
                for (int k = 1; k < AC_OUT; ++k) m = fmaxf(m, l[k]);
                float e[AC_OUT], S = 0.f;
#pragma unroll
                for (int k = 0; k < AC_OUT; ++k) { e[k] = __expf(l[k] - m); S += e[k]; }
                const float logS = __logf(S);
                // np.random.choice(p): first k whose cumulative probability exceeds one uniform draw
                const float u = (float)(philox_row(seed, step, (uint64_t)row) >> 8) * (1.0f / 16777216.0f);
                const float target = u * S;
                int a = AC_OUT - 1;
                float c = 0.f;
                bool found = false;
#pragma unroll
                for (int k = 0; k < AC_OUT; ++k) {
                    c += e[k];
                    if (!found && target < c) { a = k; found = true; }
                }
                bool live = true;
                if (n_agents) live = (int)(row % MAXV) < n_agents[row / MAXV];
                actions[row] = (int8_t)(live ? a : 1);   // absent agents idle (their rows are ignored by the env)
                if (logp_sel) logp_sel[row] = l[a] - m - logS;
                if (logp_all) {
#pragma unroll
                    for (int k = 0; k < AC_OUT; ++k) logp_all[row * AC_OUT + k] = l[k] - m - logS;
                }
            }
        }
    }
}

// R_t = r_t + gamma * R_{t+1}, restarted after a terminal step: MAPPO._discount_reward (marl/mappo.py:364-370) for every
// (env, agent) column of a rollout at once.  rewards / out [T][n_cols], dones [T][n_cols / cols_per_env] (1 where step t
// ended the episode of that env), final_value [n_cols] (critic bootstrap; ignored where the last step was terminal).
__global__ void __launch_bounds__(256) discounted_returns_kernel(const float *__restrict__ rewards,
                                                                 const uint8_t *__restrict__ dones,
                                                                 const float *__restrict__ final_value, float gamma, int T,
                                                                 int64_t n_cols, int cols_per_env, float *__restrict__ out) {
    const int64_t n_env = n_cols / cols_per_env;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cols; c += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = c / cols_per_env;
        float run = final_value ? final_value[c] : 0.f;
        for (int t = T - 1; t >= 0; --t) {
            if (dones[(int64_t)t * n_env + e]) run = 0.f;
            run = run * gamma + rewards[(int64_t)t * n_cols + c];
            out[(int64_t)t * n_cols + c] = run;
        }
    }
}

int launch_discounted_returns(const float *rewards, const uint8_t *dones, const float *final_value, float gamma, int T,
                              int64_t n_cols, int cols_per_env, float *out, void *stream) {
    if (n_cols <= 0 || T <= 0) return 0;
    int64_t blocks = (n_cols + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    discounted_returns_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rewards, dones, final_value, gamma, T, n_cols,
                                                                               cols_per_env, out);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_actor_sample(const float *obs, const int32_t *n_agents, int64_t n_rows, const float *w1, const float *b1,
                        const float *w2, const float *b2, const float *w3, const float *b3, uint64_t seed, uint64_t step,
                        int8_t *actions, float *logp_all, float *logp_sel, void *stream) {
    if (n_rows <= 0) return 0;
    static bool attr_set = false;
    const int smem = AC_SMEM_WORDS * (int)sizeof(uint32_t);
    if (!attr_set) {
        if (cudaFuncSetAttribute(actor_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
        attr_set = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t tiles = (n_rows + 15) / 16, ctas = (tiles + AC_THREADS / 32 - 1) / (AC_THREADS / 32);
    if (ctas > sms) ctas = sms;   // persistent: one CTA per SM, warps stride over the 16-row tiles
    actor_sample_kernel<<<(unsigned)ctas, AC_THREADS, smem, (cudaStream_t)stream>>>(obs, n_agents, n_rows, w1, b1, w2, b2, w3,
                                                                                     b3, seed, step, actions, logp_all, logp_sel);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace mm
