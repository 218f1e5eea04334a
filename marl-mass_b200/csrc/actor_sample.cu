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
// Two implementations, same arithmetic (TF32 operands, fp32 accumulation):
//   * actor_sample_tcgen05_kernel (default): the two 128-wide layers as tcgen05.mma kind::tf32 with the accumulators in
//     TMEM, 128 rows per tile, epilogues in the thread that owns the row - see the section further down;
//   * actor_sample_kernel (mm_set_actor_impl(1)): mma.sync.m16n8k8 per warp, kept as the independent cross-check.  The
//     accumulator fragment of layer L (thread holds rows g, g+8 and columns 2t, 2t+1 of each 8-wide tile) is reused
//     directly as the A fragment of layer L+1 (rows g, g+8, columns t, t+4) by permuting the K index of that layer:
//     A column t <-> hidden unit 8kk+2t, A column t+4 <-> 8kk+2t+1, and the weights are laid out in shared memory in
//     fragment order with the same permutation, so no shuffle or shared-memory round trip separates the layers.
// Measured at 786 432 rows: torch layers 1.89 ms, mma.sync 0.20 ms, tcgen05 0.17 ms (profiles/README.md); this op is
// ~33 GFLOP per step, so both are bound by their epilogues and staging, not by the tensor pipe.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>
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

// Invalid-action masking of the MAPPO_GI actor (marl/single_agent/Model_gi.py:63-66: logits[action_mask == 0] = -1e8
// before the log-softmax): bit k of mask_bits[row] set = action k available (the env kernel's action_mask buffer).
__device__ __forceinline__ void apply_action_mask(float (&l)[AC_OUT], const uint8_t *mask_bits, int64_t row) {
    if (!mask_bits) return;
    const uint32_t bits = mask_bits[row];
#pragma unroll
    for (int k = 0; k < AC_OUT; ++k)
        if (!((bits >> k) & 1u)) l[k] = -1e8f;
}

__global__ void __launch_bounds__(AC_THREADS, 1)
actor_sample_kernel(const float *__restrict__ obs, const int32_t *__restrict__ n_agents, int64_t n_rows,
                    const float *__restrict__ w1, const float *__restrict__ b1, const float *__restrict__ w2,
                    const float *__restrict__ b2, const float *__restrict__ w3, const float *__restrict__ b3,
                    uint64_t seed, uint64_t step, const uint8_t *__restrict__ mask_bits, int8_t *__restrict__ actions,
                    float *__restrict__ logp_all, float *__restrict__ logp_sel) {
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
                apply_action_mask(l, mask_bits, row);
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

// ------------------------------------------------------------------------------------------------
// tcgen05 version: the two 128-wide layers on the 5th-generation tensor cores
// ------------------------------------------------------------------------------------------------
// One persistent CTA per SM, 512 threads, 128 observation rows per tile (= the 128 TMEM lanes).
//   layer 1   D1[128x128] = A1[128x32]  * W1^T   4 x tcgen05.mma kind::tf32 (M128 N128 K8), accumulator in TMEM cols 0..127
//   epilogue  TMEM -> registers (tcgen05.ld 32x32b), + bias, ReLU, TF32 rounding -> A2 in shared memory
//   layer 2   D2[128x128] = A2[128x128] * W2^T  16 x tcgen05.mma, accumulator in TMEM cols 128..255
//   epilogue  TMEM -> registers, + bias, ReLU, and the 128 -> 5 output layer as FMAs in the thread that owns the row
//             (too narrow for an MMA tile), log-softmax, inverse-CDF draw, one action byte per row
// Operands are K-major, no swizzle: 16-byte rows of 4 TF32 values, core matrices of 8 rows; with the layout
// [k-chunk][row] (float4) a warp's 32 rows of one chunk are 512 contiguous bytes (conflict-free staging), the stride
// between 8-row groups is SBO = 128 B and between the two 16-byte k-chunks of one MMA LBO = 2048 B.
// Warps w, w+4, w+8, w+12 share TMEM lane quarter w%4 (hardware rule) and split the 128 columns four ways; the next
// tile's observations are requested right after a tile's first MMA is issued, so their latency hides behind the epilogues.
constexpr int T5_THREADS = 512;
constexpr int T5_CSPLIT = T5_THREADS / 128;                  // column groups (of 128 / T5_CSPLIT columns) per row
constexpr int T5_COLS = 128 / T5_CSPLIT;
constexpr int T5_ROWS = 128;
constexpr uint32_t T5_CHUNK_BYTES = T5_ROWS * 16;            // one k-chunk (4 values) of all 128 rows
// shared memory (bytes)
constexpr uint32_t T5_A1 = 0;                                // [8 chunks][128] float4
constexpr uint32_t T5_W1 = T5_A1 + 8 * T5_CHUNK_BYTES;       // [8 chunks][128] float4
constexpr uint32_t T5_A2 = T5_W1 + 8 * T5_CHUNK_BYTES;       // [32 chunks][128] float4
constexpr uint32_t T5_W2 = T5_A2 + 32 * T5_CHUNK_BYTES;      // [32 chunks][128] float4
constexpr uint32_t T5_W3 = T5_W2 + 32 * T5_CHUNK_BYTES;      // [128 hidden][8] f32 (5 used)
constexpr uint32_t T5_B1 = T5_W3 + AC_HID * 8 * 4;
constexpr uint32_t T5_B2 = T5_B1 + AC_HID * 4;
constexpr uint32_t T5_B3 = T5_B2 + AC_HID * 4;
constexpr uint32_t T5_PART = T5_B3 + 8 * 4;                  // [column groups 1..][128 rows][8] f32 partial logits
constexpr uint32_t T5_BAR = T5_PART + (T5_CSPLIT - 1) * T5_ROWS * 8 * 4;   // 2 mbarriers + the TMEM base address
constexpr uint32_t T5_SMEM = T5_BAR + 32;

__device__ __forceinline__ uint32_t t5_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t t5_desc(uint32_t smem_addr) {
    // K-major, SWIZZLE_NONE: start address, LBO (between the two k-chunks of an MMA), SBO (between 8-row groups), version 1
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(T5_CHUNK_BYTES >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) |
           (1ull << 46);
}
// kind::tf32, F32 accumulate, M = 128, N = 128, A and B K-major (cute/arch/mma_sm100_desc.hpp InstrDescriptor)
constexpr uint32_t T5_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void t5_mma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(T5_IDESC), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void t5_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void t5_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void t5_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
                 "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(T5_THREADS, 1)
actor_sample_tcgen05_kernel(const float *__restrict__ obs, const int32_t *__restrict__ n_agents, int64_t n_rows,
                            const float *__restrict__ w1, const float *__restrict__ b1, const float *__restrict__ w2,
                            const float *__restrict__ b2, const float *__restrict__ w3, const float *__restrict__ b3,
                            uint64_t seed, uint64_t step, const uint8_t *__restrict__ mask_bits,
                            int8_t *__restrict__ actions, float *__restrict__ logp_all, float *__restrict__ logp_sel) {
    extern __shared__ __align__(128) uint8_t t5_sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float4 *a1 = reinterpret_cast<float4 *>(t5_sm + T5_A1), *w1f = reinterpret_cast<float4 *>(t5_sm + T5_W1);
    float4 *a2 = reinterpret_cast<float4 *>(t5_sm + T5_A2), *w2f = reinterpret_cast<float4 *>(t5_sm + T5_W2);
    float *w3s = reinterpret_cast<float *>(t5_sm + T5_W3), *b1s = reinterpret_cast<float *>(t5_sm + T5_B1);
    float *b2s = reinterpret_cast<float *>(t5_sm + T5_B2), *b3s = reinterpret_cast<float *>(t5_sm + T5_B3);
    float *part = reinterpret_cast<float *>(t5_sm + T5_PART);
    uint64_t *bars = reinterpret_cast<uint64_t *>(t5_sm + T5_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t5_sm + T5_BAR + 16);
    const uint32_t bar1 = t5_smem(bars), bar2 = t5_smem(bars + 1);

    // ---- one-time setup: weights in operand layout (TF32-rounded), barriers, TMEM ----
    auto tf = [](float x) { return __uint_as_float(to_tf32(x)); };
    for (int idx = tid; idx < 8 * T5_ROWS; idx += T5_THREADS) {          // W1: [out n][in k], k padded 30 -> 32
        const int c = idx / T5_ROWS, n = idx % T5_ROWS, k = 4 * c;
        float4 v;
        v.x = tf(k + 0 < AC_IN ? w1[n * AC_IN + k + 0] : 0.f); v.y = tf(k + 1 < AC_IN ? w1[n * AC_IN + k + 1] : 0.f);
        v.z = tf(k + 2 < AC_IN ? w1[n * AC_IN + k + 2] : 0.f); v.w = tf(k + 3 < AC_IN ? w1[n * AC_IN + k + 3] : 0.f);
        w1f[idx] = v;
    }
    for (int idx = tid; idx < 32 * T5_ROWS; idx += T5_THREADS) {
        const int c = idx / T5_ROWS, n = idx % T5_ROWS;
        const float4 v = *reinterpret_cast<const float4 *>(w2 + n * AC_HID + 4 * c);
        w2f[idx] = make_float4(tf(v.x), tf(v.y), tf(v.z), tf(v.w));
    }
    for (int idx = tid; idx < AC_HID * 8; idx += T5_THREADS) {
        const int h = idx >> 3, k = idx & 7;
        w3s[idx] = k < AC_OUT ? w3[k * AC_HID + h] : 0.f;
    }
    for (int idx = tid; idx < AC_HID; idx += T5_THREADS) { b1s[idx] = b1[idx]; b2s[idx] = b2[idx]; }
    if (tid < 8) b3s[tid] = tid < AC_OUT ? b3[tid] : 0.f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar2));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(t5_smem(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t a1_addr = t5_smem(a1), w1_addr = t5_smem(w1f), a2_addr = t5_smem(a2), w2_addr = t5_smem(w2f);

    const int row_in_tile = 32 * (warp & 3) + lane;        // TMEM lane = row of the tile this thread reads
    const int cgrp = warp >> 2;                            // which T5_COLS of the 128 columns
    const int col0 = T5_COLS * cgrp;
    static_assert(T5_COLS == 32, "one tcgen05.ld.32x32b.x32 per thread and epilogue");
    const uint32_t t_lane = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    const int64_t n_tiles = (n_rows + T5_ROWS - 1) / T5_ROWS;
    // observation staging: thread t handles row t % 128, k-chunks 2 * (t / 128) and + 1
    const int st_r = tid & (T5_ROWS - 1), st_c0 = 2 * (tid >> 7);
    float4 pre[2];
    auto fetch = [&](int64_t tile) {
        const int64_t row = tile * T5_ROWS + st_r;
        const float2 *src = reinterpret_cast<const float2 *>(obs + row * AC_IN);   // rows are 120 B: 8-byte aligned
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int c = st_c0 + q;
            float2 lo = make_float2(0.f, 0.f), hi = lo;
            if (tile < n_tiles && row < n_rows) {
                lo = __ldg(src + 2 * c);
                if (4 * c + 2 < AC_IN) hi = __ldg(src + 2 * c + 1);
            }
            pre[q] = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
    };
    fetch(blockIdx.x);
    uint32_t parity = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, parity ^= 1) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
            a1[(st_c0 + q) * T5_ROWS + st_r] = make_float4(tf(pre[q].x), tf(pre[q].y), tf(pre[q].z), tf(pre[q].w));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 4; ++j)
                t5_mma(tmem, t5_desc(a1_addr + j * 2 * T5_CHUNK_BYTES), t5_desc(w1_addr + j * 2 * T5_CHUNK_BYTES), j > 0);
            t5_commit(bar1);
        }
        fetch(tile + gridDim.x);                           // next tile's rows: in flight during this tile's epilogues
        t5_wait(bar1, parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 1: h1 = relu(D1 + b1) -> A2 (this thread: its row, 32 columns) ----
        {
            uint32_t v[32];
            t5_ld32(t_lane + (uint32_t)col0, v);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 o;
                o.x = tf(fmaxf(__uint_as_float(v[4 * q + 0]) + b1s[col0 + 4 * q + 0], 0.f));
                o.y = tf(fmaxf(__uint_as_float(v[4 * q + 1]) + b1s[col0 + 4 * q + 1], 0.f));
                o.z = tf(fmaxf(__uint_as_float(v[4 * q + 2]) + b1s[col0 + 4 * q + 2], 0.f));
                o.w = tf(fmaxf(__uint_as_float(v[4 * q + 3]) + b1s[col0 + 4 * q + 3], 0.f));
                a2[(col0 / 4 + q) * T5_ROWS + row_in_tile] = o;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j)
                t5_mma(tmem + 128, t5_desc(a2_addr + j * 2 * T5_CHUNK_BYTES), t5_desc(w2_addr + j * 2 * T5_CHUNK_BYTES), j > 0);
            t5_commit(bar2);
        }
        t5_wait(bar2, parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 2: h2 = relu(D2 + b2), output layer as FMAs over this thread's 32 hidden units ----
        float lg[AC_OUT] = {0.f, 0.f, 0.f, 0.f, 0.f};
        {
            uint32_t v[32];
            t5_ld32(t_lane + 128u + (uint32_t)col0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float h = fmaxf(__uint_as_float(v[j]) + b2s[col0 + j], 0.f);
                const float4 wa = *reinterpret_cast<const float4 *>(w3s + (col0 + j) * 8);
                const float wb = w3s[(col0 + j) * 8 + 4];
                lg[0] = fmaf(h, wa.x, lg[0]); lg[1] = fmaf(h, wa.y, lg[1]); lg[2] = fmaf(h, wa.z, lg[2]);
                lg[3] = fmaf(h, wa.w, lg[3]); lg[4] = fmaf(h, wb, lg[4]);
            }
        }
        if (cgrp > 0) {
#pragma unroll
            for (int k = 0; k < AC_OUT; ++k) part[((cgrp - 1) * T5_ROWS + row_in_tile) * 8 + k] = lg[k];
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        const int64_t row = tile * T5_ROWS + row_in_tile;
        if (cgrp == 0 && row < n_rows) {
            float l[AC_OUT];
#pragma unroll
            for (int k = 0; k < AC_OUT; ++k) {
                float acc = lg[k];
#pragma unroll
                for (int g2 = 0; g2 < T5_CSPLIT - 1; ++g2) acc += part[(g2 * T5_ROWS + row_in_tile) * 8 + k];
                l[k] = acc + b3s[k];
            }
            apply_action_mask(l, mask_bits, row);
            float m = l[0];
#pragma unroll
            for (int k = 1; k < AC_OUT; ++k) m = fmaxf(m, l[k]);
            float e[AC_OUT], S = 0.f;
#pragma unroll
            for (int k = 0; k < AC_OUT; ++k) { e[k] = __expf(l[k] - m); S += e[k]; }
            const float logS = __logf(S);
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
            actions[row] = (int8_t)(live ? a : 1);
            if (logp_sel) logp_sel[row] = l[a] - m - logS;
            if (logp_all) {
#pragma unroll
                for (int k = 0; k < AC_OUT; ++k) logp_all[row * AC_OUT + k] = l[k] - m - logS;
            }
        }
        __syncthreads();   // `part` is reused by the next tile
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05, second generation: fp16 operands, two CTAs per SM, any first-layer width, optional value head
// ------------------------------------------------------------------------------------------------
// The TF32 kernel above runs one tile at a time per SM: stage -> MMA -> epilogue -> MMA -> epilogue, each phase waiting
// for the one before (ncu: long-scoreboard stalls on tcgen05.ld / mbarrier polls, tensor pipe 19 % busy).  This one
// halves every shared-memory operand (kind::f16: fp16 has TF32's 10-bit mantissa; observations live in [-1, 1], the
// hidden activations of these networks stay far below 65 504) so that TWO CTAs fit an SM (<= 108 KB each, 256 TMEM
// columns each): while one CTA's threads run an epilogue the other CTA's MMAs and loads proceed - the overlap of a
// software pipeline without warp specialisation.  MMAs are K = 16 (half as many instructions).  The 128 -> 5 output
// layer (+ the value column) is a third MMA (N = 16) on h2 staged as fp16: as per-thread FMAs over shared-memory weights
// it was ~55 % of the kernel's instructions (ncu: 55 M warp-instructions per launch, issue slots 46 % busy).
// 786 432 rows: TF32 kernel 0.167 ms -> 0.115 (fp16, two CTAs / SM) -> 0.093 (third MMA).
//   H1 = 128: MAPPO ActorNetwork 30-128-128-5 (Model_common.py:5-23).
//   H1 = 160: MAPPO_GI ActorCriticNetwork with state_split (Model_gi.py:137-220): its three first-layer blocks
//             (5 -> 32, 10 -> 64, 10 -> 64 over fixed column lists) are ONE 30 -> 160 layer whose weight matrix is zero
//             outside the blocks; the host scatters fc11 / fc12 / fc13 into it (rollout.py).  The optional sixth output
//             column is critic_linear (the state value V(s), mappo_gi.py:396-404).
// CTA = 256 threads: warp w reads TMEM lane quarter w % 4 (hardware rule) and column half w / 4.  D2 reuses D1's
// columns (D1 is dead once epilogue 1 has been read and the CTA has synchronised), h2 reuses h1's shared memory, D3 sits
// in columns 192..207; the final epilogue (log-softmax, draw) runs in the 128 threads of column half 0, one row each.
template <int H1>
struct M5 {
    static constexpr int THREADS = 256;
    static constexpr int ROWS = 128;
    static constexpr int CH1 = H1 / 8;                        // k-chunks (8 halves = 16 bytes) of layer 2
    static constexpr uint32_t A_CHUNK = ROWS * 16;            // one k-chunk of the 128 rows of a tile / of W2's 128 rows
    static constexpr uint32_t W1_CHUNK = H1 * 16;             // one k-chunk of W1's H1 rows
    static constexpr uint32_t W3_CHUNK = 16 * 16;             // one k-chunk of the output layer's 16 rows (5 logits, the value, 10 zero)
    static constexpr uint32_t A1 = 0;                         // [4 chunks][128] 16 B
    static constexpr uint32_t W1 = A1 + 4 * A_CHUNK;          // [4 chunks][H1]
    static constexpr uint32_t A2 = W1 + 4 * W1_CHUNK;         // [CH1][128]: h1, then (first 16 chunks) h2
    static constexpr uint32_t W2 = A2 + CH1 * A_CHUNK;        // [CH1][128]
    static constexpr uint32_t W3 = W2 + CH1 * A_CHUNK;        // [16 chunks][16]
    static constexpr uint32_t B1 = W3 + 16 * W3_CHUNK;
    static constexpr uint32_t B2 = B1 + H1 * 4;
    static constexpr uint32_t B3 = B2 + AC_HID * 4;
    static constexpr uint32_t BAR = B3 + 8 * 4;               // 3 mbarriers + the TMEM base address
    static constexpr uint32_t SMEM = BAR + 32;
    static constexpr uint32_t D3_COL = 192;                   // TMEM columns: D1 0..H1-1, D2 0..127 (reuses D1's), D3 192..207
    // instruction descriptors, kind::f16: F32 accumulate, fp16 A / B, K-major, M = 128, N = H1 / 128 / 16
    static constexpr uint32_t IDESC1 = (1u << 4) | ((uint32_t)(H1 >> 3) << 17) | ((128u >> 4) << 24);
    static constexpr uint32_t IDESC2 = (1u << 4) | ((uint32_t)(AC_HID >> 3) << 17) | ((128u >> 4) << 24);
    static constexpr uint32_t IDESC3 = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
};

__device__ __forceinline__ uint64_t m5_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    // K-major, SWIZZLE_NONE: start address, LBO (between the two 16-byte k-chunks of an MMA), SBO (between 8-row groups)
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void m5_mma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void m5_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));    // low half = a, high half = b
    return r;
}

template <int H1>
__global__ void __launch_bounds__(256, 2)
actor_mlp_tcgen05_kernel(const float *__restrict__ obs, const int32_t *__restrict__ n_agents, int64_t n_rows,
                         const float *__restrict__ w1, const float *__restrict__ b1, const float *__restrict__ w2,
                         const float *__restrict__ b2, const float *__restrict__ w3, const float *__restrict__ b3,
                         const float *__restrict__ wv, const float *__restrict__ bv, uint64_t seed, uint64_t step,
                         const uint8_t *__restrict__ mask_bits, int8_t *__restrict__ actions, float *__restrict__ logp_all,
                         float *__restrict__ logp_sel, float *__restrict__ values, float *__restrict__ obs_copy,
                         uint8_t *__restrict__ live_out) {
    using L = M5<H1>;
    extern __shared__ __align__(128) uint8_t m5_sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint4 *a1 = reinterpret_cast<uint4 *>(m5_sm + L::A1), *w1f = reinterpret_cast<uint4 *>(m5_sm + L::W1);
    uint4 *a2 = reinterpret_cast<uint4 *>(m5_sm + L::A2), *w2f = reinterpret_cast<uint4 *>(m5_sm + L::W2);
    uint4 *w3f = reinterpret_cast<uint4 *>(m5_sm + L::W3);
    float *b1s = reinterpret_cast<float *>(m5_sm + L::B1), *b2s = reinterpret_cast<float *>(m5_sm + L::B2);
    float *b3s = reinterpret_cast<float *>(m5_sm + L::B3);
    uint64_t *bars = reinterpret_cast<uint64_t *>(m5_sm + L::BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(m5_sm + L::BAR + 24);
    const uint32_t bar1 = t5_smem(bars), bar2 = t5_smem(bars + 1), bar3 = t5_smem(bars + 2);

    // ---- one-time setup: weights in operand layout (fp16), barriers, TMEM ----
    for (int idx = tid; idx < 4 * H1; idx += L::THREADS) {           // W1: [out n][in k], k padded 30 -> 32
        const int c = idx / H1, n = idx % H1, k = 8 * c;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = k + q < AC_IN ? w1[n * AC_IN + k + q] : 0.f;
        w1f[idx] = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
    }
    for (int idx = tid; idx < L::CH1 * L::ROWS; idx += L::THREADS) {  // W2: [out n = 128][in k = H1]
        const int c = idx / L::ROWS, n = idx % L::ROWS;
        const float4 lo = *reinterpret_cast<const float4 *>(w2 + n * H1 + 8 * c), hi = *reinterpret_cast<const float4 *>(w2 + n * H1 + 8 * c + 4);
        w2f[idx] = make_uint4(pack_h2(lo.x, lo.y), pack_h2(lo.z, lo.w), pack_h2(hi.x, hi.y), pack_h2(hi.z, hi.w));
    }
    for (int idx = tid; idx < 16 * 16; idx += L::THREADS) {          // output layer: rows 0..4 logits, 5 the value head, rest zero
        const int c = idx / 16, n = idx % 16;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = n < AC_OUT ? w3[n * AC_HID + 8 * c + q] : (n == AC_OUT && wv ? wv[8 * c + q] : 0.f);
        w3f[idx] = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
    }
    for (int idx = tid; idx < H1; idx += L::THREADS) b1s[idx] = b1[idx];
    for (int idx = tid; idx < AC_HID; idx += L::THREADS) b2s[idx] = b2[idx];
    if (tid < 8) b3s[tid] = tid < AC_OUT ? b3[tid] : (tid == AC_OUT && bv ? bv[0] : 0.f);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar2));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar3));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(t5_smem(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t a1_addr = t5_smem(a1), w1_addr = t5_smem(w1f), a2_addr = t5_smem(a2), w2_addr = t5_smem(w2f), w3_addr = t5_smem(w3f);

    const int row_in_tile = 32 * (warp & 3) + lane;        // TMEM lane = row of the tile this thread reads
    const int hh = warp >> 2;                              // column half
    const uint32_t t_lane = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    const int64_t n_tiles = (n_rows + L::ROWS - 1) / L::ROWS;
    // observation staging: thread t handles row t % 128, k-chunks 2 * (t / 128) and + 1 (8 floats each)
    const int st_r = tid & (L::ROWS - 1), st_c0 = 2 * (tid >> 7);
    float2 pre[8];
    auto fetch = [&](int64_t tile) {
        const int64_t row = tile * L::ROWS + st_r;
        const float2 *src = reinterpret_cast<const float2 *>(obs + row * AC_IN);   // rows are 120 B: 8-byte aligned
        const bool ok = tile < n_tiles && row < n_rows;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int f2 = 4 * st_c0 + q;                  // float2 index inside the row: 15 hold data
            pre[q] = (ok && f2 < AC_IN / 2) ? __ldg(src + f2) : make_float2(0.f, 0.f);
            // rollout buffer (MAPPO.interact appends the state it acted on, mappo.py:117-131): the rows pass through
            // this thread anyway
            if (obs_copy && ok && f2 < AC_IN / 2) __stcs(reinterpret_cast<float2 *>(obs_copy + row * AC_IN) + f2, pre[q]);
        }
    };
    fetch(blockIdx.x);
    uint32_t parity = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, parity ^= 1) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
            a1[(st_c0 + q) * L::ROWS + st_r] = make_uint4(pack_h2(pre[4 * q].x, pre[4 * q].y), pack_h2(pre[4 * q + 1].x, pre[4 * q + 1].y),
                                                         pack_h2(pre[4 * q + 2].x, pre[4 * q + 2].y), pack_h2(pre[4 * q + 3].x, pre[4 * q + 3].y));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 2; ++j)
                m5_mma(tmem, m5_desc(a1_addr + j * 2 * L::A_CHUNK, L::A_CHUNK), m5_desc(w1_addr + j * 2 * L::W1_CHUNK, L::W1_CHUNK),
                       L::IDESC1, j > 0);
            t5_commit(bar1);
        }
        fetch(tile + gridDim.x);                           // next tile's rows: in flight during this tile's epilogues
        t5_wait(bar1, parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 1: h1 = relu(D1 + b1) -> A2 (this thread: its row, H1 / 2 columns in pieces of 16) ----
#pragma unroll 1
        for (int it = 0; it < H1 / 32; ++it) {
            const int col = hh * (H1 / 2) + 16 * it;
            uint32_t v[16];
            m5_ld16(t_lane + (uint32_t)col, v);
            float h[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) h[q] = fmaxf(__uint_as_float(v[q]) + b1s[col + q], 0.f);
            a2[(col / 8) * L::ROWS + row_in_tile] = make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
            a2[(col / 8 + 1) * L::ROWS + row_in_tile] = make_uint4(pack_h2(h[8], h[9]), pack_h2(h[10], h[11]), pack_h2(h[12], h[13]), pack_h2(h[14], h[15]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < H1 / 16; ++j)
                m5_mma(tmem, m5_desc(a2_addr + j * 2 * L::A_CHUNK, L::A_CHUNK), m5_desc(w2_addr + j * 2 * L::A_CHUNK, L::A_CHUNK),
                       L::IDESC2, j > 0);
            t5_commit(bar2);
        }
        t5_wait(bar2, parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 2: h2 = relu(D2 + b2) -> the first 16 chunks of A2 (MMA2 has finished reading h1) ----
#pragma unroll 1
        for (int it = 0; it < 4; ++it) {
            const int col = hh * 64 + 16 * it;
            uint32_t v[16];
            m5_ld16(t_lane + (uint32_t)col, v);
            float h[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) h[q] = fmaxf(__uint_as_float(v[q]) + b2s[col + q], 0.f);
            a2[(col / 8) * L::ROWS + row_in_tile] = make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
            a2[(col / 8 + 1) * L::ROWS + row_in_tile] = make_uint4(pack_h2(h[8], h[9]), pack_h2(h[10], h[11]), pack_h2(h[12], h[13]), pack_h2(h[14], h[15]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // ---- output layer 128 -> 5 (+ value) as a third MMA (N = 16): D3 = h2 * W3^T ----
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < AC_HID / 16; ++j)
                m5_mma(tmem + L::D3_COL, m5_desc(a2_addr + j * 2 * L::A_CHUNK, L::A_CHUNK), m5_desc(w3_addr + j * 2 * L::W3_CHUNK, L::W3_CHUNK),
                       L::IDESC3, j > 0);
            t5_commit(bar3);
        }
        t5_wait(bar3, parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t row = tile * L::ROWS + row_in_tile;
        if (hh == 0) {
            uint32_t v[16];
            m5_ld16(t_lane + L::D3_COL, v);
            if (row < n_rows) {
                float l[AC_OUT];
#pragma unroll
                for (int k = 0; k < AC_OUT; ++k) l[k] = __uint_as_float(v[k]) + b3s[k];
                if (values) values[row] = __uint_as_float(v[AC_OUT]) + b3s[AC_OUT];
                apply_action_mask(l, mask_bits, row);
                float m = l[0];
#pragma unroll
                for (int k = 1; k < AC_OUT; ++k) m = fmaxf(m, l[k]);
                float e[AC_OUT], S = 0.f;
#pragma unroll
                for (int k = 0; k < AC_OUT; ++k) { e[k] = __expf(l[k] - m); S += e[k]; }
                const float logS = __logf(S);
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
                actions[row] = (int8_t)(live ? a : 1);
                if (live_out) live_out[row] = live ? 1 : 0;
                if (logp_sel) logp_sel[row] = l[a] - m - logS;
                if (logp_all) {
#pragma unroll
                    for (int k = 0; k < AC_OUT; ++k) logp_all[row * AC_OUT + k] = l[k] - m - logS;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // TMEM columns and A2 are reused by the next tile
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05, third generation: one warpgroup per tile, four (H1 = 128) or three (H1 = 160) tiles in flight per SM
// ------------------------------------------------------------------------------------------------
// ncu on the two-CTA kernel above (profiles/r2_w_ncu_summary.txt): issue slots 33 % busy, 47 % of the stall samples on
// the long scoreboard (mbarrier polls after each of the three MMAs, and the rollout-buffer store of a row right behind
// its load, which made the prefetch synchronous), 1 400 thread-instructions per row of which the bias add + relu +
// convert of the 256 hidden activations were two thirds.  Changes:
//   * the weights sit in shared memory ONCE per SM and every warpgroup (128 threads = the 128 TMEM lanes = the 128 rows of
//     a tile) runs its own tile pipeline on its own operand buffers, TMEM columns, named barrier and mbarrier: more
//     tiles in flight per SM, no CTA-wide barrier;
//   * biases ride in the MMAs: an extra K-step multiplies a constant block (1, 1, 0, ...) by (hi, lo) fp16 halves of the
//     bias (layer 1 uses its two padding columns k = 30, 31), so the accumulator already holds D + b to ~2^-22;
//   * the epilogue of a hidden layer is cvt.rn.relu.f16x2.f32 - one instruction per two activations - and a 16-byte
//     shared store per eight;
//   * D3 reuses the first 16 columns of the tile's TMEM region (D2 is dead once epilogue 2 has synchronised).
template <int H1>
struct M6 {
    static constexpr int NWG = H1 == 128 ? 4 : 3;            // tiles in flight: shared memory (H1 = 160) or TMEM (128) bound
    static constexpr int THREADS = 128 * NWG;
    static constexpr int ROWS = 128;
    static constexpr int CH1 = H1 / 8;
    static constexpr uint32_t A_CHUNK = ROWS * 16;
    static constexpr uint32_t W1_CHUNK = H1 * 16;
    static constexpr uint32_t W3_CHUNK = 16 * 16;
    static constexpr uint32_t W1 = 0;                             // [4 chunks][H1]: k = 30, 31 carry b1 (hi, lo)
    static constexpr uint32_t W2 = W1 + 4 * W1_CHUNK;             // [CH1 + 2][128]: the last two chunks carry b2
    static constexpr uint32_t W3 = W2 + (CH1 + 2) * A_CHUNK;      // [16 + 2][16]: the last two carry b3 and the value bias
    static constexpr uint32_t ONES = W3 + 18 * W3_CHUNK;          // [2 chunks][128]: (1, 1, 0, ..., 0) per row
    static constexpr uint32_t TILE0 = ONES + 2 * A_CHUNK;         // per warpgroup: A1 [4][128], A2 [CH1][128]
    static constexpr uint32_t TILE_BYTES = (4 + CH1) * A_CHUNK;
    static constexpr uint32_t BAR = TILE0 + NWG * TILE_BYTES;     // NWG mbarriers + the TMEM base address
    static constexpr uint32_t SMEM = BAR + 8 * NWG + 16;
    static constexpr uint32_t TMEM_COLS = 512;
    static constexpr uint32_t REGION = H1;                        // TMEM columns per warpgroup
    static constexpr uint32_t IDESC1 = (1u << 4) | ((uint32_t)(H1 >> 3) << 17) | ((128u >> 4) << 24);
    static constexpr uint32_t IDESC2 = (1u << 4) | ((uint32_t)(AC_HID >> 3) << 17) | ((128u >> 4) << 24);
    static constexpr uint32_t IDESC3 = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
};
static_assert(M6<128>::SMEM <= 227 * 1024 && M6<160>::SMEM <= 227 * 1024, "actor kernel shared memory");
static_assert(M6<128>::NWG * M6<128>::REGION <= 512 && M6<160>::NWG * M6<160>::REGION <= 512, "actor kernel TMEM columns");

__device__ __forceinline__ uint32_t pack_relu_h2(uint32_t a_bits, uint32_t b_bits) {
    uint32_t r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(b_bits)), "f"(__uint_as_float(a_bits)));
    return r;
}
__device__ __forceinline__ uint32_t hi_lo_h2(float b) {          // fp16 halves whose sum is b to ~2^-22
    const __half hi = __float2half_rn(b), lo = __float2half_rn(b - __half2float(hi));
    return (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
}
__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

template <int H1>
__global__ void __launch_bounds__(M6<H1>::THREADS, 1)
actor_mlp_wg_kernel(const float *__restrict__ obs, const int32_t *__restrict__ n_agents, int64_t n_rows,
                    const float *__restrict__ w1, const float *__restrict__ b1, const float *__restrict__ w2,
                    const float *__restrict__ b2, const float *__restrict__ w3, const float *__restrict__ b3,
                    const float *__restrict__ wv, const float *__restrict__ bv, uint64_t seed, uint64_t step,
                    const uint8_t *__restrict__ mask_bits, int8_t *__restrict__ actions, float *__restrict__ logp_all,
                    float *__restrict__ logp_sel, float *__restrict__ values, float *__restrict__ obs_copy,
                    uint8_t *__restrict__ live_out) {
    using L = M6<H1>;
    extern __shared__ __align__(128) uint8_t m6_sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wg = warp >> 2;
    uint4 *w1f = reinterpret_cast<uint4 *>(m6_sm + L::W1), *w2f = reinterpret_cast<uint4 *>(m6_sm + L::W2);
    uint4 *w3f = reinterpret_cast<uint4 *>(m6_sm + L::W3), *ones = reinterpret_cast<uint4 *>(m6_sm + L::ONES);
    uint4 *a1 = reinterpret_cast<uint4 *>(m6_sm + L::TILE0 + wg * L::TILE_BYTES), *a2 = a1 + 4 * L::ROWS;
    uint64_t *bars = reinterpret_cast<uint64_t *>(m6_sm + L::BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(m6_sm + L::BAR + 8 * L::NWG);
    const uint32_t bar = t5_smem(bars + wg);

    // ---- one-time setup: weights and biases in operand layout (fp16), the constant block, barriers, TMEM ----
    for (int idx = tid; idx < 4 * H1; idx += L::THREADS) {
        const int c = idx / H1, n = idx % H1, k = 8 * c;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = k + q < AC_IN ? w1[n * AC_IN + k + q] : 0.f;
        uint4 o = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        if (c == 3) o.w = hi_lo_h2(b1[n]);                            // k = 30, 31
        w1f[idx] = o;
    }
    for (int idx = tid; idx < (L::CH1 + 2) * L::ROWS; idx += L::THREADS) {
        const int c = idx / L::ROWS, n = idx % L::ROWS;
        if (c < L::CH1) {
            const float4 lo = *reinterpret_cast<const float4 *>(w2 + n * H1 + 8 * c), hi = *reinterpret_cast<const float4 *>(w2 + n * H1 + 8 * c + 4);
            w2f[idx] = make_uint4(pack_h2(lo.x, lo.y), pack_h2(lo.z, lo.w), pack_h2(hi.x, hi.y), pack_h2(hi.z, hi.w));
        } else {
            w2f[idx] = make_uint4(c == L::CH1 ? hi_lo_h2(b2[n]) : 0u, 0u, 0u, 0u);
        }
    }
    for (int idx = tid; idx < 18 * 16; idx += L::THREADS) {          // rows 0..4 logits, 5 the value head, rest zero
        const int c = idx / 16, n = idx % 16;
        if (c < 16) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = n < AC_OUT ? w3[n * AC_HID + 8 * c + q] : (n == AC_OUT && wv ? wv[8 * c + q] : 0.f);
            w3f[idx] = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        } else {
            const float b = n < AC_OUT ? b3[n] : (n == AC_OUT && bv ? bv[0] : 0.f);
            w3f[idx] = make_uint4(c == 16 ? hi_lo_h2(b) : 0u, 0u, 0u, 0u);
        }
    }
    for (int idx = tid; idx < 2 * L::ROWS; idx += L::THREADS) ones[idx] = make_uint4(idx < L::ROWS ? 0x3C003C00u : 0u, 0u, 0u, 0u);
    if (tid < L::NWG) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(t5_smem(bars + tid)));
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(t5_smem(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot + (uint32_t)wg * L::REGION;          // this warpgroup's columns
    const uint32_t a1_addr = t5_smem(a1), a2_addr = t5_smem(a2), w1_addr = t5_smem(w1f), w2_addr = t5_smem(w2f);
    const uint32_t w3_addr = t5_smem(w3f), ones_addr = t5_smem(ones);

    const int r = tid & 127;                                              // row of the tile = TMEM lane
    const uint32_t t_lane = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    const int64_t n_tiles = (n_rows + L::ROWS - 1) / L::ROWS;
    const int64_t stride = (int64_t)gridDim.x * L::NWG;
    // observation rows: a tile is 128 x 30 floats = 960 contiguous float4 (row-per-thread loads touched 30 cache lines per
    // request: 5.3 x the minimal sector count, L1 request stage 55 % busy - profiles/r2_x_*); thread r takes float4
    // r, r + 128, ...: coalesced, and each float4 is two in-row pairs (30 and the flat index are even) -> two half2 stores
    constexpr int TILE_F4 = L::ROWS * AC_IN / 4, PRE = (TILE_F4 + 127) / 128;
    const int64_t total_f = n_rows * AC_IN;
    float4 pre[PRE];
    auto fetch = [&](int64_t tile) {
#pragma unroll
        for (int q = 0; q < PRE; ++q) {
            const int i = r + 128 * q;
            const int64_t g = tile * (L::ROWS * AC_IN) + 4 * i;                    // flat float index
            pre[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < TILE_F4 && tile < n_tiles) {
                if (g + 4 <= total_f) pre[q] = __ldg(reinterpret_cast<const float4 *>(obs + g));
                else if (g + 2 <= total_f) { const float2 h = __ldg(reinterpret_cast<const float2 *>(obs + g)); pre[q].x = h.x; pre[q].y = h.y; }
            }
        }
    };
    uint32_t *a1w = reinterpret_cast<uint32_t *>(a1);
    a1[3 * L::ROWS + r].w = 0x3C003C00u;                       // k = 30, 31: the constant 1 that multiplies b1 (never overwritten)
    auto a1_word = [&](int f) {                                // 32-bit word of A1 that holds elements (f, f + 1) of the flat tile
        const int row = f / AC_IN, col = f - AC_IN * row;
        return ((col >> 3) * L::ROWS + row) * 4 + ((col & 7) >> 1);
    };
    // tile t belongs to warpgroup (t / gridDim.x) % NWG of CTA t % gridDim.x: a short batch spreads over the SMs first
    int64_t tile = (int64_t)wg * gridDim.x + blockIdx.x;
    fetch(tile);
    uint32_t parity = 0;
    for (; tile < n_tiles; tile += stride) {
        // ---- stage the observation rows as fp16; rollout buffer (MAPPO.interact appends the state it acted on,
        // mappo.py:117-131): the rows pass through here anyway - stored now, not behind the loads, so that the prefetch
        // stays asynchronous ----
#pragma unroll
        for (int q = 0; q < PRE; ++q) {
            const int i = r + 128 * q;
            if (i < TILE_F4) {
                a1w[a1_word(4 * i)] = pack_h2(pre[q].x, pre[q].y);
                a1w[a1_word(4 * i + 2)] = pack_h2(pre[q].z, pre[q].w);
                if (obs_copy) {
                    const int64_t g = tile * (L::ROWS * AC_IN) + 4 * i;
                    if (g + 4 <= total_f) __stcs(reinterpret_cast<float4 *>(obs_copy + g), pre[q]);
                    else if (g + 2 <= total_f) __stcs(reinterpret_cast<float2 *>(obs_copy + g), make_float2(pre[q].x, pre[q].y));
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        wg_sync(wg);
        if (r == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 2; ++j)
                m5_mma(tmem, m5_desc(a1_addr + j * 2 * L::A_CHUNK, L::A_CHUNK), m5_desc(w1_addr + j * 2 * L::W1_CHUNK, L::W1_CHUNK),
                       L::IDESC1, j > 0);
            t5_commit(bar);
        }
        fetch(tile + stride);                                  // next tile's rows: in flight during this tile's MMAs and epilogues
        t5_wait(bar, parity); parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 1: h1 = relu(D1) -> A2 ----
#pragma unroll 1
        for (int col = 0; col < H1; col += 32) {
            uint32_t v[32];
            t5_ld32(t_lane + (uint32_t)col, v);
#pragma unroll
            for (int g = 0; g < 4; ++g)
                a2[(col / 8 + g) * L::ROWS + r] = make_uint4(pack_relu_h2(v[8 * g], v[8 * g + 1]), pack_relu_h2(v[8 * g + 2], v[8 * g + 3]),
                                                            pack_relu_h2(v[8 * g + 4], v[8 * g + 5]), pack_relu_h2(v[8 * g + 6], v[8 * g + 7]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        wg_sync(wg);
        if (r == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < H1 / 16; ++j)
                m5_mma(tmem, m5_desc(a2_addr + j * 2 * L::A_CHUNK, L::A_CHUNK), m5_desc(w2_addr + j * 2 * L::A_CHUNK, L::A_CHUNK),
                       L::IDESC2, j > 0);
            m5_mma(tmem, m5_desc(ones_addr, L::A_CHUNK), m5_desc(w2_addr + (H1 / 16) * 2 * L::A_CHUNK, L::A_CHUNK), L::IDESC2, 1);
            t5_commit(bar);
        }
        t5_wait(bar, parity); parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 2: h2 = relu(D2) -> the first 16 chunks of A2 (MMA2 has finished reading h1) ----
#pragma unroll 1
        for (int col = 0; col < AC_HID; col += 32) {
            uint32_t v[32];
            t5_ld32(t_lane + (uint32_t)col, v);
#pragma unroll
            for (int g = 0; g < 4; ++g)
                a2[(col / 8 + g) * L::ROWS + r] = make_uint4(pack_relu_h2(v[8 * g], v[8 * g + 1]), pack_relu_h2(v[8 * g + 2], v[8 * g + 3]),
                                                            pack_relu_h2(v[8 * g + 4], v[8 * g + 5]), pack_relu_h2(v[8 * g + 6], v[8 * g + 7]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        wg_sync(wg);
        // ---- output layer 128 -> 5 (+ value), N = 16, into the first columns of the region ----
        if (r == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < AC_HID / 16; ++j)
                m5_mma(tmem, m5_desc(a2_addr + j * 2 * L::A_CHUNK, L::A_CHUNK), m5_desc(w3_addr + j * 2 * L::W3_CHUNK, L::W3_CHUNK),
                       L::IDESC3, j > 0);
            m5_mma(tmem, m5_desc(ones_addr, L::A_CHUNK), m5_desc(w3_addr + 16 * L::W3_CHUNK, L::W3_CHUNK), L::IDESC3, 1);
            t5_commit(bar);
        }
        t5_wait(bar, parity); parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            const int64_t row = tile * L::ROWS + r;
            uint32_t v[16];
            m5_ld16(t_lane, v);
            if (row < n_rows) {
                float l[AC_OUT];
#pragma unroll
                for (int k = 0; k < AC_OUT; ++k) l[k] = __uint_as_float(v[k]);
                if (values) values[row] = __uint_as_float(v[AC_OUT]);
                apply_action_mask(l, mask_bits, row);
                float m = l[0];
#pragma unroll
                for (int k = 1; k < AC_OUT; ++k) m = fmaxf(m, l[k]);
                float e[AC_OUT], S = 0.f;
#pragma unroll
                for (int k = 0; k < AC_OUT; ++k) { e[k] = __expf(l[k] - m); S += e[k]; }
                const float logS = __logf(S);
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
                actions[row] = (int8_t)(live ? a : 1);
                if (live_out) live_out[row] = live ? 1 : 0;
                if (logp_sel) logp_sel[row] = l[a] - m - logS;
                if (logp_all) {
#pragma unroll
                    for (int k = 0; k < AC_OUT; ++k) logp_all[row * AC_OUT + k] = l[k] - m - logS;
                }
            }
        }
        // no barrier here: the next tile's first wg_sync orders these TMEM reads before its first MMA
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(*tmem_slot) : "memory");
}

// -1: not chosen yet; 0: tcgen05 fp16, a warpgroup per tile; 1: mma.sync TF32; 2: tcgen05 TF32, one CTA per SM; 3: tcgen05 fp16, two CTAs per SM
int g_actor_impl = -1;
void set_actor_impl(int impl) { g_actor_impl = (impl >= 1 && impl <= 3) ? impl : 0; }
static void resolve_actor_impl() {
    if (g_actor_impl >= 0) return;
    const char *e = getenv("MM_ACTOR_IMPL");
    g_actor_impl = !e ? 0 : (e[0] == 'm' ? 1 : (e[0] == 't' ? 2 : (e[0] == 'c' ? 3 : 0)));
}

template <int H1>
static int launch_actor_mlp_t(const float *obs, const int32_t *n_agents, int64_t n_rows, const float *w1, const float *b1,
                              const float *w2, const float *b2, const float *w3, const float *b3, const float *wv, const float *bv,
                              uint64_t seed, uint64_t step, const uint8_t *mask_bits, int8_t *actions, float *logp_all,
                              float *logp_sel, float *values, float *obs_copy, uint8_t *live_out, void *stream) {
    static bool ready[MM_MAX_DEVICES] = {};
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MM_MAX_DEVICES) return 1;
    if (!ready[dev]) {
        if (cudaFuncSetAttribute(actor_mlp_tcgen05_kernel<H1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M5<H1>::SMEM) != cudaSuccess)
            return 1;
        ready[dev] = true;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_actor_impl != 3) {                                       // default: one warpgroup per tile, one CTA per SM
        static bool wg_ready[MM_MAX_DEVICES] = {};
        if (!wg_ready[dev]) {
            if (cudaFuncSetAttribute(actor_mlp_wg_kernel<H1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M6<H1>::SMEM) != cudaSuccess)
                return 1;
            wg_ready[dev] = true;
        }
        const int64_t n_tiles = (n_rows + 127) / 128, grid = n_tiles < sms ? n_tiles : sms;
        actor_mlp_wg_kernel<H1><<<(unsigned)grid, M6<H1>::THREADS, M6<H1>::SMEM, (cudaStream_t)stream>>>(
            obs, n_agents, n_rows, w1, b1, w2, b2, w3, b3, wv, bv, seed, step, mask_bits, actions, logp_all, logp_sel, values, obs_copy,
            live_out);
        return cudaGetLastError() == cudaSuccess ? 0 : 1;
    }
    const int64_t tiles = (n_rows + 127) / 128, ctas = tiles < 2 * sms ? tiles : 2 * sms;   // persistent: two CTAs per SM
    actor_mlp_tcgen05_kernel<H1><<<(unsigned)ctas, 256, M5<H1>::SMEM, (cudaStream_t)stream>>>(
        obs, n_agents, n_rows, w1, b1, w2, b2, w3, b3, wv, bv, seed, step, mask_bits, actions, logp_all, logp_sel, values, obs_copy,
        live_out);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_actor_mlp(const float *obs, const int32_t *n_agents, int64_t n_rows, int h1, const float *w1, const float *b1,
                     const float *w2, const float *b2, const float *w3, const float *b3, const float *wv, const float *bv,
                     uint64_t seed, uint64_t step, const uint8_t *mask_bits, int8_t *actions, float *logp_all, float *logp_sel,
                     float *values, float *obs_copy, uint8_t *live_out, void *stream) {
    if (n_rows <= 0) return 0;
    resolve_actor_impl();
    if (h1 == 128)
        return launch_actor_mlp_t<128>(obs, n_agents, n_rows, w1, b1, w2, b2, w3, b3, wv, bv, seed, step, mask_bits, actions, logp_all,
                                       logp_sel, values, obs_copy, live_out, stream);
    if (h1 == 160)
        return launch_actor_mlp_t<160>(obs, n_agents, n_rows, w1, b1, w2, b2, w3, b3, wv, bv, seed, step, mask_bits, actions, logp_all,
                                       logp_sel, values, obs_copy, live_out, stream);
    return 2;
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
                        const uint8_t *mask_bits, int8_t *actions, float *logp_all, float *logp_sel, void *stream) {
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
    // g_actor_impl (mm_set_actor_impl; initial value from MM_ACTOR_IMPL=mma): 1 selects the warp-level mma.sync kernel
    // above, kept as the independent cross-check of the tcgen05 one
    resolve_actor_impl();
    if (g_actor_impl == 0 || g_actor_impl == 3)
        return launch_actor_mlp(obs, n_agents, n_rows, 128, w1, b1, w2, b2, w3, b3, nullptr, nullptr, seed, step, mask_bits, actions,
                                logp_all, logp_sel, nullptr, nullptr, nullptr, stream);
    if (g_actor_impl == 2) {
        static bool t5_attr = false;
        if (!t5_attr) {
            if (cudaFuncSetAttribute(actor_sample_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T5_SMEM) != cudaSuccess)
                return 1;
            t5_attr = true;
        }
        int64_t t5_tiles = (n_rows + T5_ROWS - 1) / T5_ROWS, t5_ctas = t5_tiles < sms ? t5_tiles : sms;
        actor_sample_tcgen05_kernel<<<(unsigned)t5_ctas, T5_THREADS, T5_SMEM, (cudaStream_t)stream>>>(
            obs, n_agents, n_rows, w1, b1, w2, b2, w3, b3, seed, step, mask_bits, actions, logp_all, logp_sel);
        return cudaGetLastError() == cudaSuccess ? 0 : 1;
    }
    int64_t tiles = (n_rows + 15) / 16, ctas = (tiles + AC_THREADS / 32 - 1) / (AC_THREADS / 32);
    if (ctas > sms) ctas = sms;   // persistent: one CTA per SM, warps stride over the 16-row tiles
    actor_sample_kernel<<<(unsigned)ctas, AC_THREADS, smem, (cudaStream_t)stream>>>(obs, n_agents, n_rows, w1, b1, w2, b2, w3,
                                                                                     b3, seed, step, mask_bits, actions, logp_all, logp_sel);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace mm
