// mm_philox.cuh - Philox4x32-10 counter-based generator: the device-side spawn (reset_kernel) and the supervisors' draws.
// Keyed (seed; stream = global env index; episode counter [; a fourth word]): any env, episode and step can be
// generated independently, with no state carried between launches.
#pragma once
#include <stdint.h>

namespace mm {

struct Philox {
    uint32_t key[2], ctr[4], out[4];
    int have;
    __device__ Philox(uint64_t seed, uint64_t stream, uint32_t episode, uint32_t word0 = 0) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        ctr[0] = word0; ctr[1] = episode; ctr[2] = (uint32_t)stream; ctr[3] = (uint32_t)(stream >> 32);
        have = 0;
    }
    __device__ void round_(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    __device__ uint32_t next() {
        if (have == 0) {
            uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
            uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
            for (int r = 0; r < 10; ++r) {
                round_(c, k0, k1);
                k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
            }
            out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
            ctr[0]++;
            have = 4;
        }
        return out[--have];
    }
    __device__ double uniform() {  // [0, 1) with 53 bits
        uint64_t a = next(), b = next();
        return (double)(((a << 21) ^ b) & ((1ull << 53) - 1)) * (1.0 / 9007199254740992.0);
    }
    __device__ int below(int n) { return (int)(((uint64_t)next() * (uint64_t)n) >> 32); }
};

}  // namespace mm
