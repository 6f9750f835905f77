// Shared device/host helpers for the sm_100a kernels: error plumbing, Philox4x32-10, warp primitives.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/se_b200.h"

namespace se {

void set_error(const char *fmt, ...);
int check_cuda(cudaError_t err, const char *what);
int sm_count();

#define SE_REQUIRE(cond, ...)                \
    do {                                     \
        if (!(cond)) {                       \
            se::set_error(__VA_ARGS__);      \
            return SE_ERR_INVALID_ARG;       \
        }                                    \
    } while (0)

#define SE_CUDA(expr)                                         \
    do {                                                      \
        int _rc = se::check_cuda((expr), #expr);              \
        if (_rc != SE_OK) return _rc;                         \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter-based, so every (walk, step, try) / (centre, context, k) draw is
// addressable and independent of launch geometry and of how walks are sharded over GPUs.
// The same function is restated in numpy in tests/philox_ref.py.
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;

// domain separators (counter word 3, high bits)
constexpr uint32_t STREAM_WALK = 0x10000000u;
constexpr uint32_t STREAM_NEG = 0x20000000u;
constexpr uint32_t STREAM_DRAW = 0x30000000u;
constexpr uint32_t STREAM_NEG_COIN = 0x40000000u;

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = mulhi32(PHILOX_M0, ctr.x), lo0 = PHILOX_M0 * ctr.x;
        uint32_t hi1 = mulhi32(PHILOX_M1, ctr.z), lo1 = PHILOX_M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ k0, lo1, hi0 ^ ctr.w ^ k1, lo0);
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    return ctr;
}

__device__ __forceinline__ uint4 philox(uint64_t seed, uint64_t id, uint32_t sub, uint32_t stream) {
    return philox4x32_10(make_uint4((uint32_t)id, (uint32_t)(id >> 32), sub, stream), (uint32_t)seed,
                         (uint32_t)(seed >> 32));
}

// uniform in [0,1) with 24 bits (exactly representable in fp32)
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

// Alias-table / uniform draw from two 32-bit randoms.
__device__ __forceinline__ int64_t draw_row(const float *__restrict__ prob, const int32_t *__restrict__ alias,
                                            uint32_t vocab, uint32_t r0, uint32_t r1) {
    uint32_t j = mulhi32(r0, vocab);
    if (prob != nullptr) {
        // both loads are issued back to back (the alias entry is fetched whether or not the coin needs it): one memory
        // latency on the critical path of every negative instead of two
        const float pj = __ldg(prob + j);
        const uint32_t aj = (uint32_t)__ldg(alias + j);
        if (u01(r1) >= pj) j = aj;
    }
    return (int64_t)j;
}

__device__ __forceinline__ uint32_t pick_word(const uint4 &r, int i) {
    return i == 0 ? r.x : i == 1 ? r.y : i == 2 ? r.z : r.w;
}

// Negative k of context n of centre `centre` is keyed independently of launch geometry and embedding size:
//   bucket word = word (n & 3) of Philox(seed; centre, n >> 2, STREAM_NEG | k)
//   coin word   = word (n & 3) of Philox(seed; centre, n >> 2, STREAM_NEG_COIN | k)      (alias tables only)
// so one Philox call serves four consecutive contexts.  tests/philox_ref.py restates this.
__device__ __forceinline__ uint4 neg_words(uint64_t seed, uint64_t centre, int n, int k, uint32_t stream) {
    return philox(seed, centre, (uint32_t)(n >> 2), stream | (uint32_t)k);
}

constexpr unsigned FULL = 0xffffffffu;

}  // namespace se
