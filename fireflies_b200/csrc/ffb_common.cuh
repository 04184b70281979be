// Shared helpers for libffb200.so (sm_100a).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ffb200.h"

namespace ffb {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- error plumbing -------------------------------------------------------------------------
char* last_error_buf();   // thread-local, defined in ffb_api.cu

inline int fail_arg(int code, const char* msg) {
    snprintf(last_error_buf(), 256, "%s", msg);
    return code;
}
inline int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    snprintf(last_error_buf(), 256, "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
#define FFB_CUDA(call)                                        \
    do {                                                      \
        int _rc = ::ffb::check_cuda((call), #call);           \
        if (_rc) return _rc;                                  \
    } while (0)
#define FFB_LAUNCH_CHECK(name) FFB_CUDA((cudaPeekAtLastError(), cudaGetLastError()))

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- device math ------------------------------------------------------------------------------
// exp(-w) for w >= 0.  Cody-Waite reduction by log2(e) in two parts, ex2.approx on |f| <= 0.5 and an
// exponent-field add: relative error ~2 ulp independent of w (the plain __expf form grows with w and
// would eat the 1e-5 parity budget in the tails).  Inputs are clamped at 87 (result 1.6e-38).
__device__ __forceinline__ float exp_neg(float w) {
#ifdef FFB_FAST_EXP
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(w * -1.4426950408889634f));
    return r;
#else
    w = fminf(w, 87.0f);
    const float t = w * -1.4426950408889634f;
    const float tm = t + 12582912.0f;                 // 1.5 * 2^23: low mantissa bits hold rint(t)
    const float n = tm - 12582912.0f;
    float f = fmaf(w, -1.4426950216293335f, -n);      // hi part of log2(e) (exact product-difference)
    f = fmaf(w, -1.9259629911266175e-8f, f);          // lo part
    float p;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(f));
    return __int_as_float(__float_as_int(p) + (__float_as_int(tm) << 23));
#endif
}

// x / d with d's reciprocal precomputed: one Newton correction gives the correctly rounded quotient
// (bar rare double-rounding cases) without the MUFU + slow-path branch of the generic division.
__device__ __forceinline__ float div_by(float x, float d, float rcp_d) {
    const float q = x * rcp_d;
    const float r = fmaf(-q, d, x);
    return fmaf(r, rcp_d, q);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- Philox4x32-10 (counter-based RNG; Salmon et al. 2011) -----------------------------------------
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
        const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
#else
        const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    // counter = (c0,c1,c2,c3), key = seed
    __host__ __device__ static inline void gen(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t (&out)[4]) {
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            round(out, k0, k1);
            k0 += W0; k1 += W1;
        }
    }
    // uniform in [0,1) with 24 random bits -- the same lattice torch.rand uses for fp32
    __host__ __device__ static inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
};

}  // namespace ffb
