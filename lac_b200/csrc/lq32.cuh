// lq32.cuh -- the LQ32 logits -> integer CDF quantisation, device side.
//
// Everything here is defined in individually rounded IEEE fp32 operations plus exact
// integer arithmetic, so oracle/lac_oracle.c (orc_lq32_*) reproduces it bit for bit and the
// result does not depend on how a row is split across threads, warps or CTAs:
//
//   m    = max_i x_i                               (NaN dropped)
//   y_i  = max((x_i - m) * log2e, -64)
//   n_i  = rne(y_i), f_i = y_i - n_i in [-0.5, 0.5]
//   P_i  = trunc(poly4(f_i))  ~ 2^f_i * 2^31       (max rel. error 2.9e-6)
//   q_i  = P_i >> -n_i                              (integer, <= 2^31)
//   Q    = sum_i q_i,  C_i = sum_{j<i} q_j          (exact, order independent)
//   s    = bitlen(Q) - 1,  R = floor((2^32 - V) * 2^s / Q)
//   cum_i = ((C_i * R) >> s) + i,  cum_V = 2^32     => every frequency >= 1
//
// Replaces the reference's float table builders (llama_compress.py:24-30,
// arithmetic_coding.py:59-64); the quantisation differs from theirs by design, bound in
// DESIGN.md section 3.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lq {

__device__ __forceinline__ float log2e() { return __uint_as_float(0x3FB8AA3Bu); }
__device__ __forceinline__ float magic() { return __uint_as_float(0x4B400000u); }  // 1.5 * 2^23
constexpr uint32_t kC0 = 0x4f000000u, kC1 = 0x4eb17096u, kC2 = 0x4df601bcu, kC3 = 0x4ce4fe23u,
                   kC4 = 0x4b9d0163u;

__device__ __forceinline__ float neg_inf() { return __uint_as_float(0xFF800000u); }

// fmaxf drops a NaN operand, like PTX max.f32 and the oracle's lq_max.
__device__ __forceinline__ float vmax(float a, float b) { return fmaxf(a, b); }

__device__ __forceinline__ uint32_t q_of(float x, float m) {
    float d = __fsub_rn(x, m);
    float y = fmaxf(__fmul_rn(d, log2e()), -64.0f);
    float t = __fadd_rn(y, magic());
    uint32_t sh = 0x4B400000u - __float_as_uint(t);  // -n, in [0, 64]
    float r = __fsub_rn(t, magic());
    float f = __fsub_rn(y, r);
    float p = __uint_as_float(kC4);
    p = __fmaf_rn(p, f, __uint_as_float(kC3));
    p = __fmaf_rn(p, f, __uint_as_float(kC2));
    p = __fmaf_rn(p, f, __uint_as_float(kC1));
    p = __fmaf_rn(p, f, __uint_as_float(kC0));
    uint32_t P = __float2uint_rz(p);
    return __funnelshift_rc(P, 0u, sh);  // P >> min(sh, 32)
}

struct Scale {
    uint64_t Q;
    uint32_t R;
    int s;
};

__device__ __forceinline__ Scale make_scale(uint64_t Q, int V) {
    Scale k;
    k.Q = Q;
    k.R = 0;
    k.s = 0;
    if (Q != 0) {
        k.s = 63 - __clzll((long long)Q);
        unsigned __int128 M = (((unsigned __int128)1) << 32) - (unsigned __int128)V;
        k.R = (uint32_t)((M << k.s) / Q);
    }
    return k;
}

// ((C * R) >> s) + i ; C < 2^49, R < 2^32, s <= 48.
__device__ __forceinline__ uint32_t cum_of(uint64_t C, uint32_t i, const Scale& k) {
    uint64_t lo = C * (uint64_t)k.R;
    uint64_t hi = __umul64hi(C, (uint64_t)k.R);
    uint64_t v = k.s == 0 ? lo : ((lo >> k.s) | (hi << (64 - k.s)));
    return (uint32_t)v + i;
}

}  // namespace lq
