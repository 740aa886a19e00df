// lq32.cuh -- the LQ32 logits -> integer CDF quantisation, device side.
//
// Everything here is defined in individually rounded IEEE fp32 operations plus exact
// integer arithmetic, so oracle/lac_oracle.c (orc_lq32_*) reproduces it bit for bit and the
// result does not depend on how a row is split across threads, warps or CTAs:
//
//   m    = max_i x_i                               (NaN dropped)
//   d_i  = x_i - m
//   t_i  = fma(d_i, log2e, 1.5 * 2^23)             low mantissa bits = n_i = rne(d_i * log2e)
//   f_i  = fma(d_i, log2e, 1.5 * 2^23 - t_i)       in [-0.5, 0.5], single rounding
//   z_i  = fma(fma(fma(c3, f, c2), f, c1), f, 1.5 * 2^23)               c_k = minimax 2^f coefficients * 2^22
//   P_i  = bits(z_i) & 0x7FFFFF = rne(2^22 * 2^f_i)  (max rel. error 1.02e-4, i.e. < 1e-8 bits/token; no F2I: the XU pipe that
//          executes float->int conversions runs at ~4 threads/clk/SM on B200 and was the measured limiter)
//   q_i  = n_i < -31 (or NaN / -inf) ? 0 : (P_i << 9) >> -n_i    (integer, <= 2^31)
//   Q    = sum_i q_i,  C_i = sum_{j<i} q_j          (exact, order independent)
//   s    = bitlen(Q) - 1,  R = floor(((2^32 - V) << 31) / ((Q >> (s - 31)) + 1))
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
constexpr uint32_t kC1 = 0x4a317afdu, kC2 = 0x49780626u, kC3 = 0x4861510cu;  // minimax 2^f, degree 3, * 2^22

__device__ __forceinline__ float neg_inf() { return __uint_as_float(0xFF800000u); }

// fmaxf drops a NaN operand, like PTX max.f32 and the oracle's lq_max.
__device__ __forceinline__ float vmax(float a, float b) { return fmaxf(a, b); }

__device__ __forceinline__ uint32_t q_of(float x, float m) {
    float d = __fsub_rn(x, m);
    float t = __fmaf_rn(d, log2e(), magic());
    uint32_t sh = 0x4B400000u - __float_as_uint(t);  // -n for n in [-31, 0]; >= 32 otherwise
    float f = __fmaf_rn(d, log2e(), __fsub_rn(magic(), t));
    float p = __uint_as_float(kC3);
    p = __fmaf_rn(p, f, __uint_as_float(kC2));
    p = __fmaf_rn(p, f, __uint_as_float(kC1));
    float z = __fmaf_rn(p, f, magic());
    return __funnelshift_rc(__float_as_uint(z) << 9, 0u, sh);  // (P << 9) >> min(sh, 32)
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
    if (Q >= (1ull << 31)) {  // else: the degenerate Q == 0 row (no finite maximum)
        k.s = 63 - __clzll((long long)Q);  // >= 31: the row maximum contributes q = 2^31
        const uint64_t N = ((1ull << 32) - (uint64_t)V) << 31;
        const uint64_t D = (Q >> (k.s - 31)) + 1;  // (2^31, 2^32]
        // R = floor(N / D) <= M * 2^s / Q.  fp64 estimate (off by at most 1) + exact fix-up:
        // far shorter dependent chain than the generic 64-bit division.
        uint64_t r = (uint64_t)__ddiv_rz((double)N, (double)D);
        int64_t rem = (int64_t)(N - r * D);
        if (rem < 0) r -= 1;
        else if (rem >= (int64_t)D) r += 1;
        k.R = (uint32_t)r;
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
