// rowsum.cuh -- what the second passes do with a row summary (lq32.cuh): one warp turns the NW = 32 * CL summary
// words of a row into the row reference, the aligned segment weights, their prefixes and the row total, and
// evaluates symbol_to_range (arith_code.py:98-110) for one symbol by re-reading only that symbol's segment
// (<= 4 KB).  q is a function of (x, segment reference) only, so what is recomputed here is bit-identical to pass 1.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "lq32.cuh"
#include "ptx.cuh"

namespace lac {

constexpr int kPerThread = 32;  // row elements per thread of a warp segment (<= 1024 elements per segment)

// Lane l owns the CL consecutive segments l * CL .. l * CL + CL - 1 of the row.
template <int CL>
struct RowSum {
    uint64_t W[CL];     // load(): segment sums S_w;  align(): weights W_w;  scan(): exclusive prefixes of W_w
    uint32_t code[CL];  // segment reference codes r_w
    uint32_t r;         // row reference code (align())
    uint64_t Q;         // row total (scan() / total())

    __device__ __forceinline__ void load(const uint64_t* __restrict__ tab, int lane) {
#pragma unroll
        for (int j = 0; j < CL; j++) {
            const uint64_t w = __ldg(tab + lane * CL + j);
            code[j] = lq::word_code(w);
            W[j] = lq::word_sum(w);
        }
    }
    __device__ __forceinline__ void align() {
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < CL; j++) m = max(m, code[j]);
        r = __reduce_max_sync(0xffffffffu, m);
#pragma unroll
        for (int j = 0; j < CL; j++) W[j] = lq::shr64(W[j], lq::shift_of(r, code[j]));
    }
    // row total and the total of the segments in front of segment gw (masked REDUX sums; W stays as it is)
    __device__ __forceinline__ uint64_t total_and_front(int gw, int lane, uint64_t& front) {
        uint64_t qall = 0, qfront = 0;
#pragma unroll
        for (int j = 0; j < CL; j++) {
            qall += W[j];
            qfront += (lane * CL + j < gw) ? W[j] : 0ull;
        }
        front = warp_sum48(qfront);
        Q = warp_sum48(qall);
        return Q;
    }
    // W[j] := exclusive prefix of the weights at the start of segment lane * CL + j; Q := row total
    __device__ __forceinline__ void scan(int lane) {
        uint64_t tot = 0;
#pragma unroll
        for (int j = 0; j < CL; j++) tot += W[j];
        const uint64_t inc = warp_incl_scan(tot, lane);
        uint64_t run = inc - tot;
#pragma unroll
        for (int j = 0; j < CL; j++) {
            const uint64_t w = W[j];
            W[j] = run;
            run += w;
        }
        Q = __shfl_sync(0xffffffffu, inc, 31);
    }
    // reference code of segment gw, in every lane
    __device__ __forceinline__ uint32_t code_of(int gw) const {
        uint32_t c = code[0];
#pragma unroll
        for (int j = 1; j < CL; j++) c = (gw % CL == j) ? code[j] : c;
        return __shfl_sync(0xffffffffu, c, gw / CL);
    }
};

// The segment that holds 4-element group gs.
template <int CL>
__device__ __forceinline__ int seg_of_group(int gs, int G) {
    constexpr int NW = 32 * CL;
    int gw = (int)(((uint32_t)gs * (uint32_t)NW) / (uint32_t)G);  // +- 1
    gw = gw >= NW ? NW - 1 : gw;
    while (lq::seg_group<CL>(gw + 1, G) <= gs) gw++;
    while (lq::seg_group<CL>(gw, G) > gs) gw--;
    return gw;
}

// symbol_to_range on the total 2^32 for one symbol of one row, by one warp: (cum[sym], cum[sym + 1]), 0 = 2^32.
// VEC = 4: rows 16-byte aligned with V % 4 == 0 (128-bit loads); VEC = 1: anything else.  0 <= sym < V.
template <int VEC, int CL>
__device__ __forceinline__ uint2 warp_symbol_range(const float* __restrict__ row, int V, const uint64_t* __restrict__ tab,
                                                   int sym, int lane) {
    const int G = lq::groups_of(V);
    RowSum<CL> rs;
    rs.load(tab, lane);
    rs.align();
    const int gs = sym >> 2;
    const int gw = seg_of_group<CL>(gs, G);
    uint64_t front;
    const uint64_t Q = rs.total_and_front(gw, lane, front);
    const lq::Scale sc = lq::make_scale(Q, V);
    const uint32_t code = rs.code_of(gw);
    const int d = lq::shift_of(rs.r, code);
    const uint32_t nref = lq::nref_of_code(code);
    uint64_t part = 0;
    uint32_t qs = 0;
    // elements of the segment in front of (and including) the symbol: the slabs up to the symbol's (a warp-uniform
    // count, half of the segment on average), their loads issued together from addresses clamped to the symbol,
    // masked afterwards
    if (VEC == 4) {
        const int gbase = lq::seg_group<CL>(gw, G);
        const int g0 = gbase + lane;
        const int kmax = (gs - gbase) >> 5;  // slab of the symbol's group
        float4 x[kPerThread / 4];
#pragma unroll
        for (int k = 0; k < kPerThread / 4; k++)
            if (k <= kmax) x[k] = __ldg(reinterpret_cast<const float4*>(row) + min(g0 + 32 * k, gs));
#pragma unroll
        for (int k = 0; k < kPerThread / 4; k++) {
            if (k > kmax) break;
            const int g = g0 + 32 * k;
            uint32_t q0, q1, q2, q3;
            q_of2(x[k].x, x[k].y, nref, q0, q1);
            q_of2(x[k].z, x[k].w, nref, q2, q3);
            const int es = g < gs ? 4 : (g == gs ? (sym & 3) : -1);  // elements of this group in front of the symbol
            part += (uint64_t)(es > 0 ? q0 : 0u) + (es > 1 ? q1 : 0u) + (uint64_t)(es > 2 ? q2 : 0u) + (es > 3 ? q3 : 0u);
            if (g == gs) qs = es == 0 ? q0 : es == 1 ? q1 : es == 2 ? q2 : q3;
        }
    } else {
        const int ebase = 4 * lq::seg_group<CL>(gw, G);
        const int e0 = ebase + lane;
        const int kmax = (sym - ebase) >> 5;
        float x[kPerThread];
#pragma unroll
        for (int k = 0; k < kPerThread; k++)
            if (k <= kmax) x[k] = __ldg(row + min(e0 + 32 * k, sym));
#pragma unroll
        for (int k = 0; k < kPerThread; k++) {
            if (k > kmax) break;
            const int e = e0 + 32 * k;
            const uint32_t q0 = lq::q_of(x[k], nref);
            part += e < sym ? q0 : 0u;
            if (e == sym) qs = q0;
        }
    }
    part = warp_sum48(part);
    qs = __reduce_add_sync(0xffffffffu, qs);  // exactly one lane holds the symbol
    uint2 o;
    o.x = lq::cum_of(front + lq::shr64(part, d), (uint32_t)sym, sc);
    o.y = (sym == V - 1) ? 0u : lq::cum_of(front + lq::shr64(part + qs, d), (uint32_t)sym + 1, sc);
    return o;
}

}  // namespace lac
