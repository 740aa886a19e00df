// cdf_kernels.cu -- fused logits -> LQ32 CDF kernels (north-star part (a)) and the fused
// decode step (part (b), decoder side).
//
// One CTA of 1024 threads owns one logits row at a time and keeps it in registers: warp w
// owns a contiguous segment of the row, lanes stride through it with 128-bit loads.  The
// row is read from HBM exactly once:
//   phase A  row max                      (warp shuffle + one block barrier)
//   phase B  q_i, exact integer sums      (per lane -> per warp -> 32 warp totals in smem)
//   final    LOOKUP: cum[sym], cum[sym+1] from masked integer sums (no table written)
//            BUILD : whole table via in-warp exclusive scans
//            DECODE: hierarchical search warp -> 128-wide slab -> lane -> element for the
//                    symbol whose coder interval contains the code value, then the
//                    A_from_bin state update, all inside the kernel.
// The integer formulation (lq32.cuh) makes the result independent of this decomposition.
#include <cstdint>
#include <cuda_runtime.h>

#include "coder.cuh"
#include "lq32.cuh"

namespace lac {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kPerThread = 32;  // row elements held per thread

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = lq::vmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint64_t warp_sum(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// inclusive prefix over lanes
__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

struct DecShared {
    int64_t low, high, value;
    uint64_t pos;
    uint32_t status;
};

struct Smem {
    float red_max[kWarps];
    uint64_t wsum[kWarps];
    DecShared dec;
};

// Phases A and B for one row.  On return q[] holds this thread's q values (0 for slots
// outside the row), sm.wsum[] the 32 warp totals (valid after the trailing barrier).
template <int VEC>
__device__ __forceinline__ void row_reduce(const float* __restrict__ row, int G, int warp, int lane,
                                           Smem& sm, uint32_t (&q)[kPerThread], int& gbeg, int& gend) {
    constexpr int IT = kPerThread / VEC;
    gbeg = (int)(((int64_t)warp * G) / kWarps);
    gend = (int)(((int64_t)(warp + 1) * G) / kWarps);
    float x[kPerThread];
#pragma unroll
    for (int k = 0; k < IT; k++) {
        int g = gbeg + k * 32 + lane;
        if (VEC == 4) {
            float4 v = make_float4(lq::neg_inf(), lq::neg_inf(), lq::neg_inf(), lq::neg_inf());
            if (g < gend) v = ldg_stream4(row + 4 * (int64_t)g);
            x[4 * k + 0] = v.x;
            x[4 * k + 1] = v.y;
            x[4 * k + 2] = v.z;
            x[4 * k + 3] = v.w;
        } else {
            x[k] = g < gend ? ldg_stream1(row + g) : lq::neg_inf();
        }
    }
    float m = lq::neg_inf();
#pragma unroll
    for (int i = 0; i < kPerThread; i++) m = lq::vmax(m, x[i]);
    m = warp_max(m);
    if (lane == 0) sm.red_max[warp] = m;
    __syncthreads();
    m = warp_max(sm.red_max[lane]);
    uint64_t lane_sum = 0;
#pragma unroll
    for (int i = 0; i < kPerThread; i++) {
        q[i] = lq::q_of(x[i], m);
        lane_sum += q[i];
    }
    uint64_t ws = warp_sum(lane_sum);
    if (lane == 0) sm.wsum[warp] = ws;
    __syncthreads();
}

// Prefix of the warp totals: C at the start of `warp`'s segment, and Q.
__device__ __forceinline__ void warp_prefix(const Smem& sm, int warp, int lane, uint64_t& Cb, uint64_t& Q) {
    uint64_t v = sm.wsum[lane];
    uint64_t inc = warp_incl_scan(v, lane);
    Q = __shfl_sync(0xffffffffu, inc, 31);
    uint64_t exc = inc - v;
    Cb = __shfl_sync(0xffffffffu, exc, warp);
}

__device__ __forceinline__ lq::Scale bcast_scale(uint64_t Q, int V, int lane) {
    lq::Scale k;
    k.Q = Q;
    k.R = 0;
    k.s = 0;
    if (lane == 0) k = lq::make_scale(Q, V);
    k.R = __shfl_sync(0xffffffffu, k.R, 0);
    k.s = __shfl_sync(0xffffffffu, k.s, 0);
    return k;
}

// ------------------------------------------------------------------ LOOKUP
template <int VEC>
__global__ void __launch_bounds__(kThreads, 1)
lookup_kernel(const float* __restrict__ logits, int64_t rows, int V, int64_t row_stride,
              const int32_t* __restrict__ syms, uint32_t* __restrict__ pairs, uint32_t* __restrict__ status) {
    __shared__ Smem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = V / VEC;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        uint32_t q[kPerThread];
        int gbeg, gend;
        row_reduce<VEC>(logits + r * row_stride, G, warp, lane, sm, q, gbeg, gend);
        const int sym = syms[r];
        if (sym < 0 || sym >= V) {
            if (threadIdx.x == 0) {
                pairs[2 * r] = 0;
                pairs[2 * r + 1] = 0;
                if (status) atomicOr(status + r, LAC_ST_SYMBOL);
            }
            continue;
        }
        const int gs = sym / VEC, es = sym % VEC;
        if (gs < gbeg || gs >= gend) continue;  // not the owner warp
        uint64_t Cb, Q;
        warp_prefix(sm, warp, lane, Cb, Q);
        uint64_t part = 0, qs = 0;
        constexpr int IT = kPerThread / VEC;
#pragma unroll
        for (int k = 0; k < IT; k++) {
            int g = gbeg + k * 32 + lane;
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                uint32_t v = q[k * VEC + e];
                if (g < gs || (g == gs && e < es)) part += v;
                if (g == gs && e == es) qs = v;
            }
        }
        part = warp_sum(part);
        qs = warp_sum(qs);
        if (lane == 0) {
            lq::Scale sc = lq::make_scale(Q, V);
            uint64_t C = Cb + part;
            pairs[2 * r] = lq::cum_of(C, (uint32_t)sym, sc);
            pairs[2 * r + 1] = (sym == V - 1) ? 0u : lq::cum_of(C + qs, (uint32_t)sym + 1, sc);
        }
    }
}

// ------------------------------------------------------------------ BUILD
template <int VEC>
__global__ void __launch_bounds__(kThreads, 1)
build_kernel(const float* __restrict__ logits, int64_t rows, int V, int64_t row_stride,
             uint32_t* __restrict__ cum) {
    __shared__ Smem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = V / VEC;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        uint32_t q[kPerThread];
        int gbeg, gend;
        row_reduce<VEC>(logits + r * row_stride, G, warp, lane, sm, q, gbeg, gend);
        uint64_t base, Q;
        warp_prefix(sm, warp, lane, base, Q);
        lq::Scale sc = bcast_scale(Q, V, lane);
        uint32_t* out = cum + r * (int64_t)V;
        constexpr int IT = kPerThread / VEC;
#pragma unroll
        for (int k = 0; k < IT; k++) {
            if (gbeg + k * 32 >= gend) break;  // warp-uniform
            int g = gbeg + k * 32 + lane;
            uint64_t gsum = 0;
#pragma unroll
            for (int e = 0; e < VEC; e++) gsum += q[k * VEC + e];
            uint64_t inc = warp_incl_scan(gsum, lane);
            uint64_t C = base + inc - gsum;
            base += __shfl_sync(0xffffffffu, inc, 31);
            if (g < gend) {
                uint32_t o[VEC];
#pragma unroll
                for (int e = 0; e < VEC; e++) {
                    o[e] = lq::cum_of(C, (uint32_t)(g * VEC + e), sc);
                    C += q[k * VEC + e];
                }
                if (VEC == 4 && ((((uintptr_t)out) & 15) == 0)) {
                    *reinterpret_cast<uint4*>(out + 4 * (int64_t)g) = make_uint4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; e++) out[(int64_t)g * VEC + e] = o[e];
                }
            }
        }
    }
}

// ------------------------------------------------------------------ DECODE
// lowest coder offset of a symbol whose exclusive cumulative is c: ceil(c * w / 2^32)
// (symbol_to_range's l, arith_code.py:110-111); val_to_symbol (arith_code.py:94-101) picks
// the last symbol with lowpos <= value - l.
__device__ __forceinline__ uint64_t lowpos(uint32_t c, uint64_t w) { return coder::scale32_ceil(c, w); }

template <int VEC>
__global__ void __launch_bounds__(kThreads, 1)
decode_kernel(const float* __restrict__ logits, int64_t n_streams, int64_t T, int64_t stream_stride,
              int64_t tok_stride, int V, const int32_t* __restrict__ ntok, lac_dec_state* __restrict__ state,
              const uint8_t* __restrict__ bytes, const int64_t* __restrict__ offsets,
              int32_t* __restrict__ syms, int64_t sym_stride, int P) {
    __shared__ Smem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = V / VEC;
    constexpr int IT = kPerThread / VEC;
    for (int64_t s = blockIdx.x; s < n_streams; s += gridDim.x) {
        __syncthreads();  // previous stream's state fully written back
        if (threadIdx.x == 0) {
            sm.dec.low = state[s].low;
            sm.dec.high = state[s].high;
            sm.dec.value = state[s].value;
            sm.dec.pos = state[s].pos;
            sm.dec.status = state[s].status;
        }
        const uint8_t* data = bytes + offsets[s];
        const uint64_t nbytes = (uint64_t)(offsets[s + 1] - offsets[s]);
        const int64_t Ts = ntok ? (int64_t)ntok[s] : T;
        for (int64_t t = 0; t < Ts; t++) {
            uint32_t q[kPerThread];
            int gbeg, gend;
            // the two barriers inside also publish sm.dec written by the previous owner lane
            row_reduce<VEC>(logits + s * stream_stride + t * tok_stride, G, warp, lane, sm, q, gbeg, gend);
            const int64_t l = sm.dec.low, h = sm.dec.high;
            const uint64_t w = (uint64_t)(h - l + 1);
            const uint64_t xr = (uint64_t)(sm.dec.value - l);
            uint64_t Cb, Q;
            warp_prefix(sm, warp, lane, Cb, Q);
            lq::Scale sc = bcast_scale(Q, V, lane);
            const uint64_t Cn = Cb + sm.wsum[warp];
            const uint64_t low_b = lowpos(lq::cum_of(Cb, (uint32_t)(gbeg * VEC), sc), w);
            const uint64_t low_n = (gend == G) ? w : lowpos(lq::cum_of(Cn, (uint32_t)(gend * VEC), sc), w);
            if (!(gbeg < gend && low_b <= xr && xr < low_n)) continue;  // not the owner warp
            // ---- slab (k) level
            uint64_t Ck = Cb, Csel = Cb;
            int ksel = 0;
#pragma unroll
            for (int k = 0; k < IT; k++) {
                uint64_t gsum = 0;
#pragma unroll
                for (int e = 0; e < VEC; e++) gsum += q[k * VEC + e];
                uint64_t tot = warp_sum(gsum);
                int g0 = gbeg + k * 32;
                if (g0 < gend && lowpos(lq::cum_of(Ck, (uint32_t)(g0 * VEC), sc), w) <= xr) {
                    ksel = k;
                    Csel = Ck;
                }
                Ck += tot;
            }
            // ---- lane level inside slab ksel
            uint32_t qe[VEC];
            uint64_t gsum = 0;
#pragma unroll
            for (int k = 0; k < IT; k++) {
                if (k == ksel) {
#pragma unroll
                    for (int e = 0; e < VEC; e++) qe[e] = q[k * VEC + e];
                }
            }
#pragma unroll
            for (int e = 0; e < VEC; e++) gsum += qe[e];
            const int g = gbeg + ksel * 32 + lane;
            uint64_t inc = warp_incl_scan(gsum, lane);
            uint64_t C = Csel + inc - gsum;
            bool ok = g < gend && lowpos(lq::cum_of(C, (uint32_t)(g * VEC), sc), w) <= xr;
            unsigned ball = __ballot_sync(0xffffffffu, ok);
            int lsel = 31 - __clz((int)ball);
            if (lane != lsel) continue;
            // ---- element level (one lane)
            int sym = g * VEC;
            uint64_t Cs = C, Ce = C;
#pragma unroll
            for (int e = 1; e < VEC; e++) {
                Ce += qe[e - 1];
                if (lowpos(lq::cum_of(Ce, (uint32_t)(g * VEC + e), sc), w) <= xr) {
                    sym = g * VEC + e;
                    Cs = Ce;
                }
            }
            uint32_t qsym = qe[0];
#pragma unroll
            for (int e = 1; e < VEC; e++)
                if (sym == g * VEC + e) qsym = qe[e];
            const uint32_t lo = lq::cum_of(Cs, (uint32_t)sym, sc);
            const uint32_t hi = (sym == V - 1) ? 0u : lq::cum_of(Cs + qsym, (uint32_t)sym + 1, sc);
            // ---- A_from_bin.emit_symbol + emit_bit loop (arith_code.py:278-298)
            int64_t nl = l, nh = h;
            coder::ac_narrow32(nl, nh, lo, hi);
            const int64_t off = sm.dec.value - nl;  // value stays inside [nl, nh]
            const int k = coder::renorm_count((uint64_t)(nh - nl + 1), P);
            coder::renorm_apply(nl, nh, P, k);
            const uint64_t pos = sm.dec.pos;
            const uint64_t nb = coder::read_bits(data, nbytes, pos, k);
            sm.dec.low = nl;
            sm.dec.high = nh;
            sm.dec.value = nl + (off << k) + (int64_t)nb;
            sm.dec.pos = pos + (uint64_t)k;
            syms[s * sym_stride + t] = sym;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            state[s].low = sm.dec.low;
            state[s].high = sm.dec.high;
            state[s].value = sm.dec.value;
            state[s].pos = sm.dec.pos;
            state[s].status = sm.dec.status;
        }
    }
}

__global__ void dec_init_kernel(lac_dec_state* state, int64_t n, int P, const uint8_t* bytes,
                                const int64_t* offsets) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint8_t* data = bytes + offsets[s];
    uint64_t nbytes = (uint64_t)(offsets[s + 1] - offsets[s]);
    state[s].low = 0;
    state[s].high = (1ll << P) - 1;
    state[s].value = (int64_t)coder::read_bits(data, nbytes, 0, P);
    state[s].pos = (uint64_t)P;
    state[s].status = 0;
    state[s]._pad = 0;
}

// ------------------------------------------------------------------ launchers
static int grid_for(int64_t units) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)(units < sms ? (units > 0 ? units : 1) : sms);
}

static bool vec4_ok(const float* p, int V, int64_t s0, int64_t s1) {
    return (V % 4 == 0) && ((((uintptr_t)p) & 15) == 0) && (s0 % 4 == 0) && (s1 % 4 == 0);
}

cudaError_t launch_lookup(const float* logits, int64_t rows, int V, int64_t row_stride, const int32_t* syms,
                          uint32_t* pairs, uint32_t* status, cudaStream_t st) {
    if (rows == 0) return cudaSuccess;
    int grid = grid_for(rows);
    if (vec4_ok(logits, V, row_stride, 0))
        lookup_kernel<4><<<grid, kThreads, 0, st>>>(logits, rows, V, row_stride, syms, pairs, status);
    else
        lookup_kernel<1><<<grid, kThreads, 0, st>>>(logits, rows, V, row_stride, syms, pairs, status);
    return cudaGetLastError();
}

cudaError_t launch_build(const float* logits, int64_t rows, int V, int64_t row_stride, uint32_t* cum,
                         cudaStream_t st) {
    if (rows == 0) return cudaSuccess;
    int grid = grid_for(rows);
    if (vec4_ok(logits, V, row_stride, 0))
        build_kernel<4><<<grid, kThreads, 0, st>>>(logits, rows, V, row_stride, cum);
    else
        build_kernel<1><<<grid, kThreads, 0, st>>>(logits, rows, V, row_stride, cum);
    return cudaGetLastError();
}

cudaError_t launch_decode(const float* logits, int64_t n_streams, int64_t T, int64_t stream_stride,
                          int64_t tok_stride, int V, const int32_t* ntok, lac_dec_state* state,
                          const uint8_t* bytes, const int64_t* offsets, int32_t* syms, int64_t sym_stride,
                          int P, cudaStream_t st) {
    if (n_streams == 0 || T == 0) return cudaSuccess;
    int grid = grid_for(n_streams);
    if (vec4_ok(logits, V, stream_stride, tok_stride))
        decode_kernel<4><<<grid, kThreads, 0, st>>>(logits, n_streams, T, stream_stride, tok_stride, V, ntok,
                                                    state, bytes, offsets, syms, sym_stride, P);
    else
        decode_kernel<1><<<grid, kThreads, 0, st>>>(logits, n_streams, T, stream_stride, tok_stride, V, ntok,
                                                    state, bytes, offsets, syms, sym_stride, P);
    return cudaGetLastError();
}

cudaError_t launch_dec_init(lac_dec_state* state, int64_t n, int P, const uint8_t* bytes,
                            const int64_t* offsets, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    dec_init_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(state, n, P, bytes, offsets);
    return cudaGetLastError();
}

int max_vocab_single_cta() { return kThreads * kPerThread; }

}  // namespace lac
