// cdf_kernels.cu -- fused logits -> LQ32 CDF kernels (north-star part (a)) and the fused
// decode step (part (b), decoder side).
//
// One persistent CTA of 1024 threads per SM walks its rows.  Warp w owns a contiguous
// segment of the row; the row is read from HBM exactly once and then lives in registers:
//
//   staging  the row arrives as NCH TMA bulk copies (cp.async.bulk; 2 x 64 KB by default) into a
//            shared-memory ring, each chunk with its own mbarrier.  As soon as the warps of a chunk
//            have moved it to registers they re-arm the barrier and issue the bulk copy of the SAME
//            chunk of the CTA's NEXT row, so 128 KB per SM stay in flight during the compute phases.
//   phase A  row max                      (REDUX + the one block barrier of the row)
//   phase B  q_i with packed fp32x2 math (FADD2 / FFMA2), exact integer sums (lane -> warp)
//   finish   the LAST warp to finish phase B (shared-memory atomic counter) does the row-level
//            bookkeeping once -- prefix of the 32 warp totals, the scale (one division), for decode
//            the owner warp -- and arrives on the `done` mbarrier; there is no second block barrier.
//   final    LOOKUP: only the warp that owns the coded symbol waits for `done`; it writes
//                    cum[sym], cum[sym+1] from masked integer sums (no table is written); the other
//                    31 warps are already draining the next row.
//            BUILD : whole table via in-warp exclusive scans.
//            DECODE: 8 interleaved lane scans + ballots locate the 4-element group in the owner warp,
//                    one lane finishes the search and runs the A_from_bin state update.
// The integer formulation (lq32.cuh) makes the result independent of this decomposition.
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "coder.cuh"
#include "lq32.cuh"

namespace lac {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kPerThread = 32;  // row elements held per thread

// TMA chunks per row.  Measured on B200 (profiles/microbench/tma_stream.cu): every cp.async.bulk costs
// ~0.2 us of per-SM TMA time regardless of size, so 8 x 16 KB chunks cap at 4.7 TB/s while 2 x 64 KB
// reach 7.2 TB/s with the same 128 KB in flight.
constexpr int kMaxChunks = 8;
template <int NCH>
struct Ring {
    static constexpr int kWarpsPerChunk = kWarps / NCH;
    static constexpr int kSlotGroups = kPerThread / 4 * 32 * kWarpsPerChunk + 1;  // float4 groups per slot
    static constexpr int kSlotBytes = kSlotGroups * 16;
    static constexpr int kRingBytes = NCH * kSlotBytes;
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

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

// ------------------------------------------------------------------ packed fp32x2 (FADD2 / FFMA2)
__device__ __forceinline__ uint64_t pk2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// Two elements per instruction (FFMA2 / FADD2), exactly lq::q_of's operations (explicit FMAs: ptxas
// contracts packed mul+add pairs on its own, so the spec fuses them by definition):
// FFMA2, FADD2, FFMA2, 3 x FFMA2 per pair, then sub / shl / funnel-shift per element.  No F2I.
__device__ __forceinline__ void q_of2(float xa, float xb, uint32_t nref, uint32_t& qa, uint32_t& qb) {
    const uint64_t L2 = pk2(lq::log2e(), lq::log2e());
    const uint64_t MG = pk2(lq::magic(), lq::magic());
    const uint64_t MZ = pk2(lq::magicz(), lq::magicz());
    const uint64_t x2 = pk2(xa, xb);
    uint64_t t = fma2(x2, L2, MG);
    uint64_t rn = sub2(MG, t);
    uint64_t f = fma2(x2, L2, rn);
    uint64_t p = pk2(__uint_as_float(lq::kC3), __uint_as_float(lq::kC3));
    p = fma2(p, f, pk2(__uint_as_float(lq::kC2), __uint_as_float(lq::kC2)));
    p = fma2(p, f, pk2(__uint_as_float(lq::kC1), __uint_as_float(lq::kC1)));
    uint64_t z = fma2(p, f, MZ);
    float za, zb, ta, tb;
    upk2(z, za, zb);
    upk2(t, ta, tb);
    qa = __funnelshift_rc(__float_as_uint(za) << 7, 0u, nref - __float_as_uint(ta));
    qb = __funnelshift_rc(__float_as_uint(zb) << 7, 0u, nref - __float_as_uint(tb));
}

// ------------------------------------------------------------------ warp collectives (REDUX where possible)
__device__ __forceinline__ int f2ord(float f) {  // order-preserving float -> int (no NaNs reach here)
    int b = __float_as_int(f);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o ^ ((o >> 31) & 0x7fffffff)); }
// sum over the warp of values < 2^48: three 16-bit limbs through REDUX.SUM
__device__ __forceinline__ uint64_t warp_sum48(uint64_t v) {
    uint32_t a = __reduce_add_sync(0xffffffffu, (uint32_t)v & 0xffffu);
    uint32_t b = __reduce_add_sync(0xffffffffu, (uint32_t)(v >> 16) & 0xffffu);
    uint32_t c = __reduce_add_sync(0xffffffffu, (uint32_t)(v >> 32));
    return (uint64_t)a + ((uint64_t)b << 16) + ((uint64_t)c << 32);
}
__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// ------------------------------------------------------------------ shared control block
struct DecShared {
    int64_t low, high, value;
    uint64_t pos;             // next bit of the stream
    uint64_t win_hi, win_lo;  // 16 stream bytes starting at byte win_byte, big-endian
    uint64_t win_byte;
    int64_t data_off;
    uint64_t nbytes;
    uint32_t status;
};

struct Ctl {
    uint64_t full[kMaxChunks];  // TMA chunk landed
    uint64_t done;              // row-level bookkeeping published (phase = row parity)
    int red_max[2][kWarps];     // per-warp maxima (order-preserving ints), double-buffered by row parity
    uint64_t wsum[kWarps];      // per-warp totals of q
    uint64_t pref[kWarps];      // exclusive prefix of wsum   } written once per row by the last warp
    uint64_t Q;                 // sum of wsum                } to finish phase B, then `done` flips
    uint32_t R;                 // lq::Scale                  }
    int s;                      //                            }
    int owner;                  // decode: warp whose segment holds the symbol
    uint32_t arrive;            // warps done with phase B of the current row
    DecShared dec[2];
};

// Rows visited by this CTA, in order: outer index s = blockIdx.x, += gridDim.x; inner t < Ts.
// RowParams stays in the kernel-parameter constant bank; only (s, t, Ts) live in registers.
struct RowParams {
    const float* base;
    int64_t n_outer, T, so, st;
    const int32_t* ntok;
};
struct RowSeq {
    int64_t s;
    int t, Ts;
    __device__ __forceinline__ void skip_empty(const RowParams& p) {
        while (s < p.n_outer) {
            Ts = p.ntok ? p.ntok[s] : (int)p.T;
            if (Ts > 0) break;
            s += gridDim.x;
        }
    }
    __device__ __forceinline__ void init(const RowParams& p) {
        s = blockIdx.x;
        t = 0;
        Ts = 0;
        skip_empty(p);
    }
    __device__ __forceinline__ bool valid(const RowParams& p) const { return s < p.n_outer; }
    __device__ __forceinline__ const float* ptr(const RowParams& p) const { return p.base + s * p.so + t * p.st; }
    __device__ __forceinline__ void next(const RowParams& p) {
        if (++t >= Ts) {
            s += gridDim.x;
            t = 0;
            skip_empty(p);
        }
    }
};

// ------------------------------------------------------------------ row engine: staging + passes + finish
// File-scope shared objects have compile-time addresses, and everything about the row geometry is
// recomputed from (warp, V) on demand, so the engine keeps ONE register of state (the row counter):
// with 32 row elements per thread and a 64-register budget nothing else may stay live in the hot loop.
__shared__ Ctl g_ctl;
extern __shared__ __align__(128) unsigned char g_ring[];

template <int VEC, bool TMA, int NCH>
struct RowEngine {
    static constexpr int IT = kPerThread / VEC;
    static constexpr int kWarpsPerChunk = Ring<NCH>::kWarpsPerChunk;
    static constexpr int kSlotBytes = Ring<NCH>::kSlotBytes;
    uint32_t it;

    static __device__ __forceinline__ int warp() { return threadIdx.x >> 5; }
    static __device__ __forceinline__ int lane() { return threadIdx.x & 31; }
    static __device__ __forceinline__ int groups(int V) { return V / VEC; }
    static __device__ __forceinline__ int seg_begin(int w, int V) { return (int)(((int64_t)w * groups(V)) / kWarps); }
    static __device__ __forceinline__ int gbeg(int V) { return seg_begin(warp(), V); }
    static __device__ __forceinline__ int gend(int V) { return seg_begin(warp() + 1, V); }
    static __device__ __forceinline__ int chunk() { return warp() / kWarpsPerChunk; }
    static __device__ __forceinline__ int cg0(int V) { return seg_begin(chunk() * kWarpsPerChunk, V); }
    static __device__ __forceinline__ uint32_t cbytes(int V) {
        return (uint32_t)(seg_begin((chunk() + 1) * kWarpsPerChunk, V) - cg0(V)) * 16u;
    }
    static __device__ __forceinline__ bool leader() { return (threadIdx.x & (32 * kWarpsPerChunk - 1)) == 0; }
    static __device__ __forceinline__ unsigned char* slot() { return g_ring + chunk() * kSlotBytes; }

    __device__ void setup() {
        it = 0;
        if (threadIdx.x == 0) {
            g_ctl.arrive = 0;
            for (int i = 0; i < NCH; i++) mbar_init(&g_ctl.full[i], 1);
            mbar_init(&g_ctl.done, 1);
            fence_mbar_init();
        }
        __syncthreads();
    }
    static __device__ __forceinline__ void issue(const float* row, int V) {  // chunk leader: arm + bulk copy
        if (TMA && leader()) {
            const uint32_t nb = cbytes(V);
            if (nb) {
                mbar_expect_tx(&g_ctl.full[chunk()], nb);
                tma_load_1d(slot(), row + 4 * (int64_t)cg0(V), nb, &g_ctl.full[chunk()], evict_first_policy());
            }
        }
    }

    // One row: stage it into registers, phase A (maximum), the row's single block barrier, phase B (q, sums).
    // `next_row` (or nullptr) is prefetched as soon as this row has left shared memory.  On return q[] holds
    // this thread's final q values; the row-level results (g_ctl.pref, Q, R, s, owner) are valid once
    // wait_done() returns.
    __device__ __forceinline__ void reduce(const float* __restrict__ row, const float* next_row, int V,
                                           uint32_t (&q)[kPerThread], const DecShared* dec = nullptr) {
        float x[kPerThread];
        {
            const int gb = gbeg(V), ge = gend(V), ln = lane();
            if (TMA) {
                if (cbytes(V)) mbar_wait(&g_ctl.full[chunk()], it & 1);
                const unsigned char* sl = slot() - (size_t)cg0(V) * 16;
#pragma unroll
                for (int k = 0; k < IT; k++) {
                    int g = gb + k * 32 + ln;
                    float4 v = make_float4(lq::neg_inf(), lq::neg_inf(), lq::neg_inf(), lq::neg_inf());
                    if (g < ge) v = *reinterpret_cast<const float4*>(sl + (size_t)g * 16);
                    x[4 * k + 0] = v.x;
                    x[4 * k + 1] = v.y;
                    x[4 * k + 2] = v.z;
                    x[4 * k + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < IT; k++) {
                    int g = gb + k * 32 + ln;
                    if (VEC == 4) {
                        float4 v = make_float4(lq::neg_inf(), lq::neg_inf(), lq::neg_inf(), lq::neg_inf());
                        if (g < ge) v = ldg_stream4(row + 4 * (int64_t)g);
                        x[4 * k + 0] = v.x;
                        x[4 * k + 1] = v.y;
                        x[4 * k + 2] = v.z;
                        x[4 * k + 3] = v.w;
                    } else {
                        x[k] = g < ge ? ldg_stream1(row + g) : lq::neg_inf();
                    }
                }
            }
        }
        // phase A: row maximum (fmaxf drops NaNs; the running value starts at -inf, so it is never NaN)
        float m = lq::neg_inf();
#pragma unroll
        for (int i = 0; i < kPerThread; i++) m = fmaxf(m, x[i]);
        if (TMA) {
            // m depends on every shared-memory load of this thread, so after this barrier the chunk is
            // fully in registers and the slot can be overwritten by the next row
            named_bar_sync(1 + chunk(), 32 * kWarpsPerChunk);
            if (next_row) {
                if (leader()) fence_proxy_async();
                issue(next_row, V);
            }
        }
        const int mw = __reduce_max_sync(0xffffffffu, f2ord(m));
        int* red = g_ctl.red_max[it & 1];
        if (lane() == 0) red[warp()] = mw;
        __syncthreads();  // the only block-wide barrier of the row
        const int nref = lq::ref_of_max(ord2f(__reduce_max_sync(0xffffffffu, red[lane()])));
        // phase B: q against the row-wide reference, uint32 sums per 4 elements (4 q < 2^31.5), uint64 per lane
        // (a degenerate row -- no finite maximum, +inf, out of range -- gets a reference that pushes every
        // shift count past 31, i.e. q = 0 everywhere and the uniform table, without a second code path)
        const uint32_t nref_u = lq::ref_valid(nref) ? (uint32_t)nref : 0xFFFFFFFFu;
        uint64_t lane_sum = 0;
#pragma unroll
        for (int i = 0; i < kPerThread; i += 4) {
            q_of2(x[i], x[i + 1], nref_u, q[i], q[i + 1]);
            q_of2(x[i + 2], x[i + 3], nref_u, q[i + 2], q[i + 3]);
            lane_sum += (q[i] + q[i + 1]) + (q[i + 2] + q[i + 3]);
        }
        const uint64_t ws = warp_sum48(lane_sum);
        uint32_t prev = 0;
        if (lane() == 0) {
            g_ctl.wsum[warp()] = ws;
            fence_acq_rel_cta();
            prev = atomicAdd(&g_ctl.arrive, 1u);
        }
        prev = __shfl_sync(0xffffffffu, prev, 0);
        if (prev == kWarps - 1) finish_row(V, dec);  // last warp of the row: every wsum[] is visible
        it++;
    }

    // Row-level bookkeeping, run by exactly one warp per row.
    static __device__ __noinline__ void finish_row(int V, const DecShared* dec) {
        const int ln = lane();
        fence_acq_rel_cta();
        const uint64_t v = *reinterpret_cast<volatile uint64_t*>(&g_ctl.wsum[ln]);
        const uint64_t inc = warp_incl_scan(v, ln);
        const uint64_t Q = __shfl_sync(0xffffffffu, inc, 31);
        const uint64_t exc = inc - v;
        g_ctl.pref[ln] = exc;
        lq::Scale sc;
        sc.Q = Q;
        sc.R = 0;
        sc.s = 0;
        if (ln == 0) sc = lq::make_scale(Q, V);
        sc.R = __shfl_sync(0xffffffffu, sc.R, 0);
        sc.s = __shfl_sync(0xffffffffu, sc.s, 0);
        if (dec) {  // lane w tests warp w's segment start: the owner is the last non-empty one at or below the value
            const int gb = seg_begin(ln, V), ge = seg_begin(ln + 1, V);
            const uint64_t w = (uint64_t)(dec->high - dec->low + 1), xr = (uint64_t)(dec->value - dec->low);
            const bool ok = gb < ge && coder::scale32_ceil(lq::cum_of(exc, (uint32_t)(gb * VEC), sc), w) <= xr;
            const unsigned ball = __ballot_sync(0xffffffffu, ok);
            if (ln == 0) g_ctl.owner = 31 - __clz((int)ball);
        }
        __syncwarp();
        if (ln == 0) {
            g_ctl.Q = Q;
            g_ctl.R = sc.R;
            g_ctl.s = sc.s;
            g_ctl.arrive = 0;
            mbar_arrive(&g_ctl.done);  // release: publishes everything above
        }
    }
    // Block until the row-level results of the row just reduce()d are published.
    __device__ __forceinline__ void wait_done() const { mbar_wait(&g_ctl.done, (it - 1) & 1); }
    static __device__ __forceinline__ lq::Scale scale() {
        lq::Scale sc;
        sc.Q = g_ctl.Q;
        sc.R = g_ctl.R;
        sc.s = g_ctl.s;
        return sc;
    }
};

// ------------------------------------------------------------------ LOOKUP
template <int VEC, bool TMA, int NCH>
__global__ void __launch_bounds__(kThreads, 1)
lookup_kernel(const __grid_constant__ RowParams rp, int V, const int32_t* __restrict__ syms,
              uint32_t* __restrict__ pairs, uint32_t* __restrict__ status) {
    using Eng = RowEngine<VEC, TMA, NCH>;
    Ctl& ctl = g_ctl;
    Eng eng;
    eng.setup();
    RowSeq seq;
    seq.init(rp);
    if (seq.valid(rp)) Eng::issue(seq.ptr(rp), V);
    while (seq.valid(rp)) {
        const int64_t r = seq.s;
        const float* row = seq.ptr(rp);
        const int sym = __ldg(syms + r);  // issued now, consumed after the row's compute phases
        seq.next(rp);
        uint32_t q[kPerThread];
        eng.reduce(row, seq.valid(rp) ? seq.ptr(rp) : nullptr, V, q);
        const int warp = Eng::warp(), lane = Eng::lane(), gbeg = Eng::gbeg(V), gend = Eng::gend(V);
        if (sym < 0 || sym >= V) {
            if (threadIdx.x == 0) {
                *reinterpret_cast<uint2*>(pairs + 2 * r) = make_uint2(0u, 0u);
                if (status) atomicOr(status + r, LAC_ST_SYMBOL);
            }
            continue;
        }
        const int gs = sym / VEC, es = sym % VEC;
        if (gs >= gbeg && gs < gend) {  // owner warp; everyone else is already on the next row
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
            part = warp_sum48(part);
            qs = warp_sum48(qs);
            eng.wait_done();
            if (lane == 0) {
                const lq::Scale sc = Eng::scale();
                const uint64_t C = ctl.pref[warp] + part;
                uint2 o;
                o.x = lq::cum_of(C, (uint32_t)sym, sc);
                o.y = (sym == V - 1) ? 0u : lq::cum_of(C + qs, (uint32_t)sym + 1, sc);
                *reinterpret_cast<uint2*>(pairs + 2 * r) = o;
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------ BUILD
template <int VEC, bool TMA, int NCH>
__global__ void __launch_bounds__(kThreads, 1)
build_kernel(const __grid_constant__ RowParams rp, int V, uint32_t* __restrict__ cum) {
    using Eng = RowEngine<VEC, TMA, NCH>;
    Ctl& ctl = g_ctl;
    Eng eng;
    eng.setup();
    RowSeq seq;
    seq.init(rp);
    if (seq.valid(rp)) Eng::issue(seq.ptr(rp), V);
    while (seq.valid(rp)) {
        const int64_t r = seq.s;
        const float* row = seq.ptr(rp);
        seq.next(rp);
        uint32_t q[kPerThread];
        eng.reduce(row, seq.valid(rp) ? seq.ptr(rp) : nullptr, V, q);
        const int warp = Eng::warp(), lane = Eng::lane(), gbeg = Eng::gbeg(V), gend = Eng::gend(V);
        eng.wait_done();
        uint64_t base = ctl.pref[warp];
        const lq::Scale sc = Eng::scale();
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
// val_to_symbol (arith_code.py:94-101) on the total d = 2^32 picks the last symbol whose lowest
// coder offset ceil(cum * w / 2^32) (symbol_to_range's l, arith_code.py:110-111) is <= x = value - l.
// Comparing through the multiplication keeps 128-bit divisions off the per-token serial path.
__device__ __forceinline__ bool cum_le(uint32_t cum, uint64_t x, uint64_t w) { return coder::scale32_ceil(cum, w) <= x; }

// 8 stream bytes at byte offset b as a big-endian word, zeros past the end
__device__ __forceinline__ uint64_t load_be64(const uint8_t* data, uint64_t nbytes, uint64_t b) {
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) v = (v << 8) | (uint64_t)((b + i) < nbytes ? data[b + i] : 0);
    return v;
}

__device__ __forceinline__ void dec_load_state(DecShared& d, const lac_dec_state* st, const uint8_t* bytes,
                                               const int64_t* offsets, int64_t s) {
    d.low = st[s].low;
    d.high = st[s].high;
    d.value = st[s].value;
    d.pos = st[s].pos;
    d.status = st[s].status;
    d.data_off = offsets[s];
    d.nbytes = (uint64_t)(offsets[s + 1] - offsets[s]);
    d.win_byte = d.pos >> 3;
    d.win_hi = load_be64(bytes + d.data_off, d.nbytes, d.win_byte);
    d.win_lo = load_be64(bytes + d.data_off, d.nbytes, d.win_byte + 8);
}

template <int VEC, bool TMA, int NCH>
__global__ void __launch_bounds__(kThreads, 1)
decode_kernel(const __grid_constant__ RowParams rp, int V, lac_dec_state* __restrict__ state,
              const uint8_t* __restrict__ bytes, const int64_t* __restrict__ offsets,
              int32_t* __restrict__ syms, int64_t sym_stride, int P) {
    using Eng = RowEngine<VEC, TMA, NCH>;
    Ctl& ctl = g_ctl;
    Eng eng;
    eng.setup();
    constexpr int IT = kPerThread / VEC;
    RowSeq seq;
    seq.init(rp);
    if (seq.valid(rp)) Eng::issue(seq.ptr(rp), V);
    uint32_t par = 0;  // which DecShared buffer the current stream uses
    while (seq.valid(rp)) {
        const int64_t s = seq.s;
        const int t = seq.t;
        const bool first = (t == 0), last = (t == seq.Ts - 1);
        const float* row = seq.ptr(rp);
        seq.next(rp);
        if (first) {
            // double-buffered by stream parity: the previous stream's owner lane may still be
            // finishing with the other buffer; published by the block barrier inside reduce()
            par ^= 1;
            if (threadIdx.x == 0) dec_load_state(ctl.dec[par], state, bytes, offsets, s);
        }
        uint32_t q[kPerThread];
        DecShared& dec = ctl.dec[par];
        eng.reduce(row, seq.valid(rp) ? seq.ptr(rp) : nullptr, V, q, &dec);
        eng.wait_done();
        const int warp = Eng::warp(), lane = Eng::lane(), gbeg = Eng::gbeg(V), gend = Eng::gend(V);
        if (warp == ctl.owner) {
            const uint64_t w = (uint64_t)(dec.high - dec.low + 1);
            const uint64_t xr = (uint64_t)(dec.value - dec.low);
            const lq::Scale sc = Eng::scale();
            const uint64_t Cb = ctl.pref[warp];
            // ---- owner warp: IT interleaved lane scans of the group sums, then one ballot per slab.
            // Groups are ordered (slab k, lane); cum is monotone in that order, so the number of groups
            // at or below the value identifies the owner, and for the owning lane the last slab in which
            // its own group qualified is the owning slab.
            uint64_t inc[IT];
#pragma unroll
            for (int k = 0; k < IT; k++) {
                uint64_t a = 0;
#pragma unroll
                for (int e = 0; e < VEC; e++) a += q[k * VEC + e];
                inc[k] = a;
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
                for (int k = 0; k < IT; k++) {
                    uint64_t v = __shfl_up_sync(0xffffffffu, inc[k], o);
                    if (lane >= o) inc[k] += v;
                }
            }
            uint64_t base = Cb, C = 0;
            int cnt = 0, ksel = 0;
            uint32_t qe[VEC];
#pragma unroll
            for (int e = 0; e < VEC; e++) qe[e] = 0;
#pragma unroll
            for (int k = 0; k < IT; k++) {
                uint64_t a = 0;
#pragma unroll
                for (int e = 0; e < VEC; e++) a += q[k * VEC + e];
                const uint64_t Cg = base + inc[k] - a;
                base += __shfl_sync(0xffffffffu, inc[k], 31);
                const int gk = gbeg + k * 32 + lane;
                const bool ok = gk < gend && cum_le(lq::cum_of(Cg, (uint32_t)(gk * VEC), sc), xr, w);
                cnt += __popc(__ballot_sync(0xffffffffu, ok));
                if (ok) {
                    C = Cg;
                    ksel = k;
#pragma unroll
                    for (int e = 0; e < VEC; e++) qe[e] = q[k * VEC + e];
                }
            }
            const int j = cnt - 1;  // ordinal of the owning group in (slab, lane) order
            if (lane == (j & 31)) {
                // ---- element level (one lane)
                const int g = gbeg + ksel * 32 + lane;
                int sym = g * VEC;
                uint64_t Cs = C, Ce = C;
                uint32_t qsym = qe[0];
#pragma unroll
                for (int e = 1; e < VEC; e++) {
                    Ce += qe[e - 1];
                    if (cum_le(lq::cum_of(Ce, (uint32_t)(g * VEC + e), sc), xr, w)) {
                        sym = g * VEC + e;
                        Cs = Ce;
                        qsym = qe[e];
                    }
                }
                const uint32_t lo = lq::cum_of(Cs, (uint32_t)sym, sc);
                const uint32_t hi = (sym == V - 1) ? 0u : lq::cum_of(Cs + qsym, (uint32_t)sym + 1, sc);
                // ---- A_from_bin.emit_symbol + emit_bit loop (arith_code.py:278-298)
                int64_t nl = dec.low, nh = dec.high;
                const int64_t value = dec.value;
                coder::ac_narrow32(nl, nh, lo, hi);
                const int64_t off = value - nl;  // the value stays inside [nl, nh]
                const int k = coder::renorm_count((uint64_t)(nh - nl + 1), P);
                coder::renorm_apply(nl, nh, P, k);
                // next k bits from the 128-bit window (k <= 60, window offset < 64)
                const uint64_t pos = dec.pos;
                const uint64_t hi64 = dec.win_hi, lo64 = dec.win_lo, wb = dec.win_byte;
                const int o = (int)(pos - (wb << 3));
                const uint64_t comb = o ? ((hi64 << o) | (lo64 >> (64 - o))) : hi64;
                const uint64_t nb = k ? (comb >> (64 - k)) : 0;
                const int64_t nv = nl + (off << k) + (int64_t)nb;
                syms[s * sym_stride + t] = sym;
                if (last) {
                    state[s].low = nl;
                    state[s].high = nh;
                    state[s].value = nv;
                    state[s].pos = pos + (uint64_t)k;
                } else {
                    dec.low = nl;
                    dec.high = nh;
                    dec.value = nv;
                    dec.pos = pos + (uint64_t)k;
                    if (o + k >= 64) {  // slide the window by 8 bytes
                        dec.win_hi = lo64;
                        dec.win_byte = wb + 8;
                        dec.win_lo = load_be64(bytes + dec.data_off, dec.nbytes, wb + 16);
                    }
                }
            }
            __syncwarp();
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

// 0: scalar LDG (any alignment), 1: 128-bit LDG, 2 / 4 / 8: TMA bulk ring with that many chunks per row
// (default 2 when rows are 16-byte aligned; LAC_NO_TMA=1 and LAC_TMA_CHUNKS=n are measurement switches)
static int path_for(const float* p, int V, int64_t s0, int64_t s1) {
    bool v4 = (V % 4 == 0) && ((((uintptr_t)p) & 15) == 0) && (s0 % 4 == 0) && (s1 % 4 == 0);
    if (!v4) return 0;
    static const bool no_tma = getenv("LAC_NO_TMA") != nullptr;
    if (no_tma) return 1;
    static const int nch = getenv("LAC_TMA_CHUNKS") ? atoi(getenv("LAC_TMA_CHUNKS")) : 2;
    return (nch == 4 || nch == 8) ? nch : 2;
}

template <typename K, typename... A>
static cudaError_t launch_tma(K kernel, int ring_bytes, int grid, cudaStream_t st, A... args) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes);
    if (e != cudaSuccess) return e;
    kernel<<<grid, kThreads, ring_bytes, st>>>(args...);
    return cudaGetLastError();
}

cudaError_t launch_lookup(const float* logits, int64_t rows, int V, int64_t row_stride, const int32_t* syms,
                          uint32_t* pairs, uint32_t* status, cudaStream_t st) {
    if (rows == 0) return cudaSuccess;
    int grid = grid_for(rows);
    const RowParams rp{logits, rows, 1, row_stride, 0, nullptr};
    switch (path_for(logits, V, row_stride, 0)) {
        case 2: return launch_tma(lookup_kernel<4, true, 2>, Ring<2>::kRingBytes, grid, st, rp, V, syms, pairs, status);
        case 4: return launch_tma(lookup_kernel<4, true, 4>, Ring<4>::kRingBytes, grid, st, rp, V, syms, pairs, status);
        case 8: return launch_tma(lookup_kernel<4, true, 8>, Ring<8>::kRingBytes, grid, st, rp, V, syms, pairs, status);
        case 1: lookup_kernel<4, false, 8><<<grid, kThreads, 0, st>>>(rp, V, syms, pairs, status); break;
        default: lookup_kernel<1, false, 8><<<grid, kThreads, 0, st>>>(rp, V, syms, pairs, status);
    }
    return cudaGetLastError();
}

cudaError_t launch_build(const float* logits, int64_t rows, int V, int64_t row_stride, uint32_t* cum,
                         cudaStream_t st) {
    if (rows == 0) return cudaSuccess;
    int grid = grid_for(rows);
    const RowParams rp{logits, rows, 1, row_stride, 0, nullptr};
    switch (path_for(logits, V, row_stride, 0)) {
        case 2: return launch_tma(build_kernel<4, true, 2>, Ring<2>::kRingBytes, grid, st, rp, V, cum);
        case 4: return launch_tma(build_kernel<4, true, 4>, Ring<4>::kRingBytes, grid, st, rp, V, cum);
        case 8: return launch_tma(build_kernel<4, true, 8>, Ring<8>::kRingBytes, grid, st, rp, V, cum);
        case 1: build_kernel<4, false, 8><<<grid, kThreads, 0, st>>>(rp, V, cum); break;
        default: build_kernel<1, false, 8><<<grid, kThreads, 0, st>>>(rp, V, cum);
    }
    return cudaGetLastError();
}

cudaError_t launch_decode(const float* logits, int64_t n_streams, int64_t T, int64_t stream_stride,
                          int64_t tok_stride, int V, const int32_t* ntok, lac_dec_state* state,
                          const uint8_t* bytes, const int64_t* offsets, int32_t* syms, int64_t sym_stride,
                          int P, cudaStream_t st) {
    if (n_streams == 0 || T == 0) return cudaSuccess;
    int grid = grid_for(n_streams);
    const RowParams rp{logits, n_streams, T, stream_stride, tok_stride, ntok};
    switch (path_for(logits, V, stream_stride, tok_stride)) {
        case 2: return launch_tma(decode_kernel<4, true, 2>, Ring<2>::kRingBytes, grid, st, rp, V, state, bytes, offsets, syms, sym_stride, P);
        case 4: return launch_tma(decode_kernel<4, true, 4>, Ring<4>::kRingBytes, grid, st, rp, V, state, bytes, offsets, syms, sym_stride, P);
        case 8: return launch_tma(decode_kernel<4, true, 8>, Ring<8>::kRingBytes, grid, st, rp, V, state, bytes, offsets, syms, sym_stride, P);
        case 1: decode_kernel<4, false, 8><<<grid, kThreads, 0, st>>>(rp, V, state, bytes, offsets, syms, sym_stride, P); break;
        default: decode_kernel<1, false, 8><<<grid, kThreads, 0, st>>>(rp, V, state, bytes, offsets, syms, sym_stride, P);
    }
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
