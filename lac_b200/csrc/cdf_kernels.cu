// cdf_kernels.cu -- logits -> LQ32 CDF kernels (north-star part (a)) and the decoder side of part (b).
//
// Both directions are ONE bandwidth-bound pass over the logits plus small dependent passes:
//
//   encode:  summary_kernel -> pair_kernel (1 warp / row)             -> encode_pairs_kernel (coder_kernels.cu)
//   decode:  summary_kernel -> decode_serial_kernel (1 warp / stream, walks its tokens)
//
// summary_kernel: one persistent CTA of 1024 threads per SM walks its rows.  Warp w owns a contiguous segment
// of the row; the row is read from HBM exactly once and then lives in registers:
//
//   staging  the row arrives as NCH TMA bulk copies (cp.async.bulk; 4 x 32 KB, in a cluster 2 x 64 KB) into a
//            shared-memory ring, each chunk with its own mbarrier; the warps of a chunk move it to registers.
//   phase A  row max (REDUX + the ONE barrier of the row).  Right after it the chunk leaders re-arm their
//            mbarriers and issue the bulk copies of the CTA's NEXT row, so 128 KB per SM are in flight
//            during the compute phase.  In a cluster the CTA maxima then cross DSMEM (st.async).
//   phase B  q_i with packed fp32x2 math (FADD2 / FFMA2), exact integer sums (lane -> warp)
//   finish   lane 0 of EVERY warp stores the warp's total into the row summary {nref, wsum[32 * CL]} (264 bytes
//            per 32000-element row): no counter, no finishing warp, no scan, no division, no table.
//            (build_kernel, the full-table variant, keeps a last-arriving warp that derives prefixes and scale
//            and publishes them through the `done` mbarrier.)
//
// The second passes work from the summary and re-read only the one warp segment (<= 4 KB) they need; q is a
// function of (x, nref) only (lq32.cuh), so the recomputed values are bit-identical to pass 1 and the result does
// not depend on the decomposition.
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
// Row summary written by pass 1, uint64 words: [0] row reference nref, [1 + 32 c + w] the total of q of warp w of
// CTA part c.  Every warp writes its own word; nothing is reduced across warps in pass 1.
__host__ __device__ constexpr int summ_words(int cl) { return 1 + cl * 32; }

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

// ------------------------------------------------------------------ thread-block cluster helpers
// A row wider than one CTA can hold (32768 elements) is split over a cluster of CL CTAs; the two
// row-level reductions travel through distributed shared memory with cluster-scope mbarriers.
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
// One message to a peer CTA: the value lands in its shared memory and the same operation completes `bytes` of the
// transaction count of its mbarrier (st.async, SASS: STAS).  No fence on the sending side -- an
// st.shared::cluster + mbarrier.arrive.release.cluster pair costs MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR per message,
// about a microsecond under full memory load, which every CTA of the cluster then waits for.
__device__ __forceinline__ void st_async_u32(uint32_t addr, uint32_t v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(addr), "r"(v),
                 "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void st_async_u64(uint32_t addr, uint64_t v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u64 [%0], %1, [%2];" ::"r"(addr), "l"(v),
                 "r"(bar)
                 : "memory");
}
// Wait for the messages of the peer CTAs (a transaction barrier armed by this CTA, completed by their st.async).
// The data is in THIS CTA's shared memory and is read with ordinary shared-memory loads after the wait, so the
// default CTA-scope acquire is the right one (a cluster-scope acquire makes ptxas emit CCTL.IVALL, an L1D
// invalidate, after every wait).
__device__ __forceinline__ void mbar_wait_peers(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CL>
struct Clu {
    static __device__ __forceinline__ uint32_t rank() {
        if (CL == 1) return 0;
        uint32_t r;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
        return r;
    }
    static __device__ __forceinline__ uint32_t id() { return CL == 1 ? blockIdx.x : blockIdx.x / CL; }
    static __device__ __forceinline__ uint32_t count() { return CL == 1 ? gridDim.x : gridDim.x / CL; }
};

// ------------------------------------------------------------------ shared control block
constexpr int kMaxCluster = 8;
struct Ctl {
    uint64_t full[kMaxChunks];  // TMA chunk landed
    uint64_t done;              // row-level bookkeeping published (phase = row parity)
    uint64_t cl_max_bar[2];     // cluster: every CTA's maximum has arrived (transaction barrier, CL messages per
    uint64_t cl_sum_bar[2];     // cluster: every CTA's total has arrived    phase); one barrier per row parity, so a
                                //                                           peer one row ahead cannot alias
    uint64_t cl_Q[2][kMaxCluster];  // per-CTA totals, double-buffered by row parity
    int cl_max[2][kMaxCluster];     // per-CTA maxima (ordered ints)
    int red_max[2][kWarps];     // per-warp maxima (order-preserving ints), double-buffered by row parity
    uint64_t wsum[kWarps];      // per-warp totals of q
    uint64_t pref[kWarps];      // row-wide exclusive prefix at each warp } written once per row by the last
    uint64_t Q;                 // row total                               } warp to finish phase B, then
    uint32_t R;                 // lq::Scale                               } `done` flips
    int s;                      //                                         }
    uint32_t arrive;            // warps done with phase B of the current row
};

// Rows visited by this CTA / cluster, in order: outer index s = first, += stride; inner t < Ts.
// RowParams stays in the kernel-parameter constant bank; only (s, t, Ts) live in registers.
struct RowParams {
    const float* base;
    int64_t n_outer, T, so, st;
    const int32_t* ntok;  // per outer index: tokens present (nullptr: T everywhere)
    int64_t t0;           // ntok counts from t0 tokens before `base` (token-chunked decode), so ntok[s] - t0 remain
};
__device__ __forceinline__ int tokens_of(const RowParams& p, int64_t s) {
    if (!p.ntok) return (int)p.T;
    const int64_t n = (int64_t)p.ntok[s] - p.t0;
    return (int)(n < 0 ? 0 : (n > p.T ? p.T : n));
}
struct RowSeq {
    int64_t s;
    int t, Ts;
    __device__ __forceinline__ void skip_empty(const RowParams& p, uint32_t stride) {
        while (s < p.n_outer) {
            Ts = tokens_of(p, s);
            if (Ts > 0) break;
            s += stride;
        }
    }
    __device__ __forceinline__ void init(const RowParams& p, uint32_t first, uint32_t stride) {
        s = first;
        t = 0;
        Ts = 0;
        skip_empty(p, stride);
    }
    __device__ __forceinline__ bool valid(const RowParams& p) const { return s < p.n_outer; }
    __device__ __forceinline__ const float* ptr(const RowParams& p) const { return p.base + s * p.so + t * p.st; }
    __device__ __forceinline__ void next(const RowParams& p, uint32_t stride) {
        if (++t >= Ts) {
            s += stride;
            t = 0;
            skip_empty(p, stride);
        }
    }
};
// All rows (s, t) of the launch dealt round-robin to the CTAs / clusters regardless of the stream they belong to
// (flat index s * T + t = first + i * stride), rows past a stream's token count skipped.  No division per row.
struct FlatSeq {
    int64_t s, ds;
    int t, dt;
    __device__ __forceinline__ void skip_absent(const RowParams& p) {
        while (s < p.n_outer && p.ntok && t >= tokens_of(p, s)) step(p);
    }
    __device__ __forceinline__ void step(const RowParams& p) {
        s += ds;
        t += dt;
        if (t >= (int)p.T) {
            t -= (int)p.T;
            s++;
        }
    }
    __device__ __forceinline__ void init(const RowParams& p, uint32_t first, uint32_t stride) {
        s = first / p.T;
        t = (int)(first % p.T);
        ds = stride / p.T;
        dt = (int)(stride % p.T);
        skip_absent(p);
    }
    __device__ __forceinline__ bool valid(const RowParams& p) const { return s < p.n_outer; }
    __device__ __forceinline__ const float* ptr(const RowParams& p) const { return p.base + s * p.so + t * p.st; }
    __device__ __forceinline__ int64_t index(const RowParams& p) const { return s * p.T + t; }
    __device__ __forceinline__ void next(const RowParams& p) {
        step(p);
        skip_absent(p);
    }
};

// ------------------------------------------------------------------ row engine: staging + phases + finish
// File-scope shared objects have compile-time addresses, and everything about the row geometry is
// recomputed from (warp, V) on demand, so the engine keeps ONE register of state (the row counter):
// with 32 row elements per thread and a 64-register budget nothing else may stay live in the hot loop.
__shared__ Ctl g_ctl;
extern __shared__ __align__(128) unsigned char g_ring[];

struct NoSummary {  // build_kernel: the finishing warp derives prefixes and scale inside the kernel
    static constexpr bool kSummary = false;
    __device__ __forceinline__ uint64_t* operator()() const { return nullptr; }
};
struct SummaryAt {  // summary_kernel: where this row's summary goes
    static constexpr bool kSummary = true;
    uint64_t* p;
    __device__ __forceinline__ uint64_t* operator()() const { return p; }
};

template <int VEC, bool TMA, int NCH, int CL>
struct RowEngine {
    static_assert(CL == 1 || (TMA && VEC == 4), "cluster rows use the TMA path");
    static constexpr int IT = kPerThread / VEC;
    static constexpr int kWarpsPerChunk = Ring<NCH>::kWarpsPerChunk;
    static constexpr int kSlotBytes = Ring<NCH>::kSlotBytes;
    static constexpr int kTotWarps = CL * kWarps;
    uint32_t it;

    static __device__ __forceinline__ int warp() { return threadIdx.x >> 5; }
    static __device__ __forceinline__ int lane() { return threadIdx.x & 31; }
    static __device__ __forceinline__ int gwarp() { return (int)Clu<CL>::rank() * kWarps + warp(); }
    static __device__ __forceinline__ int groups(int V) { return V / VEC; }
    // first float4 group of (row-wide) warp gw: the row is cut evenly over the CL * 32 warps of the cluster
    static __device__ __forceinline__ int seg_begin(int gw, int V) { return (int)(((int64_t)gw * groups(V)) / kTotWarps); }
    static __device__ __forceinline__ int gbeg(int V) { return seg_begin(gwarp(), V); }
    static __device__ __forceinline__ int gend(int V) { return seg_begin(gwarp() + 1, V); }
    static __device__ __forceinline__ int chunk() { return warp() / kWarpsPerChunk; }
    static __device__ __forceinline__ int chunk_gw0() { return (int)Clu<CL>::rank() * kWarps + chunk() * kWarpsPerChunk; }
    static __device__ __forceinline__ int cg0(int V) { return seg_begin(chunk_gw0(), V); }
    static __device__ __forceinline__ uint32_t cbytes(int V) {
        return (uint32_t)(seg_begin(chunk_gw0() + kWarpsPerChunk, V) - cg0(V)) * 16u;
    }
    static __device__ __forceinline__ bool leader() { return (threadIdx.x & (32 * kWarpsPerChunk - 1)) == 0; }
    static __device__ __forceinline__ unsigned char* slot() { return g_ring + chunk() * kSlotBytes; }

    __device__ void setup() {
        it = 0;
        if (threadIdx.x == 0) {
            g_ctl.arrive = 0;
            for (int i = 0; i < NCH; i++) mbar_init(&g_ctl.full[i], 1);
            mbar_init(&g_ctl.done, 1);
            for (int i = 0; i < 2; i++) {
                mbar_init(&g_ctl.cl_max_bar[i], 1);  // one local arrive.expect_tx per phase + CL peer messages
                mbar_init(&g_ctl.cl_sum_bar[i], 1);
            }
            fence_mbar_init();
        }
        __syncthreads();
        if (CL > 1) cluster_sync_all();  // nobody may signal a peer whose barriers are not initialised yet
    }
    static __device__ void teardown() {
        if (CL > 1) cluster_sync_all();  // nobody may exit while a peer can still write into its shared memory
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

    // One row: stage it into registers, phase A (maximum), the row's single block barrier (plus, in a cluster,
    // the exchange of the CTA maxima), phase B (q, sums).  `next_row` (or nullptr) is prefetched as soon as this
    // row has left shared memory.  On return q[] holds this thread's q values; the row-level results
    // (g_ctl.pref, Q, R, s) are valid once wait_done() returns -- unless SummFn::kSummary, in which case every warp
    // just stores its total into the row summary and there is no row-level bookkeeping at all.
    // next_row() and summ() are evaluated lazily, by the chunk leaders resp. lane 0 of each warp only: pointer
    // arithmetic every thread would otherwise redo per row costs issue slots the row loop does not have.
    template <class NextFn, class SummFn>
    __device__ __forceinline__ void reduce(const float* __restrict__ row, NextFn next_row, int V,
                                           uint32_t (&q)[kPerThread], SummFn summ) {
        float x[kPerThread];
        {
            const int gb = gbeg(V), ge = gend(V), ln = lane();
            if (TMA) {
                if (cbytes(V)) mbar_wait(&g_ctl.full[chunk()], it & 1);
                const unsigned char* sl = slot() + (size_t)(gb + ln - cg0(V)) * 16;
                if (ge - gb > (IT - 1) * 32) {
                    // the usual shape (a warp segment longer than 7 slabs): slabs 0..6 are complete for every lane,
                    // no -inf initialisation and no predicates there (32 instructions per thread and row less)
#pragma unroll
                    for (int k = 0; k < IT - 1; k++) {
                        const float4 v = *reinterpret_cast<const float4*>(sl + k * 512);
                        x[4 * k + 0] = v.x;
                        x[4 * k + 1] = v.y;
                        x[4 * k + 2] = v.z;
                        x[4 * k + 3] = v.w;
                    }
                    float4 v = make_float4(lq::neg_inf(), lq::neg_inf(), lq::neg_inf(), lq::neg_inf());
                    if (gb + (IT - 1) * 32 + ln < ge) v = *reinterpret_cast<const float4*>(sl + (IT - 1) * 512);
                    x[4 * (IT - 1) + 0] = v.x;
                    x[4 * (IT - 1) + 1] = v.y;
                    x[4 * (IT - 1) + 2] = v.z;
                    x[4 * (IT - 1) + 3] = v.w;
                } else {
#pragma unroll
                    for (int k = 0; k < IT; k++) {
                        float4 v = make_float4(lq::neg_inf(), lq::neg_inf(), lq::neg_inf(), lq::neg_inf());
                        if (gb + k * 32 + ln < ge) v = *reinterpret_cast<const float4*>(sl + k * 512);
                        x[4 * k + 0] = v.x;
                        x[4 * k + 1] = v.y;
                        x[4 * k + 2] = v.z;
                        x[4 * k + 3] = v.w;
                    }
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
        const int mw = __reduce_max_sync(0xffffffffu, f2ord(m));
        const uint32_t par = it & 1;
        int* red = g_ctl.red_max[par];
        if (lane() == 0) red[warp()] = mw;
        __syncthreads();  // the only barrier of the row
        if (TMA && leader()) {
            // m depends on every shared-memory load of a thread, so after the barrier the row is fully in
            // registers and both slots can be overwritten by the next row
            const float* nr = next_row();
            if (nr) {
                fence_proxy_async();
                issue(nr, V);
            }
        }
        int mx = __reduce_max_sync(0xffffffffu, red[lane()]);
        if (CL > 1 && threadIdx.x < CL) {  // send this CTA's maximum to every CTA of the cluster (incl. itself)
            if (threadIdx.x == 0) mbar_expect_tx(&g_ctl.cl_max_bar[par], CL * 4);
            st_async_u32(mapa(smem_u32(&g_ctl.cl_max[par][Clu<CL>::rank()]), threadIdx.x), (uint32_t)mx,
                         mapa(smem_u32(&g_ctl.cl_max_bar[par]), threadIdx.x));
        }
        // phase B against the row-wide reference (in a cluster: after the peers' maxima have arrived)
        if (CL > 1) {
            mbar_wait_peers(&g_ctl.cl_max_bar[par], (it >> 1) & 1);
#pragma unroll
            for (int p = 0; p < CL; p++) mx = max(mx, g_ctl.cl_max[par][p]);
        }
        const int nloc = lq::ref_of_max(ord2f(mx));
        const uint32_t nref_u = lq::ref_valid(nloc) ? (uint32_t)nloc : 0xFFFFFFFFu;
        const uint32_t nrow_u = nref_u;
        uint64_t lane_sum = 0;
#pragma unroll
        for (int i = 0; i < kPerThread; i += 4) {
            q_of2(x[i], x[i + 1], nref_u, q[i], q[i + 1]);
            q_of2(x[i + 2], x[i + 3], nref_u, q[i + 2], q[i + 3]);
            lane_sum += (q[i] + q[i + 1]) + (q[i + 2] + q[i + 3]);
        }
        const uint64_t ws = warp_sum48(lane_sum);
        if (SummFn::kSummary) {
            // pass 1 of lookup / decode: every warp stores its own total; no counter, no finishing warp, no scan,
            // nothing for the other warps to wait for at the next row's barrier
            if (lane() == 0) {
                uint64_t* out = summ();
                out[1 + gwarp()] = ws;
                if (gwarp() == 0) out[0] = (uint64_t)nrow_u;
            }
            it++;
            return;
        }
        uint32_t prev = 0;
        if (lane() == 0) {
            g_ctl.wsum[warp()] = ws;
            fence_acq_rel_cta();
            // (plain PTX: atomicAdd() under a lane test is compiled into a warp-aggregation sequence)
            asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(prev) : "r"(smem_u32(&g_ctl.arrive)) : "memory");
        }
        prev = __shfl_sync(0xffffffffu, prev, 0);
        if (prev == kWarps - 1) finish_row(V, par, (it >> 1) & 1);  // last warp of the row: every wsum[] is visible
        it++;
    }

    // Row-level bookkeeping, run by exactly one warp per CTA per row.
    static __device__ __noinline__ void finish_row(int V, uint32_t par, uint32_t ph) {
        const int ln = lane();
        fence_acq_rel_cta();
        const uint64_t v = *reinterpret_cast<volatile uint64_t*>(&g_ctl.wsum[ln]);
        const uint64_t inc = warp_incl_scan(v, ln);
        uint64_t Q = __shfl_sync(0xffffffffu, inc, 31);
        uint64_t base = 0;
        if (CL > 1) {  // exchange the CTA totals; base = total of the lower-ranked CTAs
            if (ln < CL) {
                if (ln == 0) mbar_expect_tx(&g_ctl.cl_sum_bar[par], CL * 8);
                st_async_u64(mapa(smem_u32(&g_ctl.cl_Q[par][Clu<CL>::rank()]), ln), Q,
                             mapa(smem_u32(&g_ctl.cl_sum_bar[par]), ln));
            }
            mbar_wait_peers(&g_ctl.cl_sum_bar[par], ph);
            uint64_t tot = 0;
#pragma unroll
            for (int p = 0; p < CL; p++) {
                const uint64_t qp = g_ctl.cl_Q[par][p];
                if (p < (int)Clu<CL>::rank()) base += qp;
                tot += qp;
            }
            Q = tot;
        }
        const uint64_t exc = base + inc - v;  // row-wide exclusive prefix at the start of local warp `ln`
        g_ctl.pref[ln] = exc;
        const lq::Scale sc = lq::make_scale(Q, V);  // one division, the same in every lane
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

// ------------------------------------------------------------------ LOOKUP (encode side, second pass)
// symbol_to_range (arith_code.py:87-93) on the total 2^32 for one coded symbol per row: (cum[sym], cum[sym + 1]).
// Pass 1 is summary_kernel (below); this pass is one warp per row, all rows independent: find the warp segment
// of the symbol, re-read the part of it in front of the symbol (<= 4 KB, half of that on average), q against the
// row reference, masked integer sums, two multiply-shifts.  ~3 % extra HBM traffic, a few microseconds per 16k rows.
template <int VEC, int CL>
__global__ void __launch_bounds__(256)
pair_kernel(const float* __restrict__ logits, int64_t rows, int64_t row_stride, int V,
            const uint64_t* __restrict__ summ, const int32_t* __restrict__ syms, uint32_t* __restrict__ pairs,
            uint32_t* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    // everything that does not depend on the symbol is requested together with it
    const uint64_t* tab = summ + r * summ_words(CL);
    const int sym = __ldg(syms + r);
    const int nref = (int)(uint32_t)__ldg(tab);
    uint64_t wsum[CL];
#pragma unroll
    for (int c = 0; c < CL; c++) wsum[c] = __ldg(tab + 1 + 32 * c + lane);
    if (sym < 0 || sym >= V) {
        if (lane == 0) {
            *reinterpret_cast<uint2*>(pairs + 2 * r) = make_uint2(0u, 0u);
            if (status) atomicOr(status + r, LAC_ST_SYMBOL);
        }
        return;
    }
    constexpr int tw = kWarps * CL;
    const int groups = V / VEC;
    auto seg = [&](int gw) { return (int)(((int64_t)gw * groups) / tw); };  // RowEngine::seg_begin
    const int gs = sym / VEC;
    int gw = (int)(((uint32_t)gs * (uint32_t)tw) / (uint32_t)groups);  // the warp whose segment holds group gs (+- 1)
    gw = gw >= tw ? tw - 1 : gw;
    while (seg(gw + 1) <= gs) gw++;
    while (seg(gw) > gs) gw--;
    // row total and the total of the warp segments in front of gw, from the 32 * CL warp sums (masked REDUX sums)
    uint64_t qall = 0, qfront = 0;
#pragma unroll
    for (int c = 0; c < CL; c++) {
        qall += wsum[c];
        qfront += (32 * c + lane < gw) ? wsum[c] : 0ull;
    }
    const uint64_t Q = warp_sum48(qall);
    const uint64_t C = warp_sum48(qfront);
    const lq::Scale sc = lq::make_scale(Q, V);
    const float* row = logits + r * row_stride;
    uint64_t part = 0;
    uint32_t qs = 0;
    // groups in front of (and including) the symbol's group: at most 8 per lane, all loads issued together from
    // addresses clamped to the symbol's group, masked afterwards
    const int g0 = seg(gw) + lane;
    if (VEC == 4) {
        float4 x[kPerThread / 4];
#pragma unroll
        for (int k = 0; k < kPerThread / 4; k++) x[k] = __ldg(reinterpret_cast<const float4*>(row) + min(g0 + 32 * k, gs));
#pragma unroll
        for (int k = 0; k < kPerThread / 4; k++) {
            const int g = g0 + 32 * k;
            uint32_t q0, q1, q2, q3;
            q_of2(x[k].x, x[k].y, (uint32_t)nref, q0, q1);
            q_of2(x[k].z, x[k].w, (uint32_t)nref, q2, q3);
            const int es = g < gs ? 4 : (g == gs ? (sym & 3) : -1);  // elements of this group in front of the symbol
            part += (uint64_t)(es > 0 ? q0 : 0u) + (es > 1 ? q1 : 0u) + (uint64_t)(es > 2 ? q2 : 0u) + (es > 3 ? q3 : 0u);
            if (g == gs) qs = es == 0 ? q0 : es == 1 ? q1 : es == 2 ? q2 : q3;
        }
    } else {
        float x[kPerThread];
#pragma unroll
        for (int k = 0; k < kPerThread; k++) x[k] = __ldg(row + min(g0 + 32 * k, gs));
#pragma unroll
        for (int k = 0; k < kPerThread; k++) {
            const int g = g0 + 32 * k;
            const uint32_t q0 = lq::q_of(x[k], nref);
            part += g < gs ? q0 : 0u;
            if (g == gs) qs = q0;
        }
    }
    part = warp_sum48(part);
    qs = __reduce_add_sync(0xffffffffu, qs);  // exactly one lane holds the symbol
    if (lane == 0) {
        uint2 o;
        o.x = lq::cum_of(C + part, (uint32_t)sym, sc);
        o.y = (sym == V - 1) ? 0u : lq::cum_of(C + part + qs, (uint32_t)sym + 1, sc);
        *reinterpret_cast<uint2*>(pairs + 2 * r) = o;
    }
}

// ------------------------------------------------------------------ BUILD
template <int VEC, bool TMA, int NCH, int CL>
__global__ void __launch_bounds__(kThreads, 1)
build_kernel(const __grid_constant__ RowParams rp, int V, uint32_t* __restrict__ cum) {
    using Eng = RowEngine<VEC, TMA, NCH, CL>;
    Ctl& ctl = g_ctl;
    Eng eng;
    eng.setup();
    const uint32_t stride = Clu<CL>::count();
    RowSeq seq;
    seq.init(rp, Clu<CL>::id(), stride);
    if (seq.valid(rp)) Eng::issue(seq.ptr(rp), V);
    while (seq.valid(rp)) {
        const int64_t r = seq.s;
        const float* row = seq.ptr(rp);
        seq.next(rp, stride);
        uint32_t q[kPerThread];
        eng.reduce(row, [&] { return seq.valid(rp) ? seq.ptr(rp) : nullptr; }, V, q, NoSummary());
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
    Eng::teardown();
}

// ------------------------------------------------------------------ DECODE
// val_to_symbol (arith_code.py:94-101) on the total d = 2^32: bisect_right(dist, target) with
// target = ((value - l) * 2^32) // w, i.e. the last symbol whose exclusive cumulative is <= target.  The probe is
// computed once per row in finish_row; every boundary test is then a 96-bit multiply-shift and a compare.

// 8 stream bytes at byte offset b as a big-endian word, zeros past the end
// Assembled as two 32-bit halves from unconditional loads at clamped addresses plus a mask.  (The obvious
// form -- a 64-bit accumulator fed by predicated byte loads -- was observed to be miscompiled by ptxas 12.9 in
// one instantiation: a CS2R-zeroed register pair was read 5 cycles later still holding its previous content,
// which corrupted the window of streams shorter than 22 bytes.  tests/test_gpu_parity.py pins that case.)
__device__ __forceinline__ uint64_t load_be64(const uint8_t* data, uint64_t nbytes, uint64_t b) {
    if (b >= nbytes) return 0;
    uint32_t w[2] = {0u, 0u};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint64_t idx = b + i;
        const bool in = idx < nbytes;
        const uint32_t byte = (uint32_t)data[in ? idx : b] & (in ? 0xFFu : 0u);
        w[i >> 2] = (w[i >> 2] << 8) | byte;
    }
    return ((uint64_t)w[0] << 32) | w[1];
}

// ---- pass 1: row summaries.  The same engine as LOOKUP without an owner: rows are dealt to the CTAs / clusters
// regardless of their stream (a 4-stream job still fills the machine), and the only output is the summary --
// 8 + 256 * CL bytes per row next to the 4 * V bytes read, one word per warp.  No division and, in a cluster, no
// exchange of totals: the only cluster traffic left is the row maximum.
template <int VEC, bool TMA, int NCH, int CL>
__global__ void __launch_bounds__(kThreads, 1)
summary_kernel(const __grid_constant__ RowParams rp, int V, uint64_t* __restrict__ summ) {
    using Eng = RowEngine<VEC, TMA, NCH, CL>;
    Eng eng;
    eng.setup();
    FlatSeq seq;
    seq.init(rp, Clu<CL>::id(), Clu<CL>::count());
    if (seq.valid(rp)) Eng::issue(seq.ptr(rp), V);
    while (seq.valid(rp)) {
        const int64_t idx = seq.index(rp);
        const float* row = seq.ptr(rp);
        seq.next(rp);
        uint32_t q[kPerThread];
        eng.reduce(row, [&] { return seq.valid(rp) ? seq.ptr(rp) : nullptr; }, V, q,
                   SummaryAt{summ + idx * summ_words(CL)});
    }
    Eng::teardown();
}

// ---- pass 2: the serial part.  One warp per stream; every lane carries the same A_from_bin state
// (arith_code.py:233-306) in registers.  Per token:
//   probe    target = floor(((value - low) << 32) / w)                        (val_to_symbol, arith_code.py:94-101)
//   level 1  the last non-empty warp segment of the row whose first cumulative is <= target (summary prefixes)
//   level 2  that segment (<= 1024 elements, <= 4 KB) is read again, lane-major (lane l = 32 consecutive elements),
//            q recomputed against the row reference (bit-identical to pass 1 by construction of LQ32), one warp scan
//            + ballot picks the lane, the rest of the search is independent work inside each lane
//   update   narrow, renormalise by k bits at once, pull k bits from a 24-byte register window of the stream
// The logits segment comes from HBM (pass 1 streamed the rows with evict-first), ~3 % extra traffic.
// The summary of token t + 1 is loaded while token t is being searched and its scale (the one division that does
// not depend on the coder state) is computed while token t's segment is in flight.
template <int CL>
struct RowTab {
    uint64_t nref, Q, pre[CL];  // pre[j]: row-wide exclusive prefix of q at the start of row-wide warp lane * CL + j
    // load: the CL consecutive warp sums of this lane (into pre[]); scan(): one warp scan of the lane totals turns
    // them into prefixes and the row total.  Both are independent of the coder state, so they run one token ahead,
    // under the previous token's segment loads.
    __device__ __forceinline__ void load(const uint64_t* tab, int lane) {
        nref = tab[0];
#pragma unroll
        for (int j = 0; j < CL; j++) pre[j] = tab[1 + lane * CL + j];
    }
    __device__ __forceinline__ void scan(int lane) {
        uint64_t tot = 0;
#pragma unroll
        for (int j = 0; j < CL; j++) tot += pre[j];
        const uint64_t inc = warp_incl_scan(tot, lane);
        uint64_t run = inc - tot;
#pragma unroll
        for (int j = 0; j < CL; j++) {
            const uint64_t w = pre[j];
            pre[j] = run;
            run += w;
        }
        Q = __shfl_sync(0xffffffffu, inc, 31);
    }
};

template <int VEC, int CL>
__global__ void __launch_bounds__(128, 1)
decode_serial_kernel(const __grid_constant__ RowParams rp, int V, const uint64_t* __restrict__ summ,
                     lac_dec_state* __restrict__ state, const uint8_t* __restrict__ bytes,
                     const int64_t* __restrict__ offsets, int32_t* __restrict__ syms, int64_t sym_stride, int P) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= rp.n_outer) return;
    const int Ts = tokens_of(rp, s);
    if (Ts <= 0) return;
    constexpr int words = summ_words(CL);
    const int groups = V / VEC;
    auto seg = [&](int gw) { return (int)(((int64_t)gw * groups) / (kWarps * CL)); };  // RowEngine::seg_begin
    int64_t low = state[s].low, high = state[s].high, value = state[s].value;
    uint64_t pos = state[s].pos;
    const uint8_t* data = bytes + offsets[s];
    const uint64_t nbytes = (uint64_t)(offsets[s + 1] - offsets[s]);
    uint64_t wb = pos >> 3;  // the window: stream bytes [wb, wb + 24), big-endian words
    uint64_t hi64 = load_be64(data, nbytes, wb), lo64 = load_be64(data, nbytes, wb + 8);
    uint64_t nx64 = load_be64(data, nbytes, wb + 16);
    const float* row = rp.base + s * rp.so;
    const uint64_t* tab = summ + (s * rp.T) * words;
    RowTab<CL> cur, nxt;
    cur.load(tab, lane);
    cur.scan(lane);
    lq::Scale sc = lq::make_scale(cur.Q, V);
    nxt = cur;
    if (Ts > 1) nxt.load(tab + words, lane);
    for (int t = 0; t < Ts; t++, tab += words, row += rp.st) {
        const int nref = (int)(uint32_t)cur.nref;
        const uint64_t w = (uint64_t)(high - low + 1), xr = (uint64_t)(value - low);
        const uint32_t target = lq::div_q32(xr >> 32, xr << 32, w);
        // ---- level 1: lane l looks at the row-wide warps l * CL .. l * CL + CL - 1
        int best = -1;
        uint64_t Cb = 0;
#pragma unroll
        for (int c = 0; c < CL; c++) {
            const int gw = lane * CL + c;
            const uint64_t C = cur.pre[c];
            const int gb = seg(gw), ge = seg(gw + 1);
            const bool okw = (gb < ge) & (lq::cum_of(C, (uint32_t)(gb * VEC), sc) <= target);
            best = okw ? gw : best;
            Cb = okw ? C : Cb;
        }
        const int gsel = __reduce_max_sync(0xffffffffu, best);  // >= 0: the first non-empty segment starts at cum 0
        Cb = __shfl_sync(0xffffffffu, Cb, gsel / CL);
        const int e0 = seg(gsel) * VEC + 32 * lane, eend = seg(gsel + 1) * VEC;
        // ---- level 2: q of this lane's 32 consecutive elements.  All loads first (unconditional, from addresses
        // clamped into the segment), then branch-free arithmetic: the HBM latency is paid once per token.
        const lq::Scale sc_now = sc;
        auto advance = [&]() {  // while the segment is in flight: next token's scale, then the summary after that
            cur = nxt;
            cur.scan(lane);  // (harmless on the last token: it rescans values nobody reads)
            sc = lq::make_scale(cur.Q, V);
            if (t + 2 < Ts) nxt.load(tab + 2 * words, lane);
        };
        uint32_t r[kPerThread];
        if (VEC == 4) {
            const int glast = eend - 4;  // the segment is not empty
            float4 x[kPerThread / 4];
#pragma unroll
            for (int p = 0; p < kPerThread / 4; p++)
                x[p] = __ldg(reinterpret_cast<const float4*>(row + min(e0 + 4 * p, glast)));
            advance();
#pragma unroll
            for (int p = 0; p < kPerThread / 4; p++) {
                const uint32_t m = (e0 + 4 * p < eend) ? 0xFFFFFFFFu : 0u;
                q_of2(x[p].x, x[p].y, (uint32_t)nref, r[4 * p], r[4 * p + 1]);
                q_of2(x[p].z, x[p].w, (uint32_t)nref, r[4 * p + 2], r[4 * p + 3]);
#pragma unroll
                for (int e = 0; e < 4; e++) r[4 * p + e] &= m;
            }
        } else {
            const int elast = eend - 1;
            float x[kPerThread];
#pragma unroll
            for (int j = 0; j < kPerThread; j++) x[j] = __ldg(row + min(e0 + j, elast));
            advance();
#pragma unroll
            for (int j = 0; j < kPerThread; j += 2) {
                q_of2(x[j], x[j + 1], (uint32_t)nref, r[j], r[j + 1]);
                r[j] &= (e0 + j < eend) ? 0xFFFFFFFFu : 0u;
                r[j + 1] &= (e0 + j + 1 < eend) ? 0xFFFFFFFFu : 0u;
            }
        }
        uint32_t s4[kPerThread / 4];  // sums of 4 consecutive elements (< 2^31.5)
        uint64_t L = 0;
#pragma unroll
        for (int p = 0; p < kPerThread / 4; p++) {
            s4[p] = (r[4 * p] + r[4 * p + 1]) + (r[4 * p + 2] + r[4 * p + 3]);
            L += s4[p];
        }
        const uint64_t inc = warp_incl_scan(L, lane);
        const uint64_t Cl = Cb + inc - L;
        const bool ok = (e0 < eend) & (lq::cum_of(Cl, (uint32_t)e0, sc_now) <= target);
        const unsigned ball = __ballot_sync(0xffffffffu, ok);  // lane 0 always qualifies (same test as level 1)
        const int src = 31 - __clz((int)ball);
        // ---- inside each lane (only lane `src` matters), branch-free: last group of 4 whose start qualifies,
        // then the last element of that group
        uint64_t Cp = Cl, Cg = Cl;
        int psel = 0;
#pragma unroll
        for (int p = 1; p < kPerThread / 4; p++) {
            Cp += s4[p - 1];
            const bool okp = (e0 + 4 * p < eend) & (lq::cum_of(Cp, (uint32_t)(e0 + 4 * p), sc_now) <= target);
            psel = okp ? p : psel;
            Cg = okp ? Cp : Cg;
        }
        uint32_t qe[4] = {r[0], r[1], r[2], r[3]};
#pragma unroll
        for (int p = 1; p < kPerThread / 4; p++) {
#pragma unroll
            for (int e = 0; e < 4; e++) qe[e] = (p == psel) ? r[4 * p + e] : qe[e];
        }
        const int eg = e0 + 4 * psel;
        int sym = eg;
        uint64_t Cs = Cg, Ce = Cg;
        uint32_t qsym = qe[0];
#pragma unroll
        for (int e = 1; e < 4; e++) {
            Ce += qe[e - 1];
            const bool oke = (eg + e < eend) & (lq::cum_of(Ce, (uint32_t)(eg + e), sc_now) <= target);
            sym = oke ? eg + e : sym;
            Cs = oke ? Ce : Cs;
            qsym = oke ? qe[e] : qsym;
        }
        uint32_t lo = lq::cum_of(Cs, (uint32_t)sym, sc_now);
        uint32_t hi = (sym == V - 1) ? 0u : lq::cum_of(Cs + qsym, (uint32_t)sym + 1, sc_now);
        sym = __shfl_sync(0xffffffffu, sym, src);
        lo = __shfl_sync(0xffffffffu, lo, src);
        hi = __shfl_sync(0xffffffffu, hi, src);
        // ---- A_from_bin.emit_symbol + emit_bit loop (arith_code.py:278-298), the same in every lane
        int64_t nl = low, nh = high;
        coder::ac_narrow32(nl, nh, lo, hi);
        const int64_t off = value - nl;  // the value stays inside [nl, nh]
        const int k = coder::renorm_count((uint64_t)(nh - nl + 1), P);
        coder::renorm_apply(nl, nh, P, k);
        const int o = (int)(pos - (wb << 3));  // next k bits from the window (k <= 60, o < 64)
        const uint64_t comb = o ? ((hi64 << o) | (lo64 >> (64 - o))) : hi64;
        const uint64_t nb = k ? (comb >> (64 - k)) : 0;
        low = nl;
        high = nh;
        value = nl + (off << k) + (int64_t)nb;
        pos += (uint64_t)k;
        if (o + k >= 64) {  // slide by 8 bytes; the word loaded now is not needed before the next slide
            hi64 = lo64;
            lo64 = nx64;
            wb += 8;
            nx64 = load_be64(data, nbytes, wb + 16);
        }
        if (lane == 0) syms[s * sym_stride + t] = sym;
    }
    if (lane == 0) {
        state[s].low = low;
        state[s].high = high;
        state[s].value = value;
        state[s].pos = pos;
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
static int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

// Path selection.  Returns cluster size CL (1, 2, 4, 8) in *cl and the staging path:
// 0: scalar LDG (any alignment, CL = 1), 1: 128-bit LDG (CL = 1), 2 / 4 / 8: TMA bulk ring with that many chunks
// per row and CTA (default 4 for single-CTA rows, 2 in a cluster; LAC_NO_TMA=1 and LAC_TMA_CHUNKS=n are
// measurement switches).  -1: unsupported.
static int path_for(const float* p, int V, int64_t s0, int64_t s1, int* cl) {
    const bool v4 = (V % 4 == 0) && ((((uintptr_t)p) & 15) == 0) && (s0 % 4 == 0) && (s1 % 4 == 0);
    const int cap = kThreads * kPerThread;
    *cl = V <= cap ? 1 : V <= 2 * cap ? 2 : V <= 4 * cap ? 4 : 8;
    if (V > 8 * cap) return -1;
    if (*cl > 1) return v4 ? 2 : -1;  // rows split over a cluster need 16-byte aligned rows
    if (!v4) return 0;
    static const bool no_tma = getenv("LAC_NO_TMA") != nullptr;
    if (no_tma) return 1;
    static const int nch = getenv("LAC_TMA_CHUNKS") ? atoi(getenv("LAC_TMA_CHUNKS")) : 4;
    return (nch == 2 || nch == 8) ? nch : 4;  // 4 x 32 KB: 0.7 % faster than 2 x 64 KB once the row has one barrier
}

template <int CL, typename K, typename... A>
static cudaError_t launch_tma(K kernel, int ring_bytes, int64_t units, cudaStream_t st, A... args) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes);
    if (e != cudaSuccess) return e;
    if (CL > 1) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)ring_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int clusters = sm_count() / CL;
    if (CL > 1) {  // persistent kernel: exactly as many clusters as can be co-resident
        cfg.gridDim = dim3((unsigned)(clusters * CL));
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
        if (e != cudaSuccess) return e;
        if (n < 1) return cudaErrorLaunchOutOfResources;
        if (n < clusters) clusters = n;
    }
    if (units < clusters) clusters = (int)(units > 0 ? units : 1);
    cfg.gridDim = dim3((unsigned)(clusters * CL));
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

static int plain_grid(int64_t units) {
    int sms = sm_count();
    return (int)(units < sms ? (units > 0 ? units : 1) : sms);
}

#define LAC_DISPATCH(KERNEL, UNITS, ...)                                                                              \
    switch (cl * 16 + path) {                                                                                         \
        case 16 + 2: return launch_tma<1>(KERNEL<4, true, 2, 1>, Ring<2>::kRingBytes, UNITS, st, __VA_ARGS__);        \
        case 16 + 4: return launch_tma<1>(KERNEL<4, true, 4, 1>, Ring<4>::kRingBytes, UNITS, st, __VA_ARGS__);        \
        case 16 + 8: return launch_tma<1>(KERNEL<4, true, 8, 1>, Ring<8>::kRingBytes, UNITS, st, __VA_ARGS__);        \
        case 32 + 2: return launch_tma<2>(KERNEL<4, true, 2, 2>, Ring<2>::kRingBytes, UNITS, st, __VA_ARGS__);        \
        case 64 + 2: return launch_tma<4>(KERNEL<4, true, 2, 4>, Ring<2>::kRingBytes, UNITS, st, __VA_ARGS__);        \
        case 128 + 2: return launch_tma<8>(KERNEL<4, true, 2, 8>, Ring<2>::kRingBytes, UNITS, st, __VA_ARGS__);       \
        case 16 + 1: KERNEL<4, false, 8, 1><<<plain_grid(UNITS), kThreads, 0, st>>>(__VA_ARGS__); break;              \
        case 16 + 0: KERNEL<1, false, 8, 1><<<plain_grid(UNITS), kThreads, 0, st>>>(__VA_ARGS__); break;              \
        default: return cudaErrorInvalidValue;                                                                        \
    }                                                                                                                 \
    return cudaGetLastError();

cudaError_t launch_build(const float* logits, int64_t rows, int V, int64_t row_stride, uint32_t* cum,
                         cudaStream_t st) {
    if (rows == 0) return cudaSuccess;
    const RowParams rp{logits, rows, 1, row_stride, 0, nullptr};
    int cl = 1;
    const int path = path_for(logits, V, row_stride, 0, &cl);
    LAC_DISPATCH(build_kernel, rows, rp, V, cum)
}

static cudaError_t launch_summary(const RowParams& rp, int V, int cl, int path, uint64_t* summ, cudaStream_t st) {
    const int64_t rows = rp.n_outer * rp.T;
    LAC_DISPATCH(summary_kernel, rows, rp, V, summ)
}

// Row summaries live in a stream-ordered scratch allocation (cudaMallocAsync), at most ~64 MB per chunk of rows.
static cudaError_t summ_alloc(uint64_t** summ, int64_t rows, int cl, cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);
    static bool pool_ready[64] = {};
    if (dev >= 0 && dev < 64 && !pool_ready[dev]) {  // keep freed scratch cached instead of returning it to the OS
        cudaMemPool_t mp;
        uint64_t keep = 1ull << 30;
        if (cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess)
            cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
        pool_ready[dev] = true;
    }
    return cudaMallocAsync(reinterpret_cast<void**>(summ), (size_t)(rows * summ_words(cl) * 8), st);
}
static int64_t summ_chunk_rows(int cl) {  // LAC_SUMMARY_BYTES: test switch, forces many small chunks
    static const int64_t budget = getenv("LAC_SUMMARY_BYTES") ? atoll(getenv("LAC_SUMMARY_BYTES")) : (64ll << 20);
    const int64_t rows = budget / (summ_words(cl) * 8);
    return rows < 1 ? 1 : rows;
}

// Encode side: summary pass + one warp per row for the pair.
cudaError_t launch_lookup(const float* logits, int64_t rows, int V, int64_t row_stride, const int32_t* syms,
                          uint32_t* pairs, uint32_t* status, cudaStream_t st) {
    if (rows == 0) return cudaSuccess;
    int cl = 1;
    const int path = path_for(logits, V, row_stride, 0, &cl);
    if (path < 0) return cudaErrorInvalidValue;
    const int64_t chunk = rows < summ_chunk_rows(cl) ? rows : summ_chunk_rows(cl);
    uint64_t* summ = nullptr;
    cudaError_t e = summ_alloc(&summ, chunk, cl, st);
    if (e != cudaSuccess) return e;
    for (int64_t r0 = 0; r0 < rows && e == cudaSuccess; r0 += chunk) {
        const int64_t rn = rows - r0 < chunk ? rows - r0 : chunk;
        const float* base = logits + r0 * row_stride;
        const RowParams rp{base, rn, 1, row_stride, 0, nullptr, 0};
        e = launch_summary(rp, V, cl, path, summ, st);
        if (e != cudaSuccess) break;
        const unsigned blocks = (unsigned)((rn + 7) / 8);
        uint32_t* stat = status ? status + r0 : nullptr;
#define LAC_PAIR(VEC_, CL_) \
    pair_kernel<VEC_, CL_><<<blocks, 256, 0, st>>>(base, rn, row_stride, V, summ, syms + r0, pairs + 2 * r0, stat)
        if (path == 0) LAC_PAIR(1, 1);
        else if (cl == 1) LAC_PAIR(4, 1);
        else if (cl == 2) LAC_PAIR(4, 2);
        else if (cl == 4) LAC_PAIR(4, 4);
        else LAC_PAIR(4, 8);
#undef LAC_PAIR
        e = cudaGetLastError();
    }
    const cudaError_t ef = cudaFreeAsync(summ, st);
    return e != cudaSuccess ? e : ef;
}

// Decode = summary pass + serial pass per token chunk (~240k rows of a 32000-element vocabulary per chunk).
cudaError_t launch_decode(const float* logits, int64_t n_streams, int64_t T, int64_t stream_stride,
                          int64_t tok_stride, int V, const int32_t* ntok, lac_dec_state* state,
                          const uint8_t* bytes, const int64_t* offsets, int32_t* syms, int64_t sym_stride,
                          int P, cudaStream_t st) {
    if (n_streams == 0 || T == 0) return cudaSuccess;
    int cl = 1;
    const int path = path_for(logits, V, stream_stride, tok_stride, &cl);
    if (path < 0) return cudaErrorInvalidValue;
    int64_t tc = summ_chunk_rows(cl) / n_streams;
    tc = tc < 1 ? 1 : (tc > T ? T : tc);
    uint64_t* summ = nullptr;
    cudaError_t e = summ_alloc(&summ, n_streams * tc, cl, st);
    if (e != cudaSuccess) return e;
    const unsigned serial_blocks = (unsigned)((n_streams + 3) / 4);
    for (int64_t t0 = 0; t0 < T && e == cudaSuccess; t0 += tc) {
        const int64_t tn = T - t0 < tc ? T - t0 : tc;
        const RowParams rp{logits + t0 * tok_stride, n_streams, tn, stream_stride, tok_stride, ntok, t0};
        e = launch_summary(rp, V, cl, path, summ, st);
        if (e != cudaSuccess) break;
#define LAC_SERIAL(VEC_, CL_)                                                                                    \
    decode_serial_kernel<VEC_, CL_><<<serial_blocks, 128, 0, st>>>(rp, V, summ, state, bytes, offsets, syms + t0, \
                                                                    sym_stride, P)
        if (path == 0) LAC_SERIAL(1, 1);
        else if (cl == 1) LAC_SERIAL(4, 1);
        else if (cl == 2) LAC_SERIAL(4, 2);
        else if (cl == 4) LAC_SERIAL(4, 4);
        else LAC_SERIAL(4, 8);
#undef LAC_SERIAL
        e = cudaGetLastError();
    }
    const cudaError_t ef = cudaFreeAsync(summ, st);
    return e != cudaSuccess ? e : ef;
}

cudaError_t launch_dec_init(lac_dec_state* state, int64_t n, int P, const uint8_t* bytes,
                            const int64_t* offsets, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    dec_init_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(state, n, P, bytes, offsets);
    return cudaGetLastError();
}

// Largest vocabulary the CDF kernels take; rows_need_alignment: above one CTA's capacity rows must be 16-byte aligned.
int max_vocab() { return 8 * kThreads * kPerThread; }
int max_vocab_single_cta() { return kThreads * kPerThread; }

}  // namespace lac
