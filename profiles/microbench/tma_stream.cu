// How fast can a persistent CTA per SM stream HBM through a cp.async.bulk ring, as a function of
// bytes in flight per SM?  (read-only; in mode 1 each chunk is read by the CTA before its slot is re-armed)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma(void* dst, const void* src, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(n), "r"(smem_u32(b)) : "memory");
}
extern __shared__ __align__(128) unsigned char ring[];

// mode 0: one thread re-arms each slot as soon as it lands (pure streaming)
// mode 1: all 1024 threads read the slot (LDS.128) and __syncthreads before it is re-armed
__global__ void __launch_bounds__(1024, 1) stream(const unsigned char* src, size_t total, int slots, int slot_bytes, int mode, float* sink) {
    __shared__ uint64_t bar[32];
    size_t per_cta = total / gridDim.x / slot_bytes * slot_bytes;
    const unsigned char* base = src + (size_t)blockIdx.x * per_cta;
    int nchunks = (int)(per_cta / slot_bytes);
    if (threadIdx.x == 0) { for (int i = 0; i < slots; i++) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    if (threadIdx.x == 0) for (int i = 0; i < slots && i < nchunks; i++) { mbar_expect_tx(&bar[i], slot_bytes); tma(ring + (size_t)i * slot_bytes, base + (size_t)i * slot_bytes, slot_bytes, &bar[i]); }
    float acc = 0;
    for (int c = 0; c < nchunks; c++) {
        int s = c % slots; uint32_t ph = (c / slots) & 1;
        if (mode == 0) {
            if (threadIdx.x == 0) {
                mbar_wait(&bar[s], ph);
                int n = c + slots;
                if (n < nchunks) { mbar_expect_tx(&bar[s], slot_bytes); tma(ring + (size_t)s * slot_bytes, base + (size_t)n * slot_bytes, slot_bytes, &bar[s]); }
            }
        } else {
            mbar_wait(&bar[s], ph);
            const float4* p = reinterpret_cast<const float4*>(ring + (size_t)s * slot_bytes);
            for (int i = threadIdx.x; i < slot_bytes / 16; i += 1024) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
            __syncthreads();
            if (threadIdx.x == 0) {
                int n = c + slots;
                if (n < nchunks) { asm volatile("fence.proxy.async.shared::cta;"); mbar_expect_tx(&bar[s], slot_bytes); tma(ring + (size_t)s * slot_bytes, base + (size_t)n * slot_bytes, slot_bytes, &bar[s]); }
            }
        }
    }
    if (acc == 12345.f) sink[0] = acc;
}

// mode "row": emulate the product kernels' row loop.  Per row of `slots` chunks: wait for every chunk, read it
// to registers (LDS.128), barrier, re-arm all chunks with the next row, then spin `spin` clocks (the compute phase).
__global__ void __launch_bounds__(1024, 1) rowloop(const unsigned char* src, size_t total, int slots, int slot_bytes, int spin, float* sink) {
    __shared__ uint64_t bar[32];
    size_t row_bytes = (size_t)slots * slot_bytes;
    size_t per_cta = total / gridDim.x / row_bytes * row_bytes;
    const unsigned char* base = src + (size_t)blockIdx.x * per_cta;
    int nrows = (int)(per_cta / row_bytes);
    if (threadIdx.x == 0) { for (int i = 0; i < slots; i++) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    if (threadIdx.x == 0) for (int i = 0; i < slots; i++) { mbar_expect_tx(&bar[i], slot_bytes); tma(ring + (size_t)i * slot_bytes, base + (size_t)i * slot_bytes, slot_bytes, &bar[i]); }
    float acc = 0;
    for (int r = 0; r < nrows; r++) {
        for (int s = 0; s < slots; s++) {
            mbar_wait(&bar[s], r & 1);
            const float4* p = reinterpret_cast<const float4*>(ring + (size_t)s * slot_bytes);
            for (int i = threadIdx.x; i < slot_bytes / 16; i += 1024) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
        }
        __syncthreads();
        if (threadIdx.x == 0 && r + 1 < nrows) {
            asm volatile("fence.proxy.async.shared::cta;");
            for (int s = 0; s < slots; s++) { mbar_expect_tx(&bar[s], slot_bytes); tma(ring + (size_t)s * slot_bytes, base + (size_t)(r + 1) * row_bytes + (size_t)s * slot_bytes, slot_bytes, &bar[s]); }
        }
        long long t0 = clock64();
        while (clock64() - t0 < spin) {}
        __syncthreads();
    }
    if (acc == 12345.f) sink[0] = acc;
}

__global__ void ldg_read(const float4* src, size_t n, float* sink) {
    float acc = 0;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i + 7 * st < n; i += 8 * st) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w) : "l"(src + i + k * st));
#pragma unroll
        for (int k = 0; k < 8; k++) acc += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    if (acc == 12345.f) sink[0] = acc;
}

int main(int argc, char** argv) {
    const bool only_rowloop = argc > 1;
    size_t total = (size_t)2 << 30;
    unsigned char* src; float* sink;
    cudaMalloc(&src, total); cudaMalloc(&sink, 4); cudaMemset(src, 1, total);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(rowloop, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    int cfgs[][2] = {{4, 16384}, {8, 16384}, {12, 16384}, {13, 16384}, {4, 32768}, {6, 32768}, {2, 65536}, {3, 65536}, {16, 8192}, {24, 8192}};
    for (int mode = 0; mode < 2 && !only_rowloop; mode++)
        for (auto& c : cfgs) {
            size_t smem = (size_t)c[0] * c[1];
            float best = 1e9;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                stream<<<148, 1024, smem>>>(src, total, c[0], c[1], mode, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("mode %d slots %2d x %5d B (%3zu KB in flight/SM): %.3f ms  %.0f GB/s  %s\n", mode, c[0], c[1], smem / 1024, best, total / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    for (int slots : {2, 4})
        for (int spin : {0, 1000, 2000, 3000, 4000, 5000, 6000}) {
            int sb = 131072 / slots;
            float best = 1e9;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                rowloop<<<148, 1024, 131072>>>(src, total, slots, sb, spin, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            double rows = (double)(total / 148 / 131072);
            printf("rowloop %d x %6d B, spin %4d clk: %.3f ms  %.0f GB/s  %.2f us/row  %s\n", slots, sb, spin, best, total / best / 1e6, best * 1e3 / rows, cudaGetErrorString(cudaGetLastError()));
        }
    for (int blocks : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            ldg_read<<<blocks, 512>>>((const float4*)src, total / 16, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("ldg_read blocks %d x 512, 8 x LDG.128 in flight/thread: %.3f ms  %.0f GB/s\n", blocks, best, total / best / 1e6);
    }
    return 0;
}
