// Instruction-throughput microbenchmark for the ops the LQ32 inner loop can be built from.
// 148 CTAs x 1024 threads (8 warps / SMSP, like the product kernels), 8 independent chains per
// thread.  Prints warp-instructions per clock per SM for each op kind.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

enum Op { FFMA, FFMA2, FADD2, IADD3, SHF, LOP3, F2I, IMADW, FMNMX, SHL, MIX };

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, long long* cyc, int iters, float fs, uint32_t us) {
    float a[8]; uint32_t u[8]; uint64_t w[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 0.001f + i; u[i] = threadIdx.x * 7 + i; w[i] = ((uint64_t)u[i] << 32) | u[i]; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[i]) : "f"(fs));
            if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(w[i]) : "l"(w[(i + 1) & 7]));
            if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) & 7]));
            if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
            if (OP == SHF) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(us), "r"(u[(i + 1) & 7]));
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(us), "r"(u[(i + 1) & 7]));
            if (OP == F2I) { uint32_t tmp; asm volatile("cvt.rzi.u32.f32 %0, %1;" : "=r"(tmp) : "f"(a[i])); a[i] = __uint_as_float(tmp); }
            if (OP == IMADW) asm volatile("mad.wide.u32 %0, %1, 1, %0;" : "+l"(w[i]) : "r"(u[i]));
            if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(fs));
            if (OP == SHL) asm volatile("shl.b32 %0, %1, %2;" : "=r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(us));
            if (OP == MIX) {  // the LQ32 per-pair recipe on chain i: 8 packed + 2x(IADD, SHL, SHF, IMAD.WIDE)
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) & 7]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(w[i]) : "l"(w[(i + 1) & 7]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(us));
                asm volatile("shl.b32 %0, %0, 9;" : "+r"(u[i]));
                asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(us), "r"(u[(i + 1) & 7]));
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < 8; i++) acc += u[i] + __float_as_uint(a[i]) + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    int iters = 2000;
    k<OP><<<148, 1024>>>(out, cyc, 10, 1.0001f, 3);
    k<OP><<<148, 1024>>>(out, cyc, iters, 1.0001f, 3);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    double winst = 32.0 * iters * per_iter;  // warp-instructions per SM
    printf("%-8s %7.3f warp-inst/clk/SM  (%6.1f thread-ops/clk/SM)  cycles=%.0f %s\n", name, winst / avg, winst * 32 / avg, avg,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<FFMA>("FFMA", 8); run<FFMA2>("FFMA2", 8); run<FADD2>("FADD2", 8); run<IADD3>("IADD", 8); run<SHF>("SHF", 8);
    run<LOP3>("LOP3", 8); run<F2I>("F2I", 8); run<IMADW>("IMAD.W", 8); run<FMNMX>("FMNMX", 8); run<SHL>("SHL", 8);
    run<MIX>("MIX5", 40);
    return 0;
}
