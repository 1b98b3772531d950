// pipebench.cu -- issue-rate microbenchmark of the instructions the encode kernels lean on
// (B200 / sm_100a): warp-instructions per cycle per SM sub-partition for FFMA, FFMA2, FADD2.RM,
// VIMNMX3, VIADDMNMX, PRMT, IADD3, IMAD, VIADD.16x2, SHF, and an FP2 + INT mix.
//   nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -o pipebench tools/pipebench.cu && ./pipebench
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__global__ void k(unsigned *out, unsigned a0, unsigned a1) {
    unsigned x[CHAINS], y[CHAINS];
    unsigned long long p[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { x[c] = threadIdx.x + c + a0; y[c] = threadIdx.x * 3 + c; p[c] = ((unsigned long long)x[c] << 32) | y[c]; }
    const unsigned long long q = ((unsigned long long)a1 << 32) | a0;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a0), "r"(a1));
            if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[c]) : "l"(q));
            if (OP == 2) asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q));
            if (OP == 3) asm volatile("min.u32 %0, %0, %1; min.u32 %0, %0, %2;" : "+r"(x[c]) : "r"(y[c]), "r"(a1));   // fuses to VIMNMX3?
            if (OP == 4) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; min.u32 %0, %0, t; }" : "+r"(x[c]) : "r"(a1));   // VIADDMNMX
            if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(x[c]) : "r"(y[c]));
            if (OP == 6) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(a1));
            if (OP == 7) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a0), "r"(a1));
            if (OP == 8) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[c]) : "r"(y[c]));
            if (OP == 9) {   // mix: one FFMA2 + one integer min per chain
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[c]) : "l"(q));
                asm volatile("min.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(a1));
            }
            if (OP == 10) {   // mix: one FFMA + one integer min per chain
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(y[c]) : "r"(a0), "r"(a1));
                asm volatile("min.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(a1));
            }
            if (OP == 11) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(q));
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += x[c] + y[c] + (unsigned)p[c] + (unsigned)(p[c] >> 32);
    if (s == 0x12345678u) out[0] = s;
}

template <int OP>
void run(const char *name, int per_iter, unsigned *d, int warps_per_smsp) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int threads = 128 * warps_per_smsp;
    k<OP><<<sms, threads>>>(d, 1, 3);
    cudaEventRecord(a);
    k<OP><<<sms, threads>>>(d, 1, 3);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double cycles = ms * 1e-3 * khz * 1e3;
    const double winstr_per_smsp = (double)ITERS * CHAINS * per_iter * warps_per_smsp;
    printf("%-28s warps/SMSP %d: %.3f warp-instr/cycle/SMSP (%.2f cycles per instr)\n", name, warps_per_smsp,
           winstr_per_smsp / cycles, cycles / winstr_per_smsp);
}

int main() {
    unsigned *d;
    cudaMalloc(&d, 4);
    for (int w : {1, 4}) {
        run<0>("FFMA", 1, d, w);
        run<1>("FFMA2", 1, d, w);
        run<11>("FMUL2", 1, d, w);
        run<2>("FADD2.RM", 1, d, w);
        run<3>("min,min (VIMNMX3?)", 1, d, w);
        run<4>("add,min (VIADDMNMX?)", 1, d, w);
        run<5>("PRMT", 1, d, w);
        run<6>("IADD", 1, d, w);
        run<7>("IMAD", 1, d, w);
        run<8>("SHF", 1, d, w);
        run<9>("FFMA2 + min", 2, d, w);
        run<10>("FFMA + min", 2, d, w);
    }
    return 0;
}
