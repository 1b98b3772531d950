// f32x2.cuh -- the quantiser of go/group.go:319 on float PAIRS (Blackwell's packed FADD2 / FMUL2 / FFMA2): half the
// issue slots of the scalar form.  Shared by k_pipe_vec3 (pipe_vec3.cuh) and the group kernels (kernels_group.cu).
#pragma once
#include <cuda_runtime.h>

namespace mnw {
namespace {

__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void f2_bits(unsigned long long v, unsigned &a, unsigned &b) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long f2_add_rm(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// go/group.go:319 for two floats at once: raw bits of RM(y + 2^23), y the correctly rounded
// float32 quotient (x - low) / dx by two Markstein corrections (device_math.cuh quantize_fast).
// For 0 <= y < 2^23 the bits are FMAGIC + floor(y).
__device__ __forceinline__ unsigned long long quantize2(unsigned long long v, unsigned long long low,
                                                        unsigned long long rcp, unsigned long long ndx) {
    const unsigned long long t = f2_sub(v, low);
    unsigned long long y = f2_mul(t, rcp);
    unsigned long long e = f2_fma(ndx, y, t);
    y = f2_fma(e, rcp, y);
    e = f2_fma(ndx, y, t);
    y = f2_fma(e, rcp, y);
    return f2_add_rm(y, 0x4B0000004B000000ULL);
}

}  // namespace
}  // namespace mnw
