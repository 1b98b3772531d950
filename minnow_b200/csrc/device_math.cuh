// device_math.cuh -- scalar arithmetic of the minnow hot path, written so that
// host (C++) and device (sm_100a, --fmad=false) evaluate the SAME IEEE
// operations in the SAME order as the Go reference on amd64.
//
// Everything here is the product's own implementation; the CPU oracle under
// oracle/ is a separate restatement and is never included from here.
#pragma once
#include <cstdint>
#include <cmath>
#include <climits>

#if defined(__CUDACC__)
#define MNW_HD __host__ __device__ __forceinline__
#define MNW_D __device__ __forceinline__
#else
#define MNW_HD inline
#define MNW_D inline
#endif

namespace mnw {

// ---- Go math.Log / Log2 / Log10 (pure-Go algorithm, FreeBSD e_log.c lineage).
// Used by bit.PrecisionNeeded (go/bit/bit.go:19-21) and by minh's Log columns
// (go/minh/minh.go:143).  No FMA contraction: the .cu files are compiled with
// --fmad=false and host code with -ffp-contract=off.
MNW_HD double go_log(double x) {
    const double Ln2Hi = 6.93147180369123816490e-01;
    const double Ln2Lo = 1.90821492927058770002e-10;
    const double L1 = 6.666666666666735130e-01;
    const double L2 = 3.999999999940941908e-01;
    const double L3 = 2.857142874366239149e-01;
    const double L4 = 2.222219843214978396e-01;
    const double L5 = 1.818357216161805012e-01;
    const double L6 = 1.531383769920937332e-01;
    const double L7 = 1.479819860511658591e-01;
    const double Sqrt2Over2 = 0.70710678118654757;  // 0x3fe6a09e667f3bcd

    if (x != x) return x;
    if (x > 1.7976931348623157e308) return x;  // +Inf
    if (x < 0) return NAN;
    if (x == 0) return -INFINITY;

    int ki;
    double f1 = frexp(x, &ki);
    if (f1 < Sqrt2Over2) {
        f1 *= 2;
        ki--;
    }
    double f = f1 - 1;
    double k = (double)ki;

    double s = f / (2 + f);
    double s2 = s * s;
    double s4 = s2 * s2;
    double t1 = s2 * (L1 + s4 * (L3 + s4 * (L5 + s4 * L7)));
    double t2 = s4 * (L2 + s4 * (L4 + s4 * L6));
    double R = t1 + t2;
    double hfsq = 0.5 * f * f;
    return k * Ln2Hi - ((hfsq - (s * (hfsq + R) + k * Ln2Lo)) - f);
}

MNW_HD double go_log2(double x) {
    int e;
    double frac = frexp(x, &e);
    if (frac == 0.5) return (double)(e - 1);
    return go_log(frac) * 1.4426950408889634 /* 1/Ln2 = 0x3ff71547652b82fe */ + (double)e;
}

MNW_HD double go_log10(double x) {
    return go_log2(x) * 0.3010299956639812 /* Ln2/Ln10 = 0x3fd34413509f79ff */;
}

// bit.PrecisionNeeded, go/bit/bit.go:19-21: int(ceil(log2(float64(max+1)))).
// Returns -1 for max = 2^64-1 (Go: log2(0) = -Inf, conversion undefined).
MNW_HD int precision_needed(unsigned long long max) {
    unsigned long long v = max + 1ULL;
    if (v == 0ULL) return -1;
#if defined(__CUDA_ARCH__)
    double d = __ull2double_rn(v);
#else
    double d = (double)v;
#endif
    return (int)ceil(go_log2(d));
}

// bit.ArrayBytes, go/bit/bit.go:23-25 (exact while bits*n < 2^53).
MNW_HD long long array_bytes(long long bits, long long n) {
    return (bits * n + 7) >> 3;
}

// Go's float64 -> int64 conversion on amd64 (CVTTSD2SQ): NaN / out of range
// give 0x8000000000000000.  v is already integral (a floor).
MNW_HD long long go_float_to_i64(float fl) {
    if (!(fl >= -9223372036854775808.0f && fl < 9223372036854775808.0f)) return LLONG_MIN;
    return (long long)fl;
}

#if defined(__CUDACC__)
// go/group.go:319: int64(math.Floor(float64((x - low) / dx))) -- float32
// subtract and IEEE float32 divide; the floor of a float32 is the same number
// in float64, so it is taken in float32.
MNW_D long long quantize_exact(float v, float low, float dx) {
    float t = __fsub_rn(v, low);
    float r = __fdiv_rn(t, dx);
    return go_float_to_i64(floorf(r));
}

// minh processFloatGroup, go/minh/minh.go:141-149 (hi_clamp = Nextafter32(High, -Inf)).
MNW_D float minh_pre(float v, bool is_log, bool clamp, float low, float high, float hi_clamp) {
    if (is_log) v = __double2float_rn(go_log10((double)v));
    if (clamp) {
        if (v < low) v = low;
        if (v >= high) v = hi_clamp;
    }
    return v;
}
#endif

// go/group.go:412-420 periodicDistance (wrapping arithmetic like Go's int64).
MNW_HD long long periodic_distance(long long x, long long x0, long long pixels) {
    long long d = (long long)((unsigned long long)x - (unsigned long long)x0);
    if (d >= 0) {
        if (d > (long long)((unsigned long long)pixels - (unsigned long long)d))
            return (long long)((unsigned long long)d - (unsigned long long)pixels);
    } else {
        if (d < (long long)(0ULL - ((unsigned long long)d + (unsigned long long)pixels)))
            return (long long)((unsigned long long)pixels + (unsigned long long)d);
    }
    return d;
}

// go/group.go:374-382 bound, one element.
MNW_HD long long bound1(long long x, long long mn, long long pixels) {
    if (x < mn) return (long long)((unsigned long long)x + (unsigned long long)pixels);
    if (x >= (long long)((unsigned long long)mn + (unsigned long long)pixels))
        return (long long)((unsigned long long)x - (unsigned long long)pixels);
    return x;
}

// Decode jitter hash (include/minnow_cuda.h, mnw_jitter).
MNW_HD uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU;
    x ^= x >> 15; x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}
MNW_HD uint32_t jitter_hash32(unsigned long long seed, unsigned long long block, unsigned long long i) {
    uint32_t x = mix32((uint32_t)i ^ (uint32_t)seed);
    x += (uint32_t)block * 0x9E3779B9U + (uint32_t)(seed >> 32);
    return mix32(x);
}

}  // namespace mnw
