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

// ---- Go math.Exp / math.Pow (portable pure-Go algorithms: exp.go `exp` + `expmulti`, pow.go `pow`), used by
// minh's Log columns on the read side: float32(math.Pow(10, float64(x))), go/minh/minh.go:315-319.
// The reference's own test of this step is a tolerance test (go/minh/minh_test.go:110-113); on amd64 Go's
// math.Exp is an assembly routine whose result depends on the CPU's FMA support, so a bit-exact target does not
// exist: this is the portable algorithm, evaluated without contraction.
MNW_HD double go_exp(double x) {
    const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10;
    const double Log2e = 1.44269504088896338700e+00;
    const double Overflow = 7.09782712893383973096e+02, Underflow = -7.45133219101941108420e+02;
    const double NearZero = 1.0 / (1 << 28);
    if (x != x || x > 1.7976931348623157e308) return x;
    if (x < -1.7976931348623157e308) return 0;
    if (x > Overflow) return INFINITY;
    if (x < Underflow) return 0;
    if (-NearZero < x && x < NearZero) return 1 + x;
    int k = 0;
    if (x < 0) k = (int)(Log2e * x - 0.5);
    else if (x > 0) k = (int)(Log2e * x + 0.5);
    const double hi = x - (double)k * Ln2Hi, lo = (double)k * Ln2Lo;
    const double P1 = 1.66666666666666657415e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05;
    const double P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    const double r = hi - lo, t = r * r;
    const double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    const double y = 1 - ((lo - (r * c) / (2 - c)) - hi);
    return ldexp(y, k);
}

// math.Pow(10, y): pow.go with x = 10 (none of its x-dependent special cases applies).
MNW_HD double go_pow10(double y) {
    if (y == 0) return 1;
    if (y == 1) return 10;
    if (y != y) return y;
    if (y > 1.7976931348623157e308) return INFINITY;   // |x| > 1, y = +Inf
    if (y < -1.7976931348623157e308) return 0;
    if (y == 0.5) return sqrt(10.0);
    if (y == -0.5) return 1 / sqrt(10.0);
    double yi, yf = modf(fabs(y), &yi);
    if (yi >= 9.223372036854775808e18) return y > 0 ? INFINITY : 0;
    double a1 = 1.0;
    long long ae = 0;
    if (yf != 0) {
        if (yf > 0.5) { yf--; yi++; }
        a1 = go_exp(yf * go_log(10.0));
    }
    int xe0;
    double x1 = frexp(10.0, &xe0);
    long long xe = xe0;
    for (long long i = (long long)yi; i != 0; i >>= 1) {
        if (xe < -(1 << 12) || (1 << 12) < xe) {   // overflow / underflow is certain: Ldexp decides
            ae += xe;
            break;
        }
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1;
        xe <<= 1;
        if (x1 < .5) { x1 += x1; xe--; }
    }
    if (y < 0) { a1 = 1 / a1; ae = -ae; }
    if (ae > 100000) ae = 100000;
    if (ae < -100000) ae = -100000;
    return ldexp(a1, (int)ae);
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

// The same pixel index without the IEEE divide, for dx flagged F_FASTDIV (normal,
// 2^-60 < dx < 2^60): y = t * r with r = RN(1/dx), then two Markstein corrections
// y <- RN(y + RN(t - dx*y) * r); the residual is exact in an FMA, the first
// correction makes y a faithful quotient and the second one the correctly rounded
// quotient, which is what __fdiv_rn returns.  Callers accept the result only when
// it lands in [1, pixels) -- tiny, negative, huge and NaN quotients take
// quantize_exact -- and tests/test_gpu_parity.py checks it against __fdiv_rn.
MNW_D int quantize_fast(float v, float low, float rcp, float ndx /* = -dx */) {
    float t = __fsub_rn(v, low);
    float y = __fmul_rn(t, rcp);
    float e = __fmaf_rn(ndx, y, t);
    y = __fmaf_rn(e, rcp, y);
    e = __fmaf_rn(ndx, y, t);
    y = __fmaf_rn(e, rcp, y);
    return __float2int_rd(y);
}

// float32(math.Log10(float64(x))), go/minh/minh.go:143, at a fraction of the FP64 work of go_log10.  A float32 has
// 24 mantissa bits: m = c_k (1 + r) with c_k the centre of its 1/128 interval and |r| <= 2^-8, so
// log10(x) = e log10(2) + log10(c_k) + ln(1 + r) / ln(10) with a 5-term series (absolute error < 2^-46).  go_log10's own
// absolute error is below 2^-45.  The float32 rounding of BOTH is therefore the same number unless a rounding boundary
// lies within 2^-40 of the value: then (and for zero, negative, subnormal, infinite and NaN inputs) go_log10 decides.
__device__ const double2 MNW_LOG10_TAB[128] = {
#include "log10_table.inc"
};
MNW_D float go_log10_f32(float x) {
    const unsigned u = __float_as_uint(x);
    if (u - 0x00800000u < 0x7f000000u) {   // positive and normal
        const int e = (int)(u >> 23) - 127;
        const unsigned mant = u & 0x7fffffu;
        const double m = __hiloint2double((int)(0x3ff00000u | (mant >> 3)), (int)(mant << 29));   // 1.mant
        const double2 t = __ldg(&MNW_LOG10_TAB[mant >> 16]);
        const double r = fma(m, t.x, -1.0);
        double p = fma(r, 0.2, -0.25);
        p = fma(r, p, 0.33333333333333331);
        p = fma(r, p, -0.5);
        p = fma(r, p, 1.0);
        p = p * r;   // ln(1 + r)
        const double y = fma(p, 0x1.bcb7b1526e50ep-2 /* 1 / ln 10 */, fma((double)e, 0x1.34413509f79ffp-2 /* log10 2 */, t.y));
        const float lo = __double2float_rn(y - 0x1p-40), hi = __double2float_rn(y + 0x1p-40);
        if (lo == hi) return lo;
    }
    return __double2float_rn(go_log10((double)x));
}

// float32(math.Pow(10, float64(x))), go/minh/minh.go:315-319, at a fraction of the FP64 work of go_pow10:
// 10^x = 2^k 2^f with k + f = x log2(10), |f| <= 1/2, and 2^f = exp(f ln 2) by a degree-13 Taylor polynomial (relative
// error < 2^-48 including the 2^-53-relative error of the product x log2(10), |x| < 50).  go_pow10's own relative error is
// a few units in 2^-53 (one exp, one log and at most six multiplications of the square-and-multiply loop).  The float32
// rounding of both is therefore the same number unless a rounding boundary lies within 2^-40 (relative) of the value: then,
// and outside the range where the result is a normal float32 well inside its range, go_pow10 decides.
MNW_D float go_pow10_f32(float x) {
    if (x > -37.0f && x < 38.0f) {
        const double y = (double)x * 3.3219280948873622;   // log2(10)
        const double k = rint(y), f = (y - k) * 0.69314718055994529;   // f ln 2, |.| <= 0.3466
        double p = 1.6059043836821613e-10;                 // 1/13!
        p = fma(p, f, 2.08767569878681e-09);               // 1/12!
        p = fma(p, f, 2.505210838544172e-08);
        p = fma(p, f, 2.755731922398589e-07);
        p = fma(p, f, 2.7557319223985893e-06);
        p = fma(p, f, 2.48015873015873e-05);
        p = fma(p, f, 0.0001984126984126984);
        p = fma(p, f, 0.001388888888888889);
        p = fma(p, f, 0.008333333333333333);
        p = fma(p, f, 0.041666666666666664);
        p = fma(p, f, 0.16666666666666666);
        p = fma(p, f, 0.5);
        p = fma(p, f, 1.0);
        p = fma(p, f, 1.0);
        const double v = __longlong_as_double(__double_as_longlong(p) + (long long)k * (1LL << 52));   // p 2^k (p in [0.7, 1.42])
        const float lo = __double2float_rn(v * (1.0 - 0x1p-40)), hi = __double2float_rn(v * (1.0 + 0x1p-40));
        if (lo == hi) return lo;
    }
    return __double2float_rn(go_pow10((double)x));
}

// minh processFloatGroup, go/minh/minh.go:141-149 (hi_clamp = Nextafter32(High, -Inf)).
MNW_D float minh_pre(float v, bool is_log, bool clamp, float low, float high, float hi_clamp) {
    if (is_log) v = go_log10_f32(v);
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

// Rotation constant of the periodic-arc statistic.  With half = P/2 and
// K = P - half - 1, w(q) = (q - q0 + K) mod P orders pixel indices by their
// signed periodic distance to q0 (go/group.go:412-420): dist = w - K.
MNW_HD unsigned long long arc_rotation(long long q0, long long P) {
    long long K = P - P / 2 - 1;
    long long c = K - q0;
    if (c < 0) c += P;
    return (unsigned long long)c;
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
// key: once per block; hash: one lowbias32 round over a Weyl sequence in i.
MNW_HD uint32_t jitter_key(unsigned long long seed, unsigned long long block) {
    uint32_t k = mix32((uint32_t)(seed >> 32));
    k = mix32((uint32_t)seed ^ k);
    k = mix32((uint32_t)(block >> 32) ^ k);
    return mix32((uint32_t)block ^ k);
}
// two multiply-xorshift rounds over a Weyl sequence in i; the jitter is the TOP 24 bits,
// which the final xorshift of mix32 would not touch, so it is left out
MNW_HD uint32_t jitter_hash_keyed(uint32_t key, uint32_t i) {
    uint32_t x = i * 0x9E3779B1U + key;
    x ^= x >> 16; x *= 0x7feb352dU;
    x ^= x >> 15; x *= 0x846ca68bU;
    return x;
}
MNW_HD uint32_t jitter_hash32(unsigned long long seed, unsigned long long block, unsigned long long i) {
    return jitter_hash_keyed(jitter_key(seed, block), (uint32_t)i);
}

}  // namespace mnw
