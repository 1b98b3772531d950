// pack.cuh -- bit.BufferedArray (go/bit/bit.go:84-134) for one warp: 32 lanes x 32 values of B
// bits -> 32*B little-endian stream words, bit width resolved at compile time, and the
// byte-aligned write-out of those words.  Shared by the fused minp kernels and the group kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mnw {

// 32 values of B bits -> B words, all shifts resolved at compile time.  The fields do
// not overlap, so + is | and (v << sh) + o is a single LEA / IMAD.
template <int B>
__device__ __forceinline__ void pack32(const unsigned (&v)[32], unsigned (&o)[B]) {
#pragma unroll
    for (int j = 0; j < B; j++) o[j] = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const int bit = i * B, wd = bit >> 5, sh = bit & 31;
        o[wd] += v[i] << sh;
        if (sh + B > 32) o[(wd + 1) % B] = v[i] >> (32 - sh);   // (% B only keeps the index provably in range)
    }
}

// Write the group's 32*B stream words to dst (any byte alignment).  Interior words go
// out as aligned 32-bit stores, 128 bytes per instruction; the bytes of the first and last
// partial word are stored one by one, because the neighbouring groups own the rest of
// those words.  Word j of the stream sits at region[j ^ (j >> 5)].
template <int B>
__device__ __forceinline__ void write_group(uint8_t *dst, const unsigned *region, int lane) {
    const int a = (int)((uintptr_t)dst & 3);
    uint32_t *base = (uint32_t *)(dst - a) + lane;
    if (a == 0) {
#pragma unroll
        for (int m = 0; m < B; m++) base[32 * m] = region[32 * m + (lane ^ m)];
        return;
    }
    const int sh = 32 - 8 * a;
    // aligned word j (1 <= j < 32*B) = stream words j-1 and j, funnel-shifted
    unsigned prev = 0;   // stream word 32*m - 1, wanted by lane 0
#pragma unroll
    for (int m = 0; m < B; m++) {
        const unsigned hi = region[32 * m + (lane ^ m)];
        unsigned lo = __shfl_up_sync(0xffffffffu, hi, 1);
        if (lane == 0) lo = prev;
        prev = __shfl_sync(0xffffffffu, hi, 31);
        if (m > 0 || lane > 0) base[32 * m] = __funnelshift_r(lo, hi, sh);
        else {   // head: bytes a..3 of aligned word 0 = low bytes of stream word 0
            uint8_t *bp = (uint8_t *)base;
            for (int k = a; k < 4; k++) bp[k] = (uint8_t)(hi >> (8 * (k - a)));
        }
    }
    if (lane == 0) {   // tail: bytes 0..a-1 of aligned word 32*B = high bytes of the last stream word
        uint8_t *bp = (uint8_t *)(base + 32 * B);
        for (int k = 0; k < a; k++) bp[k] = (uint8_t)(prev >> (8 * (4 - a + k)));
    }
}

// The same for a group that is not full (the last 1..1023 values of a block): `nbytes` bytes of
// the stream, run-time sized.  Interior words as aligned 32-bit stores, partial first / last word
// byte by byte.
__device__ __forceinline__ void write_group_partial(uint8_t *dst, const unsigned *region, int nbytes, int lane) {
    const int a = (int)((uintptr_t)dst & 3);
    uint32_t *base = (uint32_t *)(dst - a);
    const int nwords = (a + nbytes + 3) >> 2, nsrc = (nbytes + 3) >> 2;
    for (int j = lane; j < nwords; j += 32) {
        const int jm = j - 1;
        const unsigned lo = (j > 0 && jm < nsrc) ? region[jm ^ (jm >> 5)] : 0u;
        const unsigned hi = j < nsrc ? region[j ^ (j >> 5)] : 0u;
        const unsigned w = __funnelshift_rc(lo, hi, 32 - 8 * a);
        const int t0 = 4 * j - a;
        if (t0 >= 0 && t0 + 4 <= nbytes) {
            base[j] = w;
        } else {
            uint8_t *bp = (uint8_t *)(base + j);
            for (int k = 0; k < 4; k++) {
                const int t = t0 + k;
                if (t >= 0 && t < nbytes) bp[k] = (uint8_t)(w >> (8 * k));
            }
        }
    }
}

}  // namespace mnw
