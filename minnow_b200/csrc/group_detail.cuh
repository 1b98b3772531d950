// group_detail.cuh -- device helpers shared by the group kernels (kernels_generic.cu: the generic two-pass
// path; kernels_group.cu: the fused single-read group encoder).  Anonymous namespace: one copy per translation unit.
#pragma once
#include "engine.cuh"
#include "device_math.cuh"
#include "launch.cuh"
#include "pack.cuh"

namespace mnw {
namespace {

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ int64_t find_block(const BlockDesc *descs, const BatchShape &sh,
                                              int64_t unit, int64_t units_per_block, bool tiles) {
    if (sh.uniform_n > 0) return unit / units_per_block;
    int64_t lo = 0, hi = sh.nblocks;  // last b with start(b) <= unit
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        int64_t s = tiles ? descs[mid].tile0 : descs[mid].chunk0;
        if (s <= unit) lo = mid; else hi = mid;
    }
    return lo;
}

// The integer the reference would hold in its int64 scratch for element i of
// the block, BEFORE periodicMin/bound: the int64 itself, or the pixel index.
__device__ __forceinline__ long long block_value(const BlockDesc &d, int64_t i) {
    if (d.kind == KIND_I64) {
        const long long *p = (const long long *)d.src;
        return d.access == ACC_GATHER ? p[d.idx[i]] : p[i];
    }
    const float *p = (const float *)d.src;
    float v;
    if (d.access == ACC_CONTIG) {
        v = p[i];
    } else if (d.access == ACC_GATHER) {
        v = p[d.idx[i]];
    } else {  // getSubCell index arithmetic, go/minp/minp.go:252-262
        uint32_t ii = (uint32_t)i, ns = (uint32_t)d.nsub;
        uint32_t jx = ii % ns, t = ii / ns;
        uint32_t jy = t % ns, jz = t / ns;
        int64_t idx = (int64_t)(jx + d.ix0) + (int64_t)(jy + d.iy0) * d.nfile +
                      (int64_t)(jz + d.iz0) * d.nfile * d.nfile;
        v = p[3 * idx + d.axis];
    }
    v = minh_pre(v, d.flags & F_LOG10, d.flags & F_CLAMP, d.low, d.high, d.hi_clamp);
    return quantize_exact(v, d.low, d.dx);
}

__device__ __forceinline__ long long warp_min_ll(long long v) {
    for (int o = 16; o; o >>= 1) { long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    return v;
}
__device__ __forceinline__ long long warp_max_ll(long long v) {
    for (int o = 16; o; o >>= 1) { long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}
__device__ __forceinline__ unsigned long long warp_min_ull(unsigned long long v) {
    for (int o = 16; o; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    return v;
}
__device__ __forceinline__ unsigned long long warp_max_ull(unsigned long long v) {
    for (int o = 16; o; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}

// bits / nbytes from (min, max offset); flags an error where Go is undefined.
__device__ __forceinline__ void finish_stat(BlockStat &s, int64_t n, unsigned long long maxoff, int *err) {
    int bits = precision_needed(maxoff);
    if (bits < 0) { bits = 64; atomicExch(err, 1); }
    s.bits = bits;
    s.nbytes = array_bytes(bits, n);
}

// ---------------------------------------------------------------------------
// byte-aligned stream store: write nbytes of the little-endian word stream s[]
// to dst, which may have any byte alignment.  Interior words are written as
// aligned 32-bit stores; the first/last partial word as single bytes, because
// neighbouring blocks own the other bytes of those words.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void store_stream(uint8_t *dst, const uint32_t *s, int64_t nbytes, int tid, int nthreads) {
    const uintptr_t A = (uintptr_t)dst;
    const int a = (int)(A & 3);
    uint32_t *base = (uint32_t *)(A - a);
    const int64_t nwords = (a + nbytes + 3) >> 2;
    for (int64_t j = tid; j < nwords; j += nthreads) {
        uint32_t lo = j > 0 ? s[j - 1] : 0u;
        uint32_t hi = 4 * j < nbytes ? s[j] : 0u;
        uint32_t w = __funnelshift_rc(lo, hi, 32 - 8 * a);
        int64_t t0 = 4 * j - a;  // stream index of this word's byte 0
        if (t0 >= 0 && t0 + 4 <= nbytes) {
            base[j] = w;
        } else {
            uint8_t *bp = (uint8_t *)(base + j);
            for (int k = 0; k < 4; k++) {
                int64_t t = t0 + k;
                if (t >= 0 && t < nbytes) bp[k] = (uint8_t)(w >> (8 * k));
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Fast path for contiguous float32 blocks of periodic groups with pixels < 2^31 (every
// FloatGroup the reference's Writer can create, go/writer.go:72-75): the same two passes as
// k_stats / k_pack, but with 128-bit loads, the division-free quantiser (device_math.cuh
// quantize_fast, exact IEEE redo for any element it does not vouch for), 32-bit statistics,
// and the warp-level compile-time packer of pack.cuh.
// ---------------------------------------------------------------------------
struct QuantP {
    float low, high, dx, hi_clamp, rcp, ndx;
    unsigned P, tmax;
    int flags;
    bool fast_ok;
};
__device__ __forceinline__ QuantP quant_params(const BlockDesc &d) {
    QuantP p;
    p.low = d.low; p.high = d.high; p.dx = d.dx; p.hi_clamp = d.hi_clamp;
    p.rcp = __frcp_rn(d.dx); p.ndx = -d.dx;
    p.P = (unsigned)d.pixels;
    p.tmax = __float_as_uint(__fsub_rn(d.high, d.low));
    p.flags = d.flags;
    p.fast_ok = d.dx > 0x1p-60f && d.dx < 0x1p60f && d.pixels >= 2;
    return p;
}
// pixel index of one element, folded into [0, pixels) (pixels -> 0, see k_stats); *raw gets the
// reference's unfolded int64 when the element is out of range (oob)
__device__ __forceinline__ unsigned quant_elem(float v, const QuantP &p, bool &oob, long long *raw) {
    if (p.flags & (F_LOG10 | F_CLAMP)) v = minh_pre(v, p.flags & F_LOG10, p.flags & F_CLAMP, p.low, p.high, p.hi_clamp);
    const float tt = __fsub_rn(v, p.low);
    float y = __fmul_rn(tt, p.rcp);
    float e = __fmaf_rn(p.ndx, y, tt);
    y = __fmaf_rn(e, p.rcp, y);
    e = __fmaf_rn(p.ndx, y, tt);
    y = __fmaf_rn(e, p.rcp, y);
    unsigned q = (unsigned)__float2int_rd(y);
    if (!(p.fast_ok && __float_as_uint(tt) <= p.tmax && q < p.P)) {   // rare: the IEEE divide decides
        const long long qq = quantize_exact(v, p.low, p.dx);
        if (raw) *raw = qq;
        if (qq == (long long)p.P) q = 0;
        else if ((unsigned long long)qq < (unsigned long long)p.P) q = (unsigned)qq;
        else { oob = true; q = 0; }
    } else if (raw) {
        *raw = q;
    }
    return q;
}

// The unchecked form of the same quantiser for whole float4s: raw bits of RM(y + 2^23), y the correctly rounded
// quotient (x - low) / dx; for 0 <= y < 2^23 they are FQ_MAGIC + floor(y).  The caller vouches for the result by a
// range test on the bits ([FQ_MAGIC, FQ_MAGIC + pixels) rejects negative, NaN, infinite and too large quotients) and
// falls back to quant_elem otherwise.  Needs pixels <= 2^22 and no log10 pre-transform; the clamp is applied here.
constexpr unsigned FQ_MAGIC = 0x4B000000u;
__device__ __forceinline__ unsigned quant_bits(float v, const QuantP &p, bool clamp) {
    if (clamp) {   // go/minh/minh.go:144-147
        v = v < p.low ? p.low : v;
        v = v >= p.high ? p.hi_clamp : v;
    }
    const float tt = __fsub_rn(v, p.low);
    float y = __fmul_rn(tt, p.rcp);
    float e = __fmaf_rn(p.ndx, y, tt);
    y = __fmaf_rn(e, p.rcp, y);
    e = __fmaf_rn(p.ndx, y, tt);
    y = __fmaf_rn(e, p.rcp, y);
    return __float_as_uint(__fadd_rd(y, 8388608.0f));
}
__device__ __forceinline__ bool quant_bits_ok(const QuantP &p) {
    return p.fast_ok && !(p.flags & F_LOG10) && p.P <= (1u << 22);
}


// Per block: the closed form of periodicMin, min, bits, nbytes from the accumulated statistics.
// Returns true when the block needs the exact sequential periodicMin (slow_block) instead.
__device__ __forceinline__ bool finalize_block(const BlockDesc &d, BlockStat &s, int *err) {
    s.do_bound = 0; s.slow = 0; s.pmin = 0; s.out_off = 0;
    if (d.n == 0) {  // int64Min / Bits / periodicMin of an empty slice are all 0
        s.min = 0; s.bits = 0; s.nbytes = 0;
    } else if (d.kind == KIND_I64 || !(d.flags & F_PERIODIC)) {
        s.min = s.qmin;
        finish_stat(s, d.n, (unsigned long long)s.qmax - (unsigned long long)s.qmin, err);
    } else if (s.oob & 1u) {   // (bit 1 of oob only says that some thread took the checked quantiser: k_group_fused)
        s.slow = 1;
        return true;
    } else {
        // Order-independent form of periodicMin (go/group.go:384-409), valid for
        // pixel indices in [0, pixels): the arc is [q0 + dmin, q0 + dmax] with
        // d = signed periodic distance to q0; too wide an arc returns 0.
        const long long P = d.pixels, half = P / 2, K = P - half - 1;
        unsigned long long spread = s.wmax - s.wmin + 1ULL;
        s.do_bound = 1;
        if (spread > (unsigned long long)half) {
            s.pmin = 0;
            s.min = s.qmin;
            finish_stat(s, d.n, (unsigned long long)s.qmax - (unsigned long long)s.qmin, err);
        } else {
            long long m = s.q0 + ((long long)s.wmin - K);
            if (m < 0) m += P;
            s.pmin = m;
            s.min = m;
            finish_stat(s, d.n, spread - 1ULL, err);
        }
    }
    return false;
}

// Exact periodicMin for a block with out-of-range pixel indices, by the whole CTA (any number of warps up to 8).
// Warp 0 walks the block in the reference's order; lanes whose element lies strictly inside the current arc
// (`continue` at go/group.go:395) are skipped 32 at a time with a ballot, every other element updates the arc
// exactly as the Go loop does.  Then min/max of bound() over the block.  s_pmin / s_red: shared scratch.
__device__ __forceinline__ void slow_block(const BlockDesc &d, BlockStat *sb, int *err, long long *s_pmin,
                                           long long (*s_red)[8]) {
    const long long P = d.pixels;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        long long x0 = block_value(d, 0), width = 1;
        bool returned_zero = false;
        for (int64_t base = 0; base < d.n && !returned_zero; base += 32) {
            int64_t i = base + lane;
            bool valid = i < d.n;
            long long q = valid ? block_value(d, i) : 0;
            unsigned pending = __ballot_sync(0xffffffffu, valid);
            while (pending) {
                long long x1 = (long long)((unsigned long long)x0 + (unsigned long long)width - 1ULL);
                if (x1 >= P) x1 = (long long)((unsigned long long)x1 - (unsigned long long)P);
                long long d0 = periodic_distance(q, x0, P);
                long long d1 = periodic_distance(q, x1, P);
                bool inside = d0 > 0 && d1 < 0;
                unsigned act = __ballot_sync(0xffffffffu, !inside) & pending;
                if (!act) break;
                int j = __ffs(act) - 1;
                long long e0 = __shfl_sync(0xffffffffu, d0, j);
                long long e1 = __shfl_sync(0xffffffffu, d1, j);
                if (e1 > (long long)(0ULL - (unsigned long long)e0)) {
                    width = (long long)((unsigned long long)width + (unsigned long long)e1);
                } else {
                    x0 = (long long)((unsigned long long)x0 + (unsigned long long)e0);
                    if (x0 < 0) x0 = (long long)((unsigned long long)x0 + (unsigned long long)P);
                    width = (long long)((unsigned long long)width - (unsigned long long)e0);
                }
                if (width > P / 2) { returned_zero = true; break; }
                pending &= ~((2u << j) - 1u);  // elements up to j are done
            }
        }
        if (lane == 0) *s_pmin = returned_zero ? 0 : x0;
    }
    __syncthreads();
    const long long pmin = *s_pmin;
    long long mn = LLONG_MAX, mx = LLONG_MIN;
    for (int64_t i = threadIdx.x; i < d.n; i += blockDim.x) {
        long long q = bound1(block_value(d, i), pmin, P);
        mn = q < mn ? q : mn;
        mx = q > mx ? q : mx;
    }
    mn = warp_min_ll(mn); mx = warp_max_ll(mx);
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = mn; s_red[1][threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); k++) {
            mn = s_red[0][k] < mn ? s_red[0][k] : mn;
            mx = s_red[1][k] > mx ? s_red[1][k] : mx;
        }
        BlockStat s = *sb;
        s.pmin = pmin; s.do_bound = 1; s.min = mn;
        finish_stat(s, d.n, (unsigned long long)mx - (unsigned long long)mn, err);
        *sb = s;
    }
    __syncthreads();
}

// slow_block for ONE warp (all 32 lanes of the calling warp, no CTA-level barrier): the exact sequential periodicMin and
// the min / max of bound() of a block with out-of-range pixel indices.  Values only (uniform across the warp).
__device__ __forceinline__ void slow_block_warp_values(const BlockDesc &d, long long &pmin_out, long long &mn_out, long long &mx_out) {
    const long long P = d.pixels;
    const int lane = threadIdx.x & 31;
    long long x0 = block_value(d, 0), width = 1;
    bool returned_zero = false;
    for (int64_t base0 = 0; base0 < d.n && !returned_zero; base0 += 32 * 8) {
      long long qq[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {   // 8 independent loads per lane in flight; the arc walk below is sequential
          const int64_t i = base0 + 32 * u + lane;
          qq[u] = i < d.n ? block_value(d, i) : 0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        if (returned_zero) break;
        const int64_t base = base0 + 32 * u;
        if (base >= d.n) break;
        int64_t i = base + lane;
        bool valid = i < d.n;
        long long q = qq[u];
        unsigned pending = __ballot_sync(0xffffffffu, valid);
        while (pending) {
            long long x1 = (long long)((unsigned long long)x0 + (unsigned long long)width - 1ULL);
            if (x1 >= P) x1 = (long long)((unsigned long long)x1 - (unsigned long long)P);
            long long d0 = periodic_distance(q, x0, P);
            long long d1 = periodic_distance(q, x1, P);
            bool inside = d0 > 0 && d1 < 0;
            unsigned act = __ballot_sync(0xffffffffu, !inside) & pending;
            if (!act) break;
            int j = __ffs(act) - 1;
            long long e0 = __shfl_sync(0xffffffffu, d0, j);
            long long e1 = __shfl_sync(0xffffffffu, d1, j);
            if (e1 > (long long)(0ULL - (unsigned long long)e0)) {
                width = (long long)((unsigned long long)width + (unsigned long long)e1);
            } else {
                x0 = (long long)((unsigned long long)x0 + (unsigned long long)e0);
                if (x0 < 0) x0 = (long long)((unsigned long long)x0 + (unsigned long long)P);
                width = (long long)((unsigned long long)width - (unsigned long long)e0);
            }
            if (width > P / 2) { returned_zero = true; break; }
            pending &= ~((2u << j) - 1u);  // elements up to j are done
        }
      }
    }
    const long long pmin = returned_zero ? 0 : x0;   // uniform across the warp
    long long mn = LLONG_MAX, mx = LLONG_MIN;
    for (int64_t base = 0; base < d.n; base += 32 * 16) {   // 16 independent loads per lane in flight
        long long q[16];
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const int64_t i = base + 32 * u + lane;
            q[u] = i < d.n ? block_value(d, i) : LLONG_MIN;
        }
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (base + 32 * u + lane < d.n) {
                const long long qb = bound1(q[u], pmin, P);
                mn = qb < mn ? qb : mn;
                mx = qb > mx ? qb : mx;
            }
        }
    }
    pmin_out = pmin; mn_out = warp_min_ll(mn); mx_out = warp_max_ll(mx);
}

__device__ __forceinline__ void slow_block_warp(const BlockDesc &d, BlockStat *sb, int *err) {
    long long pmin, mn, mx;
    slow_block_warp_values(d, pmin, mn, mx);
    if ((threadIdx.x & 31) == 0) {
        BlockStat s = *sb;
        s.pmin = pmin; s.do_bound = 1; s.min = mn;
        finish_stat(s, d.n, (unsigned long long)mx - (unsigned long long)mn, err);
        *sb = s;
    }
    __syncwarp();
}

// One 4096-element tile of one block through the 64-bit capable packer (any width, any access pattern), by a CTA of
// PACK_THREADS threads: bound, subtract min, LSB-first packing (go/bit/bit.go:84-134) to the byte-aligned
// destination.  s_out: PACK_THREADS * 64 + 4 words of shared scratch.  Ends with a __syncthreads().
__device__ __forceinline__ void pack_tile_generic(const BlockDesc &d, const BlockStat &st, int64_t tile_in_block,
                                                  uint8_t *chain_out, uint32_t *s_out) {
    const int bits = st.bits;
    const int64_t first = tile_in_block * PACK_TILE;
    const int64_t count = first + PACK_TILE <= d.n ? PACK_TILE : d.n - first;
    const unsigned long long mask = bits >= 64 ? ~0ULL : ((1ULL << bits) - 1ULL);  // go/bit/bit.go:104
    const long long P = d.pixels;

    const int g = threadIdx.x;
    int64_t i0 = first + 32 * (int64_t)g;
    if (32 * g < count) {
        unsigned long long acc_lo = 0, acc_hi = 0;
        int pos = 0, w = g * bits;
#pragma unroll 4
        for (int k = 0; k < 32; k++) {
            int64_t i = i0 + k;
            unsigned long long v = 0;
            if (i < d.n) {
                long long q = block_value(d, i);
                if (st.do_bound) q = bound1(q, st.pmin, P);
                v = ((unsigned long long)q - (unsigned long long)st.min) & mask;
            }
            acc_lo |= v << pos;
            if (pos) acc_hi |= v >> (64 - pos);
            pos += bits;
            while (pos >= 32) {
                s_out[w++] = (uint32_t)acc_lo;
                acc_lo = (acc_lo >> 32) | (acc_hi << 32);
                acc_hi >>= 32;
                pos -= 32;
            }
        }
    }
    __syncthreads();
    const int64_t nbytes = (count * bits + 7) >> 3;
    uint8_t *dst = chain_out + st.out_off + ((first * bits) >> 3);
    store_stream(dst, s_out, nbytes, threadIdx.x, PACK_THREADS);
    __syncthreads();   // s_out is reused by the next tile
}

constexpr int FPACK_THREADS = 128;   // 4 warps = the 4 pack groups of a 4096-element tile

// The tile's 32-bit values staged in sv (swizzled by 16-byte chunk) -> packed bytes: every warp packs one group of
// 1024 elements with the compile-time packer of pack.cuh and writes it to its byte-aligned place.  Noinline: the
// 32-way switch is instantiated once per translation unit.  Ends with a __syncthreads().
__device__ __noinline__ void pack_staged_tile(unsigned *sv, int count, int bits, int64_t first, const BlockStat &st,
                                              uint8_t *chain_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = warp;                              // group of 1024 elements
    if (g * 1024 < count) {
        unsigned v[32];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const uint4 r = *(const uint4 *)&sv[g * 1024 + lane * 32 + ((c ^ (lane & 7)) << 2)];
            v[4 * c] = r.x; v[4 * c + 1] = r.y; v[4 * c + 2] = r.z; v[4 * c + 3] = r.w;
        }
        unsigned *region = sv + g * 1024;
        const int gcount = count - g * 1024 < 1024 ? count - g * 1024 : 1024;
        uint8_t *dst = chain_out + st.out_off + (((first + g * 1024) * bits) >> 3);
        switch (bits) {
#define MNW_CASE(B)                                                                         \
    case B: {                                                                               \
        unsigned o[B];                                                                      \
        pack32<B>(v, o);                                                                    \
        __syncwarp();                                                                       \
        _Pragma("unroll") for (int j = 0; j < B; j++) {                                     \
            const int W = lane * B + j;                                                     \
            region[W ^ (W >> 5)] = o[j];                                                    \
        }                                                                                   \
        __syncwarp();                                                                       \
        if (gcount == 1024) write_group<B>(dst, region, lane);                              \
        else write_group_partial(dst, region, (gcount * B + 7) >> 3, lane);                 \
    } break;
            MNW_CASE(1) MNW_CASE(2) MNW_CASE(3) MNW_CASE(4) MNW_CASE(5) MNW_CASE(6) MNW_CASE(7) MNW_CASE(8)
            MNW_CASE(9) MNW_CASE(10) MNW_CASE(11) MNW_CASE(12) MNW_CASE(13) MNW_CASE(14) MNW_CASE(15) MNW_CASE(16)
            MNW_CASE(17) MNW_CASE(18) MNW_CASE(19) MNW_CASE(20) MNW_CASE(21) MNW_CASE(22) MNW_CASE(23) MNW_CASE(24)
            MNW_CASE(25) MNW_CASE(26) MNW_CASE(27) MNW_CASE(28) MNW_CASE(29) MNW_CASE(30) MNW_CASE(31) MNW_CASE(32)
#undef MNW_CASE
            default: break;
        }
    }
    __syncthreads();
}

// One 4096-element tile of a contiguous float32 block of a periodic group (pixels < 2^31): quantise, bound,
// subtract min, stage, pack (floatGroup.writeData, go/group.go:312-327).  bits in 1..32.
__device__ __forceinline__ void pack_tile_f32(const BlockDesc &d, const BlockStat &st, int64_t tile_in_block,
                                              uint8_t *chain_out, unsigned *sv) {
    const int bits = st.bits;
    const int64_t first = tile_in_block * PACK_TILE;
    const int count = (int)(first + PACK_TILE <= d.n ? PACK_TILE : d.n - first);
    const QuantP qp = quant_params(d);
    const unsigned mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
    const float *p = (const float *)d.src + first;
    const int a = (int)(((uintptr_t)p & 15) >> 2);
    const float4 *base4 = (const float4 *)(p - a);
    const int nvec = (a + count + 3) >> 2;
    // one element through the checked quantiser into its swizzled staging slot
    auto stage_elem = [&](float xv, int el) {
        bool oob = false;
        long long raw;
        const unsigned q = quant_elem(xv, qp, oob, &raw);
        unsigned v;
        if (!st.slow) {   // folded index; bound(q, pmin, pixels) - min (go/group.go:323, :246-247)
            const long long qb = (long long)q < st.pmin ? (long long)q + (long long)qp.P : (long long)q;
            v = (unsigned)(qb - st.min);
        } else {          // the block holds out-of-range indices: the reference's own int64 arithmetic
            const long long qb = st.do_bound ? bound1(raw, st.pmin, (long long)qp.P) : raw;
            v = (unsigned)((unsigned long long)qb - (unsigned long long)st.min) & mask;
        }
        const int L = el >> 5, i = el & 31;
        sv[(L << 5) + ((((i >> 2) ^ (L & 7)) << 2) | (i & 3))] = v;
    };
    if (a == 0 && count == PACK_TILE && !st.slow && st.do_bound && quant_bits_ok(qp)) {
        // whole aligned tile of a block whose indices all lie in [0, pixels]: unchecked quantiser, range test per
        // float4, 32-bit bound / subtract, one 128-bit staging store per float4
        const bool clamp = qp.flags & F_CLAMP;
        const unsigned dsub = 0u - FQ_MAGIC - (unsigned)st.pmin, cadd = (unsigned)(st.pmin - st.min);
#pragma unroll 2
        for (int iv = threadIdx.x; iv < PACK_TILE / 4; iv += FPACK_THREADS) {
            const float4 v4 = __ldcs(base4 + iv);
            const unsigned b0 = quant_bits(v4.x, qp, clamp), b1 = quant_bits(v4.y, qp, clamp);
            const unsigned b2 = quant_bits(v4.z, qp, clamp), b3 = quant_bits(v4.w, qp, clamp);
            const unsigned lo = __vimin3_u32(b0, b1, min(b2, b3)), hi = __vimax3_u32(b0, b1, max(b2, b3));
            if (lo >= FQ_MAGIC && hi < FQ_MAGIC + qp.P) {
                const unsigned d0 = b0 + dsub, d1 = b1 + dsub, d2 = b2 + dsub, d3 = b3 + dsub;   // q - pmin
                uint4 r;
                r.x = min(d0, d0 + qp.P) + cadd; r.y = min(d1, d1 + qp.P) + cadd;
                r.z = min(d2, d2 + qp.P) + cadd; r.w = min(d3, d3 + qp.P) + cadd;
                const int L = iv >> 3;
                *(uint4 *)&sv[(L << 5) + (((iv & 7) ^ (L & 7)) << 2)] = r;
            } else {
                stage_elem(v4.x, 4 * iv); stage_elem(v4.y, 4 * iv + 1); stage_elem(v4.z, 4 * iv + 2); stage_elem(v4.w, 4 * iv + 3);
            }
        }
    } else {
        for (int iv = threadIdx.x; iv < nvec; iv += FPACK_THREADS) {
            const float4 v4 = __ldcs(base4 + iv);
            const float x[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int el = 4 * iv + c - a;
                if (el >= 0 && el < count) stage_elem(x[c], el);
            }
        }
    }
    for (int el = count + threadIdx.x; el < ((count + 1023) & ~1023); el += FPACK_THREADS) {   // pad the last group
        const int L = el >> 5, i = el & 31;
        sv[(L << 5) + ((((i >> 2) ^ (L & 7)) << 2) | (i & 3))] = 0u;
    }
    __syncthreads();
    pack_staged_tile(sv, count, bits, first, st, chain_out);
}

// The same for a contiguous int64 block (intGroup.writeData, go/group.go:242-255) of at most 32 bits.
__device__ __forceinline__ void pack_tile_i64(const BlockDesc &d, const BlockStat &st, int64_t tile_in_block,
                                              uint8_t *chain_out, unsigned *sv) {
    const int bits = st.bits;
    const int64_t first = tile_in_block * PACK_TILE;
    const int count = (int)(first + PACK_TILE <= d.n ? PACK_TILE : d.n - first);
    const long long *p = (const long long *)d.src + first;
    const int a = (int)(((uintptr_t)p & 15) >> 3);
    const longlong2 *base2 = (const longlong2 *)(p - a);
    const int nvec = (a + count + 1) >> 1;
    const unsigned long long mn = (unsigned long long)st.min;
    auto slot = [](int el) { const int L = el >> 5, i = el & 31; return (L << 5) + ((((i >> 2) ^ (L & 7)) << 2) | (i & 3)); };
#pragma unroll 4
    for (int iv = threadIdx.x; iv < nvec; iv += FPACK_THREADS) {
        const longlong2 v = __ldcs(base2 + iv);
        const int e0 = 2 * iv - a;
        const unsigned v0 = (unsigned)((unsigned long long)v.x - mn), v1 = (unsigned)((unsigned long long)v.y - mn);   // go/group.go:246-247
        if (a == 0 && e0 + 1 < count) {
            *(uint2 *)&sv[slot(e0)] = make_uint2(v0, v1);   // an aligned pair shares a 16-byte chunk
        } else {
            if (e0 >= 0) sv[slot(e0)] = v0;
            if (e0 + 1 < count) sv[slot(e0 + 1)] = v1;
        }
    }
    for (int el = count + threadIdx.x; el < ((count + 1023) & ~1023); el += FPACK_THREADS) sv[slot(el)] = 0u;   // pad the last group
    __syncthreads();
    pack_staged_tile(sv, count, bits, first, st, chain_out);
}

}  // namespace
}  // namespace mnw
