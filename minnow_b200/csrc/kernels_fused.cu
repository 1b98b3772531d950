// kernels_fused.cu -- the single-read fused kernels of the minp path.
//
//   k_fused_vec3   minp.Writer.Vectors body (go/minp/minp.go:112-118): per sub-cell
//                  getSubCell (:246-264) + floatGroup.writeData (go/group.go:312-327)
//                  + intGroup.writeData (:242-255) + bit.BufferedArray
//                  (go/bit/bit.go:84-134) + blockIndex.addBlock (go/block_index.go:16-23)
//   k_decode_vec3  minp.Reader.Vectors body (go/minp/minp.go:191-206): per sub-cell
//                  Array.Slice (go/bit/bit.go:29-82) + intGroup.readData
//                  (go/group.go:257-263) + floatGroup.readData (:299-310) + periodic
//                  wrap + setSubCell (go/minp/minp.go:270-288)
//
// Encode data flow (one cluster of CS CTAs per sub-cell, CTA r owns rows
// [r*ROWS, (r+1)*ROWS) of the sub-cell = elements [r*CHUNK, (r+1)*CHUNK) of each
// of the three axis blocks):
//   1. every AoS row is read once with coalesced 128-bit loads; thread (row, c4)
//      always sees the same axis phase, so the per-axis parameters and the running
//      min/max live in registers without any selection;
//   2. the pixel index q is rotated to w = (q - q0 + K) mod pixels (the coordinate
//      in which the periodic arc of go/group.go:384-409 is a plain interval) and
//      its low 16 bits are kept in shared memory (3 * CHUNK * 2 bytes);
//   3. per-CTA statistics are all-gathered through distributed shared memory, one
//      cluster barrier, every CTA finalises (min, bits, nbytes) redundantly;
//   4. the block's byte offset in its group comes from a decoupled look-back over
//      the earlier sub-cells of the file (units are claimed through a ticket, so a
//      predecessor is always running or done);
//   5. each warp packs groups of 1024 elements (32 per lane, bit width resolved at
//      compile time), transposes the words in place in shared memory and writes
//      them out with coalesced stores to the BYTE-aligned destination.
// Blocks wider than 16 bits are listed and packed afterwards by k_pack from
// global memory (then the second read mostly hits L2); blocks holding pixel
// indices outside [0, pixels] raise abort_flag and the call is redone by the
// generic path.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>

#include "device_math.cuh"
#include "engine.cuh"
#include "fused.cuh"
#include "launch.cuh"

namespace cg = cooperative_groups;

namespace mnw {

namespace {

struct XStat {  // one CTA's statistics of one axis block
    unsigned wmin, wmax;
    int qmin, qmax;
    unsigned oob, pad0, pad1, pad2;
};

struct Fin {  // finalised block, identical in every CTA of the cluster
    long long off;    // exclusive byte offset of the block in its group
    int bits;
    int mode;         // 1: pack from shared memory, 0: nothing to pack here
    unsigned base;    // v = w - base, + padj when negative
    unsigned padj;
};

struct FusedArgs {
    const float *aos;
    const FloatParams *tab;
    int tab_per_file;
    int nfile, subcells;
    long long nunits, sc3;
    BlockStat *stats;
    int64_t *mins, *bits, *offsets, *out_len;
    uint8_t *out;
    long long axis_stride;
    FusedWork W;
};

constexpr unsigned long long PUB_AGG = 1ULL << 62, PUB_PREFIX = 2ULL << 62, PUB_VALUE = (1ULL << 62) - 1ULL;

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Exclusive prefix of the published byte sizes of blocks [first, b): the
// decoupled look-back of a chained scan, 32 predecessors per step.
__device__ long long lookback(const unsigned long long *pub, long long first, long long b) {
    const int lane = threadIdx.x & 31;
    long long sum = 0;
    for (long long hi = b; hi > first; hi -= 32) {
        const long long idx = hi - 1 - lane;
        const bool valid = idx >= first;
        unsigned long long v = 0;
        if (valid) {
            do { v = ld_relaxed(pub + idx); } while ((v >> 62) == 0);
        }
        const unsigned pmask = __ballot_sync(0xffffffffu, valid && (v >> 62) == 2);
        long long val = valid ? (long long)(v & PUB_VALUE) : 0;
        if (pmask) {  // the nearest predecessor with an inclusive prefix ends the walk
            const int stop = __ffs(pmask) - 1;
            if (lane > stop) val = 0;
        }
        for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        sum += val;
        if (pmask) break;
    }
    return sum;
}

// Exact lane path of the quantiser: anything the fast quotient could not vouch for.
// Returns the folded pixel index (pixels -> 0) or flags the element out of range.
__device__ __noinline__ int quantize_rare(float x, float low, float dx, int P, unsigned &oob) {
    long long q = quantize_exact(x, low, dx);
    if (q == (long long)P) return 0;
    if ((unsigned long long)q < (unsigned long long)P) return (int)q;
    oob = 1;
    return 0;
}

// 32 values of B bits -> B words, all shifts resolved at compile time.
template <int B>
__device__ __forceinline__ void pack32(const unsigned (&v)[32], unsigned (&o)[16]) {
#pragma unroll
    for (int j = 0; j < 16; j++) o[j] = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const int bit = i * B, wd = bit >> 5, sh = bit & 31;
        o[wd] |= v[i] << sh;
        if (sh + B > 32) o[wd + 1] |= v[i] >> (32 - sh);
    }
}

// Write `nbytes` of the little-endian word stream to dst (any byte alignment) with one
// warp; word j of the stream is ld(j).  Interior words go out as aligned 32-bit stores,
// the first/last partial word as single bytes (the neighbours own the other bytes).
template <class Ld>
__device__ __forceinline__ void warp_store_stream(uint8_t *dst, long long nbytes, int lane, Ld ld) {
    const uintptr_t A = (uintptr_t)dst;
    const int a = (int)(A & 3);
    uint32_t *base = (uint32_t *)(A - a);
    const long long nwords = (a + nbytes + 3) >> 2;
    const int nsrc = (int)((nbytes + 3) >> 2);
    for (long long j = lane; j < nwords; j += 32) {
        uint32_t lo = (j > 0 && j - 1 < nsrc) ? ld((int)j - 1) : 0u;
        uint32_t hi = j < nsrc ? ld((int)j) : 0u;
        uint32_t w = __funnelshift_rc(lo, hi, 32 - 8 * a);
        long long t0 = 4 * j - a;
        if (t0 >= 0 && t0 + 4 <= nbytes) {
            base[j] = w;
        } else {
            uint8_t *bp = (uint8_t *)(base + j);
            for (int k = 0; k < 4; k++) {
                long long t = t0 + k;
                if (t >= 0 && t < nbytes) bp[k] = (uint8_t)(w >> (8 * k));
            }
        }
    }
}

template <int CS>
__device__ __forceinline__ void cluster_sync_all() {
    if constexpr (CS > 1) cg::this_cluster().sync();
    else __syncthreads();
}

}  // namespace

// ---------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------
template <int NSUB, int CS, int NT, int UNROLL>
__global__ void __launch_bounds__(NT, 1) k_fused_vec3(const FusedArgs A) {
    constexpr int N = NSUB * NSUB * NSUB;   // elements per block
    constexpr int CHUNK = N / CS;           // elements per CTA and axis
    constexpr int ROWS = CHUNK / NSUB;      // sub-cell rows per CTA
    constexpr int R4 = 3 * NSUB / 4;        // float4 per row
    constexpr int RPP = NT / R4;            // rows per pass of the CTA
    constexpr int PASSES = ROWS / RPP;
    constexpr int GPA = CHUNK / 1024;       // pack groups per axis
    constexpr int NW = NT / 32;
    static_assert(NT % R4 == 0 && ROWS % RPP == 0, "threads tile the rows exactly");
    static_assert((RPP * NSUB) % 512 == 0, "swizzle term must be a per-thread constant");
    static_assert(CHUNK % 1024 == 0 && N % CS == 0, "whole pack groups per CTA");
    static_assert(PASSES % UNROLL == 0, "unroll divides the passes");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned short *stage = (unsigned short *)smem_raw;   // [3][CHUNK], swizzled
    __shared__ XStat s_x[2][CS][3];                       // all-gathered statistics, by parity
    __shared__ long long s_unit[2];                       // claimed unit, by parity
    __shared__ unsigned s_red[NW][3][5];
    __shared__ Fin s_fin[3];
    __shared__ long long s_q0[3];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned rank = 0;
    if constexpr (CS > 1) rank = cg::this_cluster().block_rank();

    // ---- thread geometry (constant for the whole kernel) ----
    const int col4 = tid % R4, rsub = tid / R4;
    const int a0 = col4 % 3;                 // axis of this thread's first float
    int soff[4];                             // staging offsets of the 4 floats, pass 0
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int ax = (a0 + c) % 3;
        const int e = rsub * NSUB + (4 * col4 + c) / 3;
        soff[c] = ax * CHUNK + (e ^ (((e >> 6) & 7) << 3));
    }

    int par = 0;
    if (rank == 0 && tid == 0) {
        const long long u = (long long)atomicAdd(A.W.ticket, 1u);
        if constexpr (CS > 1) {
            for (unsigned r = 0; r < CS; r++) *cg::this_cluster().map_shared_rank(&s_unit[0], r) = u;
        } else {
            s_unit[0] = u;
        }
    }
    cluster_sync_all<CS>();

    for (long long unit = s_unit[0]; unit < A.nunits; unit = s_unit[par], (void)0) {
        const long long f = unit / A.sc3, sc = unit - f * A.sc3;
        const int S = A.subcells, nfile = A.nfile;
        const int ix0 = NSUB * (int)(sc % S), iy0 = NSUB * (int)((sc / S) % S), iz0 = NSUB * (int)(sc / ((long long)S * S));
        const float *cube = A.aos + 3 * f * (long long)nfile * nfile * nfile;
        const FloatParams *tab = A.tab + (A.tab_per_file ? 3 * f : 0);

        // ---- per-axis parameters in this thread's axis order (relative axis j = actual (a0+j)%3) ----
        float low[3], rcp[3], ndx[3];
        int P[3];
        unsigned Pm1[3], C[3];
        unsigned oob = 0;
        long long q0[3];
        {
            const long long idx0 = ix0 + (long long)iy0 * nfile + (long long)iz0 * nfile * nfile;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int ax = (a0 + j) % 3;
                const FloatParams fp = tab[ax];
                low[j] = fp.low; rcp[j] = fp.rcp; ndx[j] = -fp.dx;
                P[j] = (int)fp.pixels;
                Pm1[j] = (fp.flags & F_FASTDIV) ? (unsigned)(P[j] - 1) : 0u;
                q0[j] = quantize_exact(__ldg(cube + 3 * idx0 + ax), fp.low, fp.dx);   // x[0] of the block
                const bool ok = (unsigned long long)q0[j] < (unsigned long long)P[j];
                C[j] = ok ? (unsigned)arc_rotation(q0[j], P[j]) : 0u;
                if (!ok) oob = 1;   // periodicMin starting outside [0, pixels): exact path only
            }
        }
        if (tid == 0) { s_q0[0] = q0[0]; s_q0[1] = q0[1]; s_q0[2] = q0[2]; }   // a0 == 0 here: actual order

        unsigned wmin[3] = {~0u, ~0u, ~0u}, wmax[3] = {0u, 0u, 0u};
        int qmin[3] = {INT_MAX, INT_MAX, INT_MAX}, qmax[3] = {INT_MIN, INT_MIN, INT_MIN};

        __syncthreads();   // the previous unit's pack phase has released the staging area

        // ---- phase 1: one read of the sub-cell rows ----
        const long long plane = (long long)nfile * nfile;
#pragma unroll 1
        for (int p0 = 0; p0 < PASSES; p0 += UNROLL) {
            float4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const int rowg = (int)rank * ROWS + rsub + RPP * (p0 + u);
                const int jz = rowg / NSUB, jy = rowg % NSUB;
                const long long idx = ix0 + (long long)(jy + iy0) * nfile + (long long)(jz + iz0) * plane;
                v[u] = __ldcs((const float4 *)(cube + 3 * idx) + col4);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const float x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int j = c % 3;
                    int qi = quantize_fast(x[c], low[j], rcp[j], ndx[j]);
                    if ((unsigned)(qi - 1) >= Pm1[j]) qi = quantize_rare(x[c], low[j], -ndx[j], P[j], oob);
                    unsigned w = (unsigned)qi + C[j];
                    w = min(w, w - (unsigned)P[j]);
                    wmin[j] = min(wmin[j], w); wmax[j] = max(wmax[j], w);
                    qmin[j] = min(qmin[j], qi); qmax[j] = max(qmax[j], qi);
                    stage[soff[c] + (p0 + u) * (RPP * NSUB)] = (unsigned short)w;
                }
            }
        }

        // ---- CTA reduction, in actual axis order ----
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int j = (k - a0 + 3) % 3;   // relative index of actual axis k
            unsigned a = j == 0 ? wmin[0] : (j == 1 ? wmin[1] : wmin[2]);
            unsigned b = j == 0 ? wmax[0] : (j == 1 ? wmax[1] : wmax[2]);
            int c = j == 0 ? qmin[0] : (j == 1 ? qmin[1] : qmin[2]);
            int d = j == 0 ? qmax[0] : (j == 1 ? qmax[1] : qmax[2]);
            a = __reduce_min_sync(0xffffffffu, a);
            b = __reduce_max_sync(0xffffffffu, b);
            c = __reduce_min_sync(0xffffffffu, c);
            d = __reduce_max_sync(0xffffffffu, d);
            if (lane == 0) { s_red[warp][k][0] = a; s_red[warp][k][1] = b; s_red[warp][k][2] = (unsigned)c; s_red[warp][k][3] = (unsigned)d; }
        }
        // out-of-range flags are not tracked per axis: any of them sends the whole unit to the exact path
        oob = __any_sync(0xffffffffu, oob);
        if (lane == 0) s_red[warp][0][4] = oob;
        __syncthreads();
        if (tid < 3 * CS) {   // thread (r, k): publish axis k of this CTA into CTA r
            const int r = tid / 3, k = tid % 3;
            XStat x;
            x.wmin = ~0u; x.wmax = 0u; x.qmin = INT_MAX; x.qmax = INT_MIN; x.oob = 0; x.pad0 = x.pad1 = x.pad2 = 0;
            for (int wi = 0; wi < NW; wi++) {
                x.wmin = min(x.wmin, s_red[wi][k][0]); x.wmax = max(x.wmax, s_red[wi][k][1]);
                x.qmin = min(x.qmin, (int)s_red[wi][k][2]); x.qmax = max(x.qmax, (int)s_red[wi][k][3]);
                x.oob |= s_red[wi][0][4];
            }
            if constexpr (CS > 1) *cg::this_cluster().map_shared_rank(&s_x[par][rank][k], r) = x;
            else s_x[par][0][k] = x;
        }
        if (rank == 0 && tid == 32) {   // claim the next unit for the whole cluster
            const long long u = (long long)atomicAdd(A.W.ticket, 1u);
            if constexpr (CS > 1) {
                for (unsigned r = 0; r < CS; r++) *cg::this_cluster().map_shared_rank(&s_unit[par ^ 1], r) = u;
            } else {
                s_unit[par ^ 1] = u;
            }
        }
        cluster_sync_all<CS>();

        // ---- finalise + look-back: warp k handles axis k (every CTA, redundantly) ----
        if (warp < 3) {
            const int k = warp;
            const long long b = f * 3 * A.sc3 + k * A.sc3 + sc;      // block id in the batch
            const long long chain0 = b - sc;                          // first block of the group
            XStat x = s_x[par][0][k];
            for (int r = 1; r < CS; r++) {
                const XStat y = s_x[par][r][k];
                x.wmin = min(x.wmin, y.wmin); x.wmax = max(x.wmax, y.wmax);
                x.qmin = min(x.qmin, y.qmin); x.qmax = max(x.qmax, y.qmax);
                x.oob |= y.oob;
            }
            const FloatParams fp = tab[k];
            const long long Pk = fp.pixels, half = Pk / 2, K = Pk - half - 1;
            const long long q0k = s_q0[k];
            long long mn, pmin;
            unsigned long long maxoff;
            unsigned base, padj;
            bool wide;
            const unsigned long long spread = (unsigned long long)x.wmax - x.wmin + 1ULL;
            if (spread > (unsigned long long)half) {   // arc too wide: periodicMin returns 0
                wide = true;
                pmin = 0; mn = x.qmin; maxoff = (unsigned long long)((long long)x.qmax - x.qmin);
                base = (unsigned)arc_rotation(q0k, Pk) + (unsigned)x.qmin; padj = (unsigned)Pk;
            } else {
                wide = false;
                long long m = q0k + ((long long)x.wmin - K);
                if (m < 0) m += Pk;
                pmin = m; mn = m; maxoff = spread - 1ULL;
                base = x.wmin; padj = (unsigned)Pk;
            }
            int bits = precision_needed(maxoff);
            long long nbytes = array_bytes(bits, N);
            const bool slow = x.oob != 0;
            if (slow) { bits = 0; nbytes = 0; }
            if (Pk > 65536) {   // staged values are only the low 16 bits of w
                base &= 0xffffu; padj = 65536u;
            }
            if (rank == 0 && lane == 0) st_relaxed(A.W.pub + b, PUB_AGG | (unsigned long long)nbytes);
            const long long off = lookback(A.W.pub, chain0, b);
            int mode = (bits >= 1 && bits <= 16 && !(wide && Pk > 65536)) ? 1 : 0;
            if (lane == 0) {
                if (off + nbytes > A.axis_stride) {   // never write past the caller's buffer
                    mode = 0;
                    if (rank == 0) atomicExch(A.W.err, 2);
                } else if (rank == 0 && !slow && bits > 0 && mode == 0) {
                    A.W.repack_list[atomicAdd(A.W.repack_count, 1)] = b;
                }
                if (rank == 0) {
                    st_relaxed(A.W.pub + b, PUB_PREFIX | (unsigned long long)(off + nbytes));
                    if (slow) atomicExch(A.W.abort_flag, 1);
                    BlockStat st = {};
                    st.pmin = pmin; st.min = mn; st.nbytes = nbytes; st.out_off = off; st.do_bound = 1; st.bits = bits;
                    st.q0 = q0k; st.oob = x.oob;
                    A.stats[b] = st;
                    if (A.mins) A.mins[b] = mn;
                    if (A.bits) A.bits[b] = bits;
                    if (A.offsets) A.offsets[b] = off;
                    if (A.out_len && sc == A.sc3 - 1) A.out_len[f * 3 + k] = off + nbytes;
                }
                Fin fin;
                fin.off = off; fin.bits = bits; fin.mode = mode; fin.base = base; fin.padj = padj;
                s_fin[k] = fin;
            }
        }
        __syncthreads();

        // ---- phase 2: pack groups of 1024 elements straight from shared memory ----
        for (int g = warp; g < 3 * GPA; g += NW) {
            const int k = g / GPA, gi = g - k * GPA;
            const Fin fin = s_fin[k];
            if (fin.mode == 0) continue;
            const int eb = gi * 1024 + 32 * lane;   // this lane's first element within the CTA's chunk
            const int sw = (eb >> 6) & 7;
            unsigned v[32];
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const int chunk = ((eb >> 3) + s) ^ sw;
                const uint4 r = *(const uint4 *)(stage + k * CHUNK + (chunk << 3));
                const unsigned rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    unsigned lo = (rr[t] & 0xffffu) - fin.base, hi = (rr[t] >> 16) - fin.base;
                    v[8 * s + 2 * t] = min(lo, lo + fin.padj) & 0xffffu;
                    v[8 * s + 2 * t + 1] = min(hi, hi + fin.padj) & 0xffffu;
                }
            }
            unsigned o[16];
            switch (fin.bits) {
#define MNW_CASE(B) case B: pack32<B>(v, o); break;
                MNW_CASE(1) MNW_CASE(2) MNW_CASE(3) MNW_CASE(4) MNW_CASE(5) MNW_CASE(6) MNW_CASE(7) MNW_CASE(8)
                MNW_CASE(9) MNW_CASE(10) MNW_CASE(11) MNW_CASE(12) MNW_CASE(13) MNW_CASE(14) MNW_CASE(15) MNW_CASE(16)
#undef MNW_CASE
                default: break;
            }
            // in-place transpose: the group's 2 KiB of staging now holds its packed words
            unsigned *region = (unsigned *)(stage + k * CHUNK + gi * 1024);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (j < fin.bits) {
                    const int W = lane * fin.bits + j;
                    region[W ^ (W >> 5)] = o[j];
                }
            }
            __syncwarp();
            const long long e0 = (long long)rank * CHUNK + (long long)gi * 1024;   // element index in the block
            uint8_t *dst = A.out + (f * 3 + k) * A.axis_stride + fin.off + ((e0 * fin.bits) >> 3);
            warp_store_stream(dst, 128LL * fin.bits, lane, [&](int j) { return region[j ^ (j >> 5)]; });
        }
        par ^= 1;
    }
}

// ---------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------
struct DecVec3Args {
    const uint8_t *data;
    long long stream_len;   // bytes reserved per (file, axis) stream
    const int64_t *offsets, *mins, *bits;
    const FloatParams *tab;
    int tab_per_file;
    float wrap_L;
    int jmode;
    unsigned long long seed, block_id0;
    int nfile, subcells;
    long long sc3;
    float *out;
};

// value i of a block whose stream starts at the 4-byte aligned word pointer `base`,
// `a8` bits into it: bits <= 32
__device__ __forceinline__ unsigned extract32(const uint32_t *__restrict__ base, unsigned bitpos, int bits, unsigned mask) {
    const unsigned wi = bitpos >> 5, sh = bitpos & 31;
    const uint32_t w0 = __ldg(base + wi);
    const uint32_t w1 = (sh + bits > 32) ? __ldg(base + wi + 1) : 0u;
    return __funnelshift_r(w0, w1, sh) & mask;
}

template <int NSUB, int NT>
__global__ void __launch_bounds__(NT) k_decode_vec3(const DecVec3Args A) {
    constexpr int N = NSUB * NSUB * NSUB;
    constexpr int SLAB = N < 4096 ? N : 4096;   // elements per CTA and axis
    constexpr int SLABS = N / SLAB;
    constexpr int ROWS = SLAB / NSUB;
    constexpr int R4 = 3 * NSUB / 4;
    constexpr int RPP = NT / R4;
    static_assert(NT % R4 == 0, "threads tile the rows exactly");
    __shared__ __align__(16) float dec[3][SLAB];

    const int tid = threadIdx.x;
    const long long unit = blockIdx.x / SLABS;
    const int slab = (int)(blockIdx.x - unit * SLABS);
    const long long f = unit / A.sc3, sc = unit - f * A.sc3;
    const FloatParams *tab = A.tab + (A.tab_per_file ? 3 * f : 0);

    // ---- phase 1: unpack + dequantise each axis into planar shared memory ----
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        const long long b = f * 3 * A.sc3 + k * A.sc3 + sc;
        const long long mn = A.mins[b];
        const int bits = (int)A.bits[b];
        const FloatParams fp = tab[k];
        const long long Pk = fp.pixels;
        const bool periodic = fp.flags & F_PERIODIC;
        const uint8_t *stream = A.data + (f * 3 + k) * A.stream_len + A.offsets[b];
        const int a = (int)((uintptr_t)stream & 3);
        const uint32_t *base = (const uint32_t *)(stream - a);
        const uint32_t key = jitter_key(A.seed, A.block_id0 + (unsigned long long)b);
        // 32-bit fast path: every q of the block is known to land in (-2^23, 2^23) after bound(),
        // so float32 holds it exactly; anything else takes the 64-bit path below
        const long long qlo = mn, qhi = mn + (long long)((bits >= 1 && bits <= 30) ? ((1u << bits) - 1u) : 0u);
        const bool small = bits <= 30 && mn > -(1LL << 30) && mn < (1LL << 30) && Pk > 0 && Pk < (1LL << 23) &&
                           (periodic ? (qlo >= -Pk && qhi < 2 * Pk) : (qlo > -(1LL << 23) && qhi < (1LL << 23)));
        if (small) {   // everything fits 32-bit integers and float32 holds q exactly
            const unsigned mask = bits ? (0xffffffffu >> (32 - bits)) : 0u;
            const int mn32 = (int)mn, P32 = (int)Pk;
#pragma unroll 4
            for (int i = tid; i < SLAB; i += NT) {
                const unsigned e = (unsigned)(slab * SLAB + i);
                const unsigned v = bits ? extract32(base, 8u * a + e * (unsigned)bits, bits, mask) : 0u;
                int q = mn32 + (int)v;                                      // go/group.go:262
                if (periodic) { if (q < 0) q += P32; else if (q >= P32) q -= P32; }   // bound(q, 0, pixels), :303
                float t;
                if (A.jmode == 1) {
                    // u = h24 * 2^-24 is exact in float32 and q + u needs at most 47 bits: one FMA
                    // rounds once, exactly like float32(float64(q) + u) (go/group.go:308)
                    const float h = (float)(jitter_hash_keyed(key, e) >> 8);
                    t = __fmaf_rn(h, 0x1p-24f, (float)q);
                } else {
                    t = __fadd_rn((float)q, 0.5f);   // q < 2^23: q + 0.5 is exact
                }
                float o = __fadd_rn(__fmul_rn(fp.dx, t), fp.low);
                if (A.wrap_L > 0.0f) {                                        // go/minp/minp.go:195-203
                    if (o < 0.0f) o = __fadd_rn(o, A.wrap_L);
                    else if (o >= A.wrap_L) o = __fsub_rn(o, A.wrap_L);
                }
                dec[k][i] = o;
            }
        } else {
            for (int i = tid; i < SLAB; i += NT) {
                const long long e = (long long)slab * SLAB + i;
                unsigned long long v = 0;
                if (bits) {
                    const unsigned long long bitpos = 8ULL * a + (unsigned long long)e * bits;
                    const long long wi = (long long)(bitpos >> 5);
                    const int sh = (int)(bitpos & 31);
                    const uint32_t w0 = __ldg(base + wi);
                    const uint32_t w1 = (sh + bits > 32) ? __ldg(base + wi + 1) : 0u;
                    v = (((unsigned long long)w1 << 32) | w0) >> sh;
                    if (sh + bits > 64) v |= (unsigned long long)__ldg(base + wi + 2) << (64 - sh);
                    if (bits < 64) v &= (1ULL << bits) - 1ULL;
                }
                long long q = (long long)((unsigned long long)mn + v);
                if (periodic) q = bound1(q, 0, Pk);
                double u = 0.5;
                if (A.jmode == 1) u = (double)(jitter_hash_keyed(key, (uint32_t)e) >> 8) * 0x1p-24;
                const float t = __double2float_rn(__dadd_rn(__ll2double_rn(q), u));
                float o = __fadd_rn(__fmul_rn(fp.dx, t), fp.low);
                if (A.wrap_L > 0.0f) {
                    if (o < 0.0f) o = __fadd_rn(o, A.wrap_L);
                    else if (o >= A.wrap_L) o = __fsub_rn(o, A.wrap_L);
                }
                dec[k][i] = o;
            }
        }
    }
    __syncthreads();

    // ---- phase 2: whole AoS rows, coalesced 128-bit stores (setSubCell, go/minp/minp.go:270-288) ----
    const int S = A.subcells, nfile = A.nfile;
    const int ix0 = NSUB * (int)(sc % S), iy0 = NSUB * (int)((sc / S) % S), iz0 = NSUB * (int)(sc / ((long long)S * S));
    float *cube = A.out + 3 * f * (long long)nfile * nfile * nfile;
    const int col4 = tid % R4, rsub = tid / R4, a0 = col4 % 3;
    int src[4];
#pragma unroll
    for (int c = 0; c < 4; c++) src[c] = ((a0 + c) % 3) * SLAB + (4 * col4 + c) / 3;
    const float *flat = &dec[0][0];
    for (int rl = rsub; rl < ROWS; rl += RPP) {
        const int rowg = slab * ROWS + rl;
        const int jz = rowg / NSUB, jy = rowg % NSUB;
        const long long idx = ix0 + (long long)(jy + iy0) * nfile + (long long)(jz + iz0) * nfile * nfile;
        float4 o;
        o.x = flat[src[0] + rl * NSUB]; o.y = flat[src[1] + rl * NSUB];
        o.z = flat[src[2] + rl * NSUB]; o.w = flat[src[3] + rl * NSUB];
        __stcs((float4 *)(cube + 3 * idx) + col4, o);
    }
}

// ---------------------------------------------------------------------------
// self-test of quantize_fast against the IEEE divide over a range of float bit patterns
// ---------------------------------------------------------------------------
__global__ void k_selftest_fastdiv(FloatParams fp, unsigned long long first, unsigned long long count,
                                   unsigned long long *mismatches, unsigned long long *accepted) {
    unsigned long long bad = 0, acc = 0;
    const int P = (int)fp.pixels;
    const unsigned Pm1 = (fp.flags & F_FASTDIV) ? (unsigned)(P - 1) : 0u;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < count;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)(first + i));
        const int qi = quantize_fast(x, fp.low, fp.rcp, -fp.dx);
        if ((unsigned)(qi - 1) < Pm1) {   // the fast result would be used as is
            acc++;
            if ((long long)qi != quantize_exact(x, fp.low, fp.dx)) bad++;
        }
    }
    for (int o = 16; o; o >>= 1) {
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(mismatches, bad); atomicAdd(accepted, acc); }
}

void launch_selftest_fastdiv(Launcher &L, const FloatParamsHost &fp, unsigned long long first,
                             unsigned long long count, unsigned long long *d_out2) {
    cudaMemsetAsync(d_out2, 0, 16, L.stream);
    if (count == 0) return;
    k_selftest_fastdiv<<<148 * 8, 256, 0, L.stream>>>(fp, first, count, d_out2, d_out2 + 1);
    L.count++;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
size_t fused_work_bytes(int64_t nblocks) { return (size_t)(16 * nblocks + 256); }

bool fused_vec3_supported(const FloatParamsHost *fp, int64_t nparams, int nfile, int subcells, const void *aos) {
    if (subcells <= 0 || nfile % subcells) return false;
    const int nsub = nfile / subcells;
    if (nsub != 16 && nsub != 32 && nsub != 64) return false;
    if (((uintptr_t)aos & 15) != 0) return false;
    for (int64_t i = 0; i < nparams; i++) {
        const FloatParamsHost &p = fp[i];
        if (!(p.flags & F_PERIODIC) || (p.flags & (F_LOG10 | F_CLAMP))) return false;
        if (p.pixels < 1 || p.pixels >= (1LL << 30)) return false;
    }
    return true;
}

template <int NSUB, int CS, int NT, int UNROLL>
static cudaError_t launch_fused_vec3_t(Launcher &L, const FusedArgs &A) {
    auto kern = k_fused_vec3<NSUB, CS, NT, UNROLL>;
    const size_t smem = (size_t)6 * (NSUB * NSUB * NSUB / CS);
    static bool configured = false;
    static int max_clusters = 0;
    cudaError_t e;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = L.stream;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (!configured) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cfg.gridDim = dim3(CS);
        e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
        if (e != cudaSuccess) return e;
        if (max_clusters < 1) return cudaErrorLaunchOutOfResources;
        if (getenv("MNW_DEBUG")) fprintf(stderr, "k_fused_vec3<%d,%d,%d>: %d co-resident clusters, %zu B dynamic smem\n", NSUB, CS, NT, max_clusters, smem);
        configured = true;
    }
    long long clusters = A.nunits < max_clusters ? A.nunits : max_clusters;
    cfg.gridDim = dim3((unsigned)(clusters * CS));
    L.begin("k_fused_vec3");
    e = cudaLaunchKernelEx(&cfg, kern, A);
    L.end();
    L.count++;
    return e;
}

cudaError_t launch_fused_vec3(Launcher &L, const FusedWork &W, const FloatParams *tab, int tab_per_file,
                              const float *aos, int nfile, int subcells, int64_t nfiles, BlockStat *stats,
                              int64_t *mins, int64_t *bits, int64_t *offsets, int64_t *out_len, uint8_t *out,
                              int64_t out_axis_stride) {
    FusedArgs A = {};
    A.aos = aos; A.tab = tab; A.tab_per_file = tab_per_file; A.nfile = nfile; A.subcells = subcells;
    A.sc3 = (long long)subcells * subcells * subcells;
    A.nunits = nfiles * A.sc3;
    A.stats = stats; A.mins = mins; A.bits = bits; A.offsets = offsets; A.out_len = out_len; A.out = out;
    A.axis_stride = out_axis_stride; A.W = W;
    if (A.nunits == 0) return cudaSuccess;
    switch (nfile / subcells) {
        case 64: {
            static const int nt = getenv("MNW_FUSED_NT") ? atoi(getenv("MNW_FUSED_NT")) : 768;   // tuning knob
            if (nt == 384) return launch_fused_vec3_t<64, 8, 384, 8>(L, A);
            return launch_fused_vec3_t<64, 8, 768, 4>(L, A);
        }
        case 32: return launch_fused_vec3_t<32, 1, 768, 4>(L, A);
        case 16: return launch_fused_vec3_t<16, 1, 384, 4>(L, A);
    }
    return cudaErrorNotSupported;
}

bool fused_decode_vec3_supported(int nfile, int subcells, const void *aos_out) {
    if (subcells <= 0 || nfile % subcells) return false;
    const int nsub = nfile / subcells;
    if (nsub != 16 && nsub != 32 && nsub != 64) return false;
    return ((uintptr_t)aos_out & 15) == 0;
}

cudaError_t launch_fused_decode_vec3(Launcher &L, const DecodeHost &h, int64_t nfiles) {
    DecVec3Args A = {};
    A.data = h.data; A.stream_len = h.stream_len; A.offsets = h.offsets; A.mins = h.mins; A.bits = h.bits;
    A.tab = h.tab; A.tab_per_file = h.tab_per_file; A.wrap_L = h.wrap_L; A.jmode = h.jmode; A.seed = h.seed;
    A.block_id0 = h.block_id0; A.nfile = h.nfile; A.subcells = h.subcells;
    A.sc3 = (long long)h.subcells * h.subcells * h.subcells;
    A.out = (float *)h.out;
    const long long units = nfiles * A.sc3;
    if (units == 0) return cudaSuccess;
    const int nsub = h.nfile / h.subcells;
    const long long n = (long long)nsub * nsub * nsub, slabs = n < 4096 ? 1 : n / 4096;
    if (units * slabs >= (1LL << 31)) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)(units * slabs);
    L.begin("k_decode_vec3");
    switch (nsub) {
        case 64: k_decode_vec3<64, 384><<<grid, 384, 0, L.stream>>>(A); break;
        case 32: k_decode_vec3<32, 384><<<grid, 384, 0, L.stream>>>(A); break;
        case 16: k_decode_vec3<16, 384><<<grid, 384, 0, L.stream>>>(A); break;
        default: L.end(); return cudaErrorNotSupported;
    }
    L.end();
    L.count++;
    return cudaGetLastError();
}

}  // namespace mnw
