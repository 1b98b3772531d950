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
#include <cstring>
#include <type_traits>

#include "fused_detail.cuh"

namespace mnw {

namespace {

// Tickets walk the files round-robin in chunks of G consecutive sub-cells (ticket t -> chunk t / G, sub-cell
// (chunk / nfiles) * G + t % G of file chunk % nfiles): the predecessor of a unit in its look-back chain (the previous
// sub-cell of its file) is then about G * nfiles tickets old for the first unit of a chunk instead of the ticket just
// before it -- and still always claimed earlier (running or done) -- while the G units of a chunk, neighbours in x
// whose rows are contiguous in memory, are read at the same time.  G = sc3 is the plain file-by-file order.
// Tickets beyond the batch map to themselves.
__device__ __forceinline__ long long claim_unit(const FusedArgs &A) {
    const long long t = (long long)atomicAdd(A.W.ticket, 1u);
    if (t >= A.nunits) return t;
    const long long nfiles = A.nunits / A.sc3, G = A.ticket_chunk;
    const long long c = t / G, w = t - c * G;
    return (c % nfiles) * A.sc3 + (c / nfiles) * G + w;
}

// Batch-local statistics of phase 1, in the thread's relative axis order.
struct LocalStat {
    unsigned wmin[3], wmax[3];
    int qmin[3], qmax[3];
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int j = 0; j < 3; j++) { wmin[j] = ~0u; wmax[j] = 0u; qmin[j] = INT_MAX; qmax[j] = INT_MIN; }
    }
};

}  // namespace

// ---------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------
template <int NSUB, int CS, int NT, int UNROLL, int MINB, bool PIPE>
__global__ void __launch_bounds__(NT, MINB) k_fused_vec3(const FusedArgs A) {
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
    static_assert(PASSES % UNROLL == 0 && UNROLL % 2 == 0, "batches of float4 pairs");
    // rows of one batch: either all in one z-plane (RPP*UNROLL divides NSUB) or one row per
    // plane step (RPP is a multiple of NSUB); both make the row offset linear in u
    static_assert(NSUB % (RPP * UNROLL) == 0 || RPP % NSUB == 0, "batch rows are equally spaced");
    if (A.W.skip && *A.W.skip) return;   // uniform over the grid: written before the launch

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned short *stage = (unsigned short *)smem_raw;   // [3][CHUNK], swizzled
    __shared__ XStat s_x[2][CS][3];                       // all-gathered statistics, by parity
    __shared__ long long s_unit[2];                       // claimed unit, by parity
    __shared__ unsigned s_red[NW][3][5];
    __shared__ Fin s_fin[3];
    __shared__ long long s_q0[3];
    __shared__ long long s_off[3];     // byte offset of the unit's blocks (-1: does not fit the output)
    __shared__ int s_offgen[3];        // == gen once s_off is valid for this unit
    __shared__ int s_gctr;             // next pack group of the unit

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
    const int S = A.subcells, nfile = A.nfile;
    const unsigned row4 = 3u * (unsigned)nfile / 4u, plane4 = row4 * (unsigned)nfile;   // float4 units

    int par = 0, gen = 1;
    if (tid < 3) s_offgen[tid] = 0;
    if (rank == 0 && tid == 0) {
        const long long u = claim_unit(A);
        if constexpr (CS > 1) {
            for (unsigned r = 0; r < CS; r++) *cg::this_cluster().map_shared_rank(&s_unit[0], r) = u;
        } else {
            s_unit[0] = u;
        }
    }
    cluster_sync_all<CS>();

    for (long long unit = s_unit[0]; unit < A.nunits; unit = s_unit[par]) {
        const long long f = unit / A.sc3, sc = unit - f * A.sc3;
        const int ix0 = NSUB * (int)(sc % S), iy0 = NSUB * (int)((sc / S) % S), iz0 = NSUB * (int)(sc / ((long long)S * S));
        const float *cube = A.aos + 3 * f * (long long)nfile * nfile * nfile;
        const FloatParams *tab = A.tab + (A.tab_per_file ? 3 * f : 0);
        // first float4 of this thread's column in row 0 of the sub-cell
        const float4 *pbase = (const float4 *)cube + ((unsigned)(3 * ix0) / 4u + (unsigned)iy0 * row4 + (unsigned)iz0 * plane4) + col4;

        // ---- per-axis parameters in this thread's axis order (relative axis j = actual (a0+j)%3) ----
        float low[3], rcp[3], ndx[3];
        int P[3];
        unsigned Pm1[3], C[3];
        unsigned oob = 0;
        {
            long long q0[3];
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
            if (tid == 0) { s_q0[0] = q0[0]; s_q0[1] = q0[1]; s_q0[2] = q0[2]; }   // a0 == 0 here: actual order
        }

        LocalStat run;
        run.reset();

        __syncthreads();   // the previous unit's pack phase has released the staging area

        // ---- phase 1: one read of the sub-cell rows ----
        // A batch = UNROLL float4 per thread.  The fast quantiser is trusted only for pixel
        // indices in [1, pixels): the batch minimum and maximum tell whether every element
        // qualified; if not (rare) the batch is redone with the IEEE divide.
        auto load = [&](int p0, float4 (&v)[UNROLL]) {
            const unsigned rowg0 = rank * ROWS + rsub + RPP * p0;
            const float4 *p = pbase + ((rowg0 / NSUB) * plane4 + (rowg0 % NSUB) * row4);
            const unsigned step = RPP % NSUB == 0 ? (RPP / NSUB) * plane4 : RPP * row4;   // float4 units per pass
#pragma unroll
            for (int u = 0; u < UNROLL; u++) v[u] = ld_stream_pinned(p + (size_t)u * step);
        };
        auto batch = [&](auto exact_tag, int p0, const float4 (&v)[UNROLL], LocalStat &ls) {
            constexpr bool EXACT = decltype(exact_tag)::value;
#pragma unroll
            for (int u = 0; u < UNROLL; u += 2) {
                int q[2][4];
                unsigned w[2][4];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const float x[4] = {v[u + h].x, v[u + h].y, v[u + h].z, v[u + h].w};
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const int j = c % 3;
                        if constexpr (EXACT) q[h][c] = quantize_rare(x[c], low[j], -ndx[j], P[j], oob);
                        else q[h][c] = quantize_fast(x[c], low[j], rcp[j], ndx[j]);
                        const unsigned t = (unsigned)q[h][c] + C[j];
                        w[h][c] = min(t, t - (unsigned)P[j]);
                        stage[soff[c] + (p0 + u + h) * (RPP * NSUB)] = (unsigned short)w[h][c];
                    }
                }
                // relative axis 0 owns floats 0 and 3 of each float4, axes 1 and 2 one float each
                ls.qmin[0] = __vimin3_s32(ls.qmin[0], q[0][0], q[0][3]); ls.qmin[0] = __vimin3_s32(ls.qmin[0], q[1][0], q[1][3]);
                ls.qmax[0] = __vimax3_s32(ls.qmax[0], q[0][0], q[0][3]); ls.qmax[0] = __vimax3_s32(ls.qmax[0], q[1][0], q[1][3]);
                ls.wmin[0] = __vimin3_u32(ls.wmin[0], w[0][0], w[0][3]); ls.wmin[0] = __vimin3_u32(ls.wmin[0], w[1][0], w[1][3]);
                ls.wmax[0] = __vimax3_u32(ls.wmax[0], w[0][0], w[0][3]); ls.wmax[0] = __vimax3_u32(ls.wmax[0], w[1][0], w[1][3]);
#pragma unroll
                for (int j = 1; j < 3; j++) {
                    ls.qmin[j] = __vimin3_s32(ls.qmin[j], q[0][j], q[1][j]);
                    ls.qmax[j] = __vimax3_s32(ls.qmax[j], q[0][j], q[1][j]);
                    ls.wmin[j] = __vimin3_u32(ls.wmin[j], w[0][j], w[1][j]);
                    ls.wmax[j] = __vimax3_u32(ls.wmax[j], w[0][j], w[1][j]);
                }
            }
        };
        auto step = [&](int p0, float4 (&v)[UNROLL]) {
            LocalStat ls;
            ls.reset();
            batch(std::false_type{}, p0, v, ls);
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 3; j++)
                ok = ok && (unsigned)(ls.qmin[j] - 1) < Pm1[j] && (unsigned)(ls.qmax[j] - 1) < Pm1[j];
            if (!ok) {
                ls.reset();
                load(p0, v);
                batch(std::true_type{}, p0, v, ls);
            }
#pragma unroll
            for (int j = 0; j < 3; j++) {
                run.wmin[j] = min(run.wmin[j], ls.wmin[j]); run.wmax[j] = max(run.wmax[j], ls.wmax[j]);
                run.qmin[j] = min(run.qmin[j], ls.qmin[j]); run.qmax[j] = max(run.qmax[j], ls.qmax[j]);
            }
        };
        if constexpr (PIPE) {
            // two register buffers, ping-pong: the loads of batch i+1 are in flight while batch i is quantised
            static_assert(!PIPE || PASSES % (2 * UNROLL) == 0, "pairs of batches");
            float4 va[UNROLL], vb[UNROLL];
            load(0, va);
#pragma unroll 1
            for (int p0 = 0; p0 < PASSES; p0 += 2 * UNROLL) {
                load(p0 + UNROLL, vb);
                step(p0, va);
                if (p0 + 2 * UNROLL < PASSES) load(p0 + 2 * UNROLL, va);
                step(p0 + UNROLL, vb);
            }
        } else {
            float4 v[UNROLL];
#pragma unroll 1
            for (int p0 = 0; p0 < PASSES; p0 += UNROLL) {
                load(p0, v);
                step(p0, v);
            }
        }

        // ---- CTA reduction, in actual axis order ----
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int j = (k - a0 + 3) % 3;   // relative index of actual axis k
            unsigned a = j == 0 ? run.wmin[0] : (j == 1 ? run.wmin[1] : run.wmin[2]);
            unsigned b = j == 0 ? run.wmax[0] : (j == 1 ? run.wmax[1] : run.wmax[2]);
            int c = j == 0 ? run.qmin[0] : (j == 1 ? run.qmin[1] : run.qmin[2]);
            int d = j == 0 ? run.qmax[0] : (j == 1 ? run.qmax[1] : run.qmax[2]);
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
            const long long u = claim_unit(A);
            if constexpr (CS > 1) {
                for (unsigned r = 0; r < CS; r++) *cg::this_cluster().map_shared_rank(&s_unit[par ^ 1], r) = u;
            } else {
                s_unit[par ^ 1] = u;
            }
        }
        cluster_sync_all<CS>();

        // ---- pull the next unit's rows towards L2 while this one is being packed ----
        if (A.prefetch) {
            const long long nu = s_unit[par ^ 1];
            if (nu < A.nunits) {
                const long long nf = nu / A.sc3, nsc = nu - nf * A.sc3;
                const unsigned nx0 = NSUB * (unsigned)(nsc % S), ny0 = NSUB * (unsigned)((nsc / S) % S), nz0 = NSUB * (unsigned)(nsc / ((long long)S * S));
                const float4 *nb = (const float4 *)(A.aos + 3 * nf * (long long)nfile * nfile * nfile) + (3u * nx0 / 4u + ny0 * row4 + nz0 * plane4);
                for (unsigned r = tid; r < (unsigned)ROWS; r += NT) {
                    const unsigned rowg = rank * ROWS + r;
                    const float4 *p = nb + ((rowg / NSUB) * plane4 + (rowg % NSUB) * row4);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(NSUB * 12) : "memory");
                }
            }
        }

        // ---- finalise: warp k combines the cluster's statistics of axis k (every CTA, redundantly) ----
        long long f_nbytes = 0, f_b = 0;   // warps 0..2 only
        if (warp < 3) {
            const int k = warp;
            f_b = f * 3 * A.sc3 + k * A.sc3 + sc;                    // block id in the batch
            XStat x;
            x.wmin = ~0u; x.wmax = 0u; x.qmin = INT_MAX; x.qmax = INT_MIN; x.oob = 0;
            if (lane < CS) x = s_x[par][lane][k];
            x.wmin = __reduce_min_sync(0xffffffffu, x.wmin); x.wmax = __reduce_max_sync(0xffffffffu, x.wmax);
            x.qmin = __reduce_min_sync(0xffffffffu, x.qmin); x.qmax = __reduce_max_sync(0xffffffffu, x.qmax);
            x.oob = __reduce_or_sync(0xffffffffu, x.oob);
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
                base = x.wmin; padj = 0;
            }
            // bit.PrecisionNeeded: below 2^48 Go's float64 log2 agrees with the integer bit length
            int bits = maxoff < (1ULL << 48) ? 64 - __clzll((long long)maxoff) : precision_needed(maxoff);
            long long nbytes = array_bytes(bits, N);
            const bool slow = x.oob != 0;
            if (slow) { bits = 0; nbytes = 0; }
            f_nbytes = nbytes;
            if (rank == 0 && lane == 0) st_relaxed(A.W.pub + f_b, PUB_AGG | (unsigned long long)nbytes);
            // staged values are the low 16 bits of w: enough when the packed value has <= 16 bits
            // and (wide arcs) w itself fits, i.e. pixels <= 65536
            const int mode = (bits >= 1 && bits <= 16 && !(wide && Pk > 65536)) ? 1 : 0;
            if (lane == 0) {
                Fin fin;
                fin.off = 0; fin.bits = bits; fin.mode = mode; fin.base = base; fin.padj = padj;
                s_fin[k] = fin;
                if (rank == 0) {
                    if (!slow && bits > 0 && mode == 0) A.W.repack_list[atomicAdd(A.W.repack_count, 1)] = f_b;
                    if (slow) atomicExch(A.W.abort_flag, 1);
                    BlockStat st = {};
                    st.pmin = pmin; st.min = mn; st.nbytes = nbytes; st.out_off = 0; st.do_bound = 1; st.bits = bits;
                    st.q0 = q0k; st.oob = x.oob;
                    A.stats[f_b] = st;
                    if (A.mins) A.mins[f_b] = mn;
                    if (A.bits) A.bits[f_b] = bits;
                }
            }
        }
        if (tid == 0) s_gctr = 0;
        __syncthreads();

        // ---- byte offsets: warps 0..2 walk back over the earlier sub-cells of the group while the
        // other warps already pack; a group is written out once its block's offset is posted ----
        if (warp < 3) {
            const int k = warp;
            long long off = lookback(A.W.pub, f_b - sc, f_b);
            if (lane == 0) {
                if (off + f_nbytes > A.axis_stride) {   // never write past the caller's buffer
                    if (rank == 0) atomicExch(A.W.err, 2);
                    s_off[k] = -1;
                } else {
                    s_off[k] = off;
                }
                if (rank == 0) {
                    st_relaxed(A.W.pub + f_b, PUB_PREFIX | (unsigned long long)(off + f_nbytes));
                    A.stats[f_b].out_off = off;
                    if (A.offsets) A.offsets[f_b] = off;
                    if (A.out_len && sc == A.sc3 - 1) A.out_len[f * 3 + k] = off + f_nbytes;
                }
                __threadfence_block();
                *(volatile int *)&s_offgen[k] = gen;
            }
        }

        // ---- phase 2: pack groups of 1024 elements straight from shared memory ----
        for (;;) {
            int g = 0;
            if (lane == 0) g = atomicAdd(&s_gctr, 1);
            g = __shfl_sync(0xffffffffu, g, 0);
            if (g >= 3 * GPA) break;
            const int k = g / GPA, gi = g - k * GPA;
            const Fin fin = s_fin[k];
            if (fin.mode == 0) continue;
            const int eb = gi * 1024 + 32 * lane;   // this lane's first element within the CTA's chunk
            const int sw = (eb >> 6) & 7;
            unsigned v[32];
            uint4 r[4];
#pragma unroll
            for (int s = 0; s < 4; s++) r[s] = *(const uint4 *)(stage + k * CHUNK + ((((eb >> 3) + s) ^ sw) << 3));
            if (fin.padj == 0) {   // narrow arc: v = w - wmin, exact modulo 2^16
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    const unsigned rr[4] = {r[s].x, r[s].y, r[s].z, r[s].w};
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        v[8 * s + 2 * t] = (rr[t] - fin.base) & 0xffffu;
                        v[8 * s + 2 * t + 1] = ((rr[t] >> 16) - fin.base) & 0xffffu;
                    }
                }
            } else {               // wide arc (pixels <= 65536): v = (w - C - qmin) mod pixels
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    const unsigned rr[4] = {r[s].x, r[s].y, r[s].z, r[s].w};
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const unsigned lo = (rr[t] & 0xffffu) - fin.base, hi = (rr[t] >> 16) - fin.base;
                        v[8 * s + 2 * t] = min(lo, lo + fin.padj);
                        v[8 * s + 2 * t + 1] = min(hi, hi + fin.padj);
                    }
                }
            }
            unsigned *region = (unsigned *)(stage + k * CHUNK + gi * 1024);
            const long long e0 = (long long)rank * CHUNK + (long long)gi * 1024;   // element index in the block
            uint8_t *dst0 = A.out + (f * 3 + k) * A.axis_stride + ((e0 * fin.bits) >> 3);
            switch (fin.bits) {
#define MNW_CASE(B) case B: pack_group_words<B>(v, region, lane, dst0, &s_off[k], &s_offgen[k], gen); break;
                MNW_CASE(1) MNW_CASE(2) MNW_CASE(3) MNW_CASE(4) MNW_CASE(5) MNW_CASE(6) MNW_CASE(7) MNW_CASE(8)
                MNW_CASE(9) MNW_CASE(10) MNW_CASE(11) MNW_CASE(12) MNW_CASE(13) MNW_CASE(14) MNW_CASE(15) MNW_CASE(16)
#undef MNW_CASE
                default: break;
            }
        }
        par ^= 1;
        gen++;
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
    unsigned long long seed, block_id0;
    int nfile, subcells;
    long long sc3, nslabs;
    float *out;
    int sync_mode;          // tuning knob MNW_DEC_SYNC: 1 = CTA barrier per slab, 0 = per-stage empty barriers
};

// What a CTA needs to know about one axis block of the slab it is about to decode
// (written by the producer thread together with the bulk copy of the packed bytes).
struct SlabAxis {
    const uint8_t *gsrc;   // slow path: first byte of the block's stream
    long long mn;
    long long pixels;
    float low, dx;
    unsigned key;          // jitter key of the block
    unsigned mask;         // fast path: (1 << bits) - 1
    int bits;
    int shift;             // fast path: bit position of the slab's first value in the staged bytes
    int fast;              // 1: 32-bit path from shared memory
    int periodic;
};
struct SlabInfo {
    long long out4;        // float4 index (in the output array) of row 0 of the slab, column 0
    unsigned e0;           // element index of the slab's first element within its blocks
    int fast;              // all three axes take the fast path
};

// Slow lane path of the decoder: any bit width, 64-bit arithmetic, straight from global memory.
__device__ __noinline__ float decode_rare(const SlabAxis &h, long long e, int jmode, float wrap_L) {
    const uint8_t *stream = h.gsrc;
    const int a = (int)((uintptr_t)stream & 3);
    const uint32_t *base = (const uint32_t *)(stream - a);
    unsigned long long v = 0;
    const int bits = h.bits;
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
    long long q = (long long)((unsigned long long)h.mn + v);            // go/group.go:262
    if (h.periodic) q = bound1(q, 0, h.pixels);                          // :303
    double u = 0.5;
    if (jmode == 1) u = (double)(jitter_hash_keyed(h.key, (uint32_t)e) >> 8) * 0x1p-24;
    const float t = __double2float_rn(__dadd_rn(__ll2double_rn(q), u));  // :308
    float o = __fadd_rn(__fmul_rn(h.dx, t), h.low);
    if (wrap_L > 0.0f) {                                                 // go/minp/minp.go:195-203
        if (o < 0.0f) o = __fadd_rn(o, wrap_L);
        else if (o >= wrap_L) o = __fsub_rn(o, wrap_L);
    }
    return o;
}

// Persistent CTAs walk the slabs (SLAB consecutive elements of the three axis blocks of
// one sub-cell).  Thread 0 fetches the packed bytes of the NEXT slab with three TMA bulk
// copies (cp.async.bulk, completion on an mbarrier) and works out everything that is
// uniform over the slab, while the CTA decodes the current one; every thread then
// produces whole float4 pieces of AoS rows, so each output row is written once with
// coalesced 128-bit stores and no axis ever touches a sector alone.
// WRAP: 0 no periodic wrap, 1 both tests of go/minp/minp.go:195-203, 2 only `x >= L` (every group has
// low >= +0, so no decoded value is negative).
template <int NSUB, int NT, int MINB, bool HASH, int WRAP>
__global__ void __launch_bounds__(NT, MINB) k_decode_vec3(const DecVec3Args A) {
    constexpr int N = NSUB * NSUB * NSUB;
    constexpr int SLAB = N < 4096 ? N : 4096;   // elements per slab and axis
    constexpr int SLABS = N / SLAB;
    constexpr int ROWS = SLAB / NSUB;
    constexpr int R4 = 3 * NSUB / 4;
    constexpr int RPP = NT / R4;
    constexpr int STAGE_BYTES = SLAB * 3 + 32;   // up to 24 bits per value, + alignment slack
    static_assert(NT % R4 == 0, "threads tile the rows exactly");
    extern __shared__ __align__(128) unsigned char dsm[];   // [2][3][STAGE_BYTES]
    __shared__ __align__(8) unsigned long long s_bar[2], s_empty[2];
    __shared__ SlabAxis s_hdr[2][3];
    __shared__ SlabInfo s_info[2];

    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned S = (unsigned)A.subcells, nfile = (unsigned)A.nfile, sc3 = (unsigned)A.sc3;
    const unsigned row4 = 3u * nfile / 4u, plane4 = row4 * nfile;
    const int col4 = tid % R4, rsub = tid / R4, a0 = col4 % 3;
    int ecol[4];
#pragma unroll
    for (int c = 0; c < 4; c++) ecol[c] = (4 * col4 + c) / 3;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_empty[0])), "r"(NW));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_empty[1])), "r"(NW));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // producer (one whole warp calls it; lane k < 3 describes axis k, so the dependent global loads of the three axes
    // overlap): describe slab g and start the copies of its packed bytes into stage st
    auto issue = [&](int st, unsigned g) {
        const unsigned unit = g / SLABS, slab = g % SLABS;
        const unsigned f = unit / sc3, sc = unit % sc3;
        unsigned nbytes = 0;
        const uint8_t *src = nullptr;
        int fast = 1;
        if (lane < 3) {
            const int k = lane;
            const long long b = ((long long)f * 3 + k) * sc3 + sc;
            const FloatParams fp = A.tab[(A.tab_per_file ? 3 * f : 0) + k];
            SlabAxis h;
            h.mn = A.mins[b]; h.bits = (int)A.bits[b]; h.pixels = fp.pixels; h.low = fp.low; h.dx = fp.dx;
            h.periodic = (fp.flags & F_PERIODIC) ? 1 : 0;
            h.key = jitter_key(A.seed, A.block_id0 + (unsigned long long)b);
            h.gsrc = A.data + ((long long)f * 3 + k) * A.stream_len + A.offsets[b];
            h.mask = (h.bits >= 1 && h.bits <= 24) ? ((1u << h.bits) - 1u) : 0u;
            // 32-bit path: q = mn + v lies in [0, 2*pixels) (periodic) or [0, 2^23), and float32 holds it exactly
            h.fast = h.bits >= 0 && h.bits <= 24 && h.mn >= 0 && fp.pixels > 0 && fp.pixels < (1LL << 23) &&
                     (h.periodic ? h.mn + (long long)h.mask < 2 * fp.pixels : h.mn + (long long)h.mask < (1LL << 23));
            h.shift = 0;
            if (h.fast && h.bits > 0) {
                const uint8_t *p = h.gsrc + (((long long)slab * SLAB * h.bits) >> 3);
                const unsigned a16 = (unsigned)((uintptr_t)p & 15);
                src = p - a16;
                nbytes = (a16 + (unsigned)(SLAB * h.bits / 8) + 15u) & ~15u;
                h.shift = 8 * (int)a16;
            }
            fast = h.fast;
            s_hdr[st][k] = h;
        }
        const unsigned total = nbytes + __shfl_down_sync(0xffffffffu, nbytes, 1) + __shfl_down_sync(0xffffffffu, nbytes, 2);   // (lane 0's)
        const int all_fast = __all_sync(0xffffffffu, fast);
        const unsigned bar = smem_u32(&s_bar[st]);
        if (lane == 0) {
            SlabInfo info;
            const unsigned ix0 = NSUB * (sc % S), iy0 = NSUB * ((sc / S) % S), iz0 = NSUB * (sc / (S * S));
            const unsigned row0 = slab * ROWS;   // first sub-cell row of the slab
            info.out4 = (long long)f * ((long long)plane4 * nfile) +
                        (long long)(3u * ix0 / 4u) + (long long)(iy0 + row0 % NSUB) * row4 + (long long)(iz0 + row0 / NSUB) * plane4;
            info.e0 = slab * SLAB;
            info.fast = all_fast;
            s_info[st] = info;
        }
        __syncwarp();   // the headers of lanes 1 and 2 are ordered before lane 0's arrival, which publishes them
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
        __syncwarp();
        if (nbytes)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(dsm + (st * 3 + lane) * STAGE_BYTES)), "l"(src), "r"(nbytes), "r"(bar) : "memory");
    };

    const unsigned nslabs = (unsigned)A.nslabs;
    if (warp == 0 && blockIdx.x < nslabs) issue(0, blockIdx.x);
    __syncthreads();

    // rows of the slab handled by this thread are rl = rsub + RPP*i: offset of row rl from the slab's row 0
    // (a slab is one z-plane of the sub-cell when NSUB = 64, several planes otherwise)
    // No CTA-wide barrier in the loop: a stage is handed back through s_empty (one arrival per warp), and the warp whose
    // turn it is to fetch the next slab (a different one every iteration: describing a slab costs a few dependent global
    // loads) is the only one that waits for the others -- for the iteration BEFORE the one they are working on.
    int it = 0;
    for (unsigned g = blockIdx.x; g < nslabs; g += gridDim.x, it++) {
        const int st = it & 1;
        if (A.sync_mode) {
            if (warp == 0 && g + gridDim.x < nslabs) issue(st ^ 1, g + gridDim.x);
        } else if (g + gridDim.x < nslabs && warp == it % NW) {
            if (it >= 1) {   // stage st ^ 1 was read in iteration it - 1: phase (it - 1) / 2 of its empty barrier
                const unsigned bar = smem_u32(&s_empty[st ^ 1]), parity = ((it - 1) >> 1) & 1;
                unsigned done = 0;
                while (!done)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            }
            issue(st ^ 1, g + gridDim.x);
        }
        {   // wait for this slab's bytes
            const unsigned bar = smem_u32(&s_bar[st]), parity = (it >> 1) & 1;
            unsigned done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        }
        const SlabInfo info = s_info[st];
        float4 *pbase = (float4 *)A.out + info.out4 + col4;

        if (info.fast) {
            // per-axis constants in this thread's axis order
            unsigned buf[3], bits[3], shift[3], mn[3], P[3], mask[3], key[3];
            float low[3], dx[3];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int k = (a0 + j) % 3;
                const SlabAxis &h = s_hdr[st][k];
                buf[j] = smem_u32(dsm + (st * 3 + k) * STAGE_BYTES);
                bits[j] = (unsigned)h.bits; shift[j] = (unsigned)h.shift; mn[j] = (unsigned)h.mn;
                P[j] = h.periodic ? (unsigned)h.pixels : 0u;
                mask[j] = h.mask; key[j] = h.key + info.e0 * 0x9E3779B1U;   // hash(key, e0 + el) = f(el * M + key')
                low[j] = h.low; dx[j] = h.dx;
            }
            // From one of this thread's rows to the next (RPP rows on) the bit position of its four values moves by
            // RPP * NSUB * bits, a whole number of 32-bit words: the shift within the word never changes and the word
            // address, the hash argument and the output pointer all advance by constants.
            unsigned addr[4], sh[4], hx[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int j = c % 3;
                const unsigned el0 = (unsigned)(rsub * NSUB + ecol[c]);
                const unsigned bp0 = shift[j] + el0 * bits[j];
                addr[c] = buf[j] + ((bp0 >> 5) << 2);
                sh[c] = bp0 & 31u;
                hx[c] = el0 * 0x9E3779B1U + key[j];
            }
            const unsigned dA[3] = {(RPP * NSUB / 8) * bits[0], (RPP * NSUB / 8) * bits[1], (RPP * NSUB / 8) * bits[2]};
#pragma unroll 2
            for (int rl = rsub; rl < ROWS; rl += RPP) {
                float o[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int j = c % 3;
                    unsigned w0, w1;
                    asm("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(addr[c]));
                    asm("ld.shared.b32 %0, [%1+4];" : "=r"(w1) : "r"(addr[c]));
                    addr[c] += dA[j];
                    const unsigned v = __funnelshift_r(w0, w1, sh[c]) & mask[j];    // Array.Slice, go/bit/bit.go:29-82
                    unsigned q = mn[j] + v;                                          // go/group.go:262
                    q = min(q, q - P[j]);                                            // bound(q, 0, pixels), :303 (P = 0: not periodic)
                    float t;
                    if constexpr (HASH) {
                        // u = h24 * 2^-24 is exact in float32 and q + u needs at most 47 bits: the FMA
                        // rounds once, exactly like float32(float64(q) + u) (go/group.go:308)
                        unsigned x = hx[c];                                          // = el * M + key
                        hx[c] += (unsigned)(RPP * NSUB) * 0x9E3779B1U;
                        x ^= x >> 16; x *= 0x7feb352dU;
                        x ^= x >> 15; x *= 0x846ca68bU;
                        t = __fmaf_rn((float)(x >> 8), 0x1p-24f, (float)q);
                    } else {
                        t = __fadd_rn((float)q, 0.5f);   // q < 2^23: exact
                    }
                    // (scalar on purpose: the packed f32x2 forms cost as many register moves as they save)
                    float x = __fadd_rn(__fmul_rn(dx[j], t), low[j]);
                    if constexpr (WRAP == 1) {                                       // go/minp/minp.go:195-203
                        const float xp = __fadd_rn(x, A.wrap_L), xm = __fsub_rn(x, A.wrap_L);
                        x = x < 0.0f ? xp : (x >= A.wrap_L ? xm : x);
                    } else if constexpr (WRAP == 2) {
                        const float xm = __fsub_rn(x, A.wrap_L);
                        x = x >= A.wrap_L ? xm : x;
                    }
                    o[c] = x;
                }
                __stcs(pbase + ((unsigned)(rl / NSUB) * plane4 + (unsigned)(rl % NSUB) * row4), make_float4(o[0], o[1], o[2], o[3]));   // setSubCell, :270-288
            }
        } else {
            for (int rl = rsub; rl < ROWS; rl += RPP) {
                float o[4];
                for (int c = 0; c < 4; c++) {
                    const int k = (a0 + c) % 3;
                    const long long e = (long long)info.e0 + rl * NSUB + ecol[c];
                    o[c] = decode_rare(s_hdr[st][k], e, HASH ? 1 : 0, WRAP != 0 ? A.wrap_L : 0.0f);
                }
                __stcs(pbase + ((unsigned)(rl / NSUB) * plane4 + (unsigned)(rl % NSUB) * row4), make_float4(o[0], o[1], o[2], o[3]));
            }
        }
        if (A.sync_mode) {
            __syncthreads();
        } else {
            __syncwarp();      // stage st and its header may be refilled once every warp has said so
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_empty[st])) : "memory");
        }
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
        // the pair / float4 form of k_pipe_vec3 and the group kernels: floor by RM(y + 2^23), accepted when the raw
        // bits lie in [2^23, 2^23 + pixels) -- no test on the offset at all
        if (Pm1 && P <= (1 << 22)) {
            const float t = __fsub_rn(x, fp.low);
            float y = __fmul_rn(t, fp.rcp);
            float e = __fmaf_rn(-fp.dx, y, t);
            y = __fmaf_rn(e, fp.rcp, y);
            e = __fmaf_rn(-fp.dx, y, t);
            y = __fmaf_rn(e, fp.rcp, y);
            const unsigned q2 = __float_as_uint(__fadd_rd(y, 8388608.0f)) - 0x4B000000u;
            if (q2 < (unsigned)P && (long long)q2 != quantize_exact(x, fp.low, fp.dx)) bad++;
        }
        // quantize_fast (F2I form): offset in [+0, high - low] and pixel index < pixels
        if (Pm1 && __float_as_uint(__fsub_rn(x, fp.low)) <= __float_as_uint(__fsub_rn(fp.high, fp.low)) &&
            (unsigned)qi < (unsigned)P) {
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

// go_log10_f32 (table + series, exact fallback) against float32(go_log10(float64 x)) over float bit patterns
__global__ void k_selftest_log10(unsigned long long first, unsigned long long count, unsigned long long *mismatches, int pow10) {
    unsigned long long bad = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < count;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)(first + i));
        const unsigned a = __float_as_uint(pow10 ? go_pow10_f32(x) : go_log10_f32(x));
        const unsigned b = __float_as_uint(__double2float_rn(pow10 ? go_pow10((double)x) : go_log10((double)x)));
        if (a != b && !((a & 0x7fffffffu) > 0x7f800000u && (b & 0x7fffffffu) > 0x7f800000u)) bad++;   // (any NaN equals any NaN)
    }
    for (int o = 16; o; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, bad);
}
void launch_selftest_log10(Launcher &L, unsigned long long first, unsigned long long count, unsigned long long *d_out, int pow10) {
    cudaMemsetAsync(d_out, 0, 8, L.stream);
    if (count == 0) return;
    k_selftest_log10<<<148 * 8, 256, 0, L.stream>>>(first, count, d_out, pow10);
    L.count++;
}

// float32(math.Pow(10, float64(x))) of a raw float32 column (minh Log columns stored as Float32Group, go/minh/minh.go:315-319)
__global__ void k_pow10_f32(const float *x, long long n, float *out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = go_pow10_f32(x[i]);
}
void launch_pow10_f32(Launcher &L, const float *x, long long n, float *out) {
    if (n == 0) return;
    long long g = (n + 255) / 256;
    k_pow10_f32<<<(unsigned)(g < 148 * 16 ? g : 148 * 16), 256, 0, L.stream>>>(x, n, out);
    L.count++;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
size_t fused_work_bytes(int64_t nblocks) { return (size_t)(16 * nblocks + 256); }

bool fused_vec3_supported(const FloatParamsHost *fp, int64_t nparams, int nfile, int subcells, const void *aos) {
    if (subcells <= 0 || nfile % subcells) return false;
    const int nsub = nfile / subcells;
    if (nsub != 16 && nsub != 32 && nsub != 64 && nsub != 128) return false;
    if (((uintptr_t)aos & 15) != 0 || nfile > 1024) return false;   // 32-bit float4 offsets inside a file
    for (int64_t i = 0; i < nparams; i++) {
        const FloatParamsHost &p = fp[i];
        if (!(p.flags & F_PERIODIC) || (p.flags & (F_LOG10 | F_CLAMP))) return false;
        if (p.pixels < 1 || p.pixels >= (1LL << 30)) return false;
        if (nsub == 128 && p.pixels > (1LL << 22)) return false;   // 128^3 sub-cells: k_pipe_vec3 only
    }
    return true;
}

template <int NSUB, int CS, int NT, int UNROLL, int MINB, bool PIPE>
static cudaError_t launch_fused_vec3_t(Launcher &L, const FusedArgs &A) {
    auto kern = k_fused_vec3<NSUB, CS, NT, UNROLL, MINB, PIPE>;
    const size_t smem = (size_t)6 * (NSUB * NSUB * NSUB / CS);
    static DevCfg cfgs[MNW_MAX_DEVICES];
    DevCfg &dc = dev_cfg(cfgs);
    cudaError_t e;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = L.stream;
    cfg.attrs = attr; cfg.numAttrs = 1;
    std::call_once(dc.once, [&] {
        dc.err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (dc.err != cudaSuccess) return;
        if (CS > 8) {
            dc.err = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (dc.err != cudaSuccess) return;
        }
        cfg.gridDim = dim3(CS);
        dc.err = cudaOccupancyMaxActiveClusters(&dc.a, kern, &cfg);
        if (dc.err == cudaSuccess && dc.a < 1) dc.err = cudaErrorLaunchOutOfResources;
        if (getenv("MNW_DEBUG")) fprintf(stderr, "k_fused_vec3<%d,%d,%d>: %d co-resident clusters, %zu B dynamic smem\n", NSUB, CS, NT, dc.a, smem);
    });
    if (dc.err != cudaSuccess) return dc.err;
    const int max_clusters = dc.a;
    long long clusters = A.nunits < max_clusters ? A.nunits : max_clusters;
    cfg.gridDim = dim3((unsigned)(clusters * CS));
    L.begin("k_fused_vec3");
    e = cudaLaunchKernelEx(&cfg, kern, A);
    L.end();
    L.count++;
    return e;
}

cudaError_t launch_pipe_vec3(Launcher &L, const FusedArgs &A);   // kernels_pipe.cu
cudaError_t launch_pipe_vec3_coop(Launcher &L, FusedArgs A, void *ws, int nsub);

cudaError_t launch_fused_vec3(Launcher &L, const FusedWork &W, const BlockDesc *descs, const FloatParams *tab, int tab_per_file,
                              const float *aos, int nfile, int subcells, int64_t nfiles, BlockStat *stats,
                              int64_t *mins, int64_t *bits, int64_t *offsets, int64_t *out_len, uint8_t *out,
                              int64_t out_axis_stride, bool pipe_ok, void *coop_ws, bool *ran_pipe) {
    if (ran_pipe) *ran_pipe = false;
    FusedArgs A = {};
    A.descs = descs;
    A.aos = aos; A.tab = tab; A.tab_per_file = tab_per_file; A.nfile = nfile; A.subcells = subcells;
    A.sc3 = (long long)subcells * subcells * subcells;
    A.nunits = nfiles * A.sc3;
    A.stats = stats; A.mins = mins; A.bits = bits; A.offsets = offsets; A.out_len = out_len; A.out = out;
    A.axis_stride = out_axis_stride; A.W = W;
    {   // chunk of the ticket order (claim_unit)
        static const int knob = getenv("MNW_TICKET_CHUNK") ? atoi(getenv("MNW_TICKET_CHUNK")) : 0;   // tuning knob
        long long G = knob > 0 ? knob : 1;   // measured best on 16^3 and 32^3 sub-cells (tools/bench_subcells.py): plain round-robin
        if (G < 1 || A.sc3 % G != 0) G = 1;
        A.ticket_chunk = (int)G;
    }
    static const int prefetch = getenv("MNW_PREFETCH") ? atoi(getenv("MNW_PREFETCH")) : 1;   // tuning knob
    A.prefetch = prefetch;
    if (A.nunits == 0) return cudaSuccess;
    switch (nfile / subcells) {
        case 64: {
            // tuning knob: 384 (default) / 385 / 768 = 8-CTA cluster, 1 CTA per SM with 384 threads
            // (pipelined loads / 8-deep batches) or 768 threads; 0 = 16-CTA cluster, 2 CTAs per SM
            static int variant = getenv("MNW_FUSED_NT") ? atoi(getenv("MNW_FUSED_NT")) : 1;
            if (variant == 1) {   // warp-specialised pipeline (default); needs pixels <= 2^22
                // cooperative, cluster-free schedule on all SMs when the caller brought a workspace (MNW_PIPE=cluster
                // keeps the 8-CTA cluster schedule); it falls back to the clusters when the device refuses the launch
                static const bool coop = !(getenv("MNW_PIPE") && !strcmp(getenv("MNW_PIPE"), "cluster"));
                if (pipe_ok && coop && coop_ws) {
                    const cudaError_t e = launch_pipe_vec3_coop(L, A, coop_ws, 64);
                    if (e == cudaSuccess) { if (ran_pipe) *ran_pipe = true; return e; }
                    (void)cudaGetLastError();
                }
                if (pipe_ok) { if (ran_pipe) *ran_pipe = true; return launch_pipe_vec3(L, A); }
                return launch_fused_vec3_t<64, 8, 384, 4, 1, true>(L, A);
            }
            if (variant == 0) {
                cudaError_t e = launch_fused_vec3_t<64, 16, 384, 4, 2, false>(L, A);
                if (e == cudaSuccess) return e;
                (void)cudaGetLastError();   // 16-CTA clusters are a non-portable size: fall back to 8
                variant = 384;
            }
            if (variant == 768) return launch_fused_vec3_t<64, 8, 768, 4, 1, false>(L, A);
            if (variant == 385) return launch_fused_vec3_t<64, 8, 384, 8, 1, false>(L, A);
            return launch_fused_vec3_t<64, 8, 384, 4, 1, true>(L, A);
        }
        case 32: {
            // a 32^3 sub-cell is one CTA's worth of the pipeline (k_pipe_vec3<COOP, 32>, one part per unit)
            static const bool pipe32 = !(getenv("MNW_PIPE") && !strcmp(getenv("MNW_PIPE"), "cluster"));
            if (pipe_ok && pipe32 && coop_ws) {
                const cudaError_t e = launch_pipe_vec3_coop(L, A, coop_ws, 32);
                if (e == cudaSuccess) { if (ran_pipe) *ran_pipe = true; return e; }
                (void)cudaGetLastError();
            }
            return launch_fused_vec3_t<32, 1, 768, 4, 1, false>(L, A);
        }
        case 128: {   // 64 CTAs per unit: the cooperative pipeline or nothing (the caller then takes the generic kernels)
            if (pipe_ok && coop_ws) {
                const cudaError_t e = launch_pipe_vec3_coop(L, A, coop_ws, 128);
                if (e == cudaSuccess) { if (ran_pipe) *ran_pipe = true; return e; }
                (void)cudaGetLastError();
            }
            return cudaErrorNotSupported;
        }
        case 16: {
            // a 16^3 unit stages 24 KB only: several CTAs per SM hide each other's read -> barrier -> pack phases
            static const int minb = getenv("MNW_FUSED16_MINB") ? atoi(getenv("MNW_FUSED16_MINB")) : 2;   // tuning knob (measured best with round-robin tickets)
            if (minb == 2) return launch_fused_vec3_t<16, 1, 384, 4, 2, true>(L, A);
            if (minb == 3) return launch_fused_vec3_t<16, 1, 384, 4, 3, false>(L, A);
            if (minb == 4) return launch_fused_vec3_t<16, 1, 384, 2, 4, false>(L, A);
            return launch_fused_vec3_t<16, 1, 384, 4, 1, true>(L, A);
        }
    }
    return cudaErrorNotSupported;
}

bool fused_decode_vec3_supported(int nfile, int subcells, const void *aos_out) {
    if (subcells <= 0 || nfile % subcells || nfile > 1024) return false;
    const int nsub = nfile / subcells;
    if (nsub != 16 && nsub != 32 && nsub != 64 && nsub != 128) return false;
    return ((uintptr_t)aos_out & 15) == 0;
}

template <int NSUB, bool HASH, int WRAP>
static cudaError_t launch_decode_vec3_t(Launcher &L, const DecVec3Args &A) {
    constexpr int NT = 384, MINB = 3;
    constexpr int N = NSUB * NSUB * NSUB, SLAB = N < 4096 ? N : 4096;
    constexpr size_t smem = (size_t)2 * 3 * (SLAB * 3 + 32);
    auto kern = k_decode_vec3<NSUB, NT, MINB, HASH, WRAP>;
    static DevCfg cfgs[MNW_MAX_DEVICES];
    DevCfg &dc = dev_cfg(cfgs);
    std::call_once(dc.once, [&] {
        dc.err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (dc.err != cudaSuccess) return;
        int per = 0;
        dc.err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, NT, smem);
        if (dc.err != cudaSuccess) return;
        if (per < 1) { dc.err = cudaErrorLaunchOutOfResources; return; }
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        dc.a = per * sms;
        if (getenv("MNW_DEBUG")) fprintf(stderr, "k_decode_vec3<%d>: %d co-resident CTAs, %zu B dynamic smem\n", NSUB, dc.a, smem);
    });
    if (dc.err != cudaSuccess) return dc.err;
    const int per_sm = dc.a;
    const long long grid = A.nslabs < per_sm ? A.nslabs : per_sm;
    L.begin("k_decode_vec3");
    kern<<<(unsigned)grid, NT, smem, L.stream>>>(A);
    L.end();
    L.count++;
    return cudaGetLastError();
}

template <int NSUB>
static cudaError_t launch_decode_vec3_n(Launcher &L, const DecVec3Args &A, bool hash, int wrap) {
    if (hash) {
        if (wrap == 2) return launch_decode_vec3_t<NSUB, true, 2>(L, A);
        return wrap ? launch_decode_vec3_t<NSUB, true, 1>(L, A) : launch_decode_vec3_t<NSUB, true, 0>(L, A);
    }
    if (wrap == 2) return launch_decode_vec3_t<NSUB, false, 2>(L, A);
    return wrap ? launch_decode_vec3_t<NSUB, false, 1>(L, A) : launch_decode_vec3_t<NSUB, false, 0>(L, A);
}

cudaError_t launch_fused_decode_vec3(Launcher &L, const DecodeHost &h, int64_t nfiles) {
    DecVec3Args A = {};
    A.data = h.data; A.stream_len = h.stream_len; A.offsets = h.offsets; A.mins = h.mins; A.bits = h.bits;
    A.tab = h.tab; A.tab_per_file = h.tab_per_file; A.wrap_L = h.wrap_L; A.seed = h.seed;
    A.block_id0 = h.block_id0; A.nfile = h.nfile; A.subcells = h.subcells;
    A.sc3 = (long long)h.subcells * h.subcells * h.subcells;
    A.out = (float *)h.out;
    static const int sync_knob = getenv("MNW_DEC_SYNC") ? atoi(getenv("MNW_DEC_SYNC")) : 0;
    A.sync_mode = sync_knob;
    const long long units = nfiles * A.sc3;
    if (units == 0) return cudaSuccess;
    const int nsub = h.nfile / h.subcells;
    const long long n = (long long)nsub * nsub * nsub;
    A.nslabs = units * (n < 4096 ? 1 : n / 4096);
    if (A.nslabs >= (1LL << 31)) return cudaErrorInvalidValue;
    const bool hash = h.jmode == 1;
    const int wrap = h.wrap_L > 0.0f ? (h.low_nonneg ? 2 : 1) : 0;
    switch (nsub) {
        case 128: return launch_decode_vec3_n<128>(L, A, hash, wrap);
        case 64: return launch_decode_vec3_n<64>(L, A, hash, wrap);
        case 32: return launch_decode_vec3_n<32>(L, A, hash, wrap);
        case 16: return launch_decode_vec3_n<16>(L, A, hash, wrap);
    }
    return cudaErrorNotSupported;
}

}  // namespace mnw
