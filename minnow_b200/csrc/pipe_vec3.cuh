// pipe_vec3.cuh -- k_pipe_vec3, the warp-specialised minp encode for 64^3 sub-cells.
// Included by kernels_fused.cu inside namespace mnw (it shares that file's helpers:
// XStat, Fin, FusedArgs, lookback, pack_group_words, quantize_rare).
//
// Same arithmetic and the same outputs as k_fused_vec3 (minp.Writer.Vectors,
// go/minp/minp.go:86-119; floatGroup.writeData, go/group.go:312-327; bit.BufferedArray,
// go/bit/bit.go:84-134; blockIndex, go/block_index.go:16-23), different schedule:
//
//   * one cluster of 8 CTAs per sub-cell, CTA r stages rows [512 r, 512 r + 512) of the
//     sub-cell as 16-bit rotated pixel indices (3 x 32768 x 2 B = 192 KB);
//   * 8 LOADER warps read the AoS rows (each lane 48 contiguous bytes = 4 particles, four
//     steps in flight in registers), quantise two floats per instruction (FADD2 / FMUL2 /
//     FFMA2, floor by FADD2.RM against 2^23), keep min/max of the raw float bits and of the
//     rotated index, and store 4 indices of one axis per STS.64;
//   * 4 PACKER warps wait for the cluster's statistics (all-gathered through distributed
//     shared memory, completion on an mbarrier: no cluster-wide barrier in steady state),
//     finalise (min, bits, nbytes), get the byte offsets by decoupled look-back and pack
//     1024-element groups in row order; every finished slot of 16 rows is handed back to
//     the loaders (mbarrier per slot), which by then are already staging the NEXT sub-cell:
//     loading unit n+1 overlaps packing unit n in the same 192 KB.
//   * units are claimed two ahead through an atomic ticket (a predecessor in the look-back
//     chain is therefore always running or done) and broadcast through DSMEM + mbarrier.
//
// The fast quantiser result is trusted when the raw bits of RM(y + 2^23) lie in
// [2^23, 2^23 + pixels): that rejects negative, NaN, infinite and too large quotients in one
// unsigned range test, made ONCE per thread and unit on the running min/max; a thread that
// fails it redoes its 384 elements with the IEEE divide (pipe_redo_thread).


#ifdef MNW_PIPE_DBG
__device__ unsigned long long g_pipe_dbg[32 * 64 * 8];
__device__ __forceinline__ void pipe_dbg(int cluster, int it, int ev) {
    if (cluster < 32 && it < 64) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g_pipe_dbg[(cluster * 64 + it) * 8 + ev] = t;
    }
}
#define PIPE_DBG(it, ev) do { if ((blockIdx.x & 7) == 0 && lane == 0) pipe_dbg((int)(blockIdx.x / PIPE_CS), it, ev); } while (0)
#else
#define PIPE_DBG(it, ev) do { } while (0)
#endif

#ifndef MNW_PIPE_EXP
#define MNW_PIPE_EXP 0
#endif

namespace {

constexpr unsigned FMAGIC = 0x4B000000u;   // float bits of 2^23
constexpr int PIPE_LW = 8, PIPE_PW = 7, PIPE_NT = 32 * (PIPE_LW + PIPE_PW + 1);   // + 1 scanner warp
constexpr int PIPE_TBUF = 560;
constexpr int PIPE_PFD = 8;   // L2 prefetch distance beyond the register pipeline, in steps (a multiple of 4)
constexpr int PIPE_LREGS = 152, PIPE_PREGS = 104;   // registers per thread of the loader / packer warpgroups (2 x 128 in all)   // words of one packer warp's transposition buffer (33 * 16 + 1, rounded up)
constexpr int PIPE_CS = 8, PIPE_CHUNK = 32768, PIPE_STEPS = 32, PIPE_USLOTS = 8;

struct PipePar {   // per-axis parameters of one unit (written once per unit, read by all threads)
    float low, rcp, ndx;
    unsigned P, Cm, fast, oob0, pad;
    long long q0;
    long long pad1;
};

// a float pair from shared memory, opaque to the optimiser: it stays one 64-bit register
// (ptxas otherwise keeps the scalars and rebuilds every pair with two MOVs per use)
__device__ __forceinline__ unsigned long long ld_shared_u64_pinned(const unsigned long long *p) {
    unsigned long long r;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return r;
}
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity) {   // acquire at CTA scope
    const unsigned a = smem_u32(b);
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long *b, unsigned parity) {   // acquire at cluster scope
    const unsigned a = smem_u32(b);
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// arrive on the same barrier of CTA `rank` of the cluster, releasing this thread's earlier
// (distributed) shared memory stores at cluster scope
__device__ __forceinline__ void mbar_arrive_remote(unsigned long long *b, unsigned rank) {
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(b)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void st_remote_u64(const void *p, unsigned rank, unsigned long long v) {
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(rank));
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(ra), "l"(v) : "memory");
}
__device__ __forceinline__ void st_remote_v4(const void *p, unsigned rank, uint4 v) {
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(rank));
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void bar_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Where a unit (file f, sub-cell sc) starts in the AoS input, in float4 units within its file.
struct PipeGeom {
    int S, nfile;
    unsigned nsub;   // 64 or 32
    unsigned row4, plane4;
    long long sc3;
    __device__ __forceinline__ const float4 *origin(const float *aos, long long unit, long long &f, unsigned &sc) const {
        f = unit / sc3;
        sc = (unsigned)(unit - f * sc3);
        const unsigned ix0 = nsub * (sc % (unsigned)S), iy0 = nsub * ((sc / (unsigned)S) % (unsigned)S), iz0 = nsub * (sc / (unsigned)(S * S));
        return (const float4 *)(aos + 3 * f * (long long)nfile * nfile * nfile) + (3u * ix0 / 4u + iy0 * row4 + iz0 * plane4);
    }
};

// Exact redo of the share of the loader threads in bad_mask (rare: some element left the domain of the
// fast quantiser).  The warp works on one such thread at a time, lane j redoing step j (4 particles) with
// the IEEE divide, so the 384 elements cost a couple of memory round trips instead of 384.  The thread's
// statistics come back in the same form as the fast path (raw-bit domain: FMAGIC + q).
__device__ __noinline__ void pipe_redo_warp(unsigned bad_mask, const float4 *tbase, unsigned step_lo, unsigned step_hi, int steps_per_hi,
                                            const PipePar *par, unsigned short *stage, int e_thread,
                                            unsigned *st /*[12]*/, unsigned *oob_out) {
    const int lane = threadIdx.x & 31;
    const unsigned long long tb64 = (unsigned long long)(uintptr_t)tbase;
    while (bad_mask) {
        const int b = __ffs(bad_mask) - 1;
        bad_mask &= bad_mask - 1;
        const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)tb64, b), hi = __shfl_sync(0xffffffffu, (unsigned)(tb64 >> 32), b);
        const float4 *tb = (const float4 *)(uintptr_t)(((unsigned long long)hi << 32) | lo);
        const int eb = __shfl_sync(0xffffffffu, e_thread, b);
        unsigned s[12], oob = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            s[4 * k + 0] = ~0u; s[4 * k + 1] = 0u; s[4 * k + 2] = ~0u; s[4 * k + 3] = 0u;
            if (par[k].oob0) oob = 1;
        }
        const int t = lane;
        const float4 *src4 = tb + ((unsigned)(t / steps_per_hi) * step_hi + (unsigned)(t % steps_per_hi) * step_lo);
        const float4 v0 = __ldg(src4), v1 = __ldg(src4 + 1), v2 = __ldg(src4 + 2);
        const float x[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
        for (int i = 0; i < 12; i++) {
            const int k = i % 3, pi = i / 3;
            const PipePar &pp = par[k];
            const int q = quantize_rare(x[i], pp.low, -pp.ndx, (int)pp.P, oob);
            const unsigned tt = (unsigned)q + (pp.Cm + FMAGIC);
            const unsigned w = min(tt, tt - pp.P);
            s[4 * k + 0] = min(s[4 * k + 0], w); s[4 * k + 1] = max(s[4 * k + 1], w);
            s[4 * k + 2] = min(s[4 * k + 2], (unsigned)q + FMAGIC); s[4 * k + 3] = max(s[4 * k + 3], (unsigned)q + FMAGIC);
            const int e = 1024 * t + eb + pi;
            stage[k * PIPE_CHUNK + (e ^ (((e >> 6) & 7) << 3))] = (unsigned short)w;
        }
#pragma unroll
        for (int i = 0; i < 12; i++) {
            const unsigned r = (i & 1) ? __reduce_max_sync(0xffffffffu, s[i]) : __reduce_min_sync(0xffffffffu, s[i]);
            if (lane == b) st[i] = r;
        }
        oob = __any_sync(0xffffffffu, oob);
        if (lane == b) *oob_out = oob;
    }
}

// The exact path of one block of a unit that raised the out-of-range flag (whole warp, from global memory).  One pass
// classifies the block's pixel indices and takes the int64 min / max of bound(q, 0, pixels):
//   code 0  every index lies in [0, pixels]: this AXIS was clean (the flag is per unit) -- the caller keeps its closed form;
//   code 1  some indices are FAR outside (q >= 3 pixels or q < -2 pixels: NaN and infinities convert to the int64 minimum)
//           and none is near: periodicMin is 0 whatever the order -- up to the first far element the walk of
//           go/group.go:384-409 sees indices in range only (a valid arc, or it has returned 0 already), and a far element
//           grows the arc beyond pixels / 2 in either branch -- so min and max of bound(q, 0, pixels) are the answer;
//   code 2  anything else (an index slightly outside the range, or element 0 itself outside): the sequential walk.
// out4 = {pmin, min, max, code}.  Out of line: it must not cost the scanner warp's hot path a register.
__device__ __noinline__ void pipe_exact_block(const BlockDesc *dp, long long *out4) {
    const BlockDesc d = *dp;
    const int lane = threadIdx.x & 31;
    const long long P = d.pixels;
    long long mn = LLONG_MAX, mx = LLONG_MIN;
    bool near = false, far = false;
    for (int64_t base = 0; base < d.n; base += 32 * 16) {   // 16 independent gathers per lane in flight
        long long q[16];
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const int64_t i = base + 32 * u + lane;
            q[u] = i < d.n ? block_value(d, i) : 0;
        }
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (base + 32 * u + lane < d.n) {
                const long long v = q[u];
                const bool inr = (unsigned long long)v <= (unsigned long long)P;
                const bool isfar = v >= 3 * P || v < -2 * P;
                far = far || isfar;
                near = near || (!inr && !isfar);
                const long long qb = bound1(v, 0, P);
                mn = qb < mn ? qb : mn;
                mx = qb > mx ? qb : mx;
            }
        }
    }
    const bool any_near = __any_sync(0xffffffffu, near), any_far = __any_sync(0xffffffffu, far);
    const long long q0 = block_value(d, 0);
    const bool q0_in = (unsigned long long)q0 < (unsigned long long)P;
    if (!any_near && !any_far && q0_in) {
        out4[3] = 0;
    } else if (!any_near && q0_in && 2 * P > 0) {
        out4[0] = 0; out4[1] = warp_min_ll(mn); out4[2] = warp_max_ll(mx); out4[3] = 1;
    } else {
        slow_block_warp_values(d, out4[0], out4[1], out4[2]);
        out4[3] = 2;
    }
}

// 16 fields of 2 B bits (two packed values each) -> B little-endian stream words, shifts resolved at
// compile time (bit.BufferedArray, go/bit/bit.go:84-134, for 32 values of one lane).
template <int B>
__device__ __forceinline__ void pack_fields16(const unsigned (&f)[16], unsigned (&o)[B]) {
    constexpr int FW = 2 * B;
#pragma unroll
    for (int j = 0; j < B; j++) o[j] = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int bit = i * FW, wd = bit >> 5, sh = bit & 31;
        if (FW == 32) {
            o[wd % B] = f[i];
        } else {
            o[wd % B] += f[i] << sh;
            if (sh + FW > 32) o[(wd + 1) % B] = f[i] >> ((32 - sh) & 31);
        }
    }
}

// One pack group = 1024 consecutive elements of one block = 32 lanes x 16 fields -> 32 B stream words
// = 128 B bytes at dst (any byte alignment).  The lanes shift their words by (dst & 3) bytes, so that
// the transposition buffer holds ALIGNED words (word W at buf[W + (W >> 5)]: at most 2-way bank
// conflicts either way), and the warp copies them out 128 bytes per store; the bytes of the first
// and last partial word are stored one by one (the neighbouring groups own the rest of those words).
template <int B>
__device__ __forceinline__ void emit_group(const unsigned (&f)[16], unsigned *buf, int lane, uint8_t *dst) {
    unsigned o[B];
    pack_fields16<B>(f, o);
    const int ab = (int)((uintptr_t)dst & 3), sh = 8 * ab;
    unsigned prev = __shfl_up_sync(0xffffffffu, o[B - 1], 1);
    if (lane == 0) prev = 0;
    const int W0 = lane * B;
#pragma unroll
    for (int j = 0; j < B; j++) {
        const unsigned lo = j == 0 ? prev : o[(j + B - 1) % B];
        const int W = W0 + j;
        buf[W + (W >> 5)] = __funnelshift_l(lo, o[j], sh);
    }
    if (lane == 31) buf[33 * B] = o[B - 1] >> ((32 - sh) & 31);   // aligned word 32 B: the top ab bytes of the stream
    __syncwarp();
    uint32_t *base = (uint32_t *)(dst - ab) + lane;
    const unsigned *src = buf + lane;
    if (ab == 0) {
#pragma unroll
        for (int m = 0; m < B; m++) base[32 * m] = src[33 * m];
    } else {
        if (lane > 0) {
            base[0] = src[0];
        } else {
            const unsigned w0 = src[0], w1 = buf[33 * B];
            uint8_t *hp = dst - ab, *tp = dst - ab + 128 * B;
            for (int k = ab; k < 4; k++) hp[k] = (uint8_t)(w0 >> (8 * k));
            for (int k = 0; k < ab; k++) tp[k] = (uint8_t)(w1 >> (8 * k));
        }
#pragma unroll
        for (int m = 1; m < B; m++) base[32 * m] = src[33 * m];
    }
    __syncwarp();
}

}  // namespace

// COOP = false: one 8-CTA cluster per unit, statistics through distributed shared memory (above).
// COOP = true : no clusters.  The grid is launched cooperatively (all CTAs co-resident), CTA b works on the
// items g = b + gridDim.x * i, item g being part g % 8 of unit g / 8 (same file-round-robin order), and the
// eight parts of a unit meet through a 64-byte record in global memory (atomicMax on the statistics, a
// counter, ld.acquire polling).  Every CTA looks its offsets up itself.  This uses all SMs (144 of 148)
// instead of the 120 that 8-CTA clusters reach; a part only ever waits for items with a smaller g.
// NSUB = 32 (COOP only): a 32^3 sub-cell is exactly one CTA's 32768 particles, PARTS = 1: rows of 32 particles, 8 lanes per row,
// one z-plane per step; the staging layout, the packers and the look-back are the same.
template <bool COOP, int NSUB>
__global__ void __launch_bounds__(PIPE_NT, 1) k_pipe_vec3(const FusedArgs A) {
    static_assert(NSUB == 64 || ((NSUB == 32 || NSUB == 128) && COOP), "32^3 and 128^3 sub-cells: cooperative schedule only");
    constexpr int N = NSUB * NSUB * NSUB, CHUNK = PIPE_CHUNK, CS = PIPE_CS;
    constexpr int PARTS = NSUB * NSUB * NSUB / PIPE_CHUNK;     // CTAs per unit: 1, 8 or 64
    constexpr int PPP = PIPE_CHUNK / (NSUB * NSUB);           // z-planes per part: 32 (the whole unit), 8 or 2
    constexpr int LPR = NSUB / 4;                             // lanes per row (4 particles each): 16 or 8
    constexpr int SPP = NSUB / (8 * 32 / LPR);                // steps per z-plane: 1, 4 or 16
    constexpr int LT = 32 * PIPE_LW, PT = 32 * (PIPE_PW + 1);   // loader threads; packer + scanner threads
    constexpr int NGROUPS = 3 * (CHUNK / 1024);   // pack groups per CTA and unit
    constexpr int UM = PIPE_USLOTS - 1;
    // COOP, a handful of parts per unit: one tagged record per (unit, part), polled by every CTA of the unit; many parts
    // (128^3 sub-cells: 64): one record per unit, combined by atomics, with a counter
    constexpr bool REC = COOP && PARTS > 1 && PARTS <= 16;
    if (A.W.skip && *A.W.skip) return;   // uniform over the grid: written before the launch

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned short *stage = (unsigned short *)smem_raw;   // [3][CHUNK], swizzled like k_fused_vec3
    unsigned *tbuf = (unsigned *)(smem_raw + 6 * CHUNK);  // [PIPE_PW][PIPE_TBUF] transposition buffers
    __shared__ __align__(8) unsigned long long bar_unit[PIPE_USLOTS];   // 1 arrival: rank 0 posted s_unit[slot]
    __shared__ __align__(8) unsigned long long bar_stats[2];            // CS arrivals: every CTA posted its XStat
    __shared__ __align__(8) unsigned long long bar_empty[PIPE_STEPS / 2];   // 6 arrivals: two slots (32 rows) read by the packers
    __shared__ long long s_unit[PIPE_USLOTS];
    __shared__ __align__(16) PipePar s_par[2][3];
    __shared__ __align__(8) unsigned long long s_pp[2][3][3];   // [slot][low, rcp, ndx][pair (x,y) (z,x) (y,z)]
    __shared__ __align__(16) XStat s_x[2][CS][3];
    __shared__ unsigned s_red[PIPE_LW][13];
    __shared__ unsigned s_cta[13];
    __shared__ Fin s_fin[3];
    __shared__ __align__(8) unsigned long long bar_off[3][3];           // 1 arrival: rank 0 posted s_off[slot][axis]
    __shared__ long long s_off[3][3];   // byte offset of the unit's blocks
    __shared__ int s_gctr;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned crank = 0;
    if constexpr (!COOP) crank = cg::this_cluster().block_rank();
    PipeGeom G;
    G.S = A.subcells; G.nfile = A.nfile; G.sc3 = A.sc3; G.nsub = NSUB;
    G.row4 = 3u * (unsigned)A.nfile / 4u; G.plane4 = G.row4 * (unsigned)A.nfile;

    // x[0] of every axis block of `unit`, its rotation constant and the quantiser parameters
    auto prepare = [&](long long unit, int slot, int k) {
        long long f; unsigned sc;
        const float4 *org = G.origin(A.aos, unit, f, sc);
        const FloatParams fp = A.tab[(A.tab_per_file ? 3 * f : 0) + k];
        const long long q0 = quantize_exact(__ldg((const float *)org + k), fp.low, fp.dx);
        const bool ok = (unsigned long long)q0 < (unsigned long long)fp.pixels;
        PipePar pp;
        pp.low = fp.low; pp.rcp = fp.rcp; pp.ndx = -fp.dx; pp.P = (unsigned)fp.pixels;
        pp.Cm = (ok ? (unsigned)arc_rotation(q0, fp.pixels) : 0u) - FMAGIC;
        pp.fast = (ok && (fp.flags & F_FASTDIV)) ? 1u : 0u;
        pp.oob0 = ok ? 0u : 1u;   // periodicMin starting outside [0, pixels): exact path only
        pp.pad = 0; pp.q0 = q0; pp.pad1 = 0;
        s_par[slot][k] = pp;
    };
    // the same parameters as float pairs in the order the floats of a step arrive (lanes 0..8 of one warp,
    // after prepare() by lanes 0..2 and a __syncwarp)
    auto prepare_pairs = [&](int slot, int i) {
        const int n = i / 3, j = i - 3 * n;
        const int a = (2 * j) % 3, b = (2 * j + 1) % 3;   // pair j holds axes (a, b): (0,1) (2,0) (1,2)
        const PipePar &pa = s_par[slot][a], &pb = s_par[slot][b];
        const float fa = n == 0 ? pa.low : (n == 1 ? pa.rcp : pa.ndx), fb = n == 0 ? pb.low : (n == 1 ? pb.rcp : pb.ndx);
        s_pp[slot][n][j] = (unsigned long long)__float_as_uint(fa) | ((unsigned long long)__float_as_uint(fb) << 32);
    };
    // rank 0 only: next unit for the whole cluster.  Tickets walk the files round-robin (ticket t -> sub-cell
    // t / nfiles of file t % nfiles): the predecessor of a block in its look-back chain, the previous sub-cell
    // of the same file, is then nfiles tickets old -- finished long ago when the batch holds many files, so the
    // look-back never waits -- and still always claimed earlier, i.e. running or done.
    const long long nfiles = A.nunits / A.sc3;
    auto claim = [&](int slot) {
        const long long t = (long long)atomicAdd(A.W.ticket, 1u);
        const unsigned long long u = t < A.nunits ? (unsigned long long)((t % nfiles) * A.sc3 + t / nfiles) : (unsigned long long)A.nunits;
        for (unsigned r = 0; r < CS; r++) {
            st_remote_u64(&s_unit[slot], r, u);
            mbar_arrive_remote(&bar_unit[slot], r);
        }
    };

    // unit and part of this CTA's it-th piece of work
    auto unit_at = [&](int i, unsigned &rk) -> long long {
        if constexpr (COOP) {
            const long long g = (long long)blockIdx.x + (long long)gridDim.x * i;
            rk = (unsigned)(g % PARTS);
            if (g >= PARTS * A.nunits) return A.nunits;
            const long long up = g / PARTS;
            return (up % nfiles) * A.sc3 + up / nfiles;
        } else {
            mbar_wait_cluster(&bar_unit[i & UM], (i / PIPE_USLOTS) & 1);
            rk = crank;
            return s_unit[i & UM];
        }
    };

    if (tid == 0) {
        for (int i = 0; i < PIPE_USLOTS; i++) mbar_init(&bar_unit[i], 1);
        mbar_init(&bar_stats[0], PARTS == 1 ? 1 : CS); mbar_init(&bar_stats[1], PARTS == 1 ? 1 : CS);
        for (int i = 0; i < PIPE_STEPS / 2; i++) mbar_init(&bar_empty[i], 6);
        for (int i = 0; i < 9; i++) mbar_init(&bar_off[0][0] + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if constexpr (COOP) __syncthreads(); else cluster_sync_all<CS>();
    if constexpr (!COOP) { if (crank == 0 && tid == 0) { claim(0); claim(1); } }
    {
        unsigned rk0;
        const long long u0 = unit_at(0, rk0);
        if (tid < 3 && u0 < A.nunits) prepare(u0, 0, tid);
        __syncwarp();
        if (tid < 9 && u0 < A.nunits) prepare_pairs(0, tid);
    }
    __syncthreads();

    if (warp < PIPE_LW) {
        // =========================== loaders ===========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PIPE_LREGS));
        // lane geometry: step t covers rows 16 t .. 16 t + 15 of the CTA's 512; warp w rows 2 w, 2 w + 1
        // of those, lane l particles 4 (l & 15) .. + 3 of row (l >> 4)
        const unsigned toff0 = (unsigned)((32 / LPR) * warp + lane / LPR) * G.row4 + 3u * (unsigned)(lane % LPR);
        auto toff = [&](unsigned rk) { return rk * (unsigned)PPP * G.plane4 + toff0; };   // rk: which part of the unit
        const int e_thread = 128 * warp + 4 * lane;   // element of the thread's first particle at step 0
        // byte offset of that element's 8-byte piece in an axis' staging array (chunk swizzle c ^ ((c >> 3) & 7))
        const unsigned sbyte = (unsigned)((((e_thread >> 3) ^ ((e_thread >> 6) & 7)) << 4) + ((lane & 1) << 3));
        const unsigned step_lo = (unsigned)(8 * 32 / LPR) * G.row4;   // rows of one step: 16 (64^3) or 32 = a whole plane (32^3)
        auto step_off = [&](int t) { return (unsigned)(t / SPP) * G.plane4 + (unsigned)(t % SPP) * step_lo; };

        // step t -> t + 1: the next rows of the plane, or (every SPP-th step) on to the next plane
        const size_t inc_rows = SPP == 1 ? (size_t)G.plane4 : (size_t)step_lo, inc_plane = (size_t)G.plane4 - (size_t)(SPP - 1) * step_lo;
        float4 buf[4][3] = {};
        const float4 *pr = nullptr;   // where the next refill comes from (step t + 5)
        const float4 *pq = nullptr;   // the same, PIPE_PFD steps further: what is pulled towards L2 now
        unsigned rank = 0, nrank = 0;
        long long unit = unit_at(0, rank);
        long long f; unsigned sc;
        const float4 *cur = nullptr;
        // per-axis parameters as pairs in the order the 12 floats of a step arrive: (x,y) (z,x) (y,z) (x,y) (z,x) (y,z)
        unsigned long long lowp[3], rcpp[3], ndxp[3];
        auto load_pairs = [&](int slot) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                lowp[j] = ld_shared_u64_pinned(&s_pp[slot][0][j]);
                rcpp[j] = ld_shared_u64_pinned(&s_pp[slot][1][j]);
                ndxp[j] = ld_shared_u64_pinned(&s_pp[slot][2][j]);
            }
        };
        // the float half of a step: raw bits of RM(quotient + 2^23) of the 12 floats in buffer u -> bb[particle][axis]
        auto quant12 = [&](int u, unsigned (&bb)[4][3]) {
            const float4 v0 = buf[u][0], v1 = buf[u][1], v2 = buf[u][2];
            const unsigned long long r0 = quantize2(f2_pack(v0.x, v0.y), lowp[0], rcpp[0], ndxp[0]);
            const unsigned long long r1 = quantize2(f2_pack(v0.z, v0.w), lowp[1], rcpp[1], ndxp[1]);
            const unsigned long long r2 = quantize2(f2_pack(v1.x, v1.y), lowp[2], rcpp[2], ndxp[2]);
            const unsigned long long r3 = quantize2(f2_pack(v1.z, v1.w), lowp[0], rcpp[0], ndxp[0]);
            const unsigned long long r4 = quantize2(f2_pack(v2.x, v2.y), lowp[1], rcpp[1], ndxp[1]);
            const unsigned long long r5 = quantize2(f2_pack(v2.z, v2.w), lowp[2], rcpp[2], ndxp[2]);
            f2_bits(r0, bb[0][0], bb[0][1]); f2_bits(r1, bb[0][2], bb[1][0]); f2_bits(r2, bb[1][1], bb[1][2]);
            f2_bits(r3, bb[2][0], bb[2][1]); f2_bits(r4, bb[2][2], bb[3][0]); f2_bits(r5, bb[3][1], bb[3][2]);
        };
        unsigned bc[4][3];   // float half of the step whose integer half comes next (software pipeline, one step deep)
        if (unit < A.nunits) {
            cur = G.origin(A.aos, unit, f, sc) + toff(rank);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float4 *p = cur + step_off(u);
                buf[u][0] = ld_stream_pinned(p); buf[u][1] = ld_stream_pinned(p + 1); buf[u][2] = ld_stream_pinned(p + 2);
            }
            pr = cur + step_off(4);
            load_pairs(0);
            quant12(0, bc);
            buf[0][0] = ld_stream_pinned(pr); buf[0][1] = ld_stream_pinned(pr + 1); buf[0][2] = ld_stream_pinned(pr + 2);
            pr += inc_rows;
            pq = pr + (size_t)(PIPE_PFD / SPP) * G.plane4 + (size_t)(PIPE_PFD % SPP) * step_lo;   // (no plane boundary in between)
        }
        for (int it = 0; unit < A.nunits; it++) {
            if (warp == 0) PIPE_DBG(it, 0);
            long long next = A.nunits;
            const float4 *nxt = nullptr;
            const PipePar *par = s_par[it & 1];
            unsigned Cm[3], nP[3];
            bool fast = true;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                Cm[j] = par[j].Cm; nP[j] = 0u - par[j].P;
                fast = fast && par[j].fast;
            }
            unsigned wmin[3] = {~0u, ~0u, ~0u}, wmax[3] = {0u, 0u, 0u};
            unsigned bmin[3] = {~0u, ~0u, ~0u}, bmax[3] = {0u, 0u, 0u};

#pragma unroll 1
            for (int t0 = 0; t0 < PIPE_STEPS; t0 += 4) {
                if (t0 == 16) {   // the next unit: its origin and its parameters
                    next = unit_at(it + 1, nrank);
                    if (next < A.nunits) {
                        long long nf; unsigned nsc;
                        const float4 *nb = G.origin(A.aos, next, nf, nsc);
                        nxt = nb + toff(nrank);
                        if (warp == 0) {
                            if (lane < 3) prepare(next, (it + 1) & 1, lane);
                            __syncwarp();
                            if (lane < 9) prepare_pairs((it + 1) & 1, lane);
                        }
                    }
                }
#pragma unroll
                for (int u2 = 0; u2 < 4; u2 += 2) {
                    uint2 pk[2][3];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int u = u2 + h, un = (u + 1) & 3;
                        // ---- float half of step t + 1 (buffer un), then that buffer goes back to step t + 5; the last
                        // one of a unit is step 0 of the NEXT unit and takes that unit's parameters ----
                        unsigned bn[4][3];
                        if (u == 3 && t0 == PIPE_STEPS - 4) {
                            bar_named(4, LT);   // s_pp of the next unit is complete (written by warp 0 since t0 == 16)
                            if (nxt) load_pairs((it + 1) & 1);
                        }
                        quant12(un, bn);
                        if (u == 3 && t0 == PIPE_STEPS - 8) pr = nxt;   // step t + 5 == 32: the refills move on to the next unit
                        if (u == 3 && t0 == PIPE_STEPS - 8 - PIPE_PFD) pq = nxt;   // and so does the L2 prefetch, PIPE_PFD steps earlier
                        if (A.prefetch && (lane % LPR) == 0 && (t0 + u + 5 + PIPE_PFD < PIPE_STEPS || nxt != nullptr))
                            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pq), "r"(NSUB * 12) : "memory");
                        pq += (SPP > 4 ? ((t0 + u + 5 + PIPE_PFD) % SPP) == SPP - 1 : u == 2) ? inc_plane : inc_rows;
                        if (t0 + u + 5 < PIPE_STEPS || nxt != nullptr) {
                            buf[un][0] = ld_stream_pinned(pr); buf[un][1] = ld_stream_pinned(pr + 1); buf[un][2] = ld_stream_pinned(pr + 2);
                        }
                        pr += (SPP > 4 ? ((t0 + u + 5) % SPP) == SPP - 1 : u == 2) ? inc_plane : inc_rows;
                        // ---- integer half of step t ----
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            unsigned w[4];
#pragma unroll
                            for (int pi = 0; pi < 4; pi++) {
                                const unsigned tt = bc[pi][k] + Cm[k];      // q + rotation
                                w[pi] = min(tt, tt + nP[k]);                // mod pixels
                            }
                            bmin[k] = __vimin3_u32(bmin[k], bc[0][k], bc[1][k]); bmin[k] = __vimin3_u32(bmin[k], bc[2][k], bc[3][k]);
                            bmax[k] = __vimax3_u32(bmax[k], bc[0][k], bc[1][k]); bmax[k] = __vimax3_u32(bmax[k], bc[2][k], bc[3][k]);
                            wmin[k] = __vimin3_u32(wmin[k], w[0], w[1]); wmin[k] = __vimin3_u32(wmin[k], w[2], w[3]);
                            wmax[k] = __vimax3_u32(wmax[k], w[0], w[1]); wmax[k] = __vimax3_u32(wmax[k], w[2], w[3]);
                            pk[h][k].x = __byte_perm(w[0], w[1], 0x5410);
                            pk[h][k].y = __byte_perm(w[2], w[3], 0x5410);
                        }
#pragma unroll
                        for (int pi = 0; pi < 4; pi++)
#pragma unroll
                            for (int k = 0; k < 3; k++) bc[pi][k] = bn[pi][k];
                    }
                    const int t = t0 + u2;
                    if (it > 0) mbar_wait(&bar_empty[t >> 1], (unsigned)(it - 1) & 1u);   // the packers have read these two slots
#pragma unroll
                    for (int h = 0; h < 2; h++)
#pragma unroll
                        for (int k = 0; k < 3; k++)
                            *(uint2 *)(smem_raw + k * (2 * CHUNK) + 2048 * (t + h) + sbyte) = pk[h][k];
                }
            }

            if (warp == 0) PIPE_DBG(it, 1);
            // ---- this thread's statistics; exact redo when the fast quantiser left its domain ----
            unsigned st[12], oob = 0;
            bool ok = fast;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                ok = ok && bmin[k] >= FMAGIC && bmax[k] < FMAGIC + par[k].P;
                st[4 * k + 0] = wmin[k]; st[4 * k + 1] = wmax[k]; st[4 * k + 2] = bmin[k]; st[4 * k + 3] = bmax[k];
            }
            {
                const unsigned bad = __ballot_sync(0xffffffffu, !ok);
                if (bad) pipe_redo_warp(bad, cur, step_lo, G.plane4, SPP, par, stage, e_thread, st, &oob);
            }
#pragma unroll
            for (int s = 0; s < 12; s++) {
                const unsigned r = (s & 1) ? __reduce_max_sync(0xffffffffu, st[s]) : __reduce_min_sync(0xffffffffu, st[s]);
                if (lane == 0) s_red[warp][s] = r;
            }
            oob = __any_sync(0xffffffffu, oob);
            if (lane == 0) s_red[warp][12] = oob;
            bar_named(1, LT);
            if (tid < 13) {
                unsigned m = s_red[0][tid];
                for (int wi = 1; wi < PIPE_LW; wi++) {
                    const unsigned v = s_red[wi][tid];
                    m = tid == 12 ? (m | v) : ((tid & 1) ? max(m, v) : min(m, v));
                }
                s_cta[tid] = m;
            }
            __syncwarp();
            if constexpr (PARTS == 1) {
                // this CTA holds the whole unit: its statistics are the unit's, the scanner reads them from s_cta
                if (tid == 0) mbar_arrive(&bar_stats[it & 1]);
            } else if constexpr (COOP) {
                // this part's record in global memory: word s = tag 1 << 32 | statistic s ([4 k + {0, 1, 2, 3}] = wmin, wmax, qmin,
                // qmax of axis k, [12] oob).  One relaxed 64-bit store per word and nothing else: value and tag arrive together,
                // so the readers need neither an atomic, nor a fence, nor a counter (three dependent round trips less per part)
                if constexpr (REC) {
                    if (tid < 13) {
                        const unsigned m = s_cta[tid];
                        const unsigned v = tid < 12 ? ((tid & 2) ? m - FMAGIC : m) : (m ? 1u : 0u);
                        st_relaxed((unsigned long long *)A.ustat + ((size_t)unit * PARTS + rank) * 16 + tid, (1ULL << 32) | v);
                    }
                } else {
                    // one record per unit: [4 k + {0, 1, 2, 3}] = ~wmin, wmax, ~qmin, qmax of axis k (atomicMax from zero),
                    // [12] oob, [13] parts arrived
                    unsigned *us = A.ustat + 16 * unit;
                    if (tid < 13) {
                        const unsigned m = s_cta[tid];
                        if (tid < 12) {
                            const unsigned v = (tid & 2) ? m - FMAGIC : m;
                            atomicMax(us + tid, (tid & 1) ? v : ~v);
                        } else if (m) {
                            atomicOr(us + 12, 1u);
                        }
                        __threadfence();
                    }
                    __syncwarp();
                    if (tid == 0) atomicAdd(us + 13, 1u);
                }
            } else if (tid < CS) {   // post this CTA's statistics to CTA `tid` and tell it
                const int par_i = it & 1;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    uint4 a, bq;
                    a.x = s_cta[4 * k + 0]; a.y = s_cta[4 * k + 1];
                    a.z = s_cta[4 * k + 2] - FMAGIC; a.w = s_cta[4 * k + 3] - FMAGIC;
                    bq.x = s_cta[12]; bq.y = bq.z = bq.w = 0;
                    st_remote_v4(&s_x[par_i][rank][k], (unsigned)tid, a);
                    st_remote_v4((const unsigned char *)&s_x[par_i][rank][k] + 16, (unsigned)tid, bq);
                }
                mbar_arrive_remote(&bar_stats[par_i], (unsigned)tid);
            }
            cur = nxt;
            unit = next;
            rank = nrank;
        }
    } else {
        // ====================== packers (7 warps) and the scanner (1 warp) ======================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PIPE_PREGS));
        const int pw = warp - PIPE_LW;
        const bool scanner = pw == PIPE_PW;
        unsigned *mybuf = tbuf + (scanner ? 0 : pw) * PIPE_TBUF;
        long long pre_off = 0;      // scanner of rank 0: this unit's offsets were posted one unit ago
        bool pre_posted = false;
        for (int it = 0;; it++) {
            unsigned rank;
            const long long unit = unit_at(it, rank);
            if (unit >= A.nunits) break;
            const long long f = unit / A.sc3, sc = unit - f * A.sc3;
            const int par_i = it & 1;
            if (scanner) {
                unsigned ured = 0;   // COOP, several parts: statistic `lane` of the unit
                if constexpr (PARTS == 1) {
                    mbar_wait(&bar_stats[par_i], (unsigned)(it >> 1) & 1u);
                } else if constexpr (COOP && !REC) {   // all parts of the unit have posted their statistics
                    const unsigned *us = A.ustat + 16 * unit;
                    unsigned c;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(c) : "l"(us + 13) : "memory");
                        if (c < (unsigned)PARTS) __nanosleep(64);
                    } while (c < (unsigned)PARTS);
                    if (lane < 13) { const unsigned v = __ldcg(us + lane); ured = (lane & 1) || lane == 12 ? v : ~v; }
                } else if constexpr (COOP) {   // all parts of the unit have posted their statistics: lane s < 13 combines word s
                    const unsigned long long *rec = (const unsigned long long *)A.ustat + (size_t)unit * PARTS * 16 + (lane < 13 ? lane : 0);
                    for (;;) {
                        bool all = true;
                        unsigned mnv = ~0u, mxv = 0u;
#pragma unroll 8
                        for (int r = 0; r < PARTS; r++) {
                            const unsigned long long v = ld_relaxed(rec + r * 16);
                            all = all && (v >> 32) != 0;
                            mnv = min(mnv, (unsigned)v); mxv = max(mxv, (unsigned)v);
                        }
                        ured = (lane & 1) || lane == 12 ? mxv : mnv;   // (oob: the maximum of 0 / 1 is the OR)
                        if (__all_sync(0xffffffffu, all)) break;
                        __nanosleep(32);
                    }
                } else {
                    mbar_wait_cluster(&bar_stats[par_i], (unsigned)(it >> 1) & 1u);
                }
                PIPE_DBG(it, 2);
                // ---- finalise: lane k < 3 combines the cluster's statistics of axis k (every CTA, redundantly) ----
                const int k = lane < 3 ? lane : 0;
                const long long f_b = f * 3 * A.sc3 + k * A.sc3 + sc;   // block id in the batch
                long long nbytes = 0, mn = 0, pmin = 0, q0k = 0;
                unsigned base = 0, padj = 0;
                int bits = 0, mode = 0;
                bool slow = false;
                const unsigned u_wmin = __shfl_sync(0xffffffffu, ured, 4 * k), u_wmax = __shfl_sync(0xffffffffu, ured, 4 * k + 1);
                const unsigned u_qmin = __shfl_sync(0xffffffffu, ured, 4 * k + 2), u_qmax = __shfl_sync(0xffffffffu, ured, 4 * k + 3);
                const unsigned u_oob = __shfl_sync(0xffffffffu, ured, 12);
                if (lane < 3) {
                    XStat x;
                    x.wmin = ~0u; x.wmax = 0u; x.qmin = INT_MAX; x.qmax = INT_MIN; x.oob = 0;
                    if constexpr (PARTS == 1) {
                        x.wmin = s_cta[4 * k]; x.wmax = s_cta[4 * k + 1];
                        x.qmin = (int)(s_cta[4 * k + 2] - FMAGIC); x.qmax = (int)(s_cta[4 * k + 3] - FMAGIC);
                        x.oob = s_cta[12];
                    } else if constexpr (COOP) {
                        x.wmin = u_wmin; x.wmax = u_wmax; x.qmin = (int)u_qmin; x.qmax = (int)u_qmax; x.oob = u_oob;
                    } else {
#pragma unroll
                        for (int r = 0; r < CS; r++) {
                            const uint4 a = *(const uint4 *)&s_x[par_i][r][k];
                            x.wmin = min(x.wmin, a.x); x.wmax = max(x.wmax, a.y);
                            x.qmin = min(x.qmin, (int)a.z); x.qmax = max(x.qmax, (int)a.w);
                            x.oob |= s_x[par_i][r][k].oob;
                        }
                    }
                    const PipePar pp = s_par[par_i][k];
                    const long long Pk = (long long)pp.P, half = Pk / 2, K = Pk - half - 1;
                    q0k = pp.q0;
                    unsigned long long maxoff;
                    bool wide;
                    const unsigned long long spread = (unsigned long long)x.wmax - x.wmin + 1ULL;
                    if (spread > (unsigned long long)half) {   // arc too wide: periodicMin returns 0
                        wide = true;
                        pmin = 0; mn = x.qmin; maxoff = (unsigned long long)((long long)x.qmax - x.qmin);
                        base = (pp.Cm + FMAGIC) + (unsigned)x.qmin; padj = (unsigned)Pk;
                    } else {
                        wide = false;
                        long long m = q0k + ((long long)x.wmin - K);
                        if (m < 0) m += Pk;
                        pmin = m; mn = m; maxoff = spread - 1ULL;
                        base = x.wmin; padj = 0;
                    }
                    bits = 64 - __clzll((long long)maxoff);   // bit.PrecisionNeeded (maxoff < 2^32 here)
                    nbytes = array_bytes(bits, N);
                    slow = x.oob != 0 || pp.oob0 != 0;
                    // staged values are the low 16 bits of w: enough when the packed value has <= 16 bits
                    // and (wide arcs) w itself fits, i.e. pixels <= 65536
                    mode = (bits >= 1 && bits <= 16 && !(wide && Pk > 65536)) ? 1 : 0;
                }
                // ---- a unit that holds NaN / out-of-range values (rare): the exact sequential periodicMin and the int64
                // statistics of its blocks, by this warp, from global memory -- the blocks then take the k_pack list like
                // any block the staging cannot represent, and nothing else in the batch is disturbed ----
                const unsigned slow_mask = __ballot_sync(0xffffffffu, slow);
                if (slow_mask) {
#pragma unroll 1
                    for (int kk = 0; kk < 3; kk++) {
                        if (!((slow_mask >> kk) & 1u)) continue;
                        const long long fb = f * 3 * A.sc3 + kk * A.sc3 + sc;
                        long long ex[4];   // pmin, min, max, code
                        pipe_exact_block(A.descs + fb, ex);
                        if (lane == kk && ex[3] != 0) {
                            pmin = ex[0]; mn = ex[1];
                            bits = precision_needed((unsigned long long)ex[2] - (unsigned long long)ex[1]);
                            if (bits < 0) { bits = 64; atomicExch(A.W.err, 1); }
                            nbytes = array_bytes(bits, N);
                            mode = 0;
                        }
                    }
                }
                if (lane < 3) {
                    if (rank == 0) st_relaxed(A.W.pub + f_b, PUB_AGG | (unsigned long long)nbytes);
                    Fin fin;
                    fin.off = nbytes; fin.bits = bits; fin.mode = mode; fin.base = base; fin.padj = padj;
                    s_fin[k] = fin;
                    if (rank == 0) {
                        if (bits > 0 && mode == 0) A.W.repack_list[atomicAdd(A.W.repack_count, 1)] = f_b;
                        BlockStat bs = {};
                        bs.pmin = pmin; bs.min = mn; bs.nbytes = nbytes; bs.out_off = 0; bs.do_bound = 1; bs.bits = bits;
                        bs.q0 = q0k; bs.oob = slow; bs.slow = slow;
                        A.stats[f_b] = bs;
                        if (A.mins) A.mins[f_b] = mn;
                        if (A.bits) A.bits[f_b] = bits;
                    }
                }
                if (lane == 0) s_gctr = 0;
                __syncwarp();
                bar_named(2, PT);   // s_fin is final, the staged indices of the unit are visible
                PIPE_DBG(it, 3);

                // offsets reach the packers of the cluster (broadcast by rank 0) or of this CTA (COOP)
                auto post_off = [&](int slot, int kk, long long o) {
                    if constexpr (COOP) {
                        if (lane == 0) { s_off[slot][kk] = o; mbar_arrive(&bar_off[slot][kk]); }
                    } else if (lane < CS) {
                        st_remote_u64(&s_off[slot][kk], (unsigned)lane, (unsigned long long)o);
                        mbar_arrive_remote(&bar_off[slot][kk], (unsigned)lane);
                    }
                };
                if (COOP || rank == 0) {
                    // the unit after next for the whole cluster (the loaders want it half way through the next one)
                    if constexpr (!COOP) { if (lane == 0) claim((it + 2) & UM); }
                    // ---- byte offsets (rank 0 only: one poller per cluster).  The offset of a block is the inclusive
                    // prefix of the previous sub-cell of its file, which does not depend on this unit at all: it was
                    // fetched and broadcast one unit ago if it had been published by then (always, when tickets walk
                    // many files round-robin); otherwise the decoupled look-back runs now, while the packers pack. ----
                    long long off = pre_off;
                    if (!pre_posted) {
                        bool have = true;
                        off = 0;
                        if (lane < 3 && sc > 0) {
                            const unsigned long long v = ld_relaxed(A.W.pub + f_b - 1);
                            have = (v >> 62) == 2;
                            off = (long long)(v & PUB_VALUE);
                        }
                        if (!__all_sync(0xffffffffu, have)) {
#pragma unroll 1
                            for (int kk = 0; kk < 3; kk++) {
                                const long long fb = f * 3 * A.sc3 + kk * A.sc3 + sc;
                                const long long o = lookback(A.W.pub, fb - sc, fb);
                                if (lane == kk) off = o;
                            }
                        }
#pragma unroll
                        for (int kk = 0; kk < 3; kk++) {
                            const long long o = __shfl_sync(0xffffffffu, off, kk);
                            post_off(it % 3, kk, o);
                        }
                    }
                    PIPE_DBG(it, 4);
                    if (lane < 3 && rank == 0) {
                        st_relaxed(A.W.pub + f_b, PUB_PREFIX | (unsigned long long)(off + nbytes));
                        if (off + nbytes > A.axis_stride) atomicExch(A.W.err, 2);   // the packers skip such a block
                        A.stats[f_b].out_off = off;
                        if (A.offsets) A.offsets[f_b] = off;
                        if (A.out_len && sc == A.sc3 - 1) A.out_len[f * 3 + k] = off + nbytes;
                    }
                    // ---- the next unit's offsets, if its predecessors are already through ----
                    pre_posted = false;
                    {
                        unsigned nrk;
                        const long long nu = unit_at(it + 1, nrk);   // claimed one unit ago (or in the prologue)
                        bool have = nu < A.nunits;
                        long long noff = 0;
                        if (have && lane < 3) {
                            const long long nf = nu / A.sc3, nsc = nu - nf * A.sc3;
                            if (nsc > 0) {
                                const unsigned long long v = ld_relaxed(A.W.pub + (nf * 3 * A.sc3 + k * A.sc3 + nsc) - 1);
                                have = (v >> 62) == 2;
                                noff = (long long)(v & PUB_VALUE);
                            }
                        }
                        if (__all_sync(0xffffffffu, have)) {
                            pre_posted = true;
                            pre_off = noff;
#pragma unroll
                            for (int kk = 0; kk < 3; kk++) {
                                const long long o = __shfl_sync(0xffffffffu, noff, kk);
                                post_off((it + 1) % 3, kk, o);
                            }
                        }
                    }
                    PIPE_DBG(it, 6);
                }
            } else {
                bar_named(2, PT);   // s_fin is final, the staged indices of the unit are visible
                // ---- pack groups of 1024 elements in row order; a slot (16 rows, one group per axis) goes
                // back to the loaders as soon as its three groups are in registers ----
                for (;;) {
                    int g = 0;
                    if (lane == 0) g = atomicAdd(&s_gctr, 1);
                    g = __shfl_sync(0xffffffffu, g, 0);
                    if (g >= NGROUPS) break;
                    const int gi = g / 3, k = g - 3 * gi;
                    const Fin fin = s_fin[k];
                    if (fin.mode == 0) {
                        if (lane == 0) mbar_arrive(&bar_empty[gi >> 1]);
                        continue;
                    }
                    const int eb = gi * 1024 + 32 * lane;   // this lane's first element within the CTA's chunk
                    const int sw = (eb >> 6) & 7;
                    uint4 r[4];
#pragma unroll
                    for (int s = 0; s < 4; s++) r[s] = *(const uint4 *)(stage + k * CHUNK + ((((eb >> 3) + s) ^ sw) << 3));
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_empty[gi >> 1]);
                    const unsigned rr[16] = {r[0].x, r[0].y, r[0].z, r[0].w, r[1].x, r[1].y, r[1].z, r[1].w,
                                             r[2].x, r[2].y, r[2].z, r[2].w, r[3].x, r[3].y, r[3].z, r[3].w};
                    unsigned fld[16];   // two values per field: v[2 i] | v[2 i + 1] << bits
                    const unsigned kf = 65536u - (1u << fin.bits);
                    if (fin.padj == 0) {             // narrow arc: v = w - wmin, exact modulo 2^16
                        const unsigned b2 = (fin.base & 0xffffu) * 0x10001u;
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const unsigned d = __vsub2(rr[i], b2);
                            fld[i] = d - (d >> 16) * kf;
                        }
                    } else if (fin.padj <= 32768u) {   // wide arc: v = (w - C - qmin) mod pixels, in 16-bit lanes
                        const unsigned b2 = (fin.base & 0xffffu) * 0x10001u, p2 = fin.padj * 0x10001u;
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const unsigned d0 = __vsub2(rr[i], b2);
                            const unsigned d = __viaddmin_u16x2(d0, p2, d0);
                            fld[i] = d - (d >> 16) * kf;
                        }
                    } else {                           // wide arc, 32768 < pixels <= 65536: 32-bit lanes
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const unsigned lo = (rr[i] & 0xffffu) - fin.base, hi = (rr[i] >> 16) - fin.base;
                            fld[i] = min(lo, lo + fin.padj) + (min(hi, hi + fin.padj) << fin.bits);
                        }
                    }
                    // the block's byte offset is posted by the scanner of rank 0 (never write past the caller's buffer)
                    mbar_wait_cluster(&bar_off[it % 3][k], (unsigned)(it / 3) & 1u);
                    const long long o64 = s_off[it % 3][k];
                    if (o64 + fin.off <= A.axis_stride) {
                        const long long e0 = (long long)rank * CHUNK + (long long)gi * 1024;   // element index in the block
                        uint8_t *dst = A.out + (f * 3 + k) * A.axis_stride + o64 + ((e0 * fin.bits) >> 3);
                        switch (fin.bits) {
#define MNW_CASE(B) case B: emit_group<B>(fld, mybuf, lane, dst); break;
                            MNW_CASE(1) MNW_CASE(2) MNW_CASE(3) MNW_CASE(4) MNW_CASE(5) MNW_CASE(6) MNW_CASE(7) MNW_CASE(8)
                            MNW_CASE(9) MNW_CASE(10) MNW_CASE(11) MNW_CASE(12) MNW_CASE(13) MNW_CASE(14) MNW_CASE(15) MNW_CASE(16)
#undef MNW_CASE
                            default: break;
                        }
                    }
                }
            }
            bar_named(3, PT);   // s_fin / s_off / s_gctr are rewritten for the next unit
            if (scanner) PIPE_DBG(it, 7);
        }
    }
    // nobody leaves while a peer may still store into its shared memory
    __syncthreads();
    if constexpr (!COOP) cluster_sync_all<CS>();
}
