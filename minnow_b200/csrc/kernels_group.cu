// kernels_group.cu -- k_group_fused: the single-DRAM-read encoder of contiguous int64 / float32 column blocks
// (intGroup.writeData go/group.go:242-255, floatGroup.writeData :312-327, blockIndex go/block_index.go:16-35).
//
// The two passes a block needs -- statistics (min, max, periodic arc -> min, bits, nbytes) and packing -- run in
// ONE persistent kernel, a whole "wave" of tiles apart, so that the packing pass finds its input in the 126 MB L2
// instead of HBM:
//
//   ticket order   S(0) S(1) P(0) S(2) P(1) ... S(G-1) P(G-2) P(G-1)     S(g) / P(g): the statistics / pack passes
//                                                                        of wave g; a ticket = up to `sup` tiles
//                                                                        (4096 elements each) of one pass
//   * a CTA is 1 PRODUCER warp + 4 CONSUMER warps around a ring of shared-memory slots.  The producer claims
//     tickets from an atomic counter, turns their tiles into jobs and fetches each tile with ONE TMA bulk copy
//     (cp.async.bulk, completion on the slot's mbarrier), several tiles ahead of the consumers: HBM / L2 latency is
//     hidden by the ring, not by occupancy or registers.
//   * every consumer warp owns a quarter (1024 elements) of every tile and never meets the other warps at a barrier:
//     statistics are accumulated in registers over the tiles of a ticket and posted per warp (atomics on the block's
//     record + a counter); the pack pass quantises its quarter in place in the slot, packs it with the compile-time
//     warp packer (pack.cuh) and writes it to its byte-aligned place.
//   * the warp that posts the last statistics of a block finalises it (closed-form periodicMin, min, bits, nbytes;
//     the exact sequential periodicMin for blocks with out-of-range values), publishes its size and gets its byte
//     offset by decoupled look-back over the earlier blocks of its group (chain).  The producer lets a pack job
//     through once its block's PREFIX word is there (normally a wave earlier).
//   * a CTA only ever waits for tiles with SMALLER tickets, which are held by CTAs that are running: no deadlock
//     under any residency, no cooperative launch.
//
// The quantiser works on float PAIRS (f32x2.cuh).  Algorithmic bytes per element: 4 + bits/8 (float32), 8 + bits/8
// (int64); DRAM traffic is the same as long as a wave stays L2-resident between its two passes.  Blocks wider than
// 32 bits (NaN, huge ranges) are listed for k_pack, the 64-bit capable packer.  Needs 16-byte aligned blocks
// (the launcher's condition); the last, partial tile of a block is read element-wise from global memory.
#include "fused_detail.cuh"
#include "group_detail.cuh"

namespace mnw {

struct GroupFusedArgs {
    const BlockDesc *descs;
    BlockStat *stats;
    BatchShape sh;
    int64_t *mins, *bits, *offsets, *out_len;
    uint8_t *out;
    long long chain_stride, chain_cap;
    unsigned long long *pub;   // [nblocks] look-back words: flag (2 bits) | bytes; PREFIX also means "block finalised"
    unsigned *done;            // [nblocks] warp-quarters of statistics tiles posted
    unsigned *ticket;          // next ticket
    int64_t *wide_list;        // blocks wider than 32 bits, for k_pack
    int *wide_count;
    int *err;
    float *logws;              // [nblocks][uniform_n] float32(log10 x) of the log10 columns' whole tiles, or null
    int wave_tiles;            // tiles per wave: a whole number of blocks
    int sup;                   // tiles per ticket
    int lag;                   // the pack pass of wave g follows the statistics pass of wave g + lag
    int dry;                   // diagnostics: 1 = consumers skip the pack work, 2 = and the statistics work (ring + TMA only),
                               // 3 = full work without the pack pass's "sure" shortcut
};

constexpr int GF_CW = 4;                          // consumer warps = quarters of a tile
constexpr int GF_THREADS = 32 * (GF_CW + 1);      // + the producer warp
enum : int { J_STAT_F32 = 0, J_STAT_I64, J_PACK_F32, J_PACK_I64, J_SKIP, J_END };

// A job = one tile of one pass, in a ring slot.  The fields that only depend on (block, pass) live in a GProto.
struct __align__(16) GJob {
    int kind, flush, count, whole;   // whole: a full tile, fetched into the slot by TMA
    int pi, pad0;                    // which GProto
    float *aux;                      // log10 column, whole tile: where the statistics pass leaves float32(log10 x) (clamped)
                                     // for the pack pass, so that the logarithm is taken once per value; else null
    const void *src;                 // global address of the tile's first element
    uint8_t *dst;                    // pack: where the tile's packed bytes start
};
struct __align__(16) GProto {
    long long b;
    float low, high, dx, hi_clamp;
    unsigned P, C;
    int flags, fast;     // fast: the unchecked quantiser applies
    long long pmin, min;
    int slow, do_bound, bits;
    int sure;            // pack pass: every value of the block passed the unchecked quantiser in the statistics pass and
                         // no pixel index needs the + pixels of bound(): v = bits + constant, no range test
};
constexpr int GF_NP = 8;   // GProto ring: more than any ring of job slots

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void gmbar_init(unsigned long long *b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void gmbar_wait(unsigned long long *b, unsigned parity) {   // acquire at CTA scope
    const unsigned a = smem_u32(b);
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ bool gmbar_try(unsigned long long *b, unsigned parity) {   // one (time-limited) try
    unsigned done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void gmbar_arrive(unsigned long long *b) {   // release at CTA scope
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// one TMA bulk copy global -> shared, completing `bytes` on the barrier (which also gets the producer's arrival)
__device__ __forceinline__ void tma_fetch(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    const unsigned b = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(b) : "memory");
}

__device__ __forceinline__ QuantP job_quant(const GProto &j) {
    QuantP p;
    p.low = j.low; p.high = j.high; p.dx = j.dx; p.hi_clamp = j.hi_clamp;
    p.rcp = __frcp_rn(j.dx); p.ndx = -j.dx;
    p.P = j.P;
    p.tmax = __float_as_uint(__fsub_rn(j.high, j.low));
    p.flags = j.flags;
    p.fast_ok = j.dx > 0x1p-60f && j.dx < 0x1p60f && j.P >= 2u;
    return p;
}

// position of element el (0..1023) of a warp's quarter in its staging region: rows of 32 values, the 16-byte chunks of
// row L rotated by L & 7 (conflict-free 128-bit stores by float4 index and 128-bit loads by row)
__device__ __forceinline__ int stage_slot(int el) {
    const int L = el >> 5, i = el & 31;
    return (L << 5) + ((((i >> 2) ^ (L & 7)) << 2) | (i & 3));
}

// 32 values per lane (1024 consecutive elements of a block per warp) -> packed bytes at dst (bit.BufferedArray,
// go/bit/bit.go:84-134).  region: the warp's 1024 words of shared memory, used for the transposition.
__device__ __forceinline__ void pack_group_warp(const unsigned (&v)[32], unsigned *region, int gcount, int bits, uint8_t *dst, int lane) {
    switch (bits) {
#define MNW_CASE(B)                                                                         \
    case B: {                                                                               \
        unsigned o[B];                                                                      \
        pack32<B>(v, o);                                                                    \
        __syncwarp();                                                                       \
        _Pragma("unroll") for (int jj = 0; jj < B; jj++) {                                  \
            const int W = lane * B + jj;                                                    \
            region[W ^ (W >> 5)] = o[jj];                                                   \
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

// Running statistics of one consumer warp over the tiles of a ticket (one block at a time).
struct WarpAcc {
    unsigned wmin, wmax, qmin, qmax;
    long long mn, mx;
    bool oob, fell;   // fell: some value went through the checked quantiser
    int n;   // tile quarters accumulated
    __device__ __forceinline__ void reset() { wmin = ~0u; wmax = 0u; qmin = ~0u; qmax = 0u; mn = LLONG_MAX; mx = LLONG_MIN; oob = false; fell = false; n = 0; }
};

}  // namespace

template <int SLOT_BYTES, int NS>
__global__ void __launch_bounds__(GF_THREADS, NS * SLOT_BYTES <= 32768 ? 6 : (NS * SLOT_BYTES <= 49152 ? 4 : (NS * SLOT_BYTES <= 65536 ? 3 : 2))) k_group_fused(GroupFusedArgs A) {
    extern __shared__ __align__(128) unsigned char gsm[];
    unsigned char *slots = gsm;
    GJob *jobs = (GJob *)(gsm + (size_t)NS * SLOT_BYTES);
    GProto *protos = (GProto *)(jobs + NS);
    unsigned long long *full = (unsigned long long *)(protos + GF_NP), *empty = full + NS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = (int)A.sh.total_tiles, Wt = A.wave_tiles, SUP = A.sup;
    const int tpb = (int)((A.sh.uniform_n + PACK_TILE - 1) / PACK_TILE);   // uniform blocks (the launcher's condition)
    const int64_t bpc = A.sh.blocks_per_chain;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; i++) { gmbar_init(&full[i], 1); gmbar_init(&empty[i], GF_CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == GF_CW) {
        // ------------------------------------------------------------------ producer
        if (lane != 0) return;
        const int nsw = (Wt + SUP - 1) / SUP, nwaves = (T + Wt - 1) / Wt;
        const int LAG = A.lag;
        const unsigned ntickets = 2u * (unsigned)(nwaves + LAG) * (unsigned)nsw;
        unsigned s = 0;
        long long cur_b = -1;
        int cur_pack = -1;
        GProto proto = {};   // the fields of a job that only depend on (block, pass)
        const char *cur_src = nullptr;
        float *cur_aux = nullptr;
        unsigned pi = 0;
        bool proto_dirty = false;
        int esz = 4, job_kind = J_SKIP;
        int64_t cur_n = 0;
        int cur_kind = 0, cur_chain = 0;
        long long cur_off = 0;
        unsigned t_next = atomicAdd(A.ticket, 1u);
        for (;;) {
            const unsigned t = t_next;
            if (t >= ntickets) break;
            t_next = atomicAdd(A.ticket, 1u);   // claimed one ticket ahead: the atomic's latency hides behind this ticket's jobs
            const int slot_i = (int)(t / (unsigned)nsw), r = (int)(t - (unsigned)slot_i * (unsigned)nsw), g = slot_i >> 1;
            const int pack = slot_i & 1, wave = pack ? g - LAG : g;
            if (wave < 0 || wave >= nwaves) continue;
            const int tile0 = wave * Wt + r * SUP;
            int tend = tile0 + SUP;
            if (tend > (wave + 1) * Wt) tend = (wave + 1) * Wt;
            if (tend > T) tend = T;
            // does the next ticket of this CTA continue the statistics of the block this ticket ends in?  Then the
            // consumers keep accumulating in registers instead of posting.
            bool next_same = false;
            if (!pack && t_next < ntickets) {
                const int ns_i = (int)(t_next / (unsigned)nsw), nr = (int)(t_next - (unsigned)ns_i * (unsigned)nsw);
                if (!(ns_i & 1) && (ns_i >> 1) < nwaves) {
                    const int ntile = (ns_i >> 1) * Wt + nr * SUP;
                    next_same = ntile < T && ntile / tpb == (tend - 1) / tpb;
                }
            }
            int b = tile0 / tpb, tib = tile0 - b * tpb - 1;
            for (int tile = tile0; tile < tend; tile++) {
                if (++tib == tpb) { tib = 0; b++; }
                if (b != cur_b || pack != cur_pack) {
                    const BlockDesc *gd = &A.descs[b];
                    cur_b = b; cur_pack = pack;
                    cur_n = gd->n; cur_kind = gd->kind; cur_chain = gd->chain;
                    proto.b = b;
                    cur_src = (const char *)gd->src;
                    esz = cur_kind == KIND_F32 ? 4 : 8;
                    cur_aux = (A.logws && cur_kind == KIND_F32 && (gd->flags & F_LOG10)) ? A.logws + (long long)b * A.sh.uniform_n : nullptr;
                    proto_dirty = true;
                    proto.low = gd->low; proto.high = gd->high; proto.dx = gd->dx; proto.hi_clamp = gd->hi_clamp;
                    proto.P = (unsigned)gd->pixels; proto.flags = gd->flags;
                    proto.C = 0; proto.fast = 0; proto.bits = 0; proto.do_bound = 0; proto.slow = 0; proto.sure = 0;
                    if (cur_kind == KIND_F32) {
                        const QuantP qp = job_quant(proto);
                        if (!pack) {
                            const long long q0 = A.stats[b].q0;   // written by the init kernel, an earlier launch
                            const bool q0_ok = (unsigned long long)q0 < (unsigned long long)qp.P;
                            proto.C = q0_ok ? (unsigned)arc_rotation(q0, qp.P) : 0u;
                            proto.fast = qp.fast_ok && qp.P <= (1u << 22) && q0_ok;   // (log10 columns too: the consumers take the logarithm first)
                            proto.do_bound = q0_ok;   // (statistics pass: "element 0 is in range")
                        } else {
                            proto.fast = qp.fast_ok && qp.P <= (1u << 22);
                        }
                    } else {
                        proto.fast = 1;
                    }
                    if (pack) {
                        while ((ld_acquire_u64(A.pub + b) >> 62) != 2) __nanosleep(64);
                        if (cur_aux) asm volatile("fence.proxy.async.global;" ::: "memory");   // the TMA reads what other CTAs stored
                        const BlockStat *gs = &A.stats[b];
                        proto.slow = __ldcg(&gs->slow); proto.pmin = __ldcg(&gs->pmin); proto.min = __ldcg(&gs->min);
                        proto.do_bound = __ldcg(&gs->do_bound); proto.bits = __ldcg(&gs->bits);
                        const long long nbytes = __ldcg(&gs->nbytes);
                        cur_off = __ldcg(&gs->out_off);
                        if (cur_kind == KIND_F32) proto.fast = proto.fast && !proto.slow && proto.do_bound;
                        proto.sure = A.dry != 3 && proto.fast && !(__ldcg(&gs->oob) & 2u) && __ldcg(&gs->qmin) >= proto.pmin;
                        if (proto.bits >= 1 && proto.bits <= 32 && cur_off + nbytes > A.chain_cap) {   // never write past the caller's buffer
                            atomicExch(A.err, 2);
                            proto.bits = 0;
                        }
                    }
                }
                const unsigned slot = s % NS;
                while (!gmbar_try(&empty[slot], ((s / NS) & 1u) ^ 1u)) __nanosleep(200);   // (ring full: the producer is ahead)
                if (proto_dirty) {   // (after the wait: the jobs that used this GProto entry NP changes ago are consumed)
                    pi = (pi + 1) % GF_NP;
                    protos[pi] = proto;
                    proto_dirty = false;
                    job_kind = !pack ? (cur_kind == KIND_F32 ? J_STAT_F32 : J_STAT_I64)
                                     : ((proto.bits >= 1 && proto.bits <= 32) ? (cur_kind == KIND_F32 ? J_PACK_F32 : J_PACK_I64) : J_SKIP);
                }
                GJob j;
                const long long first = (long long)tib * PACK_TILE;
                j.kind = job_kind;
                j.count = (int)(first + PACK_TILE <= cur_n ? PACK_TILE : cur_n - first);
                j.flush = !pack && ((tib == tpb - 1) || (tile == tend - 1 && !next_same));
                j.whole = j.count == PACK_TILE && job_kind != J_SKIP;
                j.pi = (int)pi; j.pad0 = 0;
                j.src = cur_src + first * esz;
                j.aux = nullptr;
                if (cur_aux && j.whole) {
                    float *ap = cur_aux + first;
                    if (((uintptr_t)ap & 15) == 0) j.aux = ap;
                }
                j.dst = A.out + (long long)cur_chain * A.chain_stride + cur_off + ((first * proto.bits) >> 3);
                jobs[slot] = j;
                if (j.whole) tma_fetch(slots + (size_t)slot * SLOT_BYTES, (pack && j.aux) ? (const void *)j.aux : j.src, (unsigned)(PACK_TILE * esz), &full[slot]);
                else gmbar_arrive(&full[slot]);
                s++;
            }
        }
        const unsigned slot = s % NS;
        gmbar_wait(&empty[slot], ((s / NS) & 1u) ^ 1u);
        jobs[slot].kind = J_END;
        gmbar_arrive(&full[slot]);
        return;
    }

    // ---------------------------------------------------------------------- consumers
    WarpAcc acc;
    acc.reset();
    bool aux_stored = false;      // this warp has stored log10 values since its last post
    int pend = 0;                 // stage of the post in flight (warp-uniform)
    long long pend_b = 0;
    unsigned pend_n = 0, pend_prev = 0, pr4 = 0;
    unsigned long long pr0 = 0, pr1 = 0;
    long long pr2 = 0, pr3 = 0;
    // every quarter of every tile of block b has been posted: finalise it, by this warp
    auto finalize = [&](long long b) {
        BlockStat *sb = &A.stats[b];
        const BlockDesc d = A.descs[b];
        int slow = 0;
        if (lane == 0) {
            __threadfence();
            BlockStat st;
            st.wmin = __ldcg(&sb->wmin); st.wmax = __ldcg(&sb->wmax);
            st.qmin = __ldcg(&sb->qmin); st.qmax = __ldcg(&sb->qmax);
            st.q0 = __ldcg(&sb->q0); st.oob = __ldcg(&sb->oob);
            slow = finalize_block(d, st, A.err);
            *sb = st;
        }
        slow = __shfl_sync(0xffffffffu, slow, 0);
        if (slow) slow_block_warp(d, sb, A.err);   // exact sequential periodicMin (rare)
        long long nbytes = 0, mn = 0;
        int bits = 0;
        if (lane == 0) { nbytes = sb->nbytes; mn = sb->min; bits = sb->bits; }
        nbytes = __shfl_sync(0xffffffffu, nbytes, 0);
        const int64_t first = (b / bpc) * bpc;
        long long excl = 0;
        if (b != first) {
            if (lane == 0) st_relaxed(A.pub + b, PUB_AGG | (unsigned long long)nbytes);
            excl = lookback(A.pub, first, b);
        }
        if (lane == 0) {
            sb->out_off = excl;
            if (A.offsets) A.offsets[b] = excl;
            if (A.mins) A.mins[b] = mn;
            if (A.bits) A.bits[b] = bits;
            if (A.out_len && (b == first + bpc - 1 || b == A.sh.nblocks - 1)) A.out_len[b / bpc] = excl + nbytes;
            if (bits > 32) A.wide_list[atomicAdd(A.wide_count, 1)] = b;
            st_release_u64(A.pub + b, PUB_PREFIX | (unsigned long long)(excl + nbytes));
        }
        __syncwarp();
    };
    // one stage of the post pipeline: 1 -> 2 the data atomics are back, count the post; 2 -> 0 was it the block's last?
    auto advance = [&]() {
        if (pend == 1) {
            if (lane == 0) {
                asm volatile("" ::"l"(pr0), "l"(pr1), "l"(pr2), "l"(pr3), "r"(pr4) : "memory");   // the data atomics are back
                pend_prev = atomicAdd(&A.done[pend_b], pend_n);
            }
            pend = 2;
        } else if (pend == 2) {
            int last = 0;
            if (lane == 0) last = pend_prev + pend_n == (unsigned)(GF_CW * tpb);
            last = __shfl_sync(0xffffffffu, last, 0);
            pend = 0;
            if (last) finalize(pend_b);
        }
    };
    for (unsigned s = 0;; s++) {
        const unsigned slot = s % NS;
        // wait for the job; an idle warp pushes its post on (the producer may itself be waiting for that very post:
        // a pack ticket is held back until its block is finalised)
        while (!__all_sync(0xffffffffu, gmbar_try(&full[slot], (s / NS) & 1u))) advance();
        const GJob jb = jobs[slot];
        const int kind = jb.kind;
        if (kind == J_END) break;
        const GProto &j = protos[jb.pi];
        unsigned char *tilep = slots + (size_t)slot * SLOT_BYTES;
        const int count = jb.count;
        const int wcount = count - warp * 1024 < 0 ? 0 : (count - warp * 1024 > 1024 ? 1024 : count - warp * 1024);   // this warp's elements
        const int flush = (kind == J_STAT_F32 || kind == J_STAT_I64) ? jb.flush : 0;
        const long long b = j.b;
        if (A.dry == 2 && (kind == J_STAT_F32 || kind == J_STAT_I64)) {
            acc.n++;
        } else if ((A.dry == 1 || A.dry == 2) && (kind == J_PACK_F32 || kind == J_PACK_I64)) {
        } else if (kind == J_STAT_F32) {
            QuantP qp = job_quant(j);
            const unsigned C = j.C;
            if (!j.do_bound) acc.oob = true;   // element 0 of the block is out of range
            if (jb.whole && jb.aux) {
                // log10 column: the logarithm (and the clamp) once per value -- in place in the slot, and a copy for the
                // pack pass (go/minh/minh.go:141-149); from here on the tile is a plain float tile
                float4 *w4 = (float4 *)(tilep + warp * 4096), *a4 = (float4 *)(jb.aux + warp * 1024);
                const bool clamp = qp.flags & F_CLAMP;
#pragma unroll 2
                for (int i = 0; i < 8; i++) {
                    float4 vv = w4[lane + 32 * i];
                    vv.x = go_log10_f32(vv.x); vv.y = go_log10_f32(vv.y); vv.z = go_log10_f32(vv.z); vv.w = go_log10_f32(vv.w);
                    if (clamp) {
                        vv.x = vv.x < qp.low ? qp.low : vv.x; vv.x = vv.x >= qp.high ? qp.hi_clamp : vv.x;
                        vv.y = vv.y < qp.low ? qp.low : vv.y; vv.y = vv.y >= qp.high ? qp.hi_clamp : vv.y;
                        vv.z = vv.z < qp.low ? qp.low : vv.z; vv.z = vv.z >= qp.high ? qp.hi_clamp : vv.z;
                        vv.w = vv.w < qp.low ? qp.low : vv.w; vv.w = vv.w >= qp.high ? qp.hi_clamp : vv.w;
                    }
                    w4[lane + 32 * i] = vv;
                    __stcg(a4 + lane + 32 * i, vv);
                }
                qp.flags &= ~(unsigned)(F_LOG10 | F_CLAMP);
                aux_stored = true;
            }
            if (jb.whole) {
                const float4 *s4 = (const float4 *)(tilep + warp * 4096);
                const bool clamp = qp.flags & F_CLAMP, islog = qp.flags & F_LOG10;
                const unsigned long long LOW2 = f2_pack(qp.low, qp.low), RCP2 = f2_pack(qp.rcp, qp.rcp), NDX2 = f2_pack(qp.ndx, qp.ndx);
                const unsigned Cm = C - FQ_MAGIC, nP = 0u - qp.P;
                unsigned bmin = ~0u, bmax = 0u, fwmin = ~0u, fwmax = 0u;
                if (j.fast) {
#pragma unroll 2
                for (int i = 0; i < 8; i++) {
                    const float4 vv = s4[lane + 32 * i];
                    float x0 = vv.x, x1 = vv.y, x2 = vv.z, x3 = vv.w;
                    if (islog) { x0 = go_log10_f32(x0); x1 = go_log10_f32(x1); x2 = go_log10_f32(x2); x3 = go_log10_f32(x3); }   // go/minh/minh.go:143
                    if (clamp) {   // go/minh/minh.go:144-147
                        x0 = x0 < qp.low ? qp.low : x0; x0 = x0 >= qp.high ? qp.hi_clamp : x0;
                        x1 = x1 < qp.low ? qp.low : x1; x1 = x1 >= qp.high ? qp.hi_clamp : x1;
                        x2 = x2 < qp.low ? qp.low : x2; x2 = x2 >= qp.high ? qp.hi_clamp : x2;
                        x3 = x3 < qp.low ? qp.low : x3; x3 = x3 >= qp.high ? qp.hi_clamp : x3;
                    }
                    unsigned b0, b1, b2, b3;
                    f2_bits(quantize2(f2_pack(x0, x1), LOW2, RCP2, NDX2), b0, b1);
                    f2_bits(quantize2(f2_pack(x2, x3), LOW2, RCP2, NDX2), b2, b3);
                    const unsigned t0 = b0 + Cm, t1 = b1 + Cm, t2 = b2 + Cm, t3 = b3 + Cm;
                    const unsigned w0 = min(t0, t0 + nP), w1 = min(t1, t1 + nP), w2 = min(t2, t2 + nP), w3 = min(t3, t3 + nP);
                    bmin = __vimin3_u32(bmin, b0, b1); bmin = __vimin3_u32(bmin, b2, b3);
                    bmax = __vimax3_u32(bmax, b0, b1); bmax = __vimax3_u32(bmax, b2, b3);
                    fwmin = __vimin3_u32(fwmin, w0, w1); fwmin = __vimin3_u32(fwmin, w2, w3);
                    fwmax = __vimax3_u32(fwmax, w0, w1); fwmax = __vimax3_u32(fwmax, w2, w3);
                }
                }
                if (j.fast && bmin >= FQ_MAGIC && bmax < FQ_MAGIC + qp.P) {
                    acc.wmin = min(acc.wmin, fwmin); acc.wmax = max(acc.wmax, fwmax);
                    acc.qmin = min(acc.qmin, bmin - FQ_MAGIC); acc.qmax = max(acc.qmax, bmax - FQ_MAGIC);
                } else {   // log10 columns, pixels > 2^22, or (rare) a value the unchecked quantiser does not vouch for: checked
                    acc.fell = true;
#pragma unroll 1
                    for (int i = 0; i < 8; i++) {
                        const float4 vv = s4[lane + 32 * i];   // (read again: no dynamic indexing of v[])
                        const float x[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            const unsigned q = quant_elem(x[c], qp, acc.oob, nullptr);
                            unsigned w = q + C;
                            w = min(w, w - qp.P);
                            acc.wmin = min(acc.wmin, w); acc.wmax = max(acc.wmax, w); acc.qmin = min(acc.qmin, q); acc.qmax = max(acc.qmax, q);
                        }
                    }
                }
            } else {   // partial tile: element-wise from global memory
                const float *gp = (const float *)jb.src + warp * 1024;
                acc.fell = true;
#pragma unroll 1
                for (int el = lane; el < wcount; el += 32) {
                    const unsigned q = quant_elem(__ldcg(gp + el), qp, acc.oob, nullptr);
                    unsigned w = q + C;
                    w = min(w, w - qp.P);
                    acc.wmin = min(acc.wmin, w); acc.wmax = max(acc.wmax, w); acc.qmin = min(acc.qmin, q); acc.qmax = max(acc.qmax, q);
                }
            }
            acc.n++;
        } else if (kind == J_STAT_I64) {
            if (jb.whole) {
                const longlong2 *s2 = (const longlong2 *)(tilep + warp * 8192);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    longlong2 v[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) v[i] = s2[lane + 32 * (8 * h + i)];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        acc.mn = v[i].x < acc.mn ? v[i].x : acc.mn; acc.mx = v[i].x > acc.mx ? v[i].x : acc.mx;
                        acc.mn = v[i].y < acc.mn ? v[i].y : acc.mn; acc.mx = v[i].y > acc.mx ? v[i].y : acc.mx;
                    }
                }
            } else {
                const long long *gp = (const long long *)jb.src + warp * 1024;
#pragma unroll 1
                for (int el = lane; el < wcount; el += 32) {
                    const long long x = __ldcg(gp + el);
                    acc.mn = x < acc.mn ? x : acc.mn; acc.mx = x > acc.mx ? x : acc.mx;
                }
            }
            acc.n++;
        } else if (kind == J_PACK_F32 || kind == J_PACK_I64) {
            unsigned *region = (unsigned *)(tilep + warp * (kind == J_PACK_F32 ? 4096 : 8192));   // the warp's 1024 words: staging + transposition
            const int bits = j.bits;
            uint8_t *dst = jb.dst + (((long long)warp * 1024 * bits) >> 3);
            unsigned v[32];
            if (kind == J_PACK_F32) {
                QuantP qp = job_quant(j);
                if (jb.whole && jb.aux) qp.flags &= ~(unsigned)(F_LOG10 | F_CLAMP);   // the tile came from the statistics pass's copy
                const long long pmin = j.pmin, mn = j.min;
                const int slow = j.slow, do_bound = j.do_bound;
                const unsigned mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
                // one element through the checked quantiser -> the value to pack
                auto checked_value = [&](float xv) -> unsigned {
                    bool oob = false;
                    long long raw;
                    const unsigned q = quant_elem(xv, qp, oob, &raw);
                    if (!slow) {   // folded index; bound(q, pmin, pixels) - min (go/group.go:323, :246-247)
                        const long long qb = (long long)q < pmin ? (long long)q + (long long)qp.P : (long long)q;
                        return (unsigned)(qb - mn);
                    }
                    // the block holds out-of-range indices: the reference's own int64 arithmetic
                    const long long qb = do_bound ? bound1(raw, pmin, (long long)qp.P) : raw;
                    return (unsigned)((unsigned long long)qb - (unsigned long long)mn) & mask;
                };
                if (jb.whole) {
                    const bool fast = j.fast;
                    const float4 *s4 = (const float4 *)(tilep + warp * 4096);
                    const bool clamp = qp.flags & F_CLAMP, islog = qp.flags & F_LOG10;
                    const unsigned long long LOW2 = f2_pack(qp.low, qp.low), RCP2 = f2_pack(qp.rcp, qp.rcp), NDX2 = f2_pack(qp.ndx, qp.ndx);
                    const unsigned dsub = 0u - FQ_MAGIC - (unsigned)pmin, cadd = (unsigned)(pmin - mn);
                    // step i: the warp reads the four 128-byte rows 4i .. 4i+3 of its quarter and stages their values into
                    // the same rows (the chunk rotation stays within a row): in place, one __syncwarp per step
                    if (j.sure) {   // the statistics pass vouches for every value and bound() adds nothing: v = bits + constant
                        const unsigned cst = dsub + cadd;
#pragma unroll 2
                        for (int i = 0; i < 8; i++) {
                            const float4 xx = s4[lane + 32 * i];
                            __syncwarp();
                            float x0 = xx.x, x1 = xx.y, x2 = xx.z, x3 = xx.w;
                            if (islog) { x0 = go_log10_f32(x0); x1 = go_log10_f32(x1); x2 = go_log10_f32(x2); x3 = go_log10_f32(x3); }
                            if (clamp) {
                                x0 = x0 < qp.low ? qp.low : x0; x0 = x0 >= qp.high ? qp.hi_clamp : x0;
                                x1 = x1 < qp.low ? qp.low : x1; x1 = x1 >= qp.high ? qp.hi_clamp : x1;
                                x2 = x2 < qp.low ? qp.low : x2; x2 = x2 >= qp.high ? qp.hi_clamp : x2;
                                x3 = x3 < qp.low ? qp.low : x3; x3 = x3 >= qp.high ? qp.hi_clamp : x3;
                            }
                            unsigned b0, b1, b2, b3;
                            f2_bits(quantize2(f2_pack(x0, x1), LOW2, RCP2, NDX2), b0, b1);
                            f2_bits(quantize2(f2_pack(x2, x3), LOW2, RCP2, NDX2), b2, b3);
                            const int iv = lane + 32 * i, L = iv >> 3;
                            *(uint4 *)&region[(L << 5) + (((iv & 7) ^ (L & 7)) << 2)] = make_uint4(b0 + cst, b1 + cst, b2 + cst, b3 + cst);
                        }
                    } else
#pragma unroll 2
                    for (int i = 0; i < 8; i++) {
                        const float4 xx = s4[lane + 32 * i];
                        __syncwarp();
                        float x0 = xx.x, x1 = xx.y, x2 = xx.z, x3 = xx.w;
                        if (islog) { x0 = go_log10_f32(x0); x1 = go_log10_f32(x1); x2 = go_log10_f32(x2); x3 = go_log10_f32(x3); }
                        if (clamp) {
                            x0 = x0 < qp.low ? qp.low : x0; x0 = x0 >= qp.high ? qp.hi_clamp : x0;
                            x1 = x1 < qp.low ? qp.low : x1; x1 = x1 >= qp.high ? qp.hi_clamp : x1;
                            x2 = x2 < qp.low ? qp.low : x2; x2 = x2 >= qp.high ? qp.hi_clamp : x2;
                            x3 = x3 < qp.low ? qp.low : x3; x3 = x3 >= qp.high ? qp.hi_clamp : x3;
                        }
                        unsigned b0, b1, b2, b3;
                        f2_bits(quantize2(f2_pack(x0, x1), LOW2, RCP2, NDX2), b0, b1);
                        f2_bits(quantize2(f2_pack(x2, x3), LOW2, RCP2, NDX2), b2, b3);
                        const unsigned lo = __vimin3_u32(b0, b1, min(b2, b3)), hi = __vimax3_u32(b0, b1, max(b2, b3));
                        uint4 rr;
                        if (fast && lo >= FQ_MAGIC && hi < FQ_MAGIC + qp.P) {
                            const unsigned d0 = b0 + dsub, d1 = b1 + dsub, d2 = b2 + dsub, d3 = b3 + dsub;   // q - pmin
                            rr.x = min(d0, d0 + qp.P) + cadd; rr.y = min(d1, d1 + qp.P) + cadd;
                            rr.z = min(d2, d2 + qp.P) + cadd; rr.w = min(d3, d3 + qp.P) + cadd;
                        } else {
                            rr.x = checked_value(xx.x); rr.y = checked_value(xx.y);
                            rr.z = checked_value(xx.z); rr.w = checked_value(xx.w);
                        }
                        const int iv = lane + 32 * i, L = iv >> 3;
                        *(uint4 *)&region[(L << 5) + (((iv & 7) ^ (L & 7)) << 2)] = rr;
                    }
                } else {   // partial tile: element-wise from global memory into the staging layout
                    const float *gp = (const float *)jb.src + warp * 1024;
#pragma unroll 1
                    for (int el = lane; el < 1024; el += 32) region[stage_slot(el)] = el < wcount ? checked_value(__ldcs(gp + el)) : 0u;
                }
            } else {
                const unsigned long long mn = (unsigned long long)j.min;
                if (jb.whole) {
                    // 1024 int64 = 8 KB per warp, in two halves: the 32-bit differences of a half land in the first
                    // 4 KB, which has been read completely (half 0) or never held the other half's input (half 1)
                    const longlong2 *s2 = (const longlong2 *)(tilep + warp * 8192);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        longlong2 x2[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) x2[i] = s2[lane + 32 * (8 * h + i)];
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int el = 2 * (lane + 32 * (8 * h + i));   // an aligned pair shares a 16-byte chunk
                            *(uint2 *)&region[stage_slot(el)] =
                                make_uint2((unsigned)((unsigned long long)x2[i].x - mn), (unsigned)((unsigned long long)x2[i].y - mn));   // go/group.go:246-247
                        }
                        __syncwarp();
                    }
                } else {
                    const long long *gp = (const long long *)jb.src + warp * 1024;
#pragma unroll 1
                    for (int el = lane; el < 1024; el += 32) region[stage_slot(el)] = el < wcount ? (unsigned)((unsigned long long)__ldcs(gp + el) - mn) : 0u;
                }
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; c++) {   // this lane's 32 consecutive values
                const uint4 rr = *(const uint4 *)&region[lane * 32 + ((c ^ (lane & 7)) << 2)];
                v[4 * c] = rr.x; v[4 * c + 1] = rr.y; v[4 * c + 2] = rr.z; v[4 * c + 3] = rr.w;
            }
            if (wcount > 0) pack_group_warp(v, region, wcount, bits, dst, lane);
        }
        // ---- this warp is done with the slot
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this warp's writes to the slot precede the next TMA fill
        if (lane == 0) gmbar_arrive(&empty[slot]);
        // ---- the statistics post of this warp, pipelined over the jobs that follow so that the warp never waits for
        // an atomic: stage 1 = the data atomics (with return values: once they are back the atomics are performed),
        // stage 2 = the block's counter, stage 3 = "was this the last post of the block?"
        advance();
        if (flush) {
            if (aux_stored) {   // the stored logarithms are performed (and visible to the async proxy) before this post counts
                __threadfence();
                asm volatile("fence.proxy.async.global;" ::: "memory");
                __syncwarp();
                aux_stored = false;
            }
            while (pend) advance();   // consecutive posts (short tickets, block ends): finish the earlier one first
            BlockStat *sb = &A.stats[b];
            if (kind == J_STAT_F32) {
                const unsigned wmin = __reduce_min_sync(0xffffffffu, acc.wmin), wmax = __reduce_max_sync(0xffffffffu, acc.wmax);
                const unsigned qmin = __reduce_min_sync(0xffffffffu, acc.qmin), qmax = __reduce_max_sync(0xffffffffu, acc.qmax);
                const unsigned ob = (__any_sync(0xffffffffu, acc.oob) ? 1u : 0u) | (__any_sync(0xffffffffu, acc.fell) ? 2u : 0u);
                if (lane == 0) {
                    pr0 = pr1 = 0; pr2 = pr3 = 0; pr4 = 0;
                    if (wmin <= wmax) {
                        pr0 = atomicMin(&sb->wmin, (unsigned long long)wmin); pr1 = atomicMax(&sb->wmax, (unsigned long long)wmax);
                        pr2 = atomicMin(&sb->qmin, (long long)qmin); pr3 = atomicMax(&sb->qmax, (long long)qmax);
                    }
                    if (ob) pr4 = atomicOr(&sb->oob, ob);
                }
            } else {
                const long long mn = warp_min_ll(acc.mn), mx = warp_max_ll(acc.mx);
                if (lane == 0) {
                    pr0 = pr1 = 0; pr2 = pr3 = 0; pr4 = 0;
                    if (mn <= mx) { pr2 = atomicMin(&sb->qmin, mn); pr3 = atomicMax(&sb->qmax, mx); }
                }
            }
            pend = 1; pend_b = b; pend_n = (unsigned)acc.n;
            acc.reset();
        }
    }
    while (pend) advance();   // the posts still in flight
}

size_t group_fused_ws_bytes(int64_t nblocks) { return (size_t)nblocks * 24 + 256; }

template <int SLOT_BYTES, int NS>
static cudaError_t launch_group_fused_t(Launcher &L, GroupFusedArgs &A, int64_t tile_bytes) {
    auto kern = k_group_fused<SLOT_BYTES, NS>;
    const size_t smem = (size_t)NS * SLOT_BYTES + NS * sizeof(GJob) + GF_NP * sizeof(GProto) + 2 * NS * sizeof(unsigned long long);
    static DevCfg cfgs[MNW_MAX_DEVICES];
    DevCfg &dc = dev_cfg(cfgs);
    std::call_once(dc.once, [&] {
        int dev = 0, sms = 148, per = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        dc.err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (dc.err != cudaSuccess) return;
        dc.err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, GF_THREADS, smem);
        if (dc.err != cudaSuccess) return;
        if (per < 1) { dc.err = cudaErrorLaunchOutOfResources; return; }
        dc.a = per * sms;
        if (getenv("MNW_DEBUG")) fprintf(stderr, "k_group_fused<%d, %d>: %d co-resident CTAs, %zu B dynamic smem\n", SLOT_BYTES, NS, dc.a, smem);
    });
    if (dc.err != cudaSuccess) return dc.err;
    const BatchShape &sh = A.sh;
    // tuning knobs: MNW_GROUP_SUP tiles per ticket, MNW_GROUP_WAVE_MB megabytes of input per wave
    static const int sup_knob = getenv("MNW_GROUP_SUP") ? atoi(getenv("MNW_GROUP_SUP")) : 0;
    static const int wave_knob = getenv("MNW_GROUP_WAVE_MB") ? atoi(getenv("MNW_GROUP_WAVE_MB")) : 0;
    static const int dry_knob = getenv("MNW_GROUP_DRY") ? atoi(getenv("MNW_GROUP_DRY")) : 0;
    A.dry = dry_knob;
    // a wave is a whole number of blocks: every statistics tile of a block then has a smaller ticket than any of the
    // block's pack tiles (a pack job only ever waits for smaller tickets)
    const long long tpb = (sh.uniform_n + PACK_TILE - 1) / PACK_TILE;
    long long wave = (long long)(wave_knob > 0 ? wave_knob : 128) * (1 << 20) / tile_bytes;
    wave = (wave + tpb - 1) / tpb * tpb;
    if (wave > sh.total_tiles) wave = sh.total_tiles;
    A.wave_tiles = (int)wave;
    // tiles per ticket: 4 amortise the ticket and the flush for a large batch; a small batch (fewer tickets than
    // co-resident CTAs) is spread over more CTAs instead (measured on 16 blocks of 2^16: 0.052 -> 0.043 ms)
    A.sup = A.logws ? 1 : 4;   // (log10 columns: a tile is ~3x the work, single-tile tickets balance better: 0.53 -> 0.48 ms)
    while (A.sup > 1 && wave / A.sup < dc.a) A.sup >>= 1;
    if (sup_knob > 0) A.sup = sup_knob;
    const long long nsw = (wave + A.sup - 1) / A.sup, nwaves = (sh.total_tiles + wave - 1) / wave;
    static const int lag_knob = getenv("MNW_GROUP_LAG") ? atoi(getenv("MNW_GROUP_LAG")) : 0;
    A.lag = lag_knob > 0 ? lag_knob : 1;
    const long long ntickets = 2 * (nwaves + A.lag) * nsw;
    if (ntickets >= (1LL << 31)) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)(ntickets < dc.a ? ntickets : dc.a);
    L.begin("k_group_fused");
    kern<<<grid, GF_THREADS, smem, L.stream>>>(A);
    L.end();
    L.count++;
    return cudaGetLastError();
}

// Fused encode of a batch of UNIFORM contiguous blocks (sh.uniform_n > 0; every float block periodic with
// 1 <= pixels < 2^31; every block 16-byte aligned).  descs are ready on the stream.  ws: group_fused_ws_bytes(nblocks).
cudaError_t launch_group_encode(Launcher &L, const BlockDesc *descs, BlockStat *stats, const BatchShape &sh, int *flags,
                                int64_t *mins, int64_t *bits, int64_t *offsets, int64_t *out_len, uint8_t *out,
                                int64_t chain_stride, int64_t chain_cap, void *ws, bool has_i64, bool prepared, void *logws) {
    if (sh.nblocks == 0 || sh.total_tiles == 0) return cudaSuccess;
    // prepared: the descriptor kernel has initialised the statistics records and zeroed ws already
    if (!prepared) launch_init_stats(L, descs, stats, sh.nblocks);
    GroupFusedArgs A = {};
    A.descs = descs; A.stats = stats; A.sh = sh;
    A.mins = mins; A.bits = bits; A.offsets = offsets; A.out_len = out_len; A.out = out;
    A.chain_stride = chain_stride; A.chain_cap = chain_cap;
    A.logws = (float *)logws;
    A.pub = (unsigned long long *)ws;
    A.wide_list = (int64_t *)(A.pub + sh.nblocks);
    A.done = (unsigned *)(A.wide_list + sh.nblocks);
    A.err = flags + FLAG_ERR; A.wide_count = flags + 3; A.ticket = (unsigned *)(flags + 4);
    cudaError_t e = prepared ? cudaSuccess : cudaMemsetAsync(ws, 0, group_fused_ws_bytes(sh.nblocks), L.stream);
    if (e != cudaSuccess) return e;
    static const int ns_knob = getenv("MNW_GROUP_NS") ? atoi(getenv("MNW_GROUP_NS")) : 0;   // tuning knob: ring depth
    if (has_i64) e = ns_knob == 2 ? launch_group_fused_t<32768, 2>(L, A, 32768) : launch_group_fused_t<32768, 3>(L, A, 32768);
    // (log10 columns: a ring of 3 slots = 4 CTAs per SM hides the FP64 latency a little better: 0.469 -> 0.450 ms)
    else e = ns_knob == 6 ? launch_group_fused_t<16384, 6>(L, A, 16384) : ((ns_knob == 3 || (ns_knob == 0 && A.logws)) ? launch_group_fused_t<16384, 3>(L, A, 16384) :
             (ns_knob == 2 ? launch_group_fused_t<16384, 2>(L, A, 16384) : launch_group_fused_t<16384, 4>(L, A, 16384)));
    if (e != cudaSuccess) return e;
    // blocks wider than 32 bits, with the fused kernel's (min, bits, offset): the 64-bit capable packer
    launch_pack_list(L, descs, stats, sh, A.wide_list, A.wide_count, out, chain_stride, chain_cap, flags + FLAG_ERR);
    return cudaGetLastError();
}

}  // namespace mnw
