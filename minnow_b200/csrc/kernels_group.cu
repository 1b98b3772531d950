// kernels_group.cu -- k_group_fused: the single-DRAM-read encoder of contiguous int64 / float32 column blocks
// (intGroup.writeData go/group.go:242-255, floatGroup.writeData :312-327, blockIndex go/block_index.go:16-35).
//
// The two passes a block needs -- statistics (min, max, periodic arc -> min, bits, nbytes) and packing -- run in
// ONE persistent kernel, a whole "wave" of tiles apart, so that the packing pass finds its input in the 126 MB L2
// instead of HBM:
//
//   ticket order   S(0) S(1) P(0) S(2) P(1) ... S(G-1) P(G-2) P(G-1)        S(g) / P(g): the statistics / pack
//                                                                           tiles (4096 elements) of wave g
//   * tickets are claimed from an atomic counter, one ahead of the tile being worked on: a CTA only ever waits for
//     tiles with SMALLER tickets, which are held by CTAs that are running -- no deadlock under any residency, no
//     cooperative launch;
//   * the CTA that finishes the last statistics tile of a block finalises it (closed-form periodicMin, min, bits,
//     nbytes; the exact sequential periodicMin for blocks with out-of-range values), publishes the block's size and
//     obtains its byte offset by decoupled look-back over the earlier blocks of its group (chain);
//   * pack tiles wait for their block's PREFIX word (normally published a wave earlier), then quantise again from
//     L2, bound / subtract min, and pack with the compile-time warp packer (pack.cuh).
//
// Algorithmic bytes per element: 4 + bits/8 (float32), 8 + bits/8 (int64); DRAM traffic is the same as long as a wave
// (a few tens of MB) stays L2-resident between its two passes.  Blocks wider than 32 bits (NaN, huge ranges) are
// listed for k_pack, the 64-bit capable packer.
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
    unsigned *done;            // [nblocks] statistics tiles finished
    unsigned *ticket;          // next ticket
    int64_t *wide_list;        // blocks wider than 32 bits, for k_pack
    int *wide_count;
    int *err;
    long long wave;            // tiles per wave
};

constexpr int GF_THREADS = FPACK_THREADS;   // 128: one warp per 1024-element pack group of a tile

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Statistics of one tile of a contiguous float32 block (periodic group, pixels < 2^31): rotated-arc min / max and
// index min / max, accumulated into the block's BlockStat with atomics.  Loads stay in L2 (ld.cg) for the pack pass.
__device__ __forceinline__ void stats_tile_f32(const BlockDesc &d, BlockStat *sb, int64_t tile_in_block, unsigned (*s_r)[5]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t first = tile_in_block * PACK_TILE;
    const int count = (int)((first + PACK_TILE < d.n ? first + PACK_TILE : d.n) - first);
    const QuantP qp = quant_params(d);
    const long long q0 = sb->q0;   // written by the init kernel, an earlier launch
    const bool q0_ok = (unsigned long long)q0 < (unsigned long long)qp.P;
    const unsigned C = q0_ok ? (unsigned)arc_rotation(q0, qp.P) : 0u;
    unsigned wmin = ~0u, wmax = 0u, qmin = ~0u, qmax = 0u;
    bool oob = !q0_ok;
    const float *p = (const float *)d.src + first;
    const int a = (int)(((uintptr_t)p & 15) >> 2);          // elements of the first 16 bytes that precede the tile
    const float4 *base4 = (const float4 *)(p - a);
    const int nvec = (a + count + 3) >> 2;
    auto checked4 = [&](const float4 v4, int iv) {
        const float x[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int el = 4 * iv + c - a;
            if (el >= 0 && el < count) {
                const unsigned q = quant_elem(x[c], qp, oob, nullptr);
                unsigned w = q + C;
                w = min(w, w - qp.P);
                wmin = min(wmin, w); wmax = max(wmax, w); qmin = min(qmin, q); qmax = max(qmax, q);
            }
        }
    };
    if (a == 0 && count == PACK_TILE && quant_bits_ok(qp) && q0_ok) {
        // whole aligned tile: all 8 float4 of the thread in flight, the unchecked quantiser, ONE range test per thread
        // on the min / max of the raw bits; a thread that fails it (rare) takes its vectors again, checked
        constexpr int NV = PACK_TILE / 4 / GF_THREADS;   // 8
        float4 v[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) v[i] = __ldcg(base4 + threadIdx.x + i * GF_THREADS);
        const bool clamp = qp.flags & F_CLAMP;
        const unsigned Cm = C - FQ_MAGIC, nP = 0u - qp.P;
        unsigned bmin = ~0u, bmax = 0u, fwmin = ~0u, fwmax = 0u;
#pragma unroll
        for (int i = 0; i < NV; i++) {
            const unsigned b0 = quant_bits(v[i].x, qp, clamp), b1 = quant_bits(v[i].y, qp, clamp);
            const unsigned b2 = quant_bits(v[i].z, qp, clamp), b3 = quant_bits(v[i].w, qp, clamp);
            const unsigned t0 = b0 + Cm, t1 = b1 + Cm, t2 = b2 + Cm, t3 = b3 + Cm;
            const unsigned w0 = min(t0, t0 + nP), w1 = min(t1, t1 + nP), w2 = min(t2, t2 + nP), w3 = min(t3, t3 + nP);
            bmin = __vimin3_u32(bmin, b0, b1); bmin = __vimin3_u32(bmin, b2, b3);
            bmax = __vimax3_u32(bmax, b0, b1); bmax = __vimax3_u32(bmax, b2, b3);
            fwmin = __vimin3_u32(fwmin, w0, w1); fwmin = __vimin3_u32(fwmin, w2, w3);
            fwmax = __vimax3_u32(fwmax, w0, w1); fwmax = __vimax3_u32(fwmax, w2, w3);
        }
        if (bmin >= FQ_MAGIC && bmax < FQ_MAGIC + qp.P) {
            wmin = fwmin; wmax = fwmax; qmin = bmin - FQ_MAGIC; qmax = bmax - FQ_MAGIC;
        } else {
#pragma unroll
            for (int i = 0; i < NV; i++) checked4(v[i], threadIdx.x + i * GF_THREADS);
        }
    } else {
        for (int iv = threadIdx.x; iv < nvec; iv += GF_THREADS) checked4(__ldcg(base4 + iv), iv);
    }
    wmin = __reduce_min_sync(0xffffffffu, wmin); wmax = __reduce_max_sync(0xffffffffu, wmax);
    qmin = __reduce_min_sync(0xffffffffu, qmin); qmax = __reduce_max_sync(0xffffffffu, qmax);
    const unsigned ob = __any_sync(0xffffffffu, oob);
    if (lane == 0) { s_r[warp][0] = wmin; s_r[warp][1] = wmax; s_r[warp][2] = qmin; s_r[warp][3] = qmax; s_r[warp][4] = ob; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned o = 0;
        for (int wi = 0; wi < GF_THREADS / 32; wi++) {
            wmin = min(wmin, s_r[wi][0]); wmax = max(wmax, s_r[wi][1]);
            qmin = min(qmin, s_r[wi][2]); qmax = max(qmax, s_r[wi][3]);
            o |= s_r[wi][4];
        }
        if (wmin <= wmax) {
            atomicMin(&sb->wmin, (unsigned long long)wmin); atomicMax(&sb->wmax, (unsigned long long)wmax);
            atomicMin(&sb->qmin, (long long)qmin); atomicMax(&sb->qmax, (long long)qmax);
        }
        if (o) atomicOr(&sb->oob, 1u);
    }
}

// int64Min + the max of ArrayBuffer.Bits over one tile of a contiguous int64 block.
__device__ __forceinline__ void stats_tile_i64(const BlockDesc &d, BlockStat *sb, int64_t tile_in_block, long long (*s_l)[2]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t first = tile_in_block * PACK_TILE;
    const int count = (int)((first + PACK_TILE < d.n ? first + PACK_TILE : d.n) - first);
    const long long *p = (const long long *)d.src + first;
    const int a = (int)(((uintptr_t)p & 15) >> 3);          // 1: the tile starts in the upper half of a 16-byte pair
    const longlong2 *base2 = (const longlong2 *)(p - a);
    const int nvec = (a + count + 1) >> 1;
    long long mn = LLONG_MAX, mx = LLONG_MIN;
    if (a == 0 && count == PACK_TILE) {
        constexpr int NV = PACK_TILE / 2 / GF_THREADS;   // 16
#pragma unroll
        for (int h = 0; h < 2; h++) {
            longlong2 v[NV / 2];
#pragma unroll
            for (int i = 0; i < NV / 2; i++) v[i] = __ldcg(base2 + threadIdx.x + (h * NV / 2 + i) * GF_THREADS);
#pragma unroll
            for (int i = 0; i < NV / 2; i++) {
                mn = v[i].x < mn ? v[i].x : mn; mx = v[i].x > mx ? v[i].x : mx;
                mn = v[i].y < mn ? v[i].y : mn; mx = v[i].y > mx ? v[i].y : mx;
            }
        }
    } else {
        for (int iv = threadIdx.x; iv < nvec; iv += GF_THREADS) {
            const longlong2 v = __ldcg(base2 + iv);
            const int e0 = 2 * iv - a;
            if (e0 >= 0) { mn = v.x < mn ? v.x : mn; mx = v.x > mx ? v.x : mx; }
            if (e0 + 1 < count) { mn = v.y < mn ? v.y : mn; mx = v.y > mx ? v.y : mx; }
        }
    }
    mn = warp_min_ll(mn); mx = warp_max_ll(mx);
    if (lane == 0) { s_l[warp][0] = mn; s_l[warp][1] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int wi = 1; wi < GF_THREADS / 32; wi++) {
            mn = s_l[wi][0] < mn ? s_l[wi][0] : mn;
            mx = s_l[wi][1] > mx ? s_l[wi][1] : mx;
        }
        if (count > 0) { atomicMin(&sb->qmin, mn); atomicMax(&sb->qmax, mx); }
    }
}

}  // namespace

__global__ void __launch_bounds__(GF_THREADS, 6) k_group_fused(GroupFusedArgs A) {
    __shared__ __align__(16) unsigned sv[PACK_TILE];   // pack staging (16 KB)
    __shared__ __align__(16) BlockStat s_st;
    __shared__ unsigned s_r[GF_THREADS / 32][5];
    __shared__ long long s_l[GF_THREADS / 32][2];
    __shared__ long long s_pmin;
    __shared__ long long s_red[2][8];
    __shared__ long long s_ticket[2];
    __shared__ int s_flag;   // 1: this CTA finalises the block; 2: and the block is slow
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long T = A.sh.total_tiles, W = A.wave, nwaves = (T + W - 1) / W;
    const long long ntickets = 2 * (nwaves + 1) * W;
    const int64_t tpb = (A.sh.uniform_n + PACK_TILE - 1) / PACK_TILE;   // uniform blocks only (launcher's condition)
    const int64_t bpc = A.sh.blocks_per_chain;

    if (threadIdx.x == 0) s_ticket[0] = (long long)atomicAdd(A.ticket, 1u);
    __syncthreads();
    for (int it = 0;; it ^= 1) {
        const long long t = s_ticket[it];
        if (t >= ntickets) break;
        // claim the NEXT ticket now: the atomic's latency hides behind this tile
        if (threadIdx.x == 0) s_ticket[it ^ 1] = (long long)atomicAdd(A.ticket, 1u);
        const long long slot = t / W, r = t - slot * W, g = slot >> 1;
        const bool pack = slot & 1;
        const long long tile = pack ? (g - 1) * W + r : g * W + r;
        const bool live = pack ? (g >= 1 && tile < T) : (g < nwaves && tile < T);
        if (live) {
            const int64_t b = tile / tpb;
            const BlockDesc d = A.descs[b];
            const int64_t tib = tile - b * tpb;
            if (!pack) {
                if (d.kind == KIND_F32) stats_tile_f32(d, &A.stats[b], tib, s_r);
                else stats_tile_i64(d, &A.stats[b], tib, s_l);
                if (threadIdx.x == 0) {   // thread 0 did the atomics of this tile
                    __threadfence();
                    const unsigned prev = atomicAdd(&A.done[b], 1u);
                    int flag = 0;
                    if (prev + 1 == (unsigned)tpb) {   // last tile of the block: finalise
                        __threadfence();
                        BlockStat s;
                        const BlockStat *gs = &A.stats[b];
                        s.wmin = __ldcg(&gs->wmin); s.wmax = __ldcg(&gs->wmax);
                        s.qmin = __ldcg(&gs->qmin); s.qmax = __ldcg(&gs->qmax);
                        s.q0 = __ldcg(&gs->q0); s.oob = __ldcg(&gs->oob);
                        flag = finalize_block(d, s, A.err) ? 2 : 1;
                        s_st = s;
                        A.stats[b] = s;
                    }
                    s_flag = flag;
                }
                __syncthreads();
                const int flag = s_flag;
                if (flag == 2) {   // exact sequential periodicMin, by the whole CTA (rare)
                    slow_block(d, &A.stats[b], A.err, &s_pmin, s_red);
                    if (threadIdx.x == 0) s_st = A.stats[b];
                    __syncthreads();
                }
                if (flag && warp == 0) {
                    const long long nbytes = s_st.nbytes;
                    const int64_t first = (b / bpc) * bpc;
                    long long excl = 0;
                    if (b != first) {
                        if (lane == 0) st_relaxed(A.pub + b, PUB_AGG | (unsigned long long)nbytes);
                        excl = lookback(A.pub, first, b);
                    }
                    if (lane == 0) {
                        A.stats[b].out_off = excl;
                        if (A.offsets) A.offsets[b] = excl;
                        if (A.mins) A.mins[b] = s_st.min;
                        if (A.bits) A.bits[b] = s_st.bits;
                        if (A.out_len && (b == first + bpc - 1 || b == A.sh.nblocks - 1)) A.out_len[b / bpc] = excl + nbytes;
                        if (s_st.bits > 32) A.wide_list[atomicAdd(A.wide_count, 1)] = b;
                        st_release_u64(A.pub + b, PUB_PREFIX | (unsigned long long)(excl + nbytes));
                    }
                }
                __syncthreads();
            } else {
                if (threadIdx.x == 0) {
                    while ((ld_acquire_u64(A.pub + b) >> 62) != 2) __nanosleep(64);
                    const BlockStat *gs = &A.stats[b];
                    BlockStat s;
                    s.slow = __ldcg(&gs->slow); s.pmin = __ldcg(&gs->pmin); s.min = __ldcg(&gs->min);
                    s.nbytes = __ldcg(&gs->nbytes); s.out_off = __ldcg(&gs->out_off);
                    s.do_bound = __ldcg(&gs->do_bound); s.bits = __ldcg(&gs->bits);
                    s_st = s;
                }
                __syncthreads();
                const BlockStat st = s_st;
                const int bits = st.bits;
                if (bits >= 1 && bits <= 32) {   // 0: nothing to write (go/bit/bit.go:162); > 32: k_pack's (wide_list)
                    if (st.out_off + st.nbytes > A.chain_cap) {   // never write past the caller's buffer
                        if (threadIdx.x == 0) atomicExch(A.err, 2);
                    } else if (d.kind == KIND_F32) {
                        pack_tile_f32(d, st, tib, A.out + (long long)d.chain * A.chain_stride, sv);
                    } else {
                        pack_tile_i64(d, st, tib, A.out + (long long)d.chain * A.chain_stride, sv);
                    }
                }
                __syncthreads();
            }
        }
        __syncthreads();   // s_ticket[it ^ 1] is visible, s_ticket[it] may be overwritten in the next iteration
    }
}

size_t group_fused_ws_bytes(int64_t nblocks) { return (size_t)nblocks * 24 + 256; }

// Fused encode of a batch of UNIFORM contiguous blocks (sh.uniform_n > 0; every float block periodic with
// 1 <= pixels < 2^31).  descs and stats (k_init) are ready on the stream.  ws: group_fused_ws_bytes(nblocks).
// Follow-up: the listed wide blocks through k_pack.
static cudaError_t launch_group_fused(Launcher &L, const BlockDesc *descs, BlockStat *stats, const BatchShape &sh, int *flags,
                               int64_t *mins, int64_t *bits, int64_t *offsets, int64_t *out_len, uint8_t *out,
                               int64_t chain_stride, int64_t chain_cap, void *ws) {
    if (sh.nblocks == 0 || sh.total_tiles == 0) return cudaSuccess;
    static DevCfg cfgs[MNW_MAX_DEVICES];
    DevCfg &dc = dev_cfg(cfgs);
    std::call_once(dc.once, [&] {
        int dev = 0, sms = 148, per = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        dc.err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_group_fused, GF_THREADS, 0);
        if (dc.err != cudaSuccess) return;
        if (per < 1) { dc.err = cudaErrorLaunchOutOfResources; return; }
        dc.a = per * sms;
        dc.b = getenv("MNW_GROUP_WAVE") ? atoi(getenv("MNW_GROUP_WAVE")) : 0;   // tuning knob: tiles per wave
        if (getenv("MNW_DEBUG")) fprintf(stderr, "k_group_fused: %d co-resident CTAs\n", dc.a);
    });
    if (dc.err != cudaSuccess) return dc.err;
    GroupFusedArgs A = {};
    A.descs = descs; A.stats = stats; A.sh = sh;
    A.mins = mins; A.bits = bits; A.offsets = offsets; A.out_len = out_len; A.out = out;
    A.chain_stride = chain_stride; A.chain_cap = chain_cap;
    A.pub = (unsigned long long *)ws;
    A.wide_list = (int64_t *)(A.pub + sh.nblocks);
    A.done = (unsigned *)(A.wide_list + sh.nblocks);
    A.err = flags + 1; A.wide_count = flags + 3; A.ticket = (unsigned *)(flags + 4);
    // a wave: enough tiles to keep every resident CTA busy for a couple of tiles; its two passes are then one wave apart
    // and it is a whole number of blocks: every statistics tile of a block then has a smaller ticket than any of the
    // block's pack tiles (a pack tile only ever waits for smaller tickets)
    const long long tpb = (sh.uniform_n + PACK_TILE - 1) / PACK_TILE;
    A.wave = dc.b > 0 ? dc.b : 2LL * dc.a;
    A.wave = (A.wave + tpb - 1) / tpb * tpb;
    if (A.wave > sh.total_tiles) A.wave = sh.total_tiles;
    cudaError_t e = cudaMemsetAsync(ws, 0, group_fused_ws_bytes(sh.nblocks), L.stream);
    if (e != cudaSuccess) return e;
    const long long nwaves = (sh.total_tiles + A.wave - 1) / A.wave, ntickets = 2 * (nwaves + 1) * A.wave;
    const unsigned grid = (unsigned)(ntickets < dc.a ? ntickets : dc.a);
    L.begin("k_group_fused");
    k_group_fused<<<grid, GF_THREADS, 0, L.stream>>>(A);
    L.end();
    L.count++;
    return cudaGetLastError();
}

cudaError_t launch_group_encode(Launcher &L, const BlockDesc *descs, BlockStat *stats, const BatchShape &sh, int *flags,
                                int64_t *mins, int64_t *bits, int64_t *offsets, int64_t *out_len, uint8_t *out,
                                int64_t chain_stride, int64_t chain_cap, void *ws) {
    launch_init_stats(L, descs, stats, sh.nblocks);
    cudaError_t e = launch_group_fused(L, descs, stats, sh, flags, mins, bits, offsets, out_len, out, chain_stride, chain_cap, ws);
    if (e != cudaSuccess) return e;
    // blocks wider than 32 bits, with the fused kernel's (min, bits, offset): the 64-bit capable packer
    launch_pack_list(L, descs, stats, sh, (const int64_t *)((unsigned long long *)ws + sh.nblocks), flags + 3, out, chain_stride,
                     chain_cap, flags + 1);
    return cudaGetLastError();
}

}  // namespace mnw
