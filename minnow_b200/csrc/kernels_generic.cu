// kernels_generic.cu -- the generic (any block size, 64-bit capable, gather /
// sub-cell capable) two-pass device path of libminnow_b200:
//
//   k_build_*   build BlockDesc records on the device
//   k_init      per-block stat init + value of element 0
//   k_stats     one read of the input: min/max and the order-independent
//               periodic-arc statistics             (go/group.go:244-248,384-409)
//   k_finalize  per block: periodicMin closed form, min, bits, nbytes
//   k_slow      exact sequential periodicMin for blocks holding out-of-range
//               pixel indices (rare), then min/max of bound()
//   k_scan      per chain exclusive scan of nbytes   (go/block_index.go:16-35)
//   k_pack      second read: bound, subtract min, LSB-first bit packing to a
//               BYTE-aligned offset                  (go/bit/bit.go:84-134)
//   k_decode    unpack + min + bound + dequantise    (go/bit/bit.go:29-82,
//                                                     go/group.go:257-263,299-310)
//
// The fused single-read cluster kernels live in kernels_fused.cu; this file is
// the fallback for blocks they do not cover and the int64 path.
#include "group_detail.cuh"

namespace mnw {

// ---------------------------------------------------------------------------
// descriptor builders
// ---------------------------------------------------------------------------
// With `idx` given the block is a GATHER: element i of block b is src[idx[first + i]]
// (BoundaryWriter.Column, go/minh/boundary.go:184-225).
__global__ void k_build_contig(BlockDesc *descs, int64_t nb, int32_t kind, const void *src, int64_t n,
                               const int64_t *starts, const int64_t *tile0, const int64_t *chunk0,
                               FloatParams fp, int64_t blocks_per_chain, const int64_t *idx, BlockStat *stats_init,
                               unsigned long long *ws_zero) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= nb) return;
    BlockDesc d = {};
    int64_t first = starts ? starts[b] : b * n;
    int64_t cnt = starts ? starts[b + 1] - starts[b] : n;
    if (idx) {
        d.src = src;
        d.idx = idx + first;
        d.access = ACC_GATHER;
    } else {
        d.src = kind == KIND_I64 ? (const void *)((const long long *)src + first)
                                 : (const void *)((const float *)src + first);
        d.access = ACC_CONTIG;
    }
    d.n = cnt;
    d.kind = kind;
    d.flags = fp.flags;
    d.low = fp.low; d.high = fp.high; d.dx = fp.dx; d.hi_clamp = fp.hi_clamp; d.pixels = fp.pixels;
    d.chain = (int32_t)(b / blocks_per_chain);
    int64_t tpb = (n + PACK_TILE - 1) / PACK_TILE, cpb = (n + STATS_CHUNK - 1) / STATS_CHUNK;
    d.tile0 = tile0 ? tile0[b] : b * tpb;
    d.chunk0 = chunk0 ? chunk0[b] : b * cpb;
    descs[b] = d;
    if (stats_init) {   // what k_init would do in a launch of its own (the fused group encode)
        BlockStat s = {};
        s.wmin = ~0ULL; s.wmax = 0ULL;
        s.qmin = LLONG_MAX; s.qmax = LLONG_MIN;
        s.q0 = d.n > 0 ? block_value(d, 0) : 0;
        stats_init[b] = s;
    }
    if (ws_zero) {      // k_group_fused's per-block words: look-back word, wide list entry, counter
        ws_zero[b] = 0ULL; ws_zero[nb + b] = 0ULL;
        ((unsigned *)(ws_zero + 2 * nb))[b] = 0u;
    }
}

// minp.Writer.Vectors block order: per file f, axis k, sub-cell sc (go/minp/minp.go:112-118)
__global__ void k_build_vec3(BlockDesc *descs, int64_t nfiles, const float *aos, int32_t nfile,
                             int32_t subcells, const FloatParams *tab, int tab_per_file) {
    int64_t sc3 = (int64_t)subcells * subcells * subcells;
    int64_t nb = nfiles * 3 * sc3;
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= nb) return;
    int64_t f = b / (3 * sc3);
    int32_t k = (int32_t)((b / sc3) % 3);
    int64_t sc = b % sc3;
    int32_t nsub = nfile / subcells;
    const FloatParams fp = tab[(tab_per_file ? 3 * f : 0) + k];
    BlockDesc d = {};
    d.src = aos + 3 * f * (int64_t)nfile * nfile * nfile;
    d.n = (int64_t)nsub * nsub * nsub;
    d.kind = KIND_F32;
    d.access = ACC_SUBCELL;
    d.flags = fp.flags;
    d.low = fp.low; d.high = fp.high; d.dx = fp.dx; d.hi_clamp = fp.hi_clamp; d.pixels = fp.pixels;
    d.chain = (int32_t)(b / sc3);
    d.nfile = nfile; d.nsub = nsub;
    d.ix0 = nsub * (int32_t)(sc % subcells);
    d.iy0 = nsub * (int32_t)((sc / subcells) % subcells);
    d.iz0 = nsub * (int32_t)(sc / ((int64_t)subcells * subcells));
    d.axis = k;
    int64_t tpb = (d.n + PACK_TILE - 1) / PACK_TILE, cpb = (d.n + STATS_CHUNK - 1) / STATS_CHUNK;
    d.tile0 = b * tpb;
    d.chunk0 = b * cpb;
    descs[b] = d;
}

// ---------------------------------------------------------------------------
// stats
// ---------------------------------------------------------------------------
// `run_if` (may be null): the kernels of the generic encode do nothing unless *run_if != 0.
// The fused path enqueues them behind itself as the exact redo for inputs it refuses.
__global__ void k_init(const BlockDesc *descs, BlockStat *stats, int64_t nb, const int *run_if) {
    if (run_if && *run_if == 0) return;
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= nb) return;
    BlockStat s = {};
    s.wmin = ~0ULL; s.wmax = 0ULL;
    s.qmin = LLONG_MAX; s.qmax = LLONG_MIN;
    s.q0 = descs[b].n > 0 ? block_value(descs[b], 0) : 0;
    stats[b] = s;
}

__global__ void __launch_bounds__(STATS_THREADS)
k_stats(const BlockDesc *__restrict__ descs, BlockStat *stats, BatchShape sh, const int *run_if) {
    __shared__ long long s_ll[2][STATS_THREADS / 32];
    __shared__ unsigned long long s_ull[2][STATS_THREADS / 32];
    __shared__ unsigned int s_oob;
    if (run_if && *run_if == 0) return;
    for (int64_t chunk = blockIdx.x; chunk < sh.total_chunks; chunk += gridDim.x) {
    int64_t cpb = sh.uniform_n > 0 ? (sh.uniform_n + STATS_CHUNK - 1) / STATS_CHUNK : 0;
    int64_t b = find_block(descs, sh, chunk, cpb, false);
    const BlockDesc d = descs[b];
    int64_t first = (chunk - d.chunk0) * STATS_CHUNK;
    int64_t end = first + STATS_CHUNK < d.n ? first + STATS_CHUNK : d.n;
    const bool periodic = d.kind == KIND_F32 && (d.flags & F_PERIODIC);
    const long long P = d.pixels;
    const long long q0 = stats[b].q0;
    const bool q0_ok = P > 0 && (unsigned long long)q0 < (unsigned long long)P;
    const unsigned long long C = (periodic && q0_ok) ? arc_rotation(q0, P) : 0ULL;
    if (threadIdx.x == 0) s_oob = 0;

    unsigned long long wmin = ~0ULL, wmax = 0ULL;
    long long qmin = LLONG_MAX, qmax = LLONG_MIN;
    unsigned int oob = 0;
    for (int64_t i = first + threadIdx.x; i < end; i += STATS_THREADS) {
        long long q = block_value(d, i);
        // A pixel index equal to `pixels` (x within half an ulp of `high`) behaves
        // exactly like index 0 in periodicDistance, periodicMin and bound
        // (go/group.go:374-420) as long as x[0] itself is in range: fold it.
        if (periodic && q0_ok && q == P) q = 0;
        qmin = q < qmin ? q : qmin;
        qmax = q > qmax ? q : qmax;
        if (periodic) {
            if (!q0_ok || (unsigned long long)q >= (unsigned long long)P) {
                oob = 1;
            } else {
                unsigned long long w = (unsigned long long)q + C;
                if (w >= (unsigned long long)P) w -= (unsigned long long)P;
                wmin = w < wmin ? w : wmin;
                wmax = w > wmax ? w : wmax;
            }
        }
    }
    qmin = warp_min_ll(qmin); qmax = warp_max_ll(qmax);
    wmin = warp_min_ull(wmin); wmax = warp_max_ull(wmax);
    oob = __any_sync(0xffffffffu, oob);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) {
        s_ll[0][w] = qmin; s_ll[1][w] = qmax; s_ull[0][w] = wmin; s_ull[1][w] = wmax;
        if (oob) atomicOr(&s_oob, 1u);
    }
    __syncthreads();
    if (w == 0) {
        constexpr int NW = STATS_THREADS / 32;
        qmin = l < NW ? s_ll[0][l] : LLONG_MAX;
        qmax = l < NW ? s_ll[1][l] : LLONG_MIN;
        wmin = l < NW ? s_ull[0][l] : ~0ULL;
        wmax = l < NW ? s_ull[1][l] : 0ULL;
        qmin = warp_min_ll(qmin); qmax = warp_max_ll(qmax);
        wmin = warp_min_ull(wmin); wmax = warp_max_ull(wmax);
        if (l == 0) {
            BlockStat *s = &stats[b];
            atomicMin(&s->qmin, qmin);
            atomicMax(&s->qmax, qmax);
            if (periodic) {
                atomicMin(&s->wmin, wmin);
                atomicMax(&s->wmax, wmax);
                if (s_oob || !q0_ok) atomicOr(&s->oob, 1u);
            }
        }
    }
    __syncthreads();   // shared scratch is reused by the next chunk
    }
}

__global__ void k_finalize(const BlockDesc *descs, BlockStat *stats, int64_t nb, int64_t *slow_list,
                           int *slow_count, int *err, const int *run_if) {
    if (run_if && *run_if == 0) return;
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= nb) return;
    BlockStat s = stats[b];
    if (finalize_block(descs[b], s, err)) slow_list[atomicAdd(slow_count, 1)] = b;
    stats[b] = s;
}

// One CTA per slow block (slow_block, group_detail.cuh).
__global__ void __launch_bounds__(256)
k_slow(const BlockDesc *descs, BlockStat *stats, const int64_t *slow_list, const int *slow_count, int *err,
       const int *run_if) {
    __shared__ long long s_pmin;
    __shared__ long long s_red[2][8];
    if (run_if && *run_if == 0) return;
    const int nslow = *slow_count;
    for (int si = blockIdx.x; si < nslow; si += gridDim.x) {
        const int64_t b = slow_list[si];
        const BlockDesc d = descs[b];
        slow_block(d, &stats[b], err, &s_pmin, s_red);
    }
}

// ---------------------------------------------------------------------------
// bounds() of minp.Writer.Vectors for non-periodic fields, go/minp/minp.go:291-300.
// float min/max through order-preserving uint32 keys and 32-bit atomics.
// ---------------------------------------------------------------------------
constexpr int LIMITS_CHUNK = 32768;  // particles per CTA
__device__ __forceinline__ uint32_t float_key(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void k_limits_init(uint32_t *keys, int64_t nfiles) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nfiles * 6) keys[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
}
__global__ void __launch_bounds__(256) k_vec3_limits(const float *__restrict__ aos, int64_t np, uint32_t *keys) {
    const int64_t f = blockIdx.y;
    const float *base = aos + 3 * f * np;
    int64_t p0 = (int64_t)blockIdx.x * LIMITS_CHUNK;
    int64_t p1 = p0 + LIMITS_CHUNK < np ? p0 + LIMITS_CHUNK : np;
    uint32_t mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    // thread t reads floats t, t+256, ...; component = index % 3
    for (int64_t i = 3 * p0 + threadIdx.x; i < 3 * p1; i += 256 * 3) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            int64_t ii = i + 256 * j;
            if (ii < 3 * p1) {
                float v = base[ii];
                if (v == v) {  // NaN never wins a comparison in bounds()
                    uint32_t key = float_key(v);
                    int c = (int)(ii % 3);
                    if (c == 0) { mn[0] = min(mn[0], key); mx[0] = max(mx[0], key); }
                    else if (c == 1) { mn[1] = min(mn[1], key); mx[1] = max(mx[1], key); }
                    else { mn[2] = min(mn[2], key); mx[2] = max(mx[2], key); }
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        mn[c] = __reduce_min_sync(0xffffffffu, mn[c]);
        mx[c] = __reduce_max_sync(0xffffffffu, mx[c]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            atomicMin(&keys[f * 6 + c], mn[c]);
            atomicMax(&keys[f * 6 + 3 + c], mx[c]);
        }
    }
}

// The same with 128-bit loads (np a multiple of 4, 16-byte aligned cubes): 192 threads, so that a
// thread's float4 always starts on the same axis and its 4 floats have fixed axes (a0, a0+1, a0+2, a0).
constexpr int LIMITS4_THREADS = 192, LIMITS4_PER_THREAD = 16;
__global__ void __launch_bounds__(LIMITS4_THREADS) k_vec3_limits4(const float *__restrict__ aos, int64_t np, uint32_t *keys) {
    const int64_t f = blockIdx.y;
    const float4 *base = (const float4 *)(aos + 3 * f * np);
    const int64_t n4 = 3 * np / 4;
    const int64_t c0 = (int64_t)blockIdx.x * (LIMITS4_THREADS * LIMITS4_PER_THREAD);
    const int a0 = threadIdx.x % 3;   // c0 and the thread stride are multiples of 3
    const float inf = __int_as_float(0x7f800000);
    float mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};   // relative axis j = actual (a0 + j) % 3
#pragma unroll 4
    for (int i = 0; i < LIMITS4_PER_THREAD; i++) {
        const int64_t g = c0 + threadIdx.x + (int64_t)LIMITS4_THREADS * i;
        if (g < n4) {
            const float4 v = __ldcs(base + g);
            // fminf / fmaxf drop a NaN operand: NaN never wins a comparison in bounds() either
            mn[0] = fminf(mn[0], fminf(v.x, v.w)); mx[0] = fmaxf(mx[0], fmaxf(v.x, v.w));
            mn[1] = fminf(mn[1], v.y); mx[1] = fmaxf(mx[1], v.y);
            mn[2] = fminf(mn[2], v.z); mx[2] = fmaxf(mx[2], v.z);
        }
    }
    __shared__ uint32_t s_k[6];
    if (threadIdx.x < 6) s_k[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int ax = (a0 + j) % 3;
        if (mn[j] <= mx[j]) {   // at least one non-NaN value seen
            atomicMin(&s_k[ax], float_key(mn[j]));
            atomicMax(&s_k[3 + ax], float_key(mx[j]));
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(&keys[f * 6 + threadIdx.x], s_k[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(&keys[f * 6 + threadIdx.x], s_k[threadIdx.x]);
}

// The three FloatGroups of every file of a NON-periodic field, derived on the device from the limits keys -- what
// minp.Writer.Vectors does on the host between bounds() and FloatGroup (go/minp/minp.go:92-95, go/writer.go:72-75):
// lo = min, hi = Nextafter32(max, 2 * max), pixels = int64(ceil(float64((hi - lo) / dx))) in float32, and the derived
// constants of FloatParams exactly as api.cu to_params computes them on the host.  flags[0] (skip) and flags[1] (abort)
// are raised when some group is outside what the fused minp kernels cover (they then return at once and the generic
// kernels, gated on the abort flag, encode the batch).
__global__ void k_vec3_params(const uint32_t *keys, int64_t nfiles, float dx_user, FloatParams *tab, FloatDescPod *desc_out,
                              int *skip, int *abort_flag, int need_pipe) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= 3 * nfiles) return;
    const int64_t f = i / 3;
    const int k = (int)(i % 3);
    auto key_to_float = [](uint32_t kk) { return __uint_as_float((kk & 0x80000000u) ? (kk & 0x7fffffffu) : ~kk); };
    const float lo = key_to_float(keys[6 * f + k]), mx = key_to_float(keys[6 * f + 3 + k]);
    const float hi = nextafterf(mx, __fmul_rn(2.0f, mx));                       // go/minp/minp.go:94
    const float span = __fsub_rn(hi, lo);
    const double c = ceil((double)__fdiv_rn(span, dx_user));                    // go/writer.go:73
    const long long pixels = (c >= -9223372036854775808.0 && c < 9223372036854775808.0) ? (long long)c : LLONG_MIN;
    FloatParams p = {};
    p.low = lo; p.high = hi; p.pixels = pixels;
    p.dx = __fdiv_rn(span, __ll2float_rn(pixels));                               // go/group.go:316
    p.hi_clamp = nextafterf(hi, -INFINITY);
    p.flags = F_PERIODIC;                                                        // go/writer.go:74
    p.rcp = __fdiv_rn(1.0f, p.dx);
    const bool normal_dx = fabsf(p.dx) >= 1.17549435e-38f && fabsf(p.dx) <= 3.40282347e38f;
    const bool normal_rcp = fabsf(p.rcp) >= 1.17549435e-38f && fabsf(p.rcp) <= 3.40282347e38f;
    if (normal_dx && normal_rcp && p.dx > 0x1p-60f && p.dx < 0x1p60f && pixels >= 2) p.flags |= F_FASTDIV;
    tab[i] = p;
    FloatDescPod d = {};
    d.low = lo; d.high = hi; d.pixels = pixels; d.periodic = 1;
    if (desc_out) desc_out[i] = d;
    if (pixels < 1 || pixels >= (1LL << 30) || (need_pipe && pixels > (1LL << 22))) { atomicExch(skip, 1); atomicExch(abort_flag, 1); }
}

void launch_vec3_params(Launcher &L, const uint32_t *keys, int64_t nfiles, float dx, FloatParams *tab, void *desc_out, int *skip,
                        int *abort_flag, int need_pipe) {
    if (nfiles == 0) return;
    k_vec3_params<<<(unsigned)((3 * nfiles + 127) / 128), 128, 0, L.stream>>>(keys, nfiles, dx, tab, (FloatDescPod *)desc_out, skip, abort_flag, need_pipe);
    L.count++;
}

// FloatParams from group descriptors that live on the device (the decode side of the same fields).
__global__ void k_params_from_desc(const FloatDescPod *desc, int64_t n, FloatParams *tab) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const FloatDescPod d = desc[i];
    FloatParams p = {};
    p.low = d.low; p.high = d.high; p.pixels = d.pixels;
    p.dx = __fdiv_rn(__fsub_rn(d.high, d.low), __ll2float_rn(d.pixels));
    p.hi_clamp = nextafterf(d.high, -INFINITY);
    p.flags = (d.periodic ? F_PERIODIC : 0) | (d.log10 ? F_LOG10 : 0) | (d.clamp ? F_CLAMP : 0);
    p.rcp = __fdiv_rn(1.0f, p.dx);
    tab[i] = p;
}
void launch_params_from_desc(Launcher &L, const void *desc, int64_t n, FloatParams *tab) {
    if (n == 0) return;
    k_params_from_desc<<<(unsigned)((n + 127) / 128), 128, 0, L.stream>>>((const FloatDescPod *)desc, n, tab);
    L.count++;
}

// ---------------------------------------------------------------------------
// scan: one CTA per chain (= minnow group); exclusive prefix of nbytes
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_scan(BlockStat *stats, BatchShape sh, int64_t *mins, int64_t *bits, int64_t *offsets, int64_t *out_len,
       const int *run_if) {
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    if (run_if && *run_if == 0) return;
    const int64_t chain = blockIdx.x;
    const int64_t b0 = chain * sh.blocks_per_chain;
    const int64_t b1 = b0 + sh.blocks_per_chain < sh.nblocks ? b0 + sh.blocks_per_chain : sh.nblocks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = b0; base < b1; base += 1024) {
        int64_t b = base + threadIdx.x;
        long long v = b < b1 ? stats[b].nbytes : 0;
        long long incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            long long w = s_warp[lane];
            long long wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            s_warp[lane] = wi - w;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        long long carry = s_carry;
        long long excl = carry + s_warp[warp] + incl - v;
        if (b < b1) {
            stats[b].out_off = excl;
            if (offsets) offsets[b] = excl;
            if (mins) mins[b] = stats[b].min;
            if (bits) bits[b] = stats[b].bits;
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && out_len) out_len[chain] = s_carry;
}

// ---------------------------------------------------------------------------
// pack
// ---------------------------------------------------------------------------
// Tiles are walked grid-stride.  With `list` given only the listed blocks are packed
// (*list_count of them; the fused path's blocks wider than 16 bits), else all of them.
__global__ void __launch_bounds__(PACK_THREADS)
k_pack(const BlockDesc *__restrict__ descs, const BlockStat *__restrict__ stats, BatchShape sh,
       uint8_t *out, int64_t chain_stride, int64_t chain_cap, int *err, const int *run_if,
       const int64_t *list, const int *list_count, int only_above = 0) {
    __shared__ uint32_t s_out[PACK_THREADS * 64 + 4];
    if (run_if && *run_if == 0) return;
    const int64_t tpb = sh.uniform_n > 0 ? (sh.uniform_n + PACK_TILE - 1) / PACK_TILE : 0;
    const int64_t total = list ? (int64_t)*list_count * tpb : sh.total_tiles;
    for (int64_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int64_t b, tile_in_block;
        if (list) {
            b = list[tile / tpb];
            tile_in_block = tile % tpb;
        } else {
            b = find_block(descs, sh, tile, tpb, true);
            tile_in_block = tile - descs[b].tile0;
        }
        const BlockStat st = stats[b];
        const int bits = st.bits;
        if (bits == 0) continue;  // ArrayBuffer.Write returns at once, go/bit/bit.go:162
        if (bits <= only_above) continue;   // narrower blocks were packed by the vectorised kernel
        if (st.out_off + st.nbytes > chain_cap) {  // never write past the caller's buffer
            if (threadIdx.x == 0) atomicExch(err, 2);
            continue;
        }
        const BlockDesc d = descs[b];
        pack_tile_generic(d, st, tile_in_block, out + (int64_t)d.chain * chain_stride, s_out);
    }
}

// ---------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long extract_bits(const uint8_t *stream, int64_t stream_len,
                                                           int64_t byte_off, int64_t i, int bits) {
    const uintptr_t S = (uintptr_t)stream;
    const int a = (int)(S & 3);
    const uint32_t *base = (const uint32_t *)(S - a);
    const int64_t limit = (a + stream_len + 3) >> 2;
    unsigned long long bitpos = ((unsigned long long)byte_off + a) * 8ULL + (unsigned long long)i * bits;
    int64_t wi = (int64_t)(bitpos >> 5);
    int sh = (int)(bitpos & 31);
    uint32_t w0 = wi < limit ? __ldg(base + wi) : 0u;
    uint32_t w1 = (sh + bits > 32 && wi + 1 < limit) ? __ldg(base + wi + 1) : 0u;
    unsigned long long v = (((unsigned long long)w1 << 32) | w0) >> sh;
    if (sh + bits > 64) {
        uint32_t w2 = wi + 2 < limit ? __ldg(base + wi + 2) : 0u;
        v |= (unsigned long long)w2 << (64 - sh);
    }
    if (bits < 64) v &= (1ULL << bits) - 1ULL;
    return v;
}

struct DecodeArgs {
    int mode;                 // 0: int group, 1: float group, 2: vec3 sub-cells
    const uint8_t *data;
    int64_t stream_len;       // bytes per stream (group: data_len; vec3: axis stride)
    const int64_t *offsets, *mins, *bits, *sel, *jitter_ids;
    int64_t n, nsel;
    const FloatParams *tab;
    int tab_per_file;
    float wrap_L;
    int jmode;
    unsigned long long seed, block_id0;
    const double *u;
    int32_t nfile, nsub, subcells;
    int64_t sc3;
    void *out;
    void *const *outs;        // contiguous decoders: output of selected block j (else out + j * n elements)
};

__global__ void __launch_bounds__(DEC_THREADS) k_decode(DecodeArgs A) {
    const int64_t cpb = (A.n + DEC_CHUNK - 1) / DEC_CHUNK;
    const int64_t j = blockIdx.x / cpb;          // which selected block
    const int64_t c = blockIdx.x - j * cpb;
    const int64_t b = A.sel ? A.sel[j] : j;      // block id within the group / batch
    const long long mn = A.mins[b];
    const int bits = (int)A.bits[b];
    const int64_t off = A.offsets[b];
    const uint8_t *stream = A.data;
    int k = 0;
    int64_t f = 0, sc = 0;
    if (A.mode == 2) {
        f = b / (3 * A.sc3);
        k = (int)((b / A.sc3) % 3);
        sc = b % A.sc3;
        stream += (b / A.sc3) * A.stream_len;    // stream (3f + k)
    }
    long long P = 0;
    float low = 0.f, dx = 0.f;
    bool periodic = false, islog = false;
    if (A.mode != 0) {
        const FloatParams fp = A.tab[(A.tab_per_file ? 3 * f : 0) + k];
        P = fp.pixels; low = fp.low; dx = fp.dx; periodic = fp.flags & F_PERIODIC; islog = fp.flags & F_LOG10;
    }
    const int64_t end = (c + 1) * DEC_CHUNK < A.n ? (c + 1) * DEC_CHUNK : A.n;
    for (int64_t i = c * DEC_CHUNK + threadIdx.x; i < end; i += DEC_THREADS) {
        unsigned long long v = bits ? extract_bits(stream, A.stream_len, off, i, bits) : 0ULL;
        long long q = (long long)((unsigned long long)mn + v);  // go/group.go:262
        if (A.mode == 0) {
            ((long long *)A.out)[j * A.n + i] = q;
            continue;
        }
        if (periodic) q = bound1(q, 0, P);              // go/group.go:303
        double u = 0.5;
        if (A.jmode == 1) u = (double)(jitter_hash32(A.seed, A.block_id0 + (unsigned long long)(A.jitter_ids ? A.jitter_ids[j] : b), (unsigned long long)i) >> 8) * 0x1p-24;
        else if (A.jmode == 2) u = A.u[j * A.n + i];
        float t = __double2float_rn(__dadd_rn(__ll2double_rn(q), u));  // go/group.go:308
        float o = __fadd_rn(__fmul_rn(dx, t), low);
        if (A.mode == 1) {
            if (islog) o = go_pow10_f32(o);   // minh Log column, go/minh/minh.go:315-319
            ((float *)A.out)[j * A.n + i] = o;
        } else {
            if (A.wrap_L > 0.0f) {                               // go/minp/minp.go:195-203
                if (o < 0.0f) o = __fadd_rn(o, A.wrap_L);
                else if (o >= A.wrap_L) o = __fsub_rn(o, A.wrap_L);
            }
            uint32_t ii = (uint32_t)i, ns = (uint32_t)A.nsub;   // setSubCell, go/minp/minp.go:270-288
            uint32_t jx = ii % ns, tt = ii / ns;
            uint32_t jy = tt % ns, jz = tt / ns;
            int32_t ix0 = A.nsub * (int32_t)(sc % A.subcells);
            int32_t iy0 = A.nsub * (int32_t)((sc / A.subcells) % A.subcells);
            int32_t iz0 = A.nsub * (int32_t)(sc / ((int64_t)A.subcells * A.subcells));
            int64_t idx = (int64_t)(jx + ix0) + (int64_t)(jy + iy0) * A.nfile + (int64_t)(jz + iz0) * A.nfile * A.nfile;
            float *cube = (float *)A.out + 3 * f * (int64_t)A.nfile * A.nfile * A.nfile;
            cube[3 * idx + k] = o;
        }
    }
}

// max of uint64 (ArrayBuffer.Bits, go/bit/bit.go:151-159)
__global__ void __launch_bounds__(256) k_umax(const unsigned long long *x, int64_t n, unsigned long long *out) {
    unsigned long long m = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = x[i] > m ? x[i] : m;
    m = warp_max_ull(m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// raw pack of caller-supplied uint64 values with caller-supplied bits
// (bit.BufferedArray): presets the stat so that k_pack does the work.
__global__ void k_preset_raw(BlockDesc *descs, BlockStat *stats, const void *src, int64_t n, int bits) {
    BlockDesc d = {};
    d.src = src; d.n = n; d.kind = KIND_I64; d.access = ACC_CONTIG;
    d.tile0 = 0; d.chunk0 = 0;
    descs[0] = d;
    BlockStat s = {};
    s.min = 0; s.bits = bits; s.nbytes = array_bytes(bits, n); s.out_off = 0; s.do_bound = 0;
    stats[0] = s;
}

// blockIndex.addBlock/blockOffset over a plain size array (go/block_index.go:16-35)
__global__ void __launch_bounds__(1024)
k_scan_sizes(const int64_t *sizes, int64_t n, int64_t base, int64_t *offsets, int64_t *total) {
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = base;
    __syncthreads();
    for (int64_t b0 = 0; b0 < n; b0 += 1024) {
        int64_t b = b0 + threadIdx.x;
        long long v = b < n ? sizes[b] : 0, incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            long long w = s_warp[lane], wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        long long excl = s_carry + s_warp[warp] + incl - v;
        if (b < n) offsets[b] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry - base;
}

constexpr int FSTAT_THREADS = 256;
__global__ void __launch_bounds__(FSTAT_THREADS)
k_stats_f32c(const BlockDesc *__restrict__ descs, BlockStat *stats, BatchShape sh, const int *run_if) {
    __shared__ unsigned s_r[FSTAT_THREADS / 32][5];
    if (run_if && *run_if == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t chunk = blockIdx.x; chunk < sh.total_chunks; chunk += gridDim.x) {
        const int64_t cpb = sh.uniform_n > 0 ? (sh.uniform_n + STATS_CHUNK - 1) / STATS_CHUNK : 0;
        const int64_t b = find_block(descs, sh, chunk, cpb, false);
        const BlockDesc d = descs[b];
        if (d.kind != KIND_F32) continue;   // a column batch mixes kinds: int64 blocks are k_stats_i64c's
        const int64_t first = (chunk - d.chunk0) * STATS_CHUNK;
        const int count = (int)((first + STATS_CHUNK < d.n ? first + STATS_CHUNK : d.n) - first);
        const QuantP qp = quant_params(d);
        const long long q0 = stats[b].q0;
        const bool q0_ok = (unsigned long long)q0 < (unsigned long long)qp.P;
        const unsigned C = q0_ok ? (unsigned)arc_rotation(q0, qp.P) : 0u;
        unsigned wmin = ~0u, wmax = 0u, qmin = ~0u, qmax = 0u;
        bool oob = !q0_ok;
        const float *p = (const float *)d.src + first;
        const int a = (int)(((uintptr_t)p & 15) >> 2);          // elements of the first 16 bytes that precede the chunk
        const float4 *base4 = (const float4 *)(p - a);
        const int nvec = (a + count + 3) >> 2;
        // checked path for vectors [v0, v1) of this thread (any element may lie outside the chunk)
        auto checked = [&](int v0, int v1) {
            for (int iv = v0 + threadIdx.x; iv < v1; iv += FSTAT_THREADS) {
                const float4 v4 = __ldcs(base4 + iv);
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
            }
        };
        if (quant_bits_ok(qp) && q0_ok) {
            // whole vectors [v_lo, v_hi): unchecked quantiser, ONE range test per thread on the min/max of the raw
            // bits; a thread that fails it (rare) takes its vectors again through the checked path
            const int v_lo = a ? 1 : 0, v_hi = (a + count) >> 2;
            const bool clamp = qp.flags & F_CLAMP;
            const unsigned Cm = C - FQ_MAGIC, nP = 0u - qp.P;
            unsigned bmin = ~0u, bmax = 0u, fwmin = ~0u, fwmax = 0u;
#pragma unroll 2
            for (int iv = v_lo + threadIdx.x; iv < v_hi; iv += FSTAT_THREADS) {
                const float4 v4 = __ldcs(base4 + iv);
                const unsigned b0 = quant_bits(v4.x, qp, clamp), b1 = quant_bits(v4.y, qp, clamp);
                const unsigned b2 = quant_bits(v4.z, qp, clamp), b3 = quant_bits(v4.w, qp, clamp);
                const unsigned t0 = b0 + Cm, t1 = b1 + Cm, t2 = b2 + Cm, t3 = b3 + Cm;
                const unsigned w0 = min(t0, t0 + nP), w1 = min(t1, t1 + nP), w2 = min(t2, t2 + nP), w3 = min(t3, t3 + nP);
                bmin = __vimin3_u32(bmin, b0, b1); bmin = __vimin3_u32(bmin, b2, b3);
                bmax = __vimax3_u32(bmax, b0, b1); bmax = __vimax3_u32(bmax, b2, b3);
                fwmin = __vimin3_u32(fwmin, w0, w1); fwmin = __vimin3_u32(fwmin, w2, w3);
                fwmax = __vimax3_u32(fwmax, w0, w1); fwmax = __vimax3_u32(fwmax, w2, w3);
            }
            if (bmin <= bmax) {   // this thread saw whole vectors
                if (bmin >= FQ_MAGIC && bmax < FQ_MAGIC + qp.P) {
                    wmin = fwmin; wmax = fwmax; qmin = bmin - FQ_MAGIC; qmax = bmax - FQ_MAGIC;
                } else {
                    for (int iv = v_lo + threadIdx.x; iv < v_hi; iv += FSTAT_THREADS) {
                        const float4 v4 = __ldcs(base4 + iv);
                        const float x[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            const unsigned q = quant_elem(x[c], qp, oob, nullptr);
                            unsigned w = q + C;
                            w = min(w, w - qp.P);
                            wmin = min(wmin, w); wmax = max(wmax, w); qmin = min(qmin, q); qmax = max(qmax, q);
                        }
                    }
                }
            }
            checked(0, v_lo);        // the partial first / last vector of the chunk
            checked(v_hi, nvec);
        } else {
            checked(0, nvec);
        }
        wmin = __reduce_min_sync(0xffffffffu, wmin); wmax = __reduce_max_sync(0xffffffffu, wmax);
        qmin = __reduce_min_sync(0xffffffffu, qmin); qmax = __reduce_max_sync(0xffffffffu, qmax);
        const unsigned ob = __any_sync(0xffffffffu, oob);
        if (lane == 0) { s_r[warp][0] = wmin; s_r[warp][1] = wmax; s_r[warp][2] = qmin; s_r[warp][3] = qmax; s_r[warp][4] = ob; }
        __syncthreads();
        if (threadIdx.x == 0 && count > 0) {
            for (int wi = 1; wi < FSTAT_THREADS / 32; wi++) {
                wmin = min(wmin, s_r[wi][0]); wmax = max(wmax, s_r[wi][1]);
                qmin = min(qmin, s_r[wi][2]); qmax = max(qmax, s_r[wi][3]);
            }
            unsigned o = 0;
            for (int wi = 0; wi < FSTAT_THREADS / 32; wi++) o |= s_r[wi][4];
            BlockStat *st = &stats[b];
            if (wmin <= wmax) {
                atomicMin(&st->wmin, (unsigned long long)wmin); atomicMax(&st->wmax, (unsigned long long)wmax);
                atomicMin(&st->qmin, (long long)qmin); atomicMax(&st->qmax, (long long)qmax);
            }
            if (o) atomicOr(&st->oob, 1u);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(FPACK_THREADS, 4)
k_pack_f32c(const BlockDesc *__restrict__ descs, const BlockStat *__restrict__ stats, BatchShape sh,
            uint8_t *out, int64_t chain_stride, int64_t chain_cap, int *err, const int *run_if) {
    __shared__ __align__(16) unsigned sv[PACK_TILE];   // the tile's packed-to-be values, swizzled by 16-byte chunk
    if (run_if && *run_if == 0) return;
    const int64_t tpb = sh.uniform_n > 0 ? (sh.uniform_n + PACK_TILE - 1) / PACK_TILE : 0;
    for (int64_t tile = blockIdx.x; tile < sh.total_tiles; tile += gridDim.x) {
        const int64_t b = find_block(descs, sh, tile, tpb, true);
        if (descs[b].kind != KIND_F32) continue;   // int64 blocks of a column batch are k_pack_i64c's
        const BlockStat st = stats[b];
        const int bits = st.bits;
        if (bits == 0 || bits > 32) continue;   // nothing to write / left to k_pack (not reachable for pixels < 2^31)
        if (st.out_off + st.nbytes > chain_cap) {
            if (threadIdx.x == 0) atomicExch(err, 2);
            continue;
        }
        const BlockDesc d = descs[b];
        pack_tile_f32(d, st, tile - d.tile0, out + (int64_t)d.chain * chain_stride, sv);
    }
}

// ---------------------------------------------------------------------------
// Fast path for contiguous int64 blocks (intGroup.writeData, go/group.go:242-255): the same two passes with
// coalesced 128-bit loads.  k_stats_i64c = int64Min + the max of ArrayBuffer.Bits; k_pack_i64c subtracts the
// minimum (the difference of a block of <= 32 bits fits 32 bits) and packs with the warp packer of pack.cuh.
// Blocks wider than 32 bits are left to k_pack.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(FSTAT_THREADS)
k_stats_i64c(const BlockDesc *__restrict__ descs, BlockStat *stats, BatchShape sh, const int *run_if) {
    __shared__ long long s_r[FSTAT_THREADS / 32][2];
    if (run_if && *run_if == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t chunk = blockIdx.x; chunk < sh.total_chunks; chunk += gridDim.x) {
        const int64_t cpb = sh.uniform_n > 0 ? (sh.uniform_n + STATS_CHUNK - 1) / STATS_CHUNK : 0;
        const int64_t b = find_block(descs, sh, chunk, cpb, false);
        const BlockDesc d = descs[b];
        if (d.kind != KIND_I64) continue;
        const int64_t first = (chunk - d.chunk0) * STATS_CHUNK;
        const int count = (int)((first + STATS_CHUNK < d.n ? first + STATS_CHUNK : d.n) - first);
        const long long *p = (const long long *)d.src + first;
        const int a = (int)(((uintptr_t)p & 15) >> 3);          // 1: the chunk starts in the upper half of a 16-byte pair
        const longlong2 *base2 = (const longlong2 *)(p - a);
        const int nvec = (a + count + 1) >> 1;
        long long mn = LLONG_MAX, mx = LLONG_MIN;
#pragma unroll 4
        for (int iv = threadIdx.x; iv < nvec; iv += FSTAT_THREADS) {
            const longlong2 v = __ldcs(base2 + iv);
            const int e0 = 2 * iv - a;
            if (e0 >= 0) { mn = v.x < mn ? v.x : mn; mx = v.x > mx ? v.x : mx; }
            if (e0 + 1 < count) { mn = v.y < mn ? v.y : mn; mx = v.y > mx ? v.y : mx; }
        }
        mn = warp_min_ll(mn); mx = warp_max_ll(mx);
        if (lane == 0) { s_r[warp][0] = mn; s_r[warp][1] = mx; }
        __syncthreads();
        if (threadIdx.x == 0 && count > 0) {
            for (int wi = 1; wi < FSTAT_THREADS / 32; wi++) {
                mn = s_r[wi][0] < mn ? s_r[wi][0] : mn;
                mx = s_r[wi][1] > mx ? s_r[wi][1] : mx;
            }
            atomicMin(&stats[b].qmin, mn);
            atomicMax(&stats[b].qmax, mx);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(FPACK_THREADS, 4)
k_pack_i64c(const BlockDesc *__restrict__ descs, const BlockStat *__restrict__ stats, BatchShape sh,
            uint8_t *out, int64_t chain_stride, int64_t chain_cap, int *err, const int *run_if) {
    __shared__ __align__(16) unsigned sv[PACK_TILE];   // the tile's packed-to-be values, swizzled by 16-byte chunk
    if (run_if && *run_if == 0) return;
    const int64_t tpb = sh.uniform_n > 0 ? (sh.uniform_n + PACK_TILE - 1) / PACK_TILE : 0;
    for (int64_t tile = blockIdx.x; tile < sh.total_tiles; tile += gridDim.x) {
        const int64_t b = find_block(descs, sh, tile, tpb, true);
        if (descs[b].kind != KIND_I64) continue;
        const BlockStat st = stats[b];
        const int bits = st.bits;
        if (bits == 0 || bits > 32) continue;   // nothing to write / k_pack's
        if (st.out_off + st.nbytes > chain_cap) {
            if (threadIdx.x == 0) atomicExch(err, 2);
            continue;
        }
        const BlockDesc d = descs[b];
        pack_tile_i64(d, st, tile - d.tile0, out + (int64_t)d.chain * chain_stride, sv);
    }
}

// Decode of contiguous float32 blocks: one CTA per 4096-element tile of a selected block.  The
// tile's packed bytes are staged in shared memory with 128-bit loads, every thread then extracts
// four consecutive values, dequantises them and stores one float4.
constexpr int FDEC_THREADS = 128;
// LOG: some group of the batch is a minh Log column (10^x after the dequantisation).  The plain instantiation does not carry
// the FP64 registers of go_pow10_f32 (63 -> ~40 registers: 12 CTAs per SM instead of 8).
template <bool LOG>
__global__ void __launch_bounds__(FDEC_THREADS) k_decode_f32c(DecodeArgs A) {
    __shared__ __align__(16) unsigned spk[DEC_CHUNK + 16];
    const int64_t tpb = (A.n + DEC_CHUNK - 1) / DEC_CHUNK;
    const int64_t j = blockIdx.x / tpb;
    const int64_t tile = blockIdx.x - j * tpb;
    const int64_t b = A.sel ? A.sel[j] : j;
    const long long mn = A.mins[b];
    const int bits = (int)A.bits[b];
    const FloatParams fp = A.tab[A.tab_per_file ? b : 0];   // (a batch of columns: one group per block)
    const long long P = fp.pixels;
    const bool periodic = fp.flags & F_PERIODIC, islog = LOG && (fp.flags & F_LOG10);
    const int64_t first = tile * DEC_CHUNK;
    const int count = (int)(first + DEC_CHUNK <= A.n ? DEC_CHUNK : A.n - first);
    const unsigned long long bid = A.block_id0 + (unsigned long long)(A.jitter_ids ? A.jitter_ids[j] : b);
    const unsigned key = jitter_key(A.seed, bid);
    float *outp = (A.outs ? (float *)A.outs[j] : (float *)A.out + j * A.n) + first;
    const unsigned mask = (bits >= 1 && bits <= 32) ? (0xffffffffu >> (32 - bits)) : 0u;
    // 32-bit path: q = mn + v lies in [0, 2*pixels) (periodic) or [0, 2^23), and float32 holds it exactly
    const bool fast = bits <= 32 && mn >= 0 && P > 0 && P < (1LL << 23) &&
                      (periodic ? mn + (long long)mask < 2 * P : mn + (long long)mask < (1LL << 23));
    if (fast) {
        unsigned shift0 = 0;
        if (bits > 0) {
            const uint8_t *src = A.data + A.offsets[b] + ((first * bits) >> 3);
            const int a16 = (int)((uintptr_t)src & 15);
            const uint4 *s16 = (const uint4 *)(src - a16);
            const int nvec = (a16 + ((count * bits + 7) >> 3) + 15) >> 4;
            for (int i = threadIdx.x; i < nvec; i += FDEC_THREADS) ((uint4 *)spk)[i] = __ldg(s16 + i);
            shift0 = 8u * (unsigned)a16;
        }
        __syncthreads();
        const bool vec_ok = (((uintptr_t)outp & 15) == 0);
        const unsigned Pp = periodic ? (unsigned)P : 0u, mn32 = (unsigned)mn;
        // A thread's four values move on by FDEC_THREADS * 4 elements per iteration = a whole number of 32-bit words:
        // the shift within the word is constant, the word index and the hash argument advance by constants.
        unsigned wi[4], sh[4], hx[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const unsigned el0 = (unsigned)(threadIdx.x * 4 + c);
            const unsigned bp0 = shift0 + el0 * (unsigned)bits;
            wi[c] = bp0 >> 5; sh[c] = bp0 & 31u;
            hx[c] = ((unsigned)first + el0) * 0x9E3779B1U + key;
        }
        const unsigned dW = (unsigned)(FDEC_THREADS * 4 / 32) * (unsigned)bits;
        const bool hash = A.jmode == 1;
        for (int e4 = threadIdx.x * 4; e4 < count; e4 += FDEC_THREADS * 4) {
            float o[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const unsigned v = __funnelshift_r(spk[wi[c]], spk[wi[c] + 1], sh[c]) & mask;   // Array.Slice, go/bit/bit.go:29-82
                wi[c] += dW;
                unsigned q = mn32 + v;                                                   // go/group.go:262
                q = min(q, q - Pp);                                                      // bound(q, 0, pixels), :303
                float t;
                if (hash) {
                    unsigned x = hx[c];
                    hx[c] += (unsigned)(FDEC_THREADS * 4) * 0x9E3779B1U;
                    x ^= x >> 16; x *= 0x7feb352dU;                                      // jitter_hash_keyed (device_math.cuh)
                    x ^= x >> 15; x *= 0x846ca68bU;
                    t = __fmaf_rn((float)(x >> 8), 0x1p-24f, (float)q);
                } else {
                    t = __fadd_rn((float)q, 0.5f);
                }
                o[c] = __fadd_rn(__fmul_rn(fp.dx, t), fp.low);                            // :308
                if (islog) o[c] = go_pow10_f32(o[c]);             // go/minh/minh.go:315-319
            }
            if (vec_ok && e4 + 4 <= count) {
                __stcs((float4 *)(outp + e4), make_float4(o[0], o[1], o[2], o[3]));
            } else {
                for (int c = 0; c < 4; c++) if (e4 + c < count) outp[e4 + c] = o[c];
            }
        }
    } else {   // any width / range: 64-bit arithmetic straight from global memory
        for (int el = threadIdx.x; el < count; el += FDEC_THREADS) {
            const int64_t i = first + el;
            const unsigned long long v = bits ? extract_bits(A.data, A.stream_len, A.offsets[b], i, bits) : 0ULL;
            long long q = (long long)((unsigned long long)mn + v);
            if (periodic) q = bound1(q, 0, P);
            double u = 0.5;
            if (A.jmode == 1) u = (double)(jitter_hash_keyed(key, (uint32_t)i) >> 8) * 0x1p-24;
            const float t = __double2float_rn(__dadd_rn(__ll2double_rn(q), u));
            float o = __fadd_rn(__fmul_rn(fp.dx, t), fp.low);
            if (islog) o = go_pow10_f32(o);
            outp[el] = o;
        }
    }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
static inline unsigned grid_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

void launch_build_contig(Launcher &L, BlockDesc *descs, int64_t nb, int32_t kind, const void *src, int64_t n,
                         const int64_t *starts, const int64_t *tile0, const int64_t *chunk0,
                         const FloatParamsHost &fp, int64_t blocks_per_chain, const int64_t *idx, BlockStat *stats_init,
                         void *ws_zero) {
    if (nb == 0) return;
    k_build_contig<<<grid_for(nb, 256), 256, 0, L.stream>>>(descs, nb, kind, src, n, starts, tile0, chunk0, fp, blocks_per_chain, idx,
                                                             stats_init, (unsigned long long *)ws_zero);
    L.count++;
}

void launch_build_vec3(Launcher &L, BlockDesc *descs, int64_t nfiles, const float *aos, int32_t nfile,
                       int32_t subcells, const FloatParams *tab, int tab_per_file) {
    int64_t nb = nfiles * 3 * (int64_t)subcells * subcells * subcells;
    if (nb == 0) return;
    k_build_vec3<<<grid_for(nb, 256), 256, 0, L.stream>>>(descs, nfiles, aos, nfile, subcells, tab, tab_per_file);
    L.count++;
}

void launch_vec3_limits(Launcher &L, const float *aos, int64_t np_per_file, int64_t nfiles, uint32_t *keys) {
    if (nfiles == 0) return;
    k_limits_init<<<grid_for(nfiles * 6, 256), 256, 0, L.stream>>>(keys, nfiles);
    L.count++;
    if (np_per_file == 0) return;
    if (np_per_file % 4 == 0 && ((uintptr_t)aos & 15) == 0) {
        const int64_t per_cta = LIMITS4_THREADS * LIMITS4_PER_THREAD;
        dim3 grid((unsigned)((3 * np_per_file / 4 + per_cta - 1) / per_cta), (unsigned)nfiles);
        L.begin("k_vec3_limits4");
        k_vec3_limits4<<<grid, LIMITS4_THREADS, 0, L.stream>>>(aos, np_per_file, keys);
        L.end();
    } else {
        int64_t chunks = (np_per_file + LIMITS_CHUNK - 1) / LIMITS_CHUNK;
        dim3 grid((unsigned)chunks, (unsigned)nfiles);
        k_vec3_limits<<<grid, 256, 0, L.stream>>>(aos, np_per_file, keys);
    }
    L.count++;
}

void launch_init_stats(Launcher &L, const BlockDesc *descs, BlockStat *stats, int64_t nb) {
    if (nb == 0) return;
    k_init<<<grid_for(nb, 256), 256, 0, L.stream>>>(descs, stats, nb, nullptr);
    L.count++;
}

// CTAs of the grid-stride kernels: enough to fill the chip, few enough to vanish when skipped.
static inline unsigned persistent_grid(int64_t units, int per_sm) {
    int64_t cap = 148LL * per_sm;
    return (unsigned)(units < cap ? (units > 0 ? units : 1) : cap);
}

void launch_generic_encode(Launcher &L, const BlockDesc *descs, BlockStat *stats, const BatchShape &sh,
                           int64_t *slow_list, int *slow_count, int *err, int64_t *mins, int64_t *bits,
                           int64_t *offsets, int64_t *out_len, uint8_t *out, int64_t chain_stride,
                           int64_t chain_cap, const int *run_if, bool f32c, bool i64c) {
    if (sh.nblocks == 0) return;
    cudaMemsetAsync(slow_count, 0, sizeof(int), L.stream);
    k_init<<<grid_for(sh.nblocks, 256), 256, 0, L.stream>>>(descs, stats, sh.nblocks, run_if);
    L.count++;
    if (sh.total_chunks > 0) {
        // f32c and i64c together: a batch of columns of both kinds, each kernel skips the other's blocks
        if (!run_if) L.begin(f32c ? "k_stats_f32c" : (i64c ? "k_stats_i64c" : "k_stats"));
        if (f32c) k_stats_f32c<<<persistent_grid(sh.total_chunks, 8), FSTAT_THREADS, 0, L.stream>>>(descs, stats, sh, run_if);
        if (i64c) k_stats_i64c<<<persistent_grid(sh.total_chunks, 8), FSTAT_THREADS, 0, L.stream>>>(descs, stats, sh, run_if);
        if (f32c && i64c) L.count++;
        if (!f32c && !i64c) k_stats<<<persistent_grid(sh.total_chunks, 16), STATS_THREADS, 0, L.stream>>>(descs, stats, sh, run_if);
        if (!run_if) L.end();
        L.count++;
    }
    k_finalize<<<grid_for(sh.nblocks, 256), 256, 0, L.stream>>>(descs, stats, sh.nblocks, slow_list, slow_count, err, run_if);
    L.count++;
    unsigned slow_grid = (unsigned)(sh.nblocks < 296 ? sh.nblocks : 296);
    k_slow<<<slow_grid, 256, 0, L.stream>>>(descs, stats, slow_list, slow_count, err, run_if);
    L.count++;
    k_scan<<<(unsigned)sh.nchains, 1024, 0, L.stream>>>(stats, sh, mins, bits, offsets, out_len, run_if);
    L.count++;
    if (sh.total_tiles > 0) {
        if (!run_if) L.begin(f32c ? "k_pack_f32c" : (i64c ? "k_pack_i64c" : "k_pack"));
        if (f32c) k_pack_f32c<<<persistent_grid(sh.total_tiles, 12), FPACK_THREADS, 0, L.stream>>>(descs, stats, sh, out, chain_stride, chain_cap, err, run_if);
        if (i64c) k_pack_i64c<<<persistent_grid(sh.total_tiles, 12), FPACK_THREADS, 0, L.stream>>>(descs, stats, sh, out, chain_stride, chain_cap, err, run_if);
        if (f32c && i64c) L.count++;
        if (!f32c && !i64c) k_pack<<<persistent_grid(sh.total_tiles, 12), PACK_THREADS, 0, L.stream>>>(descs, stats, sh, out, chain_stride, chain_cap, err,
                                                                                        run_if, nullptr, nullptr);
        if (!run_if) L.end();
        L.count++;
        if (i64c) {   // blocks wider than 32 bits (k_pack skips the rest at once)
            k_pack<<<persistent_grid(sh.total_tiles, 12), PACK_THREADS, 0, L.stream>>>(descs, stats, sh, out, chain_stride, chain_cap, err,
                                                                                       run_if, nullptr, nullptr, 32);
            L.count++;
        }
    }
}

// Pack only the listed blocks (uniform block size), from global memory.
void launch_pack_list(Launcher &L, const BlockDesc *descs, const BlockStat *stats, const BatchShape &sh,
                      const int64_t *list, const int *list_count, uint8_t *out, int64_t chain_stride,
                      int64_t chain_cap, int *err) {
    if (sh.nblocks == 0 || sh.uniform_n <= 0) return;
    k_pack<<<persistent_grid(sh.total_tiles, 12), PACK_THREADS, 0, L.stream>>>(descs, stats, sh, out, chain_stride, chain_cap, err,
                                                                               nullptr, list, list_count);
    L.count++;
}

void launch_raw_pack(Launcher &L, BlockDesc *descs, BlockStat *stats, const void *src, int64_t n, int bits,
                     uint8_t *out) {
    if (n == 0) return;
    k_preset_raw<<<1, 1, 0, L.stream>>>(descs, stats, src, n, bits);
    L.count++;
    BatchShape sh = {1, 1, 1, n, (n + PACK_TILE - 1) / PACK_TILE, 0};
    k_pack<<<persistent_grid(sh.total_tiles, 12), PACK_THREADS, 0, L.stream>>>(descs, stats, sh, out, 0, (int64_t)1 << 62, nullptr,
                                                                               nullptr, nullptr, nullptr);
    L.count++;
}

void launch_umax(Launcher &L, const unsigned long long *x, int64_t n, unsigned long long *out) {
    cudaMemsetAsync(out, 0, 8, L.stream);
    if (n == 0) return;
    unsigned g = grid_for(n, 256);
    if (g > 1184) g = 1184;
    k_umax<<<g, 256, 0, L.stream>>>(x, n, out);
    L.count++;
}

// Decode of contiguous int64 blocks (Array.Slice + intGroup.readData, go/bit/bit.go:29-82, go/group.go:257-263): one
// CTA per 4096-element tile; the tile's packed bytes are staged in shared memory with 128-bit loads, every thread
// extracts four consecutive values and stores two 16-byte pairs.  Blocks wider than 32 bits take the 64-bit
// extraction straight from global memory.
__global__ void __launch_bounds__(FDEC_THREADS) k_decode_i64c(DecodeArgs A) {
    __shared__ __align__(16) unsigned spk[DEC_CHUNK + 16];
    const int64_t tpb = (A.n + DEC_CHUNK - 1) / DEC_CHUNK;
    const int64_t j = blockIdx.x / tpb;
    const int64_t tile = blockIdx.x - j * tpb;
    const int64_t b = A.sel ? A.sel[j] : j;
    const unsigned long long mn = (unsigned long long)A.mins[b];
    const int bits = (int)A.bits[b];
    const int64_t first = tile * DEC_CHUNK;
    const int count = (int)(first + DEC_CHUNK <= A.n ? DEC_CHUNK : A.n - first);
    long long *outp = (A.outs ? (long long *)A.outs[j] : (long long *)A.out + j * A.n) + first;
    if (bits <= 32) {
        const unsigned mask = bits >= 1 ? (0xffffffffu >> (32 - bits)) : 0u;
        unsigned shift0 = 0;
        if (bits > 0) {
            const uint8_t *src = A.data + A.offsets[b] + ((first * bits) >> 3);
            const int a16 = (int)((uintptr_t)src & 15);
            const uint4 *s16 = (const uint4 *)(src - a16);
            const int nvec = (a16 + ((count * bits + 7) >> 3) + 15) >> 4;
            for (int i = threadIdx.x; i < nvec; i += FDEC_THREADS) ((uint4 *)spk)[i] = __ldg(s16 + i);
            shift0 = 8u * (unsigned)a16;
        }
        __syncthreads();
        const bool vec_ok = (((uintptr_t)outp & 15) == 0);
        for (int e4 = threadIdx.x * 4; e4 < count; e4 += FDEC_THREADS * 4) {
            long long o[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                unsigned v = 0;
                if (bits) {
                    const unsigned bp = shift0 + (unsigned)(e4 + c) * (unsigned)bits;
                    v = __funnelshift_r(spk[bp >> 5], spk[(bp >> 5) + 1], bp) & mask;
                }
                o[c] = (long long)(mn + v);   // wrapping, like Go's int64 add
            }
            if (vec_ok && e4 + 4 <= count) {
                __stcs((longlong2 *)(outp + e4), make_longlong2(o[0], o[1]));
                __stcs((longlong2 *)(outp + e4 + 2), make_longlong2(o[2], o[3]));
            } else {
                for (int c = 0; c < 4; c++) if (e4 + c < count) outp[e4 + c] = o[c];
            }
        }
    } else {
        for (int el = threadIdx.x; el < count; el += FDEC_THREADS)
            outp[el] = (long long)(mn + extract_bits(A.data, A.stream_len, A.offsets[b], first + el, bits));
    }
}

// Device-wide form for long size arrays (the gathered sizes of a sharded snapshot): every CTA scans a tile of 4096
// sizes locally and posts its total, k_scan_sizes scans the tile totals, a last pass adds each tile's base.
constexpr int SCAN_TILE = 4096;
__global__ void __launch_bounds__(1024) k_scan_tiles(const int64_t *sizes, int64_t n, int64_t *offsets, int64_t *tile_sums) {
    __shared__ long long s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i0 = (int64_t)blockIdx.x * SCAN_TILE + 4 * threadIdx.x;
    long long v[4], t = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = i0 + k < n ? sizes[i0 + k] : 0; t += v[k]; }
    long long incl = t;
    for (int o = 1; o < 32; o <<= 1) {
        long long u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = s_warp[lane], wi = w;
        for (int o = 1; o < 32; o <<= 1) {
            long long u = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += u;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    long long excl = s_warp[warp] + incl - t;
#pragma unroll
    for (int k = 0; k < 4; k++) { if (i0 + k < n) offsets[i0 + k] = excl; excl += v[k]; }
}
__global__ void __launch_bounds__(1024) k_scan_add(int64_t *offsets, int64_t n, const int64_t *tile_offsets) {
    const int64_t i0 = (int64_t)blockIdx.x * SCAN_TILE + 4 * threadIdx.x;
    const long long base = tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; k++) if (i0 + k < n) offsets[i0 + k] += base;
}

size_t scan_scratch_bytes(int64_t n) { return n > 2 * SCAN_TILE ? 16 * (size_t)((n + SCAN_TILE - 1) / SCAN_TILE) + 64 : 0; }

cudaError_t launch_scan_sizes(Launcher &L, const int64_t *sizes, int64_t n, int64_t base, int64_t *offsets,
                              int64_t *total, void *scratch) {
    if (n > 2 * SCAN_TILE && scratch) {
        const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
        int64_t *tile_sums = (int64_t *)scratch, *tile_offs = tile_sums + ntiles;
        k_scan_tiles<<<(unsigned)ntiles, 1024, 0, L.stream>>>(sizes, n, offsets, tile_sums);
        k_scan_sizes<<<1, 1024, 0, L.stream>>>(tile_sums, ntiles, base, tile_offs, total);
        k_scan_add<<<(unsigned)ntiles, 1024, 0, L.stream>>>(offsets, n, tile_offs);
        L.count += 3;
        return cudaGetLastError();
    }
    k_scan_sizes<<<1, 1024, 0, L.stream>>>(sizes, n, base, offsets, total);
    L.count++;
    return cudaGetLastError();
}

void launch_decode(Launcher &L, const DecodeHost &h) {
    if (h.nsel == 0 || h.n == 0) return;
    DecodeArgs A = {};
    A.mode = h.mode; A.data = h.data; A.stream_len = h.stream_len;
    A.offsets = h.offsets; A.mins = h.mins; A.bits = h.bits; A.sel = h.sel; A.jitter_ids = h.jitter_ids;
    A.n = h.n; A.nsel = h.nsel;
    A.tab = h.tab; A.tab_per_file = h.tab_per_file;
    A.wrap_L = h.wrap_L; A.jmode = h.jmode; A.seed = h.seed; A.block_id0 = h.block_id0; A.u = h.u;
    A.nfile = h.nfile; A.nsub = h.subcells ? h.nfile / h.subcells : 0; A.subcells = h.subcells;
    A.sc3 = (int64_t)h.subcells * h.subcells * h.subcells;
    A.out = h.out; A.outs = h.outs;
    int64_t cpb = (h.n + DEC_CHUNK - 1) / DEC_CHUNK;
    if (h.mode == 1 && h.jmode != 2) {   // contiguous float32 blocks: staged, vectorised decode
        L.begin("k_decode_f32c");
        if (h.any_log) k_decode_f32c<true><<<(unsigned)(h.nsel * cpb), FDEC_THREADS, 0, L.stream>>>(A);
        else k_decode_f32c<false><<<(unsigned)(h.nsel * cpb), FDEC_THREADS, 0, L.stream>>>(A);
        L.end();
        L.count++;
        return;
    }
    if (h.mode == 0) {   // contiguous int64 blocks: staged, vectorised decode
        L.begin("k_decode_i64c");
        k_decode_i64c<<<(unsigned)(h.nsel * cpb), FDEC_THREADS, 0, L.stream>>>(A);
        L.end();
        L.count++;
        return;
    }
    L.begin("k_decode");
    k_decode<<<(unsigned)(h.nsel * cpb), DEC_THREADS, 0, L.stream>>>(A);
    L.end();
    L.count++;
}

}  // namespace mnw
