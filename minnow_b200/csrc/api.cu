// api.cu -- the C ABI of libminnow_b200 (include/minnow_cuda.h): context,
// device scratch management, host<->device staging, and dispatch to the
// kernels.  There is no CPU implementation of the hot path in this library.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ctx.cuh"
#include "device_math.cuh"
#include "fused.cuh"

using namespace mnw;

namespace {
std::string g_create_error;
}  // namespace

int mnw_fail(mnw_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}
#define fail mnw_fail

int mnw_report_device_error(mnw_ctx *ctx, int err) {
    if (err == 1) return fail(ctx, MNW_ERR_ARG, "block value range is 2^64-1: bit.PrecisionNeeded is undefined there");
    if (err == 2) return fail(ctx, MNW_ERR_CAPACITY, "packed output does not fit the output buffer");
    if (err == 3) return fail(ctx, MNW_ERR_ARG, "a coordinate lies outside [0, 2 L): the reference indexes outside its cell grid there");
    if (err == 4) return fail(ctx, MNW_ERR_ARG, "a particle ID is not valid for this NCell and NSide (grid.Index panics)");
    if (err == 5) return fail(ctx, MNW_ERR_ARG, "text: a requested column does not exist in the data");
    if (err == 6) return fail(ctx, MNW_ERR_FORMAT, "text: a field does not parse as a number (strconv.Atoi / ParseFloat error)");
    if (err == 7) return fail(ctx, MNW_ERR_FORMAT, "text: a line has a different number of columns than the first one");
    return MNW_OK;
}

namespace {

int check_desc(mnw_ctx *ctx, const mnw_float_desc *d) {
    if (!d) return fail(ctx, MNW_ERR_ARG, "float group descriptor is NULL");
    // pixels <= 0 (e.g. minp's non-periodic limits of a single particle, go/minp/minp.go:92-95)
    // is degenerate in the reference too; it is carried through the exact path unchanged.
    return MNW_OK;
}

FloatParamsHost to_params(const mnw_float_desc &d) {
    FloatParamsHost p = {};
    p.low = d.low; p.high = d.high; p.pixels = d.pixels;
    volatile float span = d.high - d.low;           // go/group.go:316, float32 arithmetic
    p.dx = span / (float)d.pixels;
    p.hi_clamp = nextafterf(d.high, -INFINITY);      // go/minh/minh.go:146
    p.flags = (d.periodic ? F_PERIODIC : 0) | (d.log10 ? F_LOG10 : 0) | (d.clamp ? F_CLAMP : 0);
    volatile float rcp = 1.0f / p.dx;                // correctly rounded reciprocal (quantize_fast)
    p.rcp = rcp;
    if (std::isnormal(p.dx) && std::isnormal(p.rcp) && p.dx > 0x1p-60f && p.dx < 0x1p60f && d.pixels >= 2)
        p.flags |= F_FASTDIV;
    return p;
}

// Upload a table of group parameters; returns the device pointer through *tab.
int upload_params(mnw_ctx *ctx, const mnw_float_desc *desc, int64_t count, std::vector<FloatParams> &host,
                  const FloatParams **tab) {
    host.resize((size_t)count);
    for (int64_t i = 0; i < count; i++) host[(size_t)i] = to_params(desc[i]);
    CU(ctx->params.reserve(sizeof(FloatParams) * (size_t)count));
    // pageable source: the runtime stages it before returning, so `host` may die after the call
    CU(cudaMemcpyAsync(ctx->params.p, host.data(), sizeof(FloatParams) * (size_t)count, cudaMemcpyHostToDevice,
                       ctx->L.stream));
    *tab = ctx->params.as<FloatParams>();
    return MNW_OK;
}

// Reserve the per-batch device records.
int reserve_batch(mnw_ctx *ctx, int64_t nb, int64_t nchains) {
    CU(ctx->descs.reserve(sizeof(BlockDesc) * (size_t)(nb + 1)));
    CU(ctx->stats.reserve(sizeof(BlockStat) * (size_t)(nb + 1)));
    CU(ctx->slow.reserve(sizeof(int64_t) * (size_t)(nb + 1)));
    CU(ctx->flags.reserve(64));
    CU(ctx->meta.reserve(sizeof(int64_t) * (size_t)(3 * nb + nchains + 4)));
    return MNW_OK;
}

// Per-call reset of the device flag words.  Word 1 is the ERROR word: it is sticky -- set by the kernels
// (1: value range 2^64-1, 2: packed output does not fit), read AND cleared only by check_flags(), so that an
// error raised by a `_dev` call is still there when the caller reaches mnw_sync().
int reset_flags(mnw_ctx *ctx) {
    CU(ctx->flags.reserve(64));
    int *d_flags = ctx->flags.as<int>();
    if (!ctx->flags_init) {
        CU(cudaMemsetAsync(d_flags, 0, 64, ctx->L.stream));
        ctx->flags_init = true;
    } else {
        CU(cudaMemsetAsync(d_flags, 0, FLAG_ERR * sizeof(int), ctx->L.stream));   // everything but the error word
    }
    return MNW_OK;
}

// Synchronises the stream and surfaces (then clears) the device-side error word.
int check_flags(mnw_ctx *ctx) {
    if (!ctx->flags_init) {
        CU(cudaStreamSynchronize(ctx->L.stream));
        return MNW_OK;
    }
    CU(cudaMemcpyAsync(ctx->h_flags, ctx->flags.as<int>() + FLAG_ERR, sizeof(int), cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    const int err = ctx->h_flags[0];
    if (err) CU(cudaMemsetAsync(ctx->flags.as<int>() + FLAG_ERR, 0, sizeof(int), ctx->L.stream));
    return mnw_report_device_error(ctx, err);
}

// Device-resident group encode.  x/out/mins/bits/offsets/out_len are device
// pointers (mins.. may be null).  Chooses the fused path when it applies.
int encode_group_dev(mnw_ctx *ctx, int kind, const mnw_float_desc *desc, const void *x, int64_t n,
                     int64_t nblocks, const int64_t *d_starts, const int64_t *d_tile0, const int64_t *d_chunk0,
                     int64_t total_tiles, int64_t total_chunks, int64_t *mins, int64_t *bits, int64_t *offsets,
                     uint8_t *out, int64_t out_cap, int64_t *out_len, const int64_t *d_idx = nullptr) {
    if (nblocks < 0 || n < 0) return fail(ctx, MNW_ERR_ARG, "negative block count or length");
    FloatParamsHost fp = {};
    if (kind == KIND_F32) {
        int rc = check_desc(ctx, desc);
        if (rc) return rc;
        fp = to_params(*desc);
    }
    int rc = reserve_batch(ctx, nblocks, 1);
    if (rc) return rc;
    int *d_flags = ctx->flags.as<int>();
    { const int rcf = reset_flags(ctx); if (rcf) return rcf; }
    if (nblocks == 0) {
        if (out_len) CU(cudaMemsetAsync(out_len, 0, 8, ctx->L.stream));
        return MNW_OK;
    }
    BatchShape sh = {};
    sh.nblocks = nblocks; sh.nchains = 1; sh.blocks_per_chain = nblocks;
    sh.uniform_n = d_starts ? 0 : n;
    if (d_starts) {
        sh.total_tiles = total_tiles; sh.total_chunks = total_chunks;
    } else {
        sh.total_tiles = nblocks * ((n + PACK_TILE - 1) / PACK_TILE);
        sh.total_chunks = nblocks * ((n + STATS_CHUNK - 1) / STATS_CHUNK);
    }
    if (sh.total_tiles >= (1LL << 31) || sh.total_chunks >= (1LL << 31))
        return fail(ctx, MNW_ERR_ARG, "batch too large for one launch");

    ctx->last_path = 0;
    // contiguous float32 blocks of a periodic group with pixels < 2^31: the vectorised kernels
    const bool f32c = !ctx->force_generic && !d_idx && kind == KIND_F32 && (fp.flags & F_PERIODIC) && fp.pixels >= 1 &&
                      fp.pixels < (1LL << 31);
    // contiguous int64 blocks: the vectorised kernels (blocks wider than 32 bits fall to k_pack)
    const bool i64c = !ctx->force_generic && !d_idx && kind == KIND_I64;
    // uniform contiguous blocks: the fused single-read kernel (MNW_GROUP=twopass keeps the two-pass kernels, for A/B runs)
    static const bool twopass = getenv("MNW_GROUP") && !strcmp(getenv("MNW_GROUP"), "twopass");
    // (TMA tile fetches: every block must start on a 16-byte boundary)
    const bool aligned = ((uintptr_t)x & 15) == 0 && (nblocks == 1 || (n * (kind == KIND_I64 ? 8 : 4)) % 16 == 0);
    const bool fused = (f32c || i64c) && !d_starts && n > 0 && aligned && !twopass;
    if (fused) CU(ctx->group_ws.reserve(group_fused_ws_bytes(nblocks)));
    // log10 columns: the statistics pass leaves float32(log10 x) for the pack pass (one logarithm per value)
    const bool logcol = fused && kind == KIND_F32 && (fp.flags & F_LOG10);
    if (logcol) CU(ctx->group_log.reserve(4 * (size_t)nblocks * (size_t)n + 64));
    launch_build_contig(ctx->L, ctx->descs.as<BlockDesc>(), nblocks, kind, x, n, d_starts, d_tile0, d_chunk0, fp, nblocks, d_idx,
                        fused ? ctx->stats.as<BlockStat>() : nullptr, fused ? ctx->group_ws.p : nullptr);
    if (fused) {
        ctx->last_path = 2;
        const cudaError_t e = launch_group_encode(ctx->L, ctx->descs.as<BlockDesc>(), ctx->stats.as<BlockStat>(), sh, d_flags, mins, bits,
                                                  offsets, out_len, out, 0, out_cap, ctx->group_ws.p, kind == KIND_I64, true,
                                                  logcol ? ctx->group_log.p : nullptr);
        if (e != cudaSuccess) return fail(ctx, MNW_ERR_CUDA, "fused group encode: %s", cudaGetErrorString(e));
        return MNW_OK;
    }
    launch_generic_encode(ctx->L, ctx->descs.as<BlockDesc>(), ctx->stats.as<BlockStat>(), sh,
                          ctx->slow.as<int64_t>(), d_flags, d_flags + FLAG_ERR, mins, bits, offsets, out_len, out,
                          0, out_cap, nullptr, f32c, i64c);
    CU(cudaGetLastError());
    return MNW_OK;
}

// Host-pointer group encode: stage in, run, stage out.
int encode_group_host(mnw_ctx *ctx, int kind, const mnw_float_desc *desc, const void *x, int64_t n,
                      int64_t nblocks, const int64_t *starts, int64_t *mins, int64_t *bits, int64_t *offsets,
                      uint8_t *out, int64_t out_cap, int64_t *out_len, const int64_t *idx = nullptr, int64_t ncol = 0,
                      const int64_t *d_idx_resident = nullptr, const void *x_dev = nullptr) {
    // d_idx_resident: the gather index is on the device already (mnw_boundary_coordinates); x_dev: so are the elements
    if (nblocks < 0 || n < 0) return fail(ctx, MNW_ERR_ARG, "negative block count or length");
    const size_t esz = kind == KIND_I64 ? 8 : 4;
    int64_t total = starts ? starts[nblocks] - starts[0] : n * nblocks;
    if (starts && starts[0] != 0) return fail(ctx, MNW_ERR_ARG, "starts[0] must be 0");
    const int64_t nsrc = (idx || d_idx_resident) ? ncol : total;   // elements to upload: the whole column for a gather
    const int64_t *d_idx = d_idx_resident;
    if (idx) {
        if (!starts) return fail(ctx, MNW_ERR_ARG, "a gather needs starts[]");
        for (int64_t i = 0; i < total; i++)
            if (idx[i] < 0 || idx[i] >= ncol) return fail(ctx, MNW_ERR_ARG, "gather index %lld outside the column of %lld", (long long)idx[i], (long long)ncol);
        CU(ctx->ustream.reserve(8 * (size_t)total + 16));
        if (total > 0) CU(cudaMemcpyAsync(ctx->ustream.p, idx, 8 * (size_t)total, cudaMemcpyHostToDevice, ctx->L.stream));
        d_idx = ctx->ustream.as<int64_t>();
    }
    CU(ctx->in.reserve(esz * (size_t)nsrc + 16));
    CU(ctx->out.reserve(8 * (size_t)total + 64));
    int rc = reserve_batch(ctx, nblocks, 1);
    if (rc) return rc;
    if (nsrc > 0 && !x_dev) CU(cudaMemcpyAsync(ctx->in.p, x, esz * (size_t)nsrc, cudaMemcpyHostToDevice, ctx->L.stream));

    const int64_t *d_starts = nullptr, *d_tile0 = nullptr, *d_chunk0 = nullptr;
    int64_t total_tiles = 0, total_chunks = 0;
    std::vector<int64_t> h_aux;
    if (starts) {
        h_aux.resize(3 * (size_t)(nblocks + 1));
        int64_t *hs = h_aux.data(), *ht = hs + nblocks + 1, *hc = ht + nblocks + 1;
        for (int64_t b = 0; b <= nblocks; b++) hs[b] = starts[b];
        for (int64_t b = 0; b < nblocks; b++) {
            int64_t nb = starts[b + 1] - starts[b];
            if (nb < 0) return fail(ctx, MNW_ERR_ARG, "starts must be non-decreasing");
            ht[b] = total_tiles; hc[b] = total_chunks;
            total_tiles += (nb + PACK_TILE - 1) / PACK_TILE;
            total_chunks += (nb + STATS_CHUNK - 1) / STATS_CHUNK;
        }
        ht[nblocks] = total_tiles; hc[nblocks] = total_chunks;
        CU(ctx->aux.reserve(h_aux.size() * 8));
        CU(cudaMemcpyAsync(ctx->aux.p, h_aux.data(), h_aux.size() * 8, cudaMemcpyHostToDevice, ctx->L.stream));
        d_starts = ctx->aux.as<int64_t>();
        d_tile0 = d_starts + nblocks + 1;
        d_chunk0 = d_tile0 + nblocks + 1;
    }
    int64_t *d_meta = ctx->meta.as<int64_t>();
    int64_t *d_mins = d_meta, *d_bits = d_meta + nblocks, *d_offs = d_meta + 2 * nblocks, *d_len = d_meta + 3 * nblocks;
    rc = encode_group_dev(ctx, kind, desc, x_dev ? x_dev : ctx->in.p, n, nblocks, d_starts, d_tile0, d_chunk0, total_tiles,
                          total_chunks, d_mins, d_bits, d_offs, ctx->out.as<uint8_t>(), (int64_t)ctx->out.cap, d_len, d_idx);
    if (rc) return rc;
    std::vector<int64_t> h_meta(3 * (size_t)nblocks + 1);
    CU(cudaMemcpyAsync(h_meta.data(), d_meta, h_meta.size() * 8, cudaMemcpyDeviceToHost, ctx->L.stream));
    rc = check_flags(ctx);  // synchronises
    if (rc) return rc;
    int64_t len = h_meta[3 * (size_t)nblocks];
    if (mins) memcpy(mins, h_meta.data(), 8 * (size_t)nblocks);
    if (bits) memcpy(bits, h_meta.data() + nblocks, 8 * (size_t)nblocks);
    if (offsets) memcpy(offsets, h_meta.data() + 2 * nblocks, 8 * (size_t)nblocks);
    if (out_len) *out_len = len;
    if (len > out_cap) return fail(ctx, MNW_ERR_CAPACITY, "group needs %lld bytes, buffer has %lld", (long long)len, (long long)out_cap);
    if (len > 0) {
        CU(cudaMemcpyAsync(out, ctx->out.p, (size_t)len, cudaMemcpyDeviceToHost, ctx->L.stream));
        CU(cudaStreamSynchronize(ctx->L.stream));
    }
    return MNW_OK;
}

// minh.Writer.Block: every quantised column of one block in one batch (one chain per column, one block per chain).
int encode_columns_host(mnw_ctx *ctx, int64_t ncols, const mnw_column *cols, const void *const *data, int64_t n,
                        int64_t *mins, int64_t *bits, int64_t *nbytes, uint8_t *out, int64_t out_col_stride) {
    if (ncols < 0 || n < 0 || out_col_stride < 0) return fail(ctx, MNW_ERR_ARG, "negative column count, length or stride");
    if (ncols == 0) return MNW_OK;
    if (!cols || !data) return fail(ctx, MNW_ERR_ARG, "null column table");
    const int64_t tpb = (n + PACK_TILE - 1) / PACK_TILE, cpb = (n + STATS_CHUNK - 1) / STATS_CHUNK;
    if (ncols * tpb >= (1LL << 31)) return fail(ctx, MNW_ERR_ARG, "batch too large for one launch");
    std::vector<size_t> off((size_t)ncols);
    size_t tot = 0;
    bool any_f = false, any_i = false, degenerate = false;
    int64_t log_hi = 0;   // columns below this index may be log10 columns (scratch of the fused encoder)
    for (int64_t c = 0; c < ncols; c++) {
        off[(size_t)c] = tot;
        tot += ((size_t)(cols[c].is_float ? 4 : 8) * (size_t)n + 15) & ~(size_t)15;
        if (!data[c] && n > 0) return fail(ctx, MNW_ERR_ARG, "column %lld has no data", (long long)c);
        if (cols[c].is_float) {
            int rc = check_desc(ctx, &cols[c].desc);
            if (rc) return rc;
            if (!cols[c].desc.periodic || cols[c].desc.pixels >= (1LL << 31))
                return fail(ctx, MNW_ERR_ARG, "column %lld: mnw_encode_columns takes periodic FloatGroups with pixels < 2^31", (long long)c);
            // Low == High, or a NaN / zero / negative Dx: pixels <= 0 (or INT64_MIN).  The reference carries such a
            // group through its int64 arithmetic; so do the exact kernels (as encode_group_dev does)
            if (cols[c].desc.pixels < 1) degenerate = true;
            any_f = true;
            if (cols[c].desc.log10) log_hi = c + 1;
        } else {
            any_i = true;
        }
    }
    const size_t dstride = (((size_t)8 * (size_t)n + 15) & ~(size_t)15) + 16;   // ArrayBytes(bits <= 64, n) fits
    CU(ctx->in.reserve(tot + 16));
    CU(ctx->out.reserve(dstride * (size_t)ncols + 64));
    int rc = reserve_batch(ctx, ncols, ncols);
    if (rc) return rc;
    int *d_flags = ctx->flags.as<int>();
    { const int rcf = reset_flags(ctx); if (rcf) return rcf; }
    std::vector<BlockDesc> hd((size_t)ncols);
    for (int64_t c = 0; c < ncols; c++) {
        BlockDesc d = {};
        d.src = (const uint8_t *)ctx->in.p + off[(size_t)c];
        d.n = n; d.access = ACC_CONTIG; d.chain = (int32_t)c;
        d.tile0 = c * tpb; d.chunk0 = c * cpb;
        if (cols[c].is_float) {
            const FloatParamsHost fp = to_params(cols[c].desc);
            d.kind = KIND_F32; d.flags = fp.flags;
            d.low = fp.low; d.high = fp.high; d.dx = fp.dx; d.hi_clamp = fp.hi_clamp; d.pixels = fp.pixels;
        } else {
            d.kind = KIND_I64;
        }
        hd[(size_t)c] = d;
        if (n > 0) CU(cudaMemcpyAsync((uint8_t *)ctx->in.p + off[(size_t)c], data[c], (size_t)(cols[c].is_float ? 4 : 8) * (size_t)n,
                                      cudaMemcpyHostToDevice, ctx->L.stream));
    }
    CU(cudaMemcpyAsync(ctx->descs.p, hd.data(), sizeof(BlockDesc) * (size_t)ncols, cudaMemcpyHostToDevice, ctx->L.stream));
    BatchShape sh = {};
    sh.nblocks = ncols; sh.nchains = ncols; sh.blocks_per_chain = 1; sh.uniform_n = n;
    sh.total_tiles = ncols * tpb; sh.total_chunks = ncols * cpb;
    ctx->last_path = 0;
    int64_t *d_meta = ctx->meta.as<int64_t>();
    int64_t *d_mins = d_meta, *d_bits = d_meta + ncols, *d_offs = d_meta + 2 * ncols, *d_len = d_meta + 3 * ncols;
    const bool fast = !ctx->force_generic && !degenerate;
    static const bool twopass = getenv("MNW_GROUP") && !strcmp(getenv("MNW_GROUP"), "twopass");
    if (fast && n > 0 && !twopass) {   // every column is one uniform contiguous block of its own chain: the fused single-read kernel
        CU(ctx->group_ws.reserve(group_fused_ws_bytes(ncols)));
        if (log_hi) CU(ctx->group_log.reserve(4 * (size_t)log_hi * (size_t)n + 64));
        ctx->last_path = 2;
        const cudaError_t e = launch_group_encode(ctx->L, ctx->descs.as<BlockDesc>(), ctx->stats.as<BlockStat>(), sh, d_flags, d_mins,
                                                  d_bits, d_offs, d_len, ctx->out.as<uint8_t>(), (int64_t)dstride, (int64_t)dstride,
                                                  ctx->group_ws.p, any_i, false, log_hi ? ctx->group_log.p : nullptr);
        if (e != cudaSuccess) return fail(ctx, MNW_ERR_CUDA, "fused column encode: %s", cudaGetErrorString(e));
    } else {
        launch_generic_encode(ctx->L, ctx->descs.as<BlockDesc>(), ctx->stats.as<BlockStat>(), sh, ctx->slow.as<int64_t>(),
                              d_flags, d_flags + FLAG_ERR, d_mins, d_bits, d_offs, d_len, ctx->out.as<uint8_t>(), (int64_t)dstride,
                              (int64_t)dstride, nullptr, fast && any_f, fast && any_i);
    }
    CU(cudaGetLastError());
    std::vector<int64_t> h_meta(4 * (size_t)ncols);
    CU(cudaMemcpyAsync(h_meta.data(), d_meta, h_meta.size() * 8, cudaMemcpyDeviceToHost, ctx->L.stream));
    rc = check_flags(ctx);   // synchronises
    if (rc) return rc;
    for (int64_t c = 0; c < ncols; c++) {
        const int64_t len = h_meta[3 * (size_t)ncols + (size_t)c];
        if (mins) mins[c] = h_meta[(size_t)c];
        if (bits) bits[c] = h_meta[(size_t)ncols + (size_t)c];
        if (nbytes) nbytes[c] = len;
        if (len > out_col_stride) return fail(ctx, MNW_ERR_CAPACITY, "column %lld needs %lld bytes, stride is %lld", (long long)c, (long long)len, (long long)out_col_stride);
        if (len > 0) CU(cudaMemcpyAsync(out + c * out_col_stride, ctx->out.as<uint8_t>() + (size_t)c * dstride, (size_t)len,
                                        cudaMemcpyDeviceToHost, ctx->L.stream));
    }
    CU(cudaStreamSynchronize(ctx->L.stream));
    return MNW_OK;
}

int fill_decode_float(mnw_ctx *ctx, DecodeHost &h, const mnw_float_desc *desc, int64_t ndesc, const mnw_jitter *jitter) {
    int rc = check_desc(ctx, desc);
    if (rc) return rc;
    std::vector<FloatParams> host;
    rc = upload_params(ctx, desc, ndesc, host, &h.tab);
    if (rc) return rc;
    h.low_nonneg = true;   // every decoded value dx * t + low is then >= +0: the periodic wrap needs one test only
    h.any_log = false;
    for (const FloatParams &p : host) {
        h.low_nonneg = h.low_nonneg && p.low >= 0.0f && p.dx > 0.0f;
        h.any_log = h.any_log || (p.flags & F_LOG10);
    }
    if (jitter) {
        if (jitter->mode < 0 || jitter->mode > 2) return fail(ctx, MNW_ERR_ARG, "unknown jitter mode %d", jitter->mode);
        h.jmode = jitter->mode; h.seed = jitter->seed; h.block_id0 = jitter->block_id0;
    }
    return MNW_OK;
}

// Host-pointer group decode (int or float).
int decode_blocks_host(mnw_ctx *ctx, int mode, const mnw_float_desc *desc, const uint8_t *data, int64_t data_len,
                       const int64_t *offsets, const int64_t *mins, const int64_t *bits, int64_t n, int64_t nsel,
                       const int64_t *sel, const mnw_jitter *jitter, void *out) {
    if (n < 0 || nsel < 0 || data_len < 0) return fail(ctx, MNW_ERR_ARG, "negative length");
    if (nsel == 0 || n == 0) return MNW_OK;
    DecodeHost h;
    h.mode = mode;
    if (mode == 1) {
        int rc = fill_decode_float(ctx, h, desc, 1, jitter);
        if (rc) return rc;
    }
    // Upload only the bytes of the selected blocks, compacted back to back.
    std::vector<int64_t> m(3 * (size_t)nsel);
    int64_t *c_off = m.data(), *c_min = c_off + nsel, *c_bits = c_min + nsel;
    int64_t total = 0;
    for (int64_t j = 0; j < nsel; j++) {
        int64_t b = sel ? sel[j] : j;
        if (b < 0) return fail(ctx, MNW_ERR_ARG, "negative block id");
        if (bits[b] < 0 || bits[b] > 64) return fail(ctx, MNW_ERR_FORMAT, "block %lld has %lld bits", (long long)b, (long long)bits[b]);
        int64_t nb = array_bytes(bits[b], n);
        if (offsets[b] < 0 || offsets[b] + nb > data_len)
            return fail(ctx, MNW_ERR_FORMAT, "block %lld lies outside the group's data", (long long)b);
        c_off[j] = total; c_min[j] = mins[b]; c_bits[j] = bits[b];
        total += nb;
    }
    bool contiguous = true;
    for (int64_t j = 0; j + 1 < nsel && contiguous; j++) {
        int64_t b = sel ? sel[j] : j, b2 = sel ? sel[j + 1] : j + 1;
        contiguous = offsets[b] + (c_off[j + 1] - c_off[j]) == offsets[b2];
    }
    // The selected blocks' bytes go up in ONE copy: as they lie when they are contiguous; the whole span they cover
    // when they make up a quarter of it or more (random access to a large fraction of a group); otherwise gathered
    // back to back into pinned staging by the host first.  (One copy per block would cost microseconds per block.)
    int64_t span_lo = 0, span_hi = 0;
    if (!contiguous) {
        span_lo = INT64_MAX;
        for (int64_t j = 0; j < nsel; j++) {
            const int64_t b = sel ? sel[j] : j, nbj = (j + 1 < nsel ? c_off[j + 1] : total) - c_off[j];
            span_lo = offsets[b] < span_lo ? offsets[b] : span_lo;
            span_hi = offsets[b] + nbj > span_hi ? offsets[b] + nbj : span_hi;
        }
    }
    const bool whole_span = !contiguous && 4 * total >= span_hi - span_lo;
    CU(ctx->in.reserve((size_t)(whole_span ? span_hi - span_lo : total) + 16));
    if (contiguous) {
        int64_t b = sel ? sel[0] : 0;
        if (total > 0) CU(cudaMemcpyAsync(ctx->in.p, data + offsets[b], (size_t)total, cudaMemcpyHostToDevice, ctx->L.stream));
    } else if (whole_span) {
        CU(cudaMemcpyAsync(ctx->in.p, data + span_lo, (size_t)(span_hi - span_lo), cudaMemcpyHostToDevice, ctx->L.stream));
        for (int64_t j = 0; j < nsel; j++) c_off[j] = offsets[sel ? sel[j] : j] - span_lo;   // blocks stay where they are
        total = span_hi - span_lo;
    } else {
        if ((size_t)total > ctx->h_stage_cap) {
            if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
            ctx->h_stage = nullptr; ctx->h_stage_cap = 0;
            CU(cudaMallocHost(&ctx->h_stage, (size_t)total + (size_t)total / 4 + 4096));
            ctx->h_stage_cap = (size_t)total + (size_t)total / 4 + 4096;
        }
        CU(cudaStreamSynchronize(ctx->L.stream));   // an earlier call's copy out of the staging has finished
        for (int64_t j = 0; j < nsel; j++) {
            const int64_t nbj = (j + 1 < nsel ? c_off[j + 1] : total) - c_off[j], b = sel ? sel[j] : j;
            if (nbj > 0) memcpy((uint8_t *)ctx->h_stage + c_off[j], data + offsets[b], (size_t)nbj);
        }
        if (total > 0) CU(cudaMemcpyAsync(ctx->in.p, ctx->h_stage, (size_t)total, cudaMemcpyHostToDevice, ctx->L.stream));
    }
    CU(ctx->meta.reserve(m.size() * 8));
    CU(cudaMemcpyAsync(ctx->meta.p, m.data(), m.size() * 8, cudaMemcpyHostToDevice, ctx->L.stream));
    const size_t osz = mode == 0 ? 8 : 4;
    CU(ctx->dec_out.reserve(osz * (size_t)(n * nsel)));
    if (mode == 1 && h.jmode == 2) {
        if (!jitter->u_stream) return fail(ctx, MNW_ERR_ARG, "jitter mode STREAM without u_stream");
        CU(ctx->ustream.reserve(8 * (size_t)(n * nsel)));
        CU(cudaMemcpyAsync(ctx->ustream.p, jitter->u_stream, 8 * (size_t)(n * nsel), cudaMemcpyHostToDevice, ctx->L.stream));
        h.u = ctx->ustream.as<double>();
    }
    // jitter block ids must stay those of the ORIGINAL blocks: pass them through sel
    // as a device array of original ids when a hash jitter is asked for.
    h.data = ctx->in.as<uint8_t>();
    h.stream_len = total;
    h.offsets = ctx->meta.as<int64_t>();
    h.mins = h.offsets + nsel;
    h.bits = h.mins + nsel;
    h.n = n; h.nsel = nsel; h.out = ctx->dec_out.p;
    if (mode == 1 && h.jmode == 1 && sel) {
        CU(ctx->aux.reserve(8 * (size_t)nsel));
        CU(cudaMemcpyAsync(ctx->aux.p, sel, 8 * (size_t)nsel, cudaMemcpyHostToDevice, ctx->L.stream));
        h.jitter_ids = ctx->aux.as<int64_t>();
    }
    launch_decode(ctx->L, h);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, ctx->dec_out.p, osz * (size_t)(n * nsel), cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    return MNW_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// exported entry points
// ---------------------------------------------------------------------------
extern "C" {

const char *mnw_version(void) { return "minnow_b200 0.1 (sm_100a)"; }

int mnw_create(int device, mnw_ctx **out) {
    mnw_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, MNW_ERR_ARG, "mnw_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        return fail(nullptr, MNW_ERR_CUDA, "mnw_create: no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= count) return fail(nullptr, MNW_ERR_ARG, "mnw_create: device %d of %d", device, count);
    CU(cudaSetDevice(device));
    ctx = new mnw_ctx();
    ctx->device = device;
    e = cudaStreamCreateWithFlags(&ctx->L.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_flags, 64);
    if (e != cudaSuccess) {
        delete ctx;
        return fail(nullptr, MNW_ERR_CUDA, "mnw_create: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return MNW_OK;
}

void mnw_destroy(mnw_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    mnw_comm_destroy(ctx);
    cudaStreamSynchronize(ctx->L.stream);
    for (DevBuf *b : {&ctx->in, &ctx->out, &ctx->descs, &ctx->stats, &ctx->slow, &ctx->flags, &ctx->meta,
                      &ctx->aux, &ctx->dec_out, &ctx->ustream, &ctx->fused_ws, &ctx->params, &ctx->coop_ws, &ctx->group_ws, &ctx->group_log, &ctx->dec_cols,
                      &ctx->bnd_idx, &ctx->bnd_flags, &ctx->bnd_work, &ctx->txt_work, &ctx->txt_i, &ctx->txt_f, &ctx->txt_fb})
        b->release();
    if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    cudaStreamDestroy(ctx->L.stream);
    delete ctx;
}

const char *mnw_last_error(const mnw_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int mnw_sync(mnw_ctx *ctx) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    return check_flags(ctx);            /* waits for the stream; reports what the `_dev` calls could not */
}

void *mnw_stream(mnw_ctx *ctx) { return (void *)ctx->L.stream; }
int64_t mnw_launch_count(const mnw_ctx *ctx) { return ctx->L.count; }
int mnw_last_path(const mnw_ctx *ctx) { return ctx->last_path; }
void mnw_force_generic(mnw_ctx *ctx, int on) { ctx->force_generic = on; }

int mnw_precision_needed(uint64_t max) {
    int b = precision_needed(max);
    return b < 0 ? MNW_ERR_ARG : b;
}
int64_t mnw_array_bytes(int bits, int64_t n) { return array_bytes(bits, n); }
int64_t mnw_float_group_pixels(float lo, float hi, float dx) {
    volatile float span = hi - lo;
    volatile float r = span / dx;
    double c = ceil((double)r);
    if (!(c >= -9223372036854775808.0 && c < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)c;
}
uint32_t mnw_jitter_hash32(uint64_t seed, uint64_t block_id, uint64_t i) { return jitter_hash32(seed, block_id, i); }

int mnw_pack(mnw_ctx *ctx, int bits, const uint64_t *x, int64_t n, uint8_t *out) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    if (bits > 64) return fail(ctx, MNW_ERR_ARG, "Cannot pack more than 64 bits per element into a bit.Array");
    if (bits < 1 || n < 0) return fail(ctx, MNW_ERR_ARG, "mnw_pack: bits = %d, n = %lld", bits, (long long)n);
    if (n == 0) return MNW_OK;
    int64_t nb = array_bytes(bits, n);
    CU(ctx->in.reserve(8 * (size_t)n));
    CU(ctx->out.reserve((size_t)nb + 64));
    int rc = reserve_batch(ctx, 1, 1);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->in.p, x, 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->L.stream));
    launch_raw_pack(ctx->L, ctx->descs.as<BlockDesc>(), ctx->stats.as<BlockStat>(), ctx->in.p, n, bits, ctx->out.as<uint8_t>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, ctx->out.p, (size_t)nb, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    return MNW_OK;
}

int mnw_unpack(mnw_ctx *ctx, int bits, const uint8_t *in, int64_t n, uint64_t *out) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    if (bits < 1 || bits > 64 || n < 0) return fail(ctx, MNW_ERR_ARG, "mnw_unpack: bits = %d, n = %lld", bits, (long long)n);
    int64_t off = 0, mn = 0, bt = bits;
    return decode_blocks_host(ctx, 0, nullptr, in, array_bytes(bits, n), &off, &mn, &bt, n, 1, nullptr, nullptr, out);
}

int mnw_bits(mnw_ctx *ctx, const uint64_t *x, int64_t n, int *bits) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    if (n < 0 || !bits) return fail(ctx, MNW_ERR_ARG, "mnw_bits: bad argument");
    if (n == 0) { *bits = 0; return MNW_OK; }
    CU(ctx->in.reserve(8 * (size_t)n));
    CU(ctx->meta.reserve(64));
    CU(cudaMemcpyAsync(ctx->in.p, x, 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->L.stream));
    launch_umax(ctx->L, ctx->in.as<unsigned long long>(), n, ctx->meta.as<unsigned long long>());
    CU(cudaGetLastError());
    unsigned long long mx = 0;
    CU(cudaMemcpyAsync(&mx, ctx->meta.p, 8, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    int b = precision_needed(mx);
    if (b < 0) return fail(ctx, MNW_ERR_ARG, "bit.PrecisionNeeded is undefined for 2^64-1");
    *bits = b;
    return MNW_OK;
}

int mnw_encode_int_group(mnw_ctx *ctx, const int64_t *x, int64_t n, int64_t nblocks, const int64_t *starts,
                         int64_t *mins, int64_t *bits, int64_t *offsets, uint8_t *out, int64_t out_cap,
                         int64_t *out_len) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    return encode_group_host(ctx, KIND_I64, nullptr, x, n, nblocks, starts, mins, bits, offsets, out, out_cap, out_len);
}

int mnw_encode_float_group(mnw_ctx *ctx, const mnw_float_desc *desc, const float *x, int64_t n, int64_t nblocks,
                           const int64_t *starts, int64_t *mins, int64_t *bits, int64_t *offsets, uint8_t *out,
                           int64_t out_cap, int64_t *out_len) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    int rc = check_desc(ctx, desc);
    if (rc) return rc;
    return encode_group_host(ctx, KIND_F32, desc, x, n, nblocks, starts, mins, bits, offsets, out, out_cap, out_len);
}

int mnw_encode_columns(mnw_ctx *ctx, int64_t ncols, const mnw_column *cols, const void *const *data, int64_t n,
                       int64_t *mins, int64_t *bits, int64_t *nbytes, uint8_t *out, int64_t out_col_stride) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    return encode_columns_host(ctx, ncols, cols, data, n, mins, bits, nbytes, out, out_col_stride);
}

int mnw_encode_columns_dev(mnw_ctx *ctx, int64_t ncols, const mnw_column *cols, const void *const *data_dev, int64_t n,
                           int64_t *mins, int64_t *bits, int64_t *nbytes, uint8_t *out, int64_t out_col_stride) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);
    if (ncols < 0 || n < 0 || out_col_stride < 0) return fail(ctx, MNW_ERR_ARG, "negative column count, length or stride");
    if (ncols == 0) return MNW_OK;
    if (!cols || !data_dev) return fail(ctx, MNW_ERR_ARG, "null column table");
    const int64_t tpb = (n + PACK_TILE - 1) / PACK_TILE, cpb = (n + STATS_CHUNK - 1) / STATS_CHUNK;
    if (ncols * tpb >= (1LL << 31)) return fail(ctx, MNW_ERR_ARG, "batch too large for one launch");
    bool any_f = false, any_i = false, degenerate = false, aligned = true;
    int64_t log_hi = 0;
    std::vector<BlockDesc> hd((size_t)ncols);
    for (int64_t c = 0; c < ncols; c++) {
        if (!data_dev[c] && n > 0) return fail(ctx, MNW_ERR_ARG, "column %lld has no data", (long long)c);
        BlockDesc d = {};
        d.src = data_dev[c];
        d.n = n; d.access = ACC_CONTIG; d.chain = (int32_t)c;
        d.tile0 = c * tpb; d.chunk0 = c * cpb;
        aligned = aligned && ((uintptr_t)data_dev[c] & 15) == 0;
        if (cols[c].is_float) {
            if (!cols[c].desc.periodic || cols[c].desc.pixels >= (1LL << 31))
                return fail(ctx, MNW_ERR_ARG, "column %lld: mnw_encode_columns takes periodic FloatGroups with pixels < 2^31", (long long)c);
            if (cols[c].desc.pixels < 1) degenerate = true;
            const FloatParamsHost fp = to_params(cols[c].desc);
            d.kind = KIND_F32; d.flags = fp.flags;
            d.low = fp.low; d.high = fp.high; d.dx = fp.dx; d.hi_clamp = fp.hi_clamp; d.pixels = fp.pixels;
            any_f = true;
            if (cols[c].desc.log10) log_hi = c + 1;
        } else {
            d.kind = KIND_I64;
            any_i = true;
        }
        hd[(size_t)c] = d;
    }
    int rc = reserve_batch(ctx, ncols, ncols);
    if (rc) return rc;
    int *d_flags = ctx->flags.as<int>();
    { const int rcf = reset_flags(ctx); if (rcf) return rcf; }
    CU(cudaMemcpyAsync(ctx->descs.p, hd.data(), sizeof(BlockDesc) * (size_t)ncols, cudaMemcpyHostToDevice, ctx->L.stream));   // pageable: staged before return
    BatchShape sh = {};
    sh.nblocks = ncols; sh.nchains = ncols; sh.blocks_per_chain = 1; sh.uniform_n = n;
    sh.total_tiles = ncols * tpb; sh.total_chunks = ncols * cpb;
    CU(ctx->meta.reserve(8 * (size_t)ncols + 64));
    int64_t *d_offs = ctx->meta.as<int64_t>();   // (every column is block 0 of its own group: offsets are all 0)
    const bool fast = !ctx->force_generic && !degenerate;
    static const bool twopass = getenv("MNW_GROUP") && !strcmp(getenv("MNW_GROUP"), "twopass");
    if (fast && n > 0 && aligned && !twopass) {
        CU(ctx->group_ws.reserve(group_fused_ws_bytes(ncols)));
        if (log_hi) CU(ctx->group_log.reserve(4 * (size_t)log_hi * (size_t)n + 64));
        ctx->last_path = 2;
        const cudaError_t e = launch_group_encode(ctx->L, ctx->descs.as<BlockDesc>(), ctx->stats.as<BlockStat>(), sh, d_flags, mins, bits,
                                                  d_offs, nbytes, out, out_col_stride, out_col_stride, ctx->group_ws.p, any_i, false,
                                                  log_hi ? ctx->group_log.p : nullptr);
        if (e != cudaSuccess) return fail(ctx, MNW_ERR_CUDA, "fused column encode: %s", cudaGetErrorString(e));
    } else {
        ctx->last_path = 0;
        launch_generic_encode(ctx->L, ctx->descs.as<BlockDesc>(), ctx->stats.as<BlockStat>(), sh, ctx->slow.as<int64_t>(),
                              d_flags, d_flags + FLAG_ERR, mins, bits, d_offs, nbytes, out, out_col_stride, out_col_stride, nullptr,
                              fast && any_f, fast && any_i);
    }
    CU(cudaGetLastError());
    return MNW_OK;
}

int mnw_encode_int_group_gather(mnw_ctx *ctx, const int64_t *col, int64_t ncol, const int64_t *idx, int64_t nblocks,
                                const int64_t *starts, int64_t *mins, int64_t *bits, int64_t *offsets, uint8_t *out,
                                int64_t out_cap, int64_t *out_len) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    if (ncol < 0 || !starts) return fail(ctx, MNW_ERR_ARG, "gather: bad column length or no starts[]");
    return encode_group_host(ctx, KIND_I64, nullptr, col, 0, nblocks, starts, mins, bits, offsets, out, out_cap, out_len, idx, ncol);
}

int mnw_encode_float_group_gather(mnw_ctx *ctx, const mnw_float_desc *desc, const float *col, int64_t ncol,
                                  const int64_t *idx, int64_t nblocks, const int64_t *starts, int64_t *mins,
                                  int64_t *bits, int64_t *offsets, uint8_t *out, int64_t out_cap, int64_t *out_len) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    int rc = check_desc(ctx, desc);
    if (rc) return rc;
    if (ncol < 0 || !starts) return fail(ctx, MNW_ERR_ARG, "gather: bad column length or no starts[]");
    return encode_group_host(ctx, KIND_F32, desc, col, 0, nblocks, starts, mins, bits, offsets, out, out_cap, out_len, idx, ncol);
}

// ---- text.Reader.Block (go/text/text.go:181-200, go/text/parse.go) ---------------------------------------------------------
int mnw_text_parse_block(mnw_ctx *ctx, const char *buf, int64_t len, char sep, char comment, int n_icols, const int *icols,
                         int n_fcols, const int *fcols, int64_t *nrows, int64_t *nfallback) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);
    if (len < 0 || n_icols < 0 || n_fcols < 0 || n_icols > 64 || n_fcols > 64) return fail(ctx, MNW_ERR_ARG, "mnw_text_parse_block: bad argument");
    for (int k = 0; k + 1 < n_icols; k++) if (icols[k] >= icols[k + 1] || icols[k] < 0) return fail(ctx, MNW_ERR_ARG, "text: column numbers must ascend");
    for (int k = 0; k + 1 < n_fcols; k++) if (fcols[k] >= fcols[k + 1] || fcols[k] < 0) return fail(ctx, MNW_ERR_ARG, "text: column numbers must ascend");
    ctx->txt_rows = -1; ctx->txt_nfb = 0; ctx->txt_ni = n_icols; ctx->txt_nf = n_fcols;
    const int64_t ntiles = (int64_t)text_tiles(len);
    CU(ctx->in.reserve((size_t)len + 64));
    CU(ctx->aux.reserve(8 * (2 * (size_t)ntiles + 8) + scan_scratch_bytes(ntiles) + 64));
    { const int rcf = reset_flags(ctx); if (rcf) return rcf; }
    unsigned char *d_buf = ctx->in.as<unsigned char>();
    if (len > 0) CU(cudaMemcpyAsync(d_buf, buf, (size_t)len, cudaMemcpyHostToDevice, ctx->L.stream));
    int64_t *tile_nl = ctx->aux.as<int64_t>(), *tile_off = tile_nl + ntiles, *d_tot = tile_off + ntiles;
    CU(cudaMemsetAsync(d_tot, 0, 64, ctx->L.stream));
    launch_text_count(ctx->L, d_buf, len, tile_nl);
    if (ntiles > 0) {
        const cudaError_t e = launch_scan_sizes(ctx->L, tile_nl, ntiles, 0, tile_off, d_tot, d_tot + 8);
        if (e != cudaSuccess) return fail(ctx, MNW_ERR_CUDA, "text scan: %s", cudaGetErrorString(e));
    }
    int64_t nl = 0;
    CU(cudaMemcpyAsync(&nl, d_tot, 8, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    const int64_t nlines = nl + 1;   // split(): n separators -> n + 1 lines (go/text/parse.go:23-35)
    // per line: start, end, keep, row; then the totals and the column count
    CU(ctx->txt_work.reserve(8 * (4 * (size_t)nlines + 16) + scan_scratch_bytes(nlines) + 64));
    int64_t *line_start = ctx->txt_work.as<int64_t>(), *line_end = line_start + nlines + 1, *keep = line_end + nlines, *row_of = keep + nlines;
    int64_t *d_rows = row_of + nlines;
    int *d_ncols = (int *)(d_rows + 1), *d_fbcount = d_ncols + 1;
    long long *d_errline = (long long *)(d_rows + 2);
    void *scan2 = d_rows + 4;
    CU(cudaMemsetAsync(d_rows, 0, 32, ctx->L.stream));
    launch_text_starts(ctx->L, d_buf, len, tile_off, line_start);
    launch_text_lines(ctx->L, d_buf, len, nlines, line_start, (unsigned char)sep, (unsigned char)comment, line_end, keep, d_ncols);
    {
        const cudaError_t e = launch_scan_sizes(ctx->L, keep, nlines, 0, row_of, d_rows, scan2);
        if (e != cudaSuccess) return fail(ctx, MNW_ERR_CUDA, "text scan: %s", cudaGetErrorString(e));
    }
    int64_t rows = 0;
    CU(cudaMemcpyAsync(&rows, d_rows, 8, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    const int fb_cap = 1 << 16;
    CU(ctx->txt_i.reserve(8 * (size_t)n_icols * (size_t)rows + 64));
    CU(ctx->txt_f.reserve(4 * (size_t)n_fcols * (size_t)rows + 64));
    CU(ctx->txt_fb.reserve(24 * (size_t)fb_cap));
    if (rows > 0 && n_icols + n_fcols > 0)
        launch_text_parse(ctx->L, d_buf, nlines, line_start, line_end, keep, row_of, (unsigned char)sep, n_icols, icols, n_fcols, fcols,
                          d_ncols, rows, ctx->txt_i.as<int64_t>(), ctx->txt_f.as<float>(), ctx->txt_fb.as<int64_t>(), fb_cap, d_fbcount,
                          ctx->flags.as<int>() + FLAG_ERR, d_errline);
    CU(cudaGetLastError());
    int nfb = 0;
    long long errline = 0;
    CU(cudaMemcpyAsync(&nfb, d_fbcount, 4, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaMemcpyAsync(&errline, d_errline, 8, cudaMemcpyDeviceToHost, ctx->L.stream));
    int rc = check_flags(ctx);   // synchronises
    if (rc) {
        if (ctx->err.find("text:") != std::string::npos && rc == MNW_ERR_FORMAT) ctx->err += " (line " + std::to_string(errline + 1) + " of the block)";
        return rc;
    }
    if (nfb > fb_cap) return fail(ctx, MNW_ERR_CAPACITY, "text: more than %d fields need the host's parser", fb_cap);
    ctx->txt_rows = rows; ctx->txt_nfb = nfb;
    if (nrows) *nrows = rows;
    if (nfallback) *nfallback = nfb;
    return MNW_OK;
}

int mnw_text_columns(mnw_ctx *ctx, int64_t *iout, float *fout, int64_t *fallback) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);
    if (ctx->txt_rows < 0) return fail(ctx, MNW_ERR_ARG, "no mnw_text_parse_block call on this context");
    const size_t r = (size_t)ctx->txt_rows;
    if (iout && ctx->txt_ni && r) CU(cudaMemcpyAsync(iout, ctx->txt_i.p, 8 * r * (size_t)ctx->txt_ni, cudaMemcpyDeviceToHost, ctx->L.stream));
    if (fout && ctx->txt_nf && r) CU(cudaMemcpyAsync(fout, ctx->txt_f.p, 4 * r * (size_t)ctx->txt_nf, cudaMemcpyDeviceToHost, ctx->L.stream));
    if (fallback && ctx->txt_nfb) CU(cudaMemcpyAsync(fallback, ctx->txt_fb.p, 24 * (size_t)ctx->txt_nfb, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    return MNW_OK;
}

int mnw_text_columns_dev(mnw_ctx *ctx, const int64_t **icols_dev, const float **fcols_dev) {
    if (!ctx || ctx->txt_rows < 0) return ctx ? fail(ctx, MNW_ERR_ARG, "no mnw_text_parse_block call on this context") : MNW_ERR_ARG;
    if (icols_dev) *icols_dev = ctx->txt_i.as<int64_t>();
    if (fcols_dev) *fcols_dev = ctx->txt_f.as<float>();
    return MNW_OK;
}

// ---- Lagrangian re-gridding (go/minp/snapshot/grid.go) -----------------------------------------------------------------
int mnw_regrid_insert_dev(mnw_ctx *ctx, const int64_t *ids, const float *vec, int64_t n, int64_t ncell, int64_t nside, float *grid) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);
    if (n < 0 || ncell < 1 || nside < 1 || ncell * nside > 2097151) return fail(ctx, MNW_ERR_ARG, "regrid: n = %lld, ncell = %lld, nside = %lld", (long long)n, (long long)ncell, (long long)nside);
    CU(ctx->flags.reserve(64));
    if (!ctx->flags_init) { const int rcf = reset_flags(ctx); if (rcf) return rcf; }
    launch_regrid_insert(ctx->L, ids, vec, n, ncell, nside, grid, ctx->flags.as<int>() + FLAG_ERR);
    CU(cudaGetLastError());
    return MNW_OK;
}

int mnw_regrid_insert(mnw_ctx *ctx, const int64_t *ids, const float *vec, int64_t n, int64_t ncell, int64_t nside, float *grid_dev) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);
    if (n < 0) return fail(ctx, MNW_ERR_ARG, "negative length");
    CU(ctx->in.reserve(20 * (size_t)n + 64));
    int64_t *d_ids = ctx->in.as<int64_t>();
    float *d_vec = (float *)(d_ids + n);
    if (n > 0) {
        CU(cudaMemcpyAsync(d_ids, ids, 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->L.stream));
        CU(cudaMemcpyAsync(d_vec, vec, 12 * (size_t)n, cudaMemcpyHostToDevice, ctx->L.stream));
    }
    int rc = mnw_regrid_insert_dev(ctx, d_ids, d_vec, n, ncell, nside, grid_dev);
    if (rc) return rc;
    return check_flags(ctx);   // synchronises; an invalid ID is an argument error, as in the reference
}

// ---- minh BoundaryWriter (go/minh/boundary.go) ------------------------------------------------------------------------
int mnw_boundary_coordinates(mnw_ctx *ctx, const float *x, const float *y, const float *z, int64_t n, float L, float boundary,
                             int64_t cells, int64_t *sizes, int64_t *total) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);
    if (n < 0 || cells < 1 || cells > 1625) return fail(ctx, MNW_ERR_ARG, "mnw_boundary_coordinates: n = %lld, cells = %lld", (long long)n, (long long)cells);
    const int64_t c3 = cells * cells * cells;
    ctx->bnd_n = -1; ctx->bnd_m = 0;
    CU(ctx->in.reserve(12 * (size_t)n + 64));
    CU(ctx->bnd_work.reserve(8 * (2 * (size_t)n + (size_t)c3 + 2) + scan_scratch_bytes(n) + 64));
    float *dx = ctx->in.as<float>(), *dy = dx + n, *dz = dy + n;
    int64_t *d_cnt = ctx->bnd_work.as<int64_t>(), *d_eoff = d_cnt + n, *d_sizes = d_eoff + n, *d_total = d_sizes + c3;
    void *scan_scratch = d_total + 2;
    { const int rcf = reset_flags(ctx); if (rcf) return rcf; }
    if (n > 0) {
        CU(cudaMemcpyAsync(dx, x, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->L.stream));
        CU(cudaMemcpyAsync(dy, y, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->L.stream));
        CU(cudaMemcpyAsync(dz, z, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->L.stream));
    }
    launch_bnd_count(ctx->L, dx, dy, dz, n, L, boundary, cells, d_cnt, d_sizes, ctx->flags.as<int>() + FLAG_ERR);
    CU(cudaMemsetAsync(d_total, 0, 8, ctx->L.stream));
    if (n > 0) {
        const cudaError_t e = launch_scan_sizes(ctx->L, d_cnt, n, 0, d_eoff, d_total, scan_scratch);
        if (e != cudaSuccess) return fail(ctx, MNW_ERR_CUDA, "boundary scan: %s", cudaGetErrorString(e));
    }
    std::vector<int64_t> h((size_t)c3 + 1);
    CU(cudaMemcpyAsync(h.data(), d_sizes, 8 * ((size_t)c3 + 1), cudaMemcpyDeviceToHost, ctx->L.stream));   // sizes, then the total
    int rc = check_flags(ctx);   // synchronises; a coordinate outside the grid is an argument error (the reference panics)
    if (rc) return rc;
    const int64_t m = h[(size_t)c3];
    ctx->bnd_starts.assign((size_t)c3 + 1, 0);
    for (int64_t g = 0; g < c3; g++) ctx->bnd_starts[(size_t)g + 1] = ctx->bnd_starts[(size_t)g] + h[(size_t)g];
    if (ctx->bnd_starts[(size_t)c3] != m) return fail(ctx, MNW_ERR_CUDA, "boundary: cell sizes and entry count disagree");
    if (sizes) memcpy(sizes, h.data(), 8 * (size_t)c3);
    if (total) *total = m;
    CU(ctx->bnd_idx.reserve(8 * (size_t)m + 64));
    CU(ctx->bnd_flags.reserve(8 * (size_t)m + 64));
    CU(ctx->out.reserve(2 * 12 * (size_t)m + bnd_sort_scratch_bytes(m) + 256));   // sort buffers: released to the next encode
    uint8_t *w = ctx->out.as<uint8_t>();
    int64_t *vals = (int64_t *)w, *vals2 = vals + m;
    uint32_t *keys = (uint32_t *)(vals2 + m), *keys2 = keys + m;
    void *scratch = (void *)(((uintptr_t)(keys2 + m) + 15) & ~(uintptr_t)15);
    const cudaError_t e = launch_bnd_index(ctx->L, dx, dy, dz, n, L, boundary, cells, d_eoff, m, keys, vals, keys2, vals2, scratch,
                                           ctx->bnd_idx.as<int64_t>(), ctx->bnd_flags.as<int64_t>());
    if (e != cudaSuccess) return fail(ctx, MNW_ERR_CUDA, "boundary index: %s", cudaGetErrorString(e));
    CU(cudaStreamSynchronize(ctx->L.stream));   // (ctx->out is about to be reused by the column encoders)
    ctx->bnd_n = n; ctx->bnd_m = m;
    return MNW_OK;
}

int mnw_boundary_index(mnw_ctx *ctx, int64_t *idx, int64_t *flags) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);
    if (ctx->bnd_n < 0) return fail(ctx, MNW_ERR_ARG, "no mnw_boundary_coordinates call on this context");
    if (ctx->bnd_m > 0 && idx) CU(cudaMemcpyAsync(idx, ctx->bnd_idx.p, 8 * (size_t)ctx->bnd_m, cudaMemcpyDeviceToHost, ctx->L.stream));
    if (ctx->bnd_m > 0 && flags) CU(cudaMemcpyAsync(flags, ctx->bnd_flags.p, 8 * (size_t)ctx->bnd_m, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    return MNW_OK;
}

static int boundary_column(mnw_ctx *ctx, int kind, const mnw_float_desc *desc, const void *col, int64_t ncol, const void *x_dev,
                           int64_t *mins, int64_t *bits, int64_t *offsets, uint8_t *out, int64_t out_cap, int64_t *out_len) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);
    if (ctx->bnd_n < 0) return fail(ctx, MNW_ERR_ARG, "no mnw_boundary_coordinates call on this context");
    if (!x_dev && ncol != ctx->bnd_n) return fail(ctx, MNW_ERR_ARG, "column of %lld elements, coordinates of %lld", (long long)ncol, (long long)ctx->bnd_n);
    const int64_t nblocks = (int64_t)ctx->bnd_starts.size() - 1;
    if (x_dev)   // the boundary flags themselves: contiguous ragged blocks, already on the device
        return encode_group_host(ctx, kind, desc, nullptr, 0, nblocks, ctx->bnd_starts.data(), mins, bits, offsets, out, out_cap, out_len,
                                 nullptr, 0, nullptr, x_dev);
    return encode_group_host(ctx, kind, desc, col, 0, nblocks, ctx->bnd_starts.data(), mins, bits, offsets, out, out_cap, out_len, nullptr,
                             ncol, ctx->bnd_idx.as<int64_t>());
}

int mnw_boundary_encode_int_column(mnw_ctx *ctx, const int64_t *col, int64_t ncol, int64_t *mins, int64_t *bits, int64_t *offsets,
                                   uint8_t *out, int64_t out_cap, int64_t *out_len) {
    return boundary_column(ctx, KIND_I64, nullptr, col, ncol, nullptr, mins, bits, offsets, out, out_cap, out_len);
}

int mnw_boundary_encode_float_column(mnw_ctx *ctx, const mnw_float_desc *desc, const float *col, int64_t ncol, int64_t *mins,
                                     int64_t *bits, int64_t *offsets, uint8_t *out, int64_t out_cap, int64_t *out_len) {
    int rc = check_desc(ctx, desc);
    if (rc) return rc;
    return boundary_column(ctx, KIND_F32, desc, col, ncol, nullptr, mins, bits, offsets, out, out_cap, out_len);
}

int mnw_boundary_encode_flags(mnw_ctx *ctx, int64_t *mins, int64_t *bits, int64_t *offsets, uint8_t *out, int64_t out_cap,
                              int64_t *out_len) {
    return boundary_column(ctx, KIND_I64, nullptr, nullptr, 0, ctx ? ctx->bnd_flags.p : nullptr, mins, bits, offsets, out, out_cap, out_len);
}

int mnw_decode_int_blocks(mnw_ctx *ctx, const uint8_t *data, int64_t data_len, const int64_t *offsets,
                          const int64_t *mins, const int64_t *bits, int64_t n, int64_t nsel, const int64_t *sel,
                          int64_t *out) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    return decode_blocks_host(ctx, 0, nullptr, data, data_len, offsets, mins, bits, n, nsel, sel, nullptr, out);
}

int mnw_decode_float_blocks(mnw_ctx *ctx, const mnw_float_desc *desc, const uint8_t *data, int64_t data_len,
                            const int64_t *offsets, const int64_t *mins, const int64_t *bits, int64_t n,
                            int64_t nsel, const int64_t *sel, const mnw_jitter *jitter, float *out) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    return decode_blocks_host(ctx, 1, desc, data, data_len, offsets, mins, bits, n, nsel, sel, jitter, out);
}

int mnw_scan_offsets(mnw_ctx *ctx, const int64_t *nbytes, int64_t nblocks, int64_t base, int64_t *offsets,
                     int64_t *total) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    if (nblocks < 0) return fail(ctx, MNW_ERR_ARG, "negative block count");
    cudaError_t e = cudaSuccess;
    CU(ctx->meta.reserve(8 * (size_t)(2 * nblocks + 2)));
    int64_t *d_sizes = ctx->meta.as<int64_t>(), *d_off = d_sizes + nblocks, *d_total = d_off + nblocks;
    if (nblocks) CU(cudaMemcpyAsync(d_sizes, nbytes, 8 * (size_t)nblocks, cudaMemcpyHostToDevice, ctx->L.stream));
    CU(ctx->aux.reserve(scan_scratch_bytes(nblocks)));
    e = launch_scan_sizes(ctx->L, d_sizes, nblocks, base, d_off, d_total, ctx->aux.p);
    if (e != cudaSuccess) return fail(ctx, MNW_ERR_CUDA, "scan: %s", cudaGetErrorString(e));
    if (nblocks) CU(cudaMemcpyAsync(offsets, d_off, 8 * (size_t)nblocks, cudaMemcpyDeviceToHost, ctx->L.stream));
    int64_t t = 0;
    CU(cudaMemcpyAsync(&t, d_total, 8, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    if (total) *total = t;
    return MNW_OK;
}

// ---- device-resident variants ------------------------------------------------
int mnw_encode_int_group_dev(mnw_ctx *ctx, const int64_t *x, int64_t n, int64_t nblocks, int64_t *mins,
                             int64_t *bits, int64_t *offsets, uint8_t *out, int64_t out_cap, int64_t *out_len) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    return encode_group_dev(ctx, KIND_I64, nullptr, x, n, nblocks, nullptr, nullptr, nullptr, 0, 0, mins, bits,
                            offsets, out, out_cap, out_len);
}

int mnw_encode_float_group_dev(mnw_ctx *ctx, const mnw_float_desc *desc, const float *x, int64_t n,
                               int64_t nblocks, int64_t *mins, int64_t *bits, int64_t *offsets, uint8_t *out,
                               int64_t out_cap, int64_t *out_len) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    return encode_group_dev(ctx, KIND_F32, desc, x, n, nblocks, nullptr, nullptr, nullptr, 0, 0, mins, bits,
                            offsets, out, out_cap, out_len);
}

int mnw_decode_int_blocks_dev(mnw_ctx *ctx, const uint8_t *data, int64_t data_len, const int64_t *offsets,
                              const int64_t *mins, const int64_t *bits, int64_t n, int64_t nsel,
                              const int64_t *sel, int64_t *out) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    DecodeHost h;
    h.mode = 0; h.data = data; h.stream_len = data_len; h.offsets = offsets; h.mins = mins; h.bits = bits;
    h.sel = sel; h.n = n; h.nsel = nsel; h.out = out;
    launch_decode(ctx->L, h);
    CU(cudaGetLastError());
    return MNW_OK;
}

int mnw_decode_float_blocks_dev(mnw_ctx *ctx, const mnw_float_desc *desc, const uint8_t *data, int64_t data_len,
                                const int64_t *offsets, const int64_t *mins, const int64_t *bits, int64_t n,
                                int64_t nsel, const int64_t *sel, const mnw_jitter *jitter, float *out) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    DecodeHost h;
    h.mode = 1;
    int rc = fill_decode_float(ctx, h, desc, 1, jitter);
    if (rc) return rc;
    if (h.jmode == 2) h.u = jitter->u_stream;  // device pointer here
    h.data = data; h.stream_len = data_len; h.offsets = offsets; h.mins = mins; h.bits = bits;
    h.sel = sel; h.n = n; h.nsel = nsel; h.out = out;
    launch_decode(ctx->L, h);
    CU(cudaGetLastError());
    return MNW_OK;
}

// Every IntGroup / FloatGroup column of one minh block in at most three launches (plain float, Log, int) (minh.Reader.Block's per-column loop,
// go/minh/minh.go:296-323): column c is block 0 of its own group, packed at data + offsets[c].
int mnw_decode_columns_dev(mnw_ctx *ctx, int64_t ncols, const mnw_column *cols, const uint8_t *data, int64_t data_len,
                           const int64_t *offsets, const int64_t *mins, const int64_t *bits, int64_t n, const mnw_jitter *jitter,
                           void *const *out_dev) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);
    if (ncols < 0 || n < 0 || data_len < 0) return fail(ctx, MNW_ERR_ARG, "negative length");
    if (ncols == 0 || n == 0) return MNW_OK;
    if (!cols || !out_dev || !data || !offsets || !mins || !bits) return fail(ctx, MNW_ERR_ARG, "null argument");
    if (jitter && (jitter->mode < 0 || jitter->mode > 1))
        return fail(ctx, MNW_ERR_ARG, "mnw_decode_columns_dev takes MNW_JITTER_CENTER or MNW_JITTER_HASH");
    std::vector<FloatParams> tab((size_t)ncols);
    std::vector<int64_t> sel_f, sel_l, sel_i;   // plain float columns, Log columns (their kernel carries 10^x), int columns
    std::vector<void *> out_f, out_l, out_i;
    for (int64_t c = 0; c < ncols; c++) {
        if (!out_dev[c]) return fail(ctx, MNW_ERR_ARG, "column %lld has no output", (long long)c);
        if (cols[c].is_float) {
            int rc = check_desc(ctx, &cols[c].desc);
            if (rc) return rc;
            tab[(size_t)c] = to_params(cols[c].desc);
            if (cols[c].desc.log10) { sel_l.push_back(c); out_l.push_back(out_dev[c]); }
            else { sel_f.push_back(c); out_f.push_back(out_dev[c]); }
        } else {
            tab[(size_t)c] = FloatParams{};
            sel_i.push_back(c); out_i.push_back(out_dev[c]);
        }
    }
    // one upload: [tab | sel_f | sel_l | sel_i | out_f | out_l | out_i]
    const size_t tab_b = sizeof(FloatParams) * (size_t)ncols, o_sf = (tab_b + 15) & ~(size_t)15, o_sl = o_sf + 8 * sel_f.size(),
                 o_si = o_sl + 8 * sel_l.size(), o_of = o_si + 8 * sel_i.size(), o_ol = o_of + 8 * out_f.size(),
                 o_oi = o_ol + 8 * out_l.size(), total = o_oi + 8 * out_i.size();
    std::vector<unsigned char> blob(total);
    memcpy(blob.data(), tab.data(), tab_b);
    if (!sel_f.empty()) { memcpy(blob.data() + o_sf, sel_f.data(), 8 * sel_f.size()); memcpy(blob.data() + o_of, out_f.data(), 8 * out_f.size()); }
    if (!sel_l.empty()) { memcpy(blob.data() + o_sl, sel_l.data(), 8 * sel_l.size()); memcpy(blob.data() + o_ol, out_l.data(), 8 * out_l.size()); }
    if (!sel_i.empty()) { memcpy(blob.data() + o_si, sel_i.data(), 8 * sel_i.size()); memcpy(blob.data() + o_oi, out_i.data(), 8 * out_i.size()); }
    CU(ctx->dec_cols.reserve(total + 16));
    CU(cudaMemcpyAsync(ctx->dec_cols.p, blob.data(), total, cudaMemcpyHostToDevice, ctx->L.stream));   // pageable: staged before return
    unsigned char *d = (unsigned char *)ctx->dec_cols.p;
    DecodeHost h;
    h.data = data; h.stream_len = data_len; h.offsets = offsets; h.mins = mins; h.bits = bits; h.n = n;
    h.tab = (const FloatParams *)d; h.tab_per_file = 1;
    if (jitter) { h.jmode = jitter->mode; h.seed = jitter->seed; h.block_id0 = jitter->block_id0; }
    if (!sel_f.empty()) {
        h.mode = 1; h.any_log = false;
        h.sel = (const int64_t *)(d + o_sf); h.nsel = (int64_t)sel_f.size(); h.outs = (void *const *)(d + o_of);
        launch_decode(ctx->L, h);
    }
    if (!sel_l.empty()) {
        h.mode = 1; h.any_log = true;
        h.sel = (const int64_t *)(d + o_sl); h.nsel = (int64_t)sel_l.size(); h.outs = (void *const *)(d + o_ol);
        launch_decode(ctx->L, h);
    }
    if (!sel_i.empty()) {
        h.mode = 0; h.sel = (const int64_t *)(d + o_si); h.nsel = (int64_t)sel_i.size(); h.outs = (void *const *)(d + o_oi);
        launch_decode(ctx->L, h);
    }
    CU(cudaGetLastError());
    return MNW_OK;
}

// The minp encode of a batch of files whose group parameters are in the device table `tab`.  fp (host copy of the
// table) is null when the parameters were derived on the device (k_vec3_params): the fused kernels are then launched on
// the strength of the shape alone and return at once if the device raised the skip flag; the generic kernels, gated on
// the abort flag, take over.  The flag words have been reset by the caller.
static int encode_vec3_core(mnw_ctx *ctx, const FloatParams *tab, const std::vector<FloatParams> *fp, int desc_per_file,
                            const float *aos, int64_t nfile, int64_t subcells, int64_t nfiles, int64_t *mins, int64_t *bits,
                            int64_t *offsets, uint8_t *out, int64_t out_axis_stride, int64_t *out_len) {
    const int64_t ndesc = desc_per_file ? 3 * nfiles : 3;
    const int64_t sc3 = subcells * subcells * subcells, nsub = nfile / subcells, n = nsub * nsub * nsub;
    const int64_t nb = nfiles * 3 * sc3;
    int *d_flags = ctx->flags.as<int>();
    if (nb == 0) return MNW_OK;

    BatchShape sh = {};
    sh.nblocks = nb; sh.nchains = 3 * nfiles; sh.blocks_per_chain = sc3; sh.uniform_n = n;
    sh.total_tiles = nb * ((n + PACK_TILE - 1) / PACK_TILE);
    sh.total_chunks = nb * ((n + STATS_CHUNK - 1) / STATS_CHUNK);
    if (sh.total_tiles >= (1LL << 31)) return fail(ctx, MNW_ERR_ARG, "batch too large for one launch");
    launch_build_vec3(ctx->L, ctx->descs.as<BlockDesc>(), nfiles, aos, (int32_t)nfile, (int32_t)subcells, tab, desc_per_file);

    bool fused_ok, pipe_ok;
    if (fp) {
        fused_ok = fused_vec3_supported(fp->data(), ndesc, (int)nfile, (int)subcells, aos);
        pipe_ok = pipe_vec3_supported(fp->data(), ndesc);
    } else {   // shape and alignment only; k_vec3_params vouches for the parameters (or raises the skip flag)
        fused_ok = (nsub == 16 || nsub == 32 || nsub == 64 || nsub == 128) && ((uintptr_t)aos & 15) == 0 && nfile <= 1024;
        pipe_ok = true;
    }
    if (!ctx->force_generic && fused_ok) {
        ctx->last_path = 1;
        CU(ctx->fused_ws.reserve(fused_work_bytes(nb)));
        FusedWork W = {};
        W.pub = ctx->fused_ws.as<unsigned long long>();
        W.repack_list = (int64_t *)(W.pub + nb);
        W.err = d_flags + FLAG_ERR; W.abort_flag = d_flags + 2; W.repack_count = d_flags + 3; W.ticket = (unsigned int *)(d_flags + 4);
        W.skip = fp ? nullptr : d_flags + 5;
        CU(cudaMemsetAsync(W.pub, 0, 8 * (size_t)nb, ctx->L.stream));
        // The cooperative schedule owns the whole GPU while it runs: right for large batches, but calls on a
        // file or two (the host-pointer entry points, several contexts at a time) overlap better as clusters.
        void *coop_ws = nullptr;
        const char *cmin = getenv("MNW_PIPE_COOP_MIN");   // tuning / test knob: smallest batch (in units) that goes cooperative
        if (nfiles * sc3 >= (cmin ? atoll(cmin) : 256) || nsub == 32 || nsub == 128) {   // 32^3 and 128^3: k_pipe_vec3 has no cluster schedule
            CU(ctx->coop_ws.reserve(pipe_coop_ws_bytes(nfiles * sc3, (int)nsub)));
            coop_ws = ctx->coop_ws.p;
        }
        bool ran_pipe = false;
        const cudaError_t e = launch_fused_vec3(ctx->L, W, ctx->descs.as<BlockDesc>(), tab, desc_per_file, aos, (int)nfile, (int)subcells, nfiles,
                                                ctx->stats.as<BlockStat>(), mins, bits, offsets, out_len, out, out_axis_stride,
                                                pipe_ok, coop_ws, &ran_pipe);
        if (e != cudaSuccess && e != cudaErrorNotSupported) return fail(ctx, MNW_ERR_CUDA, "fused vec3 encode: %s", cudaGetErrorString(e));
        if (e == cudaErrorNotSupported) {   // no fused kernel for this shape on this device after all: the generic kernels
            ctx->last_path = 0;
            launch_generic_encode(ctx->L, ctx->descs.as<BlockDesc>(), ctx->stats.as<BlockStat>(), sh,
                                  ctx->slow.as<int64_t>(), d_flags, d_flags + FLAG_ERR, mins, bits, offsets, out_len, out,
                                  out_axis_stride, out_axis_stride);
            CU(cudaGetLastError());
            return MNW_OK;
        }
        // blocks wider than 16 bits: packed from global memory with the fused kernel's (min, bits, offset)
        launch_pack_list(ctx->L, ctx->descs.as<BlockDesc>(), ctx->stats.as<BlockStat>(), sh, W.repack_list, W.repack_count,
                         out, out_axis_stride, out_axis_stride, d_flags + FLAG_ERR);
        // The generic kernels, gated on the abort flag, redo the whole call when (a) k_fused_vec3 met a block that needs
        // the exact sequential periodicMin, or (b) k_vec3_params found derived parameters outside the fused kernels'
        // domain.  k_pipe_vec3 with parameters the host has checked raises neither: it resolves such blocks itself.
        if (!(fp && ran_pipe))
            launch_generic_encode(ctx->L, ctx->descs.as<BlockDesc>(), ctx->stats.as<BlockStat>(), sh,
                                  ctx->slow.as<int64_t>(), d_flags, d_flags + FLAG_ERR, mins, bits, offsets, out_len, out,
                                  out_axis_stride, out_axis_stride, W.abort_flag);
        CU(cudaGetLastError());
        return MNW_OK;
    }
    ctx->last_path = 0;
    launch_generic_encode(ctx->L, ctx->descs.as<BlockDesc>(), ctx->stats.as<BlockStat>(), sh,
                          ctx->slow.as<int64_t>(), d_flags, d_flags + FLAG_ERR, mins, bits, offsets, out_len, out,
                          out_axis_stride, out_axis_stride);
    CU(cudaGetLastError());
    return MNW_OK;
}

int mnw_encode_vec3_subcells_dev(mnw_ctx *ctx, const mnw_float_desc *desc, int desc_per_file, const float *aos,
                                 int64_t nfile, int64_t subcells, int64_t nfiles, int64_t *mins, int64_t *bits,
                                 int64_t *offsets, uint8_t *out, int64_t out_axis_stride, int64_t *out_len) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    if (nfile <= 0 || subcells <= 0 || nfile % subcells != 0 || nfiles < 0 || nfile > 2048)
        return fail(ctx, MNW_ERR_ARG, "vec3: nfile = %lld, subcells = %lld", (long long)nfile, (long long)subcells);
    int rc0 = check_desc(ctx, desc);
    if (rc0) return rc0;
    const int64_t ndesc = desc_per_file ? 3 * nfiles : 3;
    std::vector<FloatParams> fp;
    const FloatParams *tab = nullptr;
    rc0 = upload_params(ctx, desc, ndesc, fp, &tab);
    if (rc0) return rc0;
    const int64_t sc3 = subcells * subcells * subcells;
    int rc = reserve_batch(ctx, nfiles * 3 * sc3, 3 * nfiles);
    if (rc) return rc;
    { const int rcf = reset_flags(ctx); if (rcf) return rcf; }
    return encode_vec3_core(ctx, tab, &fp, desc_per_file, aos, nfile, subcells, nfiles, mins, bits, offsets, out, out_axis_stride, out_len);
}

int mnw_minp_encode_vectors_dev(mnw_ctx *ctx, const float *aos, int64_t nfile, int64_t subcells, int64_t nfiles, int periodic,
                                float L, float dx, mnw_float_desc *desc_dev, int64_t *mins, int64_t *bits, int64_t *offsets,
                                uint8_t *out, int64_t out_axis_stride, int64_t *out_len) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);
    if (nfile <= 0 || subcells <= 0 || nfile % subcells != 0 || nfiles < 0 || nfile > 2048)
        return fail(ctx, MNW_ERR_ARG, "vec3: nfile = %lld, subcells = %lld", (long long)nfile, (long long)subcells);
    const int64_t sc3 = subcells * subcells * subcells, np = nfile * nfile * nfile, nsub = nfile / subcells;
    int rc = reserve_batch(ctx, nfiles * 3 * sc3, 3 * nfiles);
    if (rc) return rc;
    { const int rcf = reset_flags(ctx); if (rcf) return rcf; }
    if (nfiles == 0) return MNW_OK;
    if (periodic) {   // [0, L) on every axis (go/minp/minp.go:88-90): the parameters are known on the host
        mnw_float_desc d[3];
        for (int k = 0; k < 3; k++) {
            memset(&d[k], 0, sizeof d[k]);
            d[k].low = 0.0f; d[k].high = L; d[k].pixels = mnw_float_group_pixels(0.0f, L, dx); d[k].periodic = 1;
        }
        std::vector<FloatParams> fp;
        const FloatParams *tab = nullptr;
        rc = upload_params(ctx, d, 3, fp, &tab);
        if (rc) return rc;
        if (desc_dev) {
            std::vector<mnw_float_desc> all((size_t)(3 * nfiles));
            for (size_t i = 0; i < all.size(); i++) all[i] = d[i % 3];
            CU(cudaMemcpyAsync(desc_dev, all.data(), all.size() * sizeof(mnw_float_desc), cudaMemcpyHostToDevice, ctx->L.stream));   // pageable: staged before return
        }
        return encode_vec3_core(ctx, tab, &fp, 0, aos, nfile, subcells, nfiles, mins, bits, offsets, out, out_axis_stride, out_len);
    }
    // non-periodic (go/minp/minp.go:92-95): bounds() per file, then the three FloatGroups' parameters, all on the device
    if ((np + 32767) / 32768 >= 65536LL * 32768 || nfiles > 65535) return fail(ctx, MNW_ERR_ARG, "too many files in one call");
    CU(ctx->aux.reserve(24 * (size_t)nfiles + 64));
    CU(ctx->params.reserve(sizeof(FloatParams) * 3 * (size_t)nfiles));
    launch_vec3_limits(ctx->L, aos, np, nfiles, ctx->aux.as<uint32_t>());
    int *d_flags = ctx->flags.as<int>();
    launch_vec3_params(ctx->L, ctx->aux.as<uint32_t>(), nfiles, dx, ctx->params.as<FloatParams>(), desc_dev, d_flags + 5, d_flags + 2,
                       nsub == 16 ? 0 : 1);
    CU(cudaGetLastError());
    return encode_vec3_core(ctx, ctx->params.as<FloatParams>(), nullptr, 1, aos, nfile, subcells, nfiles, mins, bits, offsets, out,
                            out_axis_stride, out_len);
}

int mnw_minp_decode_vectors_dev(mnw_ctx *ctx, const mnw_float_desc *desc_dev, const uint8_t *data, int64_t data_axis_stride,
                                const int64_t *offsets, const int64_t *mins, const int64_t *bits, int64_t nfile, int64_t subcells,
                                int64_t nfiles, int periodic, float L, const mnw_jitter *jitter, float *aos_out) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);
    if (nfile <= 0 || subcells <= 0 || nfile % subcells != 0 || nfiles < 0 || !desc_dev)
        return fail(ctx, MNW_ERR_ARG, "vec3: nfile = %lld, subcells = %lld", (long long)nfile, (long long)subcells);
    DecodeHost h;
    h.mode = 2;
    h.tab_per_file = 1;
    CU(ctx->params.reserve(sizeof(FloatParams) * 3 * (size_t)nfiles));
    launch_params_from_desc(ctx->L, desc_dev, 3 * nfiles, ctx->params.as<FloatParams>());
    h.tab = ctx->params.as<FloatParams>();
    h.low_nonneg = periodic != 0 && L > 0.0f;   // a periodic field's groups are [0, L): every decoded value is >= +0
    if (jitter) {
        if (jitter->mode < 0 || jitter->mode > 2) return fail(ctx, MNW_ERR_ARG, "unknown jitter mode %d", jitter->mode);
        h.jmode = jitter->mode; h.seed = jitter->seed; h.block_id0 = jitter->block_id0;
        if (h.jmode == 2) {
            if (!jitter->u_stream) return fail(ctx, MNW_ERR_ARG, "jitter mode STREAM without u_stream");
            h.u = jitter->u_stream;
        }
    }
    const int64_t sc3 = subcells * subcells * subcells, nsub = nfile / subcells;
    h.data = data; h.stream_len = data_axis_stride; h.offsets = offsets; h.mins = mins; h.bits = bits;
    h.n = nsub * nsub * nsub; h.nsel = nfiles * 3 * sc3; h.wrap_L = periodic ? L : 0.0f;
    h.nfile = (int32_t)nfile; h.subcells = (int32_t)subcells; h.out = aos_out;
    if (!ctx->force_generic && h.jmode != 2 && fused_decode_vec3_supported((int)nfile, (int)subcells, aos_out)) {
        cudaError_t e = launch_fused_decode_vec3(ctx->L, h, nfiles);
        if (e != cudaSuccess) return fail(ctx, MNW_ERR_CUDA, "fused vec3 decode: %s", cudaGetErrorString(e));
        return MNW_OK;
    }
    launch_decode(ctx->L, h);
    CU(cudaGetLastError());
    return MNW_OK;
}

int mnw_decode_vec3_subcells_dev(mnw_ctx *ctx, const mnw_float_desc *desc, int desc_per_file, const uint8_t *data,
                                 int64_t data_axis_stride, const int64_t *offsets, const int64_t *mins,
                                 const int64_t *bits, int64_t nfile, int64_t subcells, int64_t nfiles,
                                 float wrap_L, const mnw_jitter *jitter, float *aos_out) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    if (nfile <= 0 || subcells <= 0 || nfile % subcells != 0 || nfiles < 0)
        return fail(ctx, MNW_ERR_ARG, "vec3: nfile = %lld, subcells = %lld", (long long)nfile, (long long)subcells);
    DecodeHost h;
    h.mode = 2;
    h.tab_per_file = desc_per_file;
    int rc = fill_decode_float(ctx, h, desc, desc_per_file ? 3 * nfiles : 3, jitter);
    if (rc) return rc;
    if (h.jmode == 2) {   // caller's doubles, u_stream[block * n + i] (DEVICE pointer here): the generic kernel reads them
        if (!jitter->u_stream) return fail(ctx, MNW_ERR_ARG, "jitter mode STREAM without u_stream");
        h.u = jitter->u_stream;
    }
    const int64_t sc3 = subcells * subcells * subcells, nsub = nfile / subcells;
    h.data = data; h.stream_len = data_axis_stride; h.offsets = offsets; h.mins = mins; h.bits = bits;
    h.n = nsub * nsub * nsub; h.nsel = nfiles * 3 * sc3; h.wrap_L = wrap_L;
    h.nfile = (int32_t)nfile; h.subcells = (int32_t)subcells; h.out = aos_out;
    if (!ctx->force_generic && h.jmode != 2 && fused_decode_vec3_supported((int)nfile, (int)subcells, aos_out)) {
        cudaError_t e = launch_fused_decode_vec3(ctx->L, h, nfiles);
        if (e != cudaSuccess) return fail(ctx, MNW_ERR_CUDA, "fused vec3 decode: %s", cudaGetErrorString(e));
        return MNW_OK;
    }
    launch_decode(ctx->L, h);
    CU(cudaGetLastError());
    return MNW_OK;
}

// ---- minp host-pointer entry points -------------------------------------------
// Encode the cube already resident in ctx->in and stage the results out to the host.
static int encode_vec3_resident(mnw_ctx *ctx, const mnw_float_desc desc[3], int64_t nfile, int64_t subcells,
                                int64_t *mins, int64_t *bits, int64_t *offsets, uint8_t *out,
                                int64_t out_axis_stride, int64_t out_len[3]) {
    const int64_t np = nfile * nfile * nfile, sc3 = subcells * subcells * subcells, nb = 3 * sc3;
    const int64_t stride = (8 * np + 255) & ~255LL;  // worst case: 64 bits per value
    CU(ctx->out.reserve(3 * (size_t)stride + 64));
    int rc = reserve_batch(ctx, nb, 3);
    if (rc) return rc;
    int64_t *d_meta = ctx->meta.as<int64_t>();
    rc = mnw_encode_vec3_subcells_dev(ctx, desc, 0, ctx->in.as<float>(), nfile, subcells, 1, d_meta, d_meta + nb,
                                      d_meta + 2 * nb, ctx->out.as<uint8_t>(), stride, d_meta + 3 * nb);
    if (rc) return rc;
    std::vector<int64_t> h_meta(3 * (size_t)nb + 3);
    CU(cudaMemcpyAsync(h_meta.data(), d_meta, h_meta.size() * 8, cudaMemcpyDeviceToHost, ctx->L.stream));
    rc = check_flags(ctx);
    if (rc) return rc;
    if (mins) memcpy(mins, h_meta.data(), 8 * (size_t)nb);
    if (bits) memcpy(bits, h_meta.data() + nb, 8 * (size_t)nb);
    if (offsets) memcpy(offsets, h_meta.data() + 2 * nb, 8 * (size_t)nb);
    for (int k = 0; k < 3; k++) {
        int64_t len = h_meta[3 * (size_t)nb + k];
        out_len[k] = len;
        if (len > out_axis_stride)
            return fail(ctx, MNW_ERR_CAPACITY, "axis %d needs %lld bytes, stride is %lld", k, (long long)len, (long long)out_axis_stride);
        if (len > 0)
            CU(cudaMemcpyAsync(out + k * out_axis_stride, ctx->out.as<uint8_t>() + k * stride, (size_t)len,
                               cudaMemcpyDeviceToHost, ctx->L.stream));
    }
    CU(cudaStreamSynchronize(ctx->L.stream));
    return MNW_OK;
}

int mnw_encode_vec3_subcells(mnw_ctx *ctx, const mnw_float_desc desc[3], const float *aos, int64_t nfile,
                             int64_t subcells, int64_t *mins, int64_t *bits, int64_t *offsets, uint8_t *out,
                             int64_t out_axis_stride, int64_t out_len[3]) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    if (nfile <= 0 || subcells <= 0 || nfile % subcells != 0)
        return fail(ctx, MNW_ERR_ARG, "vec3: nfile = %lld, subcells = %lld", (long long)nfile, (long long)subcells);
    const int64_t np = nfile * nfile * nfile;
    CU(ctx->in.reserve(12 * (size_t)np + 16));
    CU(cudaMemcpyAsync(ctx->in.p, aos, 12 * (size_t)np, cudaMemcpyHostToDevice, ctx->L.stream));
    return encode_vec3_resident(ctx, desc, nfile, subcells, mins, bits, offsets, out, out_axis_stride, out_len);
}

int mnw_minp_encode_vectors(mnw_ctx *ctx, const float *aos, int64_t nfile, int64_t subcells, int periodic, float L,
                            float dx, mnw_float_desc desc_out[3], int64_t *mins, int64_t *bits, int64_t *offsets,
                            uint8_t *out, int64_t out_axis_stride, int64_t out_len[3]) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    if (nfile <= 0 || subcells <= 0 || nfile % subcells != 0)
        return fail(ctx, MNW_ERR_ARG, "vec3: nfile = %lld, subcells = %lld", (long long)nfile, (long long)subcells);
    const int64_t np = nfile * nfile * nfile;
    CU(ctx->in.reserve(12 * (size_t)np + 16));
    CU(cudaMemcpyAsync(ctx->in.p, aos, 12 * (size_t)np, cudaMemcpyHostToDevice, ctx->L.stream));   // the one upload
    float lo[3] = {0.0f, 0.0f, 0.0f}, hi[3] = {L, L, L};                                            // go/minp/minp.go:88-90
    if (!periodic) {                                                                                // :92-95
        int rc = mnw_vec3_limits_dev(ctx, ctx->in.as<float>(), np, 1, lo, hi);
        if (rc) return rc;
    }
    mnw_float_desc d[3];
    for (int k = 0; k < 3; k++) {
        memset(&d[k], 0, sizeof d[k]);
        d[k].low = lo[k]; d[k].high = hi[k];
        d[k].pixels = mnw_float_group_pixels(lo[k], hi[k], dx);                                     // go/writer.go:73
        d[k].periodic = 1;                                                                          // go/writer.go:74
        if (desc_out) desc_out[k] = d[k];
    }
    return encode_vec3_resident(ctx, d, nfile, subcells, mins, bits, offsets, out, out_axis_stride, out_len);
}

int mnw_decode_vec3_subcells(mnw_ctx *ctx, const mnw_float_desc desc[3], const uint8_t *const data[3],
                             const int64_t data_len[3], const int64_t *offsets, const int64_t *mins,
                             const int64_t *bits, int64_t nfile, int64_t subcells, float wrap_L,
                             const mnw_jitter *jitter, float *aos_out) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    if (nfile <= 0 || subcells <= 0 || nfile % subcells != 0)
        return fail(ctx, MNW_ERR_ARG, "vec3: nfile = %lld, subcells = %lld", (long long)nfile, (long long)subcells);
    const int64_t np = nfile * nfile * nfile, sc3 = subcells * subcells * subcells, nb = 3 * sc3;
    const int64_t nsub = nfile / subcells, n = nsub * nsub * nsub;
    int64_t stride = 0;
    for (int k = 0; k < 3; k++) stride = data_len[k] > stride ? data_len[k] : stride;
    stride = (stride + 255) & ~255LL;
    for (int64_t b = 0; b < nb; b++) {
        if (bits[b] < 0 || bits[b] > 64) return fail(ctx, MNW_ERR_FORMAT, "block %lld has %lld bits", (long long)b, (long long)bits[b]);
        if (offsets[b] < 0 || offsets[b] + array_bytes(bits[b], n) > data_len[b / sc3])
            return fail(ctx, MNW_ERR_FORMAT, "block %lld lies outside its group's data", (long long)b);
    }
    CU(ctx->in.reserve(3 * (size_t)stride + 16));
    CU(ctx->meta.reserve(8 * 3 * (size_t)nb));
    CU(ctx->dec_out.reserve(12 * (size_t)np));
    for (int k = 0; k < 3; k++)
        if (data_len[k] > 0)
            CU(cudaMemcpyAsync(ctx->in.as<uint8_t>() + k * stride, data[k], (size_t)data_len[k], cudaMemcpyHostToDevice, ctx->L.stream));
    int64_t *d_meta = ctx->meta.as<int64_t>();
    CU(cudaMemcpyAsync(d_meta, offsets, 8 * (size_t)nb, cudaMemcpyHostToDevice, ctx->L.stream));
    CU(cudaMemcpyAsync(d_meta + nb, mins, 8 * (size_t)nb, cudaMemcpyHostToDevice, ctx->L.stream));
    CU(cudaMemcpyAsync(d_meta + 2 * nb, bits, 8 * (size_t)nb, cudaMemcpyHostToDevice, ctx->L.stream));
    mnw_jitter jdev;
    if (jitter && jitter->mode == MNW_JITTER_STREAM) {   // the caller's doubles, one per decoded value: [3 * sc3][n]
        if (!jitter->u_stream) return fail(ctx, MNW_ERR_ARG, "jitter mode STREAM without u_stream");
        CU(ctx->ustream.reserve(8 * 3 * (size_t)np));
        CU(cudaMemcpyAsync(ctx->ustream.p, jitter->u_stream, 8 * 3 * (size_t)np, cudaMemcpyHostToDevice, ctx->L.stream));
        jdev = *jitter;
        jdev.u_stream = ctx->ustream.as<double>();
        jitter = &jdev;
    }
    int rc = mnw_decode_vec3_subcells_dev(ctx, desc, 0, ctx->in.as<uint8_t>(), stride, d_meta, d_meta + nb, d_meta + 2 * nb,
                                          nfile, subcells, 1, wrap_L, jitter, ctx->dec_out.as<float>());
    if (rc) return rc;
    CU(cudaMemcpyAsync(aos_out, ctx->dec_out.p, 12 * (size_t)np, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    return MNW_OK;
}

int mnw_scan_offsets_dev(mnw_ctx *ctx, const int64_t *nbytes, int64_t nblocks, int64_t base, int64_t *offsets,
                         int64_t *total) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    if (nblocks < 0) return fail(ctx, MNW_ERR_ARG, "negative block count");
    CU(ctx->aux.reserve(scan_scratch_bytes(nblocks)));
    cudaError_t e = launch_scan_sizes(ctx->L, nbytes, nblocks, base, offsets, total, ctx->aux.p);
    if (e != cudaSuccess) return fail(ctx, MNW_ERR_CUDA, "scan: %s", cudaGetErrorString(e));
    return MNW_OK;
}

int mnw_selftest_fastdiv(mnw_ctx *ctx, const mnw_float_desc *desc, uint32_t first_bits, uint64_t count,
                         uint64_t *mismatches, uint64_t *accepted) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    int rc = check_desc(ctx, desc);
    if (rc) return rc;
    if ((uint64_t)first_bits + count > (1ULL << 32)) return fail(ctx, MNW_ERR_ARG, "bit pattern range exceeds 2^32");
    CU(ctx->meta.reserve(64));
    launch_selftest_fastdiv(ctx->L, to_params(*desc), first_bits, count, ctx->meta.as<unsigned long long>());
    CU(cudaGetLastError());
    unsigned long long h[2] = {0, 0};
    CU(cudaMemcpyAsync(h, ctx->meta.p, 16, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    if (mismatches) *mismatches = h[0];
    if (accepted) *accepted = h[1];
    return MNW_OK;
}

int mnw_selftest_log10(mnw_ctx *ctx, uint32_t first_bits, uint64_t count, uint64_t *mismatches) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);
    if ((uint64_t)first_bits + count > (1ULL << 32)) return fail(ctx, MNW_ERR_ARG, "bit pattern range exceeds 2^32");
    CU(ctx->meta.reserve(64));
    launch_selftest_log10(ctx->L, first_bits, count, ctx->meta.as<unsigned long long>());
    CU(cudaGetLastError());
    unsigned long long h = 0;
    CU(cudaMemcpyAsync(&h, ctx->meta.p, 8, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    if (mismatches) *mismatches = h;
    return MNW_OK;
}

int mnw_selftest_pow10(mnw_ctx *ctx, uint32_t first_bits, uint64_t count, uint64_t *mismatches) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);
    if ((uint64_t)first_bits + count > (1ULL << 32)) return fail(ctx, MNW_ERR_ARG, "bit pattern range exceeds 2^32");
    CU(ctx->meta.reserve(64));
    launch_selftest_log10(ctx->L, first_bits, count, ctx->meta.as<unsigned long long>(), 1);
    CU(cudaGetLastError());
    unsigned long long h = 0;
    CU(cudaMemcpyAsync(&h, ctx->meta.p, 8, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    if (mismatches) *mismatches = h;
    return MNW_OK;
}

int mnw_pow10_f32(mnw_ctx *ctx, const float *x, int64_t n, float *out) {
    if (!ctx) return MNW_ERR_ARG;
    (void)cudaSetDevice(ctx->device);
    if (n < 0) return fail(ctx, MNW_ERR_ARG, "negative length");
    if (n == 0) return MNW_OK;
    CU(ctx->in.reserve(4 * (size_t)n));
    CU(ctx->dec_out.reserve(4 * (size_t)n));
    CU(cudaMemcpyAsync(ctx->in.p, x, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->L.stream));
    launch_pow10_f32(ctx->L, ctx->in.as<float>(), n, ctx->dec_out.as<float>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, ctx->dec_out.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    return MNW_OK;
}

int mnw_profile(mnw_ctx *ctx, int on) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    ctx->L.prof = on != 0;
    return MNW_OK;
}

int mnw_profile_summary(mnw_ctx *ctx, char *buf, int64_t cap) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    CU(cudaStreamSynchronize(ctx->L.stream));
    struct Agg { const char *name; int64_t n; double ms; };
    std::vector<Agg> agg;
    for (ProfRec &r : ctx->L.recs) {
        float ms = 0;
        if (r.a && r.b && cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) { (void)cudaGetLastError(); ms = 0; }
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
        bool found = false;
        for (Agg &a : agg) if (!strcmp(a.name, r.name)) { a.n++; a.ms += ms; found = true; break; }
        if (!found) agg.push_back({r.name, 1, (double)ms});
    }
    ctx->L.recs.clear();
    std::string s = "[";
    for (size_t i = 0; i < agg.size(); i++) {
        char line[256];
        snprintf(line, sizeof line, "%s{\"kernel\": \"%s\", \"launches\": %lld, \"ms\": %.6f}", i ? ", " : "",
                 agg[i].name, (long long)agg[i].n, agg[i].ms);
        s += line;
    }
    s += "]";
    if ((int64_t)s.size() + 1 > cap) return fail(ctx, MNW_ERR_CAPACITY, "profile summary needs %zu bytes", s.size() + 1);
    memcpy(buf, s.c_str(), s.size() + 1);
    return MNW_OK;
}

static inline float key_to_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

int mnw_vec3_limits_dev(mnw_ctx *ctx, const float *aos, int64_t np, int64_t nfiles, float *lo, float *hi) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    if (np <= 0 || nfiles < 0) return fail(ctx, MNW_ERR_ARG, "vec3 limits of an empty cube (the reference indexes vec[0])");
    if ((np + 32767) / 32768 >= 65536LL * 32768 || nfiles > 65535) return fail(ctx, MNW_ERR_ARG, "too many files in one call");
    CU(ctx->aux.reserve(24 * (size_t)nfiles + 64));
    launch_vec3_limits(ctx->L, aos, np, nfiles, ctx->aux.as<uint32_t>());
    CU(cudaGetLastError());
    std::vector<uint32_t> keys(6 * (size_t)nfiles);
    CU(cudaMemcpyAsync(keys.data(), ctx->aux.p, 24 * (size_t)nfiles, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(cudaStreamSynchronize(ctx->L.stream));
    for (int64_t f = 0; f < nfiles; f++)
        for (int k = 0; k < 3; k++) {
            float mn = key_to_float(keys[6 * f + k]), mx = key_to_float(keys[6 * f + 3 + k]);
            lo[3 * f + k] = mn;
            hi[3 * f + k] = nextafterf(mx, 2 * mx);  // go/minp/minp.go:94
        }
    return MNW_OK;
}

int mnw_vec3_limits(mnw_ctx *ctx, const float *aos, int64_t np, int64_t nfiles, float *lo, float *hi) {
    if (ctx) (void)cudaSetDevice(ctx->device);   /* a context may be used from any host thread */
    if (np <= 0 || nfiles < 0) return fail(ctx, MNW_ERR_ARG, "vec3 limits of an empty cube (the reference indexes vec[0])");
    CU(ctx->in.reserve(12 * (size_t)(np * nfiles) + 16));
    CU(cudaMemcpyAsync(ctx->in.p, aos, 12 * (size_t)(np * nfiles), cudaMemcpyHostToDevice, ctx->L.stream));
    return mnw_vec3_limits_dev(ctx, ctx->in.as<float>(), np, nfiles, lo, hi);
}

}  // extern "C"
