// pipe_api.cu -- mnw_pipe: pipelined host <-> device streaming of minp files through the block encoder / decoder
// (the staging either side of the kernels -- go/writer.go:107-141, go/reader.go:114-127,
// go/bit/bit.go:161-181 read and write one block at a time through one file cursor).
//
// A pipe is a ring of `depth` slots; a slot is a private context (stream + device staging) plus pinned metadata.
// Submitting a file enqueues  upload -> kernels -> download  on the slot's stream and returns at once, so that with ONE
// host thread the upload of file i+1, the kernels of file i and the download of file i-1 run at the same time (the copy
// engines of the two directions and the SMs are separate resources).  An encode has two device phases, because the
// packed bytes can only be copied out once their lengths are known: phase A (upload, limits, parameters, encode,
// metadata download) and phase B (download of exactly out_len[k] bytes per axis).  The host moves a slot from A to B
// when it next touches the pipe (submit / wait / poll): no thread, no callback.
#include <cstring>
#include <vector>

#include <sched.h>

#include "ctx.cuh"
#include "device_math.cuh"

using namespace mnw;

namespace {

enum SlotState { SLOT_FREE = 0, SLOT_ENC_A, SLOT_ENC_B, SLOT_DEC };

struct Slot {
    mnw_ctx *ctx = nullptr;
    int state = SLOT_FREE;
    int64_t ticket = -1, done_ticket = -1;
    int done_status = MNW_OK;
    cudaEvent_t ev = nullptr;
    int64_t *h_meta = nullptr;   // pinned: mins | bits | offsets | lens[3] | desc[3] (9 words) | device error word
    size_t h_meta_cap = 0;
    DevBuf desc_dev;
    // the submitted call
    int64_t nb = 0, dstride = 0, out_axis_stride = 0;
    int64_t *mins = nullptr, *bits = nullptr, *offsets = nullptr, *out_len = nullptr;
    mnw_float_desc *desc_out = nullptr;
    uint8_t *out = nullptr;
};

}  // namespace

struct mnw_pipe {
    int device = 0, depth = 0;
    std::vector<Slot> slots;
    int64_t next_ticket = 0;
    int sticky = MNW_OK;          // status of a failed ticket nobody waited for
    std::string err;
};

namespace {

int pipe_fail(mnw_pipe *p, int code, const char *msg) {
    p->err = msg;
    return code;
}

int reserve_meta(Slot &s, size_t words) {
    if (words <= s.h_meta_cap) return MNW_OK;
    if (s.h_meta) cudaFreeHost(s.h_meta);
    s.h_meta = nullptr; s.h_meta_cap = 0;
    if (cudaMallocHost((void **)&s.h_meta, 8 * (words + words / 4 + 64)) != cudaSuccess) return MNW_ERR_CUDA;
    s.h_meta_cap = words + words / 4 + 64;
    return MNW_OK;
}

// phase B of an encode: the lengths are on the host, copy exactly those bytes out
int enqueue_phase_b(mnw_pipe *p, Slot &s) {
    mnw_ctx *ctx = s.ctx;
    const int64_t nb = s.nb;
    const int64_t *lens = s.h_meta + 3 * nb;
    const int err = (int)s.h_meta[3 * nb + 3 + 9];
    if (err) { s.done_status = mnw_report_device_error(ctx, err); p->err = ctx->err; }
    for (int k = 0; k < 3 && s.done_status == MNW_OK; k++) {
        if (lens[k] > s.out_axis_stride) {
            s.done_status = MNW_ERR_CAPACITY;
            p->err = "axis needs more bytes than out_axis_stride";
        } else if (lens[k] > 0) {
            CU(cudaMemcpyAsync(s.out + k * s.out_axis_stride, ctx->out.as<uint8_t>() + k * s.dstride, (size_t)lens[k],
                               cudaMemcpyDeviceToHost, ctx->L.stream));
        }
    }
    CU(cudaEventRecord(s.ev, ctx->L.stream));
    s.state = SLOT_ENC_B;
    return MNW_OK;
}

void finalize(mnw_pipe *p, Slot &s) {
    if (s.state == SLOT_ENC_B) {
        const int64_t nb = s.nb;
        if (s.mins) memcpy(s.mins, s.h_meta, 8 * (size_t)nb);
        if (s.bits) memcpy(s.bits, s.h_meta + nb, 8 * (size_t)nb);
        if (s.offsets) memcpy(s.offsets, s.h_meta + 2 * nb, 8 * (size_t)nb);
        if (s.out_len) memcpy(s.out_len, s.h_meta + 3 * nb, 24);
        if (s.desc_out) memcpy(s.desc_out, s.h_meta + 3 * nb + 3, 3 * sizeof(mnw_float_desc));
    }
    s.done_ticket = s.ticket;
    if (s.done_status != MNW_OK && p->sticky == MNW_OK) p->sticky = s.done_status;
    s.state = SLOT_FREE;
}

// Move every slot as far as its device work allows, without blocking.
int pump(mnw_pipe *p) {
    for (Slot &s : p->slots) {
        if (s.state == SLOT_FREE) continue;
        const cudaError_t q = cudaEventQuery(s.ev);
        if (q == cudaErrorNotReady) continue;
        if (q != cudaSuccess) {
            s.done_status = MNW_ERR_CUDA;
            p->err = cudaGetErrorString(q);
            s.state = s.state == SLOT_ENC_A ? SLOT_ENC_B : s.state;
            finalize(p, s);
            continue;
        }
        if (s.state == SLOT_ENC_A) {
            mnw_ctx *ctx = s.ctx;
            (void)ctx;
            const int rc = enqueue_phase_b(p, s);
            if (rc != MNW_OK) { s.done_status = rc; p->err = s.ctx->err; s.state = SLOT_ENC_B; finalize(p, s); }
        } else {
            finalize(p, s);
        }
    }
    return MNW_OK;
}

Slot *acquire(mnw_pipe *p, int64_t *ticket) {
    Slot &s = p->slots[(size_t)(p->next_ticket % p->depth)];
    while (s.state != SLOT_FREE) {
        pump(p);
        if (s.state != SLOT_FREE) sched_yield();
    }
    s.ticket = p->next_ticket++;
    s.done_status = MNW_OK;
    if (ticket) *ticket = s.ticket;
    return &s;
}

}  // namespace

extern "C" {

int mnw_pipe_create(int device, int depth, mnw_pipe **out) {
    if (!out || depth < 1 || depth > 64) return MNW_ERR_ARG;
    mnw_pipe *p = new mnw_pipe();
    p->device = device; p->depth = depth;
    p->slots.resize((size_t)depth);
    for (Slot &s : p->slots) {
        int rc = mnw_create(device, &s.ctx);
        if (rc == MNW_OK && cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming) != cudaSuccess) rc = MNW_ERR_CUDA;
        if (rc != MNW_OK) {
            for (Slot &t : p->slots) { if (t.ev) cudaEventDestroy(t.ev); if (t.ctx) mnw_destroy(t.ctx); }
            delete p;
            return rc;
        }
    }
    *out = p;
    return MNW_OK;
}

void mnw_pipe_destroy(mnw_pipe *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (Slot &s : p->slots) {
        if (s.ctx) cudaStreamSynchronize(s.ctx->L.stream);
        if (s.ev) cudaEventDestroy(s.ev);
        if (s.h_meta) cudaFreeHost(s.h_meta);
        s.desc_dev.release();
        if (s.ctx) mnw_destroy(s.ctx);
    }
    delete p;
}

const char *mnw_pipe_last_error(const mnw_pipe *p) { return p ? p->err.c_str() : ""; }

int mnw_pipe_minp_encode_vectors(mnw_pipe *p, const float *aos, int64_t nfile, int64_t subcells, int periodic, float L, float dx,
                                 mnw_float_desc desc_out[3], int64_t *mins, int64_t *bits, int64_t *offsets, uint8_t *out,
                                 int64_t out_axis_stride, int64_t out_len[3], int64_t *ticket) {
    if (!p) return MNW_ERR_ARG;
    (void)cudaSetDevice(p->device);
    if (nfile <= 0 || subcells <= 0 || nfile % subcells != 0 || nfile > 2048 || !aos || !out)
        return pipe_fail(p, MNW_ERR_ARG, "mnw_pipe_minp_encode_vectors: bad argument");
    Slot *sp = acquire(p, ticket);
    Slot &s = *sp;
    mnw_ctx *ctx = s.ctx;
    const int64_t np = nfile * nfile * nfile, sc3 = subcells * subcells * subcells, nb = 3 * sc3;
    const int64_t dstride = (8 * np + 255) & ~255LL;   // worst case: 64 bits per value
    s.nb = nb; s.dstride = dstride; s.out_axis_stride = out_axis_stride;
    s.mins = mins; s.bits = bits; s.offsets = offsets; s.out_len = out_len; s.desc_out = desc_out; s.out = out;
    const size_t words = 3 * (size_t)nb + 3 + 9 + 1;
    int rc = reserve_meta(s, words);
    auto bail = [&](int code) { s.done_status = code; p->err = ctx->err; s.done_ticket = s.ticket; if (p->sticky == MNW_OK) p->sticky = code; return code; };
    if (rc) return bail(rc);
    if (ctx->in.reserve(12 * (size_t)np + 16) != cudaSuccess || ctx->out.reserve(3 * (size_t)dstride + 64) != cudaSuccess ||
        ctx->meta.reserve(8 * words) != cudaSuccess || s.desc_dev.reserve(3 * sizeof(mnw_float_desc)) != cudaSuccess) {
        (void)cudaGetLastError();
        return bail(mnw_fail(ctx, MNW_ERR_CUDA, "pipe: out of device memory"));
    }
    if (cudaMemcpyAsync(ctx->in.p, aos, 12 * (size_t)np, cudaMemcpyHostToDevice, ctx->L.stream) != cudaSuccess)
        return bail(mnw_fail(ctx, MNW_ERR_CUDA, "pipe: upload failed"));
    int64_t *d_meta = ctx->meta.as<int64_t>();
    rc = mnw_minp_encode_vectors_dev(ctx, ctx->in.as<float>(), nfile, subcells, 1, periodic, L, dx, (mnw_float_desc *)s.desc_dev.p, d_meta,
                                     d_meta + nb, d_meta + 2 * nb, ctx->out.as<uint8_t>(), dstride, d_meta + 3 * nb);
    if (rc) return bail(rc);
    cudaMemcpyAsync(d_meta + 3 * nb + 3, s.desc_dev.p, 3 * sizeof(mnw_float_desc), cudaMemcpyDeviceToDevice, ctx->L.stream);
    cudaMemsetAsync(d_meta + 3 * nb + 3 + 9, 0, 8, ctx->L.stream);
    cudaMemcpyAsync(d_meta + 3 * nb + 3 + 9, ctx->flags.as<int>() + FLAG_ERR, sizeof(int), cudaMemcpyDeviceToDevice, ctx->L.stream);
    cudaMemsetAsync(ctx->flags.as<int>() + FLAG_ERR, 0, sizeof(int), ctx->L.stream);   // reported through this ticket
    cudaMemcpyAsync(s.h_meta, d_meta, 8 * words, cudaMemcpyDeviceToHost, ctx->L.stream);
    if (cudaEventRecord(s.ev, ctx->L.stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
        return bail(mnw_fail(ctx, MNW_ERR_CUDA, "pipe: enqueue failed"));
    s.state = SLOT_ENC_A;
    return MNW_OK;
}

int mnw_pipe_minp_decode_vectors(mnw_pipe *p, const mnw_float_desc desc[3], const uint8_t *const data[3], const int64_t data_len[3],
                                 const int64_t *offsets, const int64_t *mins, const int64_t *bits, int64_t nfile, int64_t subcells,
                                 float wrap_L, const mnw_jitter *jitter, float *aos_out, int64_t *ticket) {
    if (!p) return MNW_ERR_ARG;
    (void)cudaSetDevice(p->device);
    if (nfile <= 0 || subcells <= 0 || nfile % subcells != 0 || !desc || !data || !aos_out)
        return pipe_fail(p, MNW_ERR_ARG, "mnw_pipe_minp_decode_vectors: bad argument");
    const int64_t np = nfile * nfile * nfile, sc3 = subcells * subcells * subcells, nb = 3 * sc3;
    const int64_t nsub = nfile / subcells, n = nsub * nsub * nsub;
    for (int64_t b = 0; b < nb; b++) {
        if (bits[b] < 0 || bits[b] > 64) return pipe_fail(p, MNW_ERR_FORMAT, "a block has more than 64 bits");
        if (offsets[b] < 0 || offsets[b] + array_bytes(bits[b], n) > data_len[b / sc3])
            return pipe_fail(p, MNW_ERR_FORMAT, "a block lies outside its group's data");
    }
    Slot *sp = acquire(p, ticket);
    Slot &s = *sp;
    mnw_ctx *ctx = s.ctx;
    auto bail = [&](int code) { s.done_status = code; p->err = ctx->err; s.done_ticket = s.ticket; if (p->sticky == MNW_OK) p->sticky = code; return code; };
    int64_t stride = 0;
    for (int k = 0; k < 3; k++) stride = data_len[k] > stride ? data_len[k] : stride;
    stride = (stride + 255) & ~255LL;
    int rc = reserve_meta(s, 3 * (size_t)nb);
    if (rc) return bail(rc);
    if (ctx->in.reserve(3 * (size_t)stride + 16) != cudaSuccess || ctx->meta.reserve(8 * 3 * (size_t)nb) != cudaSuccess ||
        ctx->dec_out.reserve(12 * (size_t)np) != cudaSuccess) {
        (void)cudaGetLastError();
        return bail(mnw_fail(ctx, MNW_ERR_CUDA, "pipe: out of device memory"));
    }
    // the metadata go through the slot's pinned words, so that the caller's (pageable) arrays are free on return
    memcpy(s.h_meta, offsets, 8 * (size_t)nb);
    memcpy(s.h_meta + nb, mins, 8 * (size_t)nb);
    memcpy(s.h_meta + 2 * nb, bits, 8 * (size_t)nb);
    int64_t *d_meta = ctx->meta.as<int64_t>();
    cudaMemcpyAsync(d_meta, s.h_meta, 8 * 3 * (size_t)nb, cudaMemcpyHostToDevice, ctx->L.stream);
    for (int k = 0; k < 3; k++)
        if (data_len[k] > 0)
            cudaMemcpyAsync(ctx->in.as<uint8_t>() + k * stride, data[k], (size_t)data_len[k], cudaMemcpyHostToDevice, ctx->L.stream);
    rc = mnw_decode_vec3_subcells_dev(ctx, desc, 0, ctx->in.as<uint8_t>(), stride, d_meta, d_meta + nb, d_meta + 2 * nb, nfile, subcells, 1,
                                      wrap_L, jitter, ctx->dec_out.as<float>());
    if (rc) return bail(rc);
    cudaMemcpyAsync(aos_out, ctx->dec_out.p, 12 * (size_t)np, cudaMemcpyDeviceToHost, ctx->L.stream);
    if (cudaEventRecord(s.ev, ctx->L.stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
        return bail(mnw_fail(ctx, MNW_ERR_CUDA, "pipe: enqueue failed"));
    s.state = SLOT_DEC;
    return MNW_OK;
}

int mnw_pipe_poll(mnw_pipe *p) {
    if (!p) return MNW_ERR_ARG;
    (void)cudaSetDevice(p->device);
    return pump(p);
}

int mnw_pipe_wait(mnw_pipe *p, int64_t ticket) {
    if (!p) return MNW_ERR_ARG;
    (void)cudaSetDevice(p->device);
    if (ticket < 0 || ticket >= p->next_ticket) return pipe_fail(p, MNW_ERR_ARG, "mnw_pipe_wait: unknown ticket");
    Slot &s = p->slots[(size_t)(ticket % p->depth)];
    while (s.ticket == ticket && s.state != SLOT_FREE) {
        pump(p);
        if (s.state != SLOT_FREE) sched_yield();
    }
    if (s.done_ticket == ticket) {
        const int st = s.done_status;
        if (st != MNW_OK && p->sticky == st) p->sticky = MNW_OK;   // reported
        return st;
    }
    return MNW_OK;   // an older ticket whose slot has been reused: it completed; failures surface through mnw_pipe_drain
}

int mnw_pipe_drain(mnw_pipe *p) {
    if (!p) return MNW_ERR_ARG;
    (void)cudaSetDevice(p->device);
    for (;;) {
        bool busy = false;
        pump(p);
        for (Slot &s : p->slots) busy = busy || s.state != SLOT_FREE;
        if (!busy) break;
        sched_yield();
    }
    const int st = p->sticky;
    p->sticky = MNW_OK;
    return st;
}

}  // extern "C"
