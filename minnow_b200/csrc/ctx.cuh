// ctx.cuh -- the context object behind the C ABI (include/minnow_cuda.h), shared by api.cu and pipe_api.cu.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/minnow_cuda.h"
#include "engine.cuh"
#include "launch.cuh"

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {  // retry with the exact size
            (void)cudaGetLastError();
            want = n;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

struct mnw_ctx {
    int device = 0;
    mnw::Launcher L;
    std::string err;
    int last_path = 0;
    int force_generic = 0;
    DevBuf in, out, descs, stats, slow, flags, meta, aux, dec_out, ustream, fused_ws, params, coop_ws, group_ws, group_log, dec_cols;
    int *h_flags = nullptr;  // pinned: [slow_count, err]
    void *h_stage = nullptr; // pinned staging for gathered uploads (grow-only)
    size_t h_stage_cap = 0;
    bool flags_init = false; // the device flag words have been zeroed once (the error word is sticky afterwards)
    // minh BoundaryWriter: the per-cell index lists of the last mnw_boundary_coordinates call, resident on the device
    DevBuf bnd_idx, bnd_flags, bnd_work;
    std::vector<int64_t> bnd_starts;   // [cells^3 + 1], host
    int64_t bnd_n = -1, bnd_m = 0;
    // text.Reader.Block: the parsed columns of the last mnw_text_parse_block call, resident on the device
    DevBuf txt_work, txt_i, txt_f, txt_fb;
    int64_t txt_rows = -1, txt_nfb = 0;
    int txt_ni = 0, txt_nf = 0;
    void *comm = nullptr;    // ncclComm_t of the sharded path (comm_api.cu); null = a world of one
    int comm_ranks = 0, comm_rank = 0;
};

// sets the context's (or, for c == NULL, the creation) error message; returns code
int mnw_fail(mnw_ctx *c, int code, const char *fmt, ...);
// internal: the stream has been synchronised by the caller or will be; surfaces the device error word copied to the host
int mnw_report_device_error(mnw_ctx *ctx, int err);

#define CU(call)                                                                           \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return mnw_fail(ctx, MNW_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__));   \
    } while (0)
