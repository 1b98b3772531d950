// launch.cuh -- host-side launcher prototypes shared by api.cu and the kernel
// translation units.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "engine.cuh"

namespace mnw {

struct Launcher {
    cudaStream_t stream = nullptr;
    int64_t count = 0;  // kernels launched (reported by mnw_launch_count)
};

struct FloatParamsHost {
    float low = 0, high = 0, dx = 0, hi_clamp = 0;
    int64_t pixels = 0;
    int32_t flags = 0;
};

struct DecodeHost {
    int mode = 0;
    const uint8_t *data = nullptr;
    int64_t stream_len = 0;
    const int64_t *offsets = nullptr, *mins = nullptr, *bits = nullptr, *sel = nullptr;
    const int64_t *jitter_ids = nullptr;  // original block ids for the hash jitter when sel was compacted away
    int64_t n = 0, nsel = 0;
    float low[3] = {0, 0, 0}, dx[3] = {0, 0, 0};
    int64_t pixels[3] = {0, 0, 0};
    int periodic[3] = {0, 0, 0};
    float wrap_L = 0;
    int jmode = 0;
    unsigned long long seed = 0, block_id0 = 0;
    const double *u = nullptr;
    int32_t nfile = 0, subcells = 0;
    void *out = nullptr;
};

// kernels_generic.cu
void launch_build_contig(Launcher &L, BlockDesc *descs, int64_t nb, int32_t kind, const void *src, int64_t n,
                         const int64_t *starts, const int64_t *tile0, const int64_t *chunk0,
                         const FloatParamsHost &fp, int64_t blocks_per_chain);
void launch_build_vec3(Launcher &L, BlockDesc *descs, int64_t nfiles, const float *aos, int32_t nfile,
                       int32_t subcells, const FloatParamsHost fp[3]);
void launch_generic_encode(Launcher &L, const BlockDesc *descs, BlockStat *stats, const BatchShape &sh,
                           int64_t *slow_list, int *slow_count, int *err, int64_t *mins, int64_t *bits,
                           int64_t *offsets, int64_t *out_len, uint8_t *out, int64_t chain_stride,
                           int64_t chain_cap);
cudaError_t launch_scan_sizes(Launcher &L, const int64_t *sizes, int64_t n, int64_t base, int64_t *offsets,
                              int64_t *total);
void launch_raw_pack(Launcher &L, BlockDesc *descs, BlockStat *stats, const void *src, int64_t n, int bits,
                     uint8_t *out);
void launch_umax(Launcher &L, const unsigned long long *x, int64_t n, unsigned long long *out);
void launch_decode(Launcher &L, const DecodeHost &h);

}  // namespace mnw
