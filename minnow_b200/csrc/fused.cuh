// fused.cuh -- single-read fused encode kernels (kernels_fused.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "engine.cuh"
#include "launch.cuh"

namespace mnw {

// One FloatGroup of nblocks equal blocks of n contiguous float32.
bool fused_group_supported(const FloatParamsHost &fp, int64_t n, int64_t nblocks);
cudaError_t launch_fused_group(Launcher &L, void *ws, size_t ws_cap, const FloatParamsHost &fp, const float *x,
                               int64_t n, int64_t nblocks, int64_t *mins, int64_t *bits, int64_t *offsets,
                               int64_t *out_len, uint8_t *out, int64_t out_cap, int *flags);

// minp sub-cell gather + 3-axis encode.
bool fused_vec3_supported(const FloatParamsHost *fp, int64_t nparams, int nfile, int subcells);
cudaError_t launch_fused_vec3(Launcher &L, const FloatParams *tab, int tab_per_file, const float *aos, int nfile, int subcells,
                              int64_t nfiles, int64_t *mins, int64_t *bits, int64_t *offsets, int64_t *out_len,
                              uint8_t *out, int64_t out_axis_stride, int *flags);

}  // namespace mnw
