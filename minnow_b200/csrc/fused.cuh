// fused.cuh -- single-read fused kernels (kernels_fused.cu).
//
// Encode: one thread-block cluster per minp sub-cell reads the sub-cell's AoS
// rows ONCE, quantises all three axes, keeps the 16-bit rotated pixel indices in
// (distributed) shared memory, reduces the per-block statistics across the
// cluster, obtains the block byte offsets by decoupled look-back over the
// sub-cells of the file, and packs straight from shared memory.
// Decode: one CTA per slab of a sub-cell unpacks all three axes and writes whole
// AoS rows.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "engine.cuh"
#include "launch.cuh"

namespace mnw {

// Device workspace of one fused encode launch (zeroed before the launch).
struct FusedWork {
    unsigned long long *pub;   // [nblocks] look-back words: flag (2 bits) | bytes (62 bits)
    unsigned int *ticket;      // next unit to claim
    int64_t *repack_list;      // blocks whose bits > 16: packed afterwards from global memory
    int *repack_count;
    int *abort_flag;           // some block needs the exact sequential periodicMin: rerun generically
    int *err;
    const int *skip;           // (may be null) set before the launch by k_vec3_params: the group parameters derived on
                               // the device are outside what the fused kernels cover -- every CTA returns at once
};

// minp sub-cell gather + 3-axis encode.  Supported when nfile/subcells is 16, 32
// or 64, every group is periodic without minh pre-transform, and pixels < 2^30.
bool fused_vec3_supported(const FloatParamsHost *fp, int64_t nparams, int nfile, int subcells, const void *aos);
size_t fused_work_bytes(int64_t nblocks);
cudaError_t launch_fused_vec3(Launcher &L, const FusedWork &W, const BlockDesc *descs, const FloatParams *tab, int tab_per_file,
                              const float *aos, int nfile, int subcells, int64_t nfiles, BlockStat *stats,
                              int64_t *mins, int64_t *bits, int64_t *offsets, int64_t *out_len, uint8_t *out,
                              int64_t out_axis_stride, bool pipe_ok, void *coop_ws, bool *ran_pipe = nullptr);
size_t pipe_coop_ws_bytes(int64_t nunits, int nsub);   // workspace of the cooperative k_pipe_vec3 schedule
// k_pipe_vec3 (the warp-specialised 64^3 encode) applies when every group has pixels <= 2^22.
bool pipe_vec3_supported(const FloatParamsHost *fp, int64_t nparams);

// 3-axis decode + periodic wrap + sub-cell scatter.
bool fused_decode_vec3_supported(int nfile, int subcells, const void *aos_out);
cudaError_t launch_fused_decode_vec3(Launcher &L, const DecodeHost &h, int64_t nfiles);

// quantize_fast vs __fdiv_rn over float bit patterns [first, first + count):
// d_out2[0] = mismatches among accepted fast results, d_out2[1] = accepted results.
void launch_selftest_fastdiv(Launcher &L, const FloatParamsHost &fp, unsigned long long first,
                             unsigned long long count, unsigned long long *d_out2);

void launch_selftest_log10(Launcher &L, unsigned long long first, unsigned long long count, unsigned long long *d_out, int pow10 = 0);
void launch_pow10_f32(Launcher &L, const float *x, long long n, float *out);

}  // namespace mnw
