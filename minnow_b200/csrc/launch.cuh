// launch.cuh -- host-side launcher prototypes shared by api.cu and the kernel
// translation units.
#pragma once
#include <cstdint>
#include <mutex>
#include <vector>
#include <cuda_runtime.h>
#include "engine.cuh"

namespace mnw {

// Launch configuration that is a property of (kernel, DEVICE): the dynamic shared memory opt-in
// (cudaFuncSetAttribute) and the occupancy / SM-count queries.  One table per kernel instantiation,
// indexed by the calling thread's current device (every entry point sets it to its context's device first);
// initialised once per device under std::call_once, so that contexts on several GPUs, used from several
// host threads of one process, each get their own opt-in and grid size.
constexpr int MNW_MAX_DEVICES = 64;
struct DevCfg {
    std::once_flag once;
    cudaError_t err = cudaSuccess;
    int a = 0, b = 0;   // kernel-specific (co-resident clusters / CTAs, grid size)
};
inline DevCfg &dev_cfg(DevCfg (&table)[MNW_MAX_DEVICES]) {
    int dev = 0;
    cudaGetDevice(&dev);
    return table[(unsigned)dev % MNW_MAX_DEVICES];
}

// Optional per-kernel timing with CUDA events on the launching stream
// (mnw_profile): the roofline figure of bench.py is read from here.
struct ProfRec {
    const char *name;
    cudaEvent_t a, b;
};

struct Launcher {
    cudaStream_t stream = nullptr;
    int64_t count = 0;  // kernels launched (reported by mnw_launch_count)
    bool prof = false;
    std::vector<ProfRec> recs;
    void begin(const char *name) {
        if (!prof) return;
        ProfRec r = {name, nullptr, nullptr};
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        cudaEventRecord(r.a, stream);
        recs.push_back(r);
    }
    void end() {
        if (!prof) return;
        cudaEventRecord(recs.back().b, stream);
    }
};

using FloatParamsHost = FloatParams;

struct DecodeHost {
    int mode = 0;
    const uint8_t *data = nullptr;
    int64_t stream_len = 0;
    const int64_t *offsets = nullptr, *mins = nullptr, *bits = nullptr, *sel = nullptr;
    const int64_t *jitter_ids = nullptr;  // original block ids for the hash jitter when sel was compacted away
    int64_t n = 0, nsel = 0;
    const FloatParams *tab = nullptr;  // device table: 1 entry (group), 3 (vec3) or 3*nfiles (vec3, per file)
    int tab_per_file = 0;
    float wrap_L = 0;
    bool low_nonneg = false;   // all groups have low >= +0 and dx > 0
    int jmode = 0;
    unsigned long long seed = 0, block_id0 = 0;
    const double *u = nullptr;
    int32_t nfile = 0, subcells = 0;
    void *out = nullptr;
    void *const *outs = nullptr;   // device array: output pointer of every selected block (contiguous group decoders)
    bool any_log = true;           // some group of the table is a Log column (false: the kernel without 10^x is enough)
};

// kernels_generic.cu
void launch_build_contig(Launcher &L, BlockDesc *descs, int64_t nb, int32_t kind, const void *src, int64_t n,
                         const int64_t *starts, const int64_t *tile0, const int64_t *chunk0,
                         const FloatParamsHost &fp, int64_t blocks_per_chain, const int64_t *idx = nullptr,
                         BlockStat *stats_init = nullptr, void *ws_zero = nullptr);
void launch_build_vec3(Launcher &L, BlockDesc *descs, int64_t nfiles, const float *aos, int32_t nfile,
                       int32_t subcells, const FloatParams *tab, int tab_per_file);
// bounds() of minp.Writer.Vectors (go/minp/minp.go:291-300): keys[f*6 + k] = min, [f*6 + 3 + k] = max,
// as order-preserving uint32 keys (decode with key_to_float on the host).
void launch_vec3_limits(Launcher &L, const float *aos, int64_t np_per_file, int64_t nfiles, uint32_t *keys);
// kernels_boundary.cu: minh BoundaryWriter.Coordinates (go/minh/boundary.go:39-180)
void launch_bnd_count(Launcher &L, const float *x, const float *y, const float *z, int64_t n, float l, float boundary, int64_t cells,
                      int64_t *cnt, int64_t *sizes, int *err);
size_t bnd_sort_scratch_bytes(int64_t m);
cudaError_t launch_bnd_index(Launcher &L, const float *x, const float *y, const float *z, int64_t n, float l, float boundary,
                             int64_t cells, const int64_t *eoff, int64_t m, uint32_t *keys, int64_t *vals, uint32_t *keys2,
                             int64_t *vals2, void *scratch, int64_t *idx, int64_t *flags);
// kernels_text.cu: text.Reader.Block on the device (go/text/text.go:181-200, go/text/parse.go)
size_t text_tiles(int64_t len);
void launch_text_count(Launcher &L, const unsigned char *buf, int64_t len, int64_t *tile_nl);
void launch_text_starts(Launcher &L, const unsigned char *buf, int64_t len, const int64_t *tile_off, int64_t *line_start);
void launch_text_lines(Launcher &L, const unsigned char *buf, int64_t len, int64_t nlines, const int64_t *line_start, unsigned char sep,
                       unsigned char comm, int64_t *line_end, int64_t *keep, int *ncols);
void launch_text_parse(Launcher &L, const unsigned char *buf, int64_t nlines, const int64_t *line_start, const int64_t *line_end,
                       const int64_t *keep, const int64_t *row_of, unsigned char sep, int n_i, const int *icol, int n_f, const int *fcol,
                       const int *ncols, int64_t nrows, int64_t *iout, float *fout, int64_t *fb_list, int fb_cap, int *fb_count,
                       int *err, long long *err_line);
// kernels_regrid.cu: vectorGrid.Insert over a batch (go/minp/snapshot/grid.go:118-137,206-211)
void launch_regrid_insert(Launcher &L, const int64_t *ids, const float *vec, int64_t n, int64_t ncell, int64_t nside, float *grid, int *err);
void launch_vec3_params(Launcher &L, const uint32_t *keys, int64_t nfiles, float dx, FloatParams *tab, void *desc_out, int *skip,
                        int *abort_flag, int need_pipe);
void launch_params_from_desc(Launcher &L, const void *desc, int64_t n, FloatParams *tab);
void launch_generic_encode(Launcher &L, const BlockDesc *descs, BlockStat *stats, const BatchShape &sh,
                           int64_t *slow_list, int *slow_count, int *err, int64_t *mins, int64_t *bits,
                           int64_t *offsets, int64_t *out_len, uint8_t *out, int64_t chain_stride,
                           int64_t chain_cap, const int *run_if = nullptr, bool f32c = false, bool i64c = false);
// kernels_group.cu: the fused single-read encoder of uniform contiguous blocks (k_init + k_group_fused + the wide
// blocks through k_pack).  flags: the context's device flag words.  ws: group_fused_ws_bytes(nblocks) of scratch.
size_t group_fused_ws_bytes(int64_t nblocks);
cudaError_t launch_group_encode(Launcher &L, const BlockDesc *descs, BlockStat *stats, const BatchShape &sh, int *flags,
                                int64_t *mins, int64_t *bits, int64_t *offsets, int64_t *out_len, uint8_t *out,
                                int64_t chain_stride, int64_t chain_cap, void *ws, bool has_i64, bool prepared = false,
                                void *logws = nullptr);   // logws: 4 * uniform_n bytes per block up to the last log10 block, or null
void launch_init_stats(Launcher &L, const BlockDesc *descs, BlockStat *stats, int64_t nb);
void launch_pack_list(Launcher &L, const BlockDesc *descs, const BlockStat *stats, const BatchShape &sh,
                      const int64_t *list, const int *list_count, uint8_t *out, int64_t chain_stride,
                      int64_t chain_cap, int *err);
size_t scan_scratch_bytes(int64_t n);   // scratch launch_scan_sizes wants for n sizes (0: none)
cudaError_t launch_scan_sizes(Launcher &L, const int64_t *sizes, int64_t n, int64_t base, int64_t *offsets,
                              int64_t *total, void *scratch = nullptr);
void launch_raw_pack(Launcher &L, BlockDesc *descs, BlockStat *stats, const void *src, int64_t n, int bits,
                     uint8_t *out);
void launch_umax(Launcher &L, const unsigned long long *x, int64_t n, unsigned long long *out);
void launch_decode(Launcher &L, const DecodeHost &h);

}  // namespace mnw
