// kernels_pipe.cu -- k_pipe_vec3: the warp-specialised, software-pipelined minp encode for 64^3
// sub-cells (see pipe_vec3.cuh), and its launcher.
#include <climits>

#include "fused_detail.cuh"
#include "group_detail.cuh"

namespace mnw {

#include "pipe_vec3.cuh"

cudaError_t launch_pipe_vec3(Launcher &L, const FusedArgs &A) {
    auto kern = k_pipe_vec3<false, 64>;
    const size_t smem = (size_t)6 * PIPE_CHUNK + (size_t)PIPE_PW * PIPE_TBUF * 4;
    static DevCfg cfgs[MNW_MAX_DEVICES];
    DevCfg &dc = dev_cfg(cfgs);
    cudaError_t e;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PIPE_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(PIPE_NT); cfg.dynamicSmemBytes = smem; cfg.stream = L.stream;
    cfg.attrs = attr; cfg.numAttrs = 1;
    std::call_once(dc.once, [&] {
        dc.err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (dc.err != cudaSuccess) return;
        cfg.gridDim = dim3(PIPE_CS);
        dc.err = cudaOccupancyMaxActiveClusters(&dc.a, kern, &cfg);
        if (dc.err == cudaSuccess && dc.a < 1) dc.err = cudaErrorLaunchOutOfResources;
        if (getenv("MNW_DEBUG")) fprintf(stderr, "k_pipe_vec3: %d co-resident clusters, %zu B dynamic smem\n", dc.a, smem);
    });
    if (dc.err != cudaSuccess) return dc.err;
    const int max_clusters = dc.a;
    const long long clusters = A.nunits < max_clusters ? A.nunits : max_clusters;
    cfg.gridDim = dim3((unsigned)(clusters * PIPE_CS));
    L.begin("k_pipe_vec3");
    e = cudaLaunchKernelEx(&cfg, kern, A);
    L.end();
    L.count++;
#ifdef MNW_PIPE_DBG
    if (getenv("MNW_PIPE_DBG")) {
        cudaDeviceSynchronize();
        static unsigned long long h[32 * 64 * 8];
        cudaMemcpyFromSymbol(h, g_pipe_dbg, sizeof(h));
        for (int c = 0; c < 15; c += 7)
            for (int it = 0; it < 40; it++) {
                const unsigned long long *r = h + (c * 64 + it) * 8, t0 = h[(c * 64) * 8];
                fprintf(stderr, "c%d it%2d top %7.1f | load %5.1f | stats +%5.1f | fin +%5.1f | lb0 +%5.1f lb1 +%5.1f lb2 +%5.1f | packed +%5.1f\n", c, it,
                        (r[0] - t0) / 1e3, (r[1] - r[0]) / 1e3, (r[2] - r[1]) / 1e3, (r[3] - r[2]) / 1e3, (r[4] - r[3]) / 1e3,
                        (r[5] - r[4]) / 1e3, (r[6] - r[5]) / 1e3, (r[7] - r[6]) / 1e3);
            }
    }
#endif
    return e;
}

// Cluster-free variant: cooperative launch of one CTA per SM (rounded down to a multiple of 8, so that the eight
// parts of a unit are always items of the same round), statistics records in `ustat` (16 words per unit, zeroed
// here).  Returns cudaErrorCooperativeLaunchTooLarge / NotSupported when the device cannot hold the grid; the
// caller then takes the cluster kernel.
// one 128-byte record per (unit, part): 13 words {tag 1 | statistic}, written once by the part's CTA, polled by all of the unit's
// (128^3 sub-cells, 64 parts: one 64-byte record per unit, combined by atomics)
size_t pipe_coop_ws_bytes(int64_t nunits, int nsub) { return (size_t)nunits * (size_t)(nsub == 64 ? 8 * 128 : 64); }

template <int NSUB>
static cudaError_t launch_pipe_vec3_coop_t(Launcher &L, FusedArgs A, void *ws) {
    constexpr int PARTS = NSUB * NSUB * NSUB / PIPE_CHUNK;   // CTAs per unit: 1, 8 or 64
    auto kern = k_pipe_vec3<true, NSUB>;
    const size_t smem = (size_t)6 * PIPE_CHUNK + (size_t)PIPE_PW * PIPE_TBUF * 4;
    static DevCfg cfgs[MNW_MAX_DEVICES];
    DevCfg &dc = dev_cfg(cfgs);
    cudaError_t e;
    std::call_once(dc.once, [&] {
        int dev = 0, sms = 0, coop = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        dc.err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (dc.err != cudaSuccess) return;
        dc.err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PIPE_NT, smem);
        if (dc.err != cudaSuccess) return;
        dc.a = (coop && per_sm >= 1) ? (sms / PARTS) * PARTS : 0;
        if (dc.a && getenv("MNW_PIPE_GRID")) dc.a = atoi(getenv("MNW_PIPE_GRID"));   // tuning knob (any value >= PARTS is safe)
        if (getenv("MNW_DEBUG")) fprintf(stderr, "k_pipe_vec3<coop, %d>: grid %d, %zu B dynamic smem\n", NSUB, dc.a, smem);
    });
    if (dc.err != cudaSuccess) return dc.err;
    const int grid_max = dc.a;
    if (grid_max < PARTS) return cudaErrorNotSupported;
    const long long items = PARTS * A.nunits;
    const unsigned grid = (unsigned)(items < grid_max ? items : grid_max);   // >= PARTS: no CTA ever holds two parts of a unit
    A.ustat = (unsigned *)ws;
    e = cudaMemsetAsync(ws, 0, pipe_coop_ws_bytes(A.nunits, NSUB), L.stream);
    if (e != cudaSuccess) return e;
    void *args[] = {(void *)&A};
    L.begin("k_pipe_vec3");
    e = cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(PIPE_NT), args, smem, L.stream);
    L.end();
    L.count++;
#ifdef MNW_PIPE_DBG
    if (getenv("MNW_PIPE_DBG")) {   // per-part timeline of CTAs 0, 56, 112 (%globaltimer stamps, microseconds; build with MNW_DEFINES=MNW_PIPE_DBG)
        cudaDeviceSynchronize();
        static unsigned long long h[32 * 64 * 8];
        cudaMemcpyFromSymbol(h, g_pipe_dbg, sizeof(h));
        for (int c = 0; c < 15; c += 7)
            for (int it = 2; it < 14; it++) {
                const unsigned long long *r = h + (c * 64 + it) * 8, t0 = h[(c * 64) * 8];
                fprintf(stderr, "c%d it%2d top %7.1f | load %5.1f | stats +%5.1f | fin +%5.1f | off +%5.1f | pre +%5.1f | packed +%5.1f\n", c, it,
                        (r[0] - t0) / 1e3, (r[1] - r[0]) / 1e3, (r[2] - r[1]) / 1e3, (r[3] - r[2]) / 1e3, (r[4] - r[3]) / 1e3,
                        (r[6] - r[4]) / 1e3, (r[7] - r[6]) / 1e3);
            }
    }
#endif
    return e;
}

cudaError_t launch_pipe_vec3_coop(Launcher &L, FusedArgs A, void *ws, int nsub) {
    if (nsub == 128) return launch_pipe_vec3_coop_t<128>(L, A, ws);
    return nsub == 32 ? launch_pipe_vec3_coop_t<32>(L, A, ws) : launch_pipe_vec3_coop_t<64>(L, A, ws);
}

bool pipe_vec3_supported(const FloatParamsHost *fp, int64_t nparams) {
    for (int64_t i = 0; i < nparams; i++)
        if (fp[i].pixels > (1LL << 22)) return false;   // q + rotation must stay below 2^23 (float-exact floor)
    return true;
}


}  // namespace mnw
