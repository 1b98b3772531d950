// kernels_boundary.cu -- minh BoundaryWriter.Coordinates on the device (go/minh/boundary.go:39-180): which cells of a
// cells^3 grid every point belongs to -- its own cell and, when it lies within `boundary` of a face, edge or corner,
// up to 7 neighbouring cells (periodic) -- and from that the per-cell index lists (with boundary flags) in exactly
// the reference's order: cell after cell, within a cell by point index, a point's own cell before its ghost cells in
// hostCells order.
//
//   k_bnd_count   per point: idxReg + region + hostCells -> number of host cells (1, 2, 4, 8); per cell: size (atomics)
//   (scan)        entry offset of every point
//   k_bnd_emit    per point: its entries (cell id, point << 1 | flag) in hostCells order
//   k_rs_hist / (scan) / k_rs_scatter   STABLE least-significant-digit radix sort of the entries by cell id, 8 bits per
//                 pass: one warp per tile of 4096 entries ranks its entries 32 at a time with match.any, so equal keys
//                 keep their order -- which is the reference's insertion order
//   k_bnd_split   (point, flag) out of the sorted entries
//
// HBM-bound integer work (24 + 12 * passes * 2 bytes per entry); it runs once per catalogue, ahead of the column
// encoders, which then gather through the index without it ever leaving the device.
#include "engine.cuh"
#include "device_math.cuh"
#include "launch.cuh"

namespace mnw {

struct BndParams {
    float l, dx, sb;   // box size, cell width l / float32(cells), scaledBoundary = boundary / dx (go/minh/boundary.go:40)
    int cells;
};

namespace {

// go/minh/boundary.go:173-180
__device__ __forceinline__ int bnd_region(const BndParams &p, long long ix, float x) {
    const float low = __ll2float_rn(ix);
    if (x < __fadd_rn(low, p.sb)) return -1;
    if (x >= __fsub_rn(__fadd_rn(low, 1.0f), p.sb)) return +1;
    return 0;
}

// idxReg (:154-165) + hostCells (:111-151) of one point.  Returns the number of host cells, or -1 when the reference
// would index outside its grid (a coordinate outside [0, 2 l)).
__device__ __forceinline__ int bnd_point(const BndParams &p, float cx, float cy, float cz, long long out[8]) {
    float vec[3] = {__fdiv_rn(cx, p.dx), __fdiv_rn(cy, p.dx), __fdiv_rn(cz, p.dx)};
    long long idx[3];
    int reg[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        idx[k] = go_float_to_i64(truncf(vec[k]));   // Go's int(float32)
        if (idx[k] >= p.cells) {
            idx[k] -= p.cells;
            vec[k] = __fsub_rn(vec[k], p.l);        // (minus l, not minus cells: as the reference has it)
        }
        if (idx[k] < 0 || idx[k] >= p.cells) return -1;
        reg[k] = bnd_region(p, idx[k], vec[k]);
    }
    const long long c = p.cells;
    out[0] = idx[0] + idx[1] * c + idx[2] * c * c;
    int j = 1;
#pragma unroll
    for (int z = 0; z < 2; z++) {
        if (reg[2] == 0 && z == 1) continue;
#pragma unroll
        for (int y = 0; y < 2; y++) {
            if (reg[1] == 0 && y == 1) continue;
#pragma unroll
            for (int x = 0; x < 2; x++) {
                if (reg[0] == 0 && x == 1) continue;
                const int diff[3] = {x * reg[0], y * reg[1], z * reg[2]};
                if (diff[0] == 0 && diff[1] == 0 && diff[2] == 0) continue;
                long long v[3];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    v[k] = idx[k] + diff[k];
                    if (v[k] < 0) v[k] += c;
                    if (v[k] >= c) v[k] -= c;
                }
                out[j++] = v[0] + v[1] * c + v[2] * c * c;
            }
        }
    }
    return j;
}

}  // namespace

__global__ void __launch_bounds__(256) k_bnd_count(const float *x, const float *y, const float *z, int64_t n, BndParams p, int64_t *cnt,
                                                   unsigned long long *sizes, int *err) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long gs[8];
    const int ng = bnd_point(p, x[i], y[i], z[i], gs);
    if (ng < 0) { atomicExch(err, 3); cnt[i] = 0; return; }
    cnt[i] = ng;
    for (int j = 0; j < ng; j++) atomicAdd(&sizes[gs[j]], 1ULL);
}

__global__ void __launch_bounds__(256) k_bnd_emit(const float *x, const float *y, const float *z, int64_t n, BndParams p,
                                                  const int64_t *eoff, uint32_t *keys, int64_t *vals) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long gs[8];
    const int ng = bnd_point(p, x[i], y[i], z[i], gs);
    const int64_t e0 = eoff[i];
    for (int j = 0; j < ng; j++) {
        keys[e0 + j] = (uint32_t)gs[j];
        vals[e0 + j] = (i << 1) | (j == 0 ? 0 : 1);   // boundary flag: 0 in the point's own cell (:81-82)
    }
}

constexpr int RS_TILE = 4096, RS_WARPS = 8;

// table[d * ntiles + tile] = number of entries of the tile whose digit is d
__global__ void __launch_bounds__(32 * RS_WARPS) k_rs_hist(const uint32_t *keys, int64_t m, int shift, int64_t ntiles, int64_t *table) {
    __shared__ unsigned hist[RS_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = (int64_t)blockIdx.x * RS_WARPS + warp;
    for (int d = lane; d < 256; d += 32) hist[warp][d] = 0;
    __syncwarp();
    if (tile < ntiles) {
        const int64_t e0 = tile * RS_TILE, e1 = e0 + RS_TILE < m ? e0 + RS_TILE : m;
        for (int64_t e = e0 + lane; e < e1; e += 32) atomicAdd(&hist[warp][(keys[e] >> shift) & 255u], 1u);
        __syncwarp();
        for (int d = lane; d < 256; d += 32) table[(int64_t)d * ntiles + tile] = hist[warp][d];
    }
}

// stable scatter: the warp walks its tile in order, 32 entries at a time; entries with the same digit are ranked by lane
__global__ void __launch_bounds__(32 * RS_WARPS) k_rs_scatter(const uint32_t *keys, const int64_t *vals, int64_t m, int shift,
                                                              int64_t ntiles, const int64_t *table_off, uint32_t *keys_out,
                                                              int64_t *vals_out) {
    __shared__ long long base[RS_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = (int64_t)blockIdx.x * RS_WARPS + warp;
    if (tile >= ntiles) return;
    for (int d = lane; d < 256; d += 32) base[warp][d] = table_off[(int64_t)d * ntiles + tile];
    __syncwarp();
    const int64_t e0 = tile * RS_TILE, e1 = e0 + RS_TILE < m ? e0 + RS_TILE : m;
    for (int64_t eb = e0; eb < e1; eb += 32) {
        const int64_t e = eb + lane;
        const bool valid = e < e1;
        const uint32_t k = valid ? keys[e] : 0u;
        const int64_t v = valid ? vals[e] : 0;
        const unsigned d = valid ? ((k >> shift) & 255u) : (256u + (unsigned)lane);   // invalid lanes match nobody
        const unsigned mask = __match_any_sync(0xffffffffu, d);
        const int rank = __popc(mask & ((1u << lane) - 1u));
        if (valid) {
            const long long pos = base[warp][d] + rank;
            keys_out[pos] = k;
            vals_out[pos] = v;
        }
        __syncwarp();
        if (valid && rank == 0) base[warp][d] += __popc(mask);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) k_bnd_split(const int64_t *vals, int64_t m, int64_t *idx, int64_t *flags) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= m) return;
    const int64_t v = vals[e];
    idx[e] = v >> 1;
    flags[e] = v & 1;
}

// ---- launchers (api.cu drives the sequence: it needs the entry total on the host to size the buffers) ----
void launch_bnd_count(Launcher &L, const float *x, const float *y, const float *z, int64_t n, float l, float boundary, int64_t cells,
                      int64_t *cnt, int64_t *sizes, int *err) {
    BndParams p;
    p.l = l; p.cells = (int)cells;
    volatile float dx = l / (float)cells;   // go/minh/boundary.go:58
    p.dx = dx;
    volatile float sb = boundary / dx;      // :40
    p.sb = sb;
    cudaMemsetAsync(sizes, 0, 8 * (size_t)(cells * cells * cells), L.stream);
    if (n == 0) return;
    k_bnd_count<<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(x, y, z, n, p, cnt, (unsigned long long *)sizes, err);
    L.count++;
}

size_t bnd_sort_scratch_bytes(int64_t m) {
    const int64_t ntiles = (m + RS_TILE - 1) / RS_TILE;
    return 2 * 8 * 256 * (size_t)ntiles + scan_scratch_bytes(256 * ntiles) + 256;
}

// entries -> sorted (idx, flags).  keys/vals and keys2/vals2: ping-pong buffers of m entries; scratch: bnd_sort_scratch_bytes(m)
cudaError_t launch_bnd_index(Launcher &L, const float *x, const float *y, const float *z, int64_t n, float l, float boundary,
                             int64_t cells, const int64_t *eoff, int64_t m, uint32_t *keys, int64_t *vals, uint32_t *keys2,
                             int64_t *vals2, void *scratch, int64_t *idx, int64_t *flags) {
    if (n == 0 || m == 0) return cudaSuccess;
    BndParams p;
    p.l = l; p.cells = (int)cells;
    volatile float dx = l / (float)cells;
    p.dx = dx;
    volatile float sb = boundary / dx;
    p.sb = sb;
    k_bnd_emit<<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(x, y, z, n, p, eoff, keys, vals);
    L.count++;
    const int64_t c3 = cells * cells * cells, ntiles = (m + RS_TILE - 1) / RS_TILE;
    int64_t *table = (int64_t *)scratch, *table_off = table + 256 * ntiles;
    void *scan_scratch = table_off + 256 * ntiles;
    int64_t *total = (int64_t *)((char *)scan_scratch + scan_scratch_bytes(256 * ntiles));
    const unsigned grid = (unsigned)((ntiles + RS_WARPS - 1) / RS_WARPS);
    for (int shift = 0; shift < 32 && (c3 - 1) >> shift; shift += 8) {
        k_rs_hist<<<grid, 32 * RS_WARPS, 0, L.stream>>>(keys, m, shift, ntiles, table);
        L.count++;
        cudaError_t e = launch_scan_sizes(L, table, 256 * ntiles, 0, table_off, total, scan_scratch);
        if (e != cudaSuccess) return e;
        k_rs_scatter<<<grid, 32 * RS_WARPS, 0, L.stream>>>(keys, vals, m, shift, ntiles, table_off, keys2, vals2);
        L.count++;
        uint32_t *tk = keys; keys = keys2; keys2 = tk;
        int64_t *tv = vals; vals = vals2; vals2 = tv;
    }
    k_bnd_split<<<(unsigned)((m + 255) / 256), 256, 0, L.stream>>>(vals, m, idx, flags);
    L.count++;
    return cudaGetLastError();
}

}  // namespace mnw
