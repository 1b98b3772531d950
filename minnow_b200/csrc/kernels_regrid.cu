// kernels_regrid.cu -- Lagrangian re-gridding (go/minp/snapshot/grid.go:118-137 grid.Index, :206-211 vectorGrid.Insert,
// :168-204 xGrid / vGrid): the producer of minp.Writer.Vectors' input.  Particle j of a snapshot file carries a 1-based
// ID; ID - 1 = idx + idy * nAll + idz * nAll^2 is its place on the nAll^3 Lagrangian lattice (nAll = ncell * nside), which
// is cut into ncell^3 cells of nside^3 particles: cell c = (cx, cy, cz), slot i = (ix, iy, iz) inside it, x fastest.
// The grid on the device is laid out [cell][slot][3] float32 -- exactly the [nfiles][nfile^3][3] AoS input of
// mnw_minp_encode_vectors_dev with FileCells = ncell, nfile = nside -- so a snapshot goes from its files to minp bytes
// without leaving the GPU.  A pure scatter: HBM-bound, 8 + 12 bytes read and 12 bytes written per particle.
#include "engine.cuh"
#include "launch.cuh"

namespace mnw {

__global__ void __launch_bounds__(256) k_regrid_insert(const int64_t *__restrict__ ids, const float *__restrict__ vec, int64_t n,
                                                       long long ncell, long long nside, float *grid, int *err) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const long long nall = ncell * nside, id = ids[j] - 1;   // xGrid: grid.Insert(id[j] - 1, x[j])
    if (id < 0 || id >= nall * nall * nall) { atomicExch(err, 4); return; }   // grid.Index panics
    const long long idx = id % nall, idy = (id / nall) % nall, idz = id / (nall * nall);
    const long long ix = idx % nside, iy = idy % nside, iz = idz % nside;
    const long long cx = idx / nside, cy = idy / nside, cz = idz / nside;
    const long long i = ix + iy * nside + iz * nside * nside;
    const long long c = cx + cy * ncell + cz * ncell * ncell;
    float *dst = grid + 3 * (c * nside * nside * nside + i);
    dst[0] = vec[3 * j]; dst[1] = vec[3 * j + 1]; dst[2] = vec[3 * j + 2];
}

void launch_regrid_insert(Launcher &L, const int64_t *ids, const float *vec, int64_t n, int64_t ncell, int64_t nside, float *grid, int *err) {
    if (n == 0) return;
    L.begin("k_regrid_insert");
    k_regrid_insert<<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(ids, vec, n, ncell, nside, grid, err);
    L.end();
    L.count++;
}

}  // namespace mnw
