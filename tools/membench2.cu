// membench2.cu -- starts from the read probe that reaches 7 TB/s on the minp tile pattern and adds
// the pieces of the encode A-phase one at a time, to see which one costs the time.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int NFILE = 256, NSUB = 64, S = 4;
constexpr unsigned ROW4 = 3 * NFILE / 4, PLANE4 = ROW4 * NFILE;
constexpr size_t FILE4 = (size_t)PLANE4 * NFILE;

__device__ __forceinline__ const float4 *tile_base(const float4 *aos, unsigned tileid) {
    unsigned unit = tileid / 64, pl = tileid % 64;
    unsigned f = unit / 64, sc = unit % 64;
    unsigned ix0 = NSUB * (sc % S), iy0 = NSUB * ((sc / S) % S), iz0 = NSUB * (sc / (S * S));
    return aos + f * FILE4 + (3u * ix0 / 4u + iy0 * ROW4 + (iz0 + pl) * PLANE4);
}
__global__ void fill(float *p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = (float)((i * 2654435761ull) % 1000000ull) * 0.001f;
}
__device__ __forceinline__ int qfast(float v, float low, float rcp, float ndx) {
    float t = __fsub_rn(v, low), y = __fmul_rn(t, rcp);
    float e = __fmaf_rn(ndx, y, t); y = __fmaf_rn(e, rcp, y);
    e = __fmaf_rn(ndx, y, t); y = __fmaf_rn(e, rcp, y);
    return __float2int_rd(y);
}
// LEVEL 0: xor only; 1: + quantise; 2: + rotate + min/max3; 3: + STS.U16 staging; 4: + tile-end reduce + atomics + fence
template <int LEVEL>
__global__ void __launch_bounds__(192, 4) probe(const float4 *aos, unsigned ntiles, unsigned *ticket, unsigned *out, float low, float rcp, float ndx, int P, unsigned C, unsigned short *scratch) {
    __shared__ unsigned s_t;
    __shared__ __align__(16) unsigned short stage[3 * 4096];
    __shared__ unsigned s_red[6][4];
    unsigned acc = 0;
    const int tid = threadIdx.x, col4 = tid % 48, rsub = tid / 48, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        if (tid == 0) s_t = atomicAdd(ticket, 1u);
        __syncthreads();
        unsigned t = s_t;
        __syncthreads();
        if (t >= ntiles) break;
        const float4 *b = tile_base(aos, t) + col4 + rsub * ROW4;
        int qmn = 0x7fffffff, qmx = -1; unsigned wmn = ~0u, wmx = 0;
        for (int p0 = 0; p0 < 16; p0 += 2) {
            float4 v[2];
#pragma unroll
            for (int u = 0; u < 2; u++) v[u] = __ldcs(b + (size_t)(p0 + u) * 4 * ROW4);
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const float x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                if (LEVEL == 0) { acc ^= __float_as_uint(x[0]) ^ __float_as_uint(x[1]) ^ __float_as_uint(x[2]) ^ __float_as_uint(x[3]); continue; }
                int q[4]; unsigned w[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    q[c] = qfast(x[c], low, rcp, ndx);
                    if (LEVEL >= 2) { unsigned tt = (unsigned)q[c] + C; w[c] = min(tt, tt - (unsigned)P); }
                    if (LEVEL >= 3) stage[(c % 3) * 4096 + (p0 + u) * 256 + rsub * 64 + (4 * col4 + c) / 3] = (unsigned short)w[c];
                }
                if (LEVEL == 1) acc ^= q[0] ^ q[1] ^ q[2] ^ q[3];
                if (LEVEL >= 2) {
                    qmn = __vimin3_s32(qmn, q[0], q[1]); qmn = __vimin3_s32(qmn, q[2], q[3]);
                    qmx = __vimax3_s32(qmx, q[0], q[1]); qmx = __vimax3_s32(qmx, q[2], q[3]);
                    wmn = __vimin3_u32(wmn, w[0], w[1]); wmn = __vimin3_u32(wmn, w[2], w[3]);
                    wmx = __vimax3_u32(wmx, w[0], w[1]); wmx = __vimax3_u32(wmx, w[2], w[3]);
                }
            }
        }
        if (LEVEL >= 2) acc ^= qmn ^ qmx ^ wmn ^ wmx;
        if (LEVEL >= 4) {
            unsigned a = __reduce_min_sync(0xffffffffu, wmn), bb = __reduce_max_sync(0xffffffffu, wmx);
            int c = __reduce_min_sync(0xffffffffu, qmn), d = __reduce_max_sync(0xffffffffu, qmx);
            if (lane == 0) { s_red[warp][0] = ~a; s_red[warp][1] = bb; s_red[warp][2] = ~(unsigned)c; s_red[warp][3] = (unsigned)d; }
            __syncthreads();
            if (tid < 4) { unsigned m = 0; for (int w = 0; w < 6; w++) m = max(m, s_red[w][tid]); atomicMax(out + 16 + (t / 64) * 4 + tid, m); }
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicAdd(out + 8, 1u);
        }
        if (LEVEL >= 5) {   // park the staged indices in a 24 MB ring (16 units x 1.5 MB), planar per axis
            unsigned short *scr = scratch + (size_t)((t / 64) % 16) * 3 * 262144 + (size_t)(t % 64) * 4096;
            for (int i = tid; i < 3 * 4096 / 8; i += 192) {
                const int k = i / 512, wdx = i - k * 512;
                __stcg((uint4 *)(scr + (size_t)k * 262144) + wdx, ((const uint4 *)(stage + k * 4096))[wdx]);
            }
            __threadfence();
            __syncthreads();
        }
        if (LEVEL >= 6 && t >= 64 * 8) {   // read back the tile parked 8 units ago (no dependency tracking)
            const unsigned tb = t - 64 * 8;
            const unsigned short *scr = scratch + (size_t)((tb / 64) % 16) * 3 * 262144 + (size_t)(tb % 64) * 4096;
            for (int g = warp; g < 12; g += 6) {
                const int k = g / 4, gi = g % 4;
                const uint4 *src = (const uint4 *)(scr + (size_t)k * 262144 + gi * 1024 + 32 * lane);
                uint4 r[4];
#pragma unroll
                for (int s4 = 0; s4 < 4; s4++) r[s4] = __ldcg(src + s4);
#pragma unroll
                for (int s4 = 0; s4 < 4; s4++) acc ^= r[s4].x ^ r[s4].y ^ r[s4].z ^ r[s4].w;
            }
        }
    }
    if (acc == 0x12345678u) out[0] = acc + stage[tid];
}

int main() {
    const int nfiles = 16;
    const size_t n4 = nfiles * FILE4;
    float4 *aos; unsigned *ticket, *out;
    CK(cudaMalloc(&aos, n4 * 16));
    fill<<<148 * 8, 256>>>((float *)aos, n4 * 4);
    CK(cudaMalloc(&ticket, 4));
    CK(cudaMalloc(&out, 1 << 20));
    CK(cudaMemset(out, 0, 1 << 20));
    const unsigned ntiles = nfiles * 64 * 64;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const float dx = 1000.0f / 200000.0f, rcp = 1.0f / dx;
    auto timeit = [&](const char *name, auto fn) {
        float best = 1e9;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaMemset(ticket, 0, 4));
            cudaEventRecord(a); fn(); cudaEventRecord(b);
            CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%-52s %7.3f ms  %7.1f GB/s\n", name, best, n4 * 16 / best / 1e6);
    };
    unsigned short *scratch; CK(cudaMalloc(&scratch, (size_t)16 * 3 * 262144 * 2));
#define RUN(L, name) timeit(name, [&] { probe<L><<<148 * 4, 192>>>(aos, ntiles, ticket, out, 0.0f, rcp, -dx, 200000, 1234u, scratch); });
    RUN(0, "L0 read + xor");
    RUN(1, "L1 + quantise (FSUB FMUL 4xFFMA F2I)");
    RUN(2, "L2 + rotate + 3-input min/max");
    RUN(3, "L3 + STS.U16 staging");
    RUN(4, "L4 + tile-end reduce, atomics, fence, counter");
    RUN(5, "L5 + park 24 KB per tile in a 24 MB ring");
    RUN(6, "L6 + read the tile parked 8 units earlier");
    return 0;
}
