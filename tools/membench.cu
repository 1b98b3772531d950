// membench.cu -- read-bandwidth probes for the minp access pattern (not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench membench.cu && ./membench
// Layout: nfiles cubes of 256^3 particles x 3 floats (AoS); a "tile" is one z-plane of a 64^3
// sub-cell = 64 rows of 768 contiguous bytes, 3072 bytes apart.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int NFILE = 256, NSUB = 64, S = 4;
constexpr unsigned ROW4 = 3 * NFILE / 4, PLANE4 = ROW4 * NFILE;
constexpr size_t FILE4 = (size_t)PLANE4 * NFILE;

__device__ __forceinline__ const float4 *tile_base(const float4 *aos, unsigned tileid) {
    // tileid -> (file, sub-cell, plane)
    unsigned unit = tileid / 64, pl = tileid % 64;
    unsigned f = unit / 64, sc = unit % 64;
    unsigned ix0 = NSUB * (sc % S), iy0 = NSUB * ((sc / S) % S), iz0 = NSUB * (sc / (S * S));
    return aos + f * FILE4 + (3u * ix0 / 4u + iy0 * ROW4 + (iz0 + pl) * PLANE4);
}

// contiguous streaming read
template <int U>
__global__ void read_contig(const float4 *p, size_t n4, unsigned *out) {
    unsigned acc = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = __ldcs(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; u++) acc ^= __float_as_uint(v[u].x) ^ __float_as_uint(v[u].y) ^ __float_as_uint(v[u].z) ^ __float_as_uint(v[u].w);
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// tile pattern, 192 threads: thread (rsub = tid/48, col4 = tid%48), 16 passes of 4 rows, DEPTH loads in flight
template <int DEPTH>
__global__ void __launch_bounds__(192) read_tiles(const float4 *aos, unsigned ntiles, unsigned *ticket, unsigned *out) {
    __shared__ unsigned s_t;
    unsigned acc = 0;
    const int tid = threadIdx.x, col4 = tid % 48, rsub = tid / 48;
    for (;;) {
        if (tid == 0) s_t = atomicAdd(ticket, 1u);
        __syncthreads();
        unsigned t = s_t;
        __syncthreads();
        if (t >= ntiles) break;
        const float4 *b = tile_base(aos, t) + col4 + rsub * ROW4;
        for (int p0 = 0; p0 < 16; p0 += DEPTH) {
            float4 v[DEPTH];
#pragma unroll
            for (int u = 0; u < DEPTH; u++) v[u] = __ldcs(b + (size_t)(p0 + u) * 4 * ROW4);
#pragma unroll
            for (int u = 0; u < DEPTH; u++) acc ^= __float_as_uint(v[u].x) ^ __float_as_uint(v[u].y) ^ __float_as_uint(v[u].z) ^ __float_as_uint(v[u].w);
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// tile pattern through TMA bulk copies: whole tile (48 KB) per mbarrier, double buffered
__global__ void __launch_bounds__(192) read_tiles_tma(const float4 *aos, unsigned ntiles, unsigned *ticket, unsigned *out) {
    extern __shared__ __align__(128) unsigned char sm[];   // 2 x 49152
    __shared__ __align__(8) unsigned long long bar[2];
    __shared__ unsigned s_t[2];
    const int tid = threadIdx.x;
    unsigned acc = 0;
    auto sa = [](const void *p) { return (unsigned)__cvta_generic_to_shared(p); };
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int st) {   // warp 0: claim a ticket and fetch its tile
        unsigned t = 0;
        if (tid == 0) { t = atomicAdd(ticket, 1u); s_t[st] = t; }
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= ntiles) return;
        const float4 *b = tile_base(aos, t);
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(&bar[st])), "r"(49152) : "memory");
        __syncwarp();
        for (int r = tid; r < 64; r += 32)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sa(sm + st * 49152 + r * 768)), "l"(b + (size_t)r * ROW4), "r"(768), "r"(sa(&bar[st])) : "memory");
    };
    if (tid < 32) issue(0);
    __syncthreads();
    for (int it = 0;; it++) {
        const int st = it & 1;
        if (tid < 32) issue(st ^ 1);
        const unsigned t = s_t[st];
        if (t >= ntiles) break;
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(sa(&bar[st])), "r"((it >> 1) & 1) : "memory");
        const float4 *s4 = (const float4 *)(sm + st * 49152);
#pragma unroll 4
        for (int i = tid; i < 3072; i += 192) {
            float4 v = s4[i];
            acc ^= __float_as_uint(v.x) ^ __float_as_uint(v.y) ^ __float_as_uint(v.z) ^ __float_as_uint(v.w);
        }
        __syncthreads();
    }
    if (acc == 0x12345678u) out[0] = acc;
}

int main() {
    const int nfiles = 16;
    const size_t n4 = nfiles * FILE4;
    float4 *aos;
    unsigned *ticket, *out;
    CK(cudaMalloc(&aos, n4 * 16));
    CK(cudaMemset(aos, 1, n4 * 16));
    CK(cudaMalloc(&ticket, 4));
    CK(cudaMalloc(&out, 4));
    const unsigned ntiles = nfiles * 64 * 64;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    auto timeit = [&](const char *name, auto fn) {
        float best = 1e9;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaMemset(ticket, 0, 4));
            cudaEventRecord(a);
            fn();
            cudaEventRecord(b);
            CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%-44s %7.3f ms  %7.1f GB/s\n", name, best, n4 * 16 / best / 1e6);
    };
    timeit("contig U=4, 148x8 CTAs x 256", [&] { read_contig<4><<<148 * 8, 256>>>(aos, n4, out); });
    timeit("contig U=8, 148x8 CTAs x 256", [&] { read_contig<8><<<148 * 8, 256>>>(aos, n4, out); });
    timeit("contig U=4, 148x4 CTAs x 192", [&] { read_contig<4><<<148 * 4, 192>>>(aos, n4, out); });
    timeit("tiles depth 2, 148x4 CTAs x 192", [&] { read_tiles<2><<<148 * 4, 192>>>(aos, ntiles, ticket, out); });
    timeit("tiles depth 4, 148x4 CTAs x 192", [&] { read_tiles<4><<<148 * 4, 192>>>(aos, ntiles, ticket, out); });
    timeit("tiles depth 8, 148x4 CTAs x 192", [&] { read_tiles<8><<<148 * 4, 192>>>(aos, ntiles, ticket, out); });
    timeit("tiles depth 16, 148x4 CTAs x 192", [&] { read_tiles<16><<<148 * 4, 192>>>(aos, ntiles, ticket, out); });
    timeit("tiles depth 16, 148x8 CTAs x 192", [&] { read_tiles<16><<<148 * 8, 192>>>(aos, ntiles, ticket, out); });
    CK(cudaFuncSetAttribute(read_tiles_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
    timeit("tiles TMA 2x48KB, 148x2 CTAs x 192", [&] { read_tiles_tma<<<148 * 2, 192, 98304>>>(aos, ntiles, ticket, out); });
    timeit("tiles TMA 2x48KB, 148x1 CTAs x 192", [&] { read_tiles_tma<<<148, 192, 98304>>>(aos, ntiles, ticket, out); });
    return 0;
}
