// fused_detail.cuh -- device helpers shared by the fused minp encode kernels
// (kernels_fused.cu: k_fused_vec3, kernels_pipe.cu: k_pipe_vec3).
// Everything sits in an anonymous namespace: each translation unit gets its own copy.
#pragma once
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>

#include "device_math.cuh"
#include "engine.cuh"
#include "f32x2.cuh"
#include "fused.cuh"
#include "launch.cuh"
#include "pack.cuh"

namespace cg = cooperative_groups;

namespace mnw {

struct XStat {  // one CTA's statistics of one axis block
    unsigned wmin, wmax;
    int qmin, qmax;
    unsigned oob, pad0, pad1, pad2;
};

struct Fin {  // finalised block, identical in every CTA of the cluster
    long long off;    // exclusive byte offset of the block in its group
    int bits;
    int mode;         // 1: pack from shared memory, 0: nothing to pack here
    unsigned base;    // v = w - base, + padj when negative
    unsigned padj;
};

struct FusedArgs {
    const BlockDesc *descs;   // the batch's block descriptors (k_build_vec3): the exact path of a unit with out-of-range values
    const float *aos;
    const FloatParams *tab;
    int tab_per_file;
    int nfile, subcells;
    long long nunits, sc3;
    BlockStat *stats;
    int64_t *mins, *bits, *offsets, *out_len;
    uint8_t *out;
    long long axis_stride;
    int prefetch;   // pull the next unit's rows into L2 during the pack phase
    int ticket_chunk;  // k_fused_vec3: consecutive sub-cells per step of the round-robin ticket order (claim_unit)
    unsigned *ustat;   // k_pipe_vec3<COOP>: [nunits][parts][16] u64 statistics records in global memory (zeroed before the launch)
    FusedWork W;
};

namespace {

constexpr unsigned long long PUB_AGG = 1ULL << 62, PUB_PREFIX = 2ULL << 62, PUB_VALUE = (1ULL << 62) - 1ULL;

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Exclusive prefix of the published byte sizes of blocks [first, b): the
// decoupled look-back of a chained scan, 32 predecessors per step.
__device__ long long lookback(const unsigned long long *pub, long long first, long long b) {
    const int lane = threadIdx.x & 31;
    long long sum = 0;
    for (long long hi = b; hi > first; hi -= 32) {
        const long long idx = hi - 1 - lane;
        const bool valid = idx >= first;
        unsigned long long v = 0;
        if (valid) {
            while (((v = ld_relaxed(pub + idx)) >> 62) == 0) __nanosleep(100);
        }
        const unsigned pmask = __ballot_sync(0xffffffffu, valid && (v >> 62) == 2);
        long long val = valid ? (long long)(v & PUB_VALUE) : 0;
        if (pmask) {  // the nearest predecessor with an inclusive prefix ends the walk
            const int stop = __ffs(pmask) - 1;
            if (lane > stop) val = 0;
        }
        for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        sum += val;
        if (pmask) break;
    }
    return sum;
}

// Streaming 128-bit load that the compiler may not move across other memory operations: the
// software pipeline below relies on the loads of the NEXT batch being issued before the
// current batch is processed (nvcc otherwise sinks them to save registers).
__device__ __forceinline__ float4 ld_stream_pinned(const float4 *p) {
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// Exact lane path of the quantiser: anything the fast quotient could not vouch for.
// Returns the folded pixel index (pixels -> 0) or flags the element out of range.
__device__ __noinline__ int quantize_rare(float x, float low, float dx, int P, unsigned &oob) {
    long long q = quantize_exact(x, low, dx);
    if (q == (long long)P) return 0;
    if ((unsigned long long)q < (unsigned long long)P) return (int)q;
    oob = 1;
    return 0;
}

// One pack group = 1024 consecutive elements of one block = 32 lanes x 32 values ->
// 32*B words.  The words are transposed in place through the group's own 2 KiB of
// staging (XOR swizzle: conflict-free both ways) so that the warp can write them out in
// stream order, 128 bytes per store instruction.
template <int B>
__device__ __forceinline__ void pack_group_words(const unsigned (&v)[32], unsigned *region, int lane, uint8_t *dst0,
                                                 const long long *off, const int *offgen, int gen) {
    unsigned o[B];
    pack32<B>(v, o);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < B; j++) {
        const int W = lane * B + j;
        region[W ^ (W >> 5)] = o[j];
    }
    // the block's byte offset is posted by the look-back warp of its axis
    while (*(const volatile int *)offgen != gen) { }
    __syncwarp();
    const long long o64 = *(const volatile long long *)off;
    if (o64 >= 0) write_group<B>(dst0 + o64, region, lane);
}

// Cluster-wide barrier with release/acquire at cluster scope: all it has to order are the
// distributed-shared-memory stores of the statistics exchange.
template <int CS>
__device__ __forceinline__ void cluster_sync_all() {
    if constexpr (CS > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
}

}  // namespace

}  // namespace mnw
