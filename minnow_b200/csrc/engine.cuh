// engine.cuh -- internal device-side data model of libminnow_b200.
//
// A call to the library encodes (or decodes) a BATCH of minnow blocks.  Each
// block is described by a BlockDesc (where its elements live, how to turn them
// into integers) and gets a BlockStat (what intGroup/floatGroup.writeData would
// have derived for it: min, bits, byte size, byte offset).  Blocks are grouped
// into CHAINS: a chain is one minnow group, i.e. one contiguous byte stream
// with its own running block offset (go/block_index.go).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mnw {

enum : int32_t { KIND_I64 = 0, KIND_F32 = 1 };
enum : int32_t { ACC_CONTIG = 0, ACC_GATHER = 1, ACC_SUBCELL = 2 };
enum : int32_t { F_PERIODIC = 1, F_LOG10 = 2, F_CLAMP = 4, F_FASTDIV = 8 };

struct __align__(16) BlockDesc {
    const void *src;      // CONTIG: first element; GATHER: column base; SUBCELL: AoS cube base
    const int64_t *idx;   // GATHER: this block's index list
    int64_t n;            // elements in the block
    int64_t pixels;       // floatGroup.pixels
    float low, high;      // floatGroup.low/high
    float dx;             // (high - low) / float32(pixels), go/group.go:316
    float hi_clamp;       // Nextafter32(high, -Inf), go/minh/minh.go:146
    int32_t kind, access, flags, chain;
    int32_t nfile, nsub, ix0, iy0, iz0, axis;  // SUBCELL geometry, go/minp/minp.go:246-264
    int64_t tile0;        // first pack tile of this block
    int64_t chunk0;       // first stats chunk of this block
};

struct __align__(16) BlockStat {
    // accumulated by k_stats (atomics)
    unsigned long long wmin, wmax;  // min/max rotated periodic coordinate
    long long qmin, qmax;           // min/max integer value
    long long q0;                   // integer value of element 0
    unsigned int oob;               // some q outside [0, pixels)
    int slow;                       // needs the exact sequential periodicMin
    // finalised
    long long pmin;                 // periodicMin result (origin of bound())
    long long min;                  // what the reference appends to g.mins
    long long nbytes;               // ArrayBytes(bits, n)
    long long out_off;              // blockOffset within the chain
    int do_bound;                   // apply bound(q, pmin, pixels)
    int bits;                       // what the reference appends to g.bits
};

// Parameters of one FloatGroup as the kernels use them (device table entry).
struct FloatParams {
    float low, high, dx, hi_clamp;
    int64_t pixels;
    int32_t flags;   // F_* bits; F_FASTDIV: rcp may replace the IEEE divide (device_math.cuh quantize_fast)
    float rcp;       // RN(1 / dx)
};

// The context's 16 device flag words: [0] slow-block count, [2] abort flag (fused minp encode), [3] repack / wide block
// count, [4] ticket, [FLAG_ERR] the sticky ERROR word (1: value range 2^64-1, 2: packed output does not fit), which
// only mnw_sync / the host-pointer entry points read and clear; every call zeroes the words before it.
constexpr int FLAG_ERR = 15;

// mnw_float_desc (include/minnow_cuda.h) as the kernels see it: same 24 bytes.
struct FloatDescPod {
    float low, high;
    int64_t pixels;
    uint8_t periodic, log10, clamp, reserved[5];
};

constexpr int STATS_THREADS = 256;
constexpr int STATS_CHUNK = 16384;   // elements per k_stats CTA
constexpr int PACK_THREADS = 128;
constexpr int PACK_TILE = 4096;      // elements per k_pack CTA (32 per thread)
constexpr int DEC_THREADS = 256;
constexpr int DEC_CHUNK = 4096;      // elements per k_decode CTA

// How a batch maps CTAs to blocks.
struct BatchShape {
    int64_t nblocks;
    int64_t nchains;
    int64_t blocks_per_chain;      // chains are runs of this many consecutive blocks
    int64_t uniform_n;             // > 0: every block has this many elements
    int64_t total_tiles, total_chunks;
};

}  // namespace mnw
