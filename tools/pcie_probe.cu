// pcie_probe.cu -- the host<->device copy ceiling the end-to-end number of bench.py is bounded by.
//
//   nvcc -O2 -o tools/pcie_probe tools/pcie_probe.cu && tools/pcie_probe [device] [MiB per copy] [repeats]
//
// Pinned host memory, one cudaMemcpyAsync per buffer, CUDA events on the copy streams:
//   h2d     one direction at a time
//   d2h
//   duplex  both directions at once on two streams (what a pipelined encode+decode step needs:
//           particles and packed bytes go down while packed bytes and decoded floats come back)
// Prints one JSON line.  With several processes at once (one per GPU) it measures the NODE's ceiling.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char **argv) {
    const int dev = argc > 1 ? atoi(argv[1]) : 0;
    const size_t mib = argc > 2 ? (size_t)atoll(argv[2]) : 512;
    const int reps = argc > 3 ? atoi(argv[3]) : 8;
    const size_t n = mib << 20;
    CK(cudaSetDevice(dev));
    void *h_a, *h_b, *d_a, *d_b;
    CK(cudaMallocHost(&h_a, n)); CK(cudaMallocHost(&h_b, n));
    CK(cudaMalloc(&d_a, n)); CK(cudaMalloc(&d_b, n));
    for (size_t i = 0; i < n; i += 4096) { ((char *)h_a)[i] = 1; ((char *)h_b)[i] = 2; }
    cudaStream_t s0, s1;
    CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    cudaEvent_t a0, b0, a1, b1;
    CK(cudaEventCreate(&a0)); CK(cudaEventCreate(&b0)); CK(cudaEventCreate(&a1)); CK(cudaEventCreate(&b1));
    float ms = 0, ms1 = 0;
    double h2d, d2h, dup_h2d, dup_d2h, dup;
    // warm-up
    CK(cudaMemcpyAsync(d_a, h_a, n, cudaMemcpyHostToDevice, s0)); CK(cudaMemcpyAsync(h_b, d_b, n, cudaMemcpyDeviceToHost, s1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a0, s0));
    for (int r = 0; r < reps; r++) CK(cudaMemcpyAsync(d_a, h_a, n, cudaMemcpyHostToDevice, s0));
    CK(cudaEventRecord(b0, s0)); CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&ms, a0, b0));
    h2d = (double)n * reps / ms / 1e6;
    CK(cudaEventRecord(a1, s1));
    for (int r = 0; r < reps; r++) CK(cudaMemcpyAsync(h_b, d_b, n, cudaMemcpyDeviceToHost, s1));
    CK(cudaEventRecord(b1, s1)); CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&ms, a1, b1));
    d2h = (double)n * reps / ms / 1e6;
    CK(cudaEventRecord(a0, s0)); CK(cudaEventRecord(a1, s1));
    for (int r = 0; r < reps; r++) {
        CK(cudaMemcpyAsync(d_a, h_a, n, cudaMemcpyHostToDevice, s0));
        CK(cudaMemcpyAsync(h_b, d_b, n, cudaMemcpyDeviceToHost, s1));
    }
    CK(cudaEventRecord(b0, s0)); CK(cudaEventRecord(b1, s1)); CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, a0, b0)); CK(cudaEventElapsedTime(&ms1, a1, b1));
    dup_h2d = (double)n * reps / ms / 1e6; dup_d2h = (double)n * reps / ms1 / 1e6;
    dup = 2.0 * n * reps / (ms > ms1 ? ms : ms1) / 1e6;
    printf("{\"device\": %d, \"mib_per_copy\": %zu, \"repeats\": %d, \"h2d_gbs\": %.2f, \"d2h_gbs\": %.2f, "
           "\"duplex_h2d_gbs\": %.2f, \"duplex_d2h_gbs\": %.2f, \"duplex_total_gbs\": %.2f}\n",
           dev, mib, reps, h2d, d2h, dup_h2d, dup_d2h, dup);
    return 0;
}
