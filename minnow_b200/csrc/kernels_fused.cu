// kernels_fused.cu -- fused single-read encode kernels (to be filled in).
#include "fused.cuh"

namespace mnw {

bool fused_group_supported(const FloatParamsHost &, int64_t, int64_t) { return false; }
cudaError_t launch_fused_group(Launcher &, void *, size_t, const FloatParamsHost &, const float *, int64_t, int64_t,
                               int64_t *, int64_t *, int64_t *, int64_t *, uint8_t *, int64_t, int *) {
    return cudaErrorNotSupported;
}
bool fused_vec3_supported(const FloatParamsHost *, int64_t, int, int) { return false; }
cudaError_t launch_fused_vec3(Launcher &, const FloatParams *, int, const float *, int, int, int64_t, int64_t *,
                              int64_t *, int64_t *, int64_t *, uint8_t *, int64_t, int *) {
    return cudaErrorNotSupported;
}

}  // namespace mnw
