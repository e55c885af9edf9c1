// tcgen05 TF32 implicit-GEMM convolution (placeholder until the tensor-core kernel lands).
#include "common.cuh"
namespace cg {
bool tc_fwd_supported(const cg_conv_geom&) { return false; }
int tc_conv_fwd(const cg_conv_geom&, const float*, const float*, const float*, float*, int, float, void*, size_t, cudaStream_t) {
    set_error("tensor-core path not built");
    return CG_ERR_ARG;
}
}  // namespace cg
