// conv_tc.cu -- BF16 tcgen05 implicit-GEMM convolution (placeholder dispatch: until the
// tensor-core kernels land, no shape is claimed and every call takes the FP32 SIMT path).
#include "common.cuh"
#include "conv_impl.cuh"

namespace b200 {

bool conv_tc_supports_fprop(const bcnn_b200_conv_desc *) { return false; }
bool conv_tc_supports_dgrad(const bcnn_b200_conv_desc *) { return false; }
bool conv_tc_supports_wgrad(const bcnn_b200_conv_desc *) { return false; }
size_t conv_tc_workspace_bytes(const bcnn_b200_conv_desc *) { return 0; }
int conv_tc_forward(const bcnn_b200_conv_desc *, const float *, const float *, const float *, int,
                    float *, void *, size_t, cudaStream_t) { return (int)cudaErrorNotSupported; }
int conv_tc_backward_data(const bcnn_b200_conv_desc *, const float *, const float *, float *, int,
                          void *, size_t, cudaStream_t) { return (int)cudaErrorNotSupported; }
int conv_tc_backward_weights(const bcnn_b200_conv_desc *, const float *, const float *, float *,
                             void *, size_t, cudaStream_t) { return (int)cudaErrorNotSupported; }

}  // namespace b200
