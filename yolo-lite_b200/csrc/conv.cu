// C-ABI dispatch for the dense convolution: tcgen05 implicit GEMM when the problem fits, otherwise the
// direct CUDA-core kernel (3-channel stem, odd sizes).  Both are this library's own sm_100a kernels;
// there is no library or CPU fallback.
#include "common.cuh"

namespace yl {
bool conv_tc_supported(const yl_conv_args* a, char* why, size_t why_len);
int launch_conv_tc(const yl_conv_args* a, cudaStream_t stream);
int launch_conv_direct(const yl_conv_args* a, cudaStream_t stream);
int conv_tc_info(const yl_conv_args* a, yl_conv_tc_plan* out);
int launch_conv_b2b(const yl_conv_args* a, const yl_conv_args* head, cudaStream_t stream);
bool conv_b2b_supported(const yl_conv_args* a, const yl_conv_args* head);
}  // namespace yl

extern "C" {

int yl_conv_tc_supported(const yl_conv_args* a) {
    if (!a) return 0;
    return yl::conv_tc_supported(a, nullptr, 0) ? 1 : 0;
}

int yl_conv_tc_info(const yl_conv_args* a, yl_conv_tc_plan* out) {
    YL_CHECK(a != nullptr && out != nullptr, YL_ERR_ARG, "null argument");
    return yl::conv_tc_info(a, out);
}

int yl_conv_b2b_det_supported(const yl_conv_args* conv, const yl_conv_args* head) {
    return conv && head && yl::conv_b2b_supported(conv, head) ? 1 : 0;
}

int yl_conv_b2b_det(const yl_conv_args* conv, const yl_conv_args* head, void* stream) {
    YL_CHECK(conv && head && conv->x.data && conv->w && conv->bias && head->w && head->bias && head->det.pred, YL_ERR_ARG,
             "null pointer");
    return yl::launch_conv_b2b(conv, head, (cudaStream_t)stream);
}

int yl_conv_bn_act(const yl_conv_args* a, void* stream) {
    YL_CHECK(a != nullptr, YL_ERR_ARG, "null conv args");
    YL_CHECK(a->x.data && (a->y.data || a->det.pred) && a->w && a->bias, YL_ERR_ARG, "null tensor pointer");
    YL_CHECK(a->x.c > 0 && a->y.c > 0 && a->x.n > 0 && a->x.h > 0 && a->x.w > 0, YL_ERR_ARG, "empty tensor");
    YL_CHECK(a->x.coff + a->x.c <= a->x.cstride && a->y.coff + a->y.c <= a->y.cstride, YL_ERR_ARG,
             "channel slice exceeds buffer");
    if (a->res.data) {
        YL_CHECK(a->res.c == a->y.c && a->res.coff + a->res.c <= a->res.cstride, YL_ERR_ARG, "residual slice mismatch");
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (a->det.pred) {
        YL_CHECK(a->impl != YL_IMPL_DIRECT, YL_ERR_UNSUPPORTED, "the Detect-decode epilogue exists on the tcgen05 path only");
        return yl::launch_conv_tc(a, s);
    }
    if (a->impl == YL_IMPL_DIRECT) return yl::launch_conv_direct(a, s);
    if (a->impl == YL_IMPL_TCGEN05) return yl::launch_conv_tc(a, s);
    if (yl::conv_tc_supported(a, nullptr, 0)) return yl::launch_conv_tc(a, s);
    return yl::launch_conv_direct(a, s);
}

}  // extern "C"
