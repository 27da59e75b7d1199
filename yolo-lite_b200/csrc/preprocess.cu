// GPU image preprocess: LetterBox + BGR->RGB + HWC->CHW + float + /255 in ONE kernel, reading
// the raw uint8 images (4x fewer PCIe bytes than the fp32 batch the reference uploads).
//
// Replaces, bit for bit: data/augment.py:612-681 (LetterBox: cv2.resize INTER_LINEAR + cv2.copyMakeBorder 114) and
// engine/predictor.py:67-85 (stack, [..., ::-1], transpose, .float(), /= 255).  The bilinear arithmetic is cv2's
// 8-bit fixed-point algorithm (imgproc/resize.cpp): coordinates in double -> float, 11-bit coefficients rounded
// half-to-even, horizontal pass in int32, vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2.
// Compiled WITHOUT fast-math; every floating-point step is an explicit round-to-nearest intrinsic so nvcc cannot
// contract a multiply-add into an FMA (the x86 baseline build of cv2 does not).
#include "common.cuh"

namespace yl {

struct LbCoef {
    int i0, i1;   // the two source indices
    int c0, c1;   // 11-bit fixed-point weights
};

// cv2 horizontal rule: sx < 0 -> (0, f = 0); sx >= ssize - 1 -> (ssize - 1, f = 0).
// cv2 vertical rule  : row indices are clamped, the weights are kept.
__device__ __forceinline__ LbCoef lb_coef(int d, int ssize, int dsize, bool horizontal) {
    const double inv_scale = __ddiv_rn((double)dsize, (double)ssize);
    const double scale = __ddiv_rn(1.0, inv_scale);
    float f = __double2float_rn(__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5));
    int s = __float2int_rd(f);
    f = __fsub_rn(f, (float)s);
    LbCoef c;
    if (horizontal) {
        if (s < 0) {
            s = 0;
            f = 0.f;
        }
        if (s >= ssize - 1) {
            s = ssize - 1;
            f = 0.f;
        }
        c.i0 = s;
        c.i1 = min(s + 1, ssize - 1);
    } else {
        c.i0 = min(max(s, 0), ssize - 1);
        c.i1 = min(max(s + 1, 0), ssize - 1);
    }
    c.c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    c.c1 = __float2int_rn(__fmul_rn(f, 2048.f));
    return c;
}

__global__ void __launch_bounds__(256) letterbox_u8_kernel(const yl_lb_image* __restrict__ imgs, float* __restrict__ dst,
                                                           int H, int W, int pad_value) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const yl_lb_image im = imgs[blockIdx.z];
    const long long plane = (long long)H * W;
    float* out = dst + (long long)blockIdx.z * 3 * plane + (long long)y * W + x;
    const int rx = x - im.left, ry = y - im.top;
    int b, g, r;
    if (rx < 0 || rx >= im.new_w || ry < 0 || ry >= im.new_h) {
        b = g = r = pad_value;
    } else if (im.new_w == im.sw && im.new_h == im.sh) {   // LetterBox skips cv2.resize when the shape already fits
        const uint8_t* s = im.src + (long long)ry * im.pitch + 3 * rx;
        b = s[0];
        g = s[1];
        r = s[2];
    } else {
        const LbCoef cx = lb_coef(rx, im.sw, im.new_w, true);
        const LbCoef cy = lb_coef(ry, im.sh, im.new_h, false);
        const uint8_t* r0 = im.src + (long long)cy.i0 * im.pitch;
        const uint8_t* r1 = im.src + (long long)cy.i1 * im.pitch;
        int v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int h0 = (int)r0[3 * cx.i0 + c] * cx.c0 + (int)r0[3 * cx.i1 + c] * cx.c1;
            const int h1 = (int)r1[3 * cx.i0 + c] * cx.c0 + (int)r1[3 * cx.i1 + c] * cx.c1;
            v[c] = min(max((((cy.c0 * (h0 >> 4)) >> 16) + ((cy.c1 * (h1 >> 4)) >> 16) + 2) >> 2, 0), 255);
        }
        b = v[0];
        g = v[1];
        r = v[2];
    }
    out[0] = __fdiv_rn((float)r, 255.f);          // channel 0 of the network input is R ([..., ::-1])
    out[plane] = __fdiv_rn((float)g, 255.f);
    out[2 * plane] = __fdiv_rn((float)b, 255.f);
}

}  // namespace yl

extern "C" int yl_letterbox_u8(const yl_lb_image* imgs_dev, int n, float* dst_nchw, int H, int W, int pad_value,
                               void* stream) {
    YL_CHECK(imgs_dev && dst_nchw, YL_ERR_ARG, "null pointer");
    YL_CHECK(n >= 1 && n <= 65535 && H >= 1 && H <= 65535 && W >= 1, YL_ERR_ARG, "bad letterbox batch dims");
    YL_CHECK(pad_value >= 0 && pad_value <= 255, YL_ERR_ARG, "pad value must be a byte");
    const dim3 grid((unsigned)yl::ceil_div(W, 256), (unsigned)H, (unsigned)n), block(256);
    yl::letterbox_u8_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(imgs_dev, dst_nchw, H, W, pad_value);
    YL_LAUNCH_OK("letterbox_u8_kernel");
    return YL_OK;
}
