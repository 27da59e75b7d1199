// CUDA-core kernels of the conv family: the direct (non tensor-core) dense convolution used for the 3-channel
// stem and for shapes the tcgen05 path rejects, the depthwise 3x3, and the one-time BN fold + weight pack.
// All are HBM/L2-bound: threads map to (pixel, 8-channel group) with channels innermost so every global
// access is a contiguous 16-byte vector inside a pixel's channel run.
#include "common.cuh"

namespace yl {

// ------------------------------------------------------------------------------------------------ direct conv
// One thread = one output pixel x 8 consecutive output channels.  Consecutive threads take consecutive
// pixels (same channel group), so weight reads are warp-uniform broadcasts and output stores are per-pixel
// 16-B vectors.
struct ConvDirectParams {
    const __nv_bfloat16* x;
    long long x_cstride;
    int x_coff, Ci, H, W;
    const __nv_bfloat16* w;  // [co_pad][k*k][ci_pad]
    int ci_pad, co_pad;
    const float* bias;
    void* y;
    long long y_cstride;
    int y_coff, y_c, y_f32;
    const __nv_bfloat16* res;
    long long res_cstride;
    int res_coff;
    int N, Ho, Wo, k, stride, pad, act, upsample;
    void* y_up;  // optional second destination (n, 2Ho, 2Wo): result replicated 2x2
    long long yu_cstride;
    int yu_coff;
};

__global__ void __launch_bounds__(256) conv_direct_kernel(const ConvDirectParams p) {
    const long long total = (long long)p.N * p.Ho * p.Wo;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= total) return;
    const int cg = blockIdx.y * 8;  // first output channel of this thread
    const int wo = (int)(pix % p.Wo);
    const int ho = (int)((pix / p.Wo) % p.Ho);
    const int n = (int)(pix / ((long long)p.Wo * p.Ho));

    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;

    const int K = p.k * p.k * p.ci_pad;
    for (int r = 0; r < p.k; ++r) {
        const int hi = ho * p.stride + r - p.pad;
        if (hi < 0 || hi >= p.H) continue;
        for (int s = 0; s < p.k; ++s) {
            const int wi = wo * p.stride + s - p.pad;
            if (wi < 0 || wi >= p.W) continue;
            const __nv_bfloat16* xp = p.x + (((long long)n * p.H + hi) * p.W + wi) * p.x_cstride + p.x_coff;
            const __nv_bfloat16* wp = p.w + (long long)cg * K + (r * p.k + s) * p.ci_pad;
            for (int c = 0; c < p.Ci; ++c) {
                const float xv = __bfloat162float(xp[c]);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(xv, __bfloat162float(wp[(long long)i * K + c]), acc[i]);
            }
        }
    }

    const int reps = p.upsample ? 4 : 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = cg + i;
        if (c >= p.y_c) break;
        float v = acc[i] + p.bias[c];
        if (p.act) v = silu_f(v);
        if (p.res) v += __bfloat162float(p.res[pix * p.res_cstride + p.res_coff + c]);
        for (int rep = 0; rep < reps; ++rep) {
            long long opix = pix;
            if (p.upsample)
                opix = ((long long)n * (2 * p.Ho) + (2 * ho + (rep >> 1))) * (2 * p.Wo) + (2 * wo + (rep & 1));
            if (p.y_f32)
                reinterpret_cast<float*>(p.y)[opix * p.y_cstride + p.y_coff + c] = v;
            else
                reinterpret_cast<__nv_bfloat16*>(p.y)[opix * p.y_cstride + p.y_coff + c] = __float2bfloat16_rn(v);
        }
        if (p.y_up) {
            for (int rep = 0; rep < 4; ++rep) {
                const long long opix =
                    ((long long)n * (2 * p.Ho) + (2 * ho + (rep >> 1))) * (2 * p.Wo) + (2 * wo + (rep & 1));
                if (p.y_f32)
                    reinterpret_cast<float*>(p.y_up)[opix * p.yu_cstride + p.yu_coff + c] = v;
                else
                    reinterpret_cast<__nv_bfloat16*>(p.y_up)[opix * p.yu_cstride + p.yu_coff + c] =
                        __float2bfloat16_rn(v);
            }
        }
    }
}

int launch_conv_direct(const yl_conv_args* a, cudaStream_t stream) {
    const yl_tensor& x = a->x;
    const yl_tensor& y = a->y;
    YL_CHECK(x.dtype == YL_BF16, YL_ERR_ARG, "conv input must be bf16");
    YL_CHECK(a->k >= 1 && a->k <= 7 && (a->k & 1), YL_ERR_ARG, "unsupported kernel size %d", a->k);
    YL_CHECK(a->stride >= 1, YL_ERR_ARG, "bad stride");
    YL_CHECK(a->ci_pad >= x.c && a->co_pad >= y.c, YL_ERR_ARG, "packed weight dims smaller than tensors");
    const int pad = a->k / 2;
    const int Ho = (x.h + 2 * pad - a->k) / a->stride + 1;
    const int Wo = (x.w + 2 * pad - a->k) / a->stride + 1;
    const int up = a->upsample2x ? 2 : 1;
    YL_CHECK(y.n == x.n && y.h == Ho * up && y.w == Wo * up, YL_ERR_ARG,
             "conv output dims mismatch: got (%d,%d,%d) expected (%d,%d,%d)", y.n, y.h, y.w, x.n, Ho * up, Wo * up);
    if (a->res.data) YL_CHECK(a->res.dtype == YL_BF16, YL_ERR_ARG, "residual must be bf16");
    // the 8-wide channel group may read weight rows up to co_pad: require the pack to cover it
    YL_CHECK(ceil_div(y.c, 8) * 8 <= a->co_pad, YL_ERR_ARG, "co_pad must cover y.c rounded up to 8");

    ConvDirectParams p;
    p.x = reinterpret_cast<const __nv_bfloat16*>(x.data);
    p.x_cstride = x.cstride;
    p.x_coff = x.coff;
    p.Ci = x.c;
    p.H = x.h;
    p.W = x.w;
    p.w = reinterpret_cast<const __nv_bfloat16*>(a->w);
    p.ci_pad = a->ci_pad;
    p.co_pad = a->co_pad;
    p.bias = a->bias;
    p.y = y.data;
    p.y_cstride = y.cstride;
    p.y_coff = y.coff;
    p.y_c = y.c;
    p.y_f32 = (y.dtype == YL_F32);
    p.res = reinterpret_cast<const __nv_bfloat16*>(a->res.data);
    p.res_cstride = a->res.cstride;
    p.res_coff = a->res.coff;
    p.N = x.n;
    p.Ho = Ho;
    p.Wo = Wo;
    p.k = a->k;
    p.stride = a->stride;
    p.pad = pad;
    p.act = a->act;
    p.upsample = a->upsample2x ? 1 : 0;
    p.y_up = a->y_up.data;
    p.yu_cstride = a->y_up.cstride;
    p.yu_coff = a->y_up.coff;
    if (a->y_up.data)
        YL_CHECK(!a->upsample2x && a->y_up.n == x.n && a->y_up.h == 2 * Ho && a->y_up.w == 2 * Wo &&
                     a->y_up.c == y.c && a->y_up.dtype == y.dtype,
                 YL_ERR_ARG, "y_up must be the (n, 2h, 2w, c) twin of y");
    const long long total = (long long)x.n * Ho * Wo;
    dim3 grid((unsigned)ceil_div64(total, 256), (unsigned)ceil_div(y.c, 8), 1);
    conv_direct_kernel<<<grid, 256, 0, stream>>>(p);
    YL_LAUNCH_OK("conv_direct_kernel");
    return YL_OK;
}

// ------------------------------------------------------------------------------------------------ depthwise 3x3
struct DwParams {
    const __nv_bfloat16* x;
    long long x_cstride;
    int x_coff;
    const __nv_bfloat16* w;  // [9][C]
    const float* bias;
    __nv_bfloat16* y;
    long long y_cstride;
    int y_coff;
    const __nv_bfloat16* add;
    long long add_cstride;
    int add_coff;
    int N, H, W, C, act;
};

__global__ void __launch_bounds__(256) dwconv3x3_kernel(const DwParams p) {
    const int groups = p.C >> 3;
    const long long total = (long long)p.N * p.H * p.W * groups;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    const long long pix = idx / groups;
    const int w = (int)(pix % p.W);
    const int h = (int)((pix / p.W) % p.H);
    const int n = (int)(pix / ((long long)p.W * p.H));
    const int c = g * 8;

    float acc[8];
    {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c + 4));
        acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
        acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int hi = h + r - 1;
        if (hi < 0 || hi >= p.H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int wi = w + s - 1;
            if (wi < 0 || wi >= p.W) continue;
            const uint4 xv = __ldg(reinterpret_cast<const uint4*>(
                p.x + (((long long)n * p.H + hi) * p.W + wi) * p.x_cstride + p.x_coff + c));
            const uint4 wv = __ldg(reinterpret_cast<const uint4*>(p.w + (r * 3 + s) * p.C + c));
            acc[0] = fmaf(bf16lo_f(xv.x), bf16lo_f(wv.x), acc[0]);
            acc[1] = fmaf(bf16hi_f(xv.x), bf16hi_f(wv.x), acc[1]);
            acc[2] = fmaf(bf16lo_f(xv.y), bf16lo_f(wv.y), acc[2]);
            acc[3] = fmaf(bf16hi_f(xv.y), bf16hi_f(wv.y), acc[3]);
            acc[4] = fmaf(bf16lo_f(xv.z), bf16lo_f(wv.z), acc[4]);
            acc[5] = fmaf(bf16hi_f(xv.z), bf16hi_f(wv.z), acc[5]);
            acc[6] = fmaf(bf16lo_f(xv.w), bf16lo_f(wv.w), acc[6]);
            acc[7] = fmaf(bf16hi_f(xv.w), bf16hi_f(wv.w), acc[7]);
        }
    }
    if (p.act) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = silu_f(acc[i]);
    }
    if (p.add) {
        const uint4 av = __ldg(reinterpret_cast<const uint4*>(p.add + pix * p.add_cstride + p.add_coff + c));
        acc[0] += bf16lo_f(av.x); acc[1] += bf16hi_f(av.x);
        acc[2] += bf16lo_f(av.y); acc[3] += bf16hi_f(av.y);
        acc[4] += bf16lo_f(av.z); acc[5] += bf16hi_f(av.z);
        acc[6] += bf16lo_f(av.w); acc[7] += bf16hi_f(av.w);
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]);
    o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]);
    o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(p.y + pix * p.y_cstride + p.y_coff + c) = o;
}

// ------------------------------------------------------------------------------------------------ BN fold + pack
__global__ void fold_pack_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, const float* __restrict__ mean,
                                 const float* __restrict__ var, const float* __restrict__ cbias, float eps, int co,
                                 int ci, int k, int co_pad, int ci_pad, int depthwise,
                                 __nv_bfloat16* __restrict__ wp, float* __restrict__ bias_out) {
    const int kk = k * k;
    const long long nW = depthwise ? (long long)kk * co_pad : (long long)co_pad * kk * ci_pad;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < nW) {
        int o, c, tap;
        if (depthwise) {
            o = (int)(idx % co_pad);
            tap = (int)(idx / co_pad);
            c = 0;
        } else {
            c = (int)(idx % ci_pad);
            tap = (int)((idx / ci_pad) % kk);
            o = (int)(idx / ((long long)ci_pad * kk));
        }
        float v = 0.f;
        if (o < co && c < ci) {
            const float scale = gamma ? gamma[o] / sqrtf(var[o] + eps) : 1.f;
            v = w[((long long)o * ci + c) * kk + tap] * scale;
        }
        wp[idx] = __float2bfloat16_rn(v);
    }
    if (idx < co_pad) {
        const int o = (int)idx;
        float b = 0.f;
        if (o < co) {
            const float cb = cbias ? cbias[o] : 0.f;
            if (gamma) {
                const float scale = gamma[o] / sqrtf(var[o] + eps);
                b = beta[o] + (cb - mean[o]) * scale;
            } else {
                b = cb;
            }
        }
        bias_out[o] = b;
    }
}

}  // namespace yl

extern "C" {

int yl_fold_bn_pack(const float* w_oihw, const float* gamma, const float* beta, const float* mean, const float* var,
                    const float* conv_bias, float eps, int co, int ci, int k, int co_pad, int ci_pad, int depthwise,
                    void* w_packed, float* bias_out, void* stream) {
    YL_CHECK(w_oihw && w_packed && bias_out, YL_ERR_ARG, "null pointer");
    YL_CHECK(co > 0 && ci > 0 && k > 0 && co_pad >= co && ci_pad >= ci, YL_ERR_ARG, "bad dims");
    YL_CHECK((gamma && beta && mean && var) || (!gamma && !beta && !mean && !var), YL_ERR_ARG,
             "BN statistics must be all present or all NULL");
    YL_CHECK(!depthwise || ci == 1, YL_ERR_ARG, "depthwise pack expects one input channel per group");
    const long long nW = depthwise ? (long long)k * k * co_pad : (long long)co_pad * k * k * ci_pad;
    const long long total = nW > co_pad ? nW : co_pad;
    yl::fold_pack_kernel<<<(unsigned)yl::ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
        w_oihw, gamma, beta, mean, var, conv_bias, eps, co, ci, k, co_pad, ci_pad, depthwise,
        reinterpret_cast<__nv_bfloat16*>(w_packed), bias_out);
    YL_LAUNCH_OK("fold_pack_kernel");
    return YL_OK;
}

int yl_dwconv3x3(const yl_tensor* x, const yl_tensor* y, const void* w, const float* bias, int act,
                 const yl_tensor* add, void* stream) {
    YL_CHECK(x && y && w && bias, YL_ERR_ARG, "null pointer");
    YL_CHECK(x->dtype == YL_BF16 && y->dtype == YL_BF16, YL_ERR_ARG, "dwconv tensors must be bf16");
    YL_CHECK(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c, YL_ERR_ARG, "dwconv shape mismatch");
    YL_CHECK(x->c % 8 == 0 && x->coff % 8 == 0 && x->cstride % 8 == 0 && y->coff % 8 == 0 && y->cstride % 8 == 0,
             YL_ERR_ARG, "dwconv needs 8-channel alignment");
    yl::DwParams p;
    p.x = reinterpret_cast<const __nv_bfloat16*>(x->data);
    p.x_cstride = x->cstride;
    p.x_coff = x->coff;
    p.w = reinterpret_cast<const __nv_bfloat16*>(w);
    p.bias = bias;
    p.y = reinterpret_cast<__nv_bfloat16*>(y->data);
    p.y_cstride = y->cstride;
    p.y_coff = y->coff;
    p.add = nullptr;
    p.add_cstride = 0;
    p.add_coff = 0;
    if (add && add->data) {
        YL_CHECK(add->dtype == YL_BF16 && add->c == x->c && add->n == x->n && add->h == x->h && add->w == x->w &&
                     add->coff % 8 == 0 && add->cstride % 8 == 0,
                 YL_ERR_ARG, "dwconv add tensor mismatch");
        p.add = reinterpret_cast<const __nv_bfloat16*>(add->data);
        p.add_cstride = add->cstride;
        p.add_coff = add->coff;
    }
    p.N = x->n;
    p.H = x->h;
    p.W = x->w;
    p.C = x->c;
    p.act = act;
    const long long total = (long long)x->n * x->h * x->w * (x->c / 8);
    yl::dwconv3x3_kernel<<<(unsigned)yl::ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(p);
    YL_LAUNCH_OK("dwconv3x3_kernel");
    return YL_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ fused stem
// First layer of the network fused with the image ingest: reads the caller's NCHW fp32 batch directly
// (predictor.py:81-84 `.to(device).float()`), 3x3 stride-2 conv with <= 4 input channels, folded BN, SiLU,
// writes NHWC bf16.  Replaces a layout pass (read 4.9 MB + write 2.5 MB per image) plus a conv pass with one
// kernel whose traffic is the algorithmic minimum (read image once, write activations once).
// One thread = one output pixel x CO channels; weights are staged in shared memory as fp32 [tap][ci][CO] and
// read as warp-uniform 16-byte broadcasts.
namespace yl {

template <int CO>
__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ x, int N, int Ci, int H, int W,
                                                        const __nv_bfloat16* __restrict__ wp, int ci_pad,
                                                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                                                        long long y_cstride, int y_coff, int Ho, int Wo, int act) {
    __shared__ __align__(16) float sw[9 * 4 * CO];
    __shared__ float sb[CO];
    for (int i = threadIdx.x + threadIdx.y * 32; i < 9 * 4 * CO; i += 256) {
        const int co = i % CO, ci = (i / CO) % 4, tap = i / (4 * CO);
        sw[i] = ci < Ci ? __bfloat162float(wp[(long long)co * 9 * ci_pad + tap * ci_pad + ci]) : 0.f;
    }
    for (int i = threadIdx.x + threadIdx.y * 32; i < CO; i += 256) sb[i] = bias[i];
    __syncthreads();
    const int wo = blockIdx.x * 32 + threadIdx.x;
    const int ho = blockIdx.y * 8 + threadIdx.y;
    const int n = blockIdx.z;
    if (wo >= Wo || ho >= Ho) return;

    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = sb[c];
    const long long plane = (long long)H * W;
    const float* xn = x + (long long)n * Ci * plane;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int hi = 2 * ho + r - 1;
        if (hi < 0 || hi >= H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int wi = 2 * wo + s - 1;
            if (wi < 0 || wi >= W) continue;
            for (int ci = 0; ci < Ci; ++ci) {
                // the model consumes bf16 images: round here exactly like the NHWC bf16 ingest did
                const float xv = __bfloat162float(__float2bfloat16_rn(__ldg(xn + ci * plane + (long long)hi * W + wi)));
                const float4* wv = reinterpret_cast<const float4*>(sw + ((r * 3 + s) * 4 + ci) * CO);
#pragma unroll
                for (int c4 = 0; c4 < CO / 4; ++c4) {
                    const float4 w4 = wv[c4];
                    acc[c4 * 4 + 0] = fmaf(xv, w4.x, acc[c4 * 4 + 0]);
                    acc[c4 * 4 + 1] = fmaf(xv, w4.y, acc[c4 * 4 + 1]);
                    acc[c4 * 4 + 2] = fmaf(xv, w4.z, acc[c4 * 4 + 2]);
                    acc[c4 * 4 + 3] = fmaf(xv, w4.w, acc[c4 * 4 + 3]);
                }
            }
        }
    }
    __nv_bfloat16* dst = y + (((long long)n * Ho + ho) * Wo + wo) * y_cstride + y_coff;
#pragma unroll
    for (int c8 = 0; c8 < CO / 8; ++c8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = act ? silu_f(acc[c8 * 8 + i]) : acc[c8 * 8 + i];
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]);
        o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]);
        o.w = pack_bf16x2(v[6], v[7]);
        reinterpret_cast<uint4*>(dst)[c8] = o;
    }
}

}  // namespace yl

extern "C" int yl_stem_conv(const float* x_nchw, int n, int ci, int h, int w, const void* w_packed, int ci_pad,
                            const float* bias, const yl_tensor* y, int act, void* stream) {
    YL_CHECK(x_nchw && w_packed && bias && y && y->data, YL_ERR_ARG, "null pointer");
    YL_CHECK(ci >= 1 && ci <= 4, YL_ERR_UNSUPPORTED, "stem kernel takes 1..4 input channels");
    const int Ho = (h + 2 - 3) / 2 + 1, Wo = (w + 2 - 3) / 2 + 1;
    YL_CHECK(y->dtype == YL_BF16 && y->n == n && y->h == Ho && y->w == Wo, YL_ERR_ARG, "stem output shape mismatch");
    YL_CHECK(y->coff % 8 == 0 && y->cstride % 8 == 0, YL_ERR_ARG, "stem output needs 8-channel alignment");
    dim3 grid((unsigned)yl::ceil_div(Wo, 32), (unsigned)yl::ceil_div(Ho, 8), (unsigned)n), block(32, 8, 1);
    __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y->data);
    const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(w_packed);
    cudaStream_t s = (cudaStream_t)stream;
    switch (y->c) {
        case 16: yl::stem_conv_kernel<16><<<grid, block, 0, s>>>(x_nchw, n, ci, h, w, wp, ci_pad, bias, yp, y->cstride, y->coff, Ho, Wo, act); break;
        case 32: yl::stem_conv_kernel<32><<<grid, block, 0, s>>>(x_nchw, n, ci, h, w, wp, ci_pad, bias, yp, y->cstride, y->coff, Ho, Wo, act); break;
        case 48: yl::stem_conv_kernel<48><<<grid, block, 0, s>>>(x_nchw, n, ci, h, w, wp, ci_pad, bias, yp, y->cstride, y->coff, Ho, Wo, act); break;
        case 64: yl::stem_conv_kernel<64><<<grid, block, 0, s>>>(x_nchw, n, ci, h, w, wp, ci_pad, bias, yp, y->cstride, y->coff, Ho, Wo, act); break;
        case 96: yl::stem_conv_kernel<96><<<grid, block, 0, s>>>(x_nchw, n, ci, h, w, wp, ci_pad, bias, yp, y->cstride, y->coff, Ho, Wo, act); break;
        default:
            yl::set_error("stem kernel is built for 16/32/48/64/96 output channels (yolo11 n/s/-/m,l/x), got %d", y->c);
            return YL_ERR_UNSUPPORTED;
    }
    YL_LAUNCH_OK("stem_conv_kernel");
    return YL_OK;
}
