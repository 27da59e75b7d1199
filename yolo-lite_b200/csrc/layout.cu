// Layout kernels either side of the NHWC bf16 core: image ingest (NCHW fp32 -> NHWC bf16), reference-layout
// export (NHWC -> NCHW fp32), channel-slice copy and nearest 2x upsample.  Pure HBM-bound byte movers.
#include <cuda_fp16.h>

#include "common.cuh"

namespace yl {

// NCHW fp32 -> NHWC bf16.  Tile of 32 pixels x C channels staged through shared memory so that both the
// global read (contiguous along W) and the global write (contiguous along C) are coalesced.
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                           long long y_cstride, int y_coff, int C, long long HW) {
    __shared__ float tile[64 * 33];  // [64 channels][32 pixels + 1]
    const int n = blockIdx.y;
    const int c0 = blockIdx.z * 64;
    const int Cb = min(64, C - c0);
    const long long p0 = (long long)blockIdx.x * 32;
    for (int i = threadIdx.x; i < Cb * 32; i += blockDim.x) {
        const int c = i >> 5, px = i & 31;
        const long long p = p0 + px;
        tile[c * 33 + px] = p < HW ? x[((long long)n * C + c0 + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Cb * 32; i += blockDim.x) {
        const int px = i / Cb, c = i - px * Cb;
        const long long p = p0 + px;
        if (p < HW) y[((long long)n * HW + p) * y_cstride + y_coff + c0 + c] = __float2bfloat16_rn(tile[c * 33 + px]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ x, long long x_cstride, int x_coff,
                                                           float* __restrict__ y, int C, long long HW) {
    __shared__ float tile[32 * 65];  // [32 pixels][64 channels + 1]
    const int n = blockIdx.y;
    const int c0 = blockIdx.z * 64;
    const int Cb = min(64, C - c0);
    const long long p0 = (long long)blockIdx.x * 32;
    for (int i = threadIdx.x; i < Cb * 32; i += blockDim.x) {
        const int px = i / Cb, c = i - px * Cb;
        const long long p = p0 + px;
        float v = 0.f;
        if (p < HW) v = (float)x[((long long)n * HW + p) * x_cstride + x_coff + c0 + c];
        tile[px * 65 + c] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Cb * 32; i += blockDim.x) {
        const int c = i >> 5, px = i & 31;
        const long long p = p0 + px;
        if (p < HW) y[((long long)n * C + c0 + c) * HW + p] = tile[px * 65 + c];
    }
}

// 16-byte vectors: one thread moves 8 bf16 channels of one pixel.
__global__ void __launch_bounds__(256) copy_slice_kernel(const __nv_bfloat16* __restrict__ x, long long x_cstride,
                                                         int x_coff, __nv_bfloat16* __restrict__ y,
                                                         long long y_cstride, int y_coff, int C, long long pixels) {
    griddep_launch_dependents();
    griddep_wait();
    const int groups = C >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= pixels * groups) return;
    const int g = (int)(idx % groups);
    const long long p = idx / groups;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + p * x_cstride + x_coff + g * 8));
    *reinterpret_cast<uint4*>(y + p * y_cstride + y_coff + g * 8) = v;
}

__global__ void __launch_bounds__(256) upsample2x_kernel(const __nv_bfloat16* __restrict__ x, long long x_cstride,
                                                         int x_coff, __nv_bfloat16* __restrict__ y,
                                                         long long y_cstride, int y_coff, int C, int N, int H, int W) {
    // thread = (output pixel, 8-channel group): reads are re-used 4x through L1/L2, writes are coalesced
    griddep_launch_dependents();
    griddep_wait();
    const int groups = C >> 3;
    const int Ho = 2 * H, Wo = 2 * W;
    const long long total = (long long)N * Ho * Wo * groups;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    const long long op = idx / groups;
    const int wo = (int)(op % Wo);
    const int ho = (int)((op / Wo) % Ho);
    const int n = (int)(op / ((long long)Wo * Ho));
    const long long ip = ((long long)n * H + (ho >> 1)) * W + (wo >> 1);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ip * x_cstride + x_coff + g * 8));
    *reinterpret_cast<uint4*>(y + op * y_cstride + y_coff + g * 8) = v;
}

// fp16 -> fp32 (8 elements per thread): the ingest of half-precision host tensors (predictor.py:81-84 `.float()`)
__global__ void __launch_bounds__(256) f16_to_f32_kernel(const __half* __restrict__ x, float* __restrict__ y, long long n8,
                                                         long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n8) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
        const float2 a = __half22float2(h[0]), b = __half22float2(h[1]), c = __half22float2(h[2]), d = __half22float2(h[3]);
        float4* o = reinterpret_cast<float4*>(y) + 2 * i;
        o[0] = make_float4(a.x, a.y, b.x, b.y);
        o[1] = make_float4(c.x, c.y, d.x, d.y);
    } else if (i == n8) {
        for (long long j = n8 * 8; j < n; ++j) y[j] = __half2float(x[j]);   // tail (< 8 elements)
    }
}

// uint8 image bytes -> fp32 in [0, 1]: float(b) / 255 with a true (round-to-nearest) division, 16 bytes per thread
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ x, float* __restrict__ y, long long n16,
                                                        long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        float4* o = reinterpret_cast<float4*>(y) + 4 * i;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            o[k] = make_float4(__fdiv_rn((float)(w[k] & 0xffu), 255.f), __fdiv_rn((float)((w[k] >> 8) & 0xffu), 255.f),
                               __fdiv_rn((float)((w[k] >> 16) & 0xffu), 255.f), __fdiv_rn((float)(w[k] >> 24), 255.f));
    } else if (i == n16) {
        for (long long j = n16 * 16; j < n; ++j) y[j] = __fdiv_rn((float)x[j], 255.f);
    }
}

static bool aligned8(const yl_tensor* t) { return t->c % 8 == 0 && t->coff % 8 == 0 && t->cstride % 8 == 0; }

}  // namespace yl

extern "C" {

int yl_to_f32(const void* x, int x_dtype, float* y, long long n, void* stream) {
    YL_CHECK(x && y && n >= 0, YL_ERR_ARG, "bad to_f32 arguments");
    YL_CHECK(x_dtype == YL_F16 || x_dtype == YL_U8, YL_ERR_ARG, "to_f32 converts YL_F16 or YL_U8 (got dtype %d)", x_dtype);
    YL_CHECK(((uintptr_t)x | (uintptr_t)y) % 16 == 0, YL_ERR_ARG, "to_f32 needs 16-byte aligned pointers");
    if (n == 0) return YL_OK;
    if (x_dtype == YL_F16) {
        const long long n8 = n / 8;
        yl::f16_to_f32_kernel<<<(unsigned)yl::ceil_div64(n8 + 1, 256), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const __half*>(x), y, n8, n);
    } else {
        const long long n16 = n / 16;
        yl::u8_to_f32_kernel<<<(unsigned)yl::ceil_div64(n16 + 1, 256), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const uint8_t*>(x), y, n16, n);
    }
    YL_LAUNCH_OK("to_f32_kernel");
    return YL_OK;
}

int yl_nchw_to_nhwc(const float* x_nchw, const yl_tensor* y, void* stream) {
    YL_CHECK(x_nchw && y && y->data, YL_ERR_ARG, "null pointer");
    YL_CHECK(y->dtype == YL_BF16, YL_ERR_ARG, "nchw_to_nhwc writes bf16");
    YL_CHECK(y->c > 0 && y->coff + y->c <= y->cstride, YL_ERR_ARG, "bad channel slice");
    const long long HW = (long long)y->h * y->w;
    dim3 grid((unsigned)yl::ceil_div64(HW, 32), (unsigned)y->n, (unsigned)yl::ceil_div(y->c, 64));
    yl::nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        x_nchw, reinterpret_cast<__nv_bfloat16*>(y->data), y->cstride, y->coff, y->c, HW);
    YL_LAUNCH_OK("nchw_to_nhwc_kernel");
    return YL_OK;
}

int yl_nhwc_to_nchw(const yl_tensor* x, float* y_nchw, void* stream) {
    YL_CHECK(x && x->data && y_nchw, YL_ERR_ARG, "null pointer");
    YL_CHECK(x->c > 0 && x->coff + x->c <= x->cstride, YL_ERR_ARG, "bad channel slice");
    const long long HW = (long long)x->h * x->w;
    dim3 grid((unsigned)yl::ceil_div64(HW, 32), (unsigned)x->n, (unsigned)yl::ceil_div(x->c, 64));
    const size_t smem = 0;
    if (x->dtype == YL_BF16)
        yl::nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, 256, smem, (cudaStream_t)stream>>>(
            reinterpret_cast<const __nv_bfloat16*>(x->data), x->cstride, x->coff, y_nchw, x->c, HW);
    else
        yl::nhwc_to_nchw_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>(
            reinterpret_cast<const float*>(x->data), x->cstride, x->coff, y_nchw, x->c, HW);
    YL_LAUNCH_OK("nhwc_to_nchw_kernel");
    return YL_OK;
}

int yl_copy_slice(const yl_tensor* x, const yl_tensor* y, void* stream) {
    YL_CHECK(x && y && x->data && y->data, YL_ERR_ARG, "null pointer");
    YL_CHECK(x->dtype == YL_BF16 && y->dtype == YL_BF16, YL_ERR_ARG, "copy_slice is bf16 only");
    YL_CHECK(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c, YL_ERR_ARG, "copy_slice shape mismatch");
    YL_CHECK(yl::aligned8(x) && yl::aligned8(y), YL_ERR_ARG, "copy_slice needs 8-channel alignment");
    const long long pixels = (long long)x->n * x->h * x->w;
    const long long total = pixels * (x->c / 8);
    YL_CUDA(yl::launch_kernel(yl::copy_slice_kernel, dim3((unsigned)yl::ceil_div64(total, 256)), dim3(256), 0,
                              (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(x->data),
                              (long long)x->cstride, x->coff, reinterpret_cast<__nv_bfloat16*>(y->data),
                              (long long)y->cstride, y->coff, x->c, pixels));
    YL_LAUNCH_OK("copy_slice_kernel");
    return YL_OK;
}

int yl_upsample2x(const yl_tensor* x, const yl_tensor* y, void* stream) {
    YL_CHECK(x && y && x->data && y->data, YL_ERR_ARG, "null pointer");
    YL_CHECK(x->dtype == YL_BF16 && y->dtype == YL_BF16, YL_ERR_ARG, "upsample2x is bf16 only");
    YL_CHECK(x->n == y->n && 2 * x->h == y->h && 2 * x->w == y->w && x->c == y->c, YL_ERR_ARG,
             "upsample2x shape mismatch");
    YL_CHECK(yl::aligned8(x) && yl::aligned8(y), YL_ERR_ARG, "upsample2x needs 8-channel alignment");
    const long long total = (long long)y->n * y->h * y->w * (y->c / 8);
    YL_CUDA(yl::launch_kernel(yl::upsample2x_kernel, dim3((unsigned)yl::ceil_div64(total, 256)), dim3(256), 0,
                              (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(x->data),
                              (long long)x->cstride, x->coff, reinterpret_cast<__nv_bfloat16*>(y->data),
                              (long long)y->cstride, y->coff, x->c, x->n, x->h, x->w));
    YL_LAUNCH_OK("upsample2x_kernel");
    return YL_OK;
}

}  // extern "C"
