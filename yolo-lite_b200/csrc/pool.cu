// SPPF's chained MaxPool2d(k, stride 1, pad k/2) (block.py:165-184): y1 = m(x), y2 = m(y1), y3 = m(y2).
// One CTA owns one image x 8-channel group; the whole H x W plane of that group (16 B per pixel) lives in
// shared memory and each pool is a separable row-max / column-max pass, so x is read from HBM once and the
// three outputs are written once, straight into their channel slices of the concat buffer.  Out-of-image
// taps are skipped, which is MaxPool2d's -inf padding.
#include "common.cuh"

namespace yl {

struct PoolParams {
    const __nv_bfloat16* x;
    long long x_cstride;
    int x_coff;
    __nv_bfloat16* y[3];
    long long y_cstride[3];
    int y_coff[3];
    int H, W, C, r;  // r = k/2
};

__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
    __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a);
    __nv_bfloat162 y = *reinterpret_cast<__nv_bfloat162*>(&b);
    __nv_bfloat162 m = __hmax2(x, y);
    return *reinterpret_cast<uint32_t*>(&m);
}
__device__ __forceinline__ uint4 max8(uint4 a, uint4 b) {
    return make_uint4(max_bf16x2(a.x, b.x), max_bf16x2(a.y, b.y), max_bf16x2(a.z, b.z), max_bf16x2(a.w, b.w));
}

__global__ void __launch_bounds__(256) sppf_pool_kernel(const PoolParams p) {
    extern __shared__ __align__(16) uint4 plane[];  // [2][H*W]
    griddep_launch_dependents();
    griddep_wait();
    const int HW = p.H * p.W;
    uint4* cur = plane;
    uint4* tmp = plane + HW;
    const int n = blockIdx.y;
    const int c = blockIdx.x * 8;
    const long long pix0 = (long long)n * HW;
    for (int i = threadIdx.x; i < HW; i += blockDim.x)
        cur[i] = __ldg(reinterpret_cast<const uint4*>(p.x + (pix0 + i) * p.x_cstride + p.x_coff + c));
    __syncthreads();
    for (int stage = 0; stage < 3; ++stage) {
        for (int i = threadIdx.x; i < HW; i += blockDim.x) {  // row pass
            const int w = i % p.W, row = i - w;
            const int lo = max(w - p.r, 0), hi = min(w + p.r, p.W - 1);
            uint4 m = cur[row + lo];
            for (int j = lo + 1; j <= hi; ++j) m = max8(m, cur[row + j]);
            tmp[i] = m;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < HW; i += blockDim.x) {  // column pass
            const int w = i % p.W, h = i / p.W;
            const int lo = max(h - p.r, 0), hi = min(h + p.r, p.H - 1);
            uint4 m = tmp[lo * p.W + w];
            for (int j = lo + 1; j <= hi; ++j) m = max8(m, tmp[j * p.W + w]);
            cur[i] = m;
            *reinterpret_cast<uint4*>(p.y[stage] + (pix0 + i) * p.y_cstride[stage] + p.y_coff[stage] + c) = m;
        }
        __syncthreads();
    }
}

static int g_pool_max_smem = 0;
int init_pool() {
    int dev = 0;
    YL_CUDA(cudaGetDevice(&dev));
    YL_CUDA(cudaDeviceGetAttribute(&g_pool_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    YL_CUDA(cudaFuncSetAttribute(sppf_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_pool_max_smem));
    return YL_OK;
}

}  // namespace yl

extern "C" int yl_sppf_pool(const yl_tensor* x, const yl_tensor* y1, const yl_tensor* y2, const yl_tensor* y3, int k,
                            void* stream) {
    YL_CHECK(x && y1 && y2 && y3 && x->data && y1->data && y2->data && y3->data, YL_ERR_ARG, "null pointer");
    YL_CHECK(k >= 1 && (k & 1), YL_ERR_ARG, "pool size must be odd");
    const yl_tensor* ys[3] = {y1, y2, y3};
    yl::PoolParams p;
    YL_CHECK(x->dtype == YL_BF16 && x->c % 8 == 0 && x->coff % 8 == 0 && x->cstride % 8 == 0, YL_ERR_ARG,
             "pool input must be bf16 with 8-channel alignment");
    for (int i = 0; i < 3; ++i) {
        const yl_tensor* y = ys[i];
        YL_CHECK(y->dtype == YL_BF16 && y->n == x->n && y->h == x->h && y->w == x->w && y->c == x->c &&
                     y->coff % 8 == 0 && y->cstride % 8 == 0,
                 YL_ERR_ARG, "pool output %d mismatch", i);
        p.y[i] = reinterpret_cast<__nv_bfloat16*>(y->data);
        p.y_cstride[i] = y->cstride;
        p.y_coff[i] = y->coff;
    }
    p.x = reinterpret_cast<const __nv_bfloat16*>(x->data);
    p.x_cstride = x->cstride;
    p.x_coff = x->coff;
    p.H = x->h;
    p.W = x->w;
    p.C = x->c;
    p.r = k / 2;
    const size_t smem = (size_t)2 * x->h * x->w * 16;
    YL_CHECK((int)smem <= yl::g_pool_max_smem, YL_ERR_UNSUPPORTED,
             "SPPF plane %dx%d needs %zu B shared memory (max %d; yl_init called?)", x->h, x->w, smem,
             yl::g_pool_max_smem);
    dim3 grid((unsigned)(x->c / 8), (unsigned)x->n, 1);
    YL_CUDA(yl::launch_kernel(yl::sppf_pool_kernel, grid, dim3(256), smem, (cudaStream_t)stream, p));
    YL_LAUNCH_OK("sppf_pool_kernel");
    return YL_OK;
}
