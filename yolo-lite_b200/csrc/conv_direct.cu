// CUDA-core kernels of the conv family: the direct (non tensor-core) dense convolution used for the 3-channel
// stem and for shapes the tcgen05 path rejects, the depthwise 3x3, and the one-time BN fold + weight pack.
// All are HBM/L2-bound: threads map to (pixel, 8-channel group) with channels innermost so every global
// access is a contiguous 16-byte vector inside a pixel's channel run.
#include <stdlib.h>

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
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

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

// One thread = a strip of P consecutive output pixels of one row x 8 channels.  The 3 x (P+2) input vectors
// are each loaded once (4.5 loads per output for P = 4 instead of 9), the 9 x 8 weights live in registers as
// fp32.  Consecutive threads take consecutive 8-channel groups, so a warp reads whole pixels (>= 128 B runs).
template <int P>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const DwParams p) {
    griddep_launch_dependents();
    // 32-bit index math only (64-bit div/mod is a ~100-instruction software routine): blockIdx.y = image,
    // the x index runs over (row, strip, channel group) of one image
    const unsigned groups = (unsigned)p.C >> 3;
    const unsigned strips = (unsigned)(p.W + P - 1) / P;
    const unsigned per_image = (unsigned)p.H * strips * groups;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_image) return;
    const int g = (int)(idx % groups);
    unsigned t = idx / groups;
    const int w0 = (int)(t % strips) * P;
    const int h = (int)(t / strips);
    const int n = blockIdx.y;
    const int c = g * 8;

    griddep_wait();  // weights / bias are constants; activations only from here on
    // phase 1: every input vector of the 3 x (P+2) window in flight at once (clamped addresses, no branches:
    // the kernel is latency-bound, so the loads must not be serialised behind bounds checks)
    const long long row0 = ((long long)n * p.H + h) * p.W;
    uint4 xv[3][P + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int hi = min(max(h + r - 1, 0), p.H - 1);
        const __nv_bfloat16* xrow = p.x + ((long long)n * p.H + hi) * p.W * p.x_cstride + p.x_coff + c;
#pragma unroll
        for (int j = 0; j < P + 2; ++j) {
            const int wi = min(max(w0 + j - 1, 0), p.W - 1);
            xv[r][j] = __ldg(reinterpret_cast<const uint4*>(xrow + (long long)wi * p.x_cstride));
        }
    }
    uint4 wv[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) wv[tap] = __ldg(reinterpret_cast<const uint4*>(p.w + tap * p.C + c));
    float acc[P][8];
    {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c + 4));
#pragma unroll
        for (int o = 0; o < P; ++o) {
            acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w;
            acc[o][4] = b1.x; acc[o][5] = b1.y; acc[o][6] = b1.z; acc[o][7] = b1.w;
        }
    }
    // phase 2: out-of-image taps contribute zero (the clamped loads fetched a valid but irrelevant pixel)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int hi = h + r - 1;
        const bool row_ok = hi >= 0 && hi < p.H;
#pragma unroll
        for (int j = 0; j < P + 2; ++j) {
            const int wi = w0 + j - 1;
            const bool ok = row_ok && wi >= 0 && wi < p.W;
            const uint4 q = xv[r][j];
            float xf[8];
            xf[0] = ok ? bf16lo_f(q.x) : 0.f; xf[1] = ok ? bf16hi_f(q.x) : 0.f;
            xf[2] = ok ? bf16lo_f(q.y) : 0.f; xf[3] = ok ? bf16hi_f(q.y) : 0.f;
            xf[4] = ok ? bf16lo_f(q.z) : 0.f; xf[5] = ok ? bf16hi_f(q.z) : 0.f;
            xf[6] = ok ? bf16lo_f(q.w) : 0.f; xf[7] = ok ? bf16hi_f(q.w) : 0.f;
            // input column j feeds output o = j - s (tap column s), 0 <= o < P
#pragma unroll
            for (int s2 = 0; s2 < 3; ++s2) {
                const int o = j - s2;
                if (o < 0 || o >= P) continue;
                const uint4 wq = wv[r * 3 + s2];
                acc[o][0] = fmaf(xf[0], bf16lo_f(wq.x), acc[o][0]);
                acc[o][1] = fmaf(xf[1], bf16hi_f(wq.x), acc[o][1]);
                acc[o][2] = fmaf(xf[2], bf16lo_f(wq.y), acc[o][2]);
                acc[o][3] = fmaf(xf[3], bf16hi_f(wq.y), acc[o][3]);
                acc[o][4] = fmaf(xf[4], bf16lo_f(wq.z), acc[o][4]);
                acc[o][5] = fmaf(xf[5], bf16hi_f(wq.z), acc[o][5]);
                acc[o][6] = fmaf(xf[6], bf16lo_f(wq.w), acc[o][6]);
                acc[o][7] = fmaf(xf[7], bf16hi_f(wq.w), acc[o][7]);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < P; ++o) {
        const int w = w0 + o;
        if (w >= p.W) break;
        const long long pix = row0 + w;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = p.act ? silu_f(acc[o][i]) : acc[o][i];
        if (p.add) {
            const uint4 av = __ldg(reinterpret_cast<const uint4*>(p.add + pix * p.add_cstride + p.add_coff + c));
            v[0] += bf16lo_f(av.x); v[1] += bf16hi_f(av.x);
            v[2] += bf16lo_f(av.y); v[3] += bf16hi_f(av.y);
            v[4] += bf16lo_f(av.z); v[5] += bf16hi_f(av.z);
            v[6] += bf16lo_f(av.w); v[7] += bf16hi_f(av.w);
        }
        uint4 ov;
        ov.x = pack_bf16x2(v[0], v[1]);
        ov.y = pack_bf16x2(v[2], v[3]);
        ov.z = pack_bf16x2(v[4], v[5]);
        ov.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(p.y + pix * p.y_cstride + p.y_coff + c) = ov;
    }
}


// Depthwise 3x3 on the (legacy) tensor-core path.  The strip kernel above is instruction-issue bound (ncu: ~50
// thread instructions per output element, mostly bf16->fp32 unpacking and FMAs, issue slots 70 % busy at 1.8
// TB/s).  Here 16 pixels x 8 channels are one m16n8k16 mma.sync step chain with K = (tap pair) x (8 channels)
// and a B operand that is diagonal in the channel: D[px, c] = sum_tap x[px + tap, c] * w[tap, c].  7/8 of the
// multiplies hit zeros, but 5 MMAs replace 144 FMAs + ~200 unpack instructions.
//   block = (8 x 16)-pixel tile of one image x up to 8 channel groups; a warp owns TWO adjacent groups;
//   smem  = the 10 x 18 halo patch, stored [group][row][col] as 16-byte pixels so that an A fragment word is
//           one conflict-free 32-bit load (lanes g = 8 consecutive pixels, t = word of the pixel);
//   a patch row feeds three output rows: the fragments live in a rolling 3-row register window, so each
//           output row costs 6 shared loads per group instead of 18 (the first version was L1-wavefront bound);
//   the results of the two groups are transposed inside each lane quad (4 shuffles) so a lane stores 16 bytes
//           and a quad writes 2 x 32 contiguous bytes: full sectors, half the store wavefronts.
template <bool ACT, bool ADD>
__global__ void __launch_bounds__(128) dwconv3x3_mma_kernel(const DwParams p, int tiles_w, int ngrp, int gstride) {
    constexpr int TH = 8, TW = 16, PH = TH + 2, PW = TW + 2;
    extern __shared__ uint4 dw_patch[];  // [ngrp][gstride] pixels of 8 channels
    griddep_launch_dependents();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int w0 = ((int)blockIdx.x % tiles_w) * TW;
    const int h0 = ((int)blockIdx.x / tiles_w) * TH;
    const int cg0 = blockIdx.y * ngrp;            // first channel group of this block
    const int n = blockIdx.z;
    const int ng = min(ngrp, (p.C >> 3) - cg0);   // groups really present (the last block may be short)
    const int myg = 2 * warp;                     // this warp's first group within the block
    const bool has_a = myg < ng, has_b = myg + 1 < ng;

    // ---- constants first (they overlap the previous kernel's tail under PDL): the diagonal B fragments
    uint32_t bq[2][5][2];
    float bia[2][2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const bool has = q ? has_b : has_a;
        const int c0 = (cg0 + myg + q) * 8;
#pragma unroll
        for (int ks = 0; ks < 5; ++ks)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int tap = 2 * ks + hh;
                uint32_t v = 0;
                if (has && tap < 9 && (g >> 1) == t) {
                    const uint32_t wbits = (uint32_t)__bfloat16_as_ushort(p.w[tap * p.C + c0 + g]);
                    v = (g & 1) ? (wbits << 16) : wbits;   // k = 2t (+1): low / high half of the register
                }
                bq[q][ks][hh] = v;
            }
        bia[q][0] = has ? __ldg(p.bias + c0 + 2 * t) : 0.f;
        bia[q][1] = has ? __ldg(p.bias + c0 + 2 * t + 1) : 0.f;
    }
    griddep_wait();

    // ---- stage the halo patch with asynchronous 16-byte copies (zero-filled where the pixel is padding):
    // consecutive threads take consecutive channel groups of one pixel, i.e. whole pixels, coalesced
    constexpr int npix = PH * PW;
    {
        const uint32_t patch_s = smem_u32(dw_patch);
        const __nv_bfloat16* src0 = p.x + (long long)n * p.H * p.W * p.x_cstride + p.x_coff + cg0 * 8;
        for (int idx = threadIdx.x; idx < npix * ng; idx += blockDim.x) {
            const int pix = idx / ng, cg = idx - pix * ng;
            const int pr = pix / PW, pc = pix - pr * PW;
            const int hi = h0 - 1 + pr, wi = w0 - 1 + pc;
            const bool ok = hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            const int off = ok ? (hi * p.W + wi) : 0;   // < 2^31 pixels per image (checked by the launcher)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(patch_s + (uint32_t)(cg * gstride + pix) * 16u),
                         "l"(src0 + (long long)off * p.x_cstride + cg * 8), "r"(ok ? 16 : 0)
                         : "memory");
        }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (!has_a) return;

    // every smem offset below is a compile-time constant relative to rp0 / rp1
    const uint32_t* rp0 = reinterpret_cast<const uint32_t*>(dw_patch + myg * gstride) + g * 4 + t;
    const uint32_t* rp1 = rp0 + (has_b ? gstride * 4 : 0);
    // after the quad transpose lane t stores: pixel g + 8 * (t >> 1), group myg + (t & 1)
    const int spx = w0 + g + 8 * (t >> 1);
    const bool st_ok = spx < p.W && ((t & 1) == 0 || has_b);
    const long long pixs = ((long long)n * p.H + h0) * p.W + spx;
    char* yp = reinterpret_cast<char*>(p.y + pixs * p.y_cstride + p.y_coff + (cg0 + myg + (t & 1)) * 8);
    const long long y_row = (long long)p.W * p.y_cstride * 2;
    // residual (ADD) is read in the accumulator layout: pixel g (+8), channels 2t, 2t+1 of each group
    const long long pixa = ((long long)n * p.H + h0) * p.W + w0 + g;
    const char* ap = ADD ? reinterpret_cast<const char*>(p.add + pixa * p.add_cstride + p.add_coff + (cg0 + myg) * 8 + 2 * t)
                         : nullptr;
    const long long a_row = (long long)p.W * p.add_cstride * 2, a_half = 16 * p.add_cstride;
    const bool ok0 = w0 + g < p.W, ok1 = w0 + g + 8 < p.W;
    const int rows = min(TH, p.H - h0);

    uint32_t win[2][3][3][2];   // [group][patch row % 3][dc][pixel half]
#define YL_DW_LOADROW(PR_)                                                   \
    _Pragma("unroll") for (int dc = 0; dc < 3; ++dc) {                        \
        win[0][(PR_) % 3][dc][0] = rp0[((PR_) * PW + dc) * 4];                \
        win[0][(PR_) % 3][dc][1] = rp0[((PR_) * PW + dc + 8) * 4];            \
        win[1][(PR_) % 3][dc][0] = rp1[((PR_) * PW + dc) * 4];                \
        win[1][(PR_) % 3][dc][1] = rp1[((PR_) * PW + dc + 8) * 4];            \
    }
    YL_DW_LOADROW(0)
    YL_DW_LOADROW(1)
#pragma unroll
    for (int r = 0; r < TH; ++r) {
        if (r < rows) {
            YL_DW_LOADROW(r + 2)
            float acc[2][4];
            uint32_t av[2][2] = {{0u, 0u}, {0u, 0u}};
            if (ADD) {
                if (ok0) av[0][0] = __ldg(reinterpret_cast<const uint32_t*>(ap));
                if (ok1) av[0][1] = __ldg(reinterpret_cast<const uint32_t*>(ap + a_half));
                if (has_b) {
                    if (ok0) av[1][0] = __ldg(reinterpret_cast<const uint32_t*>(ap + 16));
                    if (ok1) av[1][1] = __ldg(reinterpret_cast<const uint32_t*>(ap + a_half + 16));
                }
                ap += a_row;
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                acc[q][0] = bia[q][0]; acc[q][1] = bia[q][1]; acc[q][2] = bia[q][0]; acc[q][3] = bia[q][1];
#pragma unroll
                for (int ks = 0; ks < 5; ++ks) {
                    uint32_t a[4];
                    {
                        const int tap = 2 * ks, dr = tap / 3, dc = tap - 3 * dr;
                        a[0] = win[q][(r + dr) % 3][dc][0];
                        a[1] = win[q][(r + dr) % 3][dc][1];
                    }
                    if (ks < 4) {
                        const int tap = 2 * ks + 1, dr = tap / 3, dc = tap - 3 * dr;
                        a[2] = win[q][(r + dr) % 3][dc][0];
                        a[3] = win[q][(r + dr) % 3][dc][1];
                    } else {
                        a[2] = 0u;
                        a[3] = 0u;
                    }
                    mma_bf16_16816(acc[q], a, bq[q][ks][0], bq[q][ks][1]);
                }
                if (ACT) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[q][i] = silu_fast(acc[q][i]);
                }
                if (ADD) {
                    acc[q][0] += bf16lo_f(av[q][0]); acc[q][1] += bf16hi_f(av[q][0]);
                    acc[q][2] += bf16lo_f(av[q][1]); acc[q][3] += bf16hi_f(av[q][1]);
                }
            }
            // items: 0 = (pixel g, group A), 1 = (pixel g, group B), 2 = (pixel g+8, A), 3 = (pixel g+8, B); lane t
            // holds word t of every item -> 4x4 transpose inside the quad -> lane t holds the 16 bytes of item t
            uint32_t v0 = pack_bf16x2(acc[0][0], acc[0][1]), v1 = pack_bf16x2(acc[1][0], acc[1][1]);
            uint32_t v2 = pack_bf16x2(acc[0][2], acc[0][3]), v3 = pack_bf16x2(acc[1][2], acc[1][3]);
            {
                const uint32_t s0 = (t & 1) ? v0 : v1, s1 = (t & 1) ? v2 : v3;
                const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
                if (t & 1) { v0 = r0; v2 = r1; } else { v1 = r0; v3 = r1; }
            }
            {
                const uint32_t s0 = (t & 2) ? v0 : v2, s1 = (t & 2) ? v1 : v3;
                const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
                if (t & 2) { v0 = r0; v1 = r1; } else { v2 = r0; v3 = r1; }
            }
            if (st_ok) *reinterpret_cast<uint4*>(yp) = make_uint4(v0, v1, v2, v3);
            yp += y_row;
        }
    }
#undef YL_DW_LOADROW
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
    YL_CHECK(x->n <= 65535, YL_ERR_ARG, "batch too large for one launch");
    cudaStream_t s = (cudaStream_t)stream;
    {
        static const int use_mma = [] {            // A/B switch, read once per process (never on a later launch)
            const char* em = getenv("YL_DW_MMA");
            return !(em && *em && atoi(em) == 0);
        }();
        if (use_mma) {
            // tensor-core path: (8 x 16)-pixel tiles, up to 8 channel groups per block, two groups per warp
            const int groups = x->c / 8;
            const int cblocks = yl::ceil_div(groups, 8);
            int ngrp = yl::ceil_div(groups, cblocks);
            ngrp += ngrp & 1;                                      // even: warps own whole pairs
            const int npix = 10 * 18;
            const int gstride = npix + ((1 - npix) % 8 + 8) % 8;  // = 1 (mod 8): groups land 4 banks apart
            const int tiles_w = yl::ceil_div(x->w, 16), tiles_h = yl::ceil_div(x->h, 8);
            const dim3 grid((unsigned)(tiles_w * tiles_h), (unsigned)yl::ceil_div(groups, ngrp), (unsigned)x->n);
            const dim3 block(32 * (ngrp / 2));
            const size_t smem = (size_t)ngrp * gstride * 16;
#define YL_DWM(ACT_, ADD_) \
    YL_CUDA(yl::launch_kernel(yl::dwconv3x3_mma_kernel<ACT_, ADD_>, grid, block, smem, s, p, tiles_w, ngrp, gstride))
            if (act) {
                if (p.add) YL_DWM(true, true); else YL_DWM(true, false);
            } else {
                if (p.add) YL_DWM(false, true); else YL_DWM(false, false);
            }
#undef YL_DWM
            YL_LAUNCH_OK("dwconv3x3_mma_kernel");
            return YL_OK;
        }
    }
    static const int strip = [] {                  // A/B switch, read once per process
        const char* e = getenv("YL_DW_STRIP");
        return (e && *e) ? atoi(e) : 2;            // measured: 2-pixel strips are fastest (tools/bench_kernels.py)
    }();
    const int P = strip >= 4 ? 4 : (strip >= 2 ? 2 : 1);
    const long long per_image = (long long)x->h * yl::ceil_div(x->w, P) * (x->c / 8);
    YL_CHECK(per_image < (1ll << 31), YL_ERR_ARG, "image too large");
    const dim3 grid((unsigned)yl::ceil_div64(per_image, 256), (unsigned)x->n, 1), block(256);
    if (P == 4) YL_CUDA(yl::launch_kernel(yl::dwconv3x3_kernel<4>, grid, block, 0, s, p));
    else if (P == 2) YL_CUDA(yl::launch_kernel(yl::dwconv3x3_kernel<2>, grid, block, 0, s, p));
    else YL_CUDA(yl::launch_kernel(yl::dwconv3x3_kernel<1>, grid, block, 0, s, p));
    YL_LAUNCH_OK("dwconv3x3_kernel");
    return YL_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ fused stem
// First layer of the network fused with the image ingest: reads the caller's NCHW fp32 batch directly
// (predictor.py:81-84 `.to(device).float()`), 3x3 stride-2 conv with <= 4 input channels, folded BN, SiLU,
// writes NHWC bf16.  Replaces a layout pass (read 4.9 MB + write 2.5 MB per image) plus a conv pass with one
// kernel whose traffic is the algorithmic minimum (read image once, write activations once).
//
// The GEMM is M = pixels, N = CO, K = 9 taps x 4 (channel-padded) = 36: too thin for tcgen05 (no TMA im2col of
// an fp32 NCHW image), but a CUDA-core kernel is FMA/LDS-issue bound at ~4x the HBM time.  A block stages the
// 17 x 132-pixel input patch of its 8 x 64 output tile in shared memory with coalesced 16-byte loads, rounded
// to bf16 exactly like the NHWC ingest would and interleaved [row][col][4 ch], so every k-pair of an im2col
// row is one conflict-free 32-bit shared load straight into an mma.sync.m16n8k16 A fragment (bf16 x bf16 ->
// fp32).  A warp walks the 64 output pixels of one row in four 16-pixel steps; the weight (B) fragments and
// biases stay in registers for the whole walk.
namespace yl {

constexpr int kStemRowPixels = 64;  // output pixels of one row per warp
constexpr int kStemWarps = 8;       // output rows per block
// input patch of a block: rows 2*h0-1 .. 2*h0+15, columns 2*w0-4 .. 2*w0+127 (16-byte aligned start), staged in
// shared memory as bf16 pixel-interleaved [row][col][4 channels] so a k-pair of the im2col row is one 32-bit word
constexpr int kStemPatchRows = 2 * kStemWarps + 1;
constexpr int kStemPatchCols = 2 * kStemRowPixels + 4;

// K order: k = tap * 4 + ci (ci padded to 4), K = 36 -> 3 k-steps of 16.
template <int CO, int CI>
__global__ void __launch_bounds__(32 * kStemWarps, 3) stem_conv_kernel(
    const float* __restrict__ x, int N, int H, int W, const __nv_bfloat16* __restrict__ wp, int ci_pad,
    const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, long long y_cstride, int y_coff, int Ho, int Wo,
    int act) {
    // Output channels are processed in passes of NTP n-tiles (<= 32 channels): the stationary weight fragments, biases
    // and accumulators of ALL 64 / 96 channels cost 144 registers, i.e. one block of 8 warps per SM (ncu: 12 % of the
    // warp slots busy, the staging barrier fully exposed).  A pass re-reads its A fragments from the staged patch
    // (shared-memory bandwidth is plentiful here) and keeps the kernel at <= 80 registers = 3 blocks per SM.
    constexpr int NT_ALL = CO / 8;
    constexpr int NT = NT_ALL <= 4 ? NT_ALL : ((NT_ALL % 4 == 0) ? 4 : 2);
    constexpr int PASSES = NT_ALL / NT;
    static_assert(NT_ALL % NT == 0, "pass width must divide the channel count");
    constexpr int KSTEPS = 3;
    constexpr int Ci = CI;
    __shared__ __align__(16) uint2 patch[kStemPatchRows * kStemPatchCols];  // one pixel = 4 x bf16 = 8 bytes
    griddep_launch_dependents();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int h0 = blockIdx.y * kStemWarps;
    const int w0 = blockIdx.x * kStemRowPixels;
    const int n = blockIdx.z;
    const long long plane = (long long)H * W;

    // B fragments: b0 = {W[k0][n], W[k0+1][n]}, b1 = {W[k0+8][n], W[k0+9][n]}, k0 = 16*ks + 2t, n = 8*nt + g
    uint32_t bfrag[KSTEPS][NT][2];
    float bia[NT][2];
    auto load_weights = [&](int pass) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int co = (pass * NT + nt) * 8 + g;
        const __nv_bfloat16* wrow = wp + (long long)co * 9 * ci_pad;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t v = 0;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 16 * ks + 2 * t + e + hh * 8;
                    const int tap = k >> 2, ci = k & 3;
                    uint32_t bits = 0;
                    if (tap < 9 && ci < Ci) bits = (uint32_t)__bfloat16_as_ushort(wrow[tap * ci_pad + ci]);
                    v |= bits << (16 * e);
                }
                bfrag[ks][nt][hh] = v;
            }
        bia[nt][0] = __ldg(bias + (pass * NT + nt) * 8 + 2 * t);
        bia[nt][1] = __ldg(bias + (pass * NT + nt) * 8 + 2 * t + 1);
    }
    };
    load_weights(0);
    griddep_wait();

    // ---- stage the patch: a work item = 4 consecutive columns of one patch row, all channels (coalesced 16-B
    // loads per channel plane when W % 4 == 0, scalar otherwise), rounded to bf16 exactly like an NHWC ingest
    const float* xn = x + (long long)n * Ci * plane;
    const bool vec_ok = (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    constexpr int kQuads = kStemPatchCols / 4;
    for (int item = threadIdx.x; item < kStemPatchRows * kQuads; item += 32 * kStemWarps) {
        const int pr = item / kQuads, q = item - pr * kQuads;
        const int hi = 2 * h0 - 1 + pr;
        const int wi0 = 2 * w0 - 4 + 4 * q;
        float v[4][4];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
#pragma unroll
            for (int e = 0; e < 4; ++e) v[ci][e] = 0.f;
        if (hi >= 0 && hi < H) {
            const float* src0 = xn + (long long)hi * W + wi0;
            if (vec_ok && wi0 >= 0 && wi0 + 3 < W) {
                float4 f[CI];
#pragma unroll
                for (int ci = 0; ci < CI; ++ci) f[ci] = __ldg(reinterpret_cast<const float4*>(src0 + ci * plane));
#pragma unroll
                for (int ci = 0; ci < CI; ++ci) {
                    v[ci][0] = f[ci].x; v[ci][1] = f[ci].y; v[ci][2] = f[ci].z; v[ci][3] = f[ci].w;
                }
            } else {
#pragma unroll
                for (int ci = 0; ci < CI; ++ci)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (wi0 + e >= 0 && wi0 + e < W) v[ci][e] = __ldg(src0 + ci * plane + e);
            }
        }
        // 4 pixels x 8 bytes = two 16-byte shared stores
        uint4* dst = reinterpret_cast<uint4*>(patch + pr * kStemPatchCols + 4 * q);
        dst[0] = make_uint4(pack_bf16x2(v[0][0], v[1][0]), pack_bf16x2(v[2][0], v[3][0]),
                            pack_bf16x2(v[0][1], v[1][1]), pack_bf16x2(v[2][1], v[3][1]));
        dst[1] = make_uint4(pack_bf16x2(v[0][2], v[1][2]), pack_bf16x2(v[2][2], v[3][2]),
                            pack_bf16x2(v[0][3], v[1][3]), pack_bf16x2(v[2][3], v[3][3]));
    }
    __syncthreads();

    const int ho = h0 + warp;
    if (ho >= Ho) return;
    // fragment word offsets (in 32-bit words) of this thread's k slots relative to patch pixel (2*warp, 2*wl + 3):
    // k = 16*ks + 2t (+8): tap = k / 4 -> (r, s), channel pair = (k / 2) & 1
    int woff[KSTEPS][2];
    bool kval[KSTEPS][2];
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int k = 16 * ks + 2 * t + hh * 8;
            const int tap = k >> 2, r = tap / 3, s2 = tap - 3 * r;
            kval[ks][hh] = tap < 9;
            woff[ks][hh] = kval[ks][hh] ? ((r * kStemPatchCols + s2) * 2 + ((k >> 1) & 1)) : 0;
        }
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(patch) + (2 * warp * kStemPatchCols + 3) * 2;
    __nv_bfloat16* yrow = y + ((long long)n * Ho + ho) * Wo * y_cstride + y_coff;

#pragma unroll 1
    for (int pass_ = 0; pass_ < PASSES; ++pass_) {
    int pass = pass_;
    asm volatile("" : "+r"(pass));   // opaque: keeps a 2-pass loop from being unrolled back into one 144-register body
    if (pass > 0) load_weights(pass);
#pragma unroll 1
    for (int mt = 0; mt < kStemRowPixels / 16; ++mt) {
        const int wl0 = mt * 16;
        if (w0 + wl0 >= Wo) break;
        uint32_t a[KSTEPS][4];
#pragma unroll
        for (int half = 0; half < 2; ++half) {      // fragment rows g and g + 8
            const uint32_t* px = pw + (wl0 + g + half * 8) * 4;  // 2 input columns per output pixel, 2 words each
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                a[ks][half] = kval[ks][0] ? px[woff[ks][0]] : 0u;      // k = 2t, 2t+1
                a[ks][2 + half] = kval[ks][1] ? px[woff[ks][1]] : 0u;  // k = 2t+8, 2t+9
            }
        }
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            acc[nt][0] = bia[nt][0]; acc[nt][1] = bia[nt][1];
            acc[nt][2] = bia[nt][0]; acc[nt][3] = bia[nt][1];
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) mma_bf16_16816(acc[nt], a[ks], bfrag[ks][nt][0], bfrag[ks][nt][1]);
        }
        // c0,c1: row g, cols 2t,2t+1;  c2,c3: row g+8.  Pairs of n-tiles are transposed inside each lane quad (items:
        // (g, nt), (g, nt+1), (g+8, nt), (g+8, nt+1); lane t ends up with the 16 bytes of item t), so a lane stores 8
        // channels at once and a quad writes 2 x 32 contiguous bytes instead of sixteen scattered 4-byte stores
        const int swo = w0 + wl0 + g + 8 * (t >> 1);
        const bool st_ok = swo < Wo;
        __nv_bfloat16* dst = yrow + (long long)swo * y_cstride + pass * NT * 8 + 8 * (t & 1);
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
            float e[2][4];
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int i = 0; i < 4; ++i) e[q][i] = act ? silu_fast(acc[2 * np + q][i]) : acc[2 * np + q][i];
            uint32_t v0 = pack_bf16x2(e[0][0], e[0][1]), v1 = pack_bf16x2(e[1][0], e[1][1]);
            uint32_t v2 = pack_bf16x2(e[0][2], e[0][3]), v3 = pack_bf16x2(e[1][2], e[1][3]);
            {
                const uint32_t s0 = (t & 1) ? v0 : v1, s1 = (t & 1) ? v2 : v3;
                const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
                if (t & 1) { v0 = r0; v2 = r1; } else { v1 = r0; v3 = r1; }
            }
            {
                const uint32_t s0 = (t & 2) ? v0 : v2, s1 = (t & 2) ? v1 : v3;
                const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
                if (t & 2) { v0 = r0; v1 = r1; } else { v2 = r0; v3 = r1; }
            }
            if (st_ok) *reinterpret_cast<uint4*>(dst + np * 16) = make_uint4(v0, v1, v2, v3);
        }
    }
    }
}

template <int CO>
static int launch_stem(const float* x, int n, int ci, int h, int w, const __nv_bfloat16* wp, int ci_pad,
                       const float* bias, __nv_bfloat16* yp, long long cs, int coff, int Ho, int Wo, int act,
                       cudaStream_t s) {
    dim3 grid((unsigned)ceil_div(Wo, kStemRowPixels), (unsigned)ceil_div(Ho, kStemWarps), (unsigned)n);
    dim3 block(32 * kStemWarps, 1, 1);
#define YL_STEM(CI_)                                                                                              \
    YL_CUDA(launch_kernel(stem_conv_kernel<CO, CI_>, grid, block, 0, s, x, n, h, w, wp, ci_pad, bias, yp, cs, coff, Ho, \
                          Wo, act))
    switch (ci) {
        case 1: YL_STEM(1); break;
        case 2: YL_STEM(2); break;
        case 3: YL_STEM(3); break;
        default: YL_STEM(4); break;
    }
#undef YL_STEM
    YL_LAUNCH_OK("stem_conv_kernel");
    return YL_OK;
}

}  // namespace yl

extern "C" int yl_stem_conv(const float* x_nchw, int n, int ci, int h, int w, const void* w_packed, int ci_pad,
                            const float* bias, const yl_tensor* y, int act, void* stream) {
    YL_CHECK(x_nchw && w_packed && bias && y && y->data, YL_ERR_ARG, "null pointer");
    YL_CHECK(ci >= 1 && ci <= 4, YL_ERR_UNSUPPORTED, "stem kernel takes 1..4 input channels");
    const int Ho = (h + 2 - 3) / 2 + 1, Wo = (w + 2 - 3) / 2 + 1;
    YL_CHECK(y->dtype == YL_BF16 && y->n == n && y->h == Ho && y->w == Wo, YL_ERR_ARG, "stem output shape mismatch");
    YL_CHECK(y->coff % 8 == 0 && y->cstride % 8 == 0, YL_ERR_ARG, "stem output needs 8-channel alignment");
    YL_CHECK((long long)ci * h * w < (1ll << 31), YL_ERR_ARG, "image too large");
    YL_CHECK(n <= 65535, YL_ERR_ARG, "batch too large for one launch");
    __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y->data);
    const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(w_packed);
    cudaStream_t s = (cudaStream_t)stream;
    switch (y->c) {
        case 16: return yl::launch_stem<16>(x_nchw, n, ci, h, w, wp, ci_pad, bias, yp, y->cstride, y->coff, Ho, Wo, act, s);
        case 32: return yl::launch_stem<32>(x_nchw, n, ci, h, w, wp, ci_pad, bias, yp, y->cstride, y->coff, Ho, Wo, act, s);
        case 48: return yl::launch_stem<48>(x_nchw, n, ci, h, w, wp, ci_pad, bias, yp, y->cstride, y->coff, Ho, Wo, act, s);
        case 64: return yl::launch_stem<64>(x_nchw, n, ci, h, w, wp, ci_pad, bias, yp, y->cstride, y->coff, Ho, Wo, act, s);
        case 96: return yl::launch_stem<96>(x_nchw, n, ci, h, w, wp, ci_pad, bias, yp, y->cstride, y->coff, Ho, Wo, act, s);
        default:
            yl::set_error("stem kernel is built for 16/32/48/64/96 output channels (yolo11 n/s/-/m,l/x), got %d", y->c);
            return YL_ERR_UNSUPPORTED;
    }
}
