// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM), operands fed by TMA.
//
//   D[M = 128 output pixels, N = co_tile] = sum over taps (r,s) and channel blocks of
//       A[pixels shifted by the tap, kblk channels]  x  W[co_tile, kblk]
//
// * activations are NHWC bf16; the A tile of one tap is a TH x TW spatial patch of one image fetched by a
//   single 4-D TMA box {kblk, TW, TH, 1}; the conv halo and ragged image edges are TMA out-of-bounds
//   zero fill (signed coordinates), so there is no im2col buffer and no padding pass;
// * stride-2 convs read one of four "parity planes" (h%2, w%2) of the input, each its own 4-D tensor map,
//   so a tap is again a dense box;
// * 1x1 convs run in flat mode: the whole (n,h,w) extent is one axis and tiles are 128 consecutive pixels;
// * weights are packed [co][tap][ci] bf16 (K-major), fetched by a 2-D TMA box {kblk, co_tile};
// * both tiles land in the canonical K-major swizzled layout (SW128/64/32 for kblk 64/32/16) that
//   tcgen05.mma reads through shared-memory descriptors; accumulation is fp32 in TMEM;
// * epilogue (4 warps, one TMEM lane quadrant each): tcgen05.ld -> +bias -> SiLU -> +residual -> bf16/f32
//   store into a channel slice of the destination (concat = aliasing), optionally replicated 2x2
//   (nearest upsample fused into the producer).
//
// Warp roles: warp 0 = TMA producer (one lane), warp 1 = TMEM owner + MMA issuer (one lane),
// warps 2..5 = epilogue.  Pipeline: `stages`-deep smem ring with full/empty mbarriers; tcgen05.commit
// releases a stage back to the producer and finally signals the epilogue.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace yl {

struct ConvTcParams {
    CUtensorMap tmA[4];
    CUtensorMap tmB;
    int Ho, Wo, Nimg;            // conv output dims per image, images (flat mode: 1, total pixels, 1)
    int tiles_w, tiles_h, tiles_n;
    int TW, TH, TN;              // A-tile box: TW*TH*TN <= 128 rows (pixels), may span images
    int m_tiles, n_tiles, total_tiles;
    int ksize, stride, pad;
    int ci_pad;                  // K elements per tap in the packed weights
    int kblk, cin_blocks;        // channels per k-iteration, iterations per tap
    int co_tile;                 // UMMA N
    int stages, acc_stages;
    uint32_t tmem_cols;
    uint32_t a_bytes, b_bytes;   // per-stage smem footprint (1024-aligned)
    uint32_t tx_bytes;           // bytes one stage's two TMA boxes deliver
    // epilogue
    void* y;
    long long y_cstride;
    int y_coff, y_c, y_f32;
    const float* bias;
    int act;
    const __nv_bfloat16* res;
    long long res_cstride;
    int res_coff;
    int upsample;
};

constexpr int kConvTcThreads = 192;

// Persistent: gridDim.x CTAs each walk tiles blockIdx.x, +gridDim.x, ...  The TMA producer runs ahead across
// tile boundaries (the smem ring never drains), and with two TMEM accumulator stages the epilogue of tile i
// overlaps the loads and MMAs of tile i+1.
__global__ void __launch_bounds__(kConvTcThreads) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- shared memory carve-up (1024-B aligned: required by the 128-B swizzle atom)
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* sA = base;
    uint8_t* sB = base + (size_t)p.stages * p.a_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + (size_t)p.stages * p.b_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* tfull_bar = empty_bar + p.stages;   // [acc_stages] accumulator ready for the epilogue
    uint64_t* tempty_bar = tfull_bar + 2;         // [acc_stages] accumulator drained by the epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 4);  // one arrive per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, p.tmem_cols);
        tmem_relinquish();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA[0]);
        tma_prefetch_desc(&p.tmB);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int taps = p.ksize * p.ksize;
    const int kiters = taps * p.cin_blocks;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int nt = tile % p.n_tiles;
                int mt = tile / p.n_tiles;
                const int w0 = (mt % p.tiles_w) * p.TW;
                mt /= p.tiles_w;
                const int h0 = (mt % p.tiles_h) * p.TH;
                const int i0 = (mt / p.tiles_h) * p.TN;
                const int n0 = nt * p.co_tile;
                for (int tap = 0; tap < taps; ++tap) {
                    const int r = tap / p.ksize, s = tap - r * p.ksize;
                    const int offh = r - p.pad, offw = s - p.pad;
                    int map = 0, dh = offh, dw = offw;
                    if (p.stride == 2) {
                        const int ph = offh & 1, pw = offw & 1;
                        map = ph * 2 + pw;
                        dh = (offh - ph) >> 1;
                        dw = (offw - pw) >> 1;
                    }
                    for (int cb = 0; cb < p.cin_blocks; ++cb, ++it) {
                        const int st = it % p.stages;
                        const uint32_t ph_bit = (uint32_t)(it / p.stages) & 1u;
                        mbar_wait(&empty_bar[st], ph_bit ^ 1u);
                        mbar_expect_tx(&full_bar[st], p.tx_bytes);
                        tma_load_4d(sA + (size_t)st * p.a_bytes, &p.tmA[map], &full_bar[st], cb * p.kblk, w0 + dw,
                                    h0 + dh, i0);
                        tma_load_2d(sB + (size_t)st * p.b_bytes, &p.tmB, &full_bar[st],
                                    tap * p.ci_pad + cb * p.kblk, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)p.co_tile);
            const uint32_t row_bytes = (uint32_t)p.kblk * 2u;
            const int ksteps = p.kblk / 16;
            int it = 0, lt = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
                const int acc = lt % p.acc_stages;
                const uint32_t acc_ph = (uint32_t)(lt / p.acc_stages) & 1u;
                mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.co_tile);
                for (int ki = 0; ki < kiters; ++ki, ++it) {
                    const int st = it % p.stages;
                    const uint32_t ph_bit = (uint32_t)(it / p.stages) & 1u;
                    mbar_wait(&full_bar[st], ph_bit);
                    tc_fence_after();
                    const uint64_t da = umma_desc_kmajor(smem_u32(sA + (size_t)st * p.a_bytes), row_bytes);
                    const uint64_t db = umma_desc_kmajor(smem_u32(sB + (size_t)st * p.b_bytes), row_bytes);
                    for (int k = 0; k < ksteps; ++k) {
                        // advance 16 bf16 (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
                        umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                  (ki > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[st]);
                }
                umma_commit(&tfull_bar[acc]);
            }
        }
    } else {
        // ================= epilogue =================
        const int q = warp & 3;  // TMEM lane quadrant this warp may read
        const int row = q * 32 + lane;
        const int tw = row % p.TW;
        const int th = (row / p.TW) % p.TH;
        const int tn = row / (p.TW * p.TH);
        int lt = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
            const int nt = tile % p.n_tiles;
            int mt = tile / p.n_tiles;
            const int w = (mt % p.tiles_w) * p.TW + tw;
            mt /= p.tiles_w;
            const int h = (mt % p.tiles_h) * p.TH + th;
            const int n = (mt / p.tiles_h) * p.TN + tn;
            const int n0 = nt * p.co_tile;
            const bool pvalid = (tn < p.TN) && (n < p.Nimg) && (h < p.Ho) && (w < p.Wo);
            const long long pix = ((long long)n * p.Ho + h) * p.Wo + w;

            const int acc = lt % p.acc_stages;
            const uint32_t acc_ph = (uint32_t)(lt / p.acc_stages) & 1u;
            mbar_wait(&tfull_bar[acc], acc_ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.co_tile);

            for (int c0 = 0; c0 < p.co_tile; c0 += 16) {
                uint32_t accv[16];
                __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the predicated stores
                tmem_ld16(taddr + (uint32_t)c0, accv);
                tmem_ld_wait();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int cb = n0 + c0 + half * 8;
                    if (!pvalid || cb >= p.y_c) continue;
                    float v[8];
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + cb));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + cb + 4));
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float x = __uint_as_float(accv[half * 8 + i]) + bb[i];
                        v[i] = p.act ? silu_f(x) : x;
                    }
                    if (p.res) {
                        const uint4 rv =
                            __ldg(reinterpret_cast<const uint4*>(p.res + pix * p.res_cstride + p.res_coff + cb));
                        v[0] += bf16lo_f(rv.x); v[1] += bf16hi_f(rv.x);
                        v[2] += bf16lo_f(rv.y); v[3] += bf16hi_f(rv.y);
                        v[4] += bf16lo_f(rv.z); v[5] += bf16hi_f(rv.z);
                        v[6] += bf16lo_f(rv.w); v[7] += bf16hi_f(rv.w);
                    }
                    const int reps = p.upsample ? 4 : 1;
                    for (int rep = 0; rep < reps; ++rep) {
                        long long opix = pix;
                        if (p.upsample) {
                            const int dy = rep >> 1, dx = rep & 1;
                            opix = ((long long)n * (2 * p.Ho) + (2 * h + dy)) * (2 * p.Wo) + (2 * w + dx);
                        }
                        if (p.y_f32) {
                            float* dst = reinterpret_cast<float*>(p.y) + opix * p.y_cstride + p.y_coff + cb;
                            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                            *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
                        } else {
                            __nv_bfloat16* dst =
                                reinterpret_cast<__nv_bfloat16*>(p.y) + opix * p.y_cstride + p.y_coff + cb;
                            uint4 o;
                            o.x = pack_bf16x2(v[0], v[1]);
                            o.y = pack_bf16x2(v[2], v[3]);
                            o.z = pack_bf16x2(v[4], v[5]);
                            o.w = pack_bf16x2(v[6], v[7]);
                            *reinterpret_cast<uint4*>(dst) = o;
                        }
                    }
                }
            }
            // all TMEM reads of this warp are complete (wait::ld above): hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ host side
static int g_max_dyn_smem = 0;
static int g_num_sms = 148;

int init_conv_tc() {
    int dev = 0;
    YL_CUDA(cudaGetDevice(&dev));
    int max_optin = 0;
    YL_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    g_max_dyn_smem = max_optin;
    YL_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    YL_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    return YL_OK;
}

static bool encode_map(CUtensorMap* m, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, CUtensorMapSwizzle sw) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) {
        set_error("yl_init() was not called (TMA encoder unresolved)");
        return false;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides_bytes, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]", (int)r,
                  rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                  box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return false;
    }
    return true;
}

// A-tile box (TW, TH, TN) with TW*TH*TN <= 128 output pixels, possibly spanning images: minimise the number
// of 128-row MMA tiles needed to cover (Wo, Ho, N); ties prefer wide boxes (longer contiguous runs).
static void choose_patch(int Ho, int Wo, int N, int* TH, int* TW, int* TN) {
    long long best = -1;
    int bw = 1, bh = 1, bn = 1;
    for (int tw = Wo < 128 ? Wo : 128; tw >= 1; --tw) {
        if (tw != Wo && (tw & (tw - 1))) continue;  // full width or a power of two
        for (int th = 1; th * tw <= 128 && th <= Ho; ++th) {
            int tn = 128 / (tw * th);
            if (tn > N) tn = N;
            // TN > 1 stacks the same (tw, th) window of consecutive images: the box is a 4-D hyper-rectangle
            const long long tiles = (long long)ceil_div(Wo, tw) * ceil_div(Ho, th) * ceil_div(N, tn);
            if (best < 0 || tiles < best) {
                best = tiles;
                bw = tw;
                bh = th;
                bn = tn;
            }
        }
    }
    *TW = bw;
    *TH = bh;
    *TN = bn;
}

bool conv_tc_supported(const yl_conv_args* a, char* why, size_t why_len) {
#define NOPE(msg)                                   \
    do {                                            \
        if (why) snprintf(why, why_len, "%s", msg); \
        return false;                               \
    } while (0)
    const yl_tensor& x = a->x;
    const yl_tensor& y = a->y;
    if (x.dtype != YL_BF16) NOPE("input must be bf16");
    if (!(a->k == 1 || a->k == 3)) NOPE("k must be 1 or 3");
    if (!(a->stride == 1 || a->stride == 2)) NOPE("stride must be 1 or 2");
    if (x.c % 8 || x.coff % 8 || x.cstride % 8) NOPE("input channels/offset/stride must be multiples of 8");
    if (y.c % 8 || y.coff % 8 || y.cstride % 8) NOPE("output channels/offset/stride must be multiples of 8");
    if (a->ci_pad % 8 || a->ci_pad < x.c) NOPE("ci_pad must be a multiple of 8 and >= x.c");
    if (a->co_pad % 8 || a->co_pad < y.c) NOPE("co_pad must be a multiple of 8 and >= y.c");
    if (a->stride == 2 && ((x.h | x.w) & 1)) NOPE("stride 2 needs even input dims");
    if (a->res.data && (a->res.c % 8 || a->res.coff % 8 || a->res.cstride % 8 || a->res.dtype != YL_BF16))
        NOPE("residual must be bf16 with 8-channel alignment");
    if (((uintptr_t)x.data | (uintptr_t)y.data | (uintptr_t)a->w | (uintptr_t)a->bias | (uintptr_t)a->res.data) & 15)
        NOPE("pointers must be 16-byte aligned");
    return true;
#undef NOPE
}

int launch_conv_tc(const yl_conv_args* a, cudaStream_t stream) {
    char why[128];
    YL_CHECK(conv_tc_supported(a, why, sizeof(why)), YL_ERR_UNSUPPORTED, "tcgen05 conv unsupported: %s", why);
    const yl_tensor& x = a->x;
    const yl_tensor& y = a->y;
    const int pad = a->k / 2;
    const int Ho = (x.h + 2 * pad - a->k) / a->stride + 1;
    const int Wo = (x.w + 2 * pad - a->k) / a->stride + 1;
    const int up = a->upsample2x ? 2 : 1;
    YL_CHECK(y.n == x.n && y.h == Ho * up && y.w == Wo * up, YL_ERR_ARG,
             "conv output dims mismatch: got (%d,%d,%d) expected (%d,%d,%d)", y.n, y.h, y.w, x.n, Ho * up, Wo * up);

    ConvTcParams p;
    memset(&p, 0, sizeof(p));
    p.ksize = a->k;
    p.stride = a->stride;
    p.pad = pad;
    p.ci_pad = a->ci_pad;
    p.kblk = x.c >= 64 ? 64 : (x.c >= 32 ? 32 : 16);
    p.cin_blocks = ceil_div(x.c, p.kblk);
    const CUtensorMapSwizzle sw = p.kblk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                               : (p.kblk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);

    // N tiling
    const int co16 = ceil_div(y.c, 16) * 16;
    const int n_tiles = ceil_div(co16, 256);
    p.co_tile = ceil_div(ceil_div(co16, n_tiles), 16) * 16;

    // M tiling + activation tensor maps
    __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(x.data) + x.coff;
    const uint64_t es = 2;
    const bool flat = (a->k == 1 && a->stride == 1 && !a->upsample2x);
    if (flat) {
        const uint64_t M = (uint64_t)x.n * x.h * x.w;
        YL_CHECK(M < (1ull << 31), YL_ERR_ARG, "too many pixels");
        p.Ho = 1;
        p.Wo = (int)M;
        p.Nimg = 1;
        p.TH = 1;
        p.TW = 128;
        p.TN = 1;
        p.tiles_h = 1;
        p.tiles_n = 1;
        p.tiles_w = ceil_div((int)M, 128);
        uint64_t dims[4] = {(uint64_t)x.c, M, 1, 1};
        uint64_t str[3] = {(uint64_t)x.cstride * es, (uint64_t)x.cstride * es * M, (uint64_t)x.cstride * es * M};
        uint32_t box[4] = {(uint32_t)p.kblk, 128, 1, 1};
        if (!encode_map(&p.tmA[0], xb, 4, dims, str, box, sw)) return YL_ERR_CUDA;
    } else {
        p.Ho = Ho;
        p.Wo = Wo;
        p.Nimg = x.n;
        choose_patch(Ho, Wo, x.n, &p.TH, &p.TW, &p.TN);
        p.tiles_h = ceil_div(Ho, p.TH);
        p.tiles_w = ceil_div(Wo, p.TW);
        p.tiles_n = ceil_div(x.n, p.TN);
        uint32_t box[4] = {(uint32_t)p.kblk, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
        if (a->stride == 1) {
            uint64_t dims[4] = {(uint64_t)x.c, (uint64_t)x.w, (uint64_t)x.h, (uint64_t)x.n};
            uint64_t str[3] = {(uint64_t)x.cstride * es, (uint64_t)x.cstride * es * x.w,
                               (uint64_t)x.cstride * es * x.w * x.h};
            if (!encode_map(&p.tmA[0], xb, 4, dims, str, box, sw)) return YL_ERR_CUDA;
        } else {
            for (int ph = 0; ph < 2; ++ph)
                for (int pw = 0; pw < 2; ++pw) {
                    uint64_t dims[4] = {(uint64_t)x.c, (uint64_t)x.w / 2, (uint64_t)x.h / 2, (uint64_t)x.n};
                    uint64_t str[3] = {(uint64_t)x.cstride * es * 2, (uint64_t)x.cstride * es * x.w * 2,
                                       (uint64_t)x.cstride * es * x.w * x.h};
                    __nv_bfloat16* b = xb + ((size_t)ph * x.w + pw) * x.cstride;
                    if (!encode_map(&p.tmA[ph * 2 + pw], b, 4, dims, str, box, sw)) return YL_ERR_CUDA;
                }
        }
    }
    // weights: [co_pad][k*k*ci_pad] bf16
    {
        const uint64_t K = (uint64_t)a->k * a->k * a->ci_pad;
        uint64_t dims[2] = {K, (uint64_t)a->co_pad};
        uint64_t str[1] = {K * es};
        uint32_t box[2] = {(uint32_t)p.kblk, (uint32_t)p.co_tile};
        if (!encode_map(&p.tmB, const_cast<void*>(a->w), 2, dims, str, box, sw)) return YL_ERR_CUDA;
    }

    p.a_bytes = 128u * p.kblk * 2u;
    p.b_bytes = ((uint32_t)p.co_tile * p.kblk * 2u + 1023u) & ~1023u;
    p.tx_bytes = (uint32_t)(p.TW * p.TH * p.TN) * p.kblk * 2u + (uint32_t)p.co_tile * p.kblk * 2u;
    p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    p.n_tiles = n_tiles;
    p.total_tiles = p.m_tiles * n_tiles;
    // two accumulator stages whenever they fit the 512 TMEM columns; two CTAs per SM when TMEM and smem allow
    p.acc_stages = (2 * p.co_tile <= 512) ? 2 : 1;
    uint32_t cols = 32;
    while ((int)cols < p.acc_stages * p.co_tile) cols <<= 1;
    p.tmem_cols = cols;
    const int ctas_per_sm = (cols <= 256) ? 2 : 1;
    const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
    const uint32_t budget = (uint32_t)(ctas_per_sm == 2 ? 108 * 1024 : 200 * 1024);
    int stages = (int)(budget / stage_bytes);
    if (stages > 12) stages = 12;
    if (stages < 2) stages = 2;
    p.stages = stages;
    const size_t smem = 1024 + (size_t)stages * stage_bytes + (2 * stages + 4) * 8 + 16;
    YL_CHECK((int)smem <= g_max_dyn_smem, YL_ERR_UNSUPPORTED, "conv tile needs %zu B smem (max %d)", smem,
             g_max_dyn_smem);

    p.y = y.data;
    p.y_cstride = y.cstride;
    p.y_coff = y.coff;
    p.y_c = y.c;
    p.y_f32 = (y.dtype == YL_F32);
    p.bias = a->bias;
    p.act = a->act;
    p.res = reinterpret_cast<const __nv_bfloat16*>(a->res.data);
    p.res_cstride = a->res.cstride;
    p.res_coff = a->res.coff;
    p.upsample = a->upsample2x ? 1 : 0;

    int grid = g_num_sms * ctas_per_sm;
    if (grid > p.total_tiles) grid = p.total_tiles;
    conv_tc_kernel<<<grid, kConvTcThreads, smem, stream>>>(p);
    YL_LAUNCH_OK("conv_tc_kernel");
    return YL_OK;
}

}  // namespace yl
