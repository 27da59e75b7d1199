// The first two layers of the network as ONE kernel: image ingest + Conv(3, C0, 3, 2) + Conv(C0, C1, 3, 2).
//
//   reference: predictor.py:81-84 (`.to(device).float()`), cfg/yolo11.yaml:17-18 (layers 0 and 1),
//              Conv.forward nn/modules/conv.py:47-49 twice.
//
// Layer by layer these are the two most expensive launches of yolo11n at bs = 64: the stem writes a 16-channel
// 320 x 320 map (210 MB) that layer 1 immediately re-reads through 32-byte TMA rows (TMA-row-rate bound, 2.3-2.9
// TB/s).  Fused, the fp32 NCHW image is read once (315 MB) and only the 160 x 160 x 32 map is written (105 MB):
// the 210 MB intermediate lives in shared memory, as bf16, i.e. with the same rounding as the unfused path.
//
// CTA = 8 x 16 output pixels of layer 1:
//   stage 0  the 35 x 72 input patch (3 planes, fp32, coalesced float4 loads) -> bf16 [row][col][4 ch] in smem
//   stage 1  layer 0 on the 17 x 33 halo region: 36 m-tiles (16 px) x 2 n-tiles x 3 k-steps of mma.sync.m16n8k16,
//            K = tap x 4 (channel-padded) = 36 -> 48; bias + SiLU; pixels outside the layer-0 map are forced to
//            ZERO (they are layer 1's zero padding); stored bf16, de-interleaved by column parity so that the
//            stride-2 reads of stage 2 become unit-stride (conflict-free ldmatrix rows)
//   stage 2  layer 1: warp = output row, 9 taps = 9 k-steps of 16 channels, 4 n-tiles; bias + SiLU; 4x4 quad
//            transposes so each lane stores 16 B (64 contiguous bytes per pixel)
// Persistent grid over tiles; weights / biases staged once per CTA before the PDL dependency wait.
#include <cuda_fp16.h>

#include "common.cuh"

namespace yl {

struct StemFusedParams {
    const void* x;                // (N, CI, H, W) fp32 | fp16 | uint8 (image bytes, /255)
    int N, CI, H, W;
    const __nv_bfloat16* w0;      // [16][9 * ci_pad0]
    int ci_pad0;
    const float* b0;
    const __nv_bfloat16* w1;      // [32][9 * 16]
    const float* b1;
    __nv_bfloat16* y;             // (N, H/4, W/4, 32) slice
    long long y_cstride;
    int y_coff;
    int H0, W0, H1, W1;           // layer-0 / layer-1 output sizes
    int act0, act1;
    int tiles_w, tiles_h, total_tiles;
};

__device__ __forceinline__ void sf_ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void sf_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Four consecutive pixels of one channel plane, as loaded (16 / 8 / 4 bytes) and widened to fp32.  uint8 is the
// predictor's image-byte path: float(b) / 255 in fp32 with a true division (predictor.py:84), then the same bf16
// rounding as every other activation.
template <typename T> struct SfQuad;
template <> struct SfQuad<float> {
    using raw = float4;
    static __device__ __forceinline__ raw zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ raw load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ float4 widen(raw v) { return v; }
    static __device__ __forceinline__ float one(const float* p) { return __ldg(p); }
};
template <> struct SfQuad<__half> {
    using raw = uint2;
    static __device__ __forceinline__ raw zero() { return make_uint2(0u, 0u); }
    static __device__ __forceinline__ raw load(const __half* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
    static __device__ __forceinline__ float4 widen(raw v) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ float one(const __half* p) { return __half2float(*p); }
};
template <> struct SfQuad<uint8_t> {
    using raw = uint32_t;
    static __device__ __forceinline__ raw zero() { return 0u; }
    static __device__ __forceinline__ raw load(const uint8_t* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
    static __device__ __forceinline__ float4 widen(raw v) {
        return make_float4(__fdiv_rn((float)(v & 0xffu), 255.f), __fdiv_rn((float)((v >> 8) & 0xffu), 255.f),
                           __fdiv_rn((float)((v >> 16) & 0xffu), 255.f), __fdiv_rn((float)(v >> 24), 255.f));
    }
    static __device__ __forceinline__ float one(const uint8_t* p) { return __fdiv_rn((float)__ldg(p), 255.f); }
};

constexpr int SF_TH = 8, SF_TW = 16;                 // layer-1 output tile
constexpr int SF_R0 = 2 * SF_TH + 1, SF_C0 = 2 * SF_TW + 1;   // 17 x 33 layer-0 pixels
constexpr int SF_RI = 2 * SF_R0 + 1;                 // 35 input rows
constexpr int SF_CI = 72;                            // input columns staged (16-byte aligned start, covers 67)
constexpr int SF_M0 = (SF_R0 * SF_C0 + 15) / 16;     // 36 m-tiles of stage 1
constexpr int SF_EC = (SF_C0 + 1) / 2;               // 17 even / (16 odd) columns per parity plane
constexpr int SF_P0 = 48;                            // bytes per layer-0 pixel in smem (32 data + 16 pad)
constexpr int SF_W1S = 9 * 16 * 2 + 16;              // layer-1 weight row pitch (288 + 16: odd number of 16-B units)
constexpr int SF_OFF_IN = 0;                                      // uint2 [35][72]
constexpr int SF_OFF_L0 = SF_OFF_IN + SF_RI * SF_CI * 8;           // 2 planes x [17][17] x 48 B
constexpr int SF_OFF_W1 = SF_OFF_L0 + 2 * SF_R0 * SF_EC * SF_P0;
constexpr int SF_OFF_B = SF_OFF_W1 + 32 * SF_W1S;                  // b0[16], b1[32] fp32
constexpr int SF_OFF_TAB = SF_OFF_B + 48 * 4;                      // uint2 [36 m-tiles][16 rows]: stage-1 geometry
constexpr int SF_SMEM = SF_OFF_TAB + SF_M0 * 16 * 8;

template <typename TIn>
__global__ void __launch_bounds__(256, 2) stem_fused_kernel(const StemFusedParams p) {
    using Q = SfQuad<TIn>;
    const TIn* px_in = reinterpret_cast<const TIn*>(p.x);
    extern __shared__ __align__(16) uint8_t sf_smem[];
    griddep_launch_dependents();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    uint2* s_in = reinterpret_cast<uint2*>(sf_smem + SF_OFF_IN);
    const uint32_t sL0 = smem_u32(sf_smem + SF_OFF_L0), sW1 = smem_u32(sf_smem + SF_OFF_W1);
    float* sb0 = reinterpret_cast<float*>(sf_smem + SF_OFF_B);
    float* sb1 = sb0 + 16;

    // ---- constants: layer-1 weights -> smem (zero padded rows), biases, layer-0 B fragments -> registers
    for (int i = threadIdx.x; i < 32 * (SF_W1S / 16); i += blockDim.x) {
        const int r = i / (SF_W1S / 16), u = i - r * (SF_W1S / 16);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (u < 18) v = __ldg(reinterpret_cast<const uint4*>(p.w1 + (size_t)r * 144) + u);
        *reinterpret_cast<uint4*>(sf_smem + SF_OFF_W1 + r * SF_W1S + u * 16) = v;
    }
    if (threadIdx.x < 16) sb0[threadIdx.x] = __ldg(p.b0 + threadIdx.x);
    if (threadIdx.x < 32) sb1[threadIdx.x] = __ldg(p.b1 + threadIdx.x);
    // layer 0: K order k = tap * 4 + ci (ci padded to 4), K = 36 -> 3 k-steps.  B fragments:
    // b0 = {W[k0][n], W[k0+1][n]}, b1 = {W[k0+8][n], W[k0+9][n]}, k0 = 16 ks + 2t, n = 8 nt + g
    uint32_t bw0[3][2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
        const __nv_bfloat16* wrow = p.w0 + (size_t)(nt * 8 + g) * 9 * p.ci_pad0;
#pragma unroll
        for (int ks = 0; ks < 3; ++ks)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t v = 0;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 16 * ks + 2 * t + e + hh * 8;
                    const int tap = k >> 2, ci = k & 3;
                    uint32_t bits = 0;
                    if (tap < 9 && ci < p.CI) bits = (uint32_t)__bfloat16_as_ushort(wrow[tap * p.ci_pad0 + ci]);
                    v |= bits << (16 * e);
                }
                bw0[ks][nt][hh] = v;
            }
    }
    // smem word offsets of this thread's k slots relative to the patch pixel of its output pixel (stage 1 A operand);
    // slots beyond tap 8 meet zero B fragments, so they may read any valid word
    int woff[3][2];
#pragma unroll
    for (int ks = 0; ks < 3; ++ks)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int k = 16 * ks + 2 * t + hh * 8;
            const int tap = k >> 2, r = tap / 3, s2 = tap - 3 * r;
            woff[ks][hh] = tap < 9 ? ((r * SF_CI + s2) * 2 + ((k >> 1) & 1)) : 0;
        }
    // stage-1 geometry is the same for every tile: per (m-tile, row) the input-patch word offset of the pixel, its
    // destination in the parity-split layer-0 patch and its (row, col) for the border mask
    uint2* s_tab = reinterpret_cast<uint2*>(sf_smem + SF_OFF_TAB);
    for (int i = threadIdx.x; i < SF_M0 * 16; i += blockDim.x) {
        const bool valid = i < SF_R0 * SF_C0;
        const int idx = valid ? i : SF_R0 * SF_C0 - 1;
        const int er = idx / SF_C0, ec = idx - er * SF_C0;
        const uint32_t aoff = (uint32_t)(((2 * er) * SF_CI + 2 * ec + 1) * 2);
        const uint32_t doff = (uint32_t)((((ec & 1) * SF_R0 + er) * SF_EC + (ec >> 1)) * SF_P0);
        s_tab[i] = make_uint2(aoff | ((uint32_t)(valid ? er : 255) << 16) | ((uint32_t)ec << 24), doff);
    }
    griddep_wait();

    const long long plane = (long long)p.H * p.W;
    const bool vec_ok = (p.W & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.x) & (4 * sizeof(TIn) - 1)) == 0);
    constexpr int QUADS = SF_CI / 4;
    constexpr int ITEMS = SF_RI * QUADS;                       // 630 quads of 4 pixels
    constexpr int ROUNDS = (ITEMS + 255) / 256;
    typename Q::raw pf[ROUNDS][3];                             // the next tile's patch, in flight across stage 2

    // Stage 0 is split in two so that its HBM round trip hides behind stage 2 of the previous tile: `patch_issue`
    // only issues the loads (quads are 16-byte aligned in the image, so with W % 4 == 0 each one is entirely inside
    // or entirely outside), `patch_store` rounds to bf16 and writes [row][col][4 ch].
    auto patch_issue = [&](int tile) {
        const int ow0 = (tile % p.tiles_w) * SF_TW;
        const int oh0 = ((tile / p.tiles_w) % p.tiles_h) * SF_TH;
        const int n = tile / (p.tiles_w * p.tiles_h);
        const int ir0 = 4 * oh0 - 3, ic0 = 4 * ow0 - 4;   // input coords of patch (0, 0)
        const TIn* xn = px_in + (long long)n * p.CI * plane;
#pragma unroll
        for (int j = 0; j < ROUNDS; ++j) {
            const int item = threadIdx.x + j * 256;
            const int pr = item / QUADS, q = item - pr * QUADS;
            const int hi = ir0 + pr, wi0 = ic0 + 4 * q;
            const bool ok = item < ITEMS && hi >= 0 && hi < p.H && wi0 >= 0 && wi0 < p.W;
            const TIn* src0 = xn + (long long)(ok ? hi : 0) * p.W + (ok ? wi0 : 0);
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                pf[j][ci] = Q::zero();
                if (ok && ci < p.CI) pf[j][ci] = Q::load(src0 + ci * plane);
            }
        }
    };
    auto patch_store = [&]() {
#pragma unroll
        for (int j = 0; j < ROUNDS; ++j) {
            const int item = threadIdx.x + j * 256;
            if (item >= ITEMS) continue;
            const int pr = item / QUADS, q = item - pr * QUADS;
            uint4* dst = reinterpret_cast<uint4*>(s_in + pr * SF_CI + 4 * q);
            const float4 c0 = Q::widen(pf[j][0]), c1 = Q::widen(pf[j][1]), c2 = Q::widen(pf[j][2]);
            dst[0] = make_uint4(pack_bf16x2(c0.x, c1.x), pack_bf16x2(c2.x, 0.f), pack_bf16x2(c0.y, c1.y), pack_bf16x2(c2.y, 0.f));
            dst[1] = make_uint4(pack_bf16x2(c0.z, c1.z), pack_bf16x2(c2.z, 0.f), pack_bf16x2(c0.w, c1.w), pack_bf16x2(c2.w, 0.f));
        }
    };
    // general shapes (W % 4 != 0 or an unaligned base): element-wise, synchronous
    auto patch_slow = [&](int tile) {
        const int ow0 = (tile % p.tiles_w) * SF_TW;
        const int oh0 = ((tile / p.tiles_w) % p.tiles_h) * SF_TH;
        const int n = tile / (p.tiles_w * p.tiles_h);
        const int ir0 = 4 * oh0 - 3, ic0 = 4 * ow0 - 4;
        const TIn* xn = px_in + (long long)n * p.CI * plane;
        for (int item = threadIdx.x; item < ITEMS; item += blockDim.x) {
            const int pr = item / QUADS, q = item - pr * QUADS;
            const int hi = ir0 + pr, wi0 = ic0 + 4 * q;
            float v[3][4];
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    v[ci][e] = 0.f;
                    if (ci < p.CI && hi >= 0 && hi < p.H && wi0 + e >= 0 && wi0 + e < p.W)
                        v[ci][e] = Q::one(xn + ci * plane + (long long)hi * p.W + wi0 + e);
                }
            uint4* dst = reinterpret_cast<uint4*>(s_in + pr * SF_CI + 4 * q);
            dst[0] = make_uint4(pack_bf16x2(v[0][0], v[1][0]), pack_bf16x2(v[2][0], 0.f), pack_bf16x2(v[0][1], v[1][1]),
                                pack_bf16x2(v[2][1], 0.f));
            dst[1] = make_uint4(pack_bf16x2(v[0][2], v[1][2]), pack_bf16x2(v[2][2], 0.f), pack_bf16x2(v[0][3], v[1][3]),
                                pack_bf16x2(v[2][3], 0.f));
        }
    };

    if ((int)blockIdx.x < p.total_tiles) {
        if (vec_ok) {
            patch_issue(blockIdx.x);
            patch_store();
        } else {
            patch_slow(blockIdx.x);
        }
    }
    __syncthreads();
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int ow0 = (tile % p.tiles_w) * SF_TW;
        const int oh0 = ((tile / p.tiles_w) % p.tiles_h) * SF_TH;
        const int n = tile / (p.tiles_w * p.tiles_h);
        const int next = tile + (int)gridDim.x;
        // layer-0 pixel (pr, pc) of the halo region reads patch rows 2pr + dr and columns 2pc + dc + 1

        // ================= stage 1: layer 0 on the 17 x 33 halo region =================
        {
            const uint32_t* pw = reinterpret_cast<const uint32_t*>(s_in);
            // border mask of the layer-0 map in patch coordinates (uniform per tile)
            const int er_lo = oh0 == 0 ? 1 : 0, er_hi = min(SF_R0, p.H0 - (2 * oh0 - 1));
            const int ec_lo = ow0 == 0 ? 1 : 0, ec_hi = min(SF_C0, p.W0 - (2 * ow0 - 1));
            for (int mt = warp; mt < SF_M0; mt += 8) {
                uint32_t a[3][4];
                uint2 te[2];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    te[half] = s_tab[mt * 16 + g + half * 8];
                    const uint32_t* px = pw + (te[half].x & 0xffffu);
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks) {
                        a[ks][half] = px[woff[ks][0]];
                        a[ks][2 + half] = px[woff[ks][1]];
                    }
                }
                float acc[2][4];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const float2 bb = *reinterpret_cast<const float2*>(sb0 + nt * 8 + 2 * t);
                    acc[nt][0] = bb.x; acc[nt][1] = bb.y; acc[nt][2] = bb.x; acc[nt][3] = bb.y;
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks) sf_mma(acc[nt], a[ks], bw0[ks][nt][0], bw0[ks][nt][1]);
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int er = (int)((te[half].x >> 16) & 0xffu), ec = (int)(te[half].x >> 24);
                    if (er == 255) continue;                       // row beyond the 17 x 33 region
                    const bool in = er >= er_lo && er < er_hi && ec >= ec_lo && ec < ec_hi;
                    const uint32_t dst = sL0 + te[half].y + (uint32_t)(4 * t);
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        float v0 = acc[nt][half * 2], v1 = acc[nt][half * 2 + 1];
                        if (p.act0) {
                            v0 = silu_fast(v0);
                            v1 = silu_fast(v1);
                        }
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + (uint32_t)(nt * 16)), "r"(in ? pack_bf16x2(v0, v1) : 0u)
                                     : "memory");
                    }
                }
            }
        }
        __syncthreads();
        if (vec_ok && next < p.total_tiles) patch_issue(next);   // in flight during stage 2

        // ================= stage 2: layer 1, warp = output row =================
        {
            const int r = warp;
            const int lrow = lane & 15, khalf = lane >> 4;
            float acc[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float2 bb = *reinterpret_cast<const float2*>(sb1 + nt * 8 + 2 * t);
                acc[nt][0] = bb.x; acc[nt][1] = bb.y; acc[nt][2] = bb.x; acc[nt][3] = bb.y;
            }
            const uint32_t brow = sW1 + (uint32_t)(((lane >> 4) * 8 + (lane & 7)) * SF_W1S + 8 * ((lane >> 3) & 1) * 2);
#pragma unroll
            for (int t3 = 0; t3 < 9; t3 += 3) {       // three taps' operands in flight at a time
                uint32_t a[3][4], b[3][2][4];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int tap = t3 + j, dr = tap / 3, dc = tap - 3 * dr;
                    // layer-0 patch pixel (2r + dr, 2 lrow + dc): parity plane dc & 1, column lrow + (dc >> 1)
                    const uint32_t arow = sL0 + (uint32_t)((((dc & 1) * SF_R0 + 2 * r + dr) * SF_EC + lrow + (dc >> 1)) * SF_P0 +
                                                           khalf * 16);
                    sf_ldsm_x4(arow, a[j][0], a[j][1], a[j][2], a[j][3]);
#pragma unroll
                    for (int np = 0; np < 2; ++np)
                        sf_ldsm_x4(brow + (uint32_t)(np * 16 * SF_W1S + tap * 32), b[j][np][0], b[j][np][1], b[j][np][2],
                                   b[j][np][3]);
                }
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int np = 0; np < 2; ++np) {
                        sf_mma(acc[2 * np], a[j], b[j][np][0], b[j][np][1]);
                        sf_mma(acc[2 * np + 1], a[j], b[j][np][2], b[j][np][3]);
                    }
            }
            // after the quad transpose lane t stores pixel col g + 8 (t >> 1), channels 16 np + 8 (t & 1) .. +8
            const int oh = oh0 + r, ow = ow0 + g + 8 * (t >> 1);
            const bool st_ok = oh < p.H1 && ow < p.W1;
            __nv_bfloat16* yp = p.y + (((long long)n * p.H1 + oh) * p.W1 + ow) * p.y_cstride + p.y_coff + 8 * (t & 1);
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                float e[2][4];
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int i = 0; i < 4; ++i) e[q][i] = p.act1 ? silu_fast(acc[2 * np + q][i]) : acc[2 * np + q][i];
                uint32_t v0 = pack_bf16x2(e[0][0], e[0][1]);   // (col g,   n-tile 0)
                uint32_t v1 = pack_bf16x2(e[1][0], e[1][1]);   // (col g,   n-tile 1)
                uint32_t v2 = pack_bf16x2(e[0][2], e[0][3]);   // (col g+8, n-tile 0)
                uint32_t v3 = pack_bf16x2(e[1][2], e[1][3]);   // (col g+8, n-tile 1)
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
                if (st_ok) *reinterpret_cast<uint4*>(yp + np * 16) = make_uint4(v0, v1, v2, v3);
            }
        }
        // stage 1 of this tile has finished reading the input patch (barrier above): refill it for the next tile
        if (next < p.total_tiles) {
            if (vec_ok) patch_store();
            else patch_slow(next);
        }
        __syncthreads();   // next patch complete; every warp is done with this tile's layer-0 patch
    }
}

static int g_sf_sms = 148;

// per-device setup (called by yl_init)
int init_stem_fused() {
    int dev = 0;
    YL_CUDA(cudaGetDevice(&dev));
    YL_CUDA(cudaDeviceGetAttribute(&g_sf_sms, cudaDevAttrMultiProcessorCount, dev));
    YL_CUDA(cudaFuncSetAttribute(stem_fused_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM));
    YL_CUDA(cudaFuncSetAttribute(stem_fused_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM));
    YL_CUDA(cudaFuncSetAttribute(stem_fused_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM));
    return YL_OK;
}

}  // namespace yl

extern "C" int yl_stem_fused_supported(int ci, int c0, int c1) { return ci >= 1 && ci <= 3 && c0 == 16 && c1 == 32; }

extern "C" int yl_stem_fused(const void* x_nchw, int x_dtype, int n, int ci, int h, int w, const void* w0, int ci_pad0,
                             const float* b0, int act0, const void* w1, int ci_pad1, const float* b1, int act1,
                             const yl_tensor* y, void* stream) {
    YL_CHECK(x_nchw && w0 && b0 && w1 && b1 && y && y->data, YL_ERR_ARG, "null pointer");
    YL_CHECK(x_dtype == YL_F32 || x_dtype == YL_F16 || x_dtype == YL_U8, YL_ERR_ARG,
             "fused stem reads fp32, fp16 or uint8 images (got dtype %d)", x_dtype);
    YL_CHECK(y->dtype == YL_BF16, YL_ERR_ARG, "stem output must be bf16");
    YL_CHECK(yl_stem_fused_supported(ci, 16, y->c) && ci_pad1 == 16, YL_ERR_UNSUPPORTED,
             "fused stem is built for <= 3 input channels, 16 -> 32 (got ci = %d, c1 = %d, ci_pad1 = %d)", ci, y->c, ci_pad1);
    const int H0 = (h + 2 - 3) / 2 + 1, W0 = (w + 2 - 3) / 2 + 1;
    const int H1 = (H0 + 2 - 3) / 2 + 1, W1 = (W0 + 2 - 3) / 2 + 1;
    YL_CHECK(y->n == n && y->h == H1 && y->w == W1, YL_ERR_ARG, "fused stem output shape mismatch: got (%d,%d,%d) expected (%d,%d,%d)",
             y->n, y->h, y->w, n, H1, W1);
    YL_CHECK(y->coff % 8 == 0 && y->cstride % 8 == 0 && ((uintptr_t)y->data | (uintptr_t)w1) % 16 == 0, YL_ERR_ARG,
             "fused stem needs 8-channel / 16-byte alignment");
    YL_CHECK((long long)ci * h * w < (1ll << 31), YL_ERR_ARG, "image too large");
    yl::StemFusedParams p;
    p.x = x_nchw;
    p.N = n;
    p.CI = ci;
    p.H = h;
    p.W = w;
    p.w0 = reinterpret_cast<const __nv_bfloat16*>(w0);
    p.ci_pad0 = ci_pad0;
    p.b0 = b0;
    p.w1 = reinterpret_cast<const __nv_bfloat16*>(w1);
    p.b1 = b1;
    p.y = reinterpret_cast<__nv_bfloat16*>(y->data);
    p.y_cstride = y->cstride;
    p.y_coff = y->coff;
    p.H0 = H0;
    p.W0 = W0;
    p.H1 = H1;
    p.W1 = W1;
    p.act0 = act0;
    p.act1 = act1;
    p.tiles_w = yl::ceil_div(W1, yl::SF_TW);
    p.tiles_h = yl::ceil_div(H1, yl::SF_TH);
    p.total_tiles = p.tiles_w * p.tiles_h * n;
    const int sms = yl::g_sf_sms;
    int per_sm = (int)((size_t)220 * 1024 / (yl::SF_SMEM + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
    int grid = sms * per_sm;
    if (grid > p.total_tiles) grid = p.total_tiles;
    if (x_dtype == YL_F32)
        YL_CUDA(yl::launch_kernel(yl::stem_fused_kernel<float>, dim3(grid), dim3(256), (size_t)yl::SF_SMEM, (cudaStream_t)stream, p));
    else if (x_dtype == YL_F16)
        YL_CUDA(yl::launch_kernel(yl::stem_fused_kernel<__half>, dim3(grid), dim3(256), (size_t)yl::SF_SMEM, (cudaStream_t)stream, p));
    else
        YL_CUDA(yl::launch_kernel(yl::stem_fused_kernel<uint8_t>, dim3(grid), dim3(256), (size_t)yl::SF_SMEM, (cudaStream_t)stream, p));
    YL_LAUNCH_OK("stem_fused_kernel");
    return YL_OK;
}
