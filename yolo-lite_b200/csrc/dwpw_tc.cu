// DWConv 3x3 + Conv 1x1 as ONE kernel: the class branch of the Detect head (head.py:46-47, 59-65:
// nn.Sequential(DWConv(x, x, 3), Conv(x, c3, 1))), i.e. conv.py:100-105 followed by conv.py:47-49.
//
// Layer by layer the depthwise result makes a round trip through HBM / L2 (52-66 MB per 80x80 level at bs = 64) and costs
// its own launch; the depthwise kernel is instruction-bound and the 1x1 that follows is bandwidth-bound.  Here the
// depthwise conv runs INSIDE the A-operand producer of the 1x1's tcgen05 GEMM:
//
//   warp 0       TMA: the (16+2) x (8+2) halo patch of one 64-channel block, raw NHWC bf16 (zero fill = conv padding)
//   warps 10-17  depthwise 3x3 + folded BN + SiLU on CUDA cores (fp32 accumulate), one thread = 4 channels x one tile
//                column, walking down the patch rows with three rolling accumulators; results are rounded to bf16 (the
//                rounding point of the unfused path) and written as the swizzled K-major A tile of the GEMM
//   warp 1       tcgen05.mma: D[128 px, co] += A[128 px, 64 ch] x W[co, 64 ch]^T per channel block, weights resident
//   warps 2-9    conv_tc's epilogue (TMEM -> bias / SiLU -> bf16 -> swizzled staging -> TMA store)
//
// Rings: raw patches (producer -> depthwise warps), A tiles (depthwise warps -> MMA), two TMEM accumulators
// (MMA -> epilogue); persistent over (16 x 8)-pixel tiles, one CTA per SM.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "conv_tc.cuh"

namespace yl {

constexpr int kDwTW = 16, kDwTH = 8, kDwPW = kDwTW + 2, kDwPH = kDwTH + 2;
constexpr int kDwRawBytes = kDwPW * kDwPH * 128;          // [row][col][64 ch] bf16
constexpr int kDwRawStride = (kDwRawBytes + 1023) & ~1023; // keeps the regions behind the ring 1024-byte aligned
constexpr int kDwABytes = 128 * 128;                      // A tile: 128 pixels x 64 channels, SW128
constexpr int kDwWarps = 8;
constexpr int kDwThreads = kConvTcThreads + 32 * kDwWarps;
constexpr int kDwMaxStages = 4;

struct DwPwParams {
    ConvTcParams c;              // the 1x1 conv: weights map, epilogue and store geometry
    CUtensorMap tmX;             // depthwise input {C, W, H, N}: box {64, 18, 10, 1}, no swizzle
    CUtensorMap tmXt;            // the same tensor with a box of C % 64 channels: the narrow last block (dense pixels)
    const __nv_bfloat16* dw_w;   // [9][C]
    const float* dw_b;           // [C]
    int C, cblocks, dw_act, raw_stages, a_stages;
    int dbg_skip;                // timing experiments only (YL_DWPW_SKIP): 1 = depthwise warps skip their math (wrong results)
    // back-to-back mode (yl_dw_pw_det): the 1x1 result is not stored; it becomes the A operand of a SECOND 1x1 GEMM (the
    // head's last conv, weights c2.tmB) whose accumulator feeds the fused Detect epilogue (class decode / class filter)
    int b2b, k2blocks;
    ConvTcParams c2;
};

// R output rows of one tile column for 4 channels: walks R + 2 patch rows with three rolling accumulators (a patch row is
// tap row dr of output row pr - dr), finishes an output row as soon as its third patch row has been added: SiLU, bf16,
// 8-byte store into the swizzled A tile.  `rp`: shared address of patch pixel (first row, column x - 1) + channel offset,
// `row_bytes` / `pix_bytes`: patch pitches; `ap`: address of the first output row's slot (rows are 16 * 128 B apart).
template <int R>
__device__ __forceinline__ void dw_rows(const uint8_t* rp, uint32_t row_bytes, uint32_t pix_bytes, uint8_t* ap,
                                        const f32x2 (&w)[9][2], const f32x2 (&bia)[2], int act) {
    // channels (0, 1) and (2, 3) travel as fp32 pairs: one FFMA2 per tap and pair (the loop is instruction-issue bound)
    f32x2 acc[3][2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        acc[j][0] = bia[0];
        acc[j][1] = bia[1];
    }
    // (plain loads / stores: the compiler is free to software-pipeline the unrolled rows — hoist the next rows' loads and
    // FMAs over the SiLU latency of the finished row; the mbarrier waits before and the proxy fence after this function
    // carry memory clobbers)
    uint2 nxt[3];
#pragma unroll
    for (int dc = 0; dc < 3; ++dc) nxt[dc] = *reinterpret_cast<const uint2*>(rp + dc * pix_bytes);
#pragma unroll
    for (int pr = 0; pr < R + 2; ++pr) {
        f32x2 xin[3][2];
#pragma unroll
        for (int dc = 0; dc < 3; ++dc) {
            const uint2 v = nxt[dc];
            xin[dc][0] = pack_f32x2(bf16lo_f(v.x), bf16hi_f(v.x));
            xin[dc][1] = pack_f32x2(bf16lo_f(v.y), bf16hi_f(v.y));
        }
        if (pr + 1 < R + 2) {
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) nxt[dc] = *reinterpret_cast<const uint2*>(rp + (pr + 1) * row_bytes + dc * pix_bytes);
        }
#pragma unroll
        for (int dr = 0; dr < 3; ++dr) {
            const int r = pr - dr;
            if (r < 0 || r >= R) continue;
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) {
                acc[r % 3][0] = fma_f32x2(xin[dc][0], w[dr * 3 + dc][0], acc[r % 3][0]);
                acc[r % 3][1] = fma_f32x2(xin[dc][1], w[dr * 3 + dc][1], acc[r % 3][1]);
            }
        }
        const int rdone = pr - 2;
        if (rdone >= 0) {
            float o[4];
            unpack_f32x2(acc[rdone % 3][0], o[0], o[1]);
            unpack_f32x2(acc[rdone % 3][1], o[2], o[3]);
            acc[rdone % 3][0] = bia[0];
            acc[rdone % 3][1] = bia[1];
            if (act) {
#pragma unroll
                for (int c = 0; c < 4; ++c) o[c] = silu_fast(o[c]);
            }
            *reinterpret_cast<uint2*>(ap + rdone * 16 * 128) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
        }
    }
}

// Epilogue of the back-to-back mode, one group of 4 warps (thread = accumulator row) per accumulator stage:
//   stage 1  TMEM acc1 -> + bias -> SiLU -> bf16 -> this group's swizzled K-major A2 tiles (one per 64 channels; exactly the
//            layout conv_tc stages for its TMA stores) -> a2_full: the MMA warp runs the second GEMM from them;
//   stage 2  TMEM acc2 -> + bias -> Detect class decode / class filter of conv_tc.cuh (no NHWC store at all).
__device__ __forceinline__ void dwpw_epilogue_b2b(const DwPwParams& P, int g, int q, int lane, int gtid, uint32_t tmem_base,
                                                  uint64_t* tfull1, uint64_t* tempty1, uint64_t* a2_full, uint64_t* a2_empty,
                                                  uint64_t* tfull2, uint64_t* tempty2, uint8_t* a2, const float* sbias,
                                                  const float* sbias3, int tile0, int tstep) {
    const ConvTcParams& p = P.c;
    const ConvTcParams& p2 = P.c2;
    const int row = q * 32 + lane;
    const int tw = row % kDwTW, th = row / kDwTW;
    const uint32_t a2_row = smem_u32(a2) + (uint32_t)row * 128u;
    const uint32_t swz = (uint32_t)(row & 7);
    const float bscale = p.act ? 0.5f : 1.0f;
    float det_best = -INFINITY, det_dist[4] = {0.f, 0.f, 0.f, 0.f};
    int det_bestc = 0;
    uint32_t uses = 0;
    for (int tile = tile0 + g * tstep; tile < p.total_tiles; tile += 2 * tstep, ++uses) {
        int wt, ht;
        const int mt = fast_divmod(tile, p.fd_tiles_w, &wt);
        const int n = fast_divmod(mt, p.fd_tiles_h, &ht);
        const int h = ht * kDwTH + th, w = wt * kDwTW + tw;
        const long long m = (h < p.Ho && w < p.Wo) ? ((long long)n * p.Ho + h) * p.Wo + w : p2.det_M;
        const uint32_t ph = uses & 1u;
        mbar_wait(&tfull1[g], ph);
        tc_fence_after();
        mbar_wait(&a2_empty[g], ph ^ 1u);     // the second GEMM of this group's previous tile has read the A2 tiles
        const uint32_t taddr1 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.acc_stride);
        for (int c = 0; c < p.nchunks; ++c) {
            uint32_t acc[32];
            tmem_ld32(taddr1 + (uint32_t)(c * 32), acc);
            tmem_ld_wait();
            if (c == p.nchunks - 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty1[g]);
            }
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b = *reinterpret_cast<const float4*>(sbias + c * 32 + i);
                v[i + 0] = fmaf(__uint_as_float(acc[i + 0]), bscale, b.x);
                v[i + 1] = fmaf(__uint_as_float(acc[i + 1]), bscale, b.y);
                v[i + 2] = fmaf(__uint_as_float(acc[i + 2]), bscale, b.z);
                v[i + 3] = fmaf(__uint_as_float(acc[i + 3]), bscale, b.w);
            }
            if (p.act) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], tanh_approx(v[i]), v[i]);
            }
            // 32 channels = 64 bytes of this row of A2 tile (c >> 1): 16-byte pieces (c & 1) * 4 + j, XOR-swizzled by the row
            const uint32_t dst = a2_row + (uint32_t)(c >> 1) * (uint32_t)kDwABytes;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t piece = ((uint32_t)((c & 1) * 4 + j)) ^ swz;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (piece << 4)),
                             "r"(pack_bf16x2(v[8 * j], v[8 * j + 1])), "r"(pack_bf16x2(v[8 * j + 2], v[8 * j + 3])),
                             "r"(pack_bf16x2(v[8 * j + 4], v[8 * j + 5])), "r"(pack_bf16x2(v[8 * j + 6], v[8 * j + 7]))
                             : "memory");
            }
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + g, kEpiGroupThreads);
        if (gtid == 0) mbar_arrive(&a2_full[g]);

        mbar_wait(&tfull2[g], ph);
        tc_fence_after();
        const uint32_t taddr2 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(2 * p.acc_stride + g * p2.acc_stride);
        for (int c = 0; c < p2.nchunks; ++c) {
            uint32_t acc[32];
            tmem_ld32(taddr2 + (uint32_t)(c * 32), acc);
            tmem_ld_wait();
            if (c == p2.nchunks - 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty2[g]);
            }
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b = *reinterpret_cast<const float4*>(sbias3 + c * 32 + i);
                v[i + 0] = __uint_as_float(acc[i + 0]) + b.x;
                v[i + 1] = __uint_as_float(acc[i + 1]) + b.y;
                v[i + 2] = __uint_as_float(acc[i + 2]) + b.z;
                v[i + 3] = __uint_as_float(acc[i + 3]) + b.w;
            }
            if (p2.det_mode == YL_DET_CLS_FILTER) det_filter_chunk<32>(p2, v, c, m, lane, det_best, det_bestc);
            else det_decode_chunk<32>(p2, v, c, m, det_dist, p2.det_mode);
        }
    }
}

__global__ void __launch_bounds__(kDwThreads, 1) dwpw_tc_kernel(const __grid_constant__ DwPwParams P) {
    const ConvTcParams& p = P.c;
    extern __shared__ uint8_t smem_raw[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* sRaw = base;
    uint8_t* sA = sRaw + (size_t)P.raw_stages * kDwRawStride;
    uint8_t* sB = sA + (size_t)P.a_stages * kDwABytes;
    uint8_t* sB3 = sB + (size_t)P.cblocks * p.b_bytes;                 // back-to-back: the second GEMM's weight tiles
    uint8_t* sStg = sB3 + (size_t)(P.b2b ? P.k2blocks : 0) * P.c2.b_bytes;  // staging tiles / back-to-back: A2 tiles
    float* sbias = reinterpret_cast<float*>(sStg + 2 * (size_t)p.stg_bufs * p.stg_bytes);
    const int nbias = p.co_tile + 32;
    const int cpad = P.cblocks * 64;
    float* sdw = sbias + ((nbias + 3) & ~3);              // [9][cpad] depthwise weights (fp32), then [cpad] bias
    float* sbias3 = sdw + 10 * cpad;                      // back-to-back: bias of the second GEMM
    const int nbias3 = P.b2b ? P.c2.co_tile + 32 : 0;
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(sbias3 + ((nbias3 + 3) & ~3));
    uint64_t* raw_empty = raw_full + kDwMaxStages;
    uint64_t* a_full = raw_empty + kDwMaxStages;
    uint64_t* a_empty = a_full + kDwMaxStages;
    uint64_t* tfull_bar = a_empty + kDwMaxStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* w_bar = tempty_bar + 2;
    uint64_t* a2_full = w_bar + 1;          // [2] back-to-back: A2 tiles of group g written
    uint64_t* a2_empty = a2_full + 2;       // [2] ... and read by the second GEMM
    uint64_t* tfull2_bar = a2_empty + 2;    // [2] second accumulator ready / drained
    uint64_t* tempty2_bar = tfull2_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty2_bar + 2);

    griddep_launch_dependents();
    if (threadIdx.x == 0) {
        for (int s = 0; s < kDwMaxStages; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_empty[s], kDwWarps);
            mbar_init(&a_full[s], kDwWarps);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 4);
            mbar_init(&a2_full[s], 1);
            mbar_init(&a2_empty[s], 1);
            mbar_init(&tfull2_bar[s], 1);
            mbar_init(&tempty2_bar[s], 4);
        }
        mbar_init(w_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, p.tmem_cols);
        tmem_relinquish();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&P.tmX);
        tma_prefetch_desc(&p.tmB);
        tma_prefetch_desc(&p.tmY[0]);
    }
    {
        const float bs = p.act ? 0.5f : 1.0f;   // the epilogue works on h = x / 2 (see conv_tc)
        for (int i = threadIdx.x; i < nbias; i += blockDim.x) sbias[i] = i < p.n_bias ? bs * __ldg(p.bias + i) : 0.f;
        for (int i = threadIdx.x; i < 9 * cpad; i += blockDim.x) {
            const int tap = i / cpad, ch = i - tap * cpad;
            sdw[i] = ch < P.C ? __bfloat162float(P.dw_w[tap * P.C + ch]) : 0.f;
        }
        for (int i = threadIdx.x; i < cpad; i += blockDim.x) sdw[9 * cpad + i] = i < P.C ? __ldg(P.dw_b + i) : 0.f;
        for (int i = threadIdx.x; i < nbias3; i += blockDim.x) sbias3[i] = i < P.c2.n_bias ? __ldg(P.c2.bias + i) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // the 1x1 weights are constants: their loads fly while the previous kernel of the stream drains
    if (warp == 0 && elect_one()) {
        mbar_expect_tx(w_bar, (uint32_t)P.cblocks * (uint32_t)p.co_tile * 128u +
                                  (P.b2b ? (uint32_t)P.k2blocks * (uint32_t)P.c2.co_tile * 128u : 0u));
        for (int cb = 0; cb < P.cblocks; ++cb) tma_load_2d(sB + (size_t)cb * p.b_bytes, &p.tmB, w_bar, cb * 64, 0);
        if (P.b2b)
            for (int kb = 0; kb < P.k2blocks; ++kb) tma_load_2d(sB3 + (size_t)kb * P.c2.b_bytes, &P.c2.tmB, w_bar, kb * 64, 0);
    }
    griddep_wait();

    const int tile0 = (int)blockIdx.x, tstep = (int)gridDim.x;
    if (warp == 0) {
        // ================= TMA producer: raw halo patches =================
        const bool leader = elect_one();
        int st = 0;
        uint32_t ph = 0;
        for (int tile = tile0; tile < p.total_tiles; tile += tstep) {
            int wt, ht;
            const int mt = fast_divmod(tile, p.fd_tiles_w, &wt);
            const int n = fast_divmod(mt, p.fd_tiles_h, &ht);
            for (int cb = 0; cb < P.cblocks; ++cb) {
                mbar_wait(&raw_empty[st], ph ^ 1u);
                if (leader) {
                    const int cbw = min(64, P.C - cb * 64);
                    mbar_expect_tx(&raw_full[st], (uint32_t)(kDwPW * kDwPH * cbw * 2));
                    tma_load_4d(sRaw + (size_t)st * kDwRawStride, cbw == 64 ? &P.tmX : &P.tmXt, &raw_full[st], cb * 64,
                                wt * kDwTW - 1, ht * kDwTH - 1, n);
                }
                if (++st == P.raw_stages) {
                    st = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const bool leader = elect_one();
        const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)p.co_tile);
        int st = 0, acc = 0;
        uint32_t ph = 0, acc_ph = 0;
        mbar_wait(w_bar, 0);
        tc_fence_after();
        int ntile = 0;
        const uint32_t idesc2 = umma_idesc_bf16(128, (uint32_t)P.c2.co_tile);
        // D2[g] = A2[g] (this CTA's j-th tile, written by epilogue group g = j & 1) x W3^T
        auto mma2 = [&](int j) {
            const int g2 = j & 1;
            const uint32_t ph2 = (uint32_t)(j >> 1) & 1u;
            mbar_wait(&a2_full[g2], ph2);
            mbar_wait(&tempty2_bar[g2], ph2 ^ 1u);
            tc_fence_after();
            if (leader) {
                const uint32_t d2 = tmem_base + (uint32_t)(2 * p.acc_stride + g2 * P.c2.acc_stride);
                const int c2ch = p.co_tile < (int)P.c2.ci_pad ? p.co_tile : (int)P.c2.ci_pad;   // channels of the 1x1 result
                for (int kb = 0; kb < P.k2blocks; ++kb) {
                    const int ks2 = (min(64, c2ch - kb * 64) + 15) >> 4;
                    const uint64_t da = umma_desc_kmajor(smem_u32(sStg + (size_t)(g2 * P.k2blocks + kb) * kDwABytes), 128u);
                    const uint64_t db = umma_desc_kmajor(smem_u32(sB3 + (size_t)kb * P.c2.b_bytes), 128u);
                    for (int k = 0; k < ks2; ++k)
                        umma_bf16(d2, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&a2_empty[g2]);
                umma_commit(&tfull2_bar[g2]);
            }
        };
        for (int tile = tile0; tile < p.total_tiles; tile += tstep) {
            mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
            for (int cb = 0; cb < P.cblocks; ++cb) {
                const int cbw = min(64, P.C - cb * 64);
                const int ksteps = cbw >> 4;
                mbar_wait(&a_full[st], ph);
                tc_fence_after();
                if (leader) {
                    const uint64_t da = umma_desc_kmajor(smem_u32(sA + (size_t)st * kDwABytes), 128u);
                    const uint64_t db = umma_desc_kmajor(smem_u32(sB + (size_t)cb * p.b_bytes), 128u);
                    for (int k = 0; k < ksteps; ++k)
                        umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (cb > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&a_empty[st]);
                }
                if (++st == P.a_stages) {
                    st = 0;
                    ph ^= 1u;
                }
            }
            if (leader) umma_commit(&tfull_bar[acc]);
            acc ^= 1;
            if (acc == 0) acc_ph ^= 1u;
            // back-to-back: the second GEMM of the PREVIOUS tile (its A2 tiles were written by the epilogue meanwhile)
            if (P.b2b) {
                if (ntile > 0) mma2(ntile - 1);
                ++ntile;
            }
        }
        if (P.b2b && ntile > 0) mma2(ntile - 1);
    } else if (warp < 2 + 8) {
        // ================= epilogue (conv_tc.cuh) =================
        const int e = warp - 2;
        const int g = e >> 2;
        const int q = warp & 3;
        const int gtid = (e & 3) * 32 + lane;
        uint8_t* stg = sStg + (size_t)g * p.stg_bufs * p.stg_bytes;
        const TileRange tr = {tile0, p.total_tiles, tstep};
        uint32_t acc_uses = 0;
        if (P.b2b)
            dwpw_epilogue_b2b(P, g, q, lane, gtid, tmem_base, tfull_bar, tempty_bar, a2_full, a2_empty, tfull2_bar, tempty2_bar,
                              sStg + (size_t)g * P.k2blocks * kDwABytes, sbias, sbias3, tile0, tstep);
        else if (p.epi_kind == 0)
            conv_tc_epilogue<32, true, false, 0, false, false>(p, p, tr, acc_uses, g, q, lane, gtid, tmem_base, tfull_bar,
                                                               tempty_bar, stg, sbias);
        else
            conv_tc_epilogue<32, false, false, 0, false, false>(p, p, tr, acc_uses, g, q, lane, gtid, tmem_base, tfull_bar,
                                                                tempty_bar, stg, sbias);
    } else {
        // ================= depthwise 3x3 -> A tiles =================
        const int dtid = (int)threadIdx.x - kConvTcThreads;
        int rs = 0, as = 0;
        uint32_t rph = 0, aph = 0;
        for (int tile = tile0; tile < p.total_tiles; tile += tstep) {
            for (int cb = 0; cb < P.cblocks; ++cb) {
                // A full block is 16 groups of 4 channels x 16 tile columns = all 256 threads, each walking the 8 rows of
                // its column.  A narrower last block (C % 64 channels, dense 2 * cbw-byte pixels) splits every column into
                // 2 or 4 row segments instead, so all threads stay busy with (4 or 2) + 2 patch rows each.
                const int cbw = min(64, P.C - cb * 64);
                const int cgn = cbw >> 2;
                const int segs = cgn <= 4 ? 4 : (cgn <= 8 ? 2 : 1);
                int cg, x, seg;
                if (cgn == 16) {
                    cg = dtid & 15;
                    x = dtid >> 4;
                    seg = 0;
                } else {
                    cg = dtid % cgn;
                    const int t2 = dtid / cgn;
                    x = t2 & 15;
                    seg = t2 >> 4;
                }
                const bool active = seg < segs && !(P.dbg_skip & 1);
                f32x2 w[9][2], bia[2];
                if (active) {
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(sdw + t * cpad + cb * 64 + cg * 4);
                        w[t][0] = v.x;
                        w[t][1] = v.y;
                    }
                    const ulonglong2 bv = *reinterpret_cast<const ulonglong2*>(sdw + 9 * cpad + cb * 64 + cg * 4);
                    bia[0] = bv.x;
                    bia[1] = bv.y;
                }
                mbar_wait(&raw_full[rs], rph);
                mbar_wait(&a_empty[as], aph ^ 1u);
                if (active) {
                    const uint32_t pix_bytes = (uint32_t)cbw * 2u, row_bytes = (uint32_t)kDwPW * pix_bytes;
                    const int r0 = seg * (kDwTH / segs);                 // first output row of this thread
                    // destination inside a (128 rows x 128 B) SW128 tile: row m = r * 16 + x, 16-byte chunk (cg >> 1) ^ (m & 7)
                    const uint32_t a_col = ((uint32_t)(((cg >> 1) ^ (x & 7)) << 4)) | ((uint32_t)(cg & 1) << 3);
                    const uint8_t* rp = sRaw + (size_t)rs * kDwRawStride + (uint32_t)r0 * row_bytes + (uint32_t)x * pix_bytes +
                                        (uint32_t)cg * 8u;
                    uint8_t* ap = sA + (size_t)as * kDwABytes + (uint32_t)(r0 * 16 + x) * 128u + a_col;
                    if (segs == 1) dw_rows<8>(rp, row_bytes, pix_bytes, ap, w, bia, P.dw_act);
                    else if (segs == 2) dw_rows<4>(rp, row_bytes, pix_bytes, ap, w, bia, P.dw_act);
                    else dw_rows<2>(rp, row_bytes, pix_bytes, ap, w, bia, P.dw_act);
                }
                fence_proxy_async_smem();      // the A tile is read by the tensor core through the async proxy
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&a_full[as]);
                    mbar_arrive(&raw_empty[rs]);
                }
                if (++rs == P.raw_stages) {
                    rs = 0;
                    rph ^= 1u;
                }
                if (++as == P.a_stages) {
                    as = 0;
                    aph ^= 1u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

static int g_dwpw_max_smem = 0, g_dwpw_sms = 148;

int init_dwpw() {
    int dev = 0;
    YL_CUDA(cudaGetDevice(&dev));
    YL_CUDA(cudaDeviceGetAttribute(&g_dwpw_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    YL_CUDA(cudaDeviceGetAttribute(&g_dwpw_sms, cudaDevAttrMultiProcessorCount, dev));
    YL_CUDA(cudaFuncSetAttribute(dwpw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_dwpw_max_smem));
    return YL_OK;
}

// `head`: NULL, or the head's last 1x1 conv (back-to-back mode: the pointwise result is consumed on chip, `a->y` only gives
// its channel count)
static bool dwpw_ok(const yl_conv_args* a, const yl_conv_args* head, char* why, size_t n) {
#define NOPE(msg)                          \
    do {                                   \
        if (why) snprintf(why, n, "%s", msg); \
        return false;                      \
    } while (0)
    const yl_tensor& x = a->x;
    const yl_tensor& y = a->y;
    if (a->k != 1 || a->stride != 1) NOPE("the pointwise conv must be 1x1 stride 1");
    if (a->res.data || a->y_up.data || a->upsample2x || a->det.pred) NOPE("no residual / upsample / Detect epilogue");
    if (x.dtype != YL_BF16 || y.dtype != YL_BF16 || !x.data || (!y.data && !head)) NOPE("bf16 in, bf16 out");
    if (x.c % 16 || x.c > 256 || x.coff % 8 || x.cstride % 8) NOPE("depthwise channels: multiple of 16, <= 256, 8-aligned slice");
    if (y.c % 8 || y.c > 128 || y.coff % 8 || y.cstride % 8) NOPE("pointwise output channels: multiple of 8, <= 128");
    if (a->ci_pad < x.c || a->ci_pad % 8 || a->co_pad < y.c || a->co_pad % 8) NOPE("bad packed weight dims");
    if (y.n != x.n || y.h != x.h || y.w != x.w) NOPE("shape mismatch");
    if (((uintptr_t)x.data | (uintptr_t)y.data | (uintptr_t)a->w | (uintptr_t)a->bias) & 15) NOPE("16-byte alignment");
    if (head) {
        const yl_det_epilogue& d = head->det;
        if (head->k != 1 || head->stride != 1 || head->act || head->res.data || head->y_up.data || head->upsample2x)
            NOPE("the head conv must be a plain 1x1 without activation / residual / upsample");
        if (head->y.data) NOPE("back-to-back mode does not store the head conv (Detect epilogue only)");
        if (!d.pred || !(d.mode == YL_DET_CLS || d.mode == YL_DET_CLS_FILTER)) NOPE("the head conv needs a class decode / filter epilogue");
        if (head->x.c != y.c || y.c > 128) NOPE("head conv input channels must equal the pointwise output channels (<= 128)");
        if (head->y.c != d.nc || d.nc > 128 || d.nc % 8) NOPE("head conv: co == nc, multiple of 8, <= 128");
        if (head->ci_pad < head->x.c || head->ci_pad % 8 || head->co_pad < d.nc || head->co_pad % 8) NOPE("bad head weight dims");
        if (d.mode == YL_DET_CLS_FILTER && (!d.cand_ws || !(d.conf >= 0.f && d.conf <= 1.f))) NOPE("class filter needs a workspace");
        if ((long long)d.A * d.nc >= (1ll << 32) || d.A < 1 || d.anchor0 < 0 || d.anchor0 + x.h * x.w > d.A) NOPE("bad anchor range");
        if (((uintptr_t)head->w | (uintptr_t)head->bias) & 15) NOPE("16-byte alignment");
    }
    return true;
#undef NOPE
}

}  // namespace yl

extern "C" {

// co3 = 0: plain mode; else the output channels of the head conv of the back-to-back mode
static size_t dwpw_smem(int cblocks, int co_tile, int raw_stages, int a_stages, int co3) {
    const int nchunks = yl::ceil_div(co_tile, 32);
    const size_t b_bytes = ((size_t)co_tile * 128 + 1023) & ~(size_t)1023;
    const size_t stg_bytes = (128 * (size_t)((nchunks >= 2 ? 2 : 1) * 64) + 1023) & ~(size_t)1023;
    size_t stg = 4 * stg_bytes, extra = 0;
    if (co3 > 0) {
        const int k2 = yl::ceil_div(co_tile, 64), co_tile3 = yl::ceil_div(co3, 16) * 16;
        const size_t a2 = 2 * (size_t)k2 * yl::kDwABytes;
        if (a2 > stg) stg = a2;
        extra = (size_t)k2 * (((size_t)co_tile3 * 128 + 1023) & ~(size_t)1023) + (size_t)((co_tile3 + 32 + 3) & ~3) * 4;
    }
    return 1024 + (size_t)raw_stages * yl::kDwRawStride + (size_t)a_stages * yl::kDwABytes + (size_t)cblocks * b_bytes + stg +
           extra + (size_t)((co_tile + 32 + 3) & ~3) * 4 + (size_t)10 * cblocks * 64 * 4 + (4 * yl::kDwMaxStages + 13) * 8 + 16;
}
// ring depths that fit: three raw patches in flight when shared memory allows (a patch is consumed in ~1.5 us, an HBM-miss
// TMA load takes longer), else two; 0 = the shape does not fit at all
static int dwpw_raw_stages(const yl_conv_args* a, const yl_conv_args* head) {
    const int cblocks = yl::ceil_div(a->x.c, 64), co_tile = yl::ceil_div(a->y.c, 16) * 16;
    static const int k_raw = [] { const char* e = getenv("YL_DWPW_RAW"); return e && *e ? atoi(e) : 3; }();
    for (int r = k_raw < 2 ? 2 : (k_raw > yl::kDwMaxStages ? yl::kDwMaxStages : k_raw); r >= 2; --r)
        if ((int)dwpw_smem(cblocks, co_tile, r, 2, head ? head->det.nc : 0) <= yl::g_dwpw_max_smem) return r;
    return 0;
}

static int dwpw_launch(const yl_conv_args* a, const void* dw_w, const float* dw_bias, int dw_act, const yl_conv_args* head,
                       void* stream);

int yl_dw_pw_supported(const yl_conv_args* pw) {
    return pw && yl::dwpw_ok(pw, nullptr, nullptr, 0) && dwpw_raw_stages(pw, nullptr) > 0 ? 1 : 0;
}

int yl_dw_pw_det_supported(const yl_conv_args* pw, const yl_conv_args* head) {
    return pw && head && yl::dwpw_ok(pw, head, nullptr, 0) && dwpw_raw_stages(pw, head) > 0 ? 1 : 0;
}

int yl_dw_pw_conv(const yl_conv_args* a, const void* dw_w, const float* dw_bias, int dw_act, void* stream) {
    return dwpw_launch(a, dw_w, dw_bias, dw_act, nullptr, stream);
}

int yl_dw_pw_det(const yl_conv_args* a, const void* dw_w, const float* dw_bias, int dw_act, const yl_conv_args* head,
                 void* stream) {
    YL_CHECK(head, YL_ERR_ARG, "null pointer");
    return dwpw_launch(a, dw_w, dw_bias, dw_act, head, stream);
}

static int dwpw_launch(const yl_conv_args* a, const void* dw_w, const float* dw_bias, int dw_act, const yl_conv_args* head,
                       void* stream) {
    using namespace yl;
    YL_CHECK(a && dw_w && dw_bias, YL_ERR_ARG, "null pointer");
    char why[128];
    YL_CHECK(dwpw_ok(a, head, why, sizeof(why)), YL_ERR_UNSUPPORTED, "fused depthwise + pointwise conv unsupported: %s", why);
    const int raw_stages = dwpw_raw_stages(a, head);
    YL_CHECK(raw_stages > 0, YL_ERR_UNSUPPORTED, "fused depthwise + pointwise conv: %d -> %d channels do not fit shared memory",
             a->x.c, a->y.c);
    const yl_tensor& x = a->x;
    const yl_tensor& y = a->y;
    DwPwParams P;
    DwPwParams* pp = &P;
    memset(pp, 0, sizeof(*pp));
    ConvTcParams& p = pp->c;
    const int H = x.h, W = x.w, N = x.n;
    p.Ho = H;
    p.Wo = W;
    p.Nimg = N;
    p.TW = kDwTW;
    p.TH = kDwTH;
    p.TN = 1;
    p.tiles_w = ceil_div(W, kDwTW);
    p.tiles_h = ceil_div(H, kDwTH);
    p.tiles_n = N;
    p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    p.n_tiles = 1;
    p.total_tiles = p.m_tiles;
    p.fd_ntiles = make_fastdiv(1);
    p.fd_tiles_w = make_fastdiv(p.tiles_w);
    p.fd_tiles_h = make_fastdiv(p.tiles_h);
    p.ksize = 1;
    p.stride = 1;
    p.ci_pad = a->ci_pad;
    p.kblk = 64;
    p.cin_blocks = ceil_div(x.c, 64);
    p.co_tile = ceil_div(y.c, 16) * 16;
    p.cw = 32;
    p.nchunks = ceil_div(p.co_tile, 32);
    p.acc_stride = p.nchunks * 32;
    uint32_t cols = 32;
    while ((int)cols < 2 * p.acc_stride) cols <<= 1;
    p.tmem_cols = cols;
    p.b_bytes = ((uint32_t)p.co_tile * 128u + 1023u) & ~1023u;
    p.stg_sub = p.nchunks >= 2 ? 2 : 1;
    p.stg_bufs = 2;
    p.stg_row_bytes = p.stg_sub * 32 * 2;
    p.stg_bytes = (128u * (uint32_t)p.stg_row_bytes + 1023u) & ~1023u;
    p.nstore = ceil_div(p.nchunks, p.stg_sub);
    p.bias = a->bias;
    p.n_bias = a->co_pad;
    p.act = a->act;
    p.epi_kind = a->act ? 0 : 2;
    p.store_y = 1;
    pp->dw_w = reinterpret_cast<const __nv_bfloat16*>(dw_w);
    pp->dw_b = dw_bias;
    pp->C = x.c;
    pp->cblocks = p.cin_blocks;
    pp->dw_act = dw_act;
    static const int k_skip = [] { const char* e = getenv("YL_DWPW_SKIP"); return e && *e ? atoi(e) : 0; }();
    pp->dbg_skip = k_skip;
    pp->raw_stages = raw_stages;
    pp->a_stages = 2;

    bool ok = true;
    {   // depthwise input: {C, W, H, N}, one 64-channel block of the halo patch per box, plain (unswizzled) layout
        uint64_t dims[4] = {(uint64_t)x.c, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)x.cstride * 2, (uint64_t)x.cstride * 2 * W, (uint64_t)x.cstride * 2 * W * H};
        uint32_t box[4] = {64u, (uint32_t)kDwPW, (uint32_t)kDwPH, 1u};
        ok = ok && encode_map(&pp->tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                              reinterpret_cast<__nv_bfloat16*>(x.data) + x.coff, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    }
    if (x.c % 64) {   // the narrow last block: dense pixels of C % 64 channels
        uint64_t dims[4] = {(uint64_t)x.c, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)x.cstride * 2, (uint64_t)x.cstride * 2 * W, (uint64_t)x.cstride * 2 * W * H};
        uint32_t box[4] = {(uint32_t)(x.c % 64), (uint32_t)kDwPW, (uint32_t)kDwPH, 1u};
        ok = ok && encode_map(&pp->tmXt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                              reinterpret_cast<__nv_bfloat16*>(x.data) + x.coff, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    }
    {   // 1x1 weights [co_pad][ci_pad]
        uint64_t dims[2] = {(uint64_t)a->ci_pad, (uint64_t)a->co_pad};
        uint64_t str[1] = {(uint64_t)a->ci_pad * 2};
        uint32_t box[2] = {64u, (uint32_t)p.co_tile};
        ok = ok && encode_map(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, const_cast<void*>(a->w), 2, dims, str, box,
                              CU_TENSOR_MAP_SWIZZLE_128B);
    }
    if (!head) {
        uint32_t obox[4] = {(uint32_t)(p.stg_sub * 32), (uint32_t)kDwTW, (uint32_t)kDwTH, 1u};
        ok = ok && encode_out_maps(&p.tmY[0], y, H, W, N, false, false, obox, swizzle_for_bytes(p.stg_row_bytes));
        p.y_map_first = 0;
        p.y_map_last = 1;
    } else {
        // back-to-back: second GEMM D2[128 px, nc] = A2[128 px, y.c] x W3[nc, y.c]^T, Detect epilogue on D2
        ConvTcParams& q = pp->c2;
        const yl_det_epilogue& d = head->det;
        pp->b2b = 1;
        pp->k2blocks = ceil_div(p.co_tile, 64);
        p.tmY[0] = pp->tmX;    // (prefetched by the kernel; never stored through)
        q.ci_pad = head->ci_pad;
        q.co_tile = ceil_div(d.nc, 16) * 16;
        q.cw = 32;
        q.nchunks = ceil_div(q.co_tile, 32);
        q.acc_stride = q.nchunks * 32;
        q.b_bytes = ((uint32_t)q.co_tile * 128u + 1023u) & ~1023u;
        q.bias = head->bias;
        q.n_bias = head->co_pad;
        q.det_mode = d.mode;
        q.det_pred = d.pred;
        q.det_nc = d.nc;
        q.det_A = d.A;
        q.det_anchor0 = d.anchor0;
        q.det_hw = H * W;
        q.det_w = W;
        q.det_stride = d.stride;
        q.det_M = (long long)N * H * W;
        if (d.mode == YL_DET_CLS_FILTER) {
            q.det_conf = d.conf;
            q.det_cand_counts = reinterpret_cast<uint32_t*>(d.cand_ws);
            q.det_cand_keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(d.cand_ws) +
                                                                    ((size_t)N * 4 + 255) / 256 * 256);
        }
        uint32_t cols2 = 32;
        while ((int)cols2 < 2 * p.acc_stride + 2 * q.acc_stride) cols2 <<= 1;
        YL_CHECK(cols2 <= 512, YL_ERR_UNSUPPORTED, "back-to-back accumulators need %u TMEM columns", cols2);
        p.tmem_cols = cols2;
        uint64_t dims[2] = {(uint64_t)head->ci_pad, (uint64_t)head->co_pad};
        uint64_t str[1] = {(uint64_t)head->ci_pad * 2};
        uint32_t box[2] = {64u, (uint32_t)q.co_tile};
        ok = ok && encode_map(&q.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, const_cast<void*>(head->w), 2, dims, str, box,
                              CU_TENSOR_MAP_SWIZZLE_128B);
    }
    if (!ok) return YL_ERR_CUDA;
    const size_t smem = dwpw_smem(pp->cblocks, p.co_tile, pp->raw_stages, pp->a_stages, head ? head->det.nc : 0);
    YL_CHECK((int)smem <= g_dwpw_max_smem, YL_ERR_UNSUPPORTED, "fused depthwise + pointwise conv needs %zu B of shared memory",
             smem);
    int grid = g_dwpw_sms < p.total_tiles ? g_dwpw_sms : p.total_tiles;
    YL_CUDA(launch_kernel(dwpw_tc_kernel, dim3(grid), dim3(kDwThreads), smem, (cudaStream_t)stream, P));
    YL_LAUNCH_OK("dwpw_tc_kernel");
    return YL_OK;
}

}  // extern "C"
