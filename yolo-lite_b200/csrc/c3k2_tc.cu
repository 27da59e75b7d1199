// Fused tail of a C3k2 block on the 5th-gen tensor cores: three chained GEMMs per spatial tile whose intermediate
// activations never leave the SM — each epilogue writes its bf16 tile into a swizzled K-major shared-memory tile that
// IS the next tcgen05.mma's A operand.
//
//   reference (nn/modules/block.py:231-235, 330-343, 720-728), t = cv1(x) = [y0 | y1] (2c channels):
//       h   = SiLU(BN(conv3x3(y1)))            c   -> c/2       Bottleneck.cv1        stage A
//       y2  = y1 + SiLU(BN(conv3x3(h)))        c/2 -> c         Bottleneck.cv2 (+ add) stage B
//       out = SiLU(BN(conv1x1([y0, y1, y2])))  3c  -> c2        C2f.cv2               stage C
//
// The mma.sync version of this fusion (c3k2_fused.cu) removed the HBM round trips of h and y2 but is bound by the legacy
// tensor path + L1 (28-33 % of its HBM floor).  Here:
//   * a CTA owns a 14 x 16 pixel tile; the (14+4) x (16+4 -> 26) halo patch of t arrives by ONE 4-D TMA box (conv zero
//     padding = TMA out-of-bounds fill) into a 64 B / 128 B swizzled tile T, double buffered across tiles;
//   * a GEMM's M = 128 rows are 16 patch rows x 8 columns ("strip"); a 3x3 tap is the same strip of the source tile
//     shifted by (dr, dc): only the descriptor start address moves (stride between 8-row groups = the tile's row
//     pitch; the swizzle XOR works on absolute address bits, as measured for conv_tc's halo-patch mode);
//   * stage A: 3 strips (16 x 24 positions, the 16 x 18 that stage B needs rounded up to whole strips), N = 16;
//     epilogue: +bias, SiLU, ZERO outside the image (stage B's zero padding), bf16 -> tile Hs (32 B rows, SW32);
//   * stage B: 2 strips over Hs, N = c; epilogue: +bias, SiLU, + y1 (read back from T), bf16 -> tile Y2;
//   * stage C: 2 strips, K = [y0 | y1] from T + y2 from Y2, N = c2; epilogue: +bias, SiLU, bf16, 16-byte global stores of
//     the pixels inside the tile and the image.
//   Rounding points are those of the layer-by-layer path (h, y2 rounded to bf16 once), so results agree with it up to
//   fp32 accumulation order.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = two epilogue groups (a strip's 128
// rows = 4 warps).  All weights (BN folded, < 50 KB) are re-tiled once per CTA into per-(stage, tap, source) UMMA B
// tiles before the PDL dependency wait.  TMEM: 256 columns (two CTAs per SM when shared memory allows).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace yl {

constexpr int kCtTH = 14, kCtTW = 16;          // output tile
constexpr int kCtPR = kCtTH + 4;               // 18 patch rows
constexpr int kCtPP = 26;                      // patch row pitch (16 + 4 halo columns, +6 so stage A can run whole strips)
constexpr int kCtHR = 18, kCtHP = 24;          // Hs: rows allocated (16 written + 2 read by garbage rows), pitch
constexpr int kCtYR = 16, kCtYP = 16;          // Y2
constexpr int kCtThreads = 320;

struct C3k2TcParams {
    CUtensorMap tmT;
    __nv_bfloat16* y;
    long long y_cstride;
    int y_coff;
    const __nv_bfloat16 *wa, *wb, *w2;   // packed [co_pad][taps * ci_pad]
    const float *ba, *bb, *b2;
    int wa_k, wb_k, w2_k, wb_ci;         // packed row lengths (elements), ci_pad of wb
    int N, H, W, C, C2, add;
    int tiles_w, tiles_h, total_tiles;
    FastDiv fd_tw, fd_th;
    // shared-memory carve-up (byte offsets from the 1024-aligned base)
    uint32_t off_T[2], off_H, off_Y2, off_Ba, off_Bb, off_Bc, off_bias, off_bar;
    uint32_t t_bytes;                    // bytes one patch box delivers
    int rbT, rbH, rbY;                   // row bytes of T / Hs / Y2 (= their swizzle spans)
    int nb;                              // stage B / Y2 channel count (= C)
    uint32_t colA, colB, colC;           // TMEM column bases
    int alias;                           // stage C's accumulator overlaps A / B's: drain before the next tile starts
};

__device__ __forceinline__ uint32_t swz(uint32_t off, uint32_t mask) { return off ^ (((off >> 7) & mask) << 4); }

// One UMMA B tile [rows][kc] (kc * 2 = rb bytes per row, rb in {32, 64}) from packed weights w[row * wk + koff + k],
// rows >= valid_rows and k >= valid_k zero; written with the swizzle the descriptor of row pitch rb expects.
__device__ __forceinline__ void stage_b_tile(uint8_t* base, uint32_t tile_off, int rows, int rb, const __nv_bfloat16* w,
                                             int wk, int koff, int valid_rows, int valid_k) {
    const int chunks = rb / 16;
    const uint32_t mask = (uint32_t)(chunks - 1);
    for (int i = threadIdx.x; i < rows * chunks; i += blockDim.x) {
        const int r = i / chunks, j = i - r * chunks;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < valid_rows && j * 8 < valid_k) v = __ldg(reinterpret_cast<const uint4*>(w + (size_t)r * wk + koff + j * 8));
        const uint32_t off = tile_off + (uint32_t)(r * rb + j * 16);
        *reinterpret_cast<uint4*>(base + swz(off, mask)) = v;
    }
}

__global__ void __launch_bounds__(kCtThreads, 2) c3k2_tc_kernel(const __grid_constant__ C3k2TcParams p) {
    extern __shared__ uint8_t ct_smem_raw[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(ct_smem_raw);
    uint8_t* base = ct_smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const uint32_t sbase = smem_u32(base);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + p.off_bar);
    uint64_t* t_full = bars;          // [2]
    uint64_t* t_empty = bars + 2;     // [2]
    uint64_t* acc_bar = bars + 4;     // [3] accumulator of stage A / B / C complete
    uint64_t* h_ready = bars + 7;     // Hs written (8 epilogue warps)
    uint64_t* y_ready = bars + 8;     // Y2 written
    uint64_t* c_drained = bars + 9;   // stage C accumulator read out
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
    float* sbias = reinterpret_cast<float*>(base + p.off_bias);   // [16] a | [C] b | [C2] c, all pre-halved

    const int C = p.C, C2 = p.C2, ch = C / 2;
    griddep_launch_dependents();
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], 1);
        }
        for (int i = 0; i < 3; ++i) mbar_init(&acc_bar[i], 1);
        mbar_init(h_ready, 8);
        mbar_init(y_ready, 8);
        mbar_init(c_drained, 8);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    if (warp == 0 && lane == 0) tma_prefetch_desc(&p.tmT);
    // ---- constants: weights re-tiled into UMMA B tiles, biases (x 0.5: the epilogues work on h = x / 2)
    {
        // stage A: 9 taps x [16 rows (ch valid)][C]
        const int rbA = C * 2;
        for (int tap = 0; tap < 9; ++tap)
            stage_b_tile(base, p.off_Ba + (uint32_t)(tap * 16 * rbA), 16, rbA, p.wa, p.wa_k, tap * C, ch, C);
        // stage B: 9 taps x [C rows][max(ch, 16)]  (h is stored with 16 channels; the pad meets zero weights)
        const int kB = ch < 16 ? 16 : ch, rbB = kB * 2;
        for (int tap = 0; tap < 9; ++tap)
            stage_b_tile(base, p.off_Bb + (uint32_t)(tap * C * rbB), C, rbB, p.wb, p.wb_k, tap * p.wb_ci, C, ch);
        // stage C: sources y0, y1, y2: [C2 rows][C]
        for (int src = 0; src < 3; ++src)
            stage_b_tile(base, p.off_Bc + (uint32_t)(src * C2 * rbA), C2, rbA, p.w2, p.w2_k, src * C, C2, C);
        for (int i = threadIdx.x; i < 16 + C + C2; i += blockDim.x) {
            float v;
            if (i < 16) v = i < ch ? __ldg(p.ba + i) : 0.f;
            else if (i < 16 + C) v = __ldg(p.bb + (i - 16));
            else v = __ldg(p.b2 + (i - 16 - C));
            sbias[i] = 0.5f * v;
        }
    }
    fence_proxy_async_smem();   // the B tiles are read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();

    const int n_iter = ((int)p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        // ================= TMA producer: the halo patch of tile i into T[i & 1] =================
        const bool leader = elect_one();
        for (int it = 0; it < n_iter; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            int tw;
            const int q = fast_divmod(tile, p.fd_tw, &tw);
            int th;
            const int n = fast_divmod(q, p.fd_th, &th);
            const int s = it & 1;
            mbar_wait(&t_empty[s], (uint32_t)((it >> 1) & 1) ^ 1u);
            if (leader) {
                mbar_expect_tx(&t_full[s], p.t_bytes);
                tma_load_4d(base + p.off_T[s], &p.tmT, &t_full[s], 0, tw * kCtTW - 2, th * kCtTH - 2, n);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const bool leader = elect_one();
        const uint32_t rbT = (uint32_t)p.rbT, rbH = (uint32_t)p.rbH, rbY = (uint32_t)p.rbY;
        const uint32_t rbA = (uint32_t)(C * 2);                     // B row bytes of stages A and C
        const uint32_t kB = (uint32_t)(ch < 16 ? 16 : ch), rbB = kB * 2u;
        const uint32_t idA = umma_idesc_bf16(128, 16), idB = umma_idesc_bf16(128, (uint32_t)C),
                       idC = umma_idesc_bf16(128, (uint32_t)C2);
        const int ksA = C / 16, ksB = (int)kB / 16, ksC = C / 16;
        for (int it = 0; it < n_iter; ++it) {
            const int s = it & 1;
            const uint32_t ph = (uint32_t)(it & 1);
            const uint32_t T = sbase + p.off_T[s];
            if (p.alias && it > 0) {
                mbar_wait(c_drained, (uint32_t)((it - 1) & 1));
                tc_fence_after();
            }
            mbar_wait(&t_full[s], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            // ---- stage A: h(r, cc) over 16 x 24 positions (3 strips), input y1 = upper half of T's channels
            if (leader) {
                for (int st = 0; st < 3; ++st) {
                    uint32_t first = 0;
                    for (int dr = 0; dr < 3; ++dr)
                        for (int dc = 0; dc < 3; ++dc) {
                            const uint32_t a0 = T + (uint32_t)((dr * kCtPP + 8 * st + dc)) * rbT + (uint32_t)(C * 2);
                            const uint64_t da = umma_desc_kmajor_ex(a0, rbT, (uint32_t)kCtPP * rbT, 0);
                            const uint64_t db = umma_desc_kmajor(sbase + p.off_Ba + (uint32_t)((dr * 3 + dc) * 16) * rbA, rbA);
                            for (int k = 0; k < ksA; ++k) {
                                umma_bf16(tmem_base + p.colA + (uint32_t)(16 * st), da + (uint64_t)(2 * k),
                                          db + (uint64_t)(2 * k), idA, first);
                                first = 1u;
                            }
                        }
                }
                umma_commit(&acc_bar[0]);
            }
            // ---- stage B: y2 over 16 x 16 positions (2 strips), input Hs
            mbar_wait(h_ready, ph);
            tc_fence_after();
            if (leader) {
                const uint32_t Hs = sbase + p.off_H;
                for (int st = 0; st < 2; ++st) {
                    uint32_t first = 0;
                    for (int dr = 0; dr < 3; ++dr)
                        for (int dc = 0; dc < 3; ++dc) {
                            const uint32_t a0 = Hs + (uint32_t)((dr * kCtHP + 8 * st + dc)) * rbH;
                            const uint64_t da = umma_desc_kmajor_ex(a0, rbH, (uint32_t)kCtHP * rbH, 0);
                            const uint64_t db = umma_desc_kmajor(sbase + p.off_Bb + (uint32_t)((dr * 3 + dc) * C) * rbB, rbB);
                            for (int k = 0; k < ksB; ++k) {
                                umma_bf16(tmem_base + p.colB + (uint32_t)(C * st), da + (uint64_t)(2 * k),
                                          db + (uint64_t)(2 * k), idB, first);
                                first = 1u;
                            }
                        }
                }
                umma_commit(&acc_bar[1]);
            }
            // ---- stage C: out over the same 2 strips, K = [y0 | y1] (T, interior offset (2, 2)) + y2 (Y2)
            mbar_wait(y_ready, ph);
            tc_fence_after();
            if (!p.alias && it > 0) {
                mbar_wait(c_drained, (uint32_t)((it - 1) & 1));
                tc_fence_after();
            }
            if (leader) {
                const uint32_t Y2 = sbase + p.off_Y2;
                for (int st = 0; st < 2; ++st) {
                    uint32_t first = 0;
                    const uint32_t aT = T + (uint32_t)(2 * kCtPP + 8 * st + 2) * rbT;
                    const uint64_t daT = umma_desc_kmajor_ex(aT, rbT, (uint32_t)kCtPP * rbT, 0);
                    const uint64_t daY = umma_desc_kmajor_ex(Y2 + (uint32_t)(8 * st) * rbY, rbY, (uint32_t)kCtYP * rbY, 0);
                    for (int src = 0; src < 3; ++src) {
                        const uint64_t db = umma_desc_kmajor(sbase + p.off_Bc + (uint32_t)(src * C2) * rbA, rbA);
                        const uint64_t da = src < 2 ? daT + (uint64_t)(src * (C * 2 / 16)) : daY;
                        for (int k = 0; k < ksC; ++k) {
                            umma_bf16(tmem_base + p.colC + (uint32_t)(C2 * st), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k),
                                      idC, first);
                            first = 1u;
                        }
                    }
                }
                umma_commit(&acc_bar[2]);
                umma_commit(&t_empty[s]);   // every read of T[s] (stages A and C; stage B's epilogue came earlier) is done
            }
        }
    } else {
        // ================= epilogue: 2 groups x 4 warps, a group takes whole strips =================
        const int e = warp - 2, grp = e >> 2, q = warp & 3;
        const int row = q * 32 + lane;            // accumulator row of this thread within a strip
        const int rr = row >> 3, cc8 = row & 7;   // patch row / column inside the strip
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t mH = (uint32_t)(p.rbH / 16 - 1), mY = (uint32_t)(p.rbY / 16 - 1), mT = (uint32_t)(p.rbT / 16 - 1);
        for (int it = 0; it < n_iter; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            int tw;
            const int qq = fast_divmod(tile, p.fd_tw, &tw);
            int th;
            const int n = fast_divmod(qq, p.fd_th, &th);
            const int h0 = th * kCtTH, w0 = tw * kCtTW;
            const int s = it & 1;
            const uint32_t ph = (uint32_t)(it & 1);
            // ---- stage A -> Hs (16 channels, ch valid): zero outside the image = stage B's zero padding
            mbar_wait(&acc_bar[0], ph);
            tc_fence_after();
            for (int st = grp; st < 3; st += 2) {
                uint32_t acc[16];
                tmem_ld16(tmem_base + lane_addr + p.colA + (uint32_t)(16 * st), acc);
                tmem_ld_wait();
                const int ih = h0 - 1 + rr, iw = w0 - 1 + 8 * st + cc8;
                const bool in = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    float v0 = fmaf(__uint_as_float(acc[i]), 0.5f, sbias[i]);
                    float v1 = fmaf(__uint_as_float(acc[i + 1]), 0.5f, sbias[i + 1]);
                    v0 = fmaf(v0, tanh_approx(v0), v0);
                    v1 = fmaf(v1, tanh_approx(v1), v1);
                    pk[i >> 1] = in ? pack_bf16x2(v0, v1) : 0u;
                }
                const uint32_t off = p.off_H + (uint32_t)((rr * kCtHP + 8 * st + cc8) * p.rbH);
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    *reinterpret_cast<uint4*>(base + swz(off + 16u * j, mH)) =
                        make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(h_ready);
            // ---- stage B -> Y2 (+ y1 from T when the bottleneck has a shortcut)
            mbar_wait(&acc_bar[1], ph);
            tc_fence_after();
            {
                const int st = grp;   // 2 strips, one per group
                const uint32_t offT = p.off_T[s] + (uint32_t)(((rr + 2) * kCtPP + 8 * st + cc8 + 2) * p.rbT) + (uint32_t)(C * 2);
                const uint32_t offY = p.off_Y2 + (uint32_t)((rr * kCtYP + 8 * st + cc8) * p.rbY);
                for (int c0 = 0; c0 < C; c0 += 16) {
                    uint32_t acc[16];
                    tmem_ld16(tmem_base + lane_addr + p.colB + (uint32_t)(C * st + c0), acc);
                    uint4 r0 = make_uint4(0u, 0u, 0u, 0u), r1 = r0;
                    if (p.add) {
                        r0 = *reinterpret_cast<const uint4*>(base + swz(offT + (uint32_t)(c0 * 2), mT));
                        r1 = *reinterpret_cast<const uint4*>(base + swz(offT + (uint32_t)(c0 * 2) + 16u, mT));
                    }
                    tmem_ld_wait();
                    const uint32_t rv[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        float v0 = fmaf(__uint_as_float(acc[i]), 0.5f, sbias[16 + c0 + i]);
                        float v1 = fmaf(__uint_as_float(acc[i + 1]), 0.5f, sbias[16 + c0 + i + 1]);
                        v0 = fmaf(v0, tanh_approx(v0), v0) + bf16lo_f(rv[i >> 1]);
                        v1 = fmaf(v1, tanh_approx(v1), v1) + bf16hi_f(rv[i >> 1]);
                        pk[i >> 1] = pack_bf16x2(v0, v1);
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        *reinterpret_cast<uint4*>(base + swz(offY + (uint32_t)(c0 * 2) + 16u * j, mY)) =
                            make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(y_ready);
            // ---- stage C -> global
            mbar_wait(&acc_bar[2], ph);
            tc_fence_after();
            {
                const int st = grp;
                const int oh = h0 + rr, ow = w0 + 8 * st + cc8;
                const bool ok = rr < kCtTH && oh < p.H && ow < p.W;
                __nv_bfloat16* dst = p.y + (((long long)n * p.H + oh) * p.W + ow) * p.y_cstride + p.y_coff;
                const float* sb = sbias + 16 + C;
                for (int c0 = 0; c0 < C2; c0 += 32) {
                    uint32_t acc[32];
                    tmem_ld32(tmem_base + lane_addr + p.colC + (uint32_t)(C2 * st + c0), acc);
                    tmem_ld_wait();
                    if (c0 + 32 >= C2) {
                        // the whole accumulator of this strip has been read: the MMA warp may reuse the columns
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(c_drained);
                    }
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float v0 = fmaf(__uint_as_float(acc[i]), 0.5f, sb[c0 + i]);
                        float v1 = fmaf(__uint_as_float(acc[i + 1]), 0.5f, sb[c0 + i + 1]);
                        v0 = fmaf(v0, tanh_approx(v0), v0);
                        v1 = fmaf(v1, tanh_approx(v1), v1);
                        pk[i >> 1] = pack_bf16x2(v0, v1);
                    }
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<uint4*>(dst + c0 + 8 * j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

static int g_ct_sms = 148, g_ct_max_smem = 0, g_ct_enabled = 1;

int init_c3k2_tc() {
    int dev = 0;
    YL_CUDA(cudaGetDevice(&dev));
    YL_CUDA(cudaDeviceGetAttribute(&g_ct_sms, cudaDevAttrMultiProcessorCount, dev));
    YL_CUDA(cudaDeviceGetAttribute(&g_ct_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    YL_CUDA(cudaFuncSetAttribute(c3k2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_ct_max_smem));
    // Measured (profiles/r02_c3k2_tc.md): per-tile dependency chain TMA -> MMA -> epilogue -> MMA -> epilogue -> MMA ->
    // epilogue costs ~9 us for a 224-pixel tile and TMEM / shared memory allow only 2 tiles in flight per SM, so this
    // version is 1.5x SLOWER than the mma.sync kernel on the thin C3k2 blocks.  It stays available for A/B runs
    // (YL_C3K2_TC=1) and as the tested reference for the epilogue -> swizzled smem -> tcgen05 A-operand mechanism.
    const char* e = getenv("YL_C3K2_TC");
    g_ct_enabled = (e && *e) ? (atoi(e) != 0) : 0;
    return YL_OK;
}

bool c3k2_tc_supported(int c, int c2) {
    // TMEM: 48 + 2c + 2 c2 columns must fit 256 (with stage C aliased over A / B when it does not)
    return g_ct_enabled && (c == 16 || c == 32) && (c2 == 32 || c2 == 64 || c2 == 128);
}

static uint32_t align_to(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// same argument meaning as yl_c3k2_tail (c3k2_fused.cu); returns YL_ERR_UNSUPPORTED when the shape is not built
int launch_c3k2_tc(const yl_tensor* t, const yl_tensor* y, const void* wa, const float* ba, int wa_ci_pad, const void* wb,
                   const float* bb, int wb_ci_pad, const void* w2, const float* b2, int w2_ci_pad, int shortcut,
                   cudaStream_t stream) {
    const int C = t->c / 2, C2 = y->c;
    YL_CHECK((C == 16 || C == 32) && (C2 == 32 || C2 == 64 || C2 == 128), YL_ERR_UNSUPPORTED,
             "c3k2 tcgen05 tail: c = %d, c2 = %d not built", C, C2);
    EncodeTiledFn enc = get_encode_tiled();
    YL_CHECK(enc != nullptr, YL_ERR_CUDA, "yl_init() was not called (TMA encoder unresolved)");
    C3k2TcParams p;
    memset(&p, 0, sizeof(p));
    p.rbT = 2 * C * 2;                       // [y0 | y1]
    p.rbH = 32;                              // 16 channels (C/2 valid, zero padded)
    p.rbY = C * 2;
    p.nb = C;
    {
        const uint64_t es = 2;
        uint64_t dims[4] = {(uint64_t)t->c, (uint64_t)t->w, (uint64_t)t->h, (uint64_t)t->n};
        uint64_t str[3] = {(uint64_t)t->cstride * es, (uint64_t)t->cstride * es * t->w, (uint64_t)t->cstride * es * t->w * t->h};
        uint32_t box[4] = {(uint32_t)t->c, (uint32_t)kCtPP, (uint32_t)kCtPR, 1u};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUtensorMapSwizzle sw = p.rbT == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
        CUresult r = enc(&p.tmT, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, reinterpret_cast<__nv_bfloat16*>(t->data) + t->coff, dims,
                         str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        YL_CHECK(r == CUDA_SUCCESS, YL_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the c3k2 patch", (int)r);
    }
    p.t_bytes = (uint32_t)(kCtPP * kCtPR * p.rbT);
    uint32_t off = 0;
    for (int s = 0; s < 2; ++s) {
        p.off_T[s] = off;
        off = align_to(off + (uint32_t)((kCtPR * kCtPP + 8) * p.rbT), 1024);   // + 8 rows: garbage rows of the last strip
    }
    p.off_H = off;
    off = align_to(off + (uint32_t)(kCtHR * kCtHP * p.rbH), 1024);
    p.off_Y2 = off;
    off = align_to(off + (uint32_t)(kCtYR * kCtYP * p.rbY), 1024);
    const int kB = C / 2 < 16 ? 16 : C / 2;
    p.off_Ba = off;
    off = align_to(off + (uint32_t)(9 * 16 * C * 2), 1024);
    p.off_Bb = off;
    off = align_to(off + (uint32_t)(9 * C * kB * 2), 1024);
    p.off_Bc = off;
    off = align_to(off + (uint32_t)(3 * C2 * C * 2), 1024);
    p.off_bias = off;
    off = align_to(off + (uint32_t)((16 + C + C2) * 4), 16);
    p.off_bar = off;
    off += 12 * 8;
    const size_t smem = (size_t)off + 1024;
    YL_CHECK((int)smem <= g_ct_max_smem, YL_ERR_UNSUPPORTED, "c3k2 tcgen05 tail needs %zu B shared memory", smem);
    p.colA = 0;
    p.colB = 64;
    const int need_c = 2 * C2;
    p.alias = (128 + need_c > 256) ? 1 : 0;
    p.colC = p.alias ? 0u : 128u;
    YL_CHECK(64 + 2 * C <= 128 && need_c <= 256, YL_ERR_UNSUPPORTED, "c3k2 tcgen05 tail: TMEM budget exceeded");
    p.y = reinterpret_cast<__nv_bfloat16*>(y->data);
    p.y_cstride = y->cstride;
    p.y_coff = y->coff;
    p.wa = reinterpret_cast<const __nv_bfloat16*>(wa);
    p.wb = reinterpret_cast<const __nv_bfloat16*>(wb);
    p.w2 = reinterpret_cast<const __nv_bfloat16*>(w2);
    p.ba = ba;
    p.bb = bb;
    p.b2 = b2;
    p.wa_k = 9 * wa_ci_pad;
    p.wb_k = 9 * wb_ci_pad;
    p.wb_ci = wb_ci_pad;
    p.w2_k = w2_ci_pad;
    p.N = t->n;
    p.H = t->h;
    p.W = t->w;
    p.C = C;
    p.C2 = C2;
    p.add = shortcut;
    p.tiles_w = ceil_div(t->w, kCtTW);
    p.tiles_h = ceil_div(t->h, kCtTH);
    p.total_tiles = p.tiles_w * p.tiles_h * t->n;
    p.fd_tw = make_fastdiv(p.tiles_w);
    p.fd_th = make_fastdiv(p.tiles_h);
    const int per_sm = smem * 2 <= (size_t)220 * 1024 ? 2 : 1;
    int grid = g_ct_sms * per_sm;
    if (grid > p.total_tiles) grid = p.total_tiles;
    YL_CUDA(launch_kernel(c3k2_tc_kernel, dim3(grid), dim3(kCtThreads), smem, stream, p));
    YL_LAUNCH_OK("c3k2_tc_kernel");
    return YL_OK;
}

}  // namespace yl
