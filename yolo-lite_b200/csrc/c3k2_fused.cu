// Fused tail of a C3k2 block (c3k = False, one Bottleneck): everything after cv1, in ONE kernel.
//
//   reference (nn/modules/block.py:231-235, 330-343, 720-728), with t = cv1(x) = [y0 | y1] (2c channels):
//       h   = SiLU(BN(conv3x3(y1)))            c   -> c/2       Bottleneck.cv1
//       y2  = y1 + SiLU(BN(conv3x3(h)))        c/2 -> c         Bottleneck.cv2 (+ shortcut)
//       out = SiLU(BN(conv1x1([y0, y1, y2])))  3c  -> c2        C2f.cv2
//
// Run layer by layer these three convolutions are the worst-behaved launches of yolo11n: 8..32-channel tensors
// at 160x160 / 80x80 make 16..64-byte TMA rows (the tcgen05 path is TMA-row-rate bound there, profiles/r01_*),
// and h and y2 make an HBM round trip each.  Here a CTA owns an 8 x 16 pixel tile: the 12 x 20 halo patch of t is
// fetched once (cp.async, zero fill = the convs' zero padding), h (10 x 18) and y2 (8 x 16) live in shared memory
// as bf16 (the same rounding points as the unfused path, so results match it bit for bit up to accumulation
// order), and only `out` is written.  HBM traffic drops from t + 2h + 2*y2 + [y0 y1 y2] + out to t + out.
//
// The three GEMMs are thin (N = 8..128, K = 48..288) and fed from shared memory, so they run on mma.sync
// m16n8k16 (bf16 x bf16 -> fp32) with ldmatrix operand fetches; tcgen05's 128-row UMMA tiles buy nothing at
// these shapes and the kernel is bound by HBM + the legacy tensor pipe, not by issue slots.
//   stage A (h)  : 12 m-tiles over the 180 halo pixels, warps take m-tiles round-robin, B (weights) in registers
//   stage B (y2) : warp w = tile row w (16 pixels), residual y1 read from the patch
//   stage C (out): warp w = tile row w; A fragments ([y0 y1] from the patch, y2 from smem) held in registers, the
//                  output channels swept in pairs of n-tiles; a 4x4 quad transpose lets every lane store 16 bytes
// Persistent grid: weights + biases are staged once per CTA (before the PDL dependency wait), tiles are strided.
#include "common.cuh"

namespace yl {

struct C3k2TailParams {
    const __nv_bfloat16* t;   // [y0 | y1]: 2C channels at t_coff of a buffer with t_cstride channels per pixel
    long long t_cstride;
    int t_coff;
    __nv_bfloat16* y;         // out: C2 channels
    long long y_cstride;
    int y_coff;
    const __nv_bfloat16 *wa, *wb, *w2;   // packed [co_pad][k*k*ci_pad] bf16 (K-major), BN folded
    const float *ba, *bb, *b2;
    int wa_rows, wa_k, wb_k, w2_k;       // valid rows of wa (co_pad), packed row lengths in elements
    int N, H, W, C2, add;
    int tiles_w, tiles_h, total_tiles;
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// row pitch with an odd number of 16-byte units: 8 consecutive rows hit 8 distinct bank groups (ldmatrix)
__host__ __device__ constexpr int pitch16(int data_bytes) { return data_bytes + (((data_bytes / 16) & 1) ? 32 : 16); }

template <int C>
struct C3k2Smem {
    static constexpr int CH = C / 2;                  // hidden channels of the bottleneck
    // tile rows: 16 for the thin variant (less halo per output pixel, 21 of 24 stage-A slots busy, half the barriers
    // per pixel; its smem still allows the 2 CTAs/SM its registers allow), 8 otherwise (two CTAs/SM need <= 113 KB)
    static constexpr int TH = (C == 16) ? 16 : 8, TW = 16;
    static constexpr int PH = TH + 4, PW = TW + 4;    // patch of t
    static constexpr int HH = TH + 2, HW = TW + 2;    // region of h
    static constexpr int TS = pitch16(4 * C);         // bytes per patch pixel (2C bf16)
    static constexpr int HS = pitch16(2 * CH);
    static constexpr int YS = pitch16(2 * C);
    static constexpr int KA = 9 * C, KB = 9 * CH, K2 = 3 * C;
    static constexpr int WAS = pitch16(2 * KA), WBS = pitch16(2 * KB + 16), W2S = pitch16(2 * K2);
    static constexpr int NA = CH < 16 ? 8 : CH;       // rows of wa kept (n-tiles of 8)
    static constexpr int HM = (HH * HW + 15) / 16;    // m-tiles of stage A
    static constexpr int NBUF = C == 16 ? 2 : 1;      // patch buffers (the thin variant prefetches the next tile)
    static constexpr int T_BYTES = PH * PW * TS;
    static constexpr int off_t = 0;
    static constexpr int off_h = off_t + NBUF * T_BYTES;
    static constexpr int off_y2 = off_h + HM * 16 * HS;
    static constexpr int off_wa = off_y2 + TH * TW * YS;
    static constexpr int off_wb = off_wa + NA * WAS;
    static constexpr int off_w2 = off_wb + C * WBS;
    __host__ __device__ static constexpr int w2_bytes(int c2) { return c2 * W2S; }
    __host__ __device__ static constexpr int bias_floats(int c2) { return NA + C + c2; }
    __host__ __device__ static constexpr int total(int c2) { return off_w2 + w2_bytes(c2) + bias_floats(c2) * 4; }
};

// copy `rows` packed weight rows of `kbytes` bytes each into smem rows of `pitch` bytes, zero-filling the tail
__device__ __forceinline__ void stage_weights(uint8_t* dst, const __nv_bfloat16* src, int rows, int src_rows, int kbytes,
                                              int pitch) {
    const int units = pitch / 16;
    for (int i = threadIdx.x; i < rows * units; i += blockDim.x) {
        const int r = i / units, u = i - r * units;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < src_rows && u * 16 < kbytes)   // kbytes is a multiple of 16 (ci_pad % 8 == 0)
            v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(src) + (size_t)r * kbytes) + u);
        *reinterpret_cast<uint4*>(dst + (size_t)r * pitch + u * 16) = v;
    }
}

template <int C>
__global__ void __launch_bounds__(256, 2) c3k2_tail_kernel(const C3k2TailParams p) {
    using S = C3k2Smem<C>;
    constexpr int CH = S::CH, TH = S::TH, TW = S::TW, PW = S::PW, HW = S::HW, HH = S::HH;
    constexpr int TS = S::TS, HS = S::HS, YS = S::YS, WAS = S::WAS, WBS = S::WBS, W2S = S::W2S;
    extern __shared__ __align__(16) uint8_t smem[];
    griddep_launch_dependents();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int lrow = lane & 15, khalf = lane >> 4;   // ldmatrix.x4 A addressing: row of the m-tile, k half
    const uint32_t sT = smem_u32(smem + S::off_t), sH = smem_u32(smem + S::off_h), sY2 = smem_u32(smem + S::off_y2);
    const uint32_t sWa = smem_u32(smem + S::off_wa), sWb = smem_u32(smem + S::off_wb), sW2 = smem_u32(smem + S::off_w2);
    float* sba = reinterpret_cast<float*>(smem + S::off_w2 + S::w2_bytes(p.C2));
    float* sbb = sba + S::NA;
    float* sb2 = sbb + C;

    // ---- constants (weights, biases): staged once per CTA, before the dependency wait
    stage_weights(smem + S::off_wa, p.wa, S::NA, p.wa_rows, p.wa_k * 2, WAS);
    stage_weights(smem + S::off_wb, p.wb, C, C, p.wb_k * 2, WBS);
    stage_weights(smem + S::off_w2, p.w2, p.C2, p.C2, p.w2_k * 2, W2S);
    for (int i = threadIdx.x; i < S::NA; i += blockDim.x) sba[i] = i < CH ? __ldg(p.ba + i) : 0.f;
    for (int i = threadIdx.x; i < C; i += blockDim.x) sbb[i] = __ldg(p.bb + i);
    for (int i = threadIdx.x; i < p.C2; i += blockDim.x) sb2[i] = __ldg(p.b2 + i);
    __syncthreads();   // staged weights visible: the stationary B fragments below come from shared memory

    // ---- stationary B fragments (registers, loaded once per CTA)
    // stage C: warp = (channel group of 32, block of rows): B for 4 n-tiles x all k-steps never leaves registers
    constexpr int KT = 2 * C / 16, KY = C / 16, KC = KT + KY;   // k-steps from the patch ([y0 y1]), from y2, total
    const int NG = p.C2 / 32;                 // channel groups; 8 % NG == 0 (checked by the launcher)
    const int grp = warp % NG, rblk = warp / NG;
    uint32_t bC[KC][4][2];
#pragma unroll
    for (int ks = 0; ks < KC; ++ks)
#pragma unroll
        for (int np = 0; np < 2; ++np)
            ldsm_x4(sW2 + (uint32_t)((grp * 32 + np * 16 + (lane >> 4) * 8 + (lane & 7)) * W2S + (ks * 16 + 8 * ((lane >> 3) & 1)) * 2),
                    bC[ks][2 * np][0], bC[ks][2 * np][1], bC[ks][2 * np + 1][0], bC[ks][2 * np + 1][1]);
    // thin variant (C == 16): the 3x3 weights are small enough to stay in registers too
    constexpr bool THIN = (C == 16);
    constexpr int KSB = (9 * CH + 15) / 16;   // k-steps of stage B
    uint32_t bA[THIN ? 9 : 1][2];
    if (THIN) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
            ldsm_x2(sWa + (uint32_t)((lane & 7) * WAS + (tap * C + 8 * ((lane >> 3) & 1)) * 2), bA[THIN ? tap : 0][0],
                    bA[THIN ? tap : 0][1]);
    }
    griddep_wait();

    constexpr int CHUNKS = 4 * C / 16;   // 16-byte chunks per patch pixel
    auto load_patch = [&](int tile, uint32_t dst) {
        // pixels (h0-2 .. h0+TH+1, w0-2 .. w0+TW+1) of t, 2C channels, zero outside the image
        const int w0 = (tile % p.tiles_w) * TW;
        const int h0 = ((tile / p.tiles_w) % p.tiles_h) * TH;
        const int n = tile / (p.tiles_w * p.tiles_h);
        const __nv_bfloat16* src0 = p.t + (long long)n * p.H * p.W * p.t_cstride + p.t_coff;
        for (int idx = threadIdx.x; idx < S::PH * PW * CHUNKS; idx += blockDim.x) {
            const int pix = idx / CHUNKS, ck = idx - pix * CHUNKS;
            const int pr = pix / PW, pc = pix - pr * PW;
            const int hi = h0 - 2 + pr, wi = w0 - 2 + pc;
            const bool ok = hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            const int off = ok ? hi * p.W + wi : 0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)(pix * TS + ck * 16)),
                         "l"(src0 + (long long)off * p.t_cstride + ck * 8), "r"(ok ? 16 : 0)
                         : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int buf = 0;
    if (S::NBUF == 2 && (int)blockIdx.x < p.total_tiles) load_patch(blockIdx.x, sT);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int w0 = (tile % p.tiles_w) * TW;
        const int h0 = ((tile / p.tiles_w) % p.tiles_h) * TH;
        const int n = tile / (p.tiles_w * p.tiles_h);
        const uint32_t sTc = sT + (uint32_t)(buf * S::T_BYTES);
        if (S::NBUF == 1) load_patch(tile, sTc);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (S::NBUF == 2) {   // prefetch the next tile's patch into the other buffer while this one is computed
            if (tile + (int)gridDim.x < p.total_tiles) load_patch(tile + gridDim.x, sT + (uint32_t)((buf ^ 1) * S::T_BYTES));
            buf ^= 1;
        }

        // ================= stage A: h = SiLU(conv3x3(y1) + ba) on the (TH+2) x (TW+2) halo region =================
        {
            constexpr int NT = (CH + 7) / 8;            // n-tiles (1 for C = 16, 2 for C = 32)
            constexpr int KS_TAP = C / 16;              // k-steps per tap
            for (int mt = warp; mt < S::HM; mt += 8) {
                int idx = mt * 16 + lrow;
                idx = idx < HH * HW ? idx : HH * HW - 1;
                const int hr = idx / HW, hc = idx - hr * HW;
                const uint32_t arow = sTc + (uint32_t)((hr * PW + hc) * TS + (C + 8 * khalf) * 2);   // y1 = channels [C, 2C)
                float acc[NT][4];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float2 bb2 = *reinterpret_cast<const float2*>(sba + nt * 8 + 2 * t);
                    acc[nt][0] = bb2.x; acc[nt][1] = bb2.y; acc[nt][2] = bb2.x; acc[nt][3] = bb2.y;
                }
                if (THIN) {
                    // all nine A fragments in flight before the dependent MMA chain starts
                    uint32_t a9[9][4];
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap)
                        ldsm_x4(arow + (uint32_t)(((tap / 3) * PW + tap % 3) * TS), a9[tap][0], a9[tap][1], a9[tap][2], a9[tap][3]);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) mma16816(acc[0], a9[tap], bA[THIN ? tap : 0][0], bA[THIN ? tap : 0][1]);
                } else {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int dr = tap / 3, dc = tap - 3 * dr;
                        uint32_t a[KS_TAP][4], b[KS_TAP][4];
#pragma unroll
                        for (int ks = 0; ks < KS_TAP; ++ks) {
                            ldsm_x4(arow + (uint32_t)((dr * PW + dc) * TS + ks * 32), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
                            ldsm_x4(sWa + (uint32_t)((((lane >> 4) * 8) + (lane & 7)) * WAS + (tap * C + ks * 16 + 8 * ((lane >> 3) & 1)) * 2),
                                    b[ks][0], b[ks][1], b[ks][2], b[ks][3]);
                        }
#pragma unroll
                        for (int ks = 0; ks < KS_TAP; ++ks) {
                            mma16816(acc[0], a[ks], b[ks][0], b[ks][1]);
                            mma16816(acc[NT - 1], a[ks], b[ks][2], b[ks][3]);
                        }
                    }
                }
                // rows g and g + 8 of the m-tile; h is ZERO outside the image (it is stage B's zero padding)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int e = mt * 16 + g + half * 8;
                    if (e >= HH * HW) continue;
                    const int er = e / HW, ec = e - er * HW;
                    const int hi = h0 - 1 + er, wi = w0 - 1 + ec;
                    const bool in = hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const float v0 = in ? silu_fast(acc[nt][half * 2]) : 0.f;
                        const float v1 = in ? silu_fast(acc[nt][half * 2 + 1]) : 0.f;
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sH + (uint32_t)(e * HS + (nt * 8 + 2 * t) * 2)),
                                     "r"(pack_bf16x2(v0, v1))
                                     : "memory");
                    }
                }
            }
        }
        __syncthreads();

        // ================= stage B: y2 = y1 + SiLU(conv3x3(h) + bb), warp = tile row =================
        for (int r = warp; r < TH; r += 8) {   // warp = tile row(s)
            constexpr int NT = C / 8;
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 bb2 = *reinterpret_cast<const float2*>(sbb + nt * 8 + 2 * t);
                acc[nt][0] = bb2.x; acc[nt][1] = bb2.y; acc[nt][2] = bb2.x; acc[nt][3] = bb2.y;
            }
            const uint32_t hrow = sH + (uint32_t)((r * HW + lrow) * HS);
            if (CH == 8) {
                // k-step = two taps x 8 channels; lanes 16..31 (k half 1) address the second tap.  The 10th "tap"
                // of the last step meets zero weights (the smem weight rows are zero padded)
                uint32_t a[KSB][4], b[KSB][4];
#pragma unroll
                for (int ks = 0; ks < KSB; ++ks) {
                    const int tap = min(2 * ks + khalf, 8);
                    const int dr = tap / 3, dc = tap - 3 * dr;
                    ldsm_x4(hrow + (uint32_t)((dr * HW + dc) * HS), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
                    ldsm_x4(sWb + (uint32_t)(((lane >> 4) * 8 + (lane & 7)) * WBS + (ks * 16 + 8 * ((lane >> 3) & 1)) * 2),
                            b[ks][0], b[ks][1], b[ks][2], b[ks][3]);
                }
#pragma unroll
                for (int ks = 0; ks < KSB; ++ks) {
                    mma16816(acc[0], a[ks], b[ks][0], b[ks][1]);
                    mma16816(acc[1], a[ks], b[ks][2], b[ks][3]);
                }
            } else {
                constexpr int KS_TAP = CH / 16;
#pragma unroll
                for (int k3 = 0; k3 < KSB; k3 += 3) {      // three k-steps' operands in flight at a time
                    uint32_t a[3][4], b[3][NT / 2][4];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const int ks = k3 + j;
                        const int tap = ks / KS_TAP, kk = ks - tap * KS_TAP;
                        const int dr = tap / 3, dc = tap - 3 * dr;
                        ldsm_x4(hrow + (uint32_t)((dr * HW + dc) * HS + (kk * 16 + 8 * khalf) * 2), a[j][0], a[j][1], a[j][2], a[j][3]);
#pragma unroll
                        for (int np = 0; np < NT / 2; ++np)
                            ldsm_x4(sWb + (uint32_t)((np * 16 + (lane >> 4) * 8 + (lane & 7)) * WBS + (ks * 16 + 8 * ((lane >> 3) & 1)) * 2),
                                    b[j][np][0], b[j][np][1], b[j][np][2], b[j][np][3]);
                    }
#pragma unroll
                    for (int j = 0; j < 3; ++j)
#pragma unroll
                        for (int np = 0; np < NT / 2; ++np) {
                            mma16816(acc[2 * np], a[j], b[j][np][0], b[j][np][1]);
                            mma16816(acc[2 * np + 1], a[j], b[j][np][2], b[j][np][3]);
                        }
                }
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int col = g + half * 8;
                const uint32_t y1p = sTc + (uint32_t)(((r + 2) * PW + col + 2) * TS + C * 2);
                const uint32_t y2p = sY2 + (uint32_t)((r * TW + col) * YS);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    float v0 = silu_fast(acc[nt][half * 2]), v1 = silu_fast(acc[nt][half * 2 + 1]);
                    if (p.add) {
                        uint32_t rv;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(rv) : "r"(y1p + (uint32_t)((nt * 8 + 2 * t) * 2)));
                        v0 += bf16lo_f(rv);
                        v1 += bf16hi_f(rv);
                    }
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(y2p + (uint32_t)((nt * 8 + 2 * t) * 2)), "r"(pack_bf16x2(v0, v1))
                                 : "memory");
                }
            }
        }
        __syncthreads();   // stage C reads y2 rows written by other warps

        // ================= stage C: out = SiLU(conv1x1([y0 y1 y2]) + b2) =================
        // warp = (32-channel group `grp`, a block of TH*NG/8 rows): the weights are the stationary register operand,
        // the activations stream through ldmatrix (KC loads feed 4*KC MMAs)
        {
            // after the quad transpose lane t stores pixel col g + 8 * (t >> 1), channels 16 * np + 8 * (t & 1) ..+8
            const int scol = g + 8 * (t >> 1);
            const int wi = w0 + scol;
            const int rpw = TH * NG / 8;          // rows per warp: the 8 / NG warps of a channel group share the TH rows
            for (int rr = 0; rr < rpw; ++rr) {
                const int r = rblk * rpw + rr;
                uint32_t a[KC][4];
                const uint32_t trow = sTc + (uint32_t)(((r + 2) * PW + lrow + 2) * TS + 8 * khalf * 2);
                const uint32_t yrow = sY2 + (uint32_t)((r * TW + lrow) * YS + 8 * khalf * 2);
#pragma unroll
                for (int ks = 0; ks < KT; ++ks) ldsm_x4(trow + ks * 32, a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
#pragma unroll
                for (int ks = 0; ks < KY; ++ks) ldsm_x4(yrow + ks * 32, a[KT + ks][0], a[KT + ks][1], a[KT + ks][2], a[KT + ks][3]);
                float acc[4][4];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const float2 bb2 = *reinterpret_cast<const float2*>(sb2 + grp * 32 + nt * 8 + 2 * t);
                    acc[nt][0] = bb2.x; acc[nt][1] = bb2.y; acc[nt][2] = bb2.x; acc[nt][3] = bb2.y;
                }
#pragma unroll
                for (int ks = 0; ks < KC; ++ks)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma16816(acc[nt], a[ks], bC[ks][nt][0], bC[ks][nt][1]);
                const int hi = h0 + r;
                const bool st_ok = hi < p.H && wi < p.W;
                __nv_bfloat16* yp = p.y + (((long long)n * p.H + hi) * p.W + wi) * p.y_cstride + p.y_coff + grp * 32 + 8 * (t & 1);
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    uint32_t v0 = pack_bf16x2(silu_fast(acc[2 * np][0]), silu_fast(acc[2 * np][1]));           // (col g,   n-tile 0)
                    uint32_t v1 = pack_bf16x2(silu_fast(acc[2 * np + 1][0]), silu_fast(acc[2 * np + 1][1]));   // (col g,   n-tile 1)
                    uint32_t v2 = pack_bf16x2(silu_fast(acc[2 * np][2]), silu_fast(acc[2 * np][3]));           // (col g+8, n-tile 0)
                    uint32_t v3 = pack_bf16x2(silu_fast(acc[2 * np + 1][2]), silu_fast(acc[2 * np + 1][3]));   // (col g+8, n-tile 1)
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
        }
        __syncthreads();   // the next iteration overwrites sH / sY2 (and, single-buffered, the patch)
    }
}

static int g_c3k2_sms = 148, g_c3k2_max_smem = 0;

template <int C>
static int launch_c3k2_tail(C3k2TailParams p, cudaStream_t s) {
    using S = C3k2Smem<C>;
    p.tiles_w = ceil_div(p.W, S::TW);
    p.tiles_h = ceil_div(p.H, S::TH);
    p.total_tiles = p.tiles_w * p.tiles_h * p.N;
    const size_t smem = (size_t)S::total(p.C2);
    YL_CHECK((int)smem <= g_c3k2_max_smem, YL_ERR_UNSUPPORTED, "c3k2 tail needs %zu B of shared memory (device allows %d)", smem,
             g_c3k2_max_smem);
    int per_sm = (int)((size_t)220 * 1024 / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);   // __launch_bounds__(256, 2): registers cap residency at 2
    int grid = g_c3k2_sms * per_sm;
    if (grid > p.total_tiles) grid = p.total_tiles;
    YL_CUDA(launch_kernel(c3k2_tail_kernel<C>, dim3(grid), dim3(256), smem, s, p));
    YL_LAUNCH_OK("c3k2_tail_kernel");
    return YL_OK;
}

// per-device setup (called by yl_init): opt both instantiations into the device's full dynamic shared memory
int init_c3k2() {
    int dev = 0;
    YL_CUDA(cudaGetDevice(&dev));
    YL_CUDA(cudaDeviceGetAttribute(&g_c3k2_sms, cudaDevAttrMultiProcessorCount, dev));
    YL_CUDA(cudaDeviceGetAttribute(&g_c3k2_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    YL_CUDA(cudaFuncSetAttribute(c3k2_tail_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_c3k2_max_smem));
    YL_CUDA(cudaFuncSetAttribute(c3k2_tail_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_c3k2_max_smem));
    return YL_OK;
}

}  // namespace yl

namespace yl {
bool c3k2_tc_supported(int c, int c2);
int launch_c3k2_tc(const yl_tensor* t, const yl_tensor* y, const void* wa, const float* ba, int wa_ci_pad, const void* wb,
                   const float* bb, int wb_ci_pad, const void* w2, const float* b2, int w2_ci_pad, int shortcut,
                   cudaStream_t stream);
}  // namespace yl

// The tcgen05 version of the same block (c3k2_tc.cu), whatever YL_C3K2_TC says: A/B runs and parity tests.
extern "C" int yl_c3k2_tail_tc(const yl_tensor* t, const yl_tensor* y, const void* wa, const float* ba, int wa_co_pad,
                               int wa_ci_pad, const void* wb, const float* bb, int wb_ci_pad, const void* w2,
                               const float* b2, int w2_ci_pad, int shortcut, void* stream) {
    YL_CHECK(t && y && t->data && y->data && wa && wb && w2 && ba && bb && b2, YL_ERR_ARG, "null pointer");
    YL_CHECK(t->dtype == YL_BF16 && y->dtype == YL_BF16, YL_ERR_ARG, "c3k2 tail tensors must be bf16");
    YL_CHECK(t->n == y->n && t->h == y->h && t->w == y->w && t->c % 32 == 0, YL_ERR_ARG, "c3k2 tail shape mismatch");
    YL_CHECK(t->coff % 8 == 0 && t->cstride % 8 == 0 && y->coff % 8 == 0 && y->cstride % 8 == 0, YL_ERR_ARG,
             "c3k2 tail needs 8-channel alignment");
    const int c = t->c / 2;
    YL_CHECK(wa_co_pad >= c / 2 && wa_ci_pad == c && wb_ci_pad == (c / 2 + 7) / 8 * 8 && w2_ci_pad == 3 * c, YL_ERR_ARG,
             "packed weight layouts do not match the block");
    return yl::launch_c3k2_tc(t, y, wa, ba, wa_ci_pad, wb, bb, wb_ci_pad, w2, b2, w2_ci_pad, shortcut, (cudaStream_t)stream);
}

extern "C" int yl_c3k2_tail_supported(int c, int c2) {
    return (c == 16 || c == 32) && (c2 == 32 || c2 == 64 || c2 == 128 || c2 == 256);
}

extern "C" int yl_c3k2_tail(const yl_tensor* t, const yl_tensor* y, const void* wa, const float* ba, int wa_co_pad,
                            int wa_ci_pad, const void* wb, const float* bb, int wb_ci_pad, const void* w2, const float* b2,
                            int w2_ci_pad, int shortcut, void* stream) {
    YL_CHECK(t && y && t->data && y->data && wa && wb && w2 && ba && bb && b2, YL_ERR_ARG, "null pointer");
    YL_CHECK(t->dtype == YL_BF16 && y->dtype == YL_BF16, YL_ERR_ARG, "c3k2 tail tensors must be bf16");
    YL_CHECK(t->n == y->n && t->h == y->h && t->w == y->w, YL_ERR_ARG, "c3k2 tail shape mismatch");
    YL_CHECK(t->c % 32 == 0 && yl_c3k2_tail_supported(t->c / 2, y->c), YL_ERR_UNSUPPORTED,
             "c3k2 tail is built for c in {16, 32} and c2 in {32, 64, 128, 256} (got c = %d, c2 = %d)", t->c / 2, y->c);
    YL_CHECK(t->coff % 8 == 0 && t->cstride % 8 == 0 && y->coff % 8 == 0 && y->cstride % 8 == 0, YL_ERR_ARG,
             "c3k2 tail needs 8-channel alignment");
    YL_CHECK(((uintptr_t)t->data | (uintptr_t)y->data | (uintptr_t)wa | (uintptr_t)wb | (uintptr_t)w2) % 16 == 0, YL_ERR_ARG,
             "pointers must be 16-byte aligned");
    const int c = t->c / 2;
    YL_CHECK(wa_ci_pad == c && wb_ci_pad == (c / 2 + 7) / 8 * 8 && w2_ci_pad == 3 * c, YL_ERR_ARG,
             "packed weight layouts do not match the block (ci_pad %d / %d / %d for c = %d)", wa_ci_pad, wb_ci_pad,
             w2_ci_pad, c);
    YL_CHECK((long long)t->h * t->w < (1ll << 31), YL_ERR_ARG, "image too large");
    // tcgen05 version (c3k2_tc.cu) for the shapes it is built for; the mma.sync kernel below covers the rest
    if (yl::c3k2_tc_supported(c, y->c))
        return yl::launch_c3k2_tc(t, y, wa, ba, wa_ci_pad, wb, bb, wb_ci_pad, w2, b2, w2_ci_pad, shortcut, (cudaStream_t)stream);
    yl::C3k2TailParams p;
    p.t = reinterpret_cast<const __nv_bfloat16*>(t->data);
    p.t_cstride = t->cstride;
    p.t_coff = t->coff;
    p.y = reinterpret_cast<__nv_bfloat16*>(y->data);
    p.y_cstride = y->cstride;
    p.y_coff = y->coff;
    p.wa = reinterpret_cast<const __nv_bfloat16*>(wa);
    p.wb = reinterpret_cast<const __nv_bfloat16*>(wb);
    p.w2 = reinterpret_cast<const __nv_bfloat16*>(w2);
    p.ba = ba;
    p.bb = bb;
    p.b2 = b2;
    p.wa_rows = wa_co_pad;
    p.wa_k = 9 * wa_ci_pad;
    p.wb_k = 9 * wb_ci_pad;
    p.w2_k = w2_ci_pad;
    p.N = t->n;
    p.H = t->h;
    p.W = t->w;
    p.C2 = y->c;
    p.add = shortcut;
    cudaStream_t s = (cudaStream_t)stream;
    return c == 16 ? yl::launch_c3k2_tail<16>(p, s) : yl::launch_c3k2_tail<32>(p, s);
}
