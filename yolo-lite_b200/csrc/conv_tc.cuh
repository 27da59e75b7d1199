// conv_tc.cuh — what the tcgen05 convolution kernels share: the per-launch parameter block, the tile walk and the
// epilogue (TMEM -> bias / SiLU / residual / Detect decode -> swizzled staging tile -> TMA store).
// Used by conv_tc.cu (one layer per launch) and conv_chain.cu (a chain of layers per launch, one cluster per image).
#pragma once
#include "common.cuh"

namespace yl {

struct ConvTcParams {
    CUtensorMap tmA[4];
    CUtensorMap tmB;
    CUtensorMap tmY[5];          // [0] plain destination, [1..4] parity planes of the 2x-upsampled destination
    int y_map_first, y_map_last; // stores go to tmY[first..last)
    int Ho, Wo, Nimg;            // conv output dims per image, images (flat mode: 1, total pixels, 1)
    int tiles_w, tiles_h, tiles_n;
    int TW, TH, TN;              // A-tile box: TW*TH*TN <= 128 rows (pixels), may span images
    int m_tiles, n_tiles, total_tiles;
    FastDiv fd_ntiles, fd_tiles_w, fd_tiles_h;   // tile index -> (n tile, w tile, h tile, image tile)
    int ksize, stride, pad;
    int ci_pad;                  // K elements per tap in the packed weights
    int kblk, cin_blocks;        // channels per k-iteration, iterations per tap
    int co_tile;                 // UMMA N
    int acc_stride;              // TMEM columns per accumulator stage (co_tile rounded up to the chunk width)
    int stages;
    // halo-patch mode (3x3 stride 1, thin channels): ONE TMA box {kblk, patch_pw, TH+2} per tile holds every
    // tap's A operand (taps are shifted windows of it); the 9-tap weights stay resident in smem
    int patch, patch_pw, patch_bo;
    CUtensorMap tmW3;            // weights as {ci, co, tap}: box {kblk, co_tile, 9} -> smem [tap][co_tile][kblk]
    uint32_t w_bytes;
    // resident-weights mode (non-patch, one N tile, small K): all k-iterations' weight tiles are fetched once per
    // CTA into smem [kiter][co_tile][kblk]; the ring then carries activations only (TMA issue rate, not bytes, is
    // what bounds the thin layers: ~3.3 clk per box row per SM, tools/tma_bench.cu)
    int wres;
    int wearly;                  // issue the resident-weight loads before the PDL dependency wait
    // conv_chain_kernel (conv_chain.cu): ch_mode 1 = full-width patch (3x3 stride 1: one A box per tile and channel block,
    // taps are shifted start addresses); ch_na / ch_nb = stages of the activation / weight ring; ch_bres = the weight tiles
    // of the layer are resident (loaded once per CTA); ch_a_tx / ch_b_tx = bytes of one A / B box
    int ch_mode, ch_na, ch_nb, ch_bres;
    uint32_t ch_a_tx, ch_b_tx;
    uint32_t tmem_cols;
    uint32_t a_bytes, b_bytes;   // per-stage smem footprint (1024-aligned)
    uint32_t b_region;           // smem bytes of the weight area: ring (stages * b_bytes) or resident block
    uint32_t tx_bytes;           // bytes one stage's two TMA boxes deliver
    // epilogue
    int cw;                      // accumulator columns per TMEM load (16 or 32)
    int nchunks;
    int stg_sub;                 // TMEM chunks per TMA store (1 or 2): a store moves stg_sub * cw channels
    int nstore;                  // TMA stores per tile = ceil(nchunks / stg_sub)
    int stg_bufs;                // staging tiles per epilogue group (2 = the store of chunk k overlaps chunk k+1)
    int stg_row_bytes;           // stg_sub * cw * element size: 32, 64 or 128 (= the staging swizzle span)
    uint32_t stg_bytes;          // per staging tile
    int y_f32;
    const float* bias;
    int n_bias;                  // valid bias entries (co_pad)
    int act;
    int epi_kind;                // which epilogue instantiation runs (see conv_tc_kernel)
    const __nv_bfloat16* res;
    long long res_cstride;
    int res_coff, res_c;
    // fused Detect decode (flat 1x1 head convs): see yl_det_epilogue
    int store_y;                 // 0: no NHWC destination (decode only)
    int det_mode;                // yl_det_mode
    float* det_pred;
    int det_nc, det_A, det_anchor0, det_hw, det_w;
    float det_stride;
    long long det_M;             // valid pixels (rows beyond it belong to the ragged last tile)
    float det_conf;              // YL_DET_CLS_FILTER: confidence threshold (strict >)
    uint32_t* det_cand_counts;   //   per-image candidate counters of the NMS workspace
    unsigned long long* det_cand_keys;  // per-image key lists, `det_A` entries each
    // back-to-back mode (conv_tc_kernel<.., B2B = true>, yl_conv_b2b_det): the conv result is not stored; rounded to bf16 it
    // is the A operand of a SECOND 1x1 GEMM (weights tmB3, bias3: the head's last conv) whose accumulator feeds the Detect
    // decode described by det_* (box or class decode)
    int b2b, k2blocks, co_tile3, nchunks3, acc3_stride, n_bias3, c2_ch;
    uint32_t b3_bytes;
    const float* bias3;
    CUtensorMap tmB3;
    unsigned long long* dbg;     // optional timeline slot (8 x %globaltimer ns, written by CTA 0): yl_debug_timeline
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// (compiled out of the production instantiation: DBG is a template parameter of the kernel)
#define YL_STAMP(slot)                                                          \
    do {                                                                        \
        if constexpr (DBG) {                                                    \
            if (p.dbg && blockIdx.x == 0) p.dbg[slot] = globaltimer_ns();       \
        }                                                                       \
    } while (0)

// The tiles one CTA walks: begin, begin + step, ... < end (one-layer launch: blockIdx.x / total_tiles / gridDim.x).
struct TileRange {
    int begin, end, step;
};

// 16-byte load that is coherent at L2 (data another CTA wrote earlier in the same kernel)
__device__ __forceinline__ uint4 ld_cg_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

constexpr int kConvTcThreads = 320;
constexpr int kEpiGroupThreads = 128;

template <int CW>
__device__ __forceinline__ void tmem_ld_cw(uint32_t taddr, uint32_t (&r)[CW]);
template <>
__device__ __forceinline__ void tmem_ld_cw<16>(uint32_t taddr, uint32_t (&r)[16]) {
    tmem_ld16(taddr, r);
}
template <>
__device__ __forceinline__ void tmem_ld_cw<32>(uint32_t taddr, uint32_t (&r)[32]) {
    tmem_ld32(taddr, r);
}

// Fused Detect decode of one accumulator chunk (CW columns, bias already added) of pixel `m` (flat index over
// (image, h, w)): box mode turns each 16-bin group into its DFL expectation and, once the four sides are
// known, writes (cx, cy, w, h) * stride; class mode writes sigmoid(logit).  Consecutive lanes hold consecutive
// pixels = consecutive anchors, so every channel row of the (B, 4+nc, A) prediction gets 128-byte coalesced
// stores (head.py:95-126, block.py:51-69, tal.py:326-350).
// YL_DET_CLS_FILTER: running best class of this thread's pixel over the accumulator chunks; after the last chunk every
// passing anchor (best > conf, strict; ties keep the lowest class: utils/ops.py:203, 242-244) is appended to its
// image's candidate list with the same key the NMS filter kernel would build from the stored score.  Lanes are
// grouped by image (a 128-pixel tile may straddle two images) and each group does ONE atomicAdd.
template <int CW>
__device__ __forceinline__ void det_filter_chunk(const ConvTcParams& p, const float (&v)[CW], int c, long long m, int lane,
                                                 float& best, int& bestc) {
    if (c == 0) {
        best = -INFINITY;
        bestc = 0;
    }
    // The sigmoid is monotone, so the best class of the chunk is where the LOGIT is largest: two passes of min / max /
    // compare instead of 2 MUFU ops per class (the epilogue of these layers was MUFU-bound), then ONE sigmoid.  What a
    // monotone map does not preserve is the reference's tie rule (the FIRST class whose SCORE is maximal wins, also when
    // its logit is a hair smaller but rounds to the same score).  Ties in the score need the two largest logits within
    // `eps` of each other (bound derived from the error of ex2.approx / rcp.approx and the rounding of 1 + e^-v; scores
    // above sigmoid(5) sit where 1 + e^-v loses the difference): then, for the whole warp, the per-class evaluation below
    // decides exactly as before.  Random logits take that path for < 1 % of the warps.
    float m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
    for (int i = 0; i < CW; ++i) {
        if (c * CW + i < p.det_nc) {
            m2 = fmaxf(m2, fminf(m1, v[i]));
            m1 = fmaxf(m1, v[i]);
        }
    }
    const float eps = m1 <= 0.f ? 2e-6f * fmaxf(8.f, -m1) : 1e-3f;
    const bool unsure = !(m1 <= 5.f) || !(m2 < m1 - eps);      // (NaN / -inf logits land here too)
    if (!__any_sync(0xffffffffu, unsure)) {
        int idx = 0;
#pragma unroll
        for (int i = CW - 1; i >= 0; --i)
            if (v[i] == m1) idx = i;                            // (lanes beyond nc never equal m1: it came from a valid one)
        const float sc = __fdividef(1.f, 1.f + __expf(-m1));    // the value YL_DET_CLS would have stored
        if (sc > best) {
            best = sc;
            bestc = c * CW + idx;
        }
    } else {
#pragma unroll
        for (int i = 0; i < CW; ++i) {
            const int ch = c * CW + i;
            if (ch < p.det_nc) {
                const float sc = __fdividef(1.f, 1.f + __expf(-v[i]));
                if (sc > best) {
                    best = sc;
                    bestc = ch;
                }
            }
        }
    }
    if (c != p.nchunks - 1) return;
    const bool emit = (m < p.det_M) && (best > p.det_conf);
    if (__ballot_sync(0xffffffffu, emit) == 0u) return;
    const int b = emit ? (int)(m / p.det_hw) : -1;
    const unsigned grp = __match_any_sync(0xffffffffu, b);
    if (emit) {
        const int leader = __ffs(grp) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(p.det_cand_counts + b, (uint32_t)__popc(grp));
        base = __shfl_sync(grp, base, leader);
        const uint32_t pos = base + (uint32_t)__popc(grp & ((1u << lane) - 1u));
        const int al = (int)(m - (long long)b * p.det_hw);
        const uint32_t idx = (uint32_t)(p.det_anchor0 + al) * (uint32_t)p.det_nc + (uint32_t)bestc;
        p.det_cand_keys[(unsigned long long)b * (unsigned long long)p.det_A + pos] =
            ((unsigned long long)score_to_desc(best) << 32) | idx;
    }
}

template <int CW>
__device__ __forceinline__ void det_decode_chunk(const ConvTcParams& p, const float (&v)[CW], int c, long long m,
                                                 float (&dist)[4], int det_mode) {
    if (m >= p.det_M) return;
    const int b = (int)(m / p.det_hw);
    const int al = (int)(m - (long long)b * p.det_hw);
    float* out = p.det_pred + (long long)b * (4 + p.det_nc) * p.det_A + p.det_anchor0 + al;
    if (det_mode == YL_DET_BOX) {
        if (CW == 32) {  // reg_max == 16: two sides per chunk
            float d2[2];
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                float mx = v[sd * 16];
#pragma unroll
                for (int k = 1; k < 16; ++k) mx = fmaxf(mx, v[sd * 16 + k]);
                float ssum = 0.f, e = 0.f;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const float w = __expf(v[sd * 16 + k] - mx);
                    ssum += w;
                    e = fmaf(w, (float)k, e);
                }
                d2[sd] = __fdividef(e, ssum);
            }
            if (c == 0) {
                dist[0] = d2[0];
                dist[1] = d2[1];
            } else {
                dist[2] = d2[0];
                dist[3] = d2[1];
                const float ax = (float)(al % p.det_w) + 0.5f, ay = (float)(al / p.det_w) + 0.5f;
                const float x1 = ax - dist[0], y1 = ay - dist[1], x2 = ax + dist[2], y2 = ay + dist[3];
                out[0] = (x1 + x2) * 0.5f * p.det_stride;
                out[(long long)p.det_A] = (y1 + y2) * 0.5f * p.det_stride;
                out[2ll * p.det_A] = (x2 - x1) * p.det_stride;
                out[3ll * p.det_A] = (y2 - y1) * p.det_stride;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < CW; ++i) {
            const int ch = c * CW + i;
            if (ch < p.det_nc) out[(long long)(4 + ch) * p.det_A] = __fdividef(1.f, 1.f + __expf(-v[i]));
        }
    }
}

// One epilogue group (4 warps, thread = accumulator row) draining the tiles of its accumulator stage.
// A "store chunk" is stg_sub TMEM chunks (<= 128 B per pixel row) staged in one swizzled tile and written by
// one TMA store.  With two staging tiles the store of chunk k is issued after the barrier of chunk k+1, so the
// group never waits for a TMA store to drain its source: one named barrier per store chunk, and the smem read
// of store k overlaps the TMEM load + math of chunk k+1.
//
// GENERIC == false is the hot instantiation (bf16 destination, no Detect epilogue; activation / residual are compile-time
// ACT / RES): ncu showed ~300 warp instructions per 32-column chunk of which only ~160 were the math, the rest runtime
// flag tests, parameter reloads, swizzle arithmetic and the (disabled) timeline stamps.  GENERIC == true keeps every
// runtime option (fp32 destination, Detect decode / class filter, no NHWC store).
// DET (non-generic): 0 = plain bf16 NHWC store; YL_DET_BOX / YL_DET_CLS / YL_DET_CLS_FILTER = the engine path's head
// convs, whose result feeds the fused Detect epilogue only (no NHWC store at all).
//
// `tr`: the tiles of this CTA; group g takes the 1st, 3rd, ... (g = 0) or 2nd, 4th, ... (g = 1) of them.  `acc_uses`: how
// often this group's accumulator stage has been consumed so far (its mbarrier parity; 0 for a one-layer launch).
// `pm`: where the TENSOR MAPS live (kernel parameter or global memory: TMA cannot take them from shared memory); `p` may
// be a shared-memory copy of the same block.  CHAIN (conv_chain.cu): residual rows were written earlier in the SAME
// kernel, so they are read with L2-coherent loads, and the function returns only when its stores are complete (not merely
// read out of the staging tiles): the next layer of the chain, on another CTA of the cluster, reads them.
template <int CW, bool ACT, bool RES, int DET, bool GENERIC, bool DBG, bool CHAIN = false>
__device__ __forceinline__ void conv_tc_epilogue(const ConvTcParams& p, const ConvTcParams& pm, const TileRange tr,
                                                 uint32_t& acc_uses, int g, int q, int lane, int gtid,
                                                 uint32_t tmem_base, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                                 uint8_t* stg, const float* sbias) {
    const bool k_act = GENERIC ? (p.act != 0) : ACT;
    const bool k_res = GENERIC ? (p.res != nullptr) : RES;
    const bool k_f32 = GENERIC ? (p.y_f32 != 0) : false;
    const int k_det = GENERIC ? p.det_mode : DET;
    const bool k_store = GENERIC ? (p.store_y != 0) : (DET == 0);
    const float bscale = k_act ? 0.5f : 1.0f;   // see the bias staging in the kernel prologue
    const int row = q * 32 + lane;
    const int tw = row % p.TW;
    const int th = (row / p.TW) % p.TH;
    const int tn = row / (p.TW * p.TH);
    const uint32_t swz_mask = (uint32_t)(p.stg_row_bytes / 16 - 1);  // 1, 3 or 7 sixteen-byte chunks
    const uint32_t row_off = (uint32_t)row * (uint32_t)p.stg_row_bytes;
    // a row's bytes never cross a 128-byte line (row pitch 32 / 64 / 128), so the swizzle XOR is a per-thread constant
    const uint32_t swz_xor = ((row_off >> 7) & swz_mask) << 4;
    const uint32_t stg_base = smem_u32(stg);
    const uint32_t sub_bytes = (uint32_t)CW * (k_f32 ? 4u : 2u);
    const int nchunks = p.nchunks, stg_sub = p.stg_sub, nstore = p.nstore;
    const bool dbl = p.stg_bufs == 2;
    // the first warp of the group owns the TMA stores; one elected lane issues / commits / waits on them
    const bool leader = (gtid < 32) && elect_one();
    uint32_t kstore = 0;                       // running store-chunk counter (selects the staging tile)
    int pend = 0, pc0 = 0, pw0 = 0, ph0 = 0, pi0 = 0;  // filled tile whose TMA store is not issued yet
    uint32_t pbuf = 0;
    float det_dist[4] = {0.f, 0.f, 0.f, 0.f};  // decode mode: DFL distances (l, t, r, b) of this thread's pixel
    float det_best = -INFINITY;                // class-filter mode: best score / class of this thread's pixel
    int det_bestc = 0;

    bool first_tile = true;
    for (int tile = tr.begin + g * tr.step; tile < tr.end; tile += 2 * tr.step, first_tile = false) {
        int nt, wt, ht;
        int mt = fast_divmod(tile, p.fd_ntiles, &nt);
        mt = fast_divmod(mt, p.fd_tiles_w, &wt);
        const int it = fast_divmod(mt, p.fd_tiles_h, &ht);
        const int w0 = wt * p.TW, h0 = ht * p.TH, i0 = it * p.TN;
        const int n0 = nt * p.co_tile;

        const __nv_bfloat16* resrow = nullptr;
        if (k_res) {
            const int w = w0 + tw, h = h0 + th, n = i0 + tn;
            if ((tn < p.TN) && (n < p.Nimg) && (h < p.Ho) && (w < p.Wo))
                resrow = p.res + (((long long)n * p.Ho + h) * p.Wo + w) * p.res_cstride + p.res_coff;
        }

        const uint32_t ph = acc_uses & 1u;
        ++acc_uses;
        mbar_wait(&tfull_bar[g], ph);
        tc_fence_after();
        if (leader && g == 0 && first_tile) YL_STAMP(4);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.acc_stride);

        for (int s = 0; s < nstore; ++s, ++kstore) {
            const uint32_t buf = dbl ? (kstore & 1u) * p.stg_bytes : 0u;
            for (int u = 0; u < stg_sub; ++u) {
                const int c = s * stg_sub + u;
                if (c >= nchunks) break;
                const int col0 = n0 + c * CW;  // first output channel of this chunk
                uint32_t acc[CW];
                tmem_ld_cw<CW>(taddr + (uint32_t)(c * CW), acc);
                // residual rows are independent of the accumulator: issue the loads under the TMEM latency
                uint4 rv[CW / 8];
                if (k_res) {
#pragma unroll
                    for (int i = 0; i < CW / 8; ++i) {
                        rv[i] = make_uint4(0u, 0u, 0u, 0u);
                        if (resrow && col0 + i * 8 < p.res_c)
                            rv[i] = CHAIN ? ld_cg_u4(reinterpret_cast<const uint4*>(resrow + col0) + i)
                                          : __ldg(reinterpret_cast<const uint4*>(resrow + col0) + i);
                    }
                }
                tmem_ld_wait();
                if (leader && g == 0 && first_tile && c < 3) YL_STAMP(c == 0 ? 8 : (c == 1 ? 12 : 14));
                if (c == nchunks - 1) {
                    // every TMEM read of this tile has completed: hand the accumulator back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[g]);
                }
                // SiLU(x) = h + h * tanh(h) with h = x / 2 (one MUFU op).  For activated convs the staged bias is b / 2 and
                // the accumulator is scaled by 1/2 in the same FMA: h = fma(acc, 0.5, b / 2) is bit-identical to
                // 0.5 * (acc + b) (a scaling by two commutes with the rounding) and saves an instruction per element.
                float v[CW];
#pragma unroll
                for (int i = 0; i < CW; i += 4) {
                    const float4 b = *reinterpret_cast<const float4*>(sbias + col0 + i);
                    v[i + 0] = fmaf(__uint_as_float(acc[i + 0]), bscale, b.x);
                    v[i + 1] = fmaf(__uint_as_float(acc[i + 1]), bscale, b.y);
                    v[i + 2] = fmaf(__uint_as_float(acc[i + 2]), bscale, b.z);
                    v[i + 3] = fmaf(__uint_as_float(acc[i + 3]), bscale, b.w);
                }
                if (k_act) {
#pragma unroll
                    for (int i = 0; i < CW; ++i) v[i] = fmaf(v[i], tanh_approx(v[i]), v[i]);
                }
                if (k_res) {
#pragma unroll
                    for (int i = 0; i < CW / 8; ++i) {
                        v[i * 8 + 0] += bf16lo_f(rv[i].x); v[i * 8 + 1] += bf16hi_f(rv[i].x);
                        v[i * 8 + 2] += bf16lo_f(rv[i].y); v[i * 8 + 3] += bf16hi_f(rv[i].y);
                        v[i * 8 + 4] += bf16lo_f(rv[i].z); v[i * 8 + 5] += bf16hi_f(rv[i].z);
                        v[i * 8 + 6] += bf16lo_f(rv[i].w); v[i * 8 + 7] += bf16hi_f(rv[i].w);
                    }
                }
                if (GENERIC || DET != 0) {
                    if (k_det == YL_DET_CLS_FILTER)
                        det_filter_chunk<CW>(p, v, c, w0 + row, lane, det_best, det_bestc);
                    else if (k_det)
                        det_decode_chunk<CW>(p, v, c, w0 + row, det_dist, k_det);
                }
                if (!k_store) continue;
                if (leader && g == 0 && first_tile && c < 2) YL_STAMP(c == 0 ? 9 : 13);
                if (u == 0) {
                    // the staging tile about to be overwritten must have been drained by its last TMA store:
                    // every committed store has (two tiles: the newest committed one used this tile, the one
                    // filled last is still pending; one tile: the newest committed one used it)
                    if (leader) bulk_wait_read<0>();
                    named_bar_sync(1 + g, kEpiGroupThreads);
                    // ... and the barrier also says every thread has written + fenced the pending tile
                    if (pend) {
                        if (leader) {
                            for (int m = p.y_map_first; m < p.y_map_last; ++m)
                                tma_store_4d(&pm.tmY[m], stg + pbuf, pc0, pw0, ph0, pi0);
                            bulk_commit();
                        }
                        pend = 0;
                    }
                    if (leader && g == 0 && first_tile && (c == 0 || c == 2)) YL_STAMP(c == 0 ? 10 : 15);
                }
                const uint32_t base_off = row_off + (uint32_t)u * sub_bytes;
                if (k_f32) {
#pragma unroll
                    for (int j = 0; j < CW / 4; ++j) {
                        uint32_t off = base_off + (uint32_t)j * 16u;
                        off ^= ((off >> 7) & swz_mask) << 4;   // fp32 rows of 128 B: a chunk may start a new line
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg_base + buf + off),
                                     "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                                     : "memory");
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < CW / 8; ++j) {
                        const uint32_t off = (base_off + (uint32_t)j * 16u) ^ swz_xor;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_base + buf + off),
                                     "r"(pack_bf16x2(v[8 * j], v[8 * j + 1])),
                                     "r"(pack_bf16x2(v[8 * j + 2], v[8 * j + 3])),
                                     "r"(pack_bf16x2(v[8 * j + 4], v[8 * j + 5])),
                                     "r"(pack_bf16x2(v[8 * j + 6], v[8 * j + 7]))
                                     : "memory");
                    }
                }
            }
            if (!k_store) continue;
            fence_proxy_async_smem();
            if (leader && g == 0 && first_tile && s == 0) YL_STAMP(11);
            if (dbl) {
                pend = 1;
                pbuf = buf;
                pc0 = n0 + s * stg_sub * CW;
                pw0 = w0;
                ph0 = h0;
                pi0 = i0;
            } else {
                named_bar_sync(1 + g, kEpiGroupThreads);
                if (leader) {
                    for (int m = p.y_map_first; m < p.y_map_last; ++m)
                        tma_store_4d(&pm.tmY[m], stg, n0 + s * stg_sub * CW, w0, h0, i0);
                    bulk_commit();
                }
            }
        }
    }
    if (dbl) {
        named_bar_sync(1 + g, kEpiGroupThreads);
        if (leader && pend) {
            for (int m = p.y_map_first; m < p.y_map_last; ++m) tma_store_4d(&pm.tmY[m], stg + pbuf, pc0, pw0, ph0, pi0);
            bulk_commit();
        }
    }
    if (leader && g == 0) YL_STAMP(5);
    // the staging tiles must outlive the TMA engine's reads of them; the global writes themselves are flushed by
    // the grid's completion (which is what the next kernel's griddepcontrol.wait / stream order waits for)
    if (leader) {
        if (CHAIN) bulk_wait<0>();
        else bulk_wait_read<0>();
    }
    if (leader && g == 0) YL_STAMP(6);
}

// Epilogue of the back-to-back mode, one group of 4 warps (thread = accumulator row) per accumulator stage:
//   stage 1  TMEM acc1 -> + bias -> SiLU -> bf16 -> this group's swizzled K-major A2 tiles (one per 64 channels: the layout
//            of the TMA-store staging tiles IS the UMMA A-operand layout) -> a2_full: the MMA warp runs the second GEMM;
//   stage 2  TMEM acc2 -> + bias -> Detect box / class decode straight into the prediction (no NHWC store at all).
template <bool ACT>
__device__ __forceinline__ void conv_tc_epilogue_b2b(const ConvTcParams& p, const TileRange tr, int g, int q, int lane, int gtid,
                                                     uint32_t tmem_base, uint64_t* tfull1, uint64_t* tempty1,
                                                     uint64_t* a2_full, uint64_t* a2_empty, uint64_t* tfull2,
                                                     uint64_t* tempty2, uint8_t* a2, const float* sbias, const float* sbias3) {
    const int row = q * 32 + lane;
    const int tw = row % p.TW;
    const int th = (row / p.TW) % p.TH;
    const int tn = row / (p.TW * p.TH);
    const uint32_t a2_row = smem_u32(a2) + (uint32_t)row * 128u;
    const uint32_t swz = (uint32_t)(row & 7);
    const float bscale = ACT ? 0.5f : 1.0f;
    float det_dist[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t uses = 0;
    for (int tile = tr.begin + g * tr.step; tile < tr.end; tile += 2 * tr.step, ++uses) {
        int nt, wt, ht;
        int mt = fast_divmod(tile, p.fd_ntiles, &nt);
        mt = fast_divmod(mt, p.fd_tiles_w, &wt);
        const int it = fast_divmod(mt, p.fd_tiles_h, &ht);
        const int w = wt * p.TW + tw, hh = ht * p.TH + th, n = it * p.TN + tn;
        const bool inside = (tn < p.TN) && (n < p.Nimg) && (hh < p.Ho) && (w < p.Wo);
        const long long m = inside ? ((long long)n * p.Ho + hh) * p.Wo + w : p.det_M;
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
            if (ACT) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], tanh_approx(v[i]), v[i]);
            }
            const uint32_t dst = a2_row + (uint32_t)(c >> 1) * 16384u;
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
        const uint32_t taddr2 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(2 * p.acc_stride + g * p.acc3_stride);
        for (int c = 0; c < p.nchunks3; ++c) {
            uint32_t acc[32];
            tmem_ld32(taddr2 + (uint32_t)(c * 32), acc);
            tmem_ld_wait();
            if (c == p.nchunks3 - 1) {
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
            det_decode_chunk<32>(p, v, c, m, det_dist, p.det_mode);
        }
    }
}

// host side (conv_tc.cu)
bool encode_map(CUtensorMap* m, CUtensorMapDataType dt, void* base, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle sw);
CUtensorMapSwizzle swizzle_for_bytes(int row_bytes);
// tensor maps of a conv destination ([up == 0]: one map at conv resolution; [up == 1]: four parity planes of the 2x map)
bool encode_out_maps(CUtensorMap* maps, const yl_tensor& y, int Ho, int Wo, int N, bool flat, bool up, const uint32_t* box,
                     CUtensorMapSwizzle sw);
struct ConvTcPlanOpts {
    int chain = 0;          // plan for conv_chain_kernel: tiles never span images (TN = 1, no flat mode), streamed weights only
                            // (no halo patch / resident weights), N tiles <= 128 columns (two CTAs per SM), no batch-size
                            // dependent dispatch
    int chain_cluster = 4;  // CTAs per image (cluster size)
    const yl_conv_args* head = nullptr;   // back-to-back mode: the head's last 1x1 conv (Detect box / class decode, no store)
};
int plan_conv_tc(const yl_conv_args* a, ConvTcParams& p, int* grid_out, size_t* smem_out, const ConvTcPlanOpts* opts = nullptr);

}  // namespace yl
