// A chain of convolution layers in ONE launch: the small feature maps of the network (20x20 / 40x40 at 640^2 input).
//
// Layer by layer those convolutions are latency chains, not bandwidth or math: at bs = 64 a 128->128 1x1 on 20x20 maps is
// 200 tiles of work (~2 us of tensor-core + epilogue time on a fraction of the GPU) but costs 7.5-10 us as a launch
// (TMEM alloc + barrier init + descriptor fetch, dependency wait, TMA latency, MMA, epilogue, store drain, grid
// completion + flush before the next kernel may start).  yolo11n has ~45 such launches per step
// (block.py:165-184, 720-739, 999-1038 of the reference: C3k chains, SPPF, C2PSA).
//
// Every image is independent through a convolution chain, so ONE THREAD-BLOCK CLUSTER owns one image for the whole
// chain: the 4 CTAs of a cluster split the (M, N) tiles of each layer, write their results with TMA stores (they stay in
// L2), meet at a hardware cluster barrier, and start the next layer with TMA loads of what the other CTAs just wrote.
// No grid-wide synchronisation exists (clusters are gang-scheduled by the hardware, images never wait for each other),
// and per layer the launch overhead is replaced by: wait for the stores to complete + one cluster barrier (~1 us).
// TMEM, the mbarriers and the shared-memory carve-up are set up once per chain.
//
// The per-layer work is exactly conv_tc.cu's: same tiles (never spanning images here), same tcgen05.mma order, same
// epilogue code (conv_tc.cuh), so a chain is bit-identical to its layers launched one by one.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "conv_tc.cuh"

namespace yl {

constexpr int kChainParamSlot = 2304;                  // one shared-memory copy of a layer's parameter block
constexpr int kChainMaxA = 8, kChainMaxB = 12;          // stages of the activation / weight ring (mbarrier slots)
constexpr int kChainBarBytes = 384;                     // mbarriers + TMEM slot
constexpr int kChainHeader = kChainBarBytes + 2 * kChainParamSlot;   // ... and two parameter slots
constexpr int kChainHeaderBudget = 5632;                // what plan_conv_tc(chain) leaves free (incl. alignment slack)
static_assert((2 * kChainMaxA + 2 * kChainMaxB + 5) * 8 + 4 <= kChainBarBytes, "barrier area");
static_assert(sizeof(ConvTcParams) / 16 <= kConvTcThreads, "one cp.async per thread moves a parameter block");
static_assert(sizeof(ConvTcParams) <= kChainParamSlot, "parameter block outgrew its shared-memory slot");
static_assert(sizeof(ConvTcParams) % 16 == 0, "parameter block is copied in 16-byte pieces");
static_assert(kChainHeader + 128 <= kChainHeaderBudget, "chain header budget");

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
// Cluster barrier, every thread of every CTA takes part.  `publish`: this warp holds a thread whose completed TMA stores
// the other CTAs are about to read (release at cluster scope = one MEMBAR for this warp only); every other warp arrives
// relaxed: it has written nothing another CTA reads, and a release by all 320 threads costs ~2 us per layer (MEMBAR.GPU).
__device__ __forceinline__ void cluster_arrive(bool publish) {
    if (publish) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    else asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// orders generic-proxy and async-proxy (TMA) accesses of this thread to global memory
// (the .global form is a view fence only; the unqualified one adds a MEMBAR.GPU, ~1 us per layer when all threads run it)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// grid = clusters x cluster size; cluster c owns image c.  `layers`: n_layers parameter blocks in GLOBAL memory (the
// TMA unit reads tensor maps from global / const / param space only), planned by plan_conv_tc(chain).
__global__ void __launch_bounds__(kConvTcThreads, 2)
conv_chain_kernel(const ConvTcParams* __restrict__ layers, int n_layers, uint32_t tmem_cols,
                  unsigned long long* __restrict__ dbg) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* hdr = smem_raw + (((raw + 127u) & ~127u) - raw);
    uint64_t* fullA = reinterpret_cast<uint64_t*>(hdr);      // [kChainMaxA]
    uint64_t* emptyA = fullA + kChainMaxA;
    uint64_t* fullB = emptyA + kChainMaxA;                   // [kChainMaxB]
    uint64_t* emptyB = fullB + kChainMaxB;
    uint64_t* tfull_bar = emptyB + kChainMaxB;               // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                    // [2]
    uint64_t* wres_bar = tempty_bar + 2;                     // resident weights of the layer landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wres_bar + 1);
    uint8_t* const slot0 = hdr + kChainBarBytes;   // parameter slot of layer L: slot0 + (L & 1) * kChainParamSlot
    const uint32_t ring_raw = smem_u32(hdr + kChainHeader);
    uint8_t* ring = hdr + kChainHeader + (((ring_raw + 1023u) & ~1023u) - ring_raw);

    griddep_launch_dependents();
    if (threadIdx.x == 0) {
        for (int s = 0; s < kChainMaxA; ++s) {
            mbar_init(&fullA[s], 1);
            mbar_init(&emptyA[s], 1);
        }
        for (int s = 0; s < kChainMaxB; ++s) {
            mbar_init(&fullB[s], 1);
            mbar_init(&emptyB[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 4);
        }
        mbar_init(wres_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, tmem_cols);
        tmem_relinquish();
    }
    {
        const uint4* src = reinterpret_cast<const uint4*>(layers);
        uint4* dst = reinterpret_cast<uint4*>(slot0);
        for (int i = threadIdx.x; i < (int)(sizeof(ConvTcParams) / 16); i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&layers[0].tmA[0]);
        tma_prefetch_desc(&layers[0].tmB);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int rank = (int)cluster_ctarank();
    const int csize = (int)cluster_nctarank();
    const int image = (int)cluster_id_x();

    // ring / accumulator positions survive from layer to layer: one parity bit per mbarrier, kept by the warp that waits
    uint32_t pa_bits = 0, pb_bits = 0;   // producer: parity of each A / B stage's next fill
    uint32_t ca_bits = 0, cb_bits = 0;   // MMA issuer: parity of each A / B stage's next full phase
    uint32_t wres_uses = 0;              // MMA issuer: layers that used the resident-weight barrier
    uint32_t acc_uses0 = 0, acc_uses1 = 0;   // MMA issuer: uses of accumulator stage 0 / 1
    uint32_t epi_uses = 0;               // epilogue group: uses of its accumulator stage

    // activations written by the previous kernel of the stream are read from here on
    griddep_wait();

    // debug timeline (yl_conv_chain_debug): CTA 0 stamps %globaltimer per layer: [L][0] layer start, [1] first operands
    // landed, [2] last MMA issued, [3] epilogue entered, [4] epilogue returned (stores complete), [5] cluster barrier passed
    const bool stamp = dbg != nullptr && blockIdx.x == 0;
    for (int L = 0; L < n_layers; ++L) {
        if (stamp && threadIdx.x == 0) dbg[8 * L + 0] = globaltimer_ns();
        const ConvTcParams& p = *reinterpret_cast<const ConvTcParams*>(slot0 + (L & 1) * kChainParamSlot);
        const ConvTcParams& pm = layers[L];
        // the next layer's parameter block travels to the other slot while this layer runs
        if (L + 1 < n_layers && threadIdx.x < sizeof(ConvTcParams) / 16)
            cp_async16(slot0 + ((L + 1) & 1) * kChainParamSlot + 16 * threadIdx.x,
                       reinterpret_cast<const uint8_t*>(&layers[L + 1]) + 16 * threadIdx.x);

        uint8_t* sA = ring;
        uint8_t* sB = sA + (size_t)p.ch_na * p.a_bytes;
        uint8_t* sStg = sB + (size_t)p.ch_nb * p.b_bytes;
        float* sbias = reinterpret_cast<float*>(sStg + 2 * (size_t)p.stg_bufs * p.stg_bytes);
        const int nbias = p.n_tiles * p.co_tile + 32;
        const int per_image = p.tiles_w * p.tiles_h * p.n_tiles;
        const TileRange tr = {image * per_image + rank, (image + 1) * per_image, csize};
        const int taps = p.ksize * p.ksize;
        const int kiters = taps * p.cin_blocks;
        const bool patch = p.ch_mode != 0;
        const bool bres = p.ch_bres != 0;

        if (warp == 0) {
            // ================= TMA producer =================
            // loads are issued in the order the MMA warp consumes them: per tile [patch mode: its cin_blocks A boxes], then
            // per k-iteration [per-tap mode: the A box] [streamed weights: the B box]
            const bool leader = elect_one();
            if (leader && L + 1 < n_layers) {
                tma_prefetch_desc(&layers[L + 1].tmA[0]);
                tma_prefetch_desc(&layers[L + 1].tmB);
            }
            int sa = 0, sb = 0;
            bool first = true;
            for (int tile = tr.begin; tile < tr.end; tile += tr.step) {
                int nt, wt, ht;
                int mt = fast_divmod(tile, p.fd_ntiles, &nt);
                mt = fast_divmod(mt, p.fd_tiles_w, &wt);
                const int it = fast_divmod(mt, p.fd_tiles_h, &ht);
                const int w0 = wt * p.TW, h0 = ht * p.TH, i0 = it * p.TN;
                const int n0 = nt * p.co_tile;
                if (first && bres) {
                    // every tile of this CTA has the same N tile (host: cluster size % n_tiles == 0)
                    if (leader) {
                        mbar_expect_tx(wres_bar, (uint32_t)kiters * p.ch_b_tx);
                        int ki = 0;
                        for (int tap = 0; tap < taps; ++tap)
                            for (int cb = 0; cb < p.cin_blocks; ++cb, ++ki)
                                tma_load_2d(sB + (size_t)ki * p.b_bytes, &pm.tmB, wres_bar, tap * p.ci_pad + cb * p.kblk, n0);
                    }
                }
                first = false;
                if (patch) {
                    for (int cb = 0; cb < p.cin_blocks; ++cb) {
                        mbar_wait(&emptyA[sa], ((pa_bits >> sa) & 1u) ^ 1u);
                        pa_bits ^= 1u << sa;
                        if (leader) {
                            mbar_expect_tx(&fullA[sa], p.ch_a_tx);
                            tma_load_4d(sA + (size_t)sa * p.a_bytes, &pm.tmA[0], &fullA[sa], cb * p.kblk, -1, h0 - 1, i0);
                        }
                        if (++sa == p.ch_na) sa = 0;
                    }
                }
                int tap = 0;
                for (int r = 0; r < p.ksize; ++r) {
                    for (int s2 = 0; s2 < p.ksize; ++s2, ++tap) {
                        const int offh = r - p.pad, offw = s2 - p.pad;
                        int map = 0, dh = offh, dw = offw;
                        if (p.stride == 2) {
                            const int ph2 = offh & 1, pw2 = offw & 1;
                            map = ph2 * 2 + pw2;
                            dh = (offh - ph2) >> 1;
                            dw = (offw - pw2) >> 1;
                        }
                        for (int cb = 0; cb < p.cin_blocks; ++cb) {
                            if (!patch) {
                                mbar_wait(&emptyA[sa], ((pa_bits >> sa) & 1u) ^ 1u);
                                pa_bits ^= 1u << sa;
                                if (leader) {
                                    mbar_expect_tx(&fullA[sa], p.ch_a_tx);
                                    tma_load_4d(sA + (size_t)sa * p.a_bytes, &pm.tmA[map], &fullA[sa], cb * p.kblk, w0 + dw,
                                                h0 + dh, i0);
                                }
                                if (++sa == p.ch_na) sa = 0;
                            }
                            if (!bres) {
                                mbar_wait(&emptyB[sb], ((pb_bits >> sb) & 1u) ^ 1u);
                                pb_bits ^= 1u << sb;
                                if (leader) {
                                    mbar_expect_tx(&fullB[sb], p.ch_b_tx);
                                    tma_load_2d(sB + (size_t)sb * p.b_bytes, &pm.tmB, &fullB[sb],
                                                tap * p.ci_pad + cb * p.kblk, n0);
                                }
                                if (++sb == p.ch_nb) sb = 0;
                            }
                        }
                    }
                }
            }
        } else if (warp == 1) {
            // ================= MMA issuer =================
            const bool leader = elect_one();
            const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)p.co_tile);
            const uint32_t rb = (uint32_t)p.kblk * 2u;
            const int ksteps = p.kblk / 16;
            int sa = 0, sb = 0, acc = 0;
            bool first = true;
            for (int tile = tr.begin; tile < tr.end; tile += tr.step) {
                const uint32_t uses = acc ? acc_uses1 : acc_uses0;
                mbar_wait(&tempty_bar[acc], (uses & 1u) ^ 1u);
                if (acc) ++acc_uses1;
                else ++acc_uses0;
                if (first && bres) {
                    mbar_wait(wres_bar, wres_uses & 1u);
                    ++wres_uses;
                }
                const int sa_tile = sa;      // patch mode: first A stage of this tile's channel blocks
                if (patch) {
                    for (int cb = 0; cb < p.cin_blocks; ++cb) {
                        mbar_wait(&fullA[sa], (ca_bits >> sa) & 1u);
                        ca_bits ^= 1u << sa;
                        if (++sa == p.ch_na) sa = 0;
                    }
                }
                tc_fence_after();
                if (stamp && leader && first) dbg[8 * L + 1] = globaltimer_ns();
                first = false;
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
                int ki = 0;
                for (int tap = 0; tap < taps; ++tap) {
                    // patch mode: tap (r, s) = the patch read from (r * TW + s) rows further on
                    const uint32_t tap_off = patch ? (uint32_t)((tap / 3) * p.TW + (tap % 3)) * rb : 0u;
                    for (int cb = 0; cb < p.cin_blocks; ++cb, ++ki) {
                        int a_st;
                        if (patch) {
                            a_st = sa_tile + cb;
                            if (a_st >= p.ch_na) a_st -= p.ch_na;
                        } else {
                            a_st = sa;
                            mbar_wait(&fullA[sa], (ca_bits >> sa) & 1u);
                            ca_bits ^= 1u << sa;
                        }
                        const int b_st = bres ? ki : sb;
                        if (!bres) {
                            mbar_wait(&fullB[sb], (cb_bits >> sb) & 1u);
                            cb_bits ^= 1u << sb;
                        }
                        tc_fence_after();
                        if (leader) {
                            const uint64_t da = umma_desc_kmajor(smem_u32(sA + (size_t)a_st * p.a_bytes) + tap_off, rb);
                            const uint64_t db = umma_desc_kmajor(smem_u32(sB + (size_t)b_st * p.b_bytes), rb);
                            for (int k = 0; k < ksteps; ++k)
                                umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                          (ki > 0 || k > 0) ? 1u : 0u);
                            if (!patch) umma_commit(&emptyA[sa]);
                            if (!bres) umma_commit(&emptyB[sb]);
                        }
                        if (!patch && ++sa == p.ch_na) sa = 0;
                        if (!bres && ++sb == p.ch_nb) sb = 0;
                    }
                }
                if (leader) {
                    if (patch) {
                        int a_st = sa_tile;
                        for (int cb = 0; cb < p.cin_blocks; ++cb) {
                            umma_commit(&emptyA[a_st]);
                            if (++a_st == p.ch_na) a_st = 0;
                        }
                    }
                    umma_commit(&tfull_bar[acc]);
                }
                acc ^= 1;
            }
            if (stamp && leader) dbg[8 * L + 2] = globaltimer_ns();
        } else {
            // ================= epilogue =================
            const int e = warp - 2;
            const int g = e >> 2;
            const int q = warp & 3;
            const int gtid = (e & 3) * 32 + lane;
            uint8_t* stg = sStg + (size_t)g * p.stg_bufs * p.stg_bytes;
            {
                // (the previous layer's readers of sbias are behind the cluster barrier)
                const float bs = p.act ? 0.5f : 1.0f;
                for (int i = e * 32 + lane; i < nbias; i += 2 * kEpiGroupThreads)
                    sbias[i] = i < p.n_bias ? bs * __ldg(p.bias + i) : 0.f;
                named_bar_sync(3, 2 * kEpiGroupThreads);
            }
#define YL_EPI(ACT_, RES_)                                                                                          \
    conv_tc_epilogue<32, ACT_, RES_, 0, false, false, true>(p, pm, tr, epi_uses, g, q, lane, gtid, tmem_base, tfull_bar, \
                                                            tempty_bar, stg, sbias)
            if (stamp && e == 0 && lane == 0) dbg[8 * L + 3] = globaltimer_ns();
            const int kind = p.epi_kind;
            if (kind == 0) YL_EPI(true, false);
            else if (kind == 1) YL_EPI(true, true);
            else if (kind == 2) YL_EPI(false, false);
            else YL_EPI(false, true);
#undef YL_EPI
            if (stamp && e == 0 && lane == 0) dbg[8 * L + 4] = globaltimer_ns();
        }
        __syncwarp();
        // layer boundary: this CTA's stores are complete (each epilogue group's leader waited on its bulk groups: the
        // data is in L2), the next parameter block has landed; then every CTA of the cluster has finished the layer.
        // Only the two warps that hold a store leader publish (proxy fence + release); only the producer, whose TMA
        // loads read what the other CTAs stored, needs the proxy fence on the way out.
        cp_async_commit_wait_all();
        __syncthreads();     // CTA-level ordering of the shared-memory hand-overs (parameter slot, bias / staging / ring reuse)
        const bool publisher = (warp == 2 || warp == 6);
        if (publisher) fence_proxy_async_global();
        cluster_arrive(publisher);
        cluster_wait();
        if (warp == 0) fence_proxy_async_global();
        if (stamp && threadIdx.x == 0) dbg[8 * L + 5] = globaltimer_ns();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

static int g_chain_max_smem = 0;
static int g_chain_cluster = 0;   // YL_CHAIN_CLUSTER: 0 = by batch size, else 1 / 2 / 4 / 8 / 16 (read once in yl_init)

// CTAs per image: at tiny batches an image's layer is spread over up to 16 SMs (measured, model graph replay at bs = 1:
// 0.614 ms with clusters of 4, 0.528 with 8, 0.444 with 16; per-layer launches: 0.343), shrinking to 4 as the batch fills
// the GPU with one CTA per SM (bs = 16: 0.829 ms with clusters of 8, 1.050 with 16)
static int chain_cluster_for(int batch) {
    if (g_chain_cluster == 1 || g_chain_cluster == 2 || g_chain_cluster == 4 || g_chain_cluster == 8 || g_chain_cluster == 16)
        return g_chain_cluster;
    int cs = 16;
    while (cs > 4 && batch * cs > 148) cs >>= 1;
    return cs;
}
static thread_local unsigned long long* g_chain_dbg = nullptr;   // yl_conv_chain_debug: state of the calling thread

int init_conv_chain() {
    int dev = 0;
    YL_CUDA(cudaGetDevice(&dev));
    YL_CUDA(cudaDeviceGetAttribute(&g_chain_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    YL_CUDA(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_chain_max_smem));
    // clusters of 16 CTAs (one image spread over 16 SMs at tiny batch sizes) are beyond the portable limit of 8
    YL_CUDA(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    {
        const char* e = getenv("YL_CHAIN_CLUSTER");
        g_chain_cluster = (e && *e) ? atoi(e) : 0;
    }
    return YL_OK;
}

static bool chain_layer_ok(const yl_conv_args* a, ConvTcParams& p, size_t* smem, int cluster = 0) {
    ConvTcPlanOpts o;
    o.chain = 1;
    o.chain_cluster = cluster > 0 ? cluster : chain_cluster_for(a->x.n);
    int grid = 0;
    if (plan_conv_tc(a, p, &grid, smem, &o) != YL_OK) return false;
    // the chain kernel instantiates the four bf16-store epilogues on 32-column chunks only
    return p.epi_kind >= 0 && p.epi_kind <= 3 && p.TN == 1 && !p.patch && !p.wres && p.ch_na >= 1 && p.ch_na <= kChainMaxA &&
           (p.ch_bres || (p.ch_nb >= 2 && p.ch_nb <= kChainMaxB));
}

}  // namespace yl

extern "C" {

size_t yl_conv_chain_desc_bytes(int n_layers) { return n_layers > 0 ? (size_t)n_layers * sizeof(yl::ConvTcParams) : 0; }

int yl_conv_chain_supported(const yl_conv_args* a) {
    if (!a) return 0;
    yl::ConvTcParams p;
    size_t smem = 0;
    return yl::chain_layer_ok(a, p, &smem) ? 1 : 0;
}

int yl_conv_chain_build(const yl_conv_args* layers, int n_layers, void* desc_dev, size_t desc_bytes, yl_conv_chain* out,
                        void* stream) {
    YL_CHECK(layers && out && desc_dev && n_layers > 0, YL_ERR_ARG, "null pointer / empty chain");
    YL_CHECK(desc_bytes >= yl_conv_chain_desc_bytes(n_layers), YL_ERR_WORKSPACE, "chain descriptor buffer too small");
    YL_CHECK(((uintptr_t)desc_dev & 127) == 0, YL_ERR_ARG, "chain descriptor buffer must be 128-byte aligned");
    std::vector<yl::ConvTcParams> host((size_t)n_layers);
    size_t smem_max = 0;
    uint32_t cols = 32;
    const int batch = layers[0].x.n;
    const int cluster = yl::chain_cluster_for(batch);
    for (int i = 0; i < n_layers; ++i) {
        size_t smem = 0;
        YL_CHECK(layers[i].x.n == batch, YL_ERR_ARG, "chain layers must share the batch size");
        if (!yl::chain_layer_ok(&layers[i], host[i], &smem, cluster)) {
            yl::set_error("layer %d of the chain cannot run in conv_chain_kernel (%d->%d k%d)", i, layers[i].x.c,
                          layers[i].y.c, layers[i].k);
            return YL_ERR_UNSUPPORTED;
        }
        if (smem > smem_max) smem_max = smem;
        if (host[i].tmem_cols > cols) cols = host[i].tmem_cols;
    }
    const size_t smem_total = smem_max + yl::kChainHeaderBudget;
    YL_CHECK((int)smem_total <= yl::g_chain_max_smem && smem_total <= 114 * 1024, YL_ERR_UNSUPPORTED,
             "chain needs %zu B of shared memory per CTA", smem_total);
    cudaStream_t s = (cudaStream_t)stream;
    // the host vector dies with this call: the copy must have completed before we return
    YL_CUDA(cudaMemcpyAsync(desc_dev, host.data(), (size_t)n_layers * sizeof(yl::ConvTcParams), cudaMemcpyHostToDevice, s));
    YL_CUDA(cudaStreamSynchronize(s));
    out->desc = desc_dev;
    out->n_layers = n_layers;
    out->batch = batch;
    out->cluster = cluster;
    out->smem_bytes = (int32_t)smem_total;
    out->tmem_cols = (int32_t)cols;
    out->reserved = 0;
    return YL_OK;
}

int yl_conv_chain_debug(unsigned long long* device_buf) {
    yl::g_chain_dbg = device_buf;
    return YL_OK;
}

int yl_conv_chain_run(const yl_conv_chain* c, void* stream) {
    YL_CHECK(c && c->desc && c->n_layers > 0 && c->batch > 0 && c->cluster > 0, YL_ERR_ARG, "bad chain");
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(c->batch * c->cluster));
    cfg.blockDim = dim3(yl::kConvTcThreads);
    cfg.dynamicSmemBytes = (size_t)c->smem_bytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)c->cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = yl::pdl_enabled() ? 2 : 1;
    YL_CUDA(cudaLaunchKernelEx(&cfg, yl::conv_chain_kernel, reinterpret_cast<const yl::ConvTcParams*>(c->desc), c->n_layers,
                               (uint32_t)c->tmem_cols, yl::g_chain_dbg));
    YL_LAUNCH_OK("conv_chain_kernel");
    return YL_OK;
}

}  // extern "C"
