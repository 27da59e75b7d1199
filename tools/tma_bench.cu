// TMA throughput microbenchmark (sm_100a): how fast can one SM's TMA unit move boxes of `rows` x `row_bytes`
// between global memory and shared memory, as a function of the row width, the direction and the number of
// resident CTAs?  The conv kernel's tile/box geometry is chosen from these numbers.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tma_bench.bin tools/tma_bench.cu
//   ./tools/tma_bench.bin            (prints one line per configuration)
//
// Kernel: persistent CTAs; thread 0 = producer (TMA loads into a D-deep ring), thread 32 = consumer (waits for a
// stage, optionally TMA-stores it to the output tensor, releases the stage).  No math: pure data movement.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (++spins > (1u << 26)) __trap();
    }
}

struct Params {
    CUtensorMap in, out;
    int tiles;        // total boxes
    int rows_per_box; // along dim 1
    int depth;
    uint32_t box_bytes;   // stage stride (1024-aligned)
    uint32_t tx_bytes;    // bytes one box delivers
    int do_load, do_store;
    int rank4;        // in/out maps are 4-D {c, w, h, n} with box {c, tw, th, 1}: coords derived from the tile
    int tw, th, W, H;
};

__global__ void __launch_bounds__(64) tma_kernel(const __grid_constant__ Params p) {
    extern __shared__ uint8_t raw[];
    const uint32_t a = smem_u32(raw);
    uint8_t* base = raw + (((a + 1023u) & ~1023u) - a);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)p.depth * p.box_bytes);
    uint64_t* empty = full + p.depth;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.depth; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tiles_w = p.rank4 ? (p.W / p.tw) : 1, tiles_h = p.rank4 ? (p.H / p.th) : 1;
    if (threadIdx.x == 0 && p.do_load) {
        int it = 0;
        for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
            const int s = it % p.depth;
            mbar_wait(&empty[s], ((it / p.depth) & 1) ^ 1);
            mbar_expect_tx(&full[s], p.tx_bytes);
            if (p.rank4) {
                const int w0 = (t % tiles_w) * p.tw, h0 = ((t / tiles_w) % tiles_h) * p.th, n0 = t / (tiles_w * tiles_h);
                asm volatile(
                    "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                    ::"r"(smem_u32(base + (size_t)s * p.box_bytes)), "l"(reinterpret_cast<uint64_t>(&p.in)),
                    "r"(smem_u32(&full[s])), "r"(0), "r"(w0), "r"(h0), "r"(n0) : "memory");
            } else {
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                    ::"r"(smem_u32(base + (size_t)s * p.box_bytes)), "l"(reinterpret_cast<uint64_t>(&p.in)),
                    "r"(smem_u32(&full[s])), "r"(0), "r"(t * p.rows_per_box) : "memory");
            }
        }
    } else if (threadIdx.x == 32) {
        int it = 0;
        for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
            const int s = it % p.depth;
            if (p.do_load) mbar_wait(&full[s], (it / p.depth) & 1);
            if (p.do_store) {
                if (!p.do_load && it >= p.depth) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(3) : "memory");
                if (p.rank4) {
                    const int w0 = (t % tiles_w) * p.tw, h0 = ((t / tiles_w) % tiles_h) * p.th, n0 = t / (tiles_w * tiles_h);
                    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(&p.out)), "r"(smem_u32(base + (size_t)s * p.box_bytes)),
                                 "r"(0), "r"(w0), "r"(h0), "r"(n0) : "memory");
                } else {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(&p.out)), "r"(smem_u32(base + (size_t)s * p.box_bytes)),
                                 "r"(0), "r"(t * p.rows_per_box) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (p.do_load) {
                    // the stage can be refilled once the store has read it: keep one store in flight
                    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(1) : "memory");
                    if (it > 0) mbar_arrive(&empty[(it - 1) % p.depth]);
                }
            } else {
                mbar_arrive(&empty[s]);
            }
        }
        if (p.do_store) {
            asm volatile("cp.async.bulk.wait_group %0;" ::"n"(0) : "memory");
        }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn g_enc;

static CUtensorMapSwizzle swz(int rb) {
    return rb >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                       : rb == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
}

// flat: tensor {c (inner, bf16), M rows} with pitch `cstride` elements; box {cb, rows}
static void make2d(CUtensorMap* m, void* base, int c, long long M, int cstride, int cb, int rows, int promo) {
    cuuint64_t dims[2] = {(cuuint64_t)c, (cuuint64_t)M};
    cuuint64_t str[1] = {(cuuint64_t)cstride * 2};
    cuuint32_t box[2] = {(cuuint32_t)cb, (cuuint32_t)rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swz(cb * 2), (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode2d failed %d\n", (int)r); exit(1); }
}
static void make4d(CUtensorMap* m, void* base, int c, int W, int H, int N, int cstride, int cb, int tw, int th, int promo) {
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t str[3] = {(cuuint64_t)cstride * 2, (cuuint64_t)cstride * 2 * W, (cuuint64_t)cstride * 2 * W * H};
    cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)tw, (cuuint32_t)th, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swz(cb * 2), (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode4d failed %d\n", (int)r); exit(1); }
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    g_enc = (EncodeFn)fn;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

    const long long M = 64ll * 160 * 160;  // pixels (yolo11n bs=64 at 160x160)
    const size_t bytes = (size_t)M * 128 * 2;
    void *din, *dout;
    CK(cudaMalloc(&din, bytes));
    CK(cudaMalloc(&dout, bytes));
    CK(cudaMemset(din, 1, bytes));
    CK(cudaMemset(dout, 0, bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    printf("%-44s %8s %9s %9s %10s\n", "config", "us", "GB/s", "B/clk/SM", "clk/row/SM");
    struct Cfg { const char* name; int cb; int cstride; int rows; int load, store; int ctas; int depth; int promo; int rank4; int tw, th; };
    const Cfg cfgs[] = {
        {"load  128B rows x128, 1 CTA/SM d4", 64, 64, 128, 1, 0, 1, 4, 2, 0, 0, 0},
        {"load  128B rows x128, 2 CTA/SM d4", 64, 64, 128, 1, 0, 2, 4, 2, 0, 0, 0},
        {"load  128B rows x128, 2 CTA/SM d8", 64, 64, 128, 1, 0, 2, 8, 2, 0, 0, 0},
        {"load  128B rows x128, 4 CTA/SM d4", 64, 64, 128, 1, 0, 4, 4, 2, 0, 0, 0},
        {"load  128B rows x256, 2 CTA/SM d4", 64, 64, 256, 1, 0, 2, 4, 2, 0, 0, 0},
        {"load  128B rows x128 promo128 2CTA d4", 64, 64, 128, 1, 0, 2, 4, 1, 0, 0, 0},
        {"load  128B rows x128 promoNONE 2CTA d4", 64, 64, 128, 1, 0, 2, 4, 0, 0, 0, 0},
        {"load  128B of 256B pitch, 2 CTA/SM d4", 64, 128, 128, 1, 0, 2, 4, 2, 0, 0, 0},
        {"load   64B rows x128, 2 CTA/SM d8", 32, 32, 128, 1, 0, 2, 8, 2, 0, 0, 0},
        {"load   64B of 96B pitch, 2 CTA/SM d8", 32, 48, 128, 1, 0, 2, 8, 2, 0, 0, 0},
        {"load   32B rows x128, 2 CTA/SM d8", 16, 16, 128, 1, 0, 2, 8, 2, 0, 0, 0},
        {"load   32B of 96B pitch, 2 CTA/SM d8", 16, 48, 128, 1, 0, 2, 8, 2, 0, 0, 0},
        {"load   16B rows x128, 2 CTA/SM d8", 8, 8, 128, 1, 0, 2, 8, 2, 0, 0, 0},
        {"load  4D 128B {64,8,16}, 2 CTA/SM d4", 64, 64, 128, 1, 0, 2, 4, 2, 1, 8, 16},
        {"load  4D 128B {64,16,8}, 2 CTA/SM d4", 64, 64, 128, 1, 0, 2, 4, 2, 1, 16, 8},
        {"load  4D 128B {64,32,4}, 2 CTA/SM d4", 64, 64, 128, 1, 0, 2, 4, 2, 1, 32, 4},
        {"load  4D  32B {16,8,16}, 2 CTA/SM d8", 16, 16, 128, 1, 0, 2, 8, 2, 1, 8, 16},
        {"load  4D  32B {16,10,18} halo, 2 CTA d8", 16, 16, 180, 1, 0, 2, 8, 2, 1, 10, 18},
        {"store 128B rows x128, 1 CTA/SM", 64, 64, 128, 0, 1, 1, 4, 2, 0, 0, 0},
        {"store 128B rows x128, 2 CTA/SM", 64, 64, 128, 0, 1, 2, 4, 2, 0, 0, 0},
        {"store 128B rows x128, 4 CTA/SM", 64, 64, 128, 0, 1, 4, 4, 2, 0, 0, 0},
        {"store  64B rows x128, 2 CTA/SM", 32, 32, 128, 0, 1, 2, 4, 2, 0, 0, 0},
        {"store  64B of 128B pitch, 2 CTA/SM", 32, 64, 128, 0, 1, 2, 4, 2, 0, 0, 0},
        {"store  32B rows x128, 2 CTA/SM", 16, 16, 128, 0, 1, 2, 4, 2, 0, 0, 0},
        {"store  32B of 96B pitch, 2 CTA/SM", 16, 48, 128, 0, 1, 2, 4, 2, 0, 0, 0},
        {"store 4D 128B {64,8,16}, 2 CTA/SM", 64, 64, 128, 0, 1, 2, 4, 2, 1, 8, 16},
        {"copy  128B rows x128, 2 CTA/SM d4", 64, 64, 128, 1, 1, 2, 4, 2, 0, 0, 0},
        {"copy  128B rows x128, 1 CTA/SM d8", 64, 64, 128, 1, 1, 1, 8, 2, 0, 0, 0},
        {"copy   64B rows x128, 2 CTA/SM d8", 32, 32, 128, 1, 1, 2, 8, 2, 0, 0, 0},
        {"copy   32B rows x128, 2 CTA/SM d8", 16, 16, 128, 1, 1, 2, 8, 2, 0, 0, 0},
    };
    for (const Cfg& c : cfgs) {
        Params p;
        memset(&p, 0, sizeof(p));
        const int W = 160, H = 160, N = 64;
        long long boxes;
        if (c.rank4) {
            // halo config reads overlapping boxes; count tiles on the (tw-2, th-2) interior grid instead
            const int stepw = (c.tw == 10) ? 8 : c.tw, steph = (c.th == 18) ? 16 : c.th;
            make4d(&p.in, din, c.cb, W, H, N, c.cstride, c.cb, c.tw, c.th, c.promo);
            make4d(&p.out, dout, c.cb, W, H, N, c.cstride, c.cb, c.tw, c.th, c.promo);
            p.rank4 = 1;
            p.tw = stepw;
            p.th = steph;
            p.W = W;
            p.H = H;
            boxes = (long long)(W / stepw) * (H / steph) * N;
        } else {
            make2d(&p.in, din, c.cb, M, c.cstride, c.cb, c.rows, c.promo);
            make2d(&p.out, dout, c.cb, M, c.cstride, c.cb, c.rows, c.promo);
            boxes = M / c.rows;
        }
        p.tiles = (int)boxes;
        p.rows_per_box = c.rows;
        p.depth = c.depth;
        p.box_bytes = (uint32_t)(((size_t)c.rows * c.cb * 2 + 1023) & ~1023ull);
        p.tx_bytes = (uint32_t)((size_t)c.rows * c.cb * 2);
        p.do_load = c.load;
        p.do_store = c.store;
        const size_t smem = 1024 + (size_t)p.depth * p.box_bytes + 2 * p.depth * 8 + 64;
        if (smem * c.ctas > 220 * 1024) { printf("%-44s skipped (smem)\n", c.name); continue; }
        const int grid = sms * c.ctas;
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            tma_kernel<<<grid, 64, smem>>>(p);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        const double moved = (double)boxes * c.rows * c.cb * 2 * (c.load + c.store);
        const double rows = (double)boxes * c.rows * (c.load + c.store);
        const double clk = best * 1e-3 * clk_khz * 1e3;
        printf("%-44s %8.1f %9.0f %9.1f %10.2f\n", c.name, best * 1e3, moved / best / 1e6, moved / clk / sms,
               clk * sms / rows);
    }
    return 0;
}
