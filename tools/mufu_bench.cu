// Throughput of the special-function forms a SiLU epilogue can be built from (sm_100a), in elements / clk / SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mufu_bench.bin tools/mufu_bench.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int MODE>
__device__ __forceinline__ float silu_variant(float x) {
    if (MODE == 0) {  // h + h * tanh.approx.f32(h)
        float h = 0.5f * x, t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
        return fmaf(h, t, h);
    } else if (MODE == 1) {  // x * rcp(1 + ex2(-x * log2e))
        float e, r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
        return x * r;
    } else if (MODE == 4) {  // FMA-only: sigmoid via exp2 polynomial + Newton reciprocal
        float z = fminf(fmaxf(-1.4426950408889634f * x, -30.f), 30.f);
        float fl = floorf(z), f = z - fl;
        float p = fmaf(fmaf(fmaf(0.0790209f, f, 0.2241240f), f, 0.6968040f), f, 1.0f);  // 2^f, f in [0,1)
        float e = __int_as_float(__float_as_int(p) + ((int)fl << 23));
        float d = 1.0f + e;
        float r = __int_as_float(0x7EF311C7 - __float_as_int(d));  // bit-trick reciprocal seed
        r = r * fmaf(-d, r, 2.0f);
        r = r * fmaf(-d, r, 2.0f);
        r = r * fmaf(-d, r, 2.0f);
        return x * r;
    }
    return x;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 0.001f + i * 0.37f - 1.f;
    if (MODE == 2 || MODE == 3 || MODE == 5) {
        // packed: 2 elements per MUFU op
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float h0 = 0.5f * v[i], h1 = 0.5f * v[i + 1];
                uint32_t hp, tp;
                if (MODE == 2) {
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hp) : "f"(h1), "f"(h0));
                    asm("tanh.approx.f16x2 %0, %1;" : "=r"(tp) : "r"(hp));
                    __half2 t2 = *reinterpret_cast<__half2*>(&tp);
                    v[i] = fmaf(h0, __low2float(t2), h0);
                    v[i + 1] = fmaf(h1, __high2float(t2), h1);
                } else if (MODE == 3) {
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hp) : "f"(h1), "f"(h0));
                    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(tp) : "r"(hp));
                    v[i] = fmaf(h0, __uint_as_float(tp << 16), h0);
                    v[i + 1] = fmaf(h1, __uint_as_float(tp & 0xffff0000u), h1);
                } else {  // ex2.f16x2 then f32 rcp per element
                    float a0 = -1.4426950408889634f * v[i], a1 = -1.4426950408889634f * v[i + 1];
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hp) : "f"(a1), "f"(a0));
                    asm("ex2.approx.f16x2 %0, %1;" : "=r"(tp) : "r"(hp));
                    __half2 t2 = *reinterpret_cast<__half2*>(&tp);
                    float r0, r1;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(1.0f + __low2float(t2)));
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(1.0f + __high2float(t2)));
                    v[i] *= r0;
                    v[i + 1] *= r1;
                }
            }
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = silu_variant<MODE>(v[i]) + 0.25f;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* out, int sms, int clk_khz) {
    const int iters = 4096, blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 16);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double elems = (double)blocks * 256 * 8 * iters;
    const double clk = best * 1e-3 * clk_khz * 1e3;
    printf("%-44s %8.3f ms  %6.2f elements/clk/SM\n", name, best, elems / clk / sms);
}

int main() {
    int sms, clk;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    run<0>("silu: h + h*tanh.approx.f32(h)", out, sms, clk);
    run<1>("silu: x*rcp(1+ex2(-x))  (2 MUFU f32)", out, sms, clk);
    run<2>("silu: tanh.approx.f16x2 (1 MUFU / 2 elems)", out, sms, clk);
    run<3>("silu: tanh.approx.bf16x2 (1 MUFU / 2 elems)", out, sms, clk);
    run<5>("silu: ex2.f16x2 + 2 rcp.f32", out, sms, clk);
    run<4>("silu: FMA-only exp2 poly + Newton rcp", out, sms, clk);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
