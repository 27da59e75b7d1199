// C2PSA attention core (block.py:905-914): per image and head, out_i = sum_j softmax_j(scale * q_i.k_j) v_j.
// Tokens are the H*W pixels; q|k|v of a head are 2*kd+hd contiguous channels of the NHWC qkv tensor, so every
// operand row is a contiguous run (64 B q/k, 128 B v).
//
// v0 (this file): flash-style streaming softmax on CUDA cores — one thread per query, K/V tiles of 64 keys
// staged in shared memory and read as warp-wide broadcasts, fp32 accumulation, no N x N matrix in HBM.
// The tcgen05 version (S in TMEM) replaces it once the conv path is tuned; attention is <1 % of the FLOPs.
#include "common.cuh"

namespace yl {

struct AttnParams {
    const __nv_bfloat16* qkv;
    long long q_cstride;
    int q_coff;
    __nv_bfloat16* out;
    long long o_cstride;
    int o_coff;
    int N;  // tokens per image
    float scale;
};

template <int KD, int HD>
__global__ void __launch_bounds__(128) psa_attention_kernel(const AttnParams p) {
    constexpr int KT = 64;                     // keys per shared-memory tile
    constexpr int SUB = 16;                    // keys per online-softmax step
    __shared__ __align__(16) uint4 sK[KT * KD / 8];
    __shared__ __align__(16) uint4 sV[KT * HD / 8];
    const int head = blockIdx.y, b = blockIdx.z;
    const int i = blockIdx.x * 128 + threadIdx.x;  // query token
    const int iq = min(i, p.N - 1);
    const int hc = head * (2 * KD + HD);
    const __nv_bfloat16* base = p.qkv + (long long)b * p.N * p.q_cstride + p.q_coff + hc;

    float q[KD];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(base + (long long)iq * p.q_cstride);
#pragma unroll
        for (int v = 0; v < KD / 8; ++v) {
            const uint4 t = __ldg(qp + v);
            q[v * 8 + 0] = bf16lo_f(t.x) * p.scale; q[v * 8 + 1] = bf16hi_f(t.x) * p.scale;
            q[v * 8 + 2] = bf16lo_f(t.y) * p.scale; q[v * 8 + 3] = bf16hi_f(t.y) * p.scale;
            q[v * 8 + 4] = bf16lo_f(t.z) * p.scale; q[v * 8 + 5] = bf16hi_f(t.z) * p.scale;
            q[v * 8 + 6] = bf16lo_f(t.w) * p.scale; q[v * 8 + 7] = bf16hi_f(t.w) * p.scale;
        }
    }
    float o[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) o[c] = 0.f;
    float m = -INFINITY, l = 0.f;

    for (int j0 = 0; j0 < p.N; j0 += KT) {
        __syncthreads();
        for (int e = threadIdx.x; e < KT * (KD / 8); e += 128) {
            const int j = e / (KD / 8), v = e - j * (KD / 8);
            const int jj = min(j0 + j, p.N - 1);
            sK[e] = __ldg(reinterpret_cast<const uint4*>(base + (long long)jj * p.q_cstride + KD) + v);
        }
        for (int e = threadIdx.x; e < KT * (HD / 8); e += 128) {
            const int j = e / (HD / 8), v = e - j * (HD / 8);
            const int jj = min(j0 + j, p.N - 1);
            sV[e] = __ldg(reinterpret_cast<const uint4*>(base + (long long)jj * p.q_cstride + 2 * KD) + v);
        }
        __syncthreads();
        const int kmax = min(KT, p.N - j0);
        for (int s0 = 0; s0 < kmax; s0 += SUB) {
            float s[SUB];
            float mx = m;
#pragma unroll
            for (int jj = 0; jj < SUB; ++jj) {
                float acc = 0.f;
#pragma unroll
                for (int v = 0; v < KD / 8; ++v) {
                    const uint4 t = sK[(s0 + jj) * (KD / 8) + v];
                    acc = fmaf(q[v * 8 + 0], bf16lo_f(t.x), acc); acc = fmaf(q[v * 8 + 1], bf16hi_f(t.x), acc);
                    acc = fmaf(q[v * 8 + 2], bf16lo_f(t.y), acc); acc = fmaf(q[v * 8 + 3], bf16hi_f(t.y), acc);
                    acc = fmaf(q[v * 8 + 4], bf16lo_f(t.z), acc); acc = fmaf(q[v * 8 + 5], bf16hi_f(t.z), acc);
                    acc = fmaf(q[v * 8 + 6], bf16lo_f(t.w), acc); acc = fmaf(q[v * 8 + 7], bf16hi_f(t.w), acc);
                }
                s[jj] = (s0 + jj < kmax) ? acc : -INFINITY;
                mx = fmaxf(mx, s[jj]);
            }
            const float corr = __expf(m - mx);  // m = -inf on the first step -> 0
            l *= corr;
#pragma unroll
            for (int c = 0; c < HD; ++c) o[c] *= corr;
            m = mx;
#pragma unroll
            for (int jj = 0; jj < SUB; ++jj) {
                const float pj = __expf(s[jj] - m);
                l += pj;
#pragma unroll
                for (int v = 0; v < HD / 8; ++v) {
                    const uint4 t = sV[(s0 + jj) * (HD / 8) + v];
                    o[v * 8 + 0] = fmaf(pj, bf16lo_f(t.x), o[v * 8 + 0]); o[v * 8 + 1] = fmaf(pj, bf16hi_f(t.x), o[v * 8 + 1]);
                    o[v * 8 + 2] = fmaf(pj, bf16lo_f(t.y), o[v * 8 + 2]); o[v * 8 + 3] = fmaf(pj, bf16hi_f(t.y), o[v * 8 + 3]);
                    o[v * 8 + 4] = fmaf(pj, bf16lo_f(t.z), o[v * 8 + 4]); o[v * 8 + 5] = fmaf(pj, bf16hi_f(t.z), o[v * 8 + 5]);
                    o[v * 8 + 6] = fmaf(pj, bf16lo_f(t.w), o[v * 8 + 6]); o[v * 8 + 7] = fmaf(pj, bf16hi_f(t.w), o[v * 8 + 7]);
                }
            }
        }
    }
    if (i < p.N) {
        const float inv = 1.f / l;
        uint4* op = reinterpret_cast<uint4*>(p.out + ((long long)b * p.N + i) * p.o_cstride + p.o_coff + head * HD);
#pragma unroll
        for (int v = 0; v < HD / 8; ++v) {
            uint4 t;
            t.x = pack_bf16x2(o[v * 8 + 0] * inv, o[v * 8 + 1] * inv);
            t.y = pack_bf16x2(o[v * 8 + 2] * inv, o[v * 8 + 3] * inv);
            t.z = pack_bf16x2(o[v * 8 + 4] * inv, o[v * 8 + 5] * inv);
            t.w = pack_bf16x2(o[v * 8 + 6] * inv, o[v * 8 + 7] * inv);
            op[v] = t;
        }
    }
}

int init_attention() { return YL_OK; }

}  // namespace yl

extern "C" int yl_psa_attention(const yl_tensor* qkv, const yl_tensor* out, int heads, int key_dim, int head_dim,
                                float scale, void* stream) {
    YL_CHECK(qkv && out && qkv->data && out->data, YL_ERR_ARG, "null pointer");
    YL_CHECK(qkv->dtype == YL_BF16 && out->dtype == YL_BF16, YL_ERR_ARG, "attention tensors must be bf16");
    YL_CHECK(heads >= 1 && qkv->c == heads * (2 * key_dim + head_dim) && out->c == heads * head_dim, YL_ERR_ARG,
             "attention channel layout mismatch");
    YL_CHECK(qkv->n == out->n && qkv->h == out->h && qkv->w == out->w, YL_ERR_ARG, "attention shape mismatch");
    YL_CHECK(qkv->coff % 8 == 0 && qkv->cstride % 8 == 0 && out->coff % 8 == 0 && out->cstride % 8 == 0, YL_ERR_ARG,
             "attention needs 8-channel alignment");
    YL_CHECK(key_dim == 32 && head_dim == 64, YL_ERR_UNSUPPORTED,
             "attention is built for key_dim=32, head_dim=64 (every yolo11 scale); got %d/%d", key_dim, head_dim);
    yl::AttnParams p;
    p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv->data);
    p.q_cstride = qkv->cstride;
    p.q_coff = qkv->coff;
    p.out = reinterpret_cast<__nv_bfloat16*>(out->data);
    p.o_cstride = out->cstride;
    p.o_coff = out->coff;
    p.N = qkv->h * qkv->w;
    p.scale = scale;
    dim3 grid((unsigned)yl::ceil_div(p.N, 128), (unsigned)heads, (unsigned)qkv->n);
    yl::psa_attention_kernel<32, 64><<<grid, 128, 0, (cudaStream_t)stream>>>(p);
    YL_LAUNCH_OK("psa_attention_kernel");
    return YL_OK;
}
