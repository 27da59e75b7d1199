// C2PSA attention core (block.py:905-914): per image and head, out_i = sum_j softmax_j(scale * q_i.k_j) v_j.
// Tokens are the H*W pixels; q|k|v of a head are 2*kd+hd contiguous channels of the NHWC qkv tensor, so every
// operand row is a contiguous run (64 B q/k, 128 B v).
//
// Flash-style fused kernel on the tensor cores: one CTA = 64 queries of one (image, head), 4 warps x 16 query
// rows.  Key/value tiles of 64 tokens are staged in padded shared memory (conflict-free ldmatrix); S = Q K^T
// and O += P V run as mma.sync.m16n8k16 bf16 with fp32 accumulators held in registers (S never leaves the
// register file: the accumulator fragment of S is re-packed in place as the A fragment of P), online softmax
// in the log2 domain (one MUFU.EX2 per score), no N x N matrix in HBM.  The whole problem is 61 MFLOP and
// 0.3 MB per image (N = 400 tokens, 2-4 heads), i.e. latency-bound: register-resident mma.sync tiles beat a
// TMEM round trip here, which is why this kernel does not use tcgen05 (the convolutions do).
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

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                 "{%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int KD, int HD>
__global__ void __launch_bounds__(128) psa_attention_kernel(const AttnParams p) {
    constexpr int QT = 64, KT = 64;
    constexpr int KP = KD + 8, VP = HD + 8;  // padded pitches (elements): 16-B rows land on distinct banks
    __shared__ __align__(16) __nv_bfloat16 sQ[QT * KP];
    __shared__ __align__(16) __nv_bfloat16 sK[KT * KP];
    __shared__ __align__(16) __nv_bfloat16 sV[KT * VP];
    griddep_launch_dependents();
    griddep_wait();
    const int head = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * QT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const __nv_bfloat16* base = p.qkv + (long long)b * p.N * p.q_cstride + p.q_coff + head * (2 * KD + HD);

    // ---- Q tile -> smem -> A fragments (2 k-steps of 16)
    for (int e = tid; e < QT * (KD / 8); e += 128) {
        const int r = e / (KD / 8), v = e - r * (KD / 8);
        const int qi = min(q0 + r, p.N - 1);
        *reinterpret_cast<uint4*>(&sQ[r * KP + v * 8]) =
            __ldg(reinterpret_cast<const uint4*>(base + (long long)qi * p.q_cstride) + v);
    }
    __syncthreads();
    uint32_t qa[KD / 16][4];
#pragma unroll
    for (int ks = 0; ks < KD / 16; ++ks) {
        const int row = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int col = ks * 16 + 8 * (lane >> 4);
        ldsm_x4(qa[ks], smem_u32(&sQ[row * KP + col]));
    }

    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // rows g and g+8 of this warp's 16 queries
    const float sl2 = p.scale * 1.4426950408889634f;

    for (int j0 = 0; j0 < p.N; j0 += KT) {
        __syncthreads();  // previous tile fully consumed
        for (int e = tid; e < KT * (KD / 8); e += 128) {
            const int r = e / (KD / 8), v = e - r * (KD / 8);
            const int kj = min(j0 + r, p.N - 1);
            *reinterpret_cast<uint4*>(&sK[r * KP + v * 8]) =
                __ldg(reinterpret_cast<const uint4*>(base + (long long)kj * p.q_cstride + KD) + v);
        }
        for (int e = tid; e < KT * (HD / 8); e += 128) {
            const int r = e / (HD / 8), v = e - r * (HD / 8);
            const int kj = min(j0 + r, p.N - 1);
            *reinterpret_cast<uint4*>(&sV[r * VP + v * 8]) =
                __ldg(reinterpret_cast<const uint4*>(base + (long long)kj * p.q_cstride + 2 * KD) + v);
        }
        __syncthreads();

        // ---- S = Q K^T : 16 queries x 64 keys per warp
        float sc[KT / 8][4];
#pragma unroll
        for (int nt = 0; nt < KT / 8; ++nt) {
            sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
            uint32_t kb[4];  // (k 0-7, 8-15, 16-23, 24-31) of keys nt*8..nt*8+7
            ldsm_x4(kb, smem_u32(&sK[(nt * 8 + (lane & 7)) * KP + 8 * (lane >> 3)]));
            mma_bf16_16816(sc[nt], qa[0], kb[0], kb[1]);
            mma_bf16_16816(sc[nt], qa[1], kb[2], kb[3]);
        }
        // ---- online softmax (log2 domain), rows g (regs 0,1) and g+8 (regs 2,3)
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int nt = 0; nt < KT / 8; ++nt) {
            const int key = j0 + nt * 8 + 2 * t;
            sc[nt][0] = key < p.N ? sc[nt][0] * sl2 : -INFINITY;
            sc[nt][1] = key + 1 < p.N ? sc[nt][1] * sl2 : -INFINITY;
            sc[nt][2] = key < p.N ? sc[nt][2] * sl2 : -INFINITY;
            sc[nt][3] = key + 1 < p.N ? sc[nt][3] * sl2 : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float c0 = exp2f(m0 - mx0), c1 = exp2f(m1 - mx1);  // first tile: exp2(-inf) = 0
        m0 = mx0;
        m1 = mx1;
        l0 *= c0;
        l1 *= c1;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
            o[i][0] *= c0; o[i][1] *= c0;
            o[i][2] *= c1; o[i][3] *= c1;
        }
        uint32_t pa[KT / 16][4];  // P as A fragments: k-step kk covers keys 16kk..16kk+15
#pragma unroll
        for (int nt = 0; nt < KT / 8; ++nt) {
            const float p0 = exp2f(sc[nt][0] - m0), p1 = exp2f(sc[nt][1] - m0);
            const float p2 = exp2f(sc[nt][2] - m1), p3 = exp2f(sc[nt][3] - m1);
            l0 += p0 + p1;
            l1 += p2 + p3;
            pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
            pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
        // ---- O += P V
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
#pragma unroll
            for (int np = 0; np < HD / 16; ++np) {
                uint32_t vb[4];  // b0,b1 of channel tile 2np, b0,b1 of channel tile 2np+1
                const int key = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
                const int ch = np * 16 + 8 * (lane >> 4);
                ldsm_x4_trans(vb, smem_u32(&sV[key * VP + ch]));
                mma_bf16_16816(o[2 * np], pa[kk], vb[0], vb[1]);
                mma_bf16_16816(o[2 * np + 1], pa[kk], vb[2], vb[3]);
            }
        }
    }
    // ---- normalise and store (quad-reduced row sums)
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    __nv_bfloat16* ob = p.out + (long long)b * p.N * p.o_cstride + p.o_coff + head * HD + 2 * t;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
        if (r0 < p.N)
            *reinterpret_cast<uint32_t*>(ob + (long long)r0 * p.o_cstride + i * 8) = pack_bf16x2(o[i][0] * i0, o[i][1] * i0);
        if (r1 < p.N)
            *reinterpret_cast<uint32_t*>(ob + (long long)r1 * p.o_cstride + i * 8) = pack_bf16x2(o[i][2] * i1, o[i][3] * i1);
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
    dim3 grid((unsigned)yl::ceil_div(p.N, 64), (unsigned)heads, (unsigned)qkv->n);
    YL_CUDA(yl::launch_kernel(yl::psa_attention_kernel<32, 64>, grid, dim3(128), 0, (cudaStream_t)stream, p));
    YL_LAUNCH_OK("psa_attention_kernel");
    return YL_OK;
}
