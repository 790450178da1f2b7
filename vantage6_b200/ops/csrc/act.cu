// Bias / activation backward of a fused linear layer  y = act(x W^T + b)  (bf16 activations):
//
//     g = dy * act'(pre)            (act = GELU(erf) / ReLU; identity when the layer has no activation)
//     dpre = g  (bf16)              written only when there is an activation (otherwise dy IS dpre)
//     db[c] (+)= sum_r g[r, c]      fp32, optionally accumulated straight into the flat gradient buffer
//
// ONE pass over dy (and pre) instead of the ~10 ATen launches the composed PyTorch expression costs per layer
// (float casts, erf, exp, muls, adds, a column reduction and the AccumulateGrad add -- 27% + 13% + 7% + 6% of a
// BERT-base step in profiles/launches_bert_base_r1c.txt).  Column sums use the deterministic in-kernel tree of
// tree_reduce.cuh (64-column slices x row-strided CTAs, one wave, no floating-point atomics).
#include "common.cuh"
#include "api.h"
#include "tree_reduce.cuh"

namespace act {

using namespace tree;

V6_DEVINL uint4 ldg_v4(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
V6_DEVINL void unpack8(const uint4& t, float (&v)[8]) {
    float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
V6_DEVINL uint4 pack8(const float (&v)[8]) {
    return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
template <int ACT>
V6_DEVINL float act_grad(float x) {
    if (ACT == 1) return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
    if (ACT == 2) return x > 0.f ? 1.f : 0.f;
    return 1.f;
}

template <int ACT>
__global__ void __launch_bounds__(THREADS, 3) bias_act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ pre,
                                                                  __nv_bfloat16* __restrict__ dpre, Red rd, float* __restrict__ db,
                                                                  long long R, int C, int accumulate) {
    __shared__ float tot[128];
    constexpr int SW = 64, CGS = 8, RL = THREADS / CGS;
    const int cg = threadIdx.x % CGS, rl = threadIdx.x / CGS;
    const size_t c0 = (size_t)blockIdx.y * SW + cg * 8;
    float s[8], z[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] = 0.f; z[k] = 0.f; }
    const long long G = (long long)gridDim.x * RL;
    long long r = (long long)blockIdx.x * RL + rl;
    for (; r + G < R; r += 2 * G) {                               // 2 rows x (dy [+ pre]) in flight
        uint4 gr[2], pr[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long o = (r + u * G) * C + c0;
            gr[u] = ldg_v4(dy + o);
            if (ACT) pr[u] = ldg_v4(pre + o);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float g[8], p[8];
            unpack8(gr[u], g);
            if (ACT) {
                unpack8(pr[u], p);
#pragma unroll
                for (int k = 0; k < 8; ++k) g[k] *= act_grad<ACT>(p[k]);
                *reinterpret_cast<uint4*>(dpre + (r + u * G) * C + c0) = pack8(g);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] += g[k];
        }
    }
    for (; r < R; r += G) {
        const long long o = r * C + c0;
        float g[8], p[8];
        unpack8(ldg_v4(dy + o), g);
        if (ACT) {
            unpack8(ldg_v4(pre + o), p);
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] *= act_grad<ACT>(p[k]);
            *reinterpret_cast<uint4*>(dpre + o) = pack8(g);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] += g[k];
    }
    if (db == nullptr) return;                                     // activation-only (layer without bias)
    if (!slice_reduce(s, z, rd, SW, tot)) return;
    if (threadIdx.x < SW) {
        const int c = blockIdx.y * SW + threadIdx.x;
        db[c] = accumulate ? db[c] + tot[threadIdx.x] : tot[threadIdx.x];
    }
}

// ------------------------------------------------------------------------------------------ SwiGLU
// h = silu(g) * u and its backward (dg = dh * u * silu'(g), du = dh * silu(g)), 16 B vectors, one pass each:
// the composed PyTorch expression is 2 launches forward and ~5 backward over [tokens, ffn] tensors (Llama: 29 MB each).
V6_DEVINL float sigmoidf_fast(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void __launch_bounds__(THREADS) swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ u,
                                                             __nv_bfloat16* __restrict__ h, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float gv[8], uv[8], o[8];
        unpack8(ldg_v4(g + i * 8), gv);
        unpack8(ldg_v4(u + i * 8), uv);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = gv[k] * sigmoidf_fast(gv[k]) * uv[k];
        *reinterpret_cast<uint4*>(h + i * 8) = pack8(o);
    }
}

__global__ void __launch_bounds__(THREADS) swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ g,
                                                             const __nv_bfloat16* __restrict__ u, __nv_bfloat16* __restrict__ dg,
                                                             __nv_bfloat16* __restrict__ du, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float d[8], gv[8], uv[8], og[8], ou[8];
        unpack8(ldg_v4(dh + i * 8), d);
        unpack8(ldg_v4(g + i * 8), gv);
        unpack8(ldg_v4(u + i * 8), uv);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float sg = sigmoidf_fast(gv[k]);
            const float silu = gv[k] * sg;
            og[k] = d[k] * uv[k] * (sg + silu * (1.f - sg));          // silu'(g) = s + g s (1 - s)
            ou[k] = d[k] * silu;
        }
        *reinterpret_cast<uint4*>(dg + i * 8) = pack8(og);
        *reinterpret_cast<uint4*>(du + i * 8) = pack8(ou);
    }
}

static inline int ew_grid(long long n8) {
    long long b = (n8 + THREADS - 1) / THREADS;
    return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

}  // namespace act

// g, u, h (and dh, dg, du): dense bf16 tensors of n elements, n % 8 == 0, 16-byte aligned
extern "C" int v6_swiglu_fwd(const void* g, const void* u, void* h, long long n, cudaStream_t s) {
    using namespace act;
    if (n % 8 != 0 || n < 8) return (int)cudaErrorInvalidValue;
    swiglu_fwd_kernel<<<ew_grid(n / 8), THREADS, 0, s>>>((const __nv_bfloat16*)g, (const __nv_bfloat16*)u, (__nv_bfloat16*)h, n / 8);
    V6_CHECK_LAUNCH();
    return 0;
}
extern "C" int v6_swiglu_bwd(const void* dh, const void* g, const void* u, void* dg, void* du, long long n, cudaStream_t s) {
    using namespace act;
    if (n % 8 != 0 || n < 8) return (int)cudaErrorInvalidValue;
    swiglu_bwd_kernel<<<ew_grid(n / 8), THREADS, 0, s>>>((const __nv_bfloat16*)dh, (const __nv_bfloat16*)g, (const __nv_bfloat16*)u,
                                                        (__nv_bfloat16*)dg, (__nv_bfloat16*)du, n / 8);
    V6_CHECK_LAUNCH();
    return 0;
}

// dy, pre, dpre: [R, C] bf16 row-major (pre / dpre may be null when act == 0); db: [C] fp32 or null; C % 64 == 0.
// scratch: the tree-reduction scratch buffer (v6_bn_scratch_floats() floats, zero-initialised once).
extern "C" int v6_bias_act_bwd(const void* dy, const void* pre, void* dpre, float* db, float* scratch, long long R, int C, int act_kind,
                               int accumulate, cudaStream_t s) {
    using namespace act;
    if (C % 64 != 0 || C > 4096 || R < 1 || act_kind < 0 || act_kind > 2) return (int)cudaErrorInvalidValue;
    if (act_kind != 0 && (!pre || !dpre)) return (int)cudaErrorInvalidValue;
    if (act_kind == 0 && !db) return 0;
    static int wave[3] = {0, 0, 0};
    const __nv_bfloat16* d = (const __nv_bfloat16*)dy;
    const __nv_bfloat16* p = (const __nv_bfloat16*)pre;
    __nv_bfloat16* o = (__nv_bfloat16*)dpre;
    const Red rd = make_red(scratch);
    if (act_kind == 0) bias_act_bwd_kernel<0><<<reduce_grid(R, C, wave_ctas(bias_act_bwd_kernel<0>, wave[0])), THREADS, 0, s>>>(d, p, o, rd, db, R, C, accumulate);
    else if (act_kind == 1) bias_act_bwd_kernel<1><<<reduce_grid(R, C, wave_ctas(bias_act_bwd_kernel<1>, wave[1])), THREADS, 0, s>>>(d, p, o, rd, db, R, C, accumulate);
    else bias_act_bwd_kernel<2><<<reduce_grid(R, C, wave_ctas(bias_act_bwd_kernel<2>, wave[2])), THREADS, 0, s>>>(d, p, o, rd, db, R, C, accumulate);
    V6_CHECK_LAUNCH();
    return 0;
}
