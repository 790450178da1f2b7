// K6 (RoPE) and K8 (logistic-regression GLM step) of SURVEY.md 2.6.
#include "common.cuh"
#include "api.h"

// ----------------------------------------------------------------------------------------
// K6: rotary position embedding, rotate-half convention (Llama): for i < D/2
//   out[i]       = x[i]*cos[s,i] - x[i+D/2]*sin[s,i]
//   out[i+D/2]   = x[i+D/2]*cos[s,i] + x[i]*sin[s,i]
// q:[B,S,Hq,D] and k:[B,S,Hkv,D] (bf16) are rotated IN PLACE by one launch; the backward
// pass is the same kernel with sin negated (`inverse`).  cos/sin tables are fp32 [S, D/2].
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rope_kernel(__nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ k,
                                                   const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                                   const int* __restrict__ pos_ids, int B, int S, int Hq, int Hkv,
                                                   int D, float sign) {
    const int half = D / 2, vec_per_head = half / 8;
    const long long total = (long long)B * S * (Hq + Hkv) * vec_per_head;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int v = idx % vec_per_head;
        long long r = idx / vec_per_head;
        const int h = r % (Hq + Hkv);
        r /= (Hq + Hkv);
        const int s = r % S;
        const int b = r / S;
        const int pos = pos_ids ? pos_ids[b * S + s] : s;
        __nv_bfloat16* base = (h < Hq) ? q + (((size_t)b * S + s) * Hq + h) * D
                                       : k + (((size_t)b * S + s) * Hkv + (h - Hq)) * D;
        uint4 u1 = *reinterpret_cast<const uint4*>(base + v * 8);
        uint4 u2 = *reinterpret_cast<const uint4*>(base + half + v * 8);
        const float4 c0 = *reinterpret_cast<const float4*>(cos_t + (size_t)pos * half + v * 8);
        const float4 c1 = *reinterpret_cast<const float4*>(cos_t + (size_t)pos * half + v * 8 + 4);
        const float4 s0 = *reinterpret_cast<const float4*>(sin_t + (size_t)pos * half + v * 8);
        const float4 s1 = *reinterpret_cast<const float4*>(sin_t + (size_t)pos * half + v * 8 + 4);
        const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const float sn[8] = {s0.x * sign, s0.y * sign, s0.z * sign, s0.w * sign, s1.x * sign, s1.y * sign, s1.z * sign, s1.w * sign};
        const uint32_t a1[4] = {u1.x, u1.y, u1.z, u1.w}, a2[4] = {u2.x, u2.y, u2.z, u2.w};
        uint32_t o1[4], o2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 x1 = unpack_bf16x2(a1[j]), x2 = unpack_bf16x2(a2[j]);
            o1[j] = pack_bf16x2(x1.x * c[2 * j] - x2.x * sn[2 * j], x1.y * c[2 * j + 1] - x2.y * sn[2 * j + 1]);
            o2[j] = pack_bf16x2(x2.x * c[2 * j] + x1.x * sn[2 * j], x2.y * c[2 * j + 1] + x1.y * sn[2 * j + 1]);
        }
        *reinterpret_cast<uint4*>(base + v * 8) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
        *reinterpret_cast<uint4*>(base + half + v * 8) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
    }
}
extern "C" int v6_rope(void* q, void* k, const float* cos_t, const float* sin_t, const int* pos_ids, int B, int S,
                       int Hq, int Hkv, int D, int inverse, cudaStream_t st) {
    if (D % 16 != 0) return (int)cudaErrorInvalidValue;
    const long long total = (long long)B * S * (Hq + Hkv) * (D / 16);
    long long grid = (total + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    rope_kernel<<<(int)grid, 256, 0, st>>>((__nv_bfloat16*)q, (__nv_bfloat16*)k, cos_t, sin_t, pos_ids, B, S, Hq, Hkv, D,
                                          inverse ? -1.f : 1.f);
    V6_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------
// K8: logistic-regression gradient in ONE pass over X (memory-bound, X read exactly once):
//   z = X w + b ; p = sigmoid(z) ; r = p - y ; g_w = X^T r ; g_b = sum r ; loss = sum BCE
// X:[rows, F] bf16 or fp32 row-major, F == 256 (BASELINE config 5) or any multiple of 256.
// One warp per row: each lane owns 8 consecutive features per 256-chunk; the dot product is
// a shuffle reduction; the X^T r outer-product accumulates in registers across rows; CTAs fold
// through shared memory and emit one partial per CTA; `fold` sums the partials into
// out = [g_w (F), g_b, loss, n_rows] -- exactly the payload handed to the K3 small all-reduce.
// ----------------------------------------------------------------------------------------
constexpr int GLM_THREADS = 256;
constexpr int GLM_MAXCH = 4;          // up to F = 1024

template <typename T>
__global__ void __launch_bounds__(GLM_THREADS) glm_logistic_kernel(const T* __restrict__ X, const float* __restrict__ y,
                                                                    const float* __restrict__ w,
                                                                    float* __restrict__ part, int rows, int F) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nwarp = GLM_THREADS / 32;
    const int nch = F / 256;
    float wr[GLM_MAXCH][8], g[GLM_MAXCH][8];
#pragma unroll
    for (int c = 0; c < GLM_MAXCH; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) { g[c][k] = 0.f; wr[c][k] = (c < nch) ? w[c * 256 + lane * 8 + k] : 0.f; }
    float gb = 0.f, loss = 0.f;
    const float bias = w[F];          // w = [coefficients (F), intercept]
    for (long long row = (long long)blockIdx.x * nwarp + wid; row < rows; row += (long long)gridDim.x * nwarp) {
        float x[GLM_MAXCH][8];
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < GLM_MAXCH; ++c) {
            if (c < nch) {
                const T* p = X + (size_t)row * F + c * 256 + lane * 8;
                if constexpr (sizeof(T) == 2) {
                    uint4 t;
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
                    float2 a = unpack_bf16x2(t.x), b2 = unpack_bf16x2(t.y), c2 = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
                    x[c][0] = a.x; x[c][1] = a.y; x[c][2] = b2.x; x[c][3] = b2.y;
                    x[c][4] = c2.x; x[c][5] = c2.y; x[c][6] = d.x; x[c][7] = d.y;
                } else {
                    float4 a = ldg_stream_f4(reinterpret_cast<const float4*>(p));
                    float4 b2 = ldg_stream_f4(reinterpret_cast<const float4*>(p) + 1);
                    x[c][0] = a.x; x[c][1] = a.y; x[c][2] = a.z; x[c][3] = a.w;
                    x[c][4] = b2.x; x[c][5] = b2.y; x[c][6] = b2.z; x[c][7] = b2.w;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) dot = fmaf(x[c][k], wr[c][k], dot);
            }
        }
        dot = warp_sum(dot) + bias;
        const float yy = y[row];
        const float p = 1.f / (1.f + __expf(-dot));
        const float r = p - yy;
        // numerically stable BCE: max(z,0) - z*y + log(1+exp(-|z|))
        if (lane == 0) { loss += fmaxf(dot, 0.f) - dot * yy + log1pf(__expf(-fabsf(dot))); gb += r; }
#pragma unroll
        for (int c = 0; c < GLM_MAXCH; ++c)
            if (c < nch) {
#pragma unroll
                for (int k = 0; k < 8; ++k) g[c][k] = fmaf(r, x[c][k], g[c][k]);
            }
    }
    // fold the warps of this CTA
    extern __shared__ float sm[];   // [nwarp][F + 2]
    float* mine = sm + (size_t)wid * (F + 2);
#pragma unroll
    for (int c = 0; c < GLM_MAXCH; ++c)
        if (c < nch) {
#pragma unroll
            for (int k = 0; k < 8; ++k) mine[c * 256 + lane * 8 + k] = g[c][k];
        }
    if (lane == 0) { mine[F] = gb; mine[F + 1] = loss; }
    __syncthreads();
    for (int i = threadIdx.x; i < F + 2; i += GLM_THREADS) {
        float a = 0.f;
        for (int w2 = 0; w2 < nwarp; ++w2) a += sm[(size_t)w2 * (F + 2) + i];
        part[(size_t)blockIdx.x * (F + 2) + i] = a;
    }
}
__global__ void glm_fold_kernel(const float* __restrict__ part, float* __restrict__ out, int nparts, int F, float rows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F + 2) {
        float a = 0.f;
        for (int p = 0; p < nparts; ++p) a += part[(size_t)p * (F + 2) + i];
        out[i] = a;
    }
    if (i == F + 2) out[F + 2] = rows;
}
extern "C" int v6_glm_logistic_grad(const void* X, const float* y, const float* w, float* part,
                                    int max_parts, float* out, int rows, int F, int bf16, cudaStream_t s) {
    if (F % 256 != 0 || F > 256 * GLM_MAXCH) return (int)cudaErrorInvalidValue;
    int grid = (rows + 7) / 8;
    if (grid > 148 * 4) grid = 148 * 4;
    if (grid > max_parts) grid = max_parts;
    const size_t smem = (size_t)(GLM_THREADS / 32) * (F + 2) * sizeof(float);
    if (bf16) glm_logistic_kernel<__nv_bfloat16><<<grid, GLM_THREADS, smem, s>>>((const __nv_bfloat16*)X, y, w, part, rows, F);
    else glm_logistic_kernel<float><<<grid, GLM_THREADS, smem, s>>>((const float*)X, y, w, part, rows, F);
    V6_CHECK_LAUNCH();
    glm_fold_kernel<<<(F + 3 + 255) / 256, 256, 0, s>>>(part, out, grid, F, (float)rows);
    V6_CHECK_LAUNCH();
    return 0;
}
