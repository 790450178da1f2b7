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
// X:[rows, F] bf16 / fp32 row-major, F % 64 == 0, F <= 512; w = [coefficients (F), intercept].
//
// Mapping (v3; v1 used one warp per row and reached only 16% of HBM bandwidth -- one 16 B load in
// flight per lane and a 5-step shuffle per row): a row is owned by 8 lanes, so a warp handles 4
// rows at a time, 2x unrolled = 8 rows in flight per warp; lane l of the group reads vectors
// l, l+8, l+16, ... of its row (8 lanes x 16 B = one 128 B line per access).  The dot product is a
// 3-step shuffle inside the 8-lane group; X^T r accumulates in registers (F/8 per lane); warps and
// lane groups fold through shared memory into one partial per CTA; `fold` sums the partials into
// out = [g_w (F), g_b, loss, n_rows] -- exactly the payload handed to the K3 small all-reduce.
// (A v4 with a software-pipelined row stream and the coefficients in shared memory measured 5% SLOWER
// -- 163 vs 155 us, profiles/kernel_bench_r1b.json vs the r1d run -- and was reverted: the kernel is bound by
// issue slots + latency at 2 CTAs/SM, not by bytes in flight.  The tcgen05 formulation (z = X w and g = X^T r as
// two UMMA GEMVs with r split into hi/lo bf16 columns, X tiles fed by TMA, MN-major A for the second product)
// is the round-2 item.)
// ----------------------------------------------------------------------------------------
constexpr int GLM_THREADS = 256;

// Rows are kept in registers in their RAW storage format (bf16: one uint4 per 8 features) and unpacked
// twice (dot product, then outer product): the unpack is 2 ALU ops per pair, the registers saved
// (32 instead of 64 for two rows in flight) keep the kernel at 2 CTAs/SM without spills.
template <typename T> struct GlmLoad;
template <> struct GlmLoad<__nv_bfloat16> {
    using Raw = uint4;
    static constexpr int U = 2;                                // rows in flight per lane group
    V6_DEVINL static Raw load_raw(const __nv_bfloat16* p) {
        uint4 t;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
        return t;
    }
    V6_DEVINL static void unpack(const Raw& t, float (&v)[8]) {
        float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
    }
};
template <> struct GlmLoad<float> {
    struct Raw { float4 a, b; };
    static constexpr int U = 1;
    V6_DEVINL static Raw load_raw(const float* p) {
        Raw r;
        r.a = ldg_stream_f4(reinterpret_cast<const float4*>(p));
        r.b = ldg_stream_f4(reinterpret_cast<const float4*>(p) + 1);
        return r;
    }
    V6_DEVINL static void unpack(const Raw& r, float (&v)[8]) {
        v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
    }
};

template <typename T, int NV /* 8-feature vectors per lane = F/64 */>
__global__ void __launch_bounds__(GLM_THREADS, 2) glm_logistic_kernel(const T* __restrict__ X, const float* __restrict__ y,
                                                                       const float* __restrict__ w,
                                                                       float* __restrict__ part, int rows) {
    constexpr int F = NV * 64;
    constexpr int U = GlmLoad<T>::U;                           // row unroll per lane group
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gl = lane & 7, grp = lane >> 3;                  // lane in group, group in warp
    const int groups_per_cta = (GLM_THREADS / 32) * 4;
    const int my_group = wid * 4 + grp;
    float wr[NV][8], g[NV][8];
#pragma unroll
    for (int c = 0; c < NV; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) { g[c][k] = 0.f; wr[c][k] = w[(c * 8 + gl) * 8 + k]; }
    const float bias = w[F];
    float gb = 0.f, loss = 0.f;
    const long long stride = (long long)gridDim.x * groups_per_cta;
    for (long long row0 = (long long)blockIdx.x * groups_per_cta + my_group; row0 < rows; row0 += stride * U) {
        typename GlmLoad<T>::Raw xr[U][NV];
        float dot[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = row0 + u * stride;
            dot[u] = 0.f;
            if (row < rows) {
#pragma unroll
                for (int c = 0; c < NV; ++c) xr[u][c] = GlmLoad<T>::load_raw(X + (size_t)row * F + (c * 8 + gl) * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = row0 + u * stride;
            const bool ok = row < rows;                        // uniform inside the 8-lane group
            if (ok) {
#pragma unroll
                for (int c = 0; c < NV; ++c) {
                    float xv[8];
                    GlmLoad<T>::unpack(xr[u][c], xv);
#pragma unroll
                    for (int k = 0; k < 8; ++k) dot[u] = fmaf(xv[k], wr[c][k], dot[u]);
                }
            }
            float d = dot[u];
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            d += __shfl_xor_sync(0xffffffffu, d, 4);
            if (ok) {
                const float z = d + bias;
                const float yy = y[row];
                const float r = 1.f / (1.f + __expf(-z)) - yy;
                if (gl == 0) { loss += fmaxf(z, 0.f) - z * yy + log1pf(__expf(-fabsf(z))); gb += r; }
#pragma unroll
                for (int c = 0; c < NV; ++c) {
                    float xv[8];
                    GlmLoad<T>::unpack(xr[u][c], xv);
#pragma unroll
                    for (int k = 0; k < 8; ++k) g[c][k] = fmaf(r, xv[k], g[c][k]);
                }
            }
        }
    }
    // fold the 4 lane groups of each warp (same gl -> same features), then the warps through smem
#pragma unroll
    for (int c = 0; c < NV; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v = g[c][k];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            g[c][k] = v;
        }
    gb += __shfl_xor_sync(0xffffffffu, gb, 8);  gb += __shfl_xor_sync(0xffffffffu, gb, 16);
    loss += __shfl_xor_sync(0xffffffffu, loss, 8);  loss += __shfl_xor_sync(0xffffffffu, loss, 16);
    __shared__ float sm[(GLM_THREADS / 32) * (F + 2)];
    float* mine = sm + wid * (F + 2);
    if (grp == 0) {
#pragma unroll
        for (int c = 0; c < NV; ++c)
#pragma unroll
            for (int k = 0; k < 8; ++k) mine[(c * 8 + gl) * 8 + k] = g[c][k];
        if (gl == 0) { mine[F] = gb; mine[F + 1] = loss; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < F + 2; i += GLM_THREADS) {
        float a = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < GLM_THREADS / 32; ++w2) a += sm[w2 * (F + 2) + i];
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

template <typename T>
static int launch_glm(const void* X, const float* y, const float* w, float* part, int grid, int rows, int F, cudaStream_t s) {
    switch (F / 64) {
        case 1: glm_logistic_kernel<T, 1><<<grid, GLM_THREADS, 0, s>>>((const T*)X, y, w, part, rows); break;
        case 2: glm_logistic_kernel<T, 2><<<grid, GLM_THREADS, 0, s>>>((const T*)X, y, w, part, rows); break;
        case 4: glm_logistic_kernel<T, 4><<<grid, GLM_THREADS, 0, s>>>((const T*)X, y, w, part, rows); break;
        case 8: glm_logistic_kernel<T, 8><<<grid, GLM_THREADS, 0, s>>>((const T*)X, y, w, part, rows); break;
        default: return (int)cudaErrorInvalidValue;
    }
    return 0;
}

// out = [sum of the partials (F + 2), n_rows]; used by the experimental tensor-core path (glm_tc.cu)
extern "C" int v6_glm_fold(const float* part, float* out, int nparts, int F, int rows, cudaStream_t s) {
    glm_fold_kernel<<<(F + 3 + 255) / 256, 256, 0, s>>>(part, out, nparts, F, (float)rows);
    V6_CHECK_LAUNCH();
    return 0;
}

extern "C" int v6_glm_logistic_grad(const void* X, const float* y, const float* w, float* part,
                                    int max_parts, float* out, int rows, int F, int bf16, cudaStream_t s) {
    if (F % 64 != 0 || F > 512) return (int)cudaErrorInvalidValue;
    const int groups = (GLM_THREADS / 32) * 4;
    int grid = (rows + groups * 2 - 1) / (groups * 2);
    if (grid > 148 * 2) grid = 148 * 2;                          // 2 resident CTAs per SM, persistent over rows
    if (grid > max_parts) grid = max_parts;
    if (grid < 1) grid = 1;
    int rc = bf16 ? launch_glm<__nv_bfloat16>(X, y, w, part, grid, rows, F, s) : launch_glm<float>(X, y, w, part, grid, rows, F, s);
    if (rc) return rc;
    V6_CHECK_LAUNCH();
    glm_fold_kernel<<<(F + 3 + 255) / 256, 256, 0, s>>>(part, out, grid, F, (float)rows);
    V6_CHECK_LAUNCH();
    return 0;
}
