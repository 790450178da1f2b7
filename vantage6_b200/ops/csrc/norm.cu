// K5 of SURVEY.md 2.6: LayerNorm (BERT, 768) and RMSNorm (Llama, 4096) forward + backward,
// with the residual add fused into the forward and the bf16 cast fused into both directions.
//
// Memory-bound design: every activation element is read once and written once per pass.
// A row is owned by a thread group (TPR threads, a power of two <= 256) which keeps the whole
// row in registers (<= 4 x 16 B vectors per thread), so mean/var use the exact two-pass form
// without re-reading HBM.  gamma/beta stay fp32 (they live in the flat fp32 master buffer).
// Backward: persistent CTAs walk rows; dgamma/dbeta are accumulated in registers across rows,
// written as per-CTA partials and folded by a second tiny kernel (no atomics).
#include <stdlib.h>
#include "common.cuh"
#include "api.h"

template <typename T> struct IO;
template <> struct IO<__nv_bfloat16> {
    static constexpr int VEC = 8;
    V6_DEVINL static void load(const __nv_bfloat16* p, float (&v)[8]) {
        uint4 t = *reinterpret_cast<const uint4*>(p);
        float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
    }
    V6_DEVINL static void store(__nv_bfloat16* p, const float (&v)[8]) {
        *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                  pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
};
template <> struct IO<float> {
    static constexpr int VEC = 8;
    V6_DEVINL static void load(const float* p, float (&v)[8]) {
        float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    V6_DEVINL static void store(float* p, const float (&v)[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};

// sum across the TPR threads that own one row. TPR <= 32: shuffles only. TPR > 32: smem.
template <int TPR>
V6_DEVINL float group_sum(float v, float* smem, int row_in_cta, int t) {
    if constexpr (TPR <= 32) {
#pragma unroll
        for (int o = TPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    } else {
        constexpr int W = TPR / 32;
        v = warp_sum(v);
        __syncthreads();
        if ((t & 31) == 0) smem[row_in_cta * W + (t >> 5)] = v;
        __syncthreads();
        float r = 0.f;
#pragma unroll
        for (int w = 0; w < W; ++w) r += smem[row_in_cta * W + w];
        return r;
    }
}

constexpr int NORM_THREADS = 256;
constexpr int MAXV = 4;   // 16 B vectors per thread per row

template <typename T, int TPR, bool RMS>
__global__ void __launch_bounds__(NORM_THREADS) norm_fwd_kernel(
    const T* __restrict__ x, const T* __restrict__ residual, const float* __restrict__ gamma,
    const float* __restrict__ beta, T* __restrict__ y, T* __restrict__ res_out,
    float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int cols, float eps) {
    constexpr int RPC = NORM_THREADS / TPR;     // rows per CTA iteration
    __shared__ float red[NORM_THREADS / 32];
    const int t = threadIdx.x % TPR, rin = threadIdx.x / TPR;
    const int nvec = cols / 8;                  // vectors per row
    for (int row0 = blockIdx.x * RPC; row0 < rows; row0 += gridDim.x * RPC) {
        const int row = row0 + rin;
        const bool active = row < rows;
        float v[MAXV][8];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int vi = t + j * TPR;
            if (active && vi < nvec) {
                IO<T>::load(x + (size_t)row * cols + vi * 8, v[j]);
                if (residual) {
                    float r[8];
                    IO<T>::load(residual + (size_t)row * cols + vi * 8, r);
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[j][k] += r[k];
                }
                if (res_out) IO<T>::store(res_out + (size_t)row * cols + vi * 8, v[j]);
#pragma unroll
                for (int k = 0; k < 8; ++k) s += v[j][k];
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[j][k] = 0.f;
            }
        }
        float mean = 0.f;
        if constexpr (!RMS) mean = group_sum<TPR>(s, red, rin, t) / cols;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int vi = t + j * TPR;
            if (vi < nvec) {
#pragma unroll
                for (int k = 0; k < 8; ++k) { const float d = v[j][k] - mean; q += d * d; }
            }
        }
        const float var = group_sum<TPR>(q, red, rin, t) / cols;
        const float rstd = rsqrtf(var + eps);
        if (active && t == 0) { if (mean_out) mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int vi = t + j * TPR;
            if (active && vi < nvec) {
                float g[8], o[8];
                IO<float>::load(gamma + vi * 8, g);
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = (v[j][k] - mean) * rstd * g[k];
                if constexpr (!RMS) {
                    float b[8];
                    IO<float>::load(beta + vi * 8, b);
#pragma unroll
                    for (int k = 0; k < 8; ++k) o[k] += b[k];
                }
                IO<T>::store(y + (size_t)row * cols + vi * 8, o);
            }
        }
    }
}

// backward. x_in is the (post-residual) input of the norm. dres (optional) is the incoming
// gradient of the residual stream, added into dx (fused residual-gradient add).
template <typename T, int TPR, bool RMS>
__global__ void __launch_bounds__(NORM_THREADS) norm_bwd_kernel(
    const T* __restrict__ dy, const T* __restrict__ x_in, const T* __restrict__ dres,
    const float* __restrict__ gamma, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
    T* __restrict__ dx, float* __restrict__ dgamma_part, float* __restrict__ dbeta_part,
    int rows, int cols) {
    constexpr int RPC = NORM_THREADS / TPR;
    __shared__ float red[NORM_THREADS / 32];
    const int t = threadIdx.x % TPR, rin = threadIdx.x / TPR;
    const int nvec = cols / 8;
    float dg[MAXV][8], db[MAXV][8];
#pragma unroll
    for (int j = 0; j < MAXV; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) { dg[j][k] = 0.f; db[j][k] = 0.f; }

    for (int row0 = blockIdx.x * RPC; row0 < rows; row0 += gridDim.x * RPC) {
        const int row = row0 + rin;
        const bool active = row < rows;
        const float mean = (active && !RMS) ? mean_in[row] : 0.f;
        const float rstd = active ? rstd_in[row] : 0.f;
        float xh[MAXV][8], dyg[MAXV][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int vi = t + j * TPR;
            if (active && vi < nvec) {
                float d[8], g[8];
                IO<T>::load(dy + (size_t)row * cols + vi * 8, d);
                IO<T>::load(x_in + (size_t)row * cols + vi * 8, xh[j]);
                IO<float>::load(gamma + vi * 8, g);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    xh[j][k] = (xh[j][k] - mean) * rstd;
                    dyg[j][k] = d[k] * g[k];
                    s1 += dyg[j][k];
                    s2 += dyg[j][k] * xh[j][k];
                    dg[j][k] += d[k] * xh[j][k];
                    db[j][k] += d[k];
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) { xh[j][k] = 0.f; dyg[j][k] = 0.f; }
            }
        }
        float m1 = 0.f;
        if constexpr (!RMS) m1 = group_sum<TPR>(s1, red, rin, t) / cols;
        const float m2 = group_sum<TPR>(s2, red, rin, t) / cols;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int vi = t + j * TPR;
            if (active && vi < nvec) {
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = rstd * (dyg[j][k] - m1 - xh[j][k] * m2);
                if (dres) {
                    float r[8];
                    IO<T>::load(dres + (size_t)row * cols + vi * 8, r);
#pragma unroll
                    for (int k = 0; k < 8; ++k) o[k] += r[k];
                }
                IO<T>::store(dx + (size_t)row * cols + vi * 8, o);
            }
        }
    }
    // fold the RPC row-groups of this CTA through shared memory, then write one partial per CTA
    extern __shared__ float part[];    // [RPC][cols] for dgamma, then the same for dbeta
    float* pg = part;
    float* pb = part + (size_t)RPC * cols;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int vi = t + j * TPR;
        if (vi < nvec) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                pg[(size_t)rin * cols + vi * 8 + k] = dg[j][k];
                if (!RMS) pb[(size_t)rin * cols + vi * 8 + k] = db[j][k];
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cols; c += NORM_THREADS) {
        float a = 0.f, b = 0.f;
        for (int r = 0; r < RPC; ++r) { a += pg[(size_t)r * cols + c]; if (!RMS) b += pb[(size_t)r * cols + c]; }
        dgamma_part[(size_t)blockIdx.x * cols + c] = a;
        if (!RMS) dbeta_part[(size_t)blockIdx.x * cols + c] = b;
    }
}

// out_k[c] (+)= sum_p part_k[p][c] for k = blockIdx.y (dgamma, dbeta).  Block = 32 columns x 8 partial-slices: each
// thread sums every 8th partial with 4 loads in flight, the 8 slices fold through shared memory (v1: one thread per
// column looping over up to 1184 partials, 21 us per launch and two launches per LayerNorm --
// profiles/launches_bert_base_r1c.txt).
__global__ void __launch_bounds__(256) norm_param_grad_fold(const float* __restrict__ part0, const float* __restrict__ part1,
                                                            float* __restrict__ out0, float* __restrict__ out1, int nparts, int cols,
                                                            int accumulate) {
    __shared__ float sm[8][33];
    const float* __restrict__ part = blockIdx.y ? part1 : part0;
    float* __restrict__ out = blockIdx.y ? out1 : out0;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float a = 0.f;
    if (c < cols) {
#pragma unroll 4
        for (int p = ty; p < nparts; p += 8) a += part[(size_t)p * cols + c];
    }
    sm[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && c < cols) {
#pragma unroll
        for (int s2 = 1; s2 < 8; ++s2) a += sm[s2][tx];
        out[c] = accumulate ? out[c] + a : a;
    }
}

// ---------------------------------------------------------------------------------- backward, second generation
// The one-kernel backward above keeps 4 x 8 x 4 floats of row data AND 64 partial dgamma / dbeta accumulators per thread:
// 157 registers, one CTA per SM, 1.3-3.1 TB/s (profiles/ncu_misc_r1a.md).  Split by access pattern instead:
//   norm_bwd_dx_kernel     row-wise: dx (+ residual gradient); the row stays packed (bf16 vectors) in registers, no
//                          parameter-gradient state -> 4 CTAs/SM
//   norm_bwd_param_kernel  column-wise: a thread owns 8 columns and walks rows (dgamma += dy.xhat, dbeta += dy), 8 warps x
//                          2 rows in flight, per-CTA partial [slice][cols] -> norm_param_grad_fold
template <typename T, int TPR, bool RMS, int VPT>
__global__ void __launch_bounds__(NORM_THREADS, (VPT <= 2 ? 4 : (VPT == 3 ? 3 : 2))) norm_bwd_dx_kernel(
    const T* __restrict__ dy, const T* __restrict__ x_in, const T* __restrict__ dres, const float* __restrict__ gamma,
    const float* __restrict__ mean_in, const float* __restrict__ rstd_in, T* __restrict__ dx, int rows, int cols) {
    constexpr int RPC = NORM_THREADS / TPR;
    __shared__ float red[NORM_THREADS / 32 * 2];
    const int t = threadIdx.x % TPR, rin = threadIdx.x / TPR;
    const int nvec = cols / 8;
    for (int row0 = blockIdx.x * RPC; row0 < rows; row0 += gridDim.x * RPC) {
        const int row = row0 + rin;
        const bool active = row < rows;
        const float mean = (active && !RMS) ? mean_in[row] : 0.f;
        const float rstd = active ? rstd_in[row] : 0.f;
        float d[VPT][8], xv[VPT][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int vi = t + j * TPR;
            if (active && vi < nvec) {
                IO<T>::load(dy + (size_t)row * cols + vi * 8, d[j]);
                IO<T>::load(x_in + (size_t)row * cols + vi * 8, xv[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int vi = t + j * TPR;
            if (active && vi < nvec) {
                float g[8];
                IO<float>::load(gamma + vi * 8, g);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    xv[j][k] = (xv[j][k] - mean) * rstd;
                    d[j][k] *= g[k];
                    s1 += d[j][k];
                    s2 = fmaf(d[j][k], xv[j][k], s2);
                }
            }
        }
        float m1 = 0.f;
        if constexpr (!RMS) m1 = group_sum<TPR>(s1, red, rin, t) / cols;
        const float m2 = group_sum<TPR>(s2, red + NORM_THREADS / 32, rin, t) / cols;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int vi = t + j * TPR;
            if (active && vi < nvec) {
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = rstd * (d[j][k] - m1 - xv[j][k] * m2);
                if (dres) {
                    float r[8];
                    IO<T>::load(dres + (size_t)row * cols + vi * 8, r);
#pragma unroll
                    for (int k = 0; k < 8; ++k) o[k] += r[k];
                }
                IO<T>::store(dx + (size_t)row * cols + vi * 8, o);
            }
        }
    }
}

// grid (ceil(cols / 256), slices); block 256 = 8 warps; lane owns 8 columns, warps stride the slice's rows
template <typename T, bool RMS>
__global__ void __launch_bounds__(256) norm_bwd_param_kernel(
    const T* __restrict__ dy, const T* __restrict__ x_in, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
    float* __restrict__ dgamma_part, float* __restrict__ dbeta_part, int rows, int cols, int rows_per_slice) {
    __shared__ float sm[2][8][256 + 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 256 + lane * 8;
    const bool col_ok = c0 < cols;
    const int r_begin = blockIdx.y * rows_per_slice, r_end = min(rows, r_begin + rows_per_slice);
    float dg[8], db[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { dg[k] = 0.f; db[k] = 0.f; }
    for (int r = r_begin + warp; r < r_end; r += 16) {
        float d[2][8], xv[2][8], mu[2], rs[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int rr = r + 8 * u;
            mu[u] = 0.f; rs[u] = 0.f;
            if (rr < r_end && col_ok) {
                IO<T>::load(dy + (size_t)rr * cols + c0, d[u]);
                IO<T>::load(x_in + (size_t)rr * cols + c0, xv[u]);
                mu[u] = RMS ? 0.f : mean_in[rr];
                rs[u] = rstd_in[rr];
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) { d[u][k] = 0.f; xv[u][k] = 0.f; }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                dg[k] = fmaf(d[u][k], (xv[u][k] - mu[u]) * rs[u], dg[k]);
                db[k] += d[u][k];
            }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { sm[0][warp][lane * 8 + k] = dg[k]; sm[1][warp][lane * 8 + k] = db[k]; }
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < cols) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a += sm[0][w][threadIdx.x]; b += sm[1][w][threadIdx.x]; }
        dgamma_part[(size_t)blockIdx.y * cols + c] = a;
        if (!RMS) dbeta_part[(size_t)blockIdx.y * cols + c] = b;
    }
}

static inline int pick_tpr(int cols) {
    const int nvec = cols / 8;
    int tpr = 8;
    while (tpr * MAXV < nvec && tpr < 256) tpr <<= 1;
    return tpr;
}

#define DISPATCH_TPR(TPRV, ...) \
    switch (TPRV) { \
        case 8:   { constexpr int TPR = 8;   __VA_ARGS__; break; } \
        case 16:  { constexpr int TPR = 16;  __VA_ARGS__; break; } \
        case 32:  { constexpr int TPR = 32;  __VA_ARGS__; break; } \
        case 64:  { constexpr int TPR = 64;  __VA_ARGS__; break; } \
        case 128: { constexpr int TPR = 128; __VA_ARGS__; break; } \
        default:  { constexpr int TPR = 256; __VA_ARGS__; break; } \
    }

template <typename T, bool RMS>
static int launch_fwd(const void* x, const void* residual, const float* gamma, const float* beta, void* y,
                      void* res_out, float* mean, float* rstd, int rows, int cols, float eps, cudaStream_t s) {
    if (cols % 8 != 0 || cols > 8 * MAXV * 256) return (int)cudaErrorInvalidValue;
    const int tpr = pick_tpr(cols);
    const int rpc = NORM_THREADS / tpr;
    int grid = (rows + rpc - 1) / rpc;
    if (grid > 148 * 8) grid = 148 * 8;
    DISPATCH_TPR(tpr, (norm_fwd_kernel<T, TPR, RMS><<<grid, NORM_THREADS, 0, s>>>(
        (const T*)x, (const T*)residual, gamma, beta, (T*)y, (T*)res_out, mean, rstd, rows, cols, eps)));
    V6_CHECK_LAUNCH();
    return 0;
}

template <typename T, bool RMS>
static int launch_bwd(const void* dy, const void* x_in, const void* dres, const float* gamma, const float* mean,
                      const float* rstd, void* dx, float* dgamma, float* dbeta, float* scratch, int scratch_parts,
                      int rows, int cols, int accumulate, cudaStream_t s) {
    if (cols % 8 != 0 || cols > 8 * MAXV * 256) return (int)cudaErrorInvalidValue;
    static const bool force_v1 = [] { const char* e = getenv("V6B200_NORM_BWD"); return e && e[0] == '1'; }();      // the one-kernel form (A/B)
    const int tpr = pick_tpr(cols);
    // split form where a row fits one warp (cols <= 1024: shuffle-only row sums): 0.091 vs 0.105 ms at 32768 x 768, BERT-base round
    // 30.4 vs 31.6 ms; wide rows (Llama, 4096) need the shared-memory row sums twice per row and are faster in the one-kernel
    // form (0.137 vs 0.198 ms) -- profiles/kernel_bench_norm_r2.txt
    const bool v1 = force_v1 || tpr > 32;
    const int rpc = NORM_THREADS / tpr;
    float* pg = scratch;
    float* pb = scratch + (size_t)scratch_parts * cols;
    if (v1) {
        int grid = (rows + rpc - 1) / rpc;
        if (grid > scratch_parts) grid = scratch_parts;
        const size_t smem = (size_t)rpc * cols * sizeof(float) * (RMS ? 1 : 2);
        DISPATCH_TPR(tpr, {
            auto kern = norm_bwd_kernel<T, TPR, RMS>;
            if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<grid, NORM_THREADS, smem, s>>>((const T*)dy, (const T*)x_in, (const T*)dres, gamma, mean, rstd,
                                                    (T*)dx, pg, pb, rows, cols);
        });
        V6_CHECK_LAUNCH();
        if (dgamma)      // frozen norm weights (LoRA fine-tuning): nothing to fold
            norm_param_grad_fold<<<dim3((cols + 31) / 32, RMS ? 1 : 2), 256, 0, s>>>(pg, pb, dgamma, dbeta, grid, cols, accumulate);
        V6_CHECK_LAUNCH();
        return 0;
    }
    // parameter gradients first (they only read dy / x), then dx
    if (dgamma) {
        int slices = (rows + 31) / 32;
        if (slices > scratch_parts) slices = scratch_parts;
        const int rps = (rows + slices - 1) / slices;
        slices = (rows + rps - 1) / rps;
        norm_bwd_param_kernel<T, RMS><<<dim3((cols + 255) / 256, slices), 256, 0, s>>>((const T*)dy, (const T*)x_in, mean, rstd, pg, pb, rows,
                                                                                        cols, rps);
        V6_CHECK_LAUNCH();
        norm_param_grad_fold<<<dim3((cols + 31) / 32, RMS ? 1 : 2), 256, 0, s>>>(pg, pb, dgamma, dbeta, slices, cols, accumulate);
        V6_CHECK_LAUNCH();
    }
    const int nvec = cols / 8, vpt = (nvec + tpr - 1) / tpr;
    int grid = (rows + rpc - 1) / rpc;
    if (grid > 148 * 4) grid = 148 * 4;
#define V6_NORM_DX(VPTV) norm_bwd_dx_kernel<T, TPR, RMS, VPTV><<<grid, NORM_THREADS, 0, s>>>((const T*)dy, (const T*)x_in, (const T*)dres, gamma, mean, rstd, (T*)dx, rows, cols)
    DISPATCH_TPR(tpr, {
        if (vpt <= 1) V6_NORM_DX(1);
        else if (vpt == 2) V6_NORM_DX(2);
        else if (vpt == 3) V6_NORM_DX(3);
        else V6_NORM_DX(4);
    });
#undef V6_NORM_DX
    V6_CHECK_LAUNCH();
    return 0;
}

extern "C" int v6_layernorm_fwd(const void* x, const void* residual, const float* gamma, const float* beta, void* y,
                                void* res_out, float* mean, float* rstd, int rows, int cols, float eps, int bf16,
                                cudaStream_t s) {
    return bf16 ? launch_fwd<__nv_bfloat16, false>(x, residual, gamma, beta, y, res_out, mean, rstd, rows, cols, eps, s)
                : launch_fwd<float, false>(x, residual, gamma, beta, y, res_out, mean, rstd, rows, cols, eps, s);
}
extern "C" int v6_rmsnorm_fwd(const void* x, const void* residual, const float* gamma, void* y, void* res_out,
                              float* rstd, int rows, int cols, float eps, int bf16, cudaStream_t s) {
    return bf16 ? launch_fwd<__nv_bfloat16, true>(x, residual, gamma, nullptr, y, res_out, nullptr, rstd, rows, cols, eps, s)
                : launch_fwd<float, true>(x, residual, gamma, nullptr, y, res_out, nullptr, rstd, rows, cols, eps, s);
}
extern "C" int v6_layernorm_bwd(const void* dy, const void* x_in, const void* dres, const float* gamma, const float* mean,
                                const float* rstd, void* dx, float* dgamma, float* dbeta, float* scratch,
                                int scratch_parts, int rows, int cols, int accumulate, int bf16, cudaStream_t s) {
    return bf16 ? launch_bwd<__nv_bfloat16, false>(dy, x_in, dres, gamma, mean, rstd, dx, dgamma, dbeta, scratch, scratch_parts, rows, cols, accumulate, s)
                : launch_bwd<float, false>(dy, x_in, dres, gamma, mean, rstd, dx, dgamma, dbeta, scratch, scratch_parts, rows, cols, accumulate, s);
}
extern "C" int v6_rmsnorm_bwd(const void* dy, const void* x_in, const void* dres, const float* gamma, const float* rstd,
                              void* dx, float* dgamma, float* scratch, int scratch_parts, int rows, int cols,
                              int accumulate, int bf16, cudaStream_t s) {
    return bf16 ? launch_bwd<__nv_bfloat16, true>(dy, x_in, dres, gamma, nullptr, rstd, dx, dgamma, nullptr, scratch, scratch_parts, rows, cols, accumulate, s)
                : launch_bwd<float, true>(dy, x_in, dres, gamma, nullptr, rstd, dx, dgamma, nullptr, scratch, scratch_parts, rows, cols, accumulate, s);
}
