// Fused BatchNorm(+residual add)(+ReLU) forward / backward for NHWC bf16 activations (ResNet-50).
//
// The first profile of the flagship local step (profiles/launches_resnet50_eager_r1.txt) showed
// PyTorch's native channels-last BatchNorm kernels + stand-alone ReLU / add kernels taking ~3/4 of
// the GPU time of a ResNet-50 step on B200 -- the convolutions (cuDNN/CUTLASS sm100 kernels) only
// ~15%.  These kernels make that part memory-bound at HBM speed and remove the separate ReLU /
// residual-add passes:
//
//   forward :  stats (1 read of x)  ->  finalize (per channel)  ->  y = relu(x*scale + bias + res)
//   backward:  reduce (dy, y, x)    ->  finalize (dgamma, dbeta) ->  dx (and dres) in one pass
//
// x viewed as [R = N*H*W, C], C % 8 == 0 and C/8 a power of two <= 256 (64..2048 in ResNet-50).
// Thread mapping: threadIdx % (C/8) owns 8 consecutive channels (one 16 B vector), threadIdx / (C/8)
// is a row lane; CTAs stride over rows.  Per-channel sums are accumulated in registers across rows
// (shifted by x[0,c] to avoid cancellation in E[x^2]-E[x]^2), folded across row lanes in shared
// memory, written as per-CTA partials and summed by the finalize kernel (no atomics, deterministic).
#include "common.cuh"
#include "api.h"

namespace bn {

constexpr int THREADS = 256;
constexpr int MAX_PARTS = 148 * 2;

V6_DEVINL void load8(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
V6_DEVINL void store8(__nv_bfloat16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                              pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
V6_DEVINL void loadf8(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// fold the per-thread accumulators a[8], b[8] over the row lanes of the CTA and write one partial
// row [2C] = [sum_a (C) | sum_b (C)] for this CTA.
V6_DEVINL void fold_and_write(float (&a)[8], float (&b)[8], float* smem, float* part, int C, int cg, int rl, int RL) {
    const int CG = C >> 3;
    float* sa = smem;                       // [RL][C]
    float* sb = smem + (size_t)RL * C;      // [RL][C]
#pragma unroll
    for (int k = 0; k < 8; ++k) { sa[(size_t)rl * C + cg * 8 + k] = a[k]; sb[(size_t)rl * C + cg * 8 + k] = b[k]; }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += THREADS) {
        float x = 0.f, y = 0.f;
        for (int r = 0; r < RL; ++r) { x += sa[(size_t)r * C + c]; y += sb[(size_t)r * C + c]; }
        part[(size_t)blockIdx.x * 2 * C + c] = x;
        part[(size_t)blockIdx.x * 2 * C + C + c] = y;
    }
    (void)CG;
}

// ---------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(THREADS) bn_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ part,
                                                           long long R, int C) {
    extern __shared__ float smem[];
    const int CG = C >> 3, RL = THREADS / CG;
    const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
    float shift[8], s1[8], s2[8];
    load8(x + cg * 8, shift);                                   // row 0 as the per-channel shift
#pragma unroll
    for (int k = 0; k < 8; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
    const long long G = (long long)gridDim.x * RL;
    for (long long r = (long long)blockIdx.x * RL + rl; r < R; r += 4 * G) {      // 4 independent 16 B loads in flight
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (r + u * G < R) load8(x + (r + u * G) * C + cg * 8, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (r + u * G < R) {
#pragma unroll
                for (int k = 0; k < 8; ++k) { const float d = v[u][k] - shift[k]; s1[k] += d; s2[k] = fmaf(d, d, s2[k]); }
            }
    }
    fold_and_write(s1, s2, smem, part, C, cg, rl, RL);
}

// Finalize kernels: block = 32 channels x 8 partial-slices (256 threads).  Each thread sums every 8th
// partial with 4 independent loads in flight, the 8 slices are folded through shared memory, and the
// slice-0 thread of each channel does the per-channel math.  (v1 used one thread per channel looping
// over ~600 partials: latency-bound, 79 us per launch, 54% of the step in profiles/launches_*fusedbn*.)
constexpr int FIN_THREADS = 256;
V6_DEVINL int fold_parts(const float* __restrict__ part, int nparts, int C, float& a_out, float& b_out) {
    __shared__ float sa[8][33], sb[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float a = 0.f, b = 0.f;
    if (c < C) {
#pragma unroll 4
        for (int p = ty; p < nparts; p += 8) {
            a += part[(size_t)p * 2 * C + c];
            b += part[(size_t)p * 2 * C + C + c];
        }
    }
    sa[ty][tx] = a; sb[ty][tx] = b;
    __syncthreads();
    if (ty != 0 || c >= C) return -1;
#pragma unroll
    for (int s = 1; s < 8; ++s) { a += sa[s][tx]; b += sb[s][tx]; }
    a_out = a; b_out = b;
    return c;
}

__global__ void __launch_bounds__(FIN_THREADS) bn_fwd_finalize_kernel(const float* __restrict__ part, int nparts, const __nv_bfloat16* __restrict__ x,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float* __restrict__ running_mean, float* __restrict__ running_var,
                                       float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                       float* __restrict__ scale_out, float* __restrict__ bias_out,
                                       long long R, int C, float eps, float momentum) {
    float a, b;
    const int c = fold_parts(part, nparts, C, a, b);
    if (c < 0) return;
    const float shift = __bfloat162float(x[c]);
    const float invR = 1.f / (float)R;
    const float dm = a * invR;
    const float mean = shift + dm;
    const float var = fmaxf(b * invR - dm * dm, 0.f);
    const float rstd = rsqrtf(var + eps);
    mean_out[c] = mean;
    rstd_out[c] = rstd;
    const float sc = gamma[c] * rstd;
    scale_out[c] = sc;
    bias_out[c] = beta[c] - mean * sc;
    if (running_mean) {
        const float unbiased = R > 1 ? var * (float)R / (float)(R - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
}

template <bool RELU, bool RES>
__global__ void __launch_bounds__(THREADS) bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                                                           const float* __restrict__ scale, const float* __restrict__ bias,
                                                           __nv_bfloat16* __restrict__ y, long long R, int C) {
    const int CG = C >> 3, RL = THREADS / CG;
    const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
    float sc[8], bi[8];
    loadf8(scale + cg * 8, sc);
    loadf8(bias + cg * 8, bi);
    const long long G = (long long)gridDim.x * RL;
    for (long long r0 = (long long)blockIdx.x * RL + rl; r0 < R; r0 += 2 * G) {   // 2 rows x (x [+ res]) loads in flight
        float v[2][8], q[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long r = r0 + u * G;
            if (r < R) {
                load8(x + r * C + cg * 8, v[u]);
                if (RES) load8(res + r * C + cg * 8, q[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long r = r0 + u * G;
            if (r < R) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float t = fmaf(v[u][k], sc[k], bi[k]);
                    if (RES) t += q[u][k];
                    v[u][k] = RELU ? fmaxf(t, 0.f) : t;
                }
                store8(y + r * C + cg * 8, v[u]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------- backward
template <bool RELU>
__global__ void __launch_bounds__(THREADS) bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                                                const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, float* __restrict__ part,
                                                                long long R, int C) {
    extern __shared__ float smem[];
    const int CG = C >> 3, RL = THREADS / CG;
    const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
    float mu[8], rs[8], sg[8], sgx[8];
    loadf8(mean + cg * 8, mu);
    loadf8(rstd + cg * 8, rs);
#pragma unroll
    for (int k = 0; k < 8; ++k) { sg[k] = 0.f; sgx[k] = 0.f; }
    const long long G = (long long)gridDim.x * RL;
    for (long long r0 = (long long)blockIdx.x * RL + rl; r0 < R; r0 += 2 * G) {   // 2 rows x 3 tensors in flight
        float g[2][8], xv[2][8], yv[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long r = r0 + u * G;
            if (r < R) {
                load8(dy + r * C + cg * 8, g[u]);
                load8(x + r * C + cg * 8, xv[u]);
                if (RELU) load8(y + r * C + cg * 8, yv[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (r0 + u * G < R) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float gg = (RELU && !(yv[u][k] > 0.f)) ? 0.f : g[u][k];
                    sg[k] += gg;
                    sgx[k] = fmaf(gg, (xv[u][k] - mu[k]) * rs[k], sgx[k]);
                }
            }
        }
    }
    fold_and_write(sg, sgx, smem, part, C, cg, rl, RL);
}

// per channel: dgamma, dbeta (accumulated into the flat fp32 grad buffer) and the coefficients of
// dx = c0 * g + c1 * x + c2  with  c0 = gamma*rstd, c1 = -gamma*rstd^2*mean(g*xhat), c2 = -c0*mean(g) - c1*mean
__global__ void __launch_bounds__(FIN_THREADS) bn_bwd_finalize_kernel(
                                       const float* __restrict__ part, int nparts, const float* __restrict__ gamma,
                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef,
                                       long long R, int C, int accumulate) {
    float sg, sgx;
    const int c = fold_parts(part, nparts, C, sg, sgx);
    if (c < 0) return;
    dgamma[c] = accumulate ? dgamma[c] + sgx : sgx;
    dbeta[c] = accumulate ? dbeta[c] + sg : sg;
    const float invR = 1.f / (float)R;
    const float c0 = gamma[c] * rstd[c];
    const float c1 = -c0 * rstd[c] * sgx * invR;
    coef[c] = c0;
    coef[C + c] = c1;
    coef[2 * C + c] = -c0 * sg * invR - c1 * mean[c];
}

template <bool RELU, bool RES>
__global__ void __launch_bounds__(THREADS) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                                               const __nv_bfloat16* __restrict__ x, const float* __restrict__ coef,
                                                               __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dres,
                                                               long long R, int C) {
    const int CG = C >> 3, RL = THREADS / CG;
    const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
    float c0[8], c1[8], c2[8];
    loadf8(coef + cg * 8, c0);
    loadf8(coef + C + cg * 8, c1);
    loadf8(coef + 2 * C + cg * 8, c2);
    const long long G = (long long)gridDim.x * RL;
    for (long long r0 = (long long)blockIdx.x * RL + rl; r0 < R; r0 += 2 * G) {
        float g[2][8], xv[2][8], yv[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long r = r0 + u * G;
            if (r < R) {
                load8(dy + r * C + cg * 8, g[u]);
                load8(x + r * C + cg * 8, xv[u]);
                if (RELU) load8(y + r * C + cg * 8, yv[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long r = r0 + u * G;
            if (r < R) {
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (RELU && !(yv[u][k] > 0.f)) g[u][k] = 0.f;
                    o[k] = fmaf(c0[k], g[u][k], fmaf(c1[k], xv[u][k], c2[k]));
                }
                if (RES) store8(dres + r * C + cg * 8, g[u]);
                store8(dx + r * C + cg * 8, o);
            }
        }
    }
}

static inline bool shape_ok(int C) {
    const int cg = C >> 3;
    return C % 8 == 0 && cg >= 1 && cg <= THREADS && (cg & (cg - 1)) == 0;
}
static inline int grid_for(long long R, int C) {
    const int RL = THREADS / (C >> 3);
    long long g = (R + RL - 1) / RL;
    return (int)(g < 1 ? 1 : (g > MAX_PARTS ? MAX_PARTS : g));
}

// element-wise passes keep no partials: use every resident CTA slot (8 CTAs/SM x 148 SMs)
static inline int apply_grid(long long R, int C) {
    const int RL = THREADS / (C >> 3);
    long long g = (R + RL - 1) / RL;
    return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

}  // namespace bn

// scratch: >= MAX_PARTS*2*C floats (partials) ; stats: mean[C] rstd[C] scale[C] bias[C]
extern "C" int v6_bn_fwd(const void* x, const void* res, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, void* y, float* mean, float* rstd, float* scale_bias, float* scratch,
                         long long R, int C, float eps, float momentum, int relu, cudaStream_t s) {
    using namespace bn;
    if (!shape_ok(C)) return (int)cudaErrorInvalidValue;
    const int grid = grid_for(R, C);
    const int RL = THREADS / (C >> 3);
    const size_t smem = (size_t)RL * C * 2 * sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(bn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bn_stats_kernel<<<grid, THREADS, smem, s>>>((const __nv_bfloat16*)x, scratch, R, C);
    bn_fwd_finalize_kernel<<<(C + 31) / 32, FIN_THREADS, 0, s>>>(scratch, grid, (const __nv_bfloat16*)x, gamma, beta, running_mean,
                                                           running_var, mean, rstd, scale_bias, scale_bias + C, R, C, eps, momentum);
    const int ag = apply_grid(R, C);
    const __nv_bfloat16* xx = (const __nv_bfloat16*)x;
    const __nv_bfloat16* rr = (const __nv_bfloat16*)res;
    __nv_bfloat16* yy = (__nv_bfloat16*)y;
    if (relu) {
        if (res) bn_apply_kernel<true, true><<<ag, THREADS, 0, s>>>(xx, rr, scale_bias, scale_bias + C, yy, R, C);
        else bn_apply_kernel<true, false><<<ag, THREADS, 0, s>>>(xx, rr, scale_bias, scale_bias + C, yy, R, C);
    } else {
        if (res) bn_apply_kernel<false, true><<<ag, THREADS, 0, s>>>(xx, rr, scale_bias, scale_bias + C, yy, R, C);
        else bn_apply_kernel<false, false><<<ag, THREADS, 0, s>>>(xx, rr, scale_bias, scale_bias + C, yy, R, C);
    }
    V6_CHECK_LAUNCH();
    return 0;
}

// inference / eval: y = relu(x*scale + bias + res) with caller-provided per-channel affine
extern "C" int v6_bn_apply(const void* x, const void* res, const float* scale, const float* bias, void* y, long long R, int C,
                           int relu, cudaStream_t s) {
    using namespace bn;
    if (!shape_ok(C)) return (int)cudaErrorInvalidValue;
    const int ag = apply_grid(R, C);
    const __nv_bfloat16* xx = (const __nv_bfloat16*)x;
    const __nv_bfloat16* rr = (const __nv_bfloat16*)res;
    __nv_bfloat16* yy = (__nv_bfloat16*)y;
    if (relu) {
        if (res) bn_apply_kernel<true, true><<<ag, THREADS, 0, s>>>(xx, rr, scale, bias, yy, R, C);
        else bn_apply_kernel<true, false><<<ag, THREADS, 0, s>>>(xx, rr, scale, bias, yy, R, C);
    } else {
        if (res) bn_apply_kernel<false, true><<<ag, THREADS, 0, s>>>(xx, rr, scale, bias, yy, R, C);
        else bn_apply_kernel<false, false><<<ag, THREADS, 0, s>>>(xx, rr, scale, bias, yy, R, C);
    }
    V6_CHECK_LAUNCH();
    return 0;
}

// coef scratch: 3*C floats. dres may be null (no residual branch).
extern "C" int v6_bn_bwd(const void* dy, const void* y, const void* x, const float* gamma, const float* mean, const float* rstd,
                         void* dx, void* dres, float* dgamma, float* dbeta, float* coef, float* scratch, long long R, int C,
                         int relu, int accumulate, cudaStream_t s) {
    using namespace bn;
    if (!shape_ok(C)) return (int)cudaErrorInvalidValue;
    const int grid = grid_for(R, C);
    const int RL = THREADS / (C >> 3);
    const size_t smem = (size_t)RL * C * 2 * sizeof(float);
    const __nv_bfloat16* dyy = (const __nv_bfloat16*)dy;
    const __nv_bfloat16* yy = (const __nv_bfloat16*)y;
    const __nv_bfloat16* xx = (const __nv_bfloat16*)x;
    if (relu) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(bn_bwd_reduce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        bn_bwd_reduce_kernel<true><<<grid, THREADS, smem, s>>>(dyy, yy, xx, mean, rstd, scratch, R, C);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(bn_bwd_reduce_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        bn_bwd_reduce_kernel<false><<<grid, THREADS, smem, s>>>(dyy, yy, xx, mean, rstd, scratch, R, C);
    }
    bn_bwd_finalize_kernel<<<(C + 31) / 32, FIN_THREADS, 0, s>>>(scratch, grid, gamma, mean, rstd, dgamma, dbeta, coef, R, C, accumulate);
    const int ag = apply_grid(R, C);
    __nv_bfloat16* dxx = (__nv_bfloat16*)dx;
    __nv_bfloat16* drr = (__nv_bfloat16*)dres;
    if (relu) {
        if (dres) bn_bwd_apply_kernel<true, true><<<ag, THREADS, 0, s>>>(dyy, yy, xx, coef, dxx, drr, R, C);
        else bn_bwd_apply_kernel<true, false><<<ag, THREADS, 0, s>>>(dyy, yy, xx, coef, dxx, drr, R, C);
    } else {
        if (dres) bn_bwd_apply_kernel<false, true><<<ag, THREADS, 0, s>>>(dyy, yy, xx, coef, dxx, drr, R, C);
        else bn_bwd_apply_kernel<false, false><<<ag, THREADS, 0, s>>>(dyy, yy, xx, coef, dxx, drr, R, C);
    }
    V6_CHECK_LAUNCH();
    return 0;
}
