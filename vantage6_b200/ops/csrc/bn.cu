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
constexpr int MAX_PARTS = 148 * 4;

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
    for (long long r = (long long)blockIdx.x * RL + rl; r < R; r += (long long)gridDim.x * RL) {
        float v[8];
        load8(x + r * C + cg * 8, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float d = v[k] - shift[k]; s1[k] += d; s2[k] = fmaf(d, d, s2[k]); }
    }
    fold_and_write(s1, s2, smem, part, C, cg, rl, RL);
}

// one thread per channel: mean / rstd, running stats, and the affine (scale, bias) of the apply pass
__global__ void bn_fwd_finalize_kernel(const float* __restrict__ part, int nparts, const __nv_bfloat16* __restrict__ x,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float* __restrict__ running_mean, float* __restrict__ running_var,
                                       float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                       float* __restrict__ scale_out, float* __restrict__ bias_out,
                                       long long R, int C, float eps, float momentum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f, b = 0.f;
    for (int p = 0; p < nparts; ++p) { a += part[(size_t)p * 2 * C + c]; b += part[(size_t)p * 2 * C + C + c]; }
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
    for (long long r = (long long)blockIdx.x * RL + rl; r < R; r += (long long)gridDim.x * RL) {
        float v[8];
        load8(x + r * C + cg * 8, v);
        if (RES) {
            float q[8];
            load8(res + r * C + cg * 8, q);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], sc[k], bi[k]) + q[k];
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], sc[k], bi[k]);
        }
        if (RELU) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
        }
        store8(y + r * C + cg * 8, v);
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
    for (long long r = (long long)blockIdx.x * RL + rl; r < R; r += (long long)gridDim.x * RL) {
        float g[8], xv[8];
        load8(dy + r * C + cg * 8, g);
        load8(x + r * C + cg * 8, xv);
        if (RELU) {
            float yv[8];
            load8(y + r * C + cg * 8, yv);
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] = yv[k] > 0.f ? g[k] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { sg[k] += g[k]; sgx[k] = fmaf(g[k], (xv[k] - mu[k]) * rs[k], sgx[k]); }
    }
    fold_and_write(sg, sgx, smem, part, C, cg, rl, RL);
}

// per channel: dgamma, dbeta (accumulated into the flat fp32 grad buffer) and the coefficients of
// dx = c0 * g + c1 * x + c2  with  c0 = gamma*rstd, c1 = -gamma*rstd^2*mean(g*xhat), c2 = -c0*mean(g) - c1*mean
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ part, int nparts, const float* __restrict__ gamma,
                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef,
                                       long long R, int C, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float sg = 0.f, sgx = 0.f;
    for (int p = 0; p < nparts; ++p) { sg += part[(size_t)p * 2 * C + c]; sgx += part[(size_t)p * 2 * C + C + c]; }
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
    for (long long r = (long long)blockIdx.x * RL + rl; r < R; r += (long long)gridDim.x * RL) {
        float g[8], xv[8], o[8];
        load8(dy + r * C + cg * 8, g);
        load8(x + r * C + cg * 8, xv);
        if (RELU) {
            float yv[8];
            load8(y + r * C + cg * 8, yv);
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] = yv[k] > 0.f ? g[k] : 0.f;
        }
        if (RES) store8(dres + r * C + cg * 8, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf(c0[k], g[k], fmaf(c1[k], xv[k], c2[k]));
        store8(dx + r * C + cg * 8, o);
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
    bn_fwd_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(scratch, grid, (const __nv_bfloat16*)x, gamma, beta, running_mean,
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
    bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(scratch, grid, gamma, mean, rstd, dgamma, dbeta, coef, R, C, accumulate);
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
