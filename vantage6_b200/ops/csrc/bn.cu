// Fused BatchNorm(+residual add)(+ReLU) forward / backward for NHWC bf16 activations (ResNet-50).
//
// The first profile of the flagship local step (profiles/launches_resnet50_eager_r1.txt) showed
// PyTorch's native channels-last BatchNorm kernels + stand-alone ReLU / add kernels taking ~3/4 of
// the GPU time of a ResNet-50 step on B200 -- the convolutions (cuDNN/CUTLASS sm100 kernels) only
// ~15%.  These kernels make that part memory-bound at HBM speed and remove the separate ReLU /
// residual-add passes:
//
//   forward :  stats + per-channel finalize (1 read of x, one launch)  ->  y = relu(x*scale + bias + res)
//   backward:  reduce + finalize (dy, y, x -> dgamma, dbeta, coefficients)  ->  dx (and dres) in one pass
//
// x viewed as [R = N*H*W, C], C % 8 == 0 and C/8 a power of two <= 256 (64..2048 in ResNet-50).
// Element-wise passes: threadIdx % (C/8) owns 8 consecutive channels (one 16 B vector), threadIdx / (C/8)
// is a row lane; CTAs stride over rows.  Reductions: see "in-kernel tree reduction" below; per-channel
// sums are accumulated in registers across rows (shifted by x[0,c] to avoid cancellation in
// E[x^2]-E[x]^2) -- no floating-point atomics, deterministic.
#include "common.cuh"
#include "api.h"
#include "tree_reduce.cuh"

namespace bn {

using namespace tree;          // THREADS, Red, slice_reduce, reduce_grid, wave_ctas (tree_reduce.cuh)

// Programmatic dependent launch (PDL): the element-wise pass of a BN layer is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs are scheduled while the reduction kernel's tree
// tail is still running and block in griddepcontrol.wait until that grid has completed and flushed its writes.
// Without the attribute both instructions are no-ops.
V6_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
V6_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

V6_DEVINL void load8(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
V6_DEVINL void store8(__nv_bfloat16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                              pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
V6_DEVINL uint4 ldg_nc_v4(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
V6_DEVINL void unpack8(const uint4& t, float (&v)[8]) {
    float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
V6_DEVINL void acc_stats(const uint4& raw, const float (&shift)[8], float (&s1)[8], float (&s2)[8]) {
    float v[8];
    unpack8(raw, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float d = v[k] - shift[k]; s1[k] += d; s2[k] = fmaf(d, d, s2[k]); }
}
V6_DEVINL void loadf8(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ---------------------------------------------------------------------------------- forward
// stats + finalize: per-channel mean / rstd / (scale, bias) of y = x*scale + bias, running-stat update.
constexpr int STATS_UNROLL = 4;
__global__ void __launch_bounds__(THREADS, 4) bn_stats_kernel(const __nv_bfloat16* __restrict__ x, Red rd,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float* __restrict__ running_mean, float* __restrict__ running_var,
                                       long long* __restrict__ num_batches_tracked,
                                       float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                       float* __restrict__ scale_out, float* __restrict__ bias_out,
                                       long long R, int C, float eps, float momentum) {
    __shared__ float tot[128];
    const int SW = C < 64 ? C : 64, CGS = SW >> 3, RL = THREADS / CGS;
    const int cg = threadIdx.x % CGS, rl = threadIdx.x / CGS;
    const __nv_bfloat16* xc = x + blockIdx.y * SW + cg * 8;
    float shift[8], s1[8], s2[8];
    load8(xc, shift);                                           // row 0 as the per-channel shift
#pragma unroll
    for (int k = 0; k < 8; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
    const long long G = (long long)gridDim.x * RL;
    long long r = (long long)blockIdx.x * RL + rl;
    for (; r + (STATS_UNROLL - 1) * G < R; r += STATS_UNROLL * G) {            // UNROLL independent 16 B loads in flight
        uint4 raw[STATS_UNROLL];
#pragma unroll
        for (int u = 0; u < STATS_UNROLL; ++u) raw[u] = ldg_nc_v4(xc + (r + u * G) * C);
#pragma unroll
        for (int u = 0; u < STATS_UNROLL; ++u) acc_stats(raw[u], shift, s1, s2);
    }
    for (; r < R; r += G) acc_stats(ldg_nc_v4(xc + r * C), shift, s1, s2);
    pdl_launch_dependents();                                   // let the apply kernel's CTAs get scheduled
    if (!slice_reduce(s1, s2, rd, SW, tot)) return;
    if (threadIdx.x < SW) {
        const int c = blockIdx.y * SW + threadIdx.x;
        const float sh = __bfloat162float(x[c]);
        const float invR = 1.f / (float)R;
        const float dm = tot[threadIdx.x] * invR;
        const float mean = sh + dm;
        const float var = fmaxf(tot[SW + threadIdx.x] * invR - dm * dm, 0.f);
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
        if (num_batches_tracked && c == 0) *num_batches_tracked += 1;
    }
}

// Element-wise passes: U rows per thread are requested (raw 16 B vectors, kept packed until used) before the first one
// is consumed, and the grid is exactly one wave of resident CTAs (apply_grid) -- HBM3e wants >= ~64 KB of loads in flight
// per SM and a grid-stride loop has no tail.  SMEM_COEF keeps the per-channel coefficients in shared memory instead of
// 16-24 registers so that 4 CTAs/SM stay resident with U = 4.
template <bool RELU, bool RES, int U, int MINB, bool SMEM_COEF>
__global__ void __launch_bounds__(THREADS, MINB) bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                                                                 const float* __restrict__ scale, const float* __restrict__ bias,
                                                                 __nv_bfloat16* __restrict__ y, unsigned char* __restrict__ mask,
                                                                 long long R, int C) {
    extern __shared__ float s_coef[];                           // SMEM_COEF: scale[C] | bias[C]
    const int CG = C >> 3, RL = THREADS / CG;
    const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
    pdl_wait();                                                 // scale / bias come from the stats kernel
    float sc[8], bi[8];
    if (SMEM_COEF) {
        for (int i = threadIdx.x; i < C; i += THREADS) { s_coef[i] = scale[i]; s_coef[C + i] = bias[i]; }
        __syncthreads();
    } else {
        loadf8(scale + cg * 8, sc);
        loadf8(bias + cg * 8, bi);
    }
    const long long G = (long long)gridDim.x * RL;
    for (long long r0 = (long long)blockIdx.x * RL + rl; r0 < R; r0 += U * G) {
        uint4 xr[U], qr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * G;
            if (r < R) {
                xr[u] = ldg_nc_v4(x + r * C + cg * 8);
                if (RES) qr[u] = ldg_nc_v4(res + r * C + cg * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * G;
            if (r < R) {
                float v[8], q[8];
                unpack8(xr[u], v);
                if (RES) unpack8(qr[u], q);
                if (SMEM_COEF) { loadf8(s_coef + cg * 8, sc); loadf8(s_coef + C + cg * 8, bi); }
                unsigned bits = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float t = fmaf(v[k], sc[k], bi[k]);
                    if (RES) t += q[k];
                    if (RELU) bits |= (t > 0.f ? 1u : 0u) << k;
                    v[k] = RELU ? fmaxf(t, 0.f) : t;
                }
                store8(y + r * C + cg * 8, v);
                if (RELU && mask) mask[r * CG + cg] = (unsigned char)bits;      // 1 bit / element for the backward
            }
        }
    }
}

// ---------------------------------------------------------------------------------- backward
template <bool RELU>
V6_DEVINL void acc_bwd(const uint4& graw, const uint4& xraw, unsigned bits, const float (&mu)[8], const float (&rs)[8],
                       float (&sg)[8], float (&sgx)[8]) {
    float g[8], xv[8];
    unpack8(graw, g);
    unpack8(xraw, xv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float gg = (RELU && !((bits >> k) & 1u)) ? 0.f : g[k];
        sg[k] += gg;
        sgx[k] = fmaf(gg, (xv[k] - mu[k]) * rs[k], sgx[k]);
    }
}

// reduce + finalize: per channel dgamma, dbeta (optionally accumulated into the flat fp32 grad buffer) and
// the coefficients of dx = c0 * g + c1 * x + c2 with
//   c0 = gamma*rstd, c1 = -gamma*rstd^2*mean(g*xhat), c2 = -c0*mean(g) - c1*mean
constexpr int RED_UNROLL = 4;
template <bool RELU>
__global__ void __launch_bounds__(THREADS, 3) bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const unsigned char* __restrict__ mask,
                                       const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                       const float* __restrict__ mean, const float* __restrict__ rstd, Red rd,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef,
                                       long long R, int C, int accumulate) {
    __shared__ float tot[128];
    const int SW = C < 64 ? C : 64, CGS = SW >> 3, RL = THREADS / CGS;
    const int cg = threadIdx.x % CGS, rl = threadIdx.x / CGS;
    const size_t c0 = (size_t)blockIdx.y * SW + cg * 8;
    pdl_wait();                                                 // dy comes from the kernel before (programmatic dependent launch)
    const long long CGT = C >> 3, cgt = (long long)blockIdx.y * CGS + cg;       // mask: [R][C/8] bytes
    float mu[8], rs[8], sg[8], sgx[8];
    loadf8(mean + c0, mu);
    loadf8(rstd + c0, rs);
#pragma unroll
    for (int k = 0; k < 8; ++k) { sg[k] = 0.f; sgx[k] = 0.f; }
    const long long G = (long long)gridDim.x * RL;
    long long r = (long long)blockIdx.x * RL + rl;
    for (; r + (RED_UNROLL - 1) * G < R; r += RED_UNROLL * G) {                // RED_UNROLL rows x 3 tensors in flight
        uint4 g[RED_UNROLL], xv[RED_UNROLL];
        unsigned mb[RED_UNROLL];
#pragma unroll
        for (int u = 0; u < RED_UNROLL; ++u) {
            const long long o = (r + u * G) * C + c0;
            g[u] = ldg_nc_v4(dy + o);
            xv[u] = ldg_nc_v4(x + o);
            mb[u] = RELU ? (unsigned)__ldg(mask + (r + u * G) * CGT + cgt) : 0xffu;
        }
#pragma unroll
        for (int u = 0; u < RED_UNROLL; ++u) acc_bwd<RELU>(g[u], xv[u], mb[u], mu, rs, sg, sgx);
    }
    for (; r < R; r += G) {
        const long long o = r * C + c0;
        unsigned mb = 0xffu;
        if (RELU) mb = __ldg(mask + r * CGT + cgt);
        acc_bwd<RELU>(ldg_nc_v4(dy + o), ldg_nc_v4(x + o), mb, mu, rs, sg, sgx);
    }
    pdl_launch_dependents();
    if (!slice_reduce(sg, sgx, rd, SW, tot)) return;
    if (threadIdx.x < SW) {
        const int c = blockIdx.y * SW + threadIdx.x;
        const float tg = tot[threadIdx.x], tgx = tot[SW + threadIdx.x];
        dgamma[c] = accumulate ? dgamma[c] + tgx : tgx;
        dbeta[c] = accumulate ? dbeta[c] + tg : tg;
        const float invR = 1.f / (float)R;
        const float k0 = gamma[c] * rstd[c];
        const float k1 = -k0 * rstd[c] * tgx * invR;
        coef[c] = k0;
        coef[C + c] = k1;
        coef[2 * C + c] = -k0 * tg * invR - k1 * mean[c];
    }
}

template <bool RELU, bool RES, int U, int MINB, bool SMEM_COEF>
__global__ void __launch_bounds__(THREADS, MINB) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const unsigned char* __restrict__ mask,
                                                                     const __nv_bfloat16* __restrict__ x, const float* __restrict__ coef,
                                                                     __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dres,
                                                                     long long R, int C) {
    extern __shared__ float s_coef[];                           // SMEM_COEF: c0[C] | c1[C] | c2[C]
    const int CG = C >> 3, RL = THREADS / CG;
    const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
    pdl_wait();                                                 // coefficients come from the reduce kernel
    float c0[8], c1[8], c2[8];
    if (SMEM_COEF) {
        for (int i = threadIdx.x; i < 3 * C; i += THREADS) s_coef[i] = coef[i];
        __syncthreads();
    } else {
        loadf8(coef + cg * 8, c0);
        loadf8(coef + C + cg * 8, c1);
        loadf8(coef + 2 * C + cg * 8, c2);
    }
    const long long G = (long long)gridDim.x * RL;
    for (long long r0 = (long long)blockIdx.x * RL + rl; r0 < R; r0 += U * G) {
        uint4 gr[U], xr[U];
        unsigned mb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * G;
            mb[u] = 0xffu;
            if (r < R) {
                gr[u] = ldg_nc_v4(dy + r * C + cg * 8);
                xr[u] = ldg_nc_v4(x + r * C + cg * 8);
                if (RELU) mb[u] = __ldg(mask + r * CG + cg);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * G;
            if (r < R) {
                float g[8], xv[8], o[8];
                unpack8(gr[u], g);
                unpack8(xr[u], xv);
                if (SMEM_COEF) { loadf8(s_coef + cg * 8, c0); loadf8(s_coef + C + cg * 8, c1); loadf8(s_coef + 2 * C + cg * 8, c2); }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (RELU && !((mb[u] >> k) & 1u)) g[k] = 0.f;
                    o[k] = fmaf(c0[k], g[k], fmaf(c1[k], xv[k], c2[k]));
                }
                if (RES) store8(dres + r * C + cg * 8, g);
                store8(dx + r * C + cg * 8, o);
            }
        }
    }
}

// ---------------------------------------------------------------------------------- stem: BN apply + ReLU + max-pool as one pass
// y1 = relu(x * scale + bias) at 112 x 112 is only ever consumed by the 3x3 / s2 / p1 max-pool: the forward below reads the raw
// convolution output and writes the pooled tensor + arg-max bytes (y1 and its ReLU mask never exist: 2 x 103 MB less per step at
// batch 64); the backward gathers the pooled gradient on the fly (<= 4 windows per pixel), recomputes relu' from x, and runs
// the usual reduce / apply pair -- the full-resolution gradient tensor that maxpool_bwd wrote and both BN passes re-read
// never exists either.
V6_DEVINL void pooled_grad(const __nv_bfloat16* __restrict__ dp, const unsigned char* __restrict__ idx, int n, int h, int w, int cg,
                           int C, int Ho, int Wo, float (&acc)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const int ho_lo = h >> 1, ho_hi = (h + 1) >> 1, wo_lo = w >> 1, wo_hi = (w + 1) >> 1;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int ho = a ? ho_hi : ho_lo;
        if ((a && ho_hi == ho_lo) || ho >= Ho) continue;
        const int dh = h - (2 * ho - 1);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int wo = b ? wo_hi : wo_lo;
            if ((b && wo_hi == wo_lo) || wo >= Wo) continue;
            const unsigned me = (unsigned)(dh * 3 + (w - (2 * wo - 1)));
            const size_t o = ((size_t)(n * Ho + ho) * Wo + wo) * C + cg * 8;
            const uint2 am = __ldg(reinterpret_cast<const uint2*>(idx + o));
            float gv[8];
            unpack8(ldg_nc_v4(dp + o), gv);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const unsigned ak = ((k < 4 ? am.x : am.y) >> (8 * (k & 3))) & 0xffu;
                if (ak == me) acc[k] += gv[k];
            }
        }
    }
}

// one thread = 8 channels of one pooled pixel
__global__ void __launch_bounds__(THREADS) bn_relu_maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                                                                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ p,
                                                                      unsigned char* __restrict__ idx, int N, int H, int W, int C, int Ho, int Wo) {
    const int CG = C >> 3;
    const int total = N * Ho * Wo * CG;
    pdl_wait();
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int cg = t % CG;
        int q = t / CG;
        const int wo = q % Wo; q /= Wo;
        const int ho = q % Ho;
        const int n = q / Ho;
        float sc[8], bi[8];
        loadf8(scale + cg * 8, sc);
        loadf8(bias + cg * 8, bi);
        const int h0 = 2 * ho - 1, w0 = 2 * wo - 1;
        uint4 raw[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const int h = h0 + j / 3, w = w0 + j % 3;
            if (h >= 0 && h < H && w >= 0 && w < W) raw[j] = ldg_nc_v4(x + ((size_t)(n * H + h) * W + w) * C + cg * 8);
        }
        float best[8];
        unsigned arg[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; arg[k] = 0; }
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const int h = h0 + j / 3, w = w0 + j % 3;
            if (h >= 0 && h < H && w >= 0 && w < W) {
                float v[8];
                unpack8(raw[j], v);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // the value the separate passes would have stored: relu(affine) rounded to bf16; first maximum in scan order
                    const float yv = __bfloat162float(__float2bfloat16_rn(fmaxf(fmaf(v[k], sc[k], bi[k]), 0.f)));
                    if (yv > best[k]) { best[k] = yv; arg[k] = j; }
                }
            }
        }
        const size_t o = ((size_t)(n * Ho + ho) * Wo + wo) * C + cg * 8;
        store8(p + o, best);
        *reinterpret_cast<uint2*>(idx + o) = make_uint2(arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24),
                                                         arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24));
    }
}

// reduce + finalize of the BN backward with g = pooled gradient . relu'(x * scale + bias)
__global__ void __launch_bounds__(THREADS, 3) bn_pool_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dp, const unsigned char* __restrict__ idx,
                                       const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ bias,
                                       const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd, Red rd,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef,
                                       int N, int H, int W, int C, int Ho, int Wo, int accumulate) {
    __shared__ float tot[128];
    const int SW = 64, CGS = 8, RL = THREADS / CGS;
    const int cgl = threadIdx.x % CGS, rl = threadIdx.x / CGS;
    const int cg = blockIdx.y * CGS + cgl;                    // channel group of the whole tensor
    const size_t c0 = (size_t)cg * 8;
    pdl_wait();
    float mu[8], rs[8], sc[8], bi[8], sg[8], sgx[8];
    loadf8(mean + c0, mu); loadf8(rstd + c0, rs); loadf8(scale + c0, sc); loadf8(bias + c0, bi);
#pragma unroll
    for (int k = 0; k < 8; ++k) { sg[k] = 0.f; sgx[k] = 0.f; }
    const int R = N * H * W, G = gridDim.x * RL;
    for (int r = blockIdx.x * RL + rl; r < R; r += G) {
        const int w = r % W, q = r / W, h = q % H, n = q / H;
        float xv[8], g[8];
        unpack8(ldg_nc_v4(x + (size_t)r * C + c0), xv);
        pooled_grad(dp, idx, n, h, w, cg, C, Ho, Wo, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float gg = fmaf(xv[k], sc[k], bi[k]) > 0.f ? g[k] : 0.f;
            sg[k] += gg;
            sgx[k] = fmaf(gg, (xv[k] - mu[k]) * rs[k], sgx[k]);
        }
    }
    pdl_launch_dependents();
    if (!slice_reduce(sg, sgx, rd, SW, tot)) return;
    if (threadIdx.x < SW) {
        const int c = blockIdx.y * SW + threadIdx.x;
        const float tg = tot[threadIdx.x], tgx = tot[SW + threadIdx.x];
        dgamma[c] = accumulate ? dgamma[c] + tgx : tgx;
        dbeta[c] = accumulate ? dbeta[c] + tg : tg;
        const float invR = 1.f / (float)R;
        const float k0 = gamma[c] * rstd[c];
        const float k1 = -k0 * rstd[c] * tgx * invR;
        coef[c] = k0;
        coef[C + c] = k1;
        coef[2 * C + c] = -k0 * tg * invR - k1 * mean[c];
    }
}

__global__ void __launch_bounds__(THREADS, 3) bn_pool_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dp, const unsigned char* __restrict__ idx,
                                      const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ bias,
                                      const float* __restrict__ coef, __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int Ho, int Wo) {
    const int CG = C >> 3, RL = THREADS / CG;
    const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
    pdl_wait();
    float c0[8], c1[8], c2[8], sc[8], bi[8];
    loadf8(coef + cg * 8, c0); loadf8(coef + C + cg * 8, c1); loadf8(coef + 2 * C + cg * 8, c2);
    loadf8(scale + cg * 8, sc); loadf8(bias + cg * 8, bi);
    const int R = N * H * W, G = gridDim.x * RL;
    for (int r = blockIdx.x * RL + rl; r < R; r += G) {
        const int w = r % W, q = r / W, h = q % H, n = q / H;
        float xv[8], g[8], o[8];
        unpack8(ldg_nc_v4(x + (size_t)r * C + cg * 8), xv);
        pooled_grad(dp, idx, n, h, w, cg, C, Ho, Wo, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float gg = fmaf(xv[k], sc[k], bi[k]) > 0.f ? g[k] : 0.f;
            o[k] = fmaf(c0[k], gg, fmaf(c1[k], xv[k], c2[k]));
        }
        store8(dx + (size_t)r * C + cg * 8, o);
    }
}

static inline bool shape_ok(int C) {
    const int cg = C >> 3;
    return C % 8 == 0 && cg >= 1 && cg <= THREADS && (cg & (cg - 1)) == 0;     // 8, 16, ..., 2048
}
// element-wise passes keep no partials: exactly one wave of resident CTAs (occupancy queried once per instantiation),
// fewer when the tensor has fewer row groups than that
template <typename K>
static int apply_grid(K kernel, size_t smem, long long R, int C) {
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, THREADS, smem) != cudaSuccess || occ < 1) occ = 1;
    const int RL = THREADS / (C >> 3);
    const long long g = (R + RL - 1) / RL, cap = (long long)occ * sms;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// launch `kernel` as a programmatic dependent of the previous kernel in the stream (V6B200_PDL=0 or pdl=false: plain launch)
template <typename... KArgs, typename... Args>
static void launch_dependent(void (*kernel)(KArgs...), int grid, size_t smem, bool dependent, cudaStream_t s, Args... args) {
    static const bool pdl = [] { const char* e = getenv("V6B200_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && dependent) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// same, 2-D grid, no dynamic shared memory (the reduce kernels)
template <typename... KArgs, typename... Args>
static void launch_dependent_grid(void (*kernel)(KArgs...), dim3 grid, cudaStream_t s, Args... args) {
    static const bool pdl = [] { const char* e = getenv("V6B200_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(THREADS);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// tuning variant of the element-wise passes (V6B200_BN_CFG; scripts/bn_bench.py sweeps it, profiles/bn_bench_r2f.jsonl):
//   0  U = 4 rows in flight, coefficients in shared memory, 4 CTAs/SM      2.93 ms of BN passes per ResNet-50 step
//   1  U = 2, coefficients in registers, 3 CTAs/SM                          2.52 ms   (default; 5.5-6.0 TB/s on the
//   2  U = 4, coefficients in registers, 2 CTAs/SM                          2.51 ms    100 MB layers = 0.84-0.91 of copy peak)
static int bn_cfg() {
    static const int v = [] { const char* e = getenv("V6B200_BN_CFG"); return e ? atoi(e) : 1; }();
    return v;
}

template <bool RELU, bool RES>
static void launch_apply(bool dependent, cudaStream_t s, const __nv_bfloat16* x, const __nv_bfloat16* res, const float* scale, const float* bias,
                         __nv_bfloat16* y, unsigned char* mask, long long R, int C) {
    const int cfg = bn_cfg();
    if (cfg == 1) {
        auto k = bn_apply_kernel<RELU, RES, 2, 3, false>;
        launch_dependent(k, apply_grid(k, 0, R, C), 0, dependent, s, x, res, scale, bias, y, mask, R, C);
    } else if (cfg == 2) {
        auto k = bn_apply_kernel<RELU, RES, 4, 2, false>;
        launch_dependent(k, apply_grid(k, 0, R, C), 0, dependent, s, x, res, scale, bias, y, mask, R, C);
    } else {
        auto k = bn_apply_kernel<RELU, RES, 4, 4, true>;
        const size_t sm = (size_t)2 * C * sizeof(float);
        launch_dependent(k, apply_grid(k, sm, R, C), sm, dependent, s, x, res, scale, bias, y, mask, R, C);
    }
}

template <bool RELU, bool RES>
static void launch_bwd_apply(cudaStream_t s, const __nv_bfloat16* dy, const unsigned char* mask, const __nv_bfloat16* x, const float* coef,
                             __nv_bfloat16* dx, __nv_bfloat16* dres, long long R, int C) {
    const int cfg = bn_cfg();
    if (cfg == 1) {
        auto k = bn_bwd_apply_kernel<RELU, RES, 2, 3, false>;
        launch_dependent(k, apply_grid(k, 0, R, C), 0, true, s, dy, mask, x, coef, dx, dres, R, C);
    } else if (cfg == 2) {
        auto k = bn_bwd_apply_kernel<RELU, RES, 4, 2, false>;
        launch_dependent(k, apply_grid(k, 0, R, C), 0, true, s, dy, mask, x, coef, dx, dres, R, C);
    } else {
        auto k = bn_bwd_apply_kernel<RELU, RES, 4, 4, true>;
        const size_t sm = (size_t)3 * C * sizeof(float);
        launch_dependent(k, apply_grid(k, sm, R, C), sm, true, s, dy, mask, x, coef, dx, dres, R, C);
    }
}

}  // namespace bn

// scratch: v6_bn_scratch_floats() floats, zero-initialised once (partials + self-resetting counters), shared by
// every launch on one stream.  stats: mean[C] rstd[C] scale_bias[2C].
extern "C" long long v6_bn_scratch_floats() { return (long long)tree::SCR_FLOATS; }

extern "C" int v6_bn_fwd(const void* x, const void* res, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, long long* num_batches_tracked, void* y, void* relu_mask, float* mean, float* rstd,
                         float* scale_bias, float* scratch, long long R, int C, float eps, float momentum, int relu,
                         cudaStream_t s) {
    using namespace bn;
    if (!shape_ok(C)) return (int)cudaErrorInvalidValue;
    static int wave = 0;
    bn_stats_kernel<<<reduce_grid(R, C, wave_ctas(bn_stats_kernel, wave)), THREADS, 0, s>>>((const __nv_bfloat16*)x, make_red(scratch), gamma, beta,
                                                                  running_mean, running_var, num_batches_tracked, mean, rstd,
                                                                  scale_bias, scale_bias + C, R, C, eps, momentum);
    const __nv_bfloat16* xx = (const __nv_bfloat16*)x;
    const __nv_bfloat16* rr = (const __nv_bfloat16*)res;
    __nv_bfloat16* yy = (__nv_bfloat16*)y;
    unsigned char* mk = (unsigned char*)relu_mask;
    if (relu) {
        if (res) launch_apply<true, true>(true, s, xx, rr, scale_bias, scale_bias + C, yy, mk, R, C);
        else launch_apply<true, false>(true, s, xx, rr, scale_bias, scale_bias + C, yy, mk, R, C);
    } else {
        if (res) launch_apply<false, true>(true, s, xx, rr, scale_bias, scale_bias + C, yy, nullptr, R, C);
        else launch_apply<false, false>(true, s, xx, rr, scale_bias, scale_bias + C, yy, nullptr, R, C);
    }
    V6_CHECK_LAUNCH();
    return 0;
}

// inference / eval: y = relu(x*scale + bias + res) with caller-provided per-channel affine
extern "C" int v6_bn_apply(const void* x, const void* res, const float* scale, const float* bias, void* y, void* relu_mask,
                           long long R, int C, int relu, cudaStream_t s) {
    using namespace bn;
    if (!shape_ok(C)) return (int)cudaErrorInvalidValue;
    const __nv_bfloat16* xx = (const __nv_bfloat16*)x;
    const __nv_bfloat16* rr = (const __nv_bfloat16*)res;
    __nv_bfloat16* yy = (__nv_bfloat16*)y;
    unsigned char* mk = (unsigned char*)relu_mask;      // training with statistics from the convolution epilogue: 1 bit / element
    // (x and the statistics come from the convolution kernel before it, which releases its dependents at its start: this
    // kernel's CTAs are resident and past their prologue when that grid drains, and wait for it in griddepcontrol.wait)
    if (relu) {
        if (res) launch_apply<true, true>(true, s, xx, rr, scale, bias, yy, mk, R, C);
        else launch_apply<true, false>(true, s, xx, rr, scale, bias, yy, mk, R, C);
    } else {
        if (res) launch_apply<false, true>(true, s, xx, rr, scale, bias, yy, nullptr, R, C);
        else launch_apply<false, false>(true, s, xx, rr, scale, bias, yy, nullptr, R, C);
    }
    V6_CHECK_LAUNCH();
    return 0;
}

// coef scratch: 3*C floats. dres may be null (no residual branch).
extern "C" int v6_bn_bwd(const void* dy, const void* relu_mask, const void* x, const float* gamma, const float* mean, const float* rstd,
                         void* dx, void* dres, float* dgamma, float* dbeta, float* coef, float* scratch, long long R, int C,
                         int relu, int accumulate, cudaStream_t s) {
    using namespace bn;
    if (!shape_ok(C)) return (int)cudaErrorInvalidValue;
    const __nv_bfloat16* dyy = (const __nv_bfloat16*)dy;
    const unsigned char* yy = (const unsigned char*)relu_mask;     // 1 bit / element, written by the forward apply pass
    const __nv_bfloat16* xx = (const __nv_bfloat16*)x;
    if (relu && !relu_mask) return (int)cudaErrorInvalidValue;
    static int wave_relu = 0, wave_lin = 0;
    const dim3 rg = reduce_grid(R, C, relu ? wave_ctas(bn_bwd_reduce_kernel<true>, wave_relu) : wave_ctas(bn_bwd_reduce_kernel<false>, wave_lin));
    if (relu) launch_dependent_grid(bn_bwd_reduce_kernel<true>, rg, s, dyy, yy, xx, gamma, mean, rstd, make_red(scratch), dgamma, dbeta, coef, R, C, accumulate);
    else launch_dependent_grid(bn_bwd_reduce_kernel<false>, rg, s, dyy, yy, xx, gamma, mean, rstd, make_red(scratch), dgamma, dbeta, coef, R, C, accumulate);
    __nv_bfloat16* dxx = (__nv_bfloat16*)dx;
    __nv_bfloat16* drr = (__nv_bfloat16*)dres;
    if (relu) {
        if (dres) launch_bwd_apply<true, true>(s, dyy, yy, xx, coef, dxx, drr, R, C);
        else launch_bwd_apply<true, false>(s, dyy, yy, xx, coef, dxx, drr, R, C);
    } else {
        if (dres) launch_bwd_apply<false, true>(s, dyy, yy, xx, coef, dxx, drr, R, C);
        else launch_bwd_apply<false, false>(s, dyy, yy, xx, coef, dxx, drr, R, C);
    }
    V6_CHECK_LAUNCH();
    return 0;
}

// element-wise half of the backward alone: the reduction (dgamma, dbeta, coef) was done in the epilogue of the data-gradient
// kernel that produced dy (igemm.cu EPI_RED).  coef: [3C] = c0 | c1 | c2.
extern "C" int v6_bn_bwd_apply(const void* dy, const void* relu_mask, const void* x, const float* coef, void* dx, void* dres, long long R,
                               int C, int relu, cudaStream_t s) {
    using namespace bn;
    if (!shape_ok(C) || (relu && !relu_mask)) return (int)cudaErrorInvalidValue;
    const __nv_bfloat16* dyy = (const __nv_bfloat16*)dy;
    const unsigned char* yy = (const unsigned char*)relu_mask;
    const __nv_bfloat16* xx = (const __nv_bfloat16*)x;
    __nv_bfloat16* dxx = (__nv_bfloat16*)dx;
    __nv_bfloat16* drr = (__nv_bfloat16*)dres;
    if (relu) {
        if (dres) launch_bwd_apply<true, true>(s, dyy, yy, xx, coef, dxx, drr, R, C);
        else launch_bwd_apply<true, false>(s, dyy, yy, xx, coef, dxx, drr, R, C);
    } else {
        if (dres) launch_bwd_apply<false, true>(s, dyy, yy, xx, coef, dxx, drr, R, C);
        else launch_bwd_apply<false, false>(s, dyy, yy, xx, coef, dxx, drr, R, C);
    }
    V6_CHECK_LAUNCH();
    return 0;
}

// stem: y = maxpool3x3s2p1(relu(x * scale + bias)), idx = arg-max bytes.  x: [N, H, W, C] NHWC bf16, C % 64 == 0, C <= 2048 / CG <= 256.
extern "C" int v6_bn_pool_fwd(const void* x, const float* scale, const float* bias, void* p, void* idx, int N, int H, int W, int C,
                              cudaStream_t s) {
    using namespace bn;
    if (C % 64 != 0 || N < 1 || H < 2 || W < 2 || (long long)N * H * W * C >= (1LL << 31)) return (int)cudaErrorInvalidValue;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long work = (long long)N * Ho * Wo * (C >> 3);
    long long g = (work + THREADS - 1) / THREADS;
    const int grid = (int)(g > 148 * 16 ? 148 * 16 : g);
    launch_dependent(bn_relu_maxpool_fwd_kernel, grid, 0, true, s, (const __nv_bfloat16*)x, scale, bias, (__nv_bfloat16*)p, (unsigned char*)idx,
                     N, H, W, C, Ho, Wo);
    V6_CHECK_LAUNCH();
    return 0;
}

extern "C" int v6_bn_pool_bwd(const void* dp, const void* idx, const void* x, const float* scale, const float* bias, const float* gamma,
                              const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta, float* coef, float* scratch,
                              int N, int H, int W, int C, int accumulate, cudaStream_t s) {
    using namespace bn;
    if (C % 64 != 0 || !shape_ok(C) || (long long)N * H * W * C >= (1LL << 31)) return (int)cudaErrorInvalidValue;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long R = (long long)N * H * W;
    static int wave = 0;
    const dim3 rg = reduce_grid(R, C, wave_ctas(bn_pool_bwd_reduce_kernel, wave));
    launch_dependent_grid(bn_pool_bwd_reduce_kernel, rg, s, (const __nv_bfloat16*)dp, (const unsigned char*)idx, (const __nv_bfloat16*)x, scale, bias,
                          gamma, mean, rstd, make_red(scratch), dgamma, dbeta, coef, N, H, W, C, Ho, Wo, accumulate);
    auto k = bn_pool_bwd_apply_kernel;
    launch_dependent(k, apply_grid(k, 0, R, C), 0, true, s, (const __nv_bfloat16*)dp, (const unsigned char*)idx, (const __nv_bfloat16*)x, scale, bias,
                     (const float*)coef, (__nv_bfloat16*)dx, N, H, W, C, Ho, Wo);
    V6_CHECK_LAUNCH();
    return 0;
}
