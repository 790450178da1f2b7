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

namespace bn {

constexpr int THREADS = 256;

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

// ------------------------------------------------------------------ in-kernel tree reduction
// v3: the separate finalize kernels are gone.  A reduce CTA covers a 64-channel slice (blockIdx.y) and
// a strided set of rows (blockIdx.x); its [2*SW] partial goes to global memory, then a two-level
// "last CTA to arrive folds" tree (groups of G1 row-CTAs, then the groups of the slice) produces the
// per-channel totals inside the same launch and the last CTA of each slice does the per-channel math.
// Every fold sums its inputs in a fixed order, so the result does not depend on which CTA happens to
// arrive last (deterministic, no floating-point atomics); the counters reset themselves.
// (v2 profile, profiles/launches_resnet50_fusedbn_r1.txt: finalize 12.5 us per launch, latency-bound,
// 15% of the forward; stats at 2 TB/s because only 2 CTAs/SM were resident.)
constexpr int G1 = 16;
constexpr int ONE_LEVEL_MAX = 128;
constexpr int MAX_CTAS = 148 * 8;                    // upper bound on reduce CTAs per launch
constexpr int L2_SLOTS = 256;                        // >= MAX_CTAS / G1 + slices
constexpr size_t SCR_L2 = (size_t)MAX_CTAS * 128;    // float offsets into the scratch buffer
constexpr size_t SCR_CNT1 = SCR_L2 + (size_t)L2_SLOTS * 128;
constexpr size_t SCR_CNT2 = SCR_CNT1 + L2_SLOTS;
constexpr size_t SCR_FLOATS = SCR_CNT2 + 64;

struct Red {
    float* l1;      // [slice][row-CTA][2*SW]
    float* l2;      // [slice][group][2*SW]
    int* cnt1;      // [slice][group]   (zero between launches)
    int* cnt2;      // [slice]
};
static inline Red make_red(float* scratch) {
    return Red{scratch, scratch + SCR_L2, reinterpret_cast<int*>(scratch + SCR_CNT1), reinterpret_cast<int*>(scratch + SCR_CNT2)};
}

// true in exactly one CTA: the last of `expected` to arrive at `counter` (which it resets to 0).
// CUTLASS-semaphore pattern: the CTA barrier orders every thread's partial writes before thread 0's
// acq_rel arrival (release is cumulative), and the acquire side orders the last CTA's reads after it.
V6_DEVINL bool arrive_last(int* counter, int expected, int* s_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        int old;
        asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
        const int last = old == expected - 1;
        if (last) *counter = 0;                 // every expected arrival has happened: plain reset for the next launch
        *s_flag = last;
    }
    __syncthreads();
    return *s_flag != 0;
}

// dst[col] = sum_p src[p][col], p < n; cols in {16..128}.  float4 columns: cols/4 threads per row, THREADS/(cols/4)
// row lanes, up to 8 independent 16 B loads in flight per thread; fixed summation order.
V6_DEVINL void fold_rows(const float* __restrict__ src, int n, int cols, float* red, float* dst) {
    const int c4 = cols >> 2, lanes = THREADS / c4;
    const int col4 = threadIdx.x % c4, ln = threadIdx.x / c4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p0 = ln; p0 < n; p0 += 8 * lanes) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int p = p0 + u * lanes;
            v[u] = p < n ? __ldcg(reinterpret_cast<const float4*>(src + (size_t)p * cols) + col4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    reinterpret_cast<float4*>(red)[threadIdx.x] = s;          // red: [lanes][cols]
    __syncthreads();
    if (threadIdx.x < cols) {
        float t = 0.f;
        for (int l = 0; l < lanes; ++l) t += red[l * cols + threadIdx.x];
        dst[threadIdx.x] = t;
    }
}

// Reduce the per-thread accumulators a[8], b[8] (thread = 8 channels of one row lane) over the whole
// slice.  Returns true in the one CTA per slice that ends up with the totals in tot[0..SW) (sum a) and
// tot[SW..2SW) (sum b).
V6_DEVINL bool slice_reduce(float (&a)[8], float (&b)[8], const Red& rd, int SW, float* tot) {
    __shared__ __align__(16) float red[8 * 128];
    __shared__ int s_flag;
    const int CGS = SW >> 3, cols = 2 * SW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int off = CGS; off < 32; off <<= 1) {           // lanes with the same channel group hold different rows
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a[k] += __shfl_xor_sync(0xffffffffu, a[k], off);
            b[k] += __shfl_xor_sync(0xffffffffu, b[k], off);
        }
    }
    if (lane < CGS) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { red[warp * 128 + lane * 8 + k] = a[k]; red[warp * 128 + SW + lane * 8 + k] = b[k]; }
    }
    __syncthreads();
    const int slice = blockIdx.y, rc = blockIdx.x, nrc = gridDim.x;
    const int g1 = nrc <= ONE_LEVEL_MAX ? nrc : G1;          // few row-CTAs: one fold does it all
    const int ngrp = (nrc + g1 - 1) / g1;
    if (threadIdx.x < cols) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) s += red[w * 128 + threadIdx.x];
        __stcg(rd.l1 + ((size_t)slice * nrc + rc) * cols + threadIdx.x, s);
    }
    const int grp = rc / g1;
    const int gsz = min(g1, nrc - grp * g1);
    if (!arrive_last(rd.cnt1 + slice * ngrp + grp, gsz, &s_flag)) return false;
    if (ngrp == 1) {
        fold_rows(rd.l1 + (size_t)slice * nrc * cols, gsz, cols, red, tot);
    } else {
        fold_rows(rd.l1 + ((size_t)slice * nrc + (size_t)grp * g1) * cols, gsz, cols, red, rd.l2 + ((size_t)slice * ngrp + grp) * cols);
        if (!arrive_last(rd.cnt2 + slice, ngrp, &s_flag)) return false;
        fold_rows(rd.l2 + (size_t)slice * ngrp * cols, ngrp, cols, red, tot);
    }
    __syncthreads();
    return true;
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

template <bool RELU, bool RES>
__global__ void __launch_bounds__(THREADS) bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                                                           const float* __restrict__ scale, const float* __restrict__ bias,
                                                           __nv_bfloat16* __restrict__ y, unsigned char* __restrict__ mask,
                                                           long long R, int C) {
    const int CG = C >> 3, RL = THREADS / CG;
    const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
    pdl_wait();                                                 // scale / bias come from the stats kernel
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
                unsigned bits = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float t = fmaf(v[u][k], sc[k], bi[k]);
                    if (RES) t += q[u][k];
                    if (RELU) bits |= (t > 0.f ? 1u : 0u) << k;
                    v[u][k] = RELU ? fmaxf(t, 0.f) : t;
                }
                store8(y + r * C + cg * 8, v[u]);
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
    const long long CGT = C >> 3, cgt = (long long)blockIdx.y * CGS + cg;       // mask: [R][C/8] bytes
    float mu[8], rs[8], sg[8], sgx[8];
    loadf8(mean + c0, mu);
    loadf8(rstd + c0, rs);
#pragma unroll
    for (int k = 0; k < 8; ++k) { sg[k] = 0.f; sgx[k] = 0.f; }
    const long long G = (long long)gridDim.x * RL;
    long long r = (long long)blockIdx.x * RL + rl;
    for (; r + G < R; r += 2 * G) {                                            // 2 rows x 3 tensors in flight
        uint4 g[2], xv[2];
        unsigned mb[2] = {0xffu, 0xffu};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long o = (r + u * G) * C + c0;
            g[u] = ldg_nc_v4(dy + o);
            xv[u] = ldg_nc_v4(x + o);
            if (RELU) mb[u] = __ldg(mask + (r + u * G) * CGT + cgt);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) acc_bwd<RELU>(g[u], xv[u], mb[u], mu, rs, sg, sgx);
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

template <bool RELU, bool RES>
__global__ void __launch_bounds__(THREADS) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const unsigned char* __restrict__ mask,
                                                               const __nv_bfloat16* __restrict__ x, const float* __restrict__ coef,
                                                               __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dres,
                                                               long long R, int C) {
    const int CG = C >> 3, RL = THREADS / CG;
    const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
    pdl_wait();                                                 // coefficients come from the reduce kernel
    float c0[8], c1[8], c2[8];
    loadf8(coef + cg * 8, c0);
    loadf8(coef + C + cg * 8, c1);
    loadf8(coef + 2 * C + cg * 8, c2);
    const long long G = (long long)gridDim.x * RL;
    for (long long r0 = (long long)blockIdx.x * RL + rl; r0 < R; r0 += 2 * G) {
        float g[2][8], xv[2][8];
        unsigned mb[2] = {0xffu, 0xffu};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long r = r0 + u * G;
            if (r < R) {
                load8(dy + r * C + cg * 8, g[u]);
                load8(x + r * C + cg * 8, xv[u]);
                if (RELU) mb[u] = __ldg(mask + r * CG + cg);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long r = r0 + u * G;
            if (r < R) {
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (RELU && !((mb[u] >> k) & 1u)) g[u][k] = 0.f;
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
    return C % 8 == 0 && cg >= 1 && cg <= THREADS && (cg & (cg - 1)) == 0;     // 8, 16, ..., 2048
}
// one full wave of resident CTAs for a reduce kernel (occupancy queried once per kernel)
template <typename K>
static int wave_ctas(K kernel, int& cache) {
    if (cache == 0) {
        int dev = 0, sms = 148, occ = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, THREADS, 0) != cudaSuccess || occ < 1) occ = 1;
        cache = occ * sms > MAX_CTAS ? MAX_CTAS : occ * sms;
    }
    return cache;
}

// reduce grids: x = row-CTAs, y = 64-channel slices
static inline dim3 reduce_grid(long long R, int C, int target) {
    const int SW = C < 64 ? C : 64, slices = C / SW, RL = THREADS / (SW >> 3);
    long long nrc = (R + 4LL * RL - 1) / (4LL * RL);          // >= 4 rows per thread: the unrolled fast path, fewer partials
    const long long cap = target / slices > 0 ? target / slices : 1;
    if (nrc > cap) nrc = cap;
    return dim3((unsigned)(nrc < 1 ? 1 : nrc), (unsigned)slices, 1);
}

// element-wise passes keep no partials: use every resident CTA slot (8 CTAs/SM x 148 SMs)
static inline int apply_grid(long long R, int C) {
    const int RL = THREADS / (C >> 3);
    long long g = (R + RL - 1) / RL;
    return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

// launch `kernel` as a programmatic dependent of the previous kernel in the stream (V6B200_PDL=0: plain launch)
template <typename... KArgs, typename... Args>
static void launch_dependent(void (*kernel)(KArgs...), int grid, cudaStream_t s, Args... args) {
    static const bool pdl = [] { const char* e = getenv("V6B200_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(THREADS);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace bn

// scratch: v6_bn_scratch_floats() floats, zero-initialised once (partials + self-resetting counters), shared by
// every launch on one stream.  stats: mean[C] rstd[C] scale_bias[2C].
extern "C" long long v6_bn_scratch_floats() { return (long long)bn::SCR_FLOATS; }

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
    const int ag = apply_grid(R, C);
    const __nv_bfloat16* xx = (const __nv_bfloat16*)x;
    const __nv_bfloat16* rr = (const __nv_bfloat16*)res;
    __nv_bfloat16* yy = (__nv_bfloat16*)y;
    if (relu) {
        unsigned char* mk = (unsigned char*)relu_mask;
        if (res) launch_dependent(bn_apply_kernel<true, true>, ag, s, xx, rr, scale_bias, scale_bias + C, yy, mk, R, C);
        else launch_dependent(bn_apply_kernel<true, false>, ag, s, xx, rr, scale_bias, scale_bias + C, yy, mk, R, C);
    } else {
        if (res) launch_dependent(bn_apply_kernel<false, true>, ag, s, xx, rr, scale_bias, scale_bias + C, yy, nullptr, R, C);
        else launch_dependent(bn_apply_kernel<false, false>, ag, s, xx, rr, scale_bias, scale_bias + C, yy, nullptr, R, C);
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
        if (res) bn_apply_kernel<true, true><<<ag, THREADS, 0, s>>>(xx, rr, scale, bias, yy, nullptr, R, C);
        else bn_apply_kernel<true, false><<<ag, THREADS, 0, s>>>(xx, rr, scale, bias, yy, nullptr, R, C);
    } else {
        if (res) bn_apply_kernel<false, true><<<ag, THREADS, 0, s>>>(xx, rr, scale, bias, yy, nullptr, R, C);
        else bn_apply_kernel<false, false><<<ag, THREADS, 0, s>>>(xx, rr, scale, bias, yy, nullptr, R, C);
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
    if (relu) bn_bwd_reduce_kernel<true><<<rg, THREADS, 0, s>>>(dyy, yy, xx, gamma, mean, rstd, make_red(scratch), dgamma, dbeta, coef, R, C, accumulate);
    else bn_bwd_reduce_kernel<false><<<rg, THREADS, 0, s>>>(dyy, yy, xx, gamma, mean, rstd, make_red(scratch), dgamma, dbeta, coef, R, C, accumulate);
    const int ag = apply_grid(R, C);
    __nv_bfloat16* dxx = (__nv_bfloat16*)dx;
    __nv_bfloat16* drr = (__nv_bfloat16*)dres;
    if (relu) {
        if (dres) launch_dependent(bn_bwd_apply_kernel<true, true>, ag, s, dyy, yy, xx, coef, dxx, drr, R, C);
        else launch_dependent(bn_bwd_apply_kernel<true, false>, ag, s, dyy, yy, xx, coef, dxx, drr, R, C);
    } else {
        if (dres) launch_dependent(bn_bwd_apply_kernel<false, true>, ag, s, dyy, yy, xx, coef, dxx, drr, R, C);
        else launch_dependent(bn_bwd_apply_kernel<false, false>, ag, s, dyy, yy, xx, coef, dxx, drr, R, C);
    }
    V6_CHECK_LAUNCH();
    return 0;
}
