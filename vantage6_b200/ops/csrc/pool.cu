// ResNet stem helpers for NHWC bf16 activations (memory-bound, one pass each):
//
//   * image_normalize : uint8 NCHW image batch -> (x - mean) / std as bf16 NHWC (one kernel instead of the
//                       float cast, subtract, multiply and layout-conversion passes over a 38 MB fp32 tensor)
//   * maxpool 3x3/s2/p1 forward  : y + the arg-max position inside the window (uint8), 16 B vectors
//   * maxpool 3x3/s2/p1 backward : gather form -- every input pixel looks at the <= 4 windows that contain it
//                                  and takes dy where the saved arg-max points back at it (no atomics,
//                                  deterministic, dx written exactly once)
//
// The ATen NHWC max-pool kernels measured 206 us forward for the 64x112x112x64 stem activation on B200
// (profiles/launches_resnet50_fusedbn_v2_r1.txt) -- ~10x the HBM time of its 103 MB read + 26 MB write.
#include "common.cuh"
#include "api.h"

namespace pool {

constexpr int THREADS = 256;

V6_DEVINL void unpack8(const uint4& t, float (&v)[8]) {
    float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
V6_DEVINL uint4 pack8(const float (&v)[8]) {
    return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// one thread = 8 channels of one output pixel
template <typename IdxT>       // int when the tensor has < 2^31 elements (32-bit div / mod instead of the 64-bit call sequence)
__global__ void __launch_bounds__(THREADS) maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                              unsigned char* __restrict__ idx, int N, int H, int W, int C, int Ho,
                                                              int Wo) {
    const int CG = C >> 3;
    const IdxT total = (IdxT)N * Ho * Wo * CG;
    for (IdxT t = (IdxT)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (IdxT)gridDim.x * blockDim.x) {
        const int cg = (int)(t % CG);
        IdxT p = t / CG;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int n = (int)(p / Ho);
        float best[8];
        unsigned arg[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; arg[k] = 0; }
        const int h0 = 2 * ho - 1, w0 = 2 * wo - 1;
        uint4 raw[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const int h = h0 + j / 3, w = w0 + j % 3;
            if (h >= 0 && h < H && w >= 0 && w < W)
                raw[j] = __ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + w) * C + cg * 8));
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const int h = h0 + j / 3, w = w0 + j % 3;
            if (h >= 0 && h < H && w >= 0 && w < W) {
                float v[8];
                unpack8(raw[j], v);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (v[k] > best[k]) { best[k] = v[k]; arg[k] = j; }     // first maximum in scan order (ATen semantics)
            }
        }
        const long long o = (((long long)n * Ho + ho) * Wo + wo) * C + cg * 8;
        *reinterpret_cast<uint4*>(y + o) = pack8(best);
        *reinterpret_cast<uint2*>(idx + o) = make_uint2(arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24),
                                                         arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24));
    }
}

// one thread = 8 channels of one input pixel
__global__ void __launch_bounds__(THREADS) maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const unsigned char* __restrict__ idx,
                                                              __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int Ho,
                                                              int Wo) {
    const int CG = C >> 3;
    const long long total = (long long)N * H * W * CG;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(t % CG);
        long long p = t / CG;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int n = (int)(p / H);
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        // windows (ho, wo) with 2*ho-1 <= h <= 2*ho+1: ho in {h/2, (h+1)/2} (equal when h is even)
        const int ho_lo = h >> 1, ho_hi = (h + 1) >> 1, wo_lo = w >> 1, wo_hi = (w + 1) >> 1;
        for (int ho = ho_lo; ho <= ho_hi; ++ho) {
            if (ho >= Ho) continue;
            const int dh = h - (2 * ho - 1);
            for (int wo = wo_lo; wo <= wo_hi; ++wo) {
                if (wo >= Wo) continue;
                const unsigned me = (unsigned)(dh * 3 + (w - (2 * wo - 1)));
                const long long o = (((long long)n * Ho + ho) * Wo + wo) * C + cg * 8;
                const uint2 a = __ldg(reinterpret_cast<const uint2*>(idx + o));
                const uint4 g = __ldg(reinterpret_cast<const uint4*>(dy + o));
                float gv[8];
                unpack8(g, gv);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const unsigned ak = ((k < 4 ? a.x : a.y) >> (8 * (k & 3))) & 0xffu;
                    if (ak == me) acc[k] += gv[k];
                }
            }
        }
        *reinterpret_cast<uint4*>(dx + (((long long)n * H + h) * W + w) * C + cg * 8) = pack8(acc);
    }
}

// even H, W (the ResNet stem: 112 x 112): one thread = 8 channels of a 2x2 block of input pixels.  The block (2i..2i+1,
// 2j..2j+1) is covered by exactly the 4 windows (i..i+1, j..j+1), so 4 (dy, arg-max) loads -- all issued before the
// first use -- produce 4 outputs (the gather-per-pixel form above needs 2.25 loads per output), with 32-bit indexing.
__global__ void __launch_bounds__(THREADS) maxpool_bwd2x2_kernel(const __nv_bfloat16* __restrict__ dy, const unsigned char* __restrict__ idx,
                                                                 __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int Ho, int Wo) {
    const int CG = C >> 3, H2 = H >> 1, W2 = W >> 1;
    const int total = N * H2 * W2 * CG;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int cg = t % CG;
        int p = t / CG;
        const int j = p % W2; p /= W2;
        const int i = p % H2;
        const int n = p / H2;
        uint4 g[4];
        uint2 a[4];
#pragma unroll
        for (int wnd = 0; wnd < 4; ++wnd) {
            const int ho = i + (wnd >> 1), wo = j + (wnd & 1);
            g[wnd] = make_uint4(0, 0, 0, 0);
            a[wnd] = make_uint2(0xffffffffu, 0xffffffffu);                   // arg 255: matches no position
            if (ho < Ho && wo < Wo) {
                const size_t o = ((size_t)(n * Ho + ho) * Wo + wo) * C + cg * 8;
                a[wnd] = __ldg(reinterpret_cast<const uint2*>(idx + o));
                g[wnd] = __ldg(reinterpret_cast<const uint4*>(dy + o));
            }
        }
        float acc[4][8];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[q][k] = 0.f;
#pragma unroll
        for (int wnd = 0; wnd < 4; ++wnd) {
            float gv[8];
            unpack8(g[wnd], gv);
            const int dho = wnd >> 1, dwo = wnd & 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {                                     // pixel (2i + qa, 2j + qb)
                const int qa = q >> 1, qb = q & 1;
                if ((dho && !qa) || (dwo && !qb)) continue;                   // the next window only reaches the odd row / column
                const unsigned me = (unsigned)((dho ? 0 : 1 + qa) * 3 + (dwo ? 0 : 1 + qb));
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const unsigned ak = ((k < 4 ? a[wnd].x : a[wnd].y) >> (8 * (k & 3))) & 0xffu;
                    if (ak == me) acc[q][k] += gv[k];
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(dx + ((size_t)(n * H + 2 * i + (q >> 1)) * W + 2 * j + (q & 1)) * C + cg * 8) = pack8(acc[q]);
    }
}

// one thread = 4 horizontally adjacent pixels: 3 x uchar4 plane reads -> 12 bf16 (24 B) NHWC write
__global__ void __launch_bounds__(THREADS) image_normalize_kernel(const unsigned char* __restrict__ img, __nv_bfloat16* __restrict__ out,
                                                                  long long n_quads, long long HW, float m0, float m1, float m2,
                                                                  float s0, float s1, float s2) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += (long long)gridDim.x * blockDim.x) {
        const long long pix = q * 4;
        const long long n = pix / HW, r = pix % HW;
        const unsigned char* base = img + n * 3 * HW + r;
        const uchar4 c0 = *reinterpret_cast<const uchar4*>(base);
        const uchar4 c1 = *reinterpret_cast<const uchar4*>(base + HW);
        const uchar4 c2 = *reinterpret_cast<const uchar4*>(base + 2 * HW);
        const float a0 = ((float)c0.x - m0) * s0, a1 = ((float)c1.x - m1) * s1, a2 = ((float)c2.x - m2) * s2;
        const float b0 = ((float)c0.y - m0) * s0, b1 = ((float)c1.y - m1) * s1, b2 = ((float)c2.y - m2) * s2;
        const float d0 = ((float)c0.z - m0) * s0, d1 = ((float)c1.z - m1) * s1, d2 = ((float)c2.z - m2) * s2;
        const float e0 = ((float)c0.w - m0) * s0, e1 = ((float)c1.w - m1) * s1, e2 = ((float)c2.w - m2) * s2;
        uint2* o = reinterpret_cast<uint2*>(out + pix * 3);
        o[0] = make_uint2(pack_bf16x2(a0, a1), pack_bf16x2(a2, b0));
        o[1] = make_uint2(pack_bf16x2(b1, b2), pack_bf16x2(d0, d1));
        o[2] = make_uint2(pack_bf16x2(d2, e0), pack_bf16x2(e1, e2));
    }
}

// ---------------------------------------------------------------------------------- space-to-depth stem
// The 7x7 / stride-2 / pad-3 stem convolution on 3 input channels runs on cuDNN's legacy kernels (390 us
// forward for a 64x224x224 batch, profiles/launches_resnet50_v3_r1.txt) because C=3 defeats the NHWC
// tensor-core path.  It is algebraically a 4x4 / stride-1 convolution on the 2x2 space-to-depth image
// (12 real channels, padded to 16 -> the sm_100 implicit-GEMM kernels apply):
//   y[i,j] = sum_{p,q<4} sum_{r,s<2,c<3} xs[i+p-2, j+q-2, (r,s,c)] * ws[p,q,(r,s,c)]
//   xs[u,v,(r,s,c)] = x[2u+r, 2v+s, c],   ws[p,q,(r,s,c)] = w[2p+r-1, 2q+s-1, c] (0 outside the 7x7 filter)
// The asymmetric padding (2 before, 1 after) is materialised by the image kernel: xs is [N, H/2+3, W/2+3, 16].

// one thread = one output pixel of xs: 12 uint8 reads -> 16 bf16 (32 B) write
__global__ void __launch_bounds__(THREADS) image_normalize_s2d_kernel(const unsigned char* __restrict__ img, __nv_bfloat16* __restrict__ out,
                                                                      int N, int H, int W, float m0, float m1, float m2, float s0,
                                                                      float s1, float s2) {
    const int Hs = H / 2 + 3, Ws = W / 2 + 3;
    const long long total = (long long)N * Hs * Ws;
    const float mean[3] = {m0, m1, m2}, inv[3] = {s0, s1, s2};
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(t % Ws);
        const int u = (int)((t / Ws) % Hs);
        const int n = (int)(t / ((long long)Ws * Hs));
        const int uu = u - 2, vv = v - 2;
        float o[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) o[k] = 0.f;
        if (uu >= 0 && uu < H / 2 && vv >= 0 && vv < W / 2) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const unsigned char* pl = img + ((long long)n * 3 + c) * H * W + (long long)(2 * uu) * W + 2 * vv;
                const uchar2 top = *reinterpret_cast<const uchar2*>(pl);
                const uchar2 bot = *reinterpret_cast<const uchar2*>(pl + W);
                o[0 * 3 + c] = ((float)top.x - mean[c]) * inv[c];       // (r=0, s=0)
                o[1 * 3 + c] = ((float)top.y - mean[c]) * inv[c];       // (r=0, s=1)
                o[2 * 3 + c] = ((float)bot.x - mean[c]) * inv[c];       // (r=1, s=0)
                o[3 * 3 + c] = ((float)bot.y - mean[c]) * inv[c];       // (r=1, s=1)
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(out + t * 16);
        dst[0] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        dst[1] = make_uint4(pack_bf16x2(o[8], o[9]), pack_bf16x2(o[10], o[11]), 0u, 0u);
    }
}

// ws[o][p][q][16] (bf16) from w[o][kh][kw][c] (the flat buffers' [O,H,W,I] layout; fp32 master or bf16 shadow)
template <typename T>
__global__ void stem_weight_s2d_kernel(const T* __restrict__ w, __nv_bfloat16* __restrict__ ws, int O) {
    const int total = O * 256;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int ch = t & 15, q = (t >> 4) & 3, p = (t >> 6) & 3, o = t >> 8;
        float val = 0.f;
        if (ch < 12) {
            const int c = ch % 3, rs = ch / 3, r = rs >> 1, s2 = rs & 1;
            const int kh = 2 * p + r - 1, kw = 2 * q + s2 - 1;
            if (kh >= 0 && kh < 7 && kw >= 0 && kw < 7) val = (float)w[((o * 7 + kh) * 7 + kw) * 3 + c];
        }
        ws[t] = __float2bfloat16(val);
    }
}

// dw[o][kh][kw][c] (fp32) (+)= dws[o][p][q][ch]
template <typename T>
__global__ void stem_wgrad_d2s_kernel(const T* __restrict__ dws, float* __restrict__ dw, int O, int accumulate) {
    const int total = O * 147;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int c = t % 3, kw = (t / 3) % 7, kh = (t / 21) % 7, o = t / 147;
        const int p = (kh + 1) >> 1, r = (kh + 1) & 1, q = (kw + 1) >> 1, s2 = (kw + 1) & 1;
        const float g = (float)dws[((o * 4 + p) * 4 + q) * 16 + (r * 2 + s2) * 3 + c];
        dw[t] = accumulate ? dw[t] + g : g;
    }
}

static inline int grid_for(long long work) {
    long long g = (work + THREADS - 1) / THREADS;
    return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace pool

extern "C" int v6_maxpool3x3s2_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, cudaStream_t s) {
    using namespace pool;
    if (C % 8 != 0 || N < 1 || H < 1 || W < 1) return (int)cudaErrorInvalidValue;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;           // floor((H + 2 - 3) / 2) + 1
    if ((long long)N * H * W * C < (1LL << 31))
        maxpool_fwd_kernel<int><<<grid_for((long long)N * Ho * Wo * (C >> 3)), THREADS, 0, s>>>(
            (const __nv_bfloat16*)x, (__nv_bfloat16*)y, (unsigned char*)idx, N, H, W, C, Ho, Wo);
    else
        maxpool_fwd_kernel<long long><<<grid_for((long long)N * Ho * Wo * (C >> 3)), THREADS, 0, s>>>(
            (const __nv_bfloat16*)x, (__nv_bfloat16*)y, (unsigned char*)idx, N, H, W, C, Ho, Wo);
    V6_CHECK_LAUNCH();
    return 0;
}

extern "C" int v6_maxpool3x3s2_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C, cudaStream_t s) {
    using namespace pool;
    if (C % 8 != 0 || N < 1 || H < 1 || W < 1) return (int)cudaErrorInvalidValue;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    if (H % 2 == 0 && W % 2 == 0 && (long long)N * H * W * C < (1LL << 31))
        maxpool_bwd2x2_kernel<<<grid_for((long long)N * (H / 2) * (W / 2) * (C >> 3)), THREADS, 0, s>>>(
            (const __nv_bfloat16*)dy, (const unsigned char*)idx, (__nv_bfloat16*)dx, N, H, W, C, Ho, Wo);
    else
        maxpool_bwd_kernel<<<grid_for((long long)N * H * W * (C >> 3)), THREADS, 0, s>>>(
            (const __nv_bfloat16*)dy, (const unsigned char*)idx, (__nv_bfloat16*)dx, N, H, W, C, Ho, Wo);
    V6_CHECK_LAUNCH();
    return 0;
}

// img: uint8 [N,3,H,W] (dense NCHW), out: bf16 [N,H,W,3]; H*W must be a multiple of 4
extern "C" int v6_image_normalize(const void* img, void* out, long long N, long long HW, float m0, float m1, float m2, float s0,
                                  float s1, float s2, cudaStream_t s) {
    using namespace pool;
    if (HW % 4 != 0 || N < 1) return (int)cudaErrorInvalidValue;
    const long long quads = N * HW / 4;
    image_normalize_kernel<<<grid_for(quads), THREADS, 0, s>>>((const unsigned char*)img, (__nv_bfloat16*)out, quads, HW, m0, m1, m2,
                                                               s0, s1, s2);
    V6_CHECK_LAUNCH();
    return 0;
}

// img: uint8 [N,3,H,W] (H, W even) -> out: bf16 [N, H/2+3, W/2+3, 16]
extern "C" int v6_image_normalize_s2d(const void* img, void* out, int N, int H, int W, float m0, float m1, float m2, float s0,
                                      float s1, float s2, cudaStream_t s) {
    using namespace pool;
    if ((H & 1) || (W & 1) || N < 1) return (int)cudaErrorInvalidValue;
    image_normalize_s2d_kernel<<<grid_for((long long)N * (H / 2 + 3) * (W / 2 + 3)), THREADS, 0, s>>>(
        (const unsigned char*)img, (__nv_bfloat16*)out, N, H, W, m0, m1, m2, s0, s1, s2);
    V6_CHECK_LAUNCH();
    return 0;
}

extern "C" int v6_stem_weight_s2d(const void* w, int w_is_bf16, void* ws, int O, cudaStream_t s) {
    using namespace pool;
    if (w_is_bf16) stem_weight_s2d_kernel<<<grid_for((long long)O * 256), THREADS, 0, s>>>((const __nv_bfloat16*)w, (__nv_bfloat16*)ws, O);
    else stem_weight_s2d_kernel<<<grid_for((long long)O * 256), THREADS, 0, s>>>((const float*)w, (__nv_bfloat16*)ws, O);
    V6_CHECK_LAUNCH();
    return 0;
}

extern "C" int v6_stem_wgrad_d2s(const void* dws, float* dw, int O, int accumulate, cudaStream_t s) {
    using namespace pool;
    stem_wgrad_d2s_kernel<<<grid_for((long long)O * 147), THREADS, 0, s>>>((const __nv_bfloat16*)dws, dw, O, accumulate);
    V6_CHECK_LAUNCH();
    return 0;
}
// same, from the fp32 filter gradient the tcgen05 WGRAD kernel accumulates (igemm.cu)
extern "C" int v6_stem_wgrad_d2s_f32(const float* dws, float* dw, int O, int accumulate, cudaStream_t s) {
    using namespace pool;
    stem_wgrad_d2s_kernel<<<grid_for((long long)O * 147), THREADS, 0, s>>>(dws, dw, O, accumulate);
    V6_CHECK_LAUNCH();
    return 0;
}
