// K7 of SURVEY.md 2.6: local optimizers over FLAT parameter buffers, one launch per step.
//
// All parameters of a model live in one contiguous fp32 buffer (vantage6_b200.models.flat),
// so the optimizer is a single memory-bound sweep instead of torch.optim's per-tensor /
// foreach launches.  Two federated fusions ride on the same sweep:
//   * first local step of a round : save the incoming global model  (w_ref <- w)  -- free,
//     the old value is already in registers;
//   * last local step of a round  : publish the contribution for the FedAvg reduction
//     upload <- n_i * (w_new - w_ref)   (fp32 or bf16) -- "delta cast/scale" of K2 fused here.
// An optional bf16 shadow copy of the updated weights is written for bf16 compute paths.
#include "common.cuh"
#include "api.h"


template <int KIND /*0 sgd, 1 adamw*/>
__global__ void __launch_bounds__(512, 2) flat_optim_kernel(const OptimParams P) {
    const float cscale = P.contrib_scale_ptr ? __ldg(P.contrib_scale_ptr) : P.contrib_scale;   // n_i (device scalar: CUDA-graph safe)
    const float gs = P.grad_scale_ptr ? *P.grad_scale_ptr : 1.f;
    const float bias1 = P.bias_ptr ? P.bias_ptr[0] : P.bias1;
    const float bias2 = P.bias_ptr ? P.bias_ptr[1] : P.bias2;
    const long long n4 = P.n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        float4 w4 = reinterpret_cast<const float4*>(P.w)[i];
        float4 g4 = ldg_stream_f4(reinterpret_cast<const float4*>(P.g) + i);
        float4 m4 = reinterpret_cast<const float4*>(P.m)[i];
        float w[4] = {w4.x, w4.y, w4.z, w4.w}, g[4] = {g4.x * gs, g4.y * gs, g4.z * gs, g4.w * gs};
        float m[4] = {m4.x, m4.y, m4.z, m4.w};
        float wold[4] = {w[0], w[1], w[2], w[3]};
        if constexpr (KIND == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float d = fmaf(P.weight_decay, w[k], g[k]);
                if (P.momentum != 0.f) {
                    m[k] = P.first_momentum_step ? d : fmaf(P.momentum, m[k], (1.f - P.dampening) * d);
                    d = P.nesterov ? fmaf(P.momentum, m[k], d) : m[k];
                }
                w[k] = fmaf(-P.lr, d, w[k]);
            }
        } else {
            float4 v4 = reinterpret_cast<const float4*>(P.v)[i];
            float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                w[k] *= (1.f - P.lr * P.weight_decay);                 // decoupled decay
                m[k] = fmaf(P.beta1, m[k], (1.f - P.beta1) * g[k]);
                v[k] = fmaf(P.beta2, v[k], (1.f - P.beta2) * g[k] * g[k]);
                const float mh = m[k] * bias1;                       // bias1 = 1/(1-b1^t)
                const float vh = v[k] * bias2;
                w[k] = fmaf(-P.lr, mh / (sqrtf(vh) + P.eps), w[k]);
            }
            reinterpret_cast<float4*>(P.v)[i] = make_float4(v[0], v[1], v[2], v[3]);
        }
        reinterpret_cast<float4*>(P.w)[i] = make_float4(w[0], w[1], w[2], w[3]);
        reinterpret_cast<float4*>(P.m)[i] = make_float4(m[0], m[1], m[2], m[3]);
        float ref[4];
        if (P.save_ref) {
            reinterpret_cast<float4*>(P.w_ref)[i] = make_float4(wold[0], wold[1], wold[2], wold[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) ref[k] = wold[k];
        } else if (P.publish == 1 || P.publish == 2) {
            float4 r4 = reinterpret_cast<const float4*>(P.w_ref)[i];
            ref[0] = r4.x; ref[1] = r4.y; ref[2] = r4.z; ref[3] = r4.w;
        }
        if (P.publish == 1) {
            st_f4(reinterpret_cast<float4*>(P.upload) + i,
                  make_float4(cscale * (w[0] - ref[0]), cscale * (w[1] - ref[1]),
                              cscale * (w[2] - ref[2]), cscale * (w[3] - ref[3])));
        } else if (P.publish == 2) {
            uint2 d = make_uint2(pack_bf16x2(cscale * (w[0] - ref[0]), cscale * (w[1] - ref[1])),
                                 pack_bf16x2(cscale * (w[2] - ref[2]), cscale * (w[3] - ref[3])));
            reinterpret_cast<uint2*>(P.upload)[i] = d;
        } else if (P.publish == 3) {
            st_f4(reinterpret_cast<float4*>(P.upload) + i,
                  make_float4(cscale * w[0], cscale * w[1], cscale * w[2], cscale * w[3]));
        }
        if (P.shadow) {
            reinterpret_cast<uint2*>(P.shadow)[i] = make_uint2(pack_bf16x2(w[0], w[1]), pack_bf16x2(w[2], w[3]));
        }
    }
}

static inline int optim_grid(long long n) {
    long long blocks = (n / 4 + 511) / 512;
    long long cap = 148LL * 2 * 4;            // 4 waves of 2 CTAs/SM: enough MLP, short tail
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}
extern "C" int v6_flat_sgd(const OptimParams* hp, cudaStream_t s) {
    flat_optim_kernel<0><<<optim_grid(hp->n), 512, 0, s>>>(*hp);
    V6_CHECK_LAUNCH(); return 0;
}
extern "C" int v6_flat_adamw(const OptimParams* hp, cudaStream_t s) {
    flat_optim_kernel<1><<<optim_grid(hp->n), 512, 0, s>>>(*hp);
    V6_CHECK_LAUNCH(); return 0;
}

// ----------------------------------------------------------------------------------------
// standalone delta publish (when the optimizer is not ours, e.g. the NCCL baseline) and
// fp32 -> bf16 shadow cast; plus sum-of-squares for gradient clipping.
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 2) delta_publish_kernel(const float* __restrict__ w, const float* __restrict__ ref,
                                                               void* upload, long long n4, float scale, int bf16_out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 a = reinterpret_cast<const float4*>(w)[i];
        float4 b = ref ? reinterpret_cast<const float4*>(ref)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 d = make_float4(scale * (a.x - b.x), scale * (a.y - b.y), scale * (a.z - b.z), scale * (a.w - b.w));
        if (bf16_out) reinterpret_cast<uint2*>(upload)[i] = make_uint2(pack_bf16x2(d.x, d.y), pack_bf16x2(d.z, d.w));
        else st_f4(reinterpret_cast<float4*>(upload) + i, d);
    }
}
extern "C" int v6_delta_publish(const float* w, const float* ref, void* upload, long long n, float scale,
                                int bf16_out, cudaStream_t s) {
    delta_publish_kernel<<<optim_grid(n), 512, 0, s>>>(w, ref, upload, n / 4, scale, bf16_out);
    V6_CHECK_LAUNCH(); return 0;
}

__global__ void __launch_bounds__(512, 2) cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 a = reinterpret_cast<const float4*>(src)[i];
        reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
    }
}
extern "C" int v6_cast_bf16(const float* src, void* dst, long long n, cudaStream_t s) {
    cast_bf16_kernel<<<optim_grid(n), 512, 0, s>>>(src, (__nv_bfloat16*)dst, n / 4);
    V6_CHECK_LAUNCH(); return 0;
}

// Multi-tensor gradient sink: dst[off_k + i] += float(src_k[i]) for up to MULTI_MAX bf16 tensors in ONE launch
// (the per-layer bf16 weight gradients the conv backward produces -> the flat fp32 gradient buffer;
// replaces one cast + one add launch per layer).  blockIdx.y = tensor, blockIdx.x strides over it.
__global__ void __launch_bounds__(256) multi_accum_kernel(const MultiAccumParams p, float* __restrict__ dst) {
    const int k = blockIdx.y;
    const __nv_bfloat16* __restrict__ src = reinterpret_cast<const __nv_bfloat16*>(p.src[k]);
    float* __restrict__ d = dst + p.dst_off[k];
    const long long n = p.numel[k], n8 = n >> 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(src) + i);
        float4 a = reinterpret_cast<float4*>(d)[2 * i], b = reinterpret_cast<float4*>(d)[2 * i + 1];
        const float2 t0 = unpack_bf16x2(t.x), t1 = unpack_bf16x2(t.y), t2 = unpack_bf16x2(t.z), t3 = unpack_bf16x2(t.w);
        a.x += t0.x; a.y += t0.y; a.z += t1.x; a.w += t1.y;
        b.x += t2.x; b.y += t2.y; b.z += t3.x; b.w += t3.y;
        reinterpret_cast<float4*>(d)[2 * i] = a;
        reinterpret_cast<float4*>(d)[2 * i + 1] = b;
    }
    if (blockIdx.x == 0)
        for (long long i = (n8 << 3) + threadIdx.x; i < n; i += blockDim.x) d[i] += __bfloat162float(src[i]);
}
extern "C" int v6_multi_accum_bf16(const MultiAccumParams* p, float* dst, cudaStream_t s) {
    if (p->count < 1 || p->count > V6_MULTI_MAX) return (int)cudaErrorInvalidValue;
    multi_accum_kernel<<<dim3(32, p->count), 256, 0, s>>>(*p, dst);
    V6_CHECK_LAUNCH(); return 0;
}

// out[0] += sum(x^2)  (out must be zeroed by the caller); used for global-norm clipping:
// the clip coefficient is then computed on device and consumed via OptimParams::grad_scale_ptr.
__global__ void __launch_bounds__(512, 2) sumsq_kernel(const float* __restrict__ x, long long n4, float* out) {
    __shared__ float red[32];
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 a = ldg_stream_f4(reinterpret_cast<const float4*>(x) + i);
        acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}
__global__ void clip_coef_kernel(const float* sumsq, float max_norm, float* coef) {
    const float nrm = sqrtf(*sumsq);
    *coef = fminf(1.f, max_norm / (nrm + 1e-6f));
}
// step counter lives on the device so that a captured CUDA graph advances Adam's bias correction
__global__ void adam_bias_kernel(int* step, float beta1, float beta2, float* out) {
    const int t = ++(*step);
    out[0] = 1.f / (1.f - powf(beta1, (float)t));
    out[1] = 1.f / (1.f - powf(beta2, (float)t));
}
extern "C" int v6_adam_bias_update(int* step_counter, float beta1, float beta2, float* bias_out, cudaStream_t s) {
    adam_bias_kernel<<<1, 1, 0, s>>>(step_counter, beta1, beta2, bias_out);
    V6_CHECK_LAUNCH(); return 0;
}

extern "C" int v6_clip_coef(const float* g, long long n, float max_norm, float* sumsq_scratch, float* coef, cudaStream_t s) {
    cudaMemsetAsync(sumsq_scratch, 0, sizeof(float), s);
    sumsq_kernel<<<optim_grid(n), 512, 0, s>>>(g, n / 4, sumsq_scratch);
    clip_coef_kernel<<<1, 1, 0, s>>>(sumsq_scratch, max_norm, coef);
    V6_CHECK_LAUNCH(); return 0;
}
