// Cross-entropy over bf16 logits in two row-wise passes that never materialise an fp32 copy (Llama-3: 1024 x 128256 logits).
//
// The composed PyTorch expression the models used -- logits.float() -> log_softmax -> nll -> (backward) softmax - onehot ->
// .to(bf16) -- moves ~4.4 GB per step for that head and keeps two 525 MB fp32 tensors alive; here
//   forward : one CTA per row walks the row once (online max / sum of exp in fp32, 16 B loads), writes lse[row] and
//             loss[row] = lse - logit[label]                                         (reads the bf16 logits once)
//   backward: dlogits = (exp(logit - lse) - onehot) * g, written IN PLACE over the logits (they are the output of the head GEMM
//             and nothing else reads them), bf16                                      (one read + one write)
// Rows whose label equals ignore_index contribute nothing (loss 0, zero gradient).
#include "common.cuh"
#include "api.h"

namespace ce {

constexpr int THREADS = 512;

V6_DEVINL void unpack8(const uint4& t, float (&v)[8]) {
    float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

// combine two (max, sum-of-exp relative to max) pairs
V6_DEVINL void merge(float& m, float& s, float m2, float s2) {
    const float mm = fmaxf(m, m2);
    // a thread / warp that saw no element carries (-inf, 0): exp(-inf - -inf) would be NaN
    s = (m == -INFINITY ? 0.f : s * __expf(m - mm)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mm));
    m = mm;
}

__global__ void __launch_bounds__(THREADS) ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, const long long* __restrict__ labels,
                                                         float* __restrict__ lse, float* __restrict__ loss, int V, long long ld,
                                                         long long ignore_index) {
    __shared__ float sm[THREADS / 32], ss[THREADS / 32];
    const int row = blockIdx.x;
    const __nv_bfloat16* x = logits + (size_t)row * ld;
    const long long lab = labels[row];
    float m = -INFINITY, s = 0.f;
    const int nvec = V >> 3;
    for (int i = threadIdx.x; i < nvec; i += THREADS) {
        float v[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), v);
        float vm = v[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) vm = fmaxf(vm, v[k]);
        const float mm = fmaxf(m, vm);
        float acc = s * __expf(m - mm);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += __expf(v[k] - mm);
        m = mm; s = acc;
    }
    for (int i = (nvec << 3) + threadIdx.x; i < V; i += THREADS) {        // tail (V not a multiple of 8)
        const float v = __bfloat162float(x[i]);
        const float mm = fmaxf(m, v);
        s = s * __expf(m - mm) + __expf(v - mm);
        m = mm;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        merge(m, s, m2, s2);
    }
    if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ss[threadIdx.x >> 5] = s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float M = sm[0], S = ss[0];
        for (int w = 1; w < THREADS / 32; ++w) merge(M, S, sm[w], ss[w]);
        const float l = M + __logf(S);
        lse[row] = l;
        loss[row] = (lab == ignore_index) ? 0.f : l - __bfloat162float(x[lab]);
    }
}

// scale_ptr: device scalar = upstream gradient / number of counted rows
// V_alloc >= V: columns [V, V_alloc) are padding of the head GEMM (vocabulary padded to a tile multiple); their gradient is zero
__global__ void __launch_bounds__(THREADS) ce_bwd_kernel(__nv_bfloat16* __restrict__ logits, const long long* __restrict__ labels,
                                                         const float* __restrict__ lse, const float* __restrict__ scale_ptr, int V,
                                                         int V_alloc, long long ld, long long ignore_index) {
    const int row = blockIdx.x;
    __nv_bfloat16* x = logits + (size_t)row * ld;
    const long long lab = labels[row];
    const bool counted = lab != ignore_index;
    const float g = counted ? __ldg(scale_ptr) : 0.f, l = lse[row];
    const int nvec = V >> 3;
    for (int i = threadIdx.x; i < nvec; i += THREADS) {
        float v[8];
        unpack8(*(reinterpret_cast<const uint4*>(x) + i), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float p = __expf(v[k] - l);
            v[k] = g * (p - ((long long)(i * 8 + k) == lab ? 1.f : 0.f));
        }
        *(reinterpret_cast<uint4*>(x) + i) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
    for (int i = (nvec << 3) + threadIdx.x; i < V; i += THREADS) {
        const float p = __expf(__bfloat162float(x[i]) - l);
        x[i] = __float2bfloat16_rn(g * (p - ((long long)i == lab ? 1.f : 0.f)));
    }
    for (int i = V + threadIdx.x; i < V_alloc; i += THREADS) x[i] = __float2bfloat16_rn(0.f);
}

}  // namespace ce

extern "C" int v6_ce_fwd(const void* logits, const long long* labels, float* lse, float* loss, int T, int V, long long ld,
                         long long ignore_index, cudaStream_t s) {
    if (T < 1 || V < 1 || ld % 8 != 0) return (int)cudaErrorInvalidValue;
    ce::ce_fwd_kernel<<<T, ce::THREADS, 0, s>>>((const __nv_bfloat16*)logits, labels, lse, loss, V, ld, ignore_index);
    V6_CHECK_LAUNCH();
    return 0;
}

extern "C" int v6_ce_bwd(void* logits, const long long* labels, const float* lse, const float* scale_ptr, int T, int V, int V_alloc,
                         long long ld, long long ignore_index, cudaStream_t s) {
    if (T < 1 || V < 1 || ld % 8 != 0 || V_alloc < V || V_alloc > ld) return (int)cudaErrorInvalidValue;
    ce::ce_bwd_kernel<<<T, ce::THREADS, 0, s>>>((__nv_bfloat16*)logits, labels, lse, scale_ptr, V, V_alloc, ld, ignore_index);
    V6_CHECK_LAUNCH();
    return 0;
}
