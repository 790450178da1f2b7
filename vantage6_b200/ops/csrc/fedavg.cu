// K2 / K3 of SURVEY.md section 2.6: the federated aggregation hot path.
//
//   fedavg_round  : ONE kernel per round per rank that (1) signals "my local weights are
//                   final" to every reducer, (2) waits for the contributors, (3) pulls the
//                   contributions over NVLink (P2P loads from peer-mapped symmetric memory, or
//                   a single in-switch `multimem.ld_reduce`), forms the weighted FedAvg mean,
//                   (4) applies the server optimizer (FedAvg / FedAvgM / FedAdam) in registers
//                   on the fp32 master copy, (5) pushes the new global model straight into
//                   every node's parameter buffer (P2P stores or one `multimem.st`), and
//                   (6) signals / waits "broadcast complete".  No NCCL call, no separate
//                   broadcast, reduce, scale, cast or optimizer kernels.
//   small_allreduce: K3, latency-bound one-shot weighted all-reduce for <= 64 KB payloads
//                   (GLM coefficients, the 1k-parameter vector): single CTA, flag barrier,
//                   P2P loads of every peer slot, no grid sync.
//
// The reference transports these payloads as REST blobs through a SQL database
// (reference: vantage6/cli/server.py:223-228); here they never leave HBM.
#include "common.cuh"
#include "api.h"




template <typename T> struct Vec;            // 16 B vector of T
template <> struct Vec<float> { static constexpr int N = 4; };
template <> struct Vec<__nv_bfloat16> { static constexpr int N = 8; };

V6_DEVINL void load_contrib(const float* base, long long e, float (&v)[4]) {
    float4 t = ld_sys_f4(reinterpret_cast<const float4*>(base + e));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
V6_DEVINL void load_contrib(const __nv_bfloat16* base, long long e, float (&v)[8]) {
    uint4 t = ld_sys_u4(reinterpret_cast<const uint4*>(base + e));
    float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
V6_DEVINL void load_contrib_mc(const float* mc, long long e, float (&v)[4]) {
    float4 t = multimem_ld_reduce_add_f4(reinterpret_cast<const float4*>(mc + e));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
V6_DEVINL void load_contrib_mc(const __nv_bfloat16* mc, long long e, float (&v)[8]) {
    uint4 t = multimem_ld_reduce_add_bf16x8(reinterpret_cast<const uint4*>(mc + e));
    float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

template <typename UpT, int U, int WMAX>
__global__ void __launch_bounds__(512, 1)
fedavg_round_kernel(const FedAvgParams P) {
    constexpr int VN = Vec<UpT>::N;
    __shared__ int s_ok;
    __shared__ unsigned int s_miss;
    __shared__ float s_w[V6_MAX_PEERS];       // n_p of every contributor for this round (0 = not reporting / dead)
    __shared__ float s_inv;                   // 1 / sum_p n_p (1 / world for the un-weighted in-switch sum)
    __shared__ int s_use_mc;
    const bool is_red = ((P.reducer_mask >> P.rank) & 1u) != 0u;     // this rank owns a slice of the global model
    const bool reducer = is_red && P.hi > P.lo;
    uint32_t* my_pad = reinterpret_cast<uint32_t*>(P.pads.p[P.rank]);

    // (1) tell every reducer that my contribution for this epoch is final (and with which weight n_i: a rank only
    // has to know its OWN sample count).  Stream order guarantees the producing kernels completed; the release
    // makes contribution and weight visible system-wide.
    if (blockIdx.x == 0 && threadIdx.x < P.world && ((P.reducer_mask >> threadIdx.x) & 1u)) {
        uint32_t* pad = reinterpret_cast<uint32_t*>(P.pads.p[threadIdx.x]);
        if (P.dynamic_weights) pad[PAD_NI + P.rank] = __float_as_uint(P.my_weight);
        fence_acq_rel_sys();
        st_release_sys_u32(pad + PAD_UPLOAD + P.rank, P.epoch);
    }

    if (reducer) {
        // (2) wait for every participating contributor
        if (threadIdx.x == 0) { s_ok = 1; s_miss = 0u; }
        __syncthreads();
        if (threadIdx.x < P.world && ((P.live_mask >> threadIdx.x) & 1u)) {
            if (!spin_wait_ge(my_pad + PAD_UPLOAD + threadIdx.x, P.epoch, P.timeout_cycles, my_pad + PAD_ABORT)) {
                atomicExch(&s_ok, 0);
                atomicOr(&s_miss, 1u << threadIdx.x);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_ok) {
            float tot = 0.f, first = -1.f;
            bool equal = true, all = true;
            for (int p = 0; p < V6_MAX_PEERS; ++p) {
                float w = 0.f;
                if (p < P.world && ((P.live_mask >> p) & 1u))
                    w = P.dynamic_weights ? __uint_as_float(ld_relaxed_sys_u32(my_pad + PAD_NI + p)) : P.weight[p];
                if (!(w > 0.f)) w = 0.f;
                s_w[p] = w;
                if (p < P.world) {
                    tot += w;
                    if (w <= 0.f) all = false;
                    if (first < 0.f) first = w; else if (w != first) equal = false;
                }
            }
            // the in-switch reduction sums what every bound GPU holds: usable when every rank reports and either the
            // contributions are already multiplied by n_i (delta modes) or all n_i are equal (mean = sum / world)
            const bool mc = P.upload_mc != nullptr && all && (P.upload_prescaled || equal);
            s_use_mc = mc ? 1 : 0;
            s_inv = (mc && !P.upload_prescaled) ? 1.f / (float)P.world : (tot > 0.f ? 1.f / tot : 0.f);
        }
        __syncthreads();
        if (!s_ok) {
            // a contributor never arrived (dead / stopped node): nothing is reduced or pushed by this CTA; the status
            // word and the set of missing ranks are reported to the host and, in step (6), to every peer
            if (threadIdx.x == 0) { my_pad[PAD_STATUS] = 1; atomicOr(my_pad + PAD_MISSING, s_miss); }
        } else {
            // (3)-(5) reduce -> optimizer -> push
            // U vectors per thread per iteration, all peer loads issued before any is consumed:
            // U x (world-1) remote 16 B loads in flight per thread cover the ~2-3 us NVLink
            // round trip (bytes in flight per SM = 512 thr x U x (world-1) x 16 B).
            const long long nvec = (P.hi - P.lo) / VN;
            const long long stride = (long long)gridDim.x * blockDim.x;
            for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * U) {
                float accs[U][VN];
                if (s_use_mc) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const long long iu = i0 + u * stride;
                        if (iu < nvec) load_contrib_mc(reinterpret_cast<const UpT*>(P.upload_mc), P.lo + iu * VN, accs[u]);
                    }
                } else {
                    float v[U][WMAX][VN];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const long long iu = i0 + u * stride;
#pragma unroll
                        for (int p = 0; p < WMAX; ++p)
                            if (iu < nvec && p < P.world && s_w[p] > 0.f)
                                load_contrib(reinterpret_cast<const UpT*>(P.upload.p[p]), P.lo + iu * VN, v[u][p]);
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
#pragma unroll
                        for (int k = 0; k < VN; ++k) accs[u][k] = 0.f;
#pragma unroll
                        for (int p = 0; p < WMAX; ++p)
                            if (p < P.world && s_w[p] > 0.f) {
                                const float s = P.upload_prescaled ? 1.f : s_w[p];
#pragma unroll
                                for (int k = 0; k < VN; ++k) accs[u][k] = fmaf(s, v[u][p][k], accs[u][k]);
                            }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                const long long iu = i0 + u * stride;
                if (iu >= nvec) continue;
                const long long e = P.lo + iu * VN;
                float (&acc)[VN] = accs[u];
                // VN elements of master state (fp32): 1 or 2 float4
#pragma unroll
                for (int h = 0; h < VN / 4; ++h) {
                    const long long eh = e + 4 * h;
                    float4 wg = *reinterpret_cast<const float4*>(P.w_global + eh);
                    float w[4] = {wg.x, wg.y, wg.z, wg.w};
                    float d[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float mean = acc[4 * h + k] * s_inv;
                        d[k] = P.upload_is_delta ? mean : (mean - w[k]);   // pseudo-gradient (ascent dir)
                    }
                    if (P.server_opt == 1) {
                        float4 m4 = *reinterpret_cast<const float4*>(P.opt_m + eh);
                        float m[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) { m[k] = fmaf(P.beta1, m[k], d[k]); w[k] = fmaf(P.server_lr, m[k], w[k]); }
                        *reinterpret_cast<float4*>(P.opt_m + eh) = make_float4(m[0], m[1], m[2], m[3]);
                    } else if (P.server_opt == 2) {
                        float4 m4 = *reinterpret_cast<const float4*>(P.opt_m + eh);
                        float4 v4 = *reinterpret_cast<const float4*>(P.opt_v + eh);
                        float m[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            m[k] = fmaf(P.beta1, m[k], (1.f - P.beta1) * d[k]);
                            vv[k] = fmaf(P.beta2, vv[k], (1.f - P.beta2) * d[k] * d[k]);
                            const float mh = m[k] * P.bias1, vh = vv[k] * P.bias2;
                            w[k] = fmaf(P.server_lr, mh / (sqrtf(vh) + P.eps), w[k]);
                        }
                        *reinterpret_cast<float4*>(P.opt_m + eh) = make_float4(m[0], m[1], m[2], m[3]);
                        *reinterpret_cast<float4*>(P.opt_v + eh) = make_float4(vv[0], vv[1], vv[2], vv[3]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) w[k] = fmaf(P.server_lr, d[k], w[k]);
                    }
                    const float4 wn = make_float4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<float4*>(P.w_global + eh) = wn;
                    // (5) broadcast fused into the epilogue
                    if (P.param_mc) {
                        multimem_st_f4(reinterpret_cast<float4*>(reinterpret_cast<float*>(P.param_mc) + eh), wn);
                    } else {
#pragma unroll
                        for (int p = 0; p < V6_MAX_PEERS; ++p)
                            if (p < P.world && ((P.live_mask >> p) & 1u))
                                st_f4(reinterpret_cast<float4*>(reinterpret_cast<float*>(P.param_out.p[p]) + eh), wn);
                    }
                    acc[4 * h + 0] = w[0]; acc[4 * h + 1] = w[1]; acc[4 * h + 2] = w[2]; acc[4 * h + 3] = w[3];
                }
                if (P.shadow_out.p[0] != nullptr || P.shadow_mc != nullptr) {
                    // bf16 shadow copy of the new global for bf16 compute paths (VN elements).  Elements in the K1 range stay
                    // on the owner: the first GEMM that consumes them multicasts them tile by tile (gemm.cu, fused_bcast == 2)
                    const bool local_only = e >= P.shadow_skip_lo && e < P.shadow_skip_hi;
                    if constexpr (VN == 8) {
                        uint4 s = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                                             pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
                        if (local_only) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(P.shadow_out.p[P.rank]) + e) = s;
                        else if (P.shadow_mc) multimem_st_u4(reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(P.shadow_mc) + e), s);
                        else {
#pragma unroll
                            for (int p = 0; p < V6_MAX_PEERS; ++p)
                                if (p < P.world && ((P.live_mask >> p) & 1u))
                                    st_u4(reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(P.shadow_out.p[p]) + e), s);
                        }
                    } else {
                        uint2 s = make_uint2(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]));
#pragma unroll
                        for (int p = 0; p < V6_MAX_PEERS; ++p)
                            if (p < P.world && ((P.live_mask >> p) & 1u) && (!local_only || p == P.rank))
                                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(P.shadow_out.p[p]) + e) = s;
                    }
                }
                }   // u
            }
        }
    }

    // (6) completion: last CTA signals "my slice is pushed" and waits for all other reducers
    __shared__ unsigned int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(P.cta_counter, 1u);
        s_last = (t == gridDim.x - 1) ? 1u : 0u;
        if (s_last) *P.cta_counter = 0u;        // self-reset for the next launch
    }
    __syncthreads();
    if (s_last) {
        if (is_red && threadIdx.x < P.world && ((P.live_mask >> threadIdx.x) & 1u)) {
            uint32_t* pad = reinterpret_cast<uint32_t*>(P.pads.p[threadIdx.x]);
            // a failed round (some CTA of this reducer timed out in step 2) is announced to every live peer together
            // with the contributors that were missed, BEFORE the completion flag: nobody leaves the round believing
            // that a slice was pushed when it was not.
            const uint32_t failed = reducer ? ld_relaxed_sys_u32(my_pad + PAD_STATUS) : 0u;
            if (failed == 1u) {
                pad[PAD_RMISS + P.rank] = ld_relaxed_sys_u32(my_pad + PAD_MISSING);
                pad[PAD_RFAIL + P.rank] = P.epoch;
            }
            fence_acq_rel_sys();
            st_release_sys_u32(pad + PAD_BCAST + P.rank, P.epoch);
        }
        if (threadIdx.x < P.world && ((P.reducer_mask >> threadIdx.x) & (P.live_mask >> threadIdx.x) & 1u)) {
            if (!spin_wait_ge(my_pad + PAD_BCAST + threadIdx.x, P.epoch, P.timeout_cycles, my_pad + PAD_ABORT)) {
                my_pad[PAD_STATUS] = 1;
                atomicOr(my_pad + PAD_MISSING, 1u << threadIdx.x);
            } else if (ld_relaxed_sys_u32(my_pad + PAD_RFAIL + threadIdx.x) == P.epoch) {
                atomicCAS(my_pad + PAD_STATUS, 0u, 2u);
                atomicOr(my_pad + PAD_MISSING, ld_relaxed_sys_u32(my_pad + PAD_RMISS + threadIdx.x));
            }
        }
    }
}

extern "C" int v6_fedavg_round(const FedAvgParams* hp, int upload_dtype /*0 f32, 1 bf16*/, int grid,
                               cudaStream_t stream) {
    FedAvgParams P = *hp;
    if (grid <= 0) grid = 148;
    // few peers -> deeper unroll (keep ~8 remote loads in flight per thread)
#define V6_LAUNCH(T) \
    do { \
        if (P.world <= 2) fedavg_round_kernel<T, 4, 2><<<grid, 512, 0, stream>>>(P); \
        else if (P.world <= 4) fedavg_round_kernel<T, 2, 4><<<grid, 512, 0, stream>>>(P); \
        else fedavg_round_kernel<T, 1, 8><<<grid, 512, 0, stream>>>(P); \
    } while (0)
    if (upload_dtype == 0) V6_LAUNCH(float);
    else V6_LAUNCH(__nv_bfloat16);
#undef V6_LAUNCH
    V6_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------
// Cross-GPU barrier kernel (bring-up, tests, and round boundaries of the NCCL-free path)
// ----------------------------------------------------------------------------------------
__global__ void symm_barrier_kernel(PeerTable pads, int rank, int world, uint32_t epoch,
                                    long long timeout_cycles) {
    uint32_t* my_pad = reinterpret_cast<uint32_t*>(pads.p[rank]);
    if (threadIdx.x < world) {
        uint32_t* pad = reinterpret_cast<uint32_t*>(pads.p[threadIdx.x]);
        fence_acq_rel_sys();
        st_release_sys_u32(pad + PAD_BARRIER + rank, epoch);
        if (!spin_wait_ge(my_pad + PAD_BARRIER + threadIdx.x, epoch, timeout_cycles, my_pad + PAD_ABORT))
            my_pad[PAD_STATUS] = 1;
    }
}
extern "C" int v6_symm_barrier(const PeerTable* pads, int rank, int world, uint32_t epoch,
                               long long timeout_cycles, cudaStream_t stream) {
    symm_barrier_kernel<<<1, 32, 0, stream>>>(*pads, rank, world, epoch, timeout_cycles);
    V6_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------
// K3: small-message one-shot weighted all-reduce (+ optional server optimizer step).
//   every rank: slot[rank] holds its payload (n floats, n*4 <= 64 KB) in symmetric memory.
//   out[e] = sum_p weight[p] * slot_p[e] * inv_total       (written locally on every rank)
// Single CTA: signal -> wait -> P2P loads from all peers -> reduce -> local store.
// Double-buffered by epoch parity on the host side, so no trailing barrier is needed.
// ----------------------------------------------------------------------------------------

__global__ void __launch_bounds__(1024, 1) small_allreduce_kernel(const SmallParams P) {
    __shared__ int s_ok;
    uint32_t* my_pad = reinterpret_cast<uint32_t*>(P.pads.p[P.rank]);
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (threadIdx.x < P.world) {
        uint32_t* pad = reinterpret_cast<uint32_t*>(P.pads.p[threadIdx.x]);
        fence_acq_rel_sys();
        st_release_sys_u32(pad + PAD_SMALL + P.rank, P.epoch);
        if (P.weight[threadIdx.x] > 0.f &&
            !spin_wait_ge(my_pad + PAD_SMALL + threadIdx.x, P.epoch, P.timeout_cycles, my_pad + PAD_ABORT))
            atomicExch(&s_ok, 0);
    }
    __syncthreads();
    if (!s_ok) { if (threadIdx.x == 0) my_pad[PAD_STATUS] = 1; return; }
    for (int i = threadIdx.x; i < P.n / 4; i += blockDim.x) {
        float4 v[V6_MAX_PEERS];
#pragma unroll
        for (int p = 0; p < V6_MAX_PEERS; ++p)
            if (p < P.world && P.weight[p] > 0.f)
                v[p] = ld_sys_f4(reinterpret_cast<const float4*>(P.slots.p[p]) + i);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < V6_MAX_PEERS; ++p)
            if (p < P.world && P.weight[p] > 0.f) {
                a.x = fmaf(P.weight[p], v[p].x, a.x); a.y = fmaf(P.weight[p], v[p].y, a.y);
                a.z = fmaf(P.weight[p], v[p].z, a.z); a.w = fmaf(P.weight[p], v[p].w, a.w);
            }
        a.x *= P.inv_total; a.y *= P.inv_total; a.z *= P.inv_total; a.w *= P.inv_total;
        reinterpret_cast<float4*>(P.out)[i] = a;
    }
}
// ----------------------------------------------------------------------------------------
// K8 + K3 fused tail of a federated GLM iteration (default; V6B200_GLM_FUSED=0 = composed path; 58.4 vs 74.0 us at 2 GPUs):
//   fold the per-CTA partials of the gradient kernel into this rank's payload slot  (was: fold kernel)
//   -> signal / wait / P2P loads of every node's payload, sum                        (was: K3)
//   -> w -= lr * g / n,  loss = l / n                                               (was: three PyTorch kernels)
// in ONE single-CTA launch.  payload = [g_w (F), g_b, loss, n_rows | pad]; P.n = padded payload length (<= 1024).
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) glm_aggregate_update_kernel(const SmallParams P, const float* __restrict__ part,
                                                                       int nparts, int F, float rows, float lr,
                                                                       float* __restrict__ w, float* __restrict__ loss_out) {
    __shared__ int s_ok;
    __shared__ float tot[1024];
    uint32_t* my_pad = reinterpret_cast<uint32_t*>(P.pads.p[P.rank]);
    float* my_slot = reinterpret_cast<float*>(P.slots.p[P.rank]);
    if (threadIdx.x == 0) s_ok = 1;
    // 1. fold the gradient kernel's partials into this rank's slot (fixed order: deterministic)
    for (int i = threadIdx.x; i < P.n; i += blockDim.x) {
        float a = 0.f;
        if (i < F + 2) {
            for (int p = 0; p < nparts; ++p) a += part[(size_t)p * (F + 2) + i];
        } else if (i == F + 2) {
            a = rows;
        }
        my_slot[i] = a;
    }
    __syncthreads();
    // 2. publish / wait / reduce (as small_allreduce_kernel)
    if (threadIdx.x < P.world) {
        uint32_t* pad = reinterpret_cast<uint32_t*>(P.pads.p[threadIdx.x]);
        fence_acq_rel_sys();
        st_release_sys_u32(pad + PAD_SMALL + P.rank, P.epoch);
        if (P.weight[threadIdx.x] > 0.f &&
            !spin_wait_ge(my_pad + PAD_SMALL + threadIdx.x, P.epoch, P.timeout_cycles, my_pad + PAD_ABORT))
            atomicExch(&s_ok, 0);
    }
    __syncthreads();
    if (!s_ok) { if (threadIdx.x == 0) my_pad[PAD_STATUS] = 1; return; }
    for (int i = threadIdx.x; i < P.n / 4; i += blockDim.x) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < V6_MAX_PEERS; ++p)
            if (p < P.world && P.weight[p] > 0.f) {
                const float4 v = ld_sys_f4(reinterpret_cast<const float4*>(P.slots.p[p]) + i);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
        reinterpret_cast<float4*>(P.out)[i] = a;                    // the summed payload stays readable by the host side
        tot[4 * i] = a.x; tot[4 * i + 1] = a.y; tot[4 * i + 2] = a.z; tot[4 * i + 3] = a.w;
    }
    __syncthreads();
    // 3. the gradient step on this rank's copy of the coefficients
    const float inv_n = 1.f / tot[F + 2];
    for (int j = threadIdx.x; j <= F; j += blockDim.x) w[j] -= lr * tot[j] * inv_n;
    if (threadIdx.x == 0) *loss_out = tot[F + 1] * inv_n;
}
extern "C" int v6_glm_aggregate_update(const SmallParams* hp, const float* part, int nparts, int F, float rows, float lr, float* w,
                                       float* loss_out, cudaStream_t stream) {
    if (hp->n > 1024 || hp->n % 4 != 0 || F + 3 > hp->n || nparts < 1) return (int)cudaErrorInvalidValue;
    glm_aggregate_update_kernel<<<1, 1024, 0, stream>>>(*hp, part, nparts, F, rows, lr, w, loss_out);
    V6_CHECK_LAUNCH();
    return 0;
}

extern "C" int v6_small_allreduce(const SmallParams* hp, cudaStream_t stream) {
    small_allreduce_kernel<<<1, 1024, 0, stream>>>(*hp);
    V6_CHECK_LAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------------------
// Plain NVLink copy kernels used for bandwidth bring-up ("bus GB/s vs 900"):
//   pull: dst_local[i] = src_peer[i];   push_mc: multimem.st of a local buffer to all peers.
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 2) p2p_pull_kernel(const float4* __restrict__ src, float4* __restrict__ dst, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x * 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { long long j = i + (long long)u * gridDim.x * blockDim.x; if (j < n4) v[u] = ld_sys_f4(src + j); }
#pragma unroll
        for (int u = 0; u < 4; ++u) { long long j = i + (long long)u * gridDim.x * blockDim.x; if (j < n4) dst[j] = v[u]; }
    }
}
__global__ void __launch_bounds__(512, 2) mc_push_kernel(const float4* __restrict__ src, float4* mc_dst, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        multimem_st_f4(mc_dst + i, src[i]);
}
__global__ void __launch_bounds__(512, 2) mc_reduce_kernel(const float4* mc_src, float4* __restrict__ dst, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        dst[i] = multimem_ld_reduce_add_f4(mc_src + i);
}
extern "C" int v6_p2p_pull(const void* src_peer, void* dst_local, long long nbytes, cudaStream_t s) {
    p2p_pull_kernel<<<148 * 2, 512, 0, s>>>((const float4*)src_peer, (float4*)dst_local, nbytes / 16);
    V6_CHECK_LAUNCH(); return 0;
}
extern "C" int v6_mc_push(const void* src_local, void* mc_dst, long long nbytes, cudaStream_t s) {
    mc_push_kernel<<<148 * 2, 512, 0, s>>>((const float4*)src_local, (float4*)mc_dst, nbytes / 16);
    V6_CHECK_LAUNCH(); return 0;
}
extern "C" int v6_mc_reduce(const void* mc_src, void* dst_local, long long nbytes, cudaStream_t s) {
    mc_reduce_kernel<<<148 * 2, 512, 0, s>>>((const float4*)mc_src, (float4*)dst_local, nbytes / 16);
    V6_CHECK_LAUNCH(); return 0;
}
