// K4 backward: flash-attention gradients on tcgen05 / TMEM / TMA (sm_100a), transpose-free MMA form.
//
// Two kernels (S and dP are recomputed in both -- tensor-core FLOPs are cheap, shared-memory
// transposes are not):
//
//  dQ kernel   (CTA = 128 query rows of one head; loop over KV tiles j)      thread == query row
//      S  = Q K_j^T            (SS)        dP = dO V_j^T          (SS)          -> TMEM
//      dS = P o (dP - delta) * scale,  P = exp2(S*c - lse2)      registers -> bf16 over S's columns
//      dQ += dS K_j            (TS: A = dS from TMEM, B = K_j^T tile, keys contiguous)
//
//  dK/dV kernel (CTA = 128 keys of one KV head; loop over the q-heads of the group x Q tiles i)
//                                                                              thread == key row
//      S^T  = K_j Q_i^T        (SS)        dP^T = V_j dO_i^T      (SS)          -> TMEM
//      P^T  = exp2(S^T*c - lse2[q]),  dS^T = P^T o (dP^T - delta[q]) * scale   (per-column lse/delta)
//      dV += P^T dO_i          (TS, B = dO_i^T tile, queries contiguous)
//      dK += dS^T Q_i          (TS, B = Q_i^T tile)
//
// Pre-transposed operands (Kt, Qt, dOt: [B,H,D,S]) make every B operand K-major, exactly like Vt in
// the forward kernel.  lse2 = lse * log2(e) and delta = rowsum(dO o O) are [B,Hq,S] fp32.
// GQA needs no atomics: the CTA that owns a KV tile walks all query heads of its group.
#include <cuda.h>
#include "common.cuh"
#include "api.h"

namespace attn_bwd {

constexpr int BM = 128, BN = 128, kThreads = 256, kTmemCols = 512;

struct Params {
    __nv_bfloat16* dq;            // [B,S,Hq,D]
    __nv_bfloat16* dk;            // [B,S,Hkv,D]
    __nv_bfloat16* dv;            // [B,S,Hkv,D]
    const float* lse2;            // [B,Hq,S]
    const float* delta;           // [B,Hq,S]
    int B, S, Hq, Hkv;
    float scale_log2, scale;
};

V6_DEVINL void store_row_bf16(__nv_bfloat16* dst, const uint32_t (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 32; i += 8)
        *reinterpret_cast<uint4*>(dst + i) = make_uint4(
            pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])),
            pack_bf16x2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])), pack_bf16x2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
}

// ============================================================================== dQ kernel
template <int D, bool CAUSAL>
__global__ void __launch_bounds__(kThreads, 1)
flash_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmap_q,      // [B*S, Hq*D]   box 128 x 64
                    const __grid_constant__ CUtensorMap tmap_do,     // [B*S, Hq*D]   box 128 x 64
                    const __grid_constant__ CUtensorMap tmap_k,      // [B*S, Hkv*D]  box 128 x 64
                    const __grid_constant__ CUtensorMap tmap_v,      // [B*S, Hkv*D]  box 128 x 64
                    const __grid_constant__ CUtensorMap tmap_kt,     // [B*Hkv*D, S]  box D x 64
                    const Params P) {
    constexpr int NH = D / 64;
    constexpr int T_BYTES = BM * D * 2;                  // one 128 x D (or D x 128) bf16 tile
    constexpr int S_COL = 0, DP_COL = 128, DQ_COL = 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-B aligned; derived by pointer arithmetic so that the compiler keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* sQ = smem;
    uint8_t* sdO = sQ + T_BYTES;
    uint8_t* sK = sdO + T_BYTES;
    uint8_t* sV = sK + T_BYTES;
    uint8_t* sKt = sV + T_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sKt + T_BYTES);
    uint64_t* qdo_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = bars + 2;
    uint64_t* sdp_full = bars + 3;
    uint64_t* ds_full = bars + 4;
    uint64_t* dq_done = bars + 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_qt = (P.S + BM - 1) / BM;
    const int qt = n_qt - 1 - blockIdx.x;
    const int h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (P.Hq / P.Hkv);
    const int m0 = qt * BM;
    const int n_kv_all = (P.S + BN - 1) / BN;
    const int nkv = CAUSAL ? min(n_kv_all, qt + 1) : n_kv_all;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_do); tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_kt);
    }
    if (warp == 5 && lane == 0) {
        mbar_init(qdo_full, 1); mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
        mbar_init(sdp_full, 1); mbar_init(ds_full, 4); mbar_init(dq_done, 1);
        mbar_fence_init();
    }
    if (warp == 6) tmem_alloc<kTmemCols>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(qdo_full, 2 * T_BYTES);
#pragma unroll
            for (int hh = 0; hh < NH; ++hh) {
                tma_load_2d(sQ + hh * (BM * 128), &tmap_q, qdo_full, h * D + hh * 64, b * P.S + m0);
                tma_load_2d(sdO + hh * (BM * 128), &tmap_do, qdo_full, h * D + hh * 64, b * P.S + m0);
            }
            for (int j = 0; j < nkv; ++j) {
                mbar_wait(kv_empty, (j & 1) ^ 1);
                mbar_expect_tx(kv_full, 3 * T_BYTES);
#pragma unroll
                for (int hh = 0; hh < NH; ++hh) {
                    tma_load_2d(sK + hh * (BN * 128), &tmap_k, kv_full, hk * D + hh * 64, b * P.S + j * BN);
                    tma_load_2d(sV + hh * (BN * 128), &tmap_v, kv_full, hk * D + hh * 64, b * P.S + j * BN);
                }
#pragma unroll
                for (int kh = 0; kh < 2; ++kh)
                    tma_load_2d(sKt + kh * (D * 128), &tmap_kt, kv_full, j * BN + kh * 64, (b * P.Hkv + hk) * D);
            }
        }
    } else if (warp == 5) {
        constexpr uint32_t idesc_s = make_idesc_bf16(BM, BN);
        constexpr uint32_t idesc_q = make_idesc_bf16(BM, D);
        mbar_wait(qdo_full, 0);
        for (int j = 0; j < nkv; ++j) {
            mbar_wait(kv_full, j & 1);
            tcgen05_fence_after();
            if (lane == 0) {
                const uint32_t qa = smem_u32(sQ), da = smem_u32(sdO), kb = smem_u32(sK), vb = smem_u32(sV);
#pragma unroll
                for (int hh = 0; hh < NH; ++hh)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_bf16_ss(tmem_base + S_COL, make_smem_desc_sw128(qa + hh * (BM * 128) + k * 32),
                                     make_smem_desc_sw128(kb + hh * (BN * 128) + k * 32), idesc_s, (hh | k) ? 1u : 0u);
                        umma_bf16_ss(tmem_base + DP_COL, make_smem_desc_sw128(da + hh * (BM * 128) + k * 32),
                                     make_smem_desc_sw128(vb + hh * (BN * 128) + k * 32), idesc_s, (hh | k) ? 1u : 0u);
                    }
                umma_commit(sdp_full);
            }
            __syncwarp();
            mbar_wait(ds_full, j & 1);
            tcgen05_fence_after();
            if (lane == 0) {
                const uint32_t ktb = smem_u32(sKt);
#pragma unroll
                for (int kk = 0; kk < BN / 16; ++kk)
                    umma_bf16_ts(tmem_base + DQ_COL, tmem_base + S_COL + kk * 8,
                                 make_smem_desc_sw128(ktb + (kk >> 2) * (D * 128) + (kk & 3) * 32), idesc_q,
                                 (j > 0 || kk > 0) ? 1u : 0u);
                umma_commit(kv_empty);
                umma_commit(dq_done);
            }
            __syncwarp();
        }
    } else if (warp < 4) {
        const int ew = warp;                                  // compute warps are 0-3: the role warps sit in the highest ids (issue priority)
        const int row = m0 + ew * 32 + lane;
        const bool row_ok = row < P.S;
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        const size_t ridx = ((size_t)b * P.Hq + h) * P.S + (row_ok ? row : 0);
        const float lse2 = row_ok ? P.lse2[ridx] : INFINITY;
        const float delta = row_ok ? P.delta[ridx] : 0.f;
        for (int j = 0; j < nkv; ++j) {
            mbar_wait(sdp_full, j & 1);
            tcgen05_fence_after();
            const int key0 = j * BN;
            const bool need_mask = (key0 + BN > P.S) || (CAUSAL && key0 + BN - 1 > m0);
#pragma unroll 1
            for (int c2 = 0; c2 < BN; c2 += 64) {
                uint32_t pk[32];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int c = c2 + half * 32;
                    uint32_t sv[32], dv[32];
                    tmem_ld_32x32b_x32(tmem_base + lane_addr + S_COL + c, sv);
                    tmem_ld_32x32b_x32(tmem_base + lane_addr + DP_COL + c, dv);
                    tmem_ld_wait();
                    float ds[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float p = exp2f(fmaf(__uint_as_float(sv[i]), P.scale_log2, -lse2));
                        if (need_mask) {
                            const int key = key0 + c + i;
                            if (!(key < P.S && (!CAUSAL || key <= row))) p = 0.f;
                        }
                        ds[i] = p * (__uint_as_float(dv[i]) - delta) * P.scale;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[half * 16 + i] = pack_bf16x2(ds[2 * i], ds[2 * i + 1]);
                }
                tmem_st_32x32b_x32(tmem_base + lane_addr + S_COL + (c2 >> 1), pk);
            }
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ds_full);
        }
        mbar_wait(dq_done, (nkv - 1) & 1);
        tcgen05_fence_after();
        __nv_bfloat16* dst = P.dq + (((size_t)b * P.S + (row_ok ? row : 0)) * P.Hq + h) * D;
#pragma unroll 1
        for (int c = 0; c < D; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + DQ_COL + c, v);
            tmem_ld_wait();
            if (row_ok) store_row_bf16(dst + c, v);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 6) { tcgen05_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}

// ============================================================================== dK / dV kernel
template <int D, bool CAUSAL>
__global__ void __launch_bounds__(kThreads, 1)
flash_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmap_q,     // [B*S, Hq*D]   box 128 x 64
                     const __grid_constant__ CUtensorMap tmap_do,    // [B*S, Hq*D]   box 128 x 64
                     const __grid_constant__ CUtensorMap tmap_k,     // [B*S, Hkv*D]  box 128 x 64
                     const __grid_constant__ CUtensorMap tmap_v,     // [B*S, Hkv*D]  box 128 x 64
                     const __grid_constant__ CUtensorMap tmap_qt,    // [B*Hq*D, S]   box D x 64
                     const __grid_constant__ CUtensorMap tmap_dot,   // [B*Hq*D, S]   box D x 64
                     const Params P) {
    constexpr int NH = D / 64;
    constexpr int T_BYTES = BM * D * 2;
    constexpr int ST_COL = 0, DPT_COL = 128, DV_COL = 256, DK_COL = 256 + D;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-B aligned; derived by pointer arithmetic so that the compiler keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* sK = smem;
    uint8_t* sV = sK + T_BYTES;
    uint8_t* sQ = sV + T_BYTES;
    uint8_t* sdO = sQ + T_BYTES;
    uint8_t* sQt = sdO + T_BYTES;
    uint8_t* sdOt = sQt + T_BYTES;
    float* sLse = reinterpret_cast<float*>(sdOt + T_BYTES);       // [128] lse2 of the current q tile
    float* sDelta = sLse + BM;                                     // [128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + BM);
    uint64_t* kv_full = bars;
    uint64_t* q_full = bars + 1;
    uint64_t* q_empty = bars + 2;
    uint64_t* st_full = bars + 3;
    uint64_t* pt_full = bars + 4;
    uint64_t* acc_done = bars + 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jt = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
    const int group = P.Hq / P.Hkv;
    const int n_qt = (P.S + BM - 1) / BM;
    const int i0 = CAUSAL ? jt : 0;
    const int n_i = n_qt - i0;                          // q tiles that see this KV tile
    const int n_iter = group * n_i;
    const int k0 = jt * BN;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_do); tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_qt); tma_prefetch_desc(&tmap_dot);
    }
    if (warp == 5 && lane == 0) {
        mbar_init(kv_full, 1); mbar_init(q_full, 1); mbar_init(q_empty, 1);
        mbar_init(st_full, 1); mbar_init(pt_full, 4); mbar_init(acc_done, 1);
        mbar_fence_init();
    }
    if (warp == 6) tmem_alloc<kTmemCols>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(kv_full, 2 * T_BYTES);
#pragma unroll
            for (int hh = 0; hh < NH; ++hh) {
                tma_load_2d(sK + hh * (BN * 128), &tmap_k, kv_full, hk * D + hh * 64, b * P.S + k0);
                tma_load_2d(sV + hh * (BN * 128), &tmap_v, kv_full, hk * D + hh * 64, b * P.S + k0);
            }
            for (int t = 0; t < n_iter; ++t) {
                const int hq = hk * group + t / n_i, i = i0 + t % n_i;
                mbar_wait(q_empty, (t & 1) ^ 1);
                mbar_expect_tx(q_full, 4 * T_BYTES);
#pragma unroll
                for (int hh = 0; hh < NH; ++hh) {
                    tma_load_2d(sQ + hh * (BM * 128), &tmap_q, q_full, hq * D + hh * 64, b * P.S + i * BM);
                    tma_load_2d(sdO + hh * (BM * 128), &tmap_do, q_full, hq * D + hh * 64, b * P.S + i * BM);
                }
#pragma unroll
                for (int qh = 0; qh < 2; ++qh) {
                    tma_load_2d(sQt + qh * (D * 128), &tmap_qt, q_full, i * BM + qh * 64, (b * P.Hq + hq) * D);
                    tma_load_2d(sdOt + qh * (D * 128), &tmap_dot, q_full, i * BM + qh * 64, (b * P.Hq + hq) * D);
                }
            }
        }
    } else if (warp == 5) {
        constexpr uint32_t idesc_s = make_idesc_bf16(BN, BM);
        constexpr uint32_t idesc_d = make_idesc_bf16(BN, D);
        mbar_wait(kv_full, 0);
        for (int t = 0; t < n_iter; ++t) {
            mbar_wait(q_full, t & 1);
            tcgen05_fence_after();
            if (lane == 0) {
                const uint32_t ka = smem_u32(sK), va = smem_u32(sV), qb = smem_u32(sQ), ob = smem_u32(sdO);
#pragma unroll
                for (int hh = 0; hh < NH; ++hh)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_bf16_ss(tmem_base + ST_COL, make_smem_desc_sw128(ka + hh * (BN * 128) + k * 32),
                                     make_smem_desc_sw128(qb + hh * (BM * 128) + k * 32), idesc_s, (hh | k) ? 1u : 0u);
                        umma_bf16_ss(tmem_base + DPT_COL, make_smem_desc_sw128(va + hh * (BN * 128) + k * 32),
                                     make_smem_desc_sw128(ob + hh * (BM * 128) + k * 32), idesc_s, (hh | k) ? 1u : 0u);
                    }
                umma_commit(st_full);
            }
            __syncwarp();
            mbar_wait(pt_full, t & 1);
            tcgen05_fence_after();
            if (lane == 0) {
                const uint32_t dotb = smem_u32(sdOt), qtb = smem_u32(sQt);
#pragma unroll
                for (int kk = 0; kk < BM / 16; ++kk) {
                    umma_bf16_ts(tmem_base + DV_COL, tmem_base + ST_COL + kk * 8,
                                 make_smem_desc_sw128(dotb + (kk >> 2) * (D * 128) + (kk & 3) * 32), idesc_d, (t > 0 || kk > 0) ? 1u : 0u);
                    umma_bf16_ts(tmem_base + DK_COL, tmem_base + DPT_COL + kk * 8,
                                 make_smem_desc_sw128(qtb + (kk >> 2) * (D * 128) + (kk & 3) * 32), idesc_d, (t > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(q_empty);
                umma_commit(acc_done);
            }
            __syncwarp();
        }
    } else if (warp < 4) {
        const int ew = warp;                                  // compute warps are 0-3: the role warps sit in the highest ids (issue priority)
        const int tid = ew * 32 + lane;                   // 0..127
        const int key = k0 + tid;
        const bool key_ok = key < P.S;
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        for (int t = 0; t < n_iter; ++t) {
            const int hq = hk * group + t / n_i, i = i0 + t % n_i;
            const int q0 = i * BM;
            // stage the per-query lse2 / delta of this q tile (previous tile's readers are done: they
            // arrived on pt_full before the MMA that precedes st_full of this iteration could retire)
            {
                const int q = q0 + tid;
                const size_t idx = ((size_t)b * P.Hq + hq) * P.S + (q < P.S ? q : 0);
                sLse[tid] = q < P.S ? P.lse2[idx] : INFINITY;
                sDelta[tid] = q < P.S ? P.delta[idx] : 0.f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");          // the 4 softmax warps only
            mbar_wait(st_full, t & 1);
            tcgen05_fence_after();
            const bool need_mask = (q0 + BM > P.S) || (k0 + BN > P.S) || (CAUSAL && k0 + BN - 1 > q0);
#pragma unroll 1
            for (int c2 = 0; c2 < BM; c2 += 64) {
                uint32_t pk[32], dk[32];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int c = c2 + half * 32;
                    uint32_t sv[32], dv[32];
                    tmem_ld_32x32b_x32(tmem_base + lane_addr + ST_COL + c, sv);
                    tmem_ld_32x32b_x32(tmem_base + lane_addr + DPT_COL + c, dv);
                    tmem_ld_wait();
                    float p[32], ds[32];
#pragma unroll
                    for (int x = 0; x < 32; ++x) {
                        float pv = exp2f(fmaf(__uint_as_float(sv[x]), P.scale_log2, -sLse[c + x]));
                        if (need_mask) {
                            const int q = q0 + c + x;
                            if (!(key_ok && q < P.S && (!CAUSAL || key <= q))) pv = 0.f;
                        }
                        p[x] = pv;
                        ds[x] = pv * (__uint_as_float(dv[x]) - sDelta[c + x]) * P.scale;
                    }
#pragma unroll
                    for (int x = 0; x < 16; ++x) {
                        pk[half * 16 + x] = pack_bf16x2(p[2 * x], p[2 * x + 1]);
                        dk[half * 16 + x] = pack_bf16x2(ds[2 * x], ds[2 * x + 1]);
                    }
                }
                tmem_st_32x32b_x32(tmem_base + lane_addr + ST_COL + (c2 >> 1), pk);
                tmem_st_32x32b_x32(tmem_base + lane_addr + DPT_COL + (c2 >> 1), dk);
            }
            tmem_st_wait();
            tcgen05_fence_before();
            asm volatile("bar.sync 1, 128;" ::: "memory");          // everyone is done with sLse / sDelta
            if (lane == 0) mbar_arrive(pt_full);
        }
        mbar_wait(acc_done, (n_iter - 1) & 1);
        tcgen05_fence_after();
        __nv_bfloat16* dvp = P.dv + (((size_t)b * P.S + (key_ok ? key : 0)) * P.Hkv + hk) * D;
        __nv_bfloat16* dkp = P.dk + (((size_t)b * P.S + (key_ok ? key : 0)) * P.Hkv + hk) * D;
#pragma unroll 1
        for (int c = 0; c < D; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + DV_COL + c, v);
            tmem_ld_wait();
            if (key_ok) store_row_bf16(dvp + c, v);
            tmem_ld_32x32b_x32(tmem_base + lane_addr + DK_COL + c, v);
            tmem_ld_wait();
            if (key_ok) store_row_bf16(dkp + c, v);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 6) { tcgen05_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}

template <int D> constexpr int dq_smem() { return 5 * BM * D * 2 + 1024 + 256; }
template <int D> constexpr int dkv_smem() { return 6 * BM * D * 2 + 2 * BM * 4 + 1024 + 256; }

}  // namespace attn_bwd

template <int D, bool CAUSAL>
static int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& tk, const CUtensorMap& tv,
                      const CUtensorMap& tkt, const CUtensorMap& tqt, const CUtensorMap& tdot, const attn_bwd::Params& P,
                      cudaStream_t stream) {
    using namespace attn_bwd;
    auto kq = flash_bwd_dq_kernel<D, CAUSAL>;
    auto kkv = flash_bwd_dkv_kernel<D, CAUSAL>;
    cudaError_t e = cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, dq_smem<D>());
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, dkv_smem<D>());
    if (e != cudaSuccess) return (int)e;
    const int nt = (P.S + BM - 1) / BM;
    kq<<<dim3(nt, P.Hq, P.B), kThreads, dq_smem<D>(), stream>>>(tq, tdo, tk, tv, tkt, P);
    V6_CHECK_LAUNCH();
    kkv<<<dim3(nt, P.Hkv, P.B), kThreads, dkv_smem<D>(), stream>>>(tq, tdo, tk, tv, tqt, tdot, P);
    V6_CHECK_LAUNCH();
    return 0;
}

// q,do:[B,S,Hq,D]  k,v:[B,S,Hkv,D]  kt:[B,Hkv,D,S]  qt,dot:[B,Hq,D,S]  lse2,delta:[B,Hq,S] fp32
extern "C" int v6_flash_attn_bwd(const void* q, const void* k, const void* v, const void* dout, const void* kt, const void* qt,
                                 const void* dot, const float* lse2, const float* delta, void* dq, void* dk, void* dv, int B,
                                 int S, int Hq, int Hkv, int D, float softmax_scale, int causal, cudaStream_t stream) {
    if ((D != 64 && D != 128) || Hq % Hkv != 0 || S % 8 != 0) return (int)cudaErrorInvalidValue;
    alignas(64) CUtensorMap tq, tdo, tk, tv, tkt, tqt, tdot;
    const uint64_t BS = (uint64_t)B * S;
    if (v6_make_tmap_2d_bf16(&tq, (uint64_t)q, BS, (uint64_t)Hq * D, (uint64_t)Hq * D * 2, 128, 64, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tdo, (uint64_t)dout, BS, (uint64_t)Hq * D, (uint64_t)Hq * D * 2, 128, 64, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tk, (uint64_t)k, BS, (uint64_t)Hkv * D, (uint64_t)Hkv * D * 2, 128, 64, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tv, (uint64_t)v, BS, (uint64_t)Hkv * D, (uint64_t)Hkv * D * 2, 128, 64, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tkt, (uint64_t)kt, (uint64_t)B * Hkv * D, (uint64_t)S, (uint64_t)S * 2, (uint32_t)D, 64, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tqt, (uint64_t)qt, (uint64_t)B * Hq * D, (uint64_t)S, (uint64_t)S * 2, (uint32_t)D, 64, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tdot, (uint64_t)dot, (uint64_t)B * Hq * D, (uint64_t)S, (uint64_t)S * 2, (uint32_t)D, 64, 1)) return -2;
    attn_bwd::Params P;
    P.dq = (__nv_bfloat16*)dq; P.dk = (__nv_bfloat16*)dk; P.dv = (__nv_bfloat16*)dv; P.lse2 = lse2; P.delta = delta;
    P.B = B; P.S = S; P.Hq = Hq; P.Hkv = Hkv; P.scale = softmax_scale; P.scale_log2 = softmax_scale * 1.4426950408889634f;
    if (D == 64) return causal ? launch_bwd<64, true>(tq, tdo, tk, tv, tkt, tqt, tdot, P, stream)
                               : launch_bwd<64, false>(tq, tdo, tk, tv, tkt, tqt, tdot, P, stream);
    return causal ? launch_bwd<128, true>(tq, tdo, tk, tv, tkt, tqt, tdot, P, stream)
                  : launch_bwd<128, false>(tq, tdo, tk, tv, tkt, tqt, tdot, P, stream);
}
