// K4 forward, second design: TWO co-resident CTAs per SM instead of an intra-CTA software pipeline.
//
// The first kernel (attention.cu) keeps one 128-query tile per SM and overlaps S_{j+1} = Q K^T with the
// softmax of tile j; its single softmax warpgroup (one warp per SM sub-partition, no latency hiding) is the
// bottleneck: 563 TFLOP/s non-causal, 218-251 causal (profiles/kernel_bench_r1b.json).  Here a CTA is
// deliberately serial -- S_j -> softmax_j -> O += P_j V_j -- and small enough that two CTAs share an SM:
//
//   TMEM  256 columns per CTA : S (128 fp32 columns; P_j, packed bf16, overwrites the first 64 in place)
//                               | O (D columns)                               -> 2 x 256 = the SM's 512
//   smem  Q + ONE K tile + ONE Vt tile (96 KB at D=128)                       -> 2 x 97 KB <= 227 KB
//   regs  <= 128 per thread (__launch_bounds__(256, 2))
//
// so while one CTA's softmax warpgroup works (two warps per sub-partition now hide each other's TMEM-load
// and MUFU latency) the other CTA's MMAs own the tensor pipe; the hardware CTA scheduler interleaves them.
// K_{j+1} streams in during softmax_j / PV_j, Vt_{j+1} during S_{j+1} / softmax_{j+1} (separate full/empty
// mbarriers per operand).  Softmax per tile: pass 1 = exact row max from TMEM (no optimistic guess needed
// any more because P aliases S), lazy rescale of O only when the running max grows by > 8 (log2 units),
// pass 2 = exp2 / row sum / bf16 pack / tcgen05.st.  Same interface and numerics contract as attention.cu.
#include <cuda.h>
#include "common.cuh"
#include "api.h"

namespace attn2 {

constexpr int BM = 128;          // query rows per CTA
constexpr int BN = 128;          // keys per KV tile
constexpr int kThreads = 256;
constexpr int kTmemCols = 256;
constexpr int S_COL = 0, O_COL = 128;

struct Params {
    __nv_bfloat16* O;     // [B,S,Hq,D]
    float* lse;           // [B,Hq,S]
    int B, S, Hq, Hkv;
    float scale_log2;     // softmax_scale * log2(e)
};

// MN-major SWIZZLE_128B operand descriptor (see glm_tc.cu / docs/ROUND2_PLAN.md appendix): 64 MN-elements contiguous,
// 8 K-rows 128 B apart, sbo = bytes between 8-row K groups, lbo = bytes between 64-element MN chunks
V6_DEVINL uint64_t make_smem_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// VMN (default; V6B200_ATTN_V=t = transposed copy; validated, BERT round 27.6 vs 28.3 ms): V is read in its natural [B,S,Hkv,D] layout -- tiles
// of [128 keys x 64 d] loaded like K tiles and consumed by the PV MMA as an MN-major B operand -- instead of from a
// pre-transposed Vt copy; `tmap_vt` is then a map over V ([B*S, Hkv*D], box 128 x 64).
template <int D, bool CAUSAL, bool VMN>
__global__ void __launch_bounds__(kThreads, 2)
flash_fwd2_kernel(const __grid_constant__ CUtensorMap tmap_q,     // [B*S, Hq*D]   box 128 x 64
                  const __grid_constant__ CUtensorMap tmap_k,     // [B*S, Hkv*D]  box 128 x 64
                  const __grid_constant__ CUtensorMap tmap_vt,    // [B*Hkv*D, S]  box D x 64
                  const Params P) {
    constexpr int NH = D / 64;
    constexpr int Q_BYTES = BM * D * 2;
    constexpr int K_BYTES = BN * D * 2;
    constexpr int V_BYTES = D * BN * 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-B aligned; derived by pointer arithmetic so that the compiler keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + Q_BYTES;
    uint8_t* sV = sK + K_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + V_BYTES);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = bars + 2;
    uint64_t* v_full = bars + 3;
    uint64_t* v_empty = bars + 4;
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;
    uint64_t* pv_done = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_qt = (P.S + BM - 1) / BM;
    const int qt = n_qt - 1 - blockIdx.x;            // heaviest (causal) tiles first
    const int h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (P.Hq / P.Hkv);
    const int m0 = qt * BM;
    const int n_kv_all = (P.S + BN - 1) / BN;
    const int nkv = CAUSAL ? min(n_kv_all, qt + 1) : n_kv_all;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_vt);
    }
    if (warp == 5 && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(k_full, 1); mbar_init(k_empty, 1);
        mbar_init(v_full, 1); mbar_init(v_empty, 1);
        mbar_init(s_full, 1); mbar_init(p_full, 4);
        mbar_init(pv_done, 1);
        mbar_fence_init();
    }
    if (warp == 6) tmem_alloc<kTmemCols>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
            for (int hh = 0; hh < NH; ++hh)
                tma_load_2d(sQ + hh * (BM * 128), &tmap_q, q_full, h * D + hh * 64, b * P.S + m0);
            for (int j = 0; j < nkv; ++j) {
                const uint32_t ph = j & 1;
                mbar_wait(k_empty, ph ^ 1);                       // S_{j-1} = Q K_{j-1}^T has retired
                mbar_expect_tx(k_full, K_BYTES);
#pragma unroll
                for (int hh = 0; hh < NH; ++hh)
                    tma_load_2d(sK + hh * (BN * 128), &tmap_k, k_full, hk * D + hh * 64, b * P.S + j * BN);
                mbar_wait(v_empty, ph ^ 1);                       // O += P_{j-1} V_{j-1} has retired
                mbar_expect_tx(v_full, V_BYTES);
                if (VMN) {
#pragma unroll
                    for (int hh = 0; hh < NH; ++hh)
                        tma_load_2d(sV + hh * (BN * 128), &tmap_vt, v_full, hk * D + hh * 64, b * P.S + j * BN);
                } else {
#pragma unroll
                    for (int kh = 0; kh < 2; ++kh)
                        tma_load_2d(sV + kh * (D * 128), &tmap_vt, v_full, j * BN + kh * 64, (b * P.Hkv + hk) * D);
                }
            }
        }
    } else if (warp == 5) {
        // ================================ MMA issuer ==================================
        constexpr uint32_t idesc_s = make_idesc_bf16(BM, BN);
        constexpr uint32_t idesc_o = make_idesc_bf16(BM, D) | (VMN ? (1u << 16) : 0u);      // bit 16: B MN-major
        mbar_wait(q_full, 0);
        for (int j = 0; j < nkv; ++j) {
            const uint32_t ph = j & 1;
            // S_j overwrites the columns P_{j-1} lives in: wait until PV_{j-1} has retired (the other
            // resident CTA keeps the tensor pipe busy meanwhile).
            if (j > 0) mbar_wait(pv_done, (j - 1) & 1);
            mbar_wait(k_full, ph);
            tcgen05_fence_after();
            if (lane == 0) {
                const uint32_t qa = smem_u32(sQ), kb = smem_u32(sK);
#pragma unroll
                for (int hh = 0; hh < NH; ++hh)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_base + S_COL, make_smem_desc_sw128(qa + hh * (BM * 128) + k * 32),
                                     make_smem_desc_sw128(kb + hh * (BN * 128) + k * 32), idesc_s, (hh | k) ? 1u : 0u);
                umma_commit(s_full);
                umma_commit(k_empty);
            }
            __syncwarp();
            mbar_wait(p_full, ph);                                // P_j is in TMEM, O rescaled if needed
            mbar_wait(v_full, ph);
            tcgen05_fence_after();
            if (lane == 0) {
                const uint32_t vb = smem_u32(sV);
#pragma unroll
                for (int kk = 0; kk < BN / 16; ++kk)                  // 16 keys per step
                    umma_bf16_ts(tmem_base + O_COL, tmem_base + S_COL + kk * 8,
                                 VMN ? make_smem_desc_sw128_mn(vb + kk * 2048, BN * 128, 1024)
                                     : make_smem_desc_sw128(vb + (kk >> 2) * (D * 128) + (kk & 3) * 32),
                                 idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
                umma_commit(pv_done);
                umma_commit(v_empty);
            }
            __syncwarp();
        }
    } else if (warp < 4) {
        // ================================ softmax / epilogue ==========================
        const int ew = warp;                                  // compute warps are 0-3: the role warps sit in the highest ids (issue priority)
        const int row = m0 + ew * 32 + lane;                 // query position of this thread
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        const uint32_t s_tmem = tmem_base + lane_addr + S_COL;
        float m_used = 0.f, l = 0.f;
        for (int j = 0; j < nkv; ++j) {
            mbar_wait(s_full, j & 1);
            tcgen05_fence_after();
            const int key0 = j * BN;
            const bool need_mask = (key0 + BN > P.S) || (CAUSAL && key0 + BN - 1 > m0);
            // ---- pass 1: exact row max of this tile
            float mx = -INFINITY;
#pragma unroll 1
            for (int c2 = 0; c2 < BN; c2 += 64) {
                uint32_t v0[32], v1[32];
                tmem_ld_32x32b_x32(s_tmem + c2, v0);
                tmem_ld_32x32b_x32(s_tmem + c2 + 32, v1);
                tmem_ld_wait();
                if (need_mask) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int ka = key0 + c2 + i, kb2 = ka + 32;
                        if (ka < P.S && (!CAUSAL || ka <= row)) mx = fmaxf(mx, __uint_as_float(v0[i]));
                        if (kb2 < P.S && (!CAUSAL || kb2 <= row)) mx = fmaxf(mx, __uint_as_float(v1[i]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
                }
            }
            const float m_new = mx * P.scale_log2;
            if (j == 0) m_used = (m_new == -INFINITY) ? 0.f : m_new;
            const bool raise = j > 0 && m_new > m_used + 8.f;
            if (__any_sync(0xffffffffu, raise)) {
                // O must be rescaled: wait until PV_{j-1} has retired, then scale this thread's row
                mbar_wait(pv_done, (j - 1) & 1);
                tcgen05_fence_after();
                const float alpha = raise ? exp2f(m_used - m_new) : 1.f;
                if (raise) { m_used = m_new; l *= alpha; }
#pragma unroll 1
                for (int c = 0; c < D; c += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_base + lane_addr + O_COL + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                    tmem_st_32x32b_x32(tmem_base + lane_addr + O_COL + c, v);
                }
                tmem_st_wait();
            }
            // ---- pass 2: P = exp2(S*c - m), row sum, packed bf16 written over the consumed S columns
            float lsum = 0.f;
#pragma unroll 1
            for (int c2 = 0; c2 < BN; c2 += 64) {
                uint32_t v0[32], v1[32], pk[32];
                tmem_ld_32x32b_x32(s_tmem + c2, v0);
                tmem_ld_32x32b_x32(s_tmem + c2 + 32, v1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float e[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int ii = i + (t & 1);
                        const float sv = __uint_as_float(t < 2 ? v0[ii] : v1[ii]);
                        bool ok = true;
                        if (need_mask) {
                            const int key = key0 + c2 + (t < 2 ? 0 : 32) + ii;
                            ok = key < P.S && (!CAUSAL || key <= row);
                        }
                        e[t] = ok ? exp2f(fmaf(sv, P.scale_log2, -m_used)) : 0.f;
                    }
                    lsum += (e[0] + e[1]) + (e[2] + e[3]);
                    pk[i >> 1] = pack_bf16x2(e[0], e[1]);               // columns c2 + i, c2 + i + 1
                    pk[16 + (i >> 1)] = pack_bf16x2(e[2], e[3]);        // columns c2 + 32 + i, ...
                }
                tmem_st_32x32b_x32(s_tmem + (c2 >> 1), pk);             // 64 bf16 = 32 TMEM columns, inside consumed S
            }
            l += lsum;
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // ---- epilogue: O / l -> bf16 -> global, LSE
        mbar_wait(pv_done, (nkv - 1) & 1);
        tcgen05_fence_after();
        const float inv_l = l > 0.f ? 1.f / l : 0.f;
        const bool row_ok = row < P.S;                          // tcgen05.ld is .sync.aligned: every lane executes it
        __nv_bfloat16* dst = P.O + (((size_t)b * P.S + (row_ok ? row : 0)) * P.Hq + h) * D;
#pragma unroll 1
        for (int c = 0; c < D; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + O_COL + c, v);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 8)
                    *reinterpret_cast<uint4*>(dst + c + i) = make_uint4(
                        pack_bf16x2(__uint_as_float(v[i]) * inv_l, __uint_as_float(v[i + 1]) * inv_l),
                        pack_bf16x2(__uint_as_float(v[i + 2]) * inv_l, __uint_as_float(v[i + 3]) * inv_l),
                        pack_bf16x2(__uint_as_float(v[i + 4]) * inv_l, __uint_as_float(v[i + 5]) * inv_l),
                        pack_bf16x2(__uint_as_float(v[i + 6]) * inv_l, __uint_as_float(v[i + 7]) * inv_l));
            }
        }
        if (row_ok) P.lse[((size_t)b * P.Hq + h) * P.S + row] = m_used * 0.6931471805599453f + logf(l);
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 6) { tcgen05_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}

template <int D>
constexpr int smem_bytes() { return BM * D * 2 + BN * D * 2 + D * BN * 2 + 1024 + 256; }

}  // namespace attn2

template <int D, bool CAUSAL, bool VMN = false>
static int launch_attn2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const attn2::Params& P,
                        cudaStream_t stream) {
    auto kern = attn2::flash_fwd2_kernel<D, CAUSAL, VMN>;
    constexpr int smem = attn2::smem_bytes<D>();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((P.S + attn2::BM - 1) / attn2::BM, P.Hq, P.B);
    kern<<<grid, attn2::kThreads, smem, stream>>>(tq, tk, tv, P);
    V6_CHECK_LAUNCH();
    return 0;
}

// q:[B,S,Hq,D] k:[B,S,Hkv,D] vt:[B,Hkv,D,S] (all bf16, contiguous) -> o:[B,S,Hq,D], lse:[B,Hq,S]
extern "C" int v6_flash_attn_fwd2(const void* q, const void* k, const void* vt, void* o, float* lse, int B, int S, int Hq,
                                  int Hkv, int D, long long ldq, long long ldk, float softmax_scale, int causal, cudaStream_t stream) {
    if ((D != 64 && D != 128) || Hq % Hkv != 0 || S % 8 != 0) return (int)cudaErrorInvalidValue;
    // ldq / ldk: token-row strides in elements (0 = dense).  A packed QKV projection output [B,S,3,H,D] is consumed in
    // place: q = base, k = base + H*D, ldq = ldk = 3*H*D -- no .contiguous() copies.
    if (ldq <= 0) ldq = (long long)Hq * D;
    if (ldk <= 0) ldk = (long long)Hkv * D;
    if (ldq % 8 != 0 || ldk % 8 != 0 || ldq < (long long)Hq * D || ldk < (long long)Hkv * D) return (int)cudaErrorInvalidValue;
    alignas(64) CUtensorMap tq, tk, tv;
    if (v6_make_tmap_2d_bf16(&tq, (uint64_t)q, (uint64_t)B * S, (uint64_t)Hq * D, (uint64_t)ldq * 2, 128, 64, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tk, (uint64_t)k, (uint64_t)B * S, (uint64_t)Hkv * D, (uint64_t)ldk * 2, 128, 64, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tv, (uint64_t)vt, (uint64_t)B * Hkv * D, (uint64_t)S, (uint64_t)S * 2, (uint32_t)D, 64, 1)) return -2;
    attn2::Params P;
    P.O = (__nv_bfloat16*)o; P.lse = lse; P.B = B; P.S = S; P.Hq = Hq; P.Hkv = Hkv;
    P.scale_log2 = softmax_scale * 1.4426950408889634f;
    if (D == 64) return causal ? launch_attn2<64, true>(tq, tk, tv, P, stream) : launch_attn2<64, false>(tq, tk, tv, P, stream);
    return causal ? launch_attn2<128, true>(tq, tk, tv, P, stream) : launch_attn2<128, false>(tq, tk, tv, P, stream);
}

// Opt-in variant: v:[B,S,Hkv,D] in its natural layout (token stride ldv elements, 0 = dense); no Vt copy.
extern "C" int v6_flash_attn_fwd2_vmn(const void* q, const void* k, const void* v, void* o, float* lse, int B, int S, int Hq,
                                      int Hkv, int D, long long ldq, long long ldk, long long ldv, float softmax_scale, int causal,
                                      cudaStream_t stream) {
    if ((D != 64 && D != 128) || Hq % Hkv != 0 || S % 8 != 0) return (int)cudaErrorInvalidValue;
    if (ldq <= 0) ldq = (long long)Hq * D;
    if (ldk <= 0) ldk = (long long)Hkv * D;
    if (ldv <= 0) ldv = (long long)Hkv * D;
    if (ldq % 8 || ldk % 8 || ldv % 8 || ldq < (long long)Hq * D || ldk < (long long)Hkv * D || ldv < (long long)Hkv * D)
        return (int)cudaErrorInvalidValue;
    alignas(64) CUtensorMap tq, tk, tv;
    if (v6_make_tmap_2d_bf16(&tq, (uint64_t)q, (uint64_t)B * S, (uint64_t)Hq * D, (uint64_t)ldq * 2, 128, 64, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tk, (uint64_t)k, (uint64_t)B * S, (uint64_t)Hkv * D, (uint64_t)ldk * 2, 128, 64, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tv, (uint64_t)v, (uint64_t)B * S, (uint64_t)Hkv * D, (uint64_t)ldv * 2, 128, 64, 1)) return -2;
    attn2::Params P;
    P.O = (__nv_bfloat16*)o; P.lse = lse; P.B = B; P.S = S; P.Hq = Hq; P.Hkv = Hkv;
    P.scale_log2 = softmax_scale * 1.4426950408889634f;
    if (D == 64) return causal ? launch_attn2<64, true, true>(tq, tk, tv, P, stream) : launch_attn2<64, false, true>(tq, tk, tv, P, stream);
    return causal ? launch_attn2<128, true, true>(tq, tk, tv, P, stream) : launch_attn2<128, false, true>(tq, tk, tv, P, stream);
}
