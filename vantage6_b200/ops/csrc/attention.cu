// K4 of SURVEY.md 2.6: flash-attention forward on tcgen05 / TMEM / TMA (sm_100a).
//
//   O = softmax(Q K^T * scale [+ causal mask]) V        bf16 in/out, fp32 softmax + accumulate
//   Q:[B,S,Hq,D]  K:[B,S,Hkv,D]  Vt:[B,Hkv,D,S] (V pre-transposed so the PV B-operand is K-major)
//   O:[B,S,Hq,D]  LSE:[B,Hq,S] (natural log; consumed by the backward pass)
//   D in {64,128}; GQA (Hq multiple of Hkv); BERT: 12x64 non-causal, Llama: 32/8 x128 causal.
//
// One CTA per (128-query tile, head, batch), 256 threads:
//   warp 4   TMA producer: Q once, then a 2-stage ring of K_j and Vt_j tiles (128 keys each)
//   warp 5   MMA issuer (one lane):  S_j = Q K_j^T  (SS, 128x128xD)  into TMEM S[j%2]
//                                    O  += P_j V_j  (TS: A = P_j from TMEM, 128xDx128) into TMEM O
//            S_{j+1} is issued before softmax_j finishes, so the tensor pipe overlaps the softmax.
//   warp 6   TMEM allocator (512 columns: S0 | S1 | O | P0 | P1)
//   warps 0-3  softmax: thread == query row (tcgen05.ld 32x32b gives each thread its own row ->
//            row max / row sum need no shuffles); online softmax with LAZY rescaling: the running
//            max is only raised when the tile max exceeds it by > 8 (log2 units); only then is O
//            (in TMEM) rescaled, after waiting for the previous PV MMA.  P_j is written as packed
//            bf16 into its own TMEM columns (tcgen05.st) and consumed by the PV MMA from TMEM; one
//            sweep over S produces both P and the tile max (optimistic single-pass softmax).
#include <cuda.h>
#include "common.cuh"
#include "api.h"

namespace attn {

constexpr int BM = 128;          // query rows per CTA
constexpr int BN = 128;          // keys per KV tile
constexpr int kThreads = 256;
constexpr int kTmemCols = 512;
constexpr int S_COL0 = 0, S_COL1 = 128, O_COL = 256, P_COL0 = 384, P_COL1 = 448;   // P: 64 packed-bf16 columns each

struct Params {
    __nv_bfloat16* O;     // [B,S,Hq,D]
    float* lse;           // [B,Hq,S]
    int B, S, Hq, Hkv;
    float scale_log2;     // softmax_scale * log2(e)
};

template <int D, bool CAUSAL>
__global__ void __launch_bounds__(kThreads, 1)
flash_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q,     // [B*S, Hq*D]   box 128 x 64
                 const __grid_constant__ CUtensorMap tmap_k,     // [B*S, Hkv*D]  box 128 x 64
                 const __grid_constant__ CUtensorMap tmap_vt,    // [B*Hkv*D, S]  box D x 64
                 const Params P) {
    constexpr int NH = D / 64;                       // 64-column halves of the head dimension
    constexpr int Q_BYTES = BM * D * 2;
    constexpr int K_BYTES = BN * D * 2;
    constexpr int V_BYTES = D * BN * 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-B aligned; derived by pointer arithmetic so that the compiler keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + Q_BYTES;                      // 2 stages
    uint8_t* sV = sK + 2 * K_BYTES;                  // 2 stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * V_BYTES);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;      // [2]
    uint64_t* k_empty = bars + 3;     // [2]
    uint64_t* v_full = bars + 5;      // [2]
    uint64_t* v_empty = bars + 7;     // [2]
    uint64_t* s_full = bars + 9;      // [2]
    uint64_t* p_full = bars + 11;     // [2]
    uint64_t* pv_done = bars + 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

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
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
            mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 4);
        }
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
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_expect_tx(&k_full[st], K_BYTES);
#pragma unroll
                for (int hh = 0; hh < NH; ++hh)
                    tma_load_2d(sK + st * K_BYTES + hh * (BN * 128), &tmap_k, &k_full[st], hk * D + hh * 64, b * P.S + j * BN);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_expect_tx(&v_full[st], V_BYTES);
#pragma unroll
                for (int kh = 0; kh < 2; ++kh)
                    tma_load_2d(sV + st * V_BYTES + kh * (D * 128), &tmap_vt, &v_full[st], j * BN + kh * 64, (b * P.Hkv + hk) * D);
            }
        }
    } else if (warp == 5) {
        // ================================ MMA issuer ==================================
        constexpr uint32_t idesc_s = make_idesc_bf16(BM, BN);
        constexpr uint32_t idesc_o = make_idesc_bf16(BM, D);
        auto issue_s = [&](int j) {
            const int st = j & 1;
            mbar_wait(&k_full[st], (j >> 1) & 1);
            tcgen05_fence_after();
            if (lane == 0) {
                const uint32_t d_tmem = tmem_base + (st ? S_COL1 : S_COL0);
                const uint32_t qa = smem_u32(sQ), kb = smem_u32(sK + st * K_BYTES);
#pragma unroll
                for (int hh = 0; hh < NH; ++hh)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(d_tmem, make_smem_desc_sw128(qa + hh * (BM * 128) + k * 32),
                                     make_smem_desc_sw128(kb + hh * (BN * 128) + k * 32), idesc_s, (hh | k) ? 1u : 0u);
                umma_commit(&s_full[st]);
                umma_commit(&k_empty[st]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        issue_s(0);
        if (nkv > 1) issue_s(1);
        for (int j = 0; j < nkv; ++j) {
            const int st = j & 1;
            mbar_wait(&p_full[st], (j >> 1) & 1);            // P_j is in TMEM, O has been rescaled if needed
            mbar_wait(&v_full[st], (j >> 1) & 1);
            tcgen05_fence_after();
            if (lane == 0) {
                const uint32_t p_tmem = tmem_base + (st ? P_COL1 : P_COL0);
                const uint32_t vb = smem_u32(sV + st * V_BYTES);
#pragma unroll
                for (int kk = 0; kk < BN / 16; ++kk)
                    umma_bf16_ts(tmem_base + O_COL, p_tmem + kk * 8,
                                 make_smem_desc_sw128(vb + (kk >> 2) * (D * 128) + (kk & 3) * 32), idesc_o,
                                 (j > 0 || kk > 0) ? 1u : 0u);
                umma_commit(pv_done);
                umma_commit(&v_empty[st]);
            }
            __syncwarp();
            if (j + 2 < nkv) issue_s(j + 2);                 // executes after PV_j in the tensor pipe (in order)
        }
    } else if (warp < 4) {
        // ================================ softmax / epilogue ==========================
        const int ew = warp;                                  // compute warps are 0-3: the role warps sit in the highest ids (issue priority)
        const int row = m0 + ew * 32 + lane;                 // query position of this thread
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        float m_used = -INFINITY, l = 0.f;
        for (int j = 0; j < nkv; ++j) {
            const int st = j & 1;
            const uint32_t s_tmem = tmem_base + lane_addr + (st ? S_COL1 : S_COL0);
            mbar_wait(&s_full[st], (j >> 1) & 1);
            tcgen05_fence_after();
            const int key0 = j * BN;
            const bool need_mask = (key0 + BN > P.S) || (CAUSAL && key0 + BN - 1 > m0);
            const uint32_t p_tmem = tmem_base + lane_addr + (st ? P_COL1 : P_COL0);
            // One sweep over S computes P = exp2(S*c - m) AND the tile max.  m is the running
            // ("lazy") max: it is only raised when the tile max exceeds it by more than 8 (log2 units);
            // then -- rarely, after the first tiles -- O is rescaled and the sweep is repeated with the
            // new m (P lives in its own TMEM columns, so S is still intact for the second sweep).
            // The very first tile has no estimate yet: it gets a max-only sweep first.
            auto sweep = [&](bool want_p, float m_ref, float& mx_out, float& l_out) {
                float mx = -INFINITY, lsum = 0.f;
#pragma unroll 1
                for (int c2 = 0; c2 < BN; c2 += 64) {
                    uint32_t pk[32];
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int c = c2 + half * 32;
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(s_tmem + c, v);
                        tmem_ld_wait();
                        float p[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float sv = __uint_as_float(v[i]);
                            bool ok = true;
                            if (need_mask) {
                                const int key = key0 + c + i;
                                ok = key < P.S && (!CAUSAL || key <= row);
                            }
                            mx = fmaxf(mx, ok ? sv : -INFINITY);
                            if (want_p) {
                                const float e = ok ? exp2f(fmaf(sv, P.scale_log2, -m_ref)) : 0.f;
                                p[i] = e;
                                lsum += e;
                            }
                        }
                        if (want_p) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) pk[half * 16 + i] = pack_bf16x2(p[2 * i], p[2 * i + 1]);
                        }
                    }
                    if (want_p) tmem_st_32x32b_x32(p_tmem + (c2 >> 1), pk);
                }
                mx_out = mx; l_out = lsum;
            };
            float mx, l_tile;
            if (j == 0) {
                sweep(false, 0.f, mx, l_tile);
                const float m0v = mx * P.scale_log2;
                m_used = (m0v == -INFINITY) ? 0.f : m0v;
            }
            sweep(true, m_used, mx, l_tile);
            const float m_new = mx * P.scale_log2;
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
                sweep(true, m_used, mx, l_tile);              // redo P with the raised max
            }
            l += l_tile;
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[st]);
        }
        // ---- epilogue: O / l -> bf16 -> global, LSE
        mbar_wait(pv_done, (nkv - 1) & 1);
        tcgen05_fence_after();
        const float inv_l = l > 0.f ? 1.f / l : 0.f;
        // tcgen05.ld is .sync.aligned: every lane executes it; only the global stores are predicated
        const bool row_ok = row < P.S;
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
constexpr int smem_bytes() { return BM * D * 2 + 2 * BN * D * 2 + 2 * D * BN * 2 + 1024 + 256; }

}  // namespace attn

template <int D, bool CAUSAL>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const attn::Params& P,
                       cudaStream_t stream) {
    auto kern = attn::flash_fwd_kernel<D, CAUSAL>;
    constexpr int smem = attn::smem_bytes<D>();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((P.S + attn::BM - 1) / attn::BM, P.Hq, P.B);
    kern<<<grid, attn::kThreads, smem, stream>>>(tq, tk, tv, P);
    V6_CHECK_LAUNCH();
    return 0;
}

// q:[B,S,Hq,D] k:[B,S,Hkv,D] vt:[B,Hkv,D,S] (all bf16, contiguous) -> o:[B,S,Hq,D], lse:[B,Hq,S]
extern "C" int v6_flash_attn_fwd(const void* q, const void* k, const void* vt, void* o, float* lse, int B, int S, int Hq,
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
    attn::Params P;
    P.O = (__nv_bfloat16*)o; P.lse = lse; P.B = B; P.S = S; P.Hq = Hq; P.Hkv = Hkv;
    P.scale_log2 = softmax_scale * 1.4426950408889634f;
    if (D == 64) return causal ? launch_attn<64, true>(tq, tk, tv, P, stream) : launch_attn<64, false>(tq, tk, tv, P, stream);
    return causal ? launch_attn<128, true>(tq, tk, tv, P, stream) : launch_attn<128, false>(tq, tk, tv, P, stream);
}
