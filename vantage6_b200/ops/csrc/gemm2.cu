// 2-CTA (cta_group::2) tcgen05 GEMM:  C[M,N] = act(A[M,K] . B[N,K]^T + bias)   bf16 -> fp32 (TMEM) -> bf16
//
// Two SMs of one TPC (a 2-CTA cluster) cooperate on a 256x256 output tile:
//   * each CTA TMA-loads ITS half of the operands (128 rows of A, 128 rows of B: 32 KB per stage
//     instead of 48 KB for the 1-CTA 128x256 tile -> 1.5x less L2->SM traffic per FLOP, which is what
//     bounds the 1-CTA kernel at ~75% tensor-pipe utilisation, see profiles/ncu_gemm_r1a.md),
//   * the LEADER CTA's MMA thread issues tcgen05.mma.cta_group::2 (UMMA 256x256x16): the tensor cores
//     of both SMs read A from their own shared memory and B from both, accumulating 128 rows each in
//     their own TMEM,
//   * completion is multicast: tcgen05.commit...multicast::cluster arrives on the same-offset
//     mbarrier in both CTAs (frees the smem stage / publishes the accumulator),
//   * both CTAs' TMA loads complete_tx on the LEADER's full barrier (cp.async.bulk.tensor .cta_group::2),
//   * each CTA's 4 epilogue warps drain their own TMEM half and arrive (remotely for the peer) on the
//     leader's tmem-empty barrier.
// Persistent: 74 clusters walk the (M/256 x N/256) tile grid; 6-stage smem ring; 2 accumulator stages.
#include <cuda.h>
#include "common.cuh"
#include "api.h"

namespace gemm2 {

constexpr int TILE_M = 256, TILE_N = 256, BLOCK_K = 64, UMMA_K = 16;
constexpr int HALF_M = 128, HALF_N = 128;
constexpr int kStages = 6;
constexpr int kAccStages = 2;
constexpr int kTmemCols = 512;
constexpr int A_BYTES = HALF_M * BLOCK_K * 2;        // 16 KB
constexpr int B_BYTES = HALF_N * BLOCK_K * 2;        // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;       // 32 KB per CTA
constexpr int STG_OFF = kStages * STAGE_BYTES;       // 192 KB: epilogue staging, 2 x (32 rows x 64 cols bf16, SWIZZLE_128B) per warp
constexpr int STG_BOX_BYTES = 32 * 128;
constexpr int BAR_OFF = STG_OFF + 4 * 2 * STG_BOX_BYTES;     // 224 KB
constexpr int SMEM_BYTES = BAR_OFF + 1024 + 256;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
constexpr int kThreads = 256;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;       // clears the CTA-rank bit of a shared::cluster address

struct Params {
    int M, N, K;
    __nv_bfloat16* C;
    int ldc;
    const float* bias;
    int act;
};

V6_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

V6_DEVINL uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
V6_DEVINL void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
V6_DEVINL uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
V6_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
template <int kCols>
V6_DEVINL void tmem_alloc_2sm(uint32_t* smem_result) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(smem_result)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
V6_DEVINL void tmem_dealloc_2sm(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(kCols) : "memory");
}
// TMA load whose completion is signalled on the LEADER CTA's mbarrier
V6_DEVINL void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint64_t* bar_local, int c0, int c1) {
    const uint32_t bar = smem_u32(bar_local) & kPeerBitMask;
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(smem_dst)), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
V6_DEVINL void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
V6_DEVINL void umma_commit_2sm(uint64_t* bar) {       // arrive on the same-offset barrier of BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"((uint16_t)0x3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a,     // A [M,K] box 128 x 64
                  const __grid_constant__ CUtensorMap tmap_b,     // B [N,K] box 128 x 64
                  const __grid_constant__ CUtensorMap tmap_c,     // C [M,N] box 32 x 64 (epilogue TMA stores)
                  const Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-B aligned; derived by pointer arithmetic so that the compiler keeps the shared address space (LDS/STS, not generic LD/ST)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + kAccStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + kAccStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int num_m = (P.M + TILE_M - 1) / TILE_M;
    const int num_n = (P.N + TILE_N - 1) / TILE_N;
    const int num_k = (P.K + BLOCK_K - 1) / BLOCK_K;
    const int num_tiles = num_m * num_n;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (warp == 4 && lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); }
    if (warp == 5 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 8); }
        mbar_fence_init();
    }
    cluster_sync_all();                                  // barriers of both CTAs are initialised
    if (warp == 6) tmem_alloc_2sm<kTmemCols>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync_all();

    // Tile order.  Few row tiles: m fastest (concurrent CTAs share one weight block).  Many row tiles (A does not stay in
    // L2 between passes): grouped rasterisation -- a wave covers GROUP_N column blocks x ~148/GROUP_N row blocks, so every
    // A tile is reused GROUP_N times and every B tile ~18 times out of L2 instead of A being re-streamed from HBM once
    // per column block (8192^3: 32 passes over 134 MB).
    constexpr int GROUP_N = 8;
    const bool grouped = num_m >= 32 && num_n >= 2 * GROUP_N;
    auto tile_coords = [&](int t, int& m_blk, int& n_blk) {
        if (!grouped) { m_blk = t % num_m; n_blk = t / num_m; return; }
        const int per_group = GROUP_N * num_m, g = t / per_group, first = g * GROUP_N, in_g = t - g * per_group;
        const int gsz = min(GROUP_N, num_n - first);
        n_blk = first + in_g % gsz;
        m_blk = in_g / gsz;
    };

    if (warp == 4) {
        // ============================ TMA producer (both CTAs) ============================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = pair; t < num_tiles; t += num_pairs) {
                int m_blk, n_blk;
                tile_coords(t, m_blk, n_blk);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);     // bytes of BOTH CTAs
                    tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * TILE_M + (int)rank * HALF_M);
                    tma_load_2d_2sm(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * TILE_N + (int)rank * HALF_N);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 5) {
        // ============================ MMA issuer (leader CTA only) ============================
        if (leader) {
            constexpr uint32_t idesc = make_idesc_bf16(TILE_M, TILE_N);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int t = pair; t < num_tiles; t += num_pairs) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);        // both CTAs' epilogues drained this accumulator
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * TILE_N;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    if (lane == 0) {
                        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                        const uint32_t sb = sa + A_BYTES;
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_bf16_ss_2sm(d_tmem, make_smem_desc_sw128(sa + k * UMMA_K * 2),
                                             make_smem_desc_sw128(sb + k * UMMA_K * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_commit_2sm(&empty_bar[stage]);
                        if (kb == num_k - 1) umma_commit_2sm(&tfull_bar[acc]);
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp < 4) {       // epilogue warps 0-3: the scheduler prefers the highest warp id, so the pacing single-thread roles sit in warps 4-7
        // ============================ epilogue (both CTAs, own 128 rows) ============================
        const int ew = warp;
        int acc = 0; uint32_t acc_phase = 0;
        int stg_slot = 0;
        for (int t = pair; t < num_tiles; t += num_pairs) {
            int m_blk, n_blk;
            tile_coords(t, m_blk, n_blk);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tcgen05_fence_after();
            // 64-column chunks: TMEM -> registers (bias / activation / bf16) -> swizzled 4 KB staging box (two per
            // warp, ping-pong) -> one coalesced TMA store per chunk; TMA clips rows >= M and columns >= N.
            const int row0 = m_blk * TILE_M + (int)rank * HALF_M + ew * 32;
            const uint32_t t_row = tmem_base + acc * TILE_N + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < TILE_N; c += 64) {
                uint32_t v[2][32];
                tmem_ld_32x32b_x32(t_row + c, v[0]);
                tmem_ld_32x32b_x32(t_row + c + 32, v[1]);
                tmem_ld_wait();
                const int col0 = n_blk * TILE_N + c;
                if (row0 < P.M && col0 < P.N) {                       // warp-uniform
                    uint32_t packed[32];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[h][j]);
                        if (P.bias) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (col0 + h * 32 + j < P.N) f[j] += __ldg(P.bias + col0 + h * 32 + j);
                        }
                        if (P.act == 1) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
                        } else if (P.act == 2) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) packed[h * 16 + j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
                    }
                    uint8_t* stg = smem + STG_OFF + (ew * 2 + stg_slot) * STG_BOX_BYTES;
                    if (lane == 0) tma_store_wait_read_1();           // the store issued two chunks ago has read this box
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                            make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) { tma_store_2d(&tmap_c, stg, col0, row0); tma_store_commit(); }
                    stg_slot ^= 1;
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa(smem_u32(&tempty_bar[acc]), 0));     // leader's barrier
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) tma_store_wait_all();                          // staging must outlive the last store
    }

    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // the peer may still be reading / being signalled
    if (warp == 6) { tcgen05_fence_after(); tmem_dealloc_2sm<kTmemCols>(tmem_base); }
}

}  // namespace gemm2

extern "C" int v6_gemm2_bf16(const void* A, const void* B, void* C, const float* bias, int M, int N, int K, int lda, int ldb,
                             int ldc, int act, cudaStream_t stream) {
    using namespace gemm2;
    if (K % 8 != 0 || lda % 8 != 0 || ldb % 8 != 0 || ldc % 8 != 0) return (int)cudaErrorInvalidValue;
    alignas(64) CUtensorMap ta, tb;
    if (v6_make_tmap_2d_bf16(&ta, (uint64_t)A, M, K, (uint64_t)lda * 2, HALF_M, BLOCK_K, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tb, (uint64_t)B, N, K, (uint64_t)ldb * 2, HALF_N, BLOCK_K, 1)) return -2;
    alignas(64) CUtensorMap tc;
    if (v6_make_tmap_2d_bf16(&tc, (uint64_t)C, M, N, (uint64_t)ldc * 2, 32, 64, 1)) return -2;
    Params P;
    P.M = M; P.N = N; P.K = K; P.C = (__nv_bfloat16*)C; P.ldc = ldc; P.bias = bias; P.act = act;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm2_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int num_tiles = ((M + TILE_M - 1) / TILE_M) * ((N + TILE_N - 1) / TILE_N);
    int pairs = num_tiles < 74 ? num_tiles : 74;
    gemm2_bf16_kernel<<<pairs * 2, kThreads, SMEM_BYTES, stream>>>(ta, tb, tc, P);
    V6_CHECK_LAUNCH();
    return 0;
}
