// tcgen05 / TMEM / TMA GEMM for sm_100a, hand-written (no CUTLASS):
//
//     C[M,N] = act( A[M,K] . B[N,K]^T + bias[N] )        bf16 in, fp32 accumulate in TMEM
//
// i.e. y = x @ W^T exactly as nn.Linear stores W.  Two entry points share the kernel:
//   * v6_gemm_bf16        : plain fused linear (bias + optional GELU epilogue)
//   * v6_bcast_gemm_bf16  : K1 of SURVEY.md 2.6 -- "global-model broadcast fused with the first
//     forward GEMM that consumes it": the B operand (weights) is fetched by TMA straight from
//     the SERVER GPU's copy through a peer-mapped VA over NVLink; the m_blk==0 tile of every
//     weight column-block pulls the tile, feeds tcgen05.mma from shared memory AND writes the
//     tile back to the local weight buffer with a TMA store (so the weights are resident for
//     the other row-blocks, the backward pass and later steps); other row-blocks wait on a
//     per-(n_blk,k_blk) ready flag and read the local copy.  NVLink traffic = 1x weights and
//     the transfer overlaps the MMA tile by tile; no NCCL broadcast, no separate copy kernel.
//
// Structure (persistent, warp-specialised, one CTA per SM, 256 threads):
//   warp 0 : TMA producer (one elected lane)        smem ring: kStages x (A 128x64 | B 256x64)
//   warp 1 : MMA issuer  (one elected lane)         UMMA 128x256x16, kind::f16, cta_group::1
//   warp 2 : TMEM allocator (512 columns = 2 accumulator stages of 128x256 fp32)
//   warp 3 : K1 write-back warp (TMA store of pulled weight tiles + ready flags)
//   warps 4-7 : epilogue (tcgen05.ld 32x32b.x32 -> bias/act -> bf16 -> swizzled smem box -> TMA store)
// Pipelines: full/empty mbarriers per smem stage; tmem_full/tmem_empty per accumulator stage,
// so the epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include "common.cuh"
#include "api.h"

namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;           // 64 bf16 = 128 B = one SWIZZLE_128B atom
constexpr int UMMA_K = 16;
constexpr int kStages = 4;            // plain GEMM: 4-stage ring; K1: 3 stages + a 2-slot pull staging area
constexpr int kStagesK1 = 3;
constexpr int kAccStages = 2;
constexpr int kTmemCols = 512;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;      // 16 KB
constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;      // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;      // 48 KB
constexpr int PULL_OFF = kStagesK1 * STAGE_BYTES;   // K1 pull staging: 2 x 32 KB after the 3-stage ring
constexpr int STG_OFF = PULL_OFF + 2 * B_BYTES;     // 208 KB (>= 4 x 48 KB of the plain ring): epilogue staging,
constexpr int STG_WARP_BYTES = 32 * 128;            //   one 32-row x 64-col bf16 SWIZZLE_128B box (4 KB) per epilogue warp
constexpr int BAR_OFF = STG_OFF + 4 * STG_WARP_BYTES;
constexpr int SMEM_BYTES = BAR_OFF + 1024 /*align slack*/ + 256 /*barriers*/;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
constexpr int kThreads = 256;

struct Params {
    int M, N, K;
    __nv_bfloat16* C;          // [M, ldc]
    int ldc;
    const float* bias;         // [N] or null
    int act;                   // 0 none, 1 gelu(erf), 2 relu
    // K1 fusion (all null/0 for the plain GEMM)
    int fused_bcast;
    int stages;                // smem ring depth (kStages or kStagesK1)
    uint32_t* ready_flags;     // [num_n_blk * num_k_blk] local, zero before round 1
    uint32_t epoch;            // flags are compared against this monotonically increasing value
    // K1 v3 ("push", fused_bcast == 2): the OWNER of the weights multicasts every tile once through the switch
    // (multimem.st: egress 1x instead of one pull per consumer) and raises the tile's flag on every rank; all ranks
    // (the owner included) consume tiles from their local copy as the flags arrive.
    int is_owner, world, ldb;
    int push_kgroup;                   // push: k-blocks per push unit (1, 2 or 4)
    int own_nb_lo, own_nb_hi;          // push: this rank owns (multicasts) the 256-row weight blocks [lo, hi)
    const uint32_t* epoch_ptr;         // non-null: the epoch is read from device memory (a captured graph replays with the current round)
    uint32_t* status_ptr;              // non-null: a flag wait that times out sets *status_ptr = 1 and goes on (instead of trapping)
    const __nv_bfloat16* b_src;      // owner's source (its own copy of the weights)
    __nv_bfloat16* b_mc;             // multicast VA over every rank's weight buffer
    PeerTable flag_peers;            // every rank's flag array (peer VAs)
};

V6_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a,        // A  [M,K]  box 128x64
                 const __grid_constant__ CUtensorMap tmap_b,        // B  [N,K]  box 256x64 (local copy)
                 const __grid_constant__ CUtensorMap tmap_b_src,    // B on the server GPU (peer VA); K1 only
                 const __grid_constant__ CUtensorMap tmap_c,        // C  [M,N]  box 32x64 (epilogue TMA stores)
                 const Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-B aligned; derived by pointer arithmetic so that the compiler keeps the shared address space (LDS/STS, not generic LD/ST)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + kAccStages;
    uint64_t* pull_bar = tempty_bar + kAccStages;       // [2] K1 pull staging
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pull_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = (P.M + BLOCK_M - 1) / BLOCK_M;
    const int num_n = (P.N + BLOCK_N - 1) / BLOCK_N;
    const int num_k = (P.K + BLOCK_K - 1) / BLOCK_K;
    const int num_tiles = num_m * num_n;
    const int empty_count = 1;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if (P.fused_bcast) tma_prefetch_desc(&tmap_b_src);
    }
    if (warp == 5 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], empty_count); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
        mbar_init(&pull_bar[0], 1); mbar_init(&pull_bar[1], 1);
        mbar_fence_init();
    }
    if (warp == 6) tmem_alloc<kTmemCols>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile order: m fastest, so CTAs running concurrently share the same weight column block (L2 reuse)
    // Tile order.  Few row tiles: m fastest (concurrent CTAs share one weight block).  Many row tiles (A does not stay in
    // L2 between passes): grouped rasterisation -- a wave covers GROUP_N column blocks x ~148/GROUP_N row blocks, so every
    // A tile is reused GROUP_N times and every B tile ~18 times out of L2 instead of A being re-streamed from HBM once
    // per column block (8192^3: 32 passes over 134 MB).
    constexpr int GROUP_N = 8;
    const bool grouped = num_m >= 32 && num_n >= 2 * GROUP_N && !P.fused_bcast;
    auto tile_coords = [&](int t, int& m_blk, int& n_blk) {
        if (!grouped) { m_blk = t % num_m; n_blk = t / num_m; return; }
        const int per_group = GROUP_N * num_m, g = t / per_group, first = g * GROUP_N, in_g = t - g * per_group;
        const int gsz = min(GROUP_N, num_n - first);
        n_blk = first + in_g % gsz;
        m_blk = in_g / gsz;
    };

    // K1 v3 push (owner side), run by the K1 warp of every CTA AND -- before their first accumulator exists -- by the four
    // epilogue warps: 5 pusher warps per CTA keep ~6 MB of multicast stores in flight (one warp per CTA: 1.2 MB, a third of
    // what the link takes -- profiles/comm_2gpu_r2c.json).
    auto k1_push = [&](int pusher, int n_pushers) {
        const uint32_t epoch = P.epoch_ptr ? ld_acquire_sys_u32(P.epoch_ptr) : P.epoch;
        // push unit = one 256-row weight block x `kg` consecutive k-blocks: a row of the unit is kg x 128 contiguous
        // bytes (8 * kg lanes x 16 B, fully coalesced), 16 row-chunks in flight per lane, ONE fence + kg flags per
        // unit.  (One 32 KB tile per fence with 4 rows per instruction reached 170 GB/s from inside the GEMM --
        // profiles/comm_2gpu_r2b.json -- a third of what the link takes.)
        const int kg = P.push_kgroup, num_kg = (num_k + kg - 1) / kg;
        const int lpr = 8 * kg, rpi = 32 / lpr;                          // lanes per row, rows per instruction
        const int rsub = lane / lpr, ch = lane % lpr;
        const int n_units = num_kg * num_n;
        for (int idx = pusher; idx < n_units; idx += n_pushers) {
            // n-slowest, like the consumers' tile order (m fastest, then n): the column blocks the first wave of output tiles
            // needs are complete after 1/num_n of the transfer, and the GEMM follows the push block by block (k-major order
            // let no tile finish before the LAST k-block of everything had arrived: push and GEMM ran back to back)
            const int nb = idx / num_kg, kgi = idx % num_kg;
            if (nb < P.own_nb_lo || nb >= P.own_nb_hi) continue;          // another rank's shard
            const int rows = min(BLOCK_N, P.N - nb * BLOCK_N);
            const int col = kgi * kg * BLOCK_K + ch * 8;
            const size_t base = (size_t)(nb * BLOCK_N) * P.ldb + (size_t)col;
            const bool col_ok = col < P.K;
            constexpr int PU = 16;
#pragma unroll 1
            for (int r0 = 0; r0 < rows; r0 += PU * rpi) {
                uint4 v[PU];
#pragma unroll
                for (int u = 0; u < PU; ++u) {
                    const int r = r0 + u * rpi + rsub;
                    if (r < rows && col_ok) v[u] = *reinterpret_cast<const uint4*>(P.b_src + base + (size_t)r * P.ldb);
                }
#pragma unroll
                for (int u = 0; u < PU; ++u) {
                    const int r = r0 + u * rpi + rsub;
                    if (r < rows && col_ok) multimem_st_u4(reinterpret_cast<uint4*>(P.b_mc + base + (size_t)r * P.ldb), v[u]);
                }
            }
            fence_acq_rel_sys();                       // every lane's multicast stores before the flags
            __syncwarp();
            for (int j = lane; j < P.world * kg; j += 32) {
                const int peer = j / kg, kb = kgi * kg + j % kg;
                if (kb < num_k) st_release_sys_u32(reinterpret_cast<uint32_t*>(P.flag_peers.p[peer]) + nb * num_k + kb, epoch);
            }
            __syncwarp();
        }
    };

    if (warp == 4) {
        // ============================ TMA producer ============================
        int stage = 0; uint32_t phase = 0;                       // (advanced by lane 0 only)
        auto issue = [&](int kb, int m_blk, int n_blk) {            // lane 0: one k-block of operands into the ring
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * STAGE_BYTES;
            uint8_t* sb = sa + A_BYTES;
            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
            if (++stage == P.stages) { stage = 0; phase ^= 1; }
        };
        if (!P.fused_bcast) {
            if (lane == 0) {
                for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                    int m_blk, n_blk;
                    tile_coords(t, m_blk, n_blk);
                    for (int kb = 0; kb < num_k; ++kb) issue(kb, m_blk, n_blk);
                }
            }
        } else {
            // K1: a weight tile may only be loaded once its ready flag carries this round's epoch.  The WHOLE warp polls: 32
            // flags per L2 round trip (a system-scope load costs ~1 us; one dependent load per k-block in front of every TMA
            // issue made the producer 3x slower than the MMA -- push and GEMM ran back to back, profiles/comm_8gpu_r2.json),
            // lane 0 issues the k-blocks of the ready prefix; column blocks this CTA has seen complete are not polled again.
            const uint32_t epoch = P.epoch_ptr ? ld_acquire_sys_u32(P.epoch_ptr) : P.epoch;
            unsigned long long done_n = 0ull;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                int m_blk, n_blk;
                tile_coords(t, m_blk, n_blk);
                if (n_blk < 64 && ((done_n >> n_blk) & 1ull)) {
                    if (lane == 0) for (int kb = 0; kb < num_k; ++kb) issue(kb, m_blk, n_blk);
                    __syncwarp();
                    continue;
                }
                for (int kb0 = 0; kb0 < num_k; kb0 += 32) {
                    const int nb = min(32, num_k - kb0);
                    const uint32_t* f = P.ready_flags + n_blk * num_k + kb0;
                    int issued = 0;
                    const long long t0 = clock64();
                    while (issued < nb) {
                        const uint32_t v = lane < nb ? ld_acquire_sys_u32(f + lane) : epoch;    // acquire: the tile behind a set flag is visible; the warp vote below carries that to lane 0
                        const unsigned m = __ballot_sync(0xffffffffu, (int32_t)(v - epoch) >= 0);
                        int cnt = m == 0xffffffffu ? 32 : __ffs(~m) - 1;          // length of the ready prefix
                        cnt = min(cnt, nb);
                        if (cnt == issued) {
                            __nanosleep(64);
                            if (clock64() - t0 > 4000000000LL) {
                                if (!P.status_ptr) asm volatile("trap;");
                                if (lane == 0) *P.status_ptr = 1u;                  // dead owner: reported, the round is redone
                                cnt = nb;
                            } else {
                                continue;
                            }
                        }
                        asm volatile("fence.proxy.async.global;" ::: "memory");          // generic-proxy stores -> the TMA (async proxy) reads
                        if (lane == 0) for (int k = issued; k < cnt; ++k) issue(kb0 + k, m_blk, n_blk);
                        issued = cnt;
                        __syncwarp();
                    }
                }
                if (n_blk < 64) done_n |= 1ull << n_blk;
            }
        }
    } else if (warp == 5) {
        // ============================ MMA issuer ==============================
        constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);          // epilogue has drained this accumulator
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            for (int kb = 0; kb < num_k; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t da = make_smem_desc_sw128(sa + k * UMMA_K * 2);
                        const uint64_t db = make_smem_desc_sw128(sb + k * UMMA_K * 2);
                        umma_bf16_ss(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);               // frees the smem stage when the MMAs retire
                    if (kb == num_k - 1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == P.stages) { stage = 0; phase ^= 1; }
            }
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp == 7) {
        // ============================ K1 pull warp ============================
        // Every CTA pulls its share of the weight tiles from the server GPU (TMA load through the
        // peer-mapped VA, over NVLink) into a 2-slot staging area, TMA-stores them into the local
        // weight buffer and publishes a per-tile ready flag.  Pull order is k-major so that the
        // first k-blocks of every column land first; all CTAs pull in parallel (NVLink saturated)
        // and never wait on anything but their own staging slots, so the GEMM tiles can start as
        // soon as their first weight tiles are local: transfer and MMA overlap tile by tile.
        if (P.fused_bcast == 2) {
            if (P.is_owner) k1_push(blockIdx.x * 5 + 4, gridDim.x * 5);
        } else if (P.fused_bcast && lane == 0) {
            const int n_pull = num_k * num_n;
            uint8_t* stg = smem + PULL_OFF;
            int slot = 0; uint32_t ph[2] = {0, 0};
            int issued = 0;
            int first = blockIdx.x;
            // software pipeline: load(i+1) is in flight while tile i is stored
            auto issue_load = [&](int idx, int s) {
                const int kb = idx / num_n, nb = idx % num_n;
                mbar_expect_tx(&pull_bar[s], B_BYTES);
                tma_load_2d(stg + s * B_BYTES, &tmap_b_src, &pull_bar[s], kb * BLOCK_K, nb * BLOCK_N);
            };
            if (first < n_pull) { issue_load(first, 0); issued = 1; }
            for (int idx = first; idx < n_pull; idx += gridDim.x) {
                const int nxt = idx + gridDim.x;
                if (nxt < n_pull) {
                    tma_store_wait_read();                    // staging slot (slot^1) no longer read by a store
                    issue_load(nxt, slot ^ 1);
                    ++issued;
                }
                mbar_wait(&pull_bar[slot], ph[slot]);
                ph[slot] ^= 1;
                const int kb = idx / num_n, nb = idx % num_n;
                tma_store_2d(&tmap_b, stg + slot * B_BYTES, kb * BLOCK_K, nb * BLOCK_N);   // un-swizzles on the way out
                tma_store_commit();
                tma_store_wait_all();                         // bytes are in the local copy
                asm volatile("fence.proxy.async.global;" ::: "memory");
                __threadfence();
                st_release_sys_u32(P.ready_flags + nb * num_k + kb, P.epoch);
                slot ^= 1;
            }
            (void)issued;
        }
    } else if (warp < 4) {       // epilogue warps 0-3: the scheduler prefers the highest warp id, so the pacing single-thread roles sit in warps 4-7
        if (P.fused_bcast == 2 && P.is_owner) k1_push(blockIdx.x * 5 + warp, gridDim.x * 5);     // idle until the first tile: help push
        // ============================ epilogue ================================
        const int ew = warp;                      // == warp % 4 -> TMEM lanes [32*ew, 32*ew+32)
        int acc = 0; uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            int m_blk, n_blk;
            tile_coords(t, m_blk, n_blk);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tcgen05_fence_after();
            // v2 epilogue: 64-column chunks go TMEM -> registers (bias / activation / bf16) -> a swizzled 4 KB
            // staging box in shared memory -> ONE coalesced TMA store per chunk (v1 wrote 16 B per lane at a
            // row stride: partial-sector writes, 77 us for a 200704x256x64 problem cuBLAS does in 26 us --
            // profiles/conv1x1_bench_r1.json).  TMA clips rows >= M and columns >= N.
            const int row0 = m_blk * BLOCK_M + ew * 32;
            const uint32_t t_row = tmem_base + acc * BLOCK_N + ((uint32_t)(ew * 32) << 16);
            uint8_t* stg = smem + STG_OFF + ew * STG_WARP_BYTES;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 64) {
                uint32_t v[2][32];
                tmem_ld_32x32b_x32(t_row + c, v[0]);
                tmem_ld_32x32b_x32(t_row + c + 32, v[1]);
                tmem_ld_wait();
                const int col0 = n_blk * BLOCK_N + c;
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
                    if (lane == 0) tma_store_wait_read();             // previous chunk's store has read the staging box
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)                        // 16 B chunk j of row `lane`, SWIZZLE_128B position
                        *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                            make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) { tma_store_2d(&tmap_c, stg, col0, row0); tma_store_commit(); }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) tma_store_wait_all();                          // staging must outlive the last store
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 6) { tcgen05_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}

}  // namespace gemm


static int launch_gemm(const void* A, const void* B, const void* B_src, void* C, const float* bias, int M, int N, int K,
                       int lda, int ldb, int ldc, int act, uint32_t* ready_flags, uint32_t epoch, int max_ctas,
                       cudaStream_t stream) {
    using namespace gemm;
    if (K % 8 != 0 || lda % 8 != 0 || ldb % 8 != 0 || ldc % 8 != 0) return (int)cudaErrorInvalidValue;
    alignas(64) CUtensorMap ta, tb, tbs;
    if (v6_make_tmap_2d_bf16(&ta, (uint64_t)A, M, K, (uint64_t)lda * 2, BLOCK_M, BLOCK_K, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tb, (uint64_t)B, N, K, (uint64_t)ldb * 2, BLOCK_N, BLOCK_K, 1)) return -2;
    if (B_src) { if (v6_make_tmap_2d_bf16(&tbs, (uint64_t)B_src, N, K, (uint64_t)ldb * 2, BLOCK_N, BLOCK_K, 1)) return -2; }
    else tbs = tb;
    alignas(64) CUtensorMap tc;
    if (v6_make_tmap_2d_bf16(&tc, (uint64_t)C, M, N, (uint64_t)ldc * 2, 32, 64, 1)) return -2;
    Params P = {};
    P.M = M; P.N = N; P.K = K; P.C = (__nv_bfloat16*)C; P.ldc = ldc; P.bias = bias; P.act = act;
    P.fused_bcast = B_src ? 1 : 0; P.ready_flags = ready_flags; P.epoch = epoch;
    P.stages = B_src ? kStagesK1 : kStages;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int num_tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BLOCK_N - 1) / BLOCK_N);
    int sms = max_ctas > 0 ? max_ctas : 148;
    const int grid = num_tiles < sms ? num_tiles : sms;
    gemm_bf16_kernel<<<grid, kThreads, SMEM_BYTES, stream>>>(ta, tb, tbs, tc, P);
    V6_CHECK_LAUNCH();
    return 0;
}

extern "C" int v6_gemm_bf16(const void* A, const void* B, void* C, const float* bias, int M, int N, int K, int lda,
                            int ldb, int ldc, int act, cudaStream_t stream) {
    return launch_gemm(A, B, nullptr, C, bias, M, N, K, lda, ldb, ldc, act, nullptr, 0, 0, stream);
}
extern "C" int v6_bcast_gemm_bf16(const void* A, void* B_local, const void* B_server_peer, void* C, const float* bias,
                                  int M, int N, int K, int lda, int ldb, int ldc, int act, uint32_t* ready_flags,
                                  uint32_t epoch, cudaStream_t stream) {
    return launch_gemm(A, B_local, B_server_peer, C, bias, M, N, K, lda, ldb, ldc, act, ready_flags, epoch, 0, stream);
}
// K1 v3: C = A . W^T where W (the global model's weight) is multicast by its owner from inside this kernel.
//   B_local: this rank's copy (a buffer of the symmetric heap, bound to `B_mc`); flags: this rank's flag array;
//   flag_peers: every rank's flag array.  The owner passes is_owner = 1 (its B_local holds the new weights).
extern "C" int v6_bcast_push_gemm_bf16(const void* A, void* B_local, void* B_mc, void* C, const float* bias, int M, int N, int K,
                                       int lda, int ldb, int ldc, int act, uint32_t* ready_flags, const PeerTable* flag_peers,
                                       int world, int is_owner, uint32_t epoch, int own_nb_lo, int own_nb_hi, const uint32_t* epoch_ptr,
                                       uint32_t* status_ptr, cudaStream_t stream) {
    using namespace gemm;
    if (K % 8 != 0 || lda % 8 != 0 || ldb % 8 != 0 || ldc % 8 != 0 || !B_mc) return (int)cudaErrorInvalidValue;
    alignas(64) CUtensorMap ta, tb, tc;
    if (v6_make_tmap_2d_bf16(&ta, (uint64_t)A, M, K, (uint64_t)lda * 2, BLOCK_M, BLOCK_K, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tb, (uint64_t)B_local, N, K, (uint64_t)ldb * 2, BLOCK_N, BLOCK_K, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tc, (uint64_t)C, M, N, (uint64_t)ldc * 2, 32, 64, 1)) return -2;
    Params P = {};
    P.M = M; P.N = N; P.K = K; P.C = (__nv_bfloat16*)C; P.ldc = ldc; P.bias = bias; P.act = act;
    P.fused_bcast = 2; P.ready_flags = ready_flags; P.epoch = epoch; P.stages = kStages;
    P.is_owner = is_owner; P.own_nb_lo = own_nb_lo; P.own_nb_hi = own_nb_hi < 0 ? (N + BLOCK_N - 1) / BLOCK_N : own_nb_hi;
    P.epoch_ptr = epoch_ptr; P.status_ptr = status_ptr;
    {   // widest push unit that still gives every CTA's pusher warp two units of the blocks this rank owns
        const int nn = P.own_nb_hi - P.own_nb_lo, nk = (K + BLOCK_K - 1) / BLOCK_K;
        P.push_kgroup = 1;
        for (int kg : {4, 2}) if ((long long)nn * ((nk + kg - 1) / kg) >= 148 * 5 / (kg == 2 ? 2 : 1)) { P.push_kgroup = kg; break; }
    }
    P.world = world; P.ldb = ldb; P.b_src = (const __nv_bfloat16*)B_local; P.b_mc = (__nv_bfloat16*)B_mc;
    P.flag_peers = *flag_peers;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int num_tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BLOCK_N - 1) / BLOCK_N);
    // the owner needs enough CTAs to push every weight tile promptly even when the GEMM itself has few output tiles
    int grid = num_tiles < 148 ? (is_owner ? 148 : num_tiles) : 148;
    gemm_bf16_kernel<<<grid, kThreads, SMEM_BYTES, stream>>>(ta, tb, tb, tc, P);
    V6_CHECK_LAUNCH();
    return 0;
}
extern "C" int v6_gemm_smem_bytes() { return gemm::SMEM_BYTES; }
