// Implicit-GEMM convolution family on tcgen05 / TMEM / TMA for sm_100a, hand-written (no CUTLASS, no cuDNN):
//
//   FPROP   Y [pix, Cout]        = im2col(X)[pix, (r,s,ci)] . W[Cout, (r,s,ci)]^T          (+ fused BatchNorm statistics)
//   DGRAD   dX[pix, Cin]         = im2col(dY)[pix, (r,s,co)] . W[co, (R-1-r,S-1-s), Cin]    (stride 1)
//   WGRAD   dW[Cout, (r,s,ci)]  += dY[pix, Cout]^T . im2col(X)[pix, (r,s,ci)]               (fp32, split-K over pixels)
//
// Activations are NHWC bf16, filters are [Cout, R, S, Cin] bf16 (the channels-last bf16 shadow that the fused
// optimizer / aggregation kernels maintain), filter gradients accumulate in fp32 straight into the flat gradient
// buffer of the model.  1x1 / stride-1 layers and nn.Linear are the degenerate case (plain 2-D TMA tiles), so the same
// kernel is the dX = dY.W and dW = dY^T.X GEMM of the transformer layers.
//
// What makes it "implicit": the A operand of FPROP / DGRAD and the B operand of WGRAD are gathered by TMA **im2col**
// tensor maps (cuTensorMapEncodeIm2col): one `cp.async.bulk.tensor.4d...im2col` per (filter tap, 64-channel block)
// lands 128 (or 64) consecutive output pixels x 64 channels in shared memory, already SWIZZLE_128B, padding
// zero-filled and the convolution stride applied by the copy engine.  Nothing is ever unfolded in HBM.
//
// Operand layouts (UMMA shared-memory descriptors):
//   FPROP : A K-major (pixels x channels),  B K-major  (W rows = Cout, K contiguous)
//   DGRAD : A K-major,                      B MN-major (W[co][tap][ci]: ci = N contiguous, co = K rows)
//   WGRAD : A MN-major (dY[pix][co]: co = M contiguous, pix = K rows),  B MN-major (X[pix][ci])
// -> no transposed copy of the filter or of the activations exists anywhere.
//
// Structure (persistent, warp-specialised, 256 threads, one CTA per SM): warp 0 TMA producer, warp 1 MMA issuer,
// warp 2 TMEM allocator, warps 4-7 epilogue; 4-stage smem ring (48 KB / stage), 2 TMEM accumulator stages.
// FPROP epilogue = bf16 tile through a swizzled staging box + TMA store, and (optionally) the BatchNorm batch
// statistics of the layer: per-column sum / sum-of-squares of the bf16-rounded outputs, accumulated per CTA over its
// run of row-tiles, folded in a fixed order by the last CTA to finish a column block, which also does the per-channel
// finalize (mean, rstd, scale/bias, running statistics).  The separate statistics pass over Y (bn.cu::bn_stats_kernel,
// 9 % of a ResNet-50 round) disappears; the result is deterministic.
// WGRAD epilogue = red.global.add.v4.f32 of the fp32 accumulators into dW (split-K partial sums meet in L2).
#include <cuda.h>
#include "common.cuh"
#include "api.h"

namespace igemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr int kTmemCols = 512;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;        // 16 KB
constexpr int B_BYTES = 256 * BLOCK_K * 2;            // 32 KB (N tile <= 256)
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int BOX_BYTES = 64 * 128;                   // one 64-row x 64-element SWIZZLE_128B box (MN-major operands)
constexpr int STG_OFF = kStages * STAGE_BYTES;        // epilogue staging: one 32x64 bf16 box per epilogue warp
constexpr int STG_WARP_BYTES = 32 * 128;
constexpr int STAT_OFF = STG_OFF + 4 * STG_WARP_BYTES;
constexpr int STAT_FLOATS = 4 * 2 * 256;              // [warp][sum | sumsq][256 columns]
constexpr int STAT_BYTES = 2 * STAT_FLOATS * 4;       // double-buffered by tile parity
constexpr int BAR_OFF = STAT_OFF + STAT_BYTES;
constexpr int SMEM_BYTES = BAR_OFF + 1024 /*align slack*/ + 256 /*barriers*/;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
constexpr int kThreads = 256;
constexpr int MAX_SLOTS = 148;

enum { FPROP = 0, DGRAD = 1, WGRAD = 2 };

struct Params {
    int mode;
    int M, N;                  // output space: FPROP/DGRAD [pixels, channels]; WGRAD [Cout, taps*Cin]
    int num_kb;                // 64-wide blocks of the whole reduction dimension
    int block_n;               // N tile (64/128/192/256); WGRAD: 256 (4 chunks of 64 columns)
    int a_im2col, b_im2col;    // operand gathered through an im2col tensor map (else plain tiled map)
    int PQ, Q;                 // output pixels per image / per image row of the im2col traversal
    int stride, pad;           // base pixel of output (p,q) = (p*stride - pad, q*stride - pad)
    int S;                     // filter width (tap = r*S + s)
    int cblocks;               // 64-channel blocks per filter tap
    int taps;                  // R*S
    int flip;                  // DGRAD: filter tap used = taps-1-tap
    int act;                   // bf16 epilogue: 0 none, 1 gelu(erf), 2 relu
    const float* bias;         // [N] or null
    // WGRAD
    float* dw; long long ldw; int splits, kb_per_split; float out_scale;
    // FPROP BatchNorm statistics (null gamma -> off)
    const float* gamma; const float* beta; float* running_mean; float* running_var; long long* num_batches_tracked;
    float* mean_out; float* rstd_out; float* scale_out; float* bias_out;
    float* part; int* counters; float eps, momentum;
};

V6_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// MN-major SWIZZLE_128B operand: 64 MN-elements (128 B) contiguous per K row, 8-row K groups `sbo` bytes apart,
// 64-element MN chunks `lbo` bytes apart (validated by glm_tc.cu, docs/ROUND2_PLAN.md appendix).
V6_DEVINL uint64_t make_smem_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

V6_DEVINL void tma_load_im2col_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c, int w, int h, int n, int off_w, int off_h) {
    const uint16_t ow = (uint16_t)off_w, oh = (uint16_t)off_h;
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                 :: "r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh)
                 : "memory");
}
V6_DEVINL void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
V6_DEVINL void red_add_f4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
V6_DEVINL void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }      // the 4 epilogue warps only

struct Item { int m_blk, n_blk, kb0, kb1; };

__global__ void __launch_bounds__(kThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
             const __grid_constant__ CUtensorMap tmap_c, const Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + kAccStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + kAccStages);
    int* s_flag = reinterpret_cast<int*>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = (P.M + BLOCK_M - 1) / BLOCK_M;
    const int num_n = (P.N + P.block_n - 1) / P.block_n;
    const int num_tiles = num_m * num_n;
    const int num_items = P.mode == WGRAD ? num_tiles * P.splits : num_tiles;

    auto decode = [&](int item) {
        Item it;
        const int tile = P.mode == WGRAD ? item % num_tiles : item;
        it.m_blk = tile % num_m;                                  // m fastest: concurrent CTAs share the filter block
        it.n_blk = tile / num_m;
        if (P.mode == WGRAD) {
            const int split = item / num_tiles;
            it.kb0 = split * P.kb_per_split;
            it.kb1 = min(P.num_kb, it.kb0 + P.kb_per_split);
        } else { it.kb0 = 0; it.kb1 = P.num_kb; }
        return it;
    };
    // WGRAD: 64-column chunks of this N tile that exist (each chunk = one (tap, channel block))
    auto wgrad_chunks = [&](int n_blk) { return min(4, (P.N >> 6) - n_blk * 4); };

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); if (P.mode != WGRAD) tma_prefetch_desc(&tmap_c); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================ TMA producer ============================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const Item it = decode(item);
                int pn = 0, pw = 0, ph = 0;                      // base pixel of the tile's first row (FPROP / DGRAD im2col)
                const int m0 = it.m_blk * BLOCK_M;
                if (P.mode != WGRAD && P.a_im2col) {
                    pn = m0 / P.PQ;
                    const int rem = m0 - pn * P.PQ;
                    ph = (rem / P.Q) * P.stride - P.pad;
                    pw = (rem % P.Q) * P.stride - P.pad;
                }
                const int na = P.mode == WGRAD ? min(2, (P.M - m0 + 63) >> 6) : 0;
                const int nch = P.mode == WGRAD ? wgrad_chunks(it.n_blk) : P.block_n >> 6;
                for (int kb = it.kb0; kb < it.kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    uint64_t* fb = &full_bar[stage];
                    if (P.mode == FPROP) {
                        mbar_expect_tx(fb, A_BYTES + P.block_n * 128);
                        if (P.a_im2col) {
                            const int tap = kb / P.cblocks, cb = kb - tap * P.cblocks;
                            tma_load_im2col_4d(sa, &tmap_a, fb, cb * 64, pw, ph, pn, tap % P.S, tap / P.S);
                        } else {
                            tma_load_2d(sa, &tmap_a, fb, kb * BLOCK_K, m0);
                        }
                        tma_load_2d(sb, &tmap_b, fb, kb * BLOCK_K, it.n_blk * P.block_n);
                    } else if (P.mode == DGRAD) {
                        mbar_expect_tx(fb, A_BYTES + nch * BOX_BYTES);
                        const int tap = kb / P.cblocks, cb = kb - tap * P.cblocks;
                        if (P.a_im2col) tma_load_im2col_4d(sa, &tmap_a, fb, cb * 64, pw, ph, pn, tap % P.S, tap / P.S);
                        else tma_load_2d(sa, &tmap_a, fb, kb * BLOCK_K, m0);
                        const int wtap = P.flip ? P.taps - 1 - tap : tap;
                        for (int j = 0; j < nch; ++j)
                            tma_load_3d(sb + j * BOX_BYTES, &tmap_b, fb, it.n_blk * P.block_n + j * 64, wtap, cb * 64);
                    } else {
                        mbar_expect_tx(fb, (na + nch) * BOX_BYTES);
                        const int pix0 = kb * BLOCK_K;
                        for (int i = 0; i < na; ++i) tma_load_2d(sa + i * BOX_BYTES, &tmap_a, fb, m0 + i * 64, pix0);
                        int bn = 0, bw = 0, bh = 0;
                        if (P.b_im2col) {
                            bn = pix0 / P.PQ;
                            const int rem = pix0 - bn * P.PQ;
                            bh = (rem / P.Q) * P.stride - P.pad;
                            bw = (rem % P.Q) * P.stride - P.pad;
                        }
                        for (int j = 0; j < nch; ++j) {
                            const int chunk = it.n_blk * 4 + j;
                            if (P.b_im2col) {
                                const int tap = chunk / P.cblocks, cb = chunk - tap * P.cblocks;
                                tma_load_im2col_4d(sb + j * BOX_BYTES, &tmap_b, fb, cb * 64, bw, bh, bn, tap % P.S, tap / P.S);
                            } else {
                                tma_load_2d(sb + j * BOX_BYTES, &tmap_b, fb, chunk * 64, pix0);
                            }
                        }
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ==============================
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const Item it = decode(item);
            const int n_umma = P.mode == WGRAD ? wgrad_chunks(it.n_blk) * 64 : P.block_n;
            uint32_t idesc = make_idesc_bf16(BLOCK_M, n_umma);
            if (P.mode == DGRAD) idesc |= 1u << 16;                       // B MN-major
            if (P.mode == WGRAD) idesc |= (1u << 15) | (1u << 16);        // A and B MN-major
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            for (int kb = it.kb0; kb < it.kb1; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t da = P.mode == WGRAD ? make_smem_desc_sw128_mn(sa + k * 2048, BOX_BYTES, 1024)
                                                            : make_smem_desc_sw128(sa + k * UMMA_K * 2);
                        const uint64_t db = P.mode == FPROP ? make_smem_desc_sw128(sb + k * UMMA_K * 2)
                                                            : make_smem_desc_sw128_mn(sb + k * 2048, BOX_BYTES, 1024);
                        umma_bf16_ss(d_tmem, da, db, idesc, (kb > it.kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == it.kb1 - 1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ============================ epilogue ================================
        const int ew = warp - 4;                      // TMEM lanes [32*ew, 32*ew+32)
        const int et = ew * 32 + lane;                // epilogue thread id: owns columns 2*et, 2*et+1 of the statistics
        const bool stats = P.mode == FPROP && P.gamma != nullptr;
        int acc = 0; uint32_t acc_phase = 0;
        float run1[2] = {0.f, 0.f}, run2[2] = {0.f, 0.f};
        int run_first_m = -1, tile_par = 0;
        float* stat = reinterpret_cast<float*>(smem + STAT_OFF);
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const Item it = decode(item);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tcgen05_fence_after();
            const uint32_t t_row = tmem_base + acc * 256 + ((uint32_t)(ew * 32) << 16);
            if (P.mode == WGRAD) {
                // fp32 accumulators -> red.add into dW[row][n_blk*256 + col]; row = output channel
                const int row = it.m_blk * BLOCK_M + ew * 32 + lane;
                const int ncols = wgrad_chunks(it.n_blk) * 64;
                float* drow = P.dw + (long long)row * P.ldw + it.n_blk * 256;
#pragma unroll 1
                for (int c = 0; c < ncols; c += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + c, v);
                    tmem_ld_wait();
                    if (row < P.M) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            red_add_f4(drow + c + j, __uint_as_float(v[j]) * P.out_scale, __uint_as_float(v[j + 1]) * P.out_scale,
                                       __uint_as_float(v[j + 2]) * P.out_scale, __uint_as_float(v[j + 3]) * P.out_scale);
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
                continue;
            }
            const int row0 = it.m_blk * BLOCK_M + ew * 32;
            uint8_t* stg = smem + STG_OFF + ew * STG_WARP_BYTES;
            float cs1[4][2], cs2[4][2];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                cs1[ch][0] = cs1[ch][1] = cs2[ch][0] = cs2[ch][1] = 0.f;
                const int c = ch * 64;
                if (c < P.block_n) {                                       // uniform
                    uint32_t v[2][32];
                    tmem_ld_32x32b_x32(t_row + c, v[0]);
                    tmem_ld_32x32b_x32(t_row + c + 32, v[1]);
                    tmem_ld_wait();
                    const int col0 = it.n_blk * P.block_n + c;
                    if (row0 < P.M && col0 < P.N) {                        // warp-uniform
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
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                                make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) { tma_store_2d(&tmap_c, stg, col0, row0); tma_store_commit(); }
                        if (stats) {
                            // column sums of the bf16-rounded tile: lane owns columns 2*lane, 2*lane+1 of the box
                            const int chunk16 = lane >> 2, within = (lane & 3) * 4;
                            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 8
                            for (int r = 0; r < 32; ++r) {
                                const uint32_t wv = *reinterpret_cast<const uint32_t*>(stg + r * 128 + ((chunk16 ^ (r & 7)) << 4) + within);
                                const float x0 = __uint_as_float(wv << 16), x1 = __uint_as_float(wv & 0xffff0000u);
                                a0 += x0; a1 += x1; b0 = fmaf(x0, x0, b0); b1 = fmaf(x1, x1, b1);
                            }
                            cs1[ch][0] = a0; cs1[ch][1] = a1; cs2[ch][0] = b0; cs2[ch][1] = b1;
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }

            if (stats) {
                // combine the 4 row-quadrants of the tile, accumulate over this CTA's run of row tiles of the column block
                float* sp = stat + tile_par * STAT_FLOATS + ew * 512;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    if (ch * 64 < P.block_n) {
                        *reinterpret_cast<float2*>(sp + ch * 64 + 2 * lane) = make_float2(cs1[ch][0], cs1[ch][1]);
                        *reinterpret_cast<float2*>(sp + 256 + ch * 64 + 2 * lane) = make_float2(cs2[ch][0], cs2[ch][1]);
                    }
                }
                epi_bar_sync();
                if (run_first_m < 0) run_first_m = it.m_blk;
                if (2 * et < P.block_n) {
                    const float* sq = stat + tile_par * STAT_FLOATS;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const float2 u1 = *reinterpret_cast<const float2*>(sq + w * 512 + 2 * et);
                        const float2 u2 = *reinterpret_cast<const float2*>(sq + w * 512 + 256 + 2 * et);
                        run1[0] += u1.x; run1[1] += u1.y; run2[0] += u2.x; run2[1] += u2.y;
                    }
                }
                tile_par ^= 1;
                const int nxt = item + gridDim.x;
                const bool flush = nxt >= num_items || decode(nxt).n_blk != it.n_blk;
                if (flush) {
                    const int slots = min(num_m, (int)gridDim.x);
                    float* pb = P.part + ((size_t)(it.n_blk * slots + run_first_m) * 2) * P.block_n;
                    if (2 * et < P.block_n) {
                        __stcg(reinterpret_cast<float2*>(pb + 2 * et), make_float2(run1[0], run1[1]));
                        __stcg(reinterpret_cast<float2*>(pb + P.block_n + 2 * et), make_float2(run2[0], run2[1]));
                    }
                    run1[0] = run1[1] = run2[0] = run2[1] = 0.f;
                    run_first_m = -1;
                    __threadfence();
                    epi_bar_sync();
                    if (et == 0) {
                        int old;
                        asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(old) : "l"(P.counters + it.n_blk) : "memory");
                        const int last = old == slots - 1;
                        if (last) P.counters[it.n_blk] = 0;              // every expected arrival has happened
                        *s_flag = last;
                    }
                    epi_bar_sync();
                    const bool last = *s_flag != 0;
                    epi_bar_sync();                                      // s_flag may be rewritten by the next flush
                    if (last && 2 * et < P.block_n) {
                        // fixed-order fold over the slots (deterministic) + per-channel finalize
                        const float* fb = P.part + ((size_t)it.n_blk * slots * 2) * P.block_n;
                        double t1[2] = {0.0, 0.0}, t2[2] = {0.0, 0.0};
                        for (int s0 = 0; s0 < slots; s0 += 4) {
                            float2 u1[4], u2[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int s = s0 + u;
                                u1[u] = s < slots ? __ldcg(reinterpret_cast<const float2*>(fb + (size_t)s * 2 * P.block_n + 2 * et)) : make_float2(0.f, 0.f);
                                u2[u] = s < slots ? __ldcg(reinterpret_cast<const float2*>(fb + (size_t)s * 2 * P.block_n + P.block_n + 2 * et)) : make_float2(0.f, 0.f);
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) { t1[0] += u1[u].x; t1[1] += u1[u].y; t2[0] += u2[u].x; t2[1] += u2[u].y; }
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int c = it.n_blk * P.block_n + 2 * et + q;
                            if (c < P.N) {
                                const double invR = 1.0 / (double)P.M;
                                const double mean_d = t1[q] * invR;
                                double var_d = t2[q] * invR - mean_d * mean_d;
                                if (var_d < 0.0) var_d = 0.0;
                                const float mean = (float)mean_d, var = (float)var_d;
                                const float rstd = rsqrtf(var + P.eps);
                                P.mean_out[c] = mean;
                                P.rstd_out[c] = rstd;
                                const float sc = P.gamma[c] * rstd;
                                P.scale_out[c] = sc;
                                P.bias_out[c] = P.beta[c] - mean * sc;
                                if (P.running_mean) {
                                    const float unbiased = P.M > 1 ? var * (float)P.M / (float)(P.M - 1) : var;
                                    P.running_mean[c] = (1.f - P.momentum) * P.running_mean[c] + P.momentum * mean;
                                    P.running_var[c] = (1.f - P.momentum) * P.running_var[c] + P.momentum * unbiased;
                                }
                                if (P.num_batches_tracked && c == 0) *P.num_batches_tracked += 1;
                            }
                        }
                    }
                }
            }
        }
        if (P.mode != WGRAD && lane == 0) tma_store_wait_all();        // staging must outlive the last store
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) { tcgen05_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}

}  // namespace igemm

// ------------------------------------------------------------------------------------------------- host side
extern "C" long long v6_igemm_scratch_floats() { return 64 + (long long)2048 * 2 * igemm::MAX_SLOTS; }   // counters | partials

namespace {

struct ConvGeom {
    int N, H, W, C;          // the im2col-side activation tensor (NHWC)
    int R, S, stride, pad;   // filter window / traversal over that tensor
    int P, Q;                // traversal output size
};

int set_smem_attr() {
    static bool done = false;
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(igemm::igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, igemm::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        done = true;
    }
    return 0;
}

int im2col_map(void* out, const void* ptr, const ConvGeom& g, int pixels) {
    return v6_make_tmap_im2col_bf16(out, (uint64_t)ptr, g.C, g.W, g.H, g.N, -g.pad, -g.pad, g.pad - (g.S - 1), g.pad - (g.R - 1), 64, pixels,
                                    g.stride, g.stride);
}

int pick_block_n(int M, int N) {
    const int num_m = (M + 127) / 128;
    if (N <= 64) return 64;
    if (N <= 128) return 128;
    // keep >= ~1 wave of tiles: prefer the widest tile that still yields >= 120 tiles
    for (int bn : {256, 128}) if (num_m * ((N + bn - 1) / bn) >= 120) return bn;
    return 64;
}

int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const igemm::Params& P, cudaStream_t s) {
    using namespace igemm;
    int rc = set_smem_attr();
    if (rc) return rc;
    const int num_m = (P.M + BLOCK_M - 1) / BLOCK_M, num_n = (P.N + P.block_n - 1) / P.block_n;
    const int items = P.mode == WGRAD ? num_m * num_n * P.splits : num_m * num_n;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const int cap = sms < MAX_SLOTS ? sms : MAX_SLOTS;
    const int grid = items < cap ? items : cap;
    igemm_kernel<<<grid, kThreads, SMEM_BYTES, s>>>(ta, tb, tc, P);
    V6_CHECK_LAUNCH();
    return 0;
}

}  // namespace

// y[N,P,Q,Cout] = conv(x[N,H,W,Cin], w[Cout,R,S,Cin]); bn != 0: also the BatchNorm batch statistics of y (see Params)
extern "C" int v6_conv_fprop(const void* x, const void* w, void* y, const float* bias, int act, int N, int H, int W, int Cin, int Cout,
                             int R, int S, int stride, int pad, const float* gamma, const float* beta, float* running_mean,
                             float* running_var, long long* num_batches_tracked, float* mean_out, float* rstd_out,
                             float* scale_bias_out, float* scratch, float eps, float momentum, int force_im2col, cudaStream_t stream) {
    using namespace igemm;
    if (Cin % 64 != 0 || Cout % 8 != 0 || R != S) return (int)cudaErrorInvalidValue;
    const int Pp = (H + 2 * pad - R) / stride + 1, Qq = (W + 2 * pad - S) / stride + 1;
    const long long Mll = (long long)N * Pp * Qq;
    if (Mll >= (1LL << 31)) return (int)cudaErrorInvalidValue;
    const int M = (int)Mll, K = R * S * Cin;
    const bool plain = R == 1 && stride == 1 && pad == 0 && !force_im2col;
    Params P = {};
    P.mode = FPROP; P.M = M; P.N = Cout; P.num_kb = K / 64; P.block_n = pick_block_n(M, Cout);
    P.a_im2col = plain ? 0 : 1; P.PQ = Pp * Qq; P.Q = Qq; P.stride = stride; P.pad = pad; P.S = S; P.cblocks = Cin / 64; P.taps = R * S;
    P.act = act; P.bias = bias;
    if (gamma) {
        if (Cout % 64 != 0) return (int)cudaErrorInvalidValue;
        P.gamma = gamma; P.beta = beta; P.running_mean = running_mean; P.running_var = running_var;
        P.num_batches_tracked = num_batches_tracked; P.mean_out = mean_out; P.rstd_out = rstd_out;
        P.scale_out = scale_bias_out; P.bias_out = scale_bias_out + Cout; P.eps = eps; P.momentum = momentum;
        P.counters = reinterpret_cast<int*>(scratch); P.part = scratch + 64;
        const int num_n = (Cout + P.block_n - 1) / P.block_n;
        if (num_n > 64 || (long long)num_n * P.block_n > 2048) return (int)cudaErrorInvalidValue;
    }
    alignas(64) CUtensorMap ta, tb, tc;
    if (plain) { if (v6_make_tmap_2d_bf16(&ta, (uint64_t)x, M, Cin, (uint64_t)Cin * 2, BLOCK_M, BLOCK_K, 1)) return -2; }
    else { ConvGeom g{N, H, W, Cin, R, S, stride, pad, Pp, Qq}; if (im2col_map(&ta, x, g, BLOCK_M)) return -2; }
    if (v6_make_tmap_2d_bf16(&tb, (uint64_t)w, Cout, K, (uint64_t)K * 2, P.block_n, BLOCK_K, 1)) return -2;
    if (v6_make_tmap_2d_bf16(&tc, (uint64_t)y, M, Cout, (uint64_t)Cout * 2, 32, 64, 1)) return -2;
    return launch(ta, tb, tc, P, stream);
}

// dx[N,H,W,Cin] = conv_transpose(dy[N,P,Q,Cout], w[Cout,R,S,Cin]) for stride 1 (P = H + 2*pad - R + 1)
extern "C" int v6_conv_dgrad(const void* dy, const void* w, void* dx, int N, int H, int W, int Cin, int Cout, int R, int S, int pad,
                             int force_im2col, cudaStream_t stream) {
    using namespace igemm;
    if (Cout % 64 != 0 || Cin % 8 != 0 || R != S) return (int)cudaErrorInvalidValue;
    const int Pp = H + 2 * pad - R + 1, Qq = W + 2 * pad - S + 1;        // dY spatial size
    const long long Mll = (long long)N * H * W;
    if (Mll >= (1LL << 31)) return (int)cudaErrorInvalidValue;
    const int M = (int)Mll;
    const bool plain = R == 1 && pad == 0 && !force_im2col;
    Params P = {};
    P.mode = DGRAD; P.M = M; P.N = Cin; P.num_kb = R * S * (Cout / 64); P.block_n = pick_block_n(M, Cin);
    P.a_im2col = plain ? 0 : 1; P.PQ = H * W; P.Q = W; P.stride = 1; P.pad = R - 1 - pad; P.S = S; P.cblocks = Cout / 64; P.taps = R * S; P.flip = 1;
    alignas(64) CUtensorMap ta, tb, tc;
    if (plain) { if (v6_make_tmap_2d_bf16(&ta, (uint64_t)dy, M, Cout, (uint64_t)Cout * 2, BLOCK_M, BLOCK_K, 1)) return -2; }
    else { ConvGeom g{N, Pp, Qq, Cout, R, S, 1, R - 1 - pad, H, W}; if (im2col_map(&ta, dy, g, BLOCK_M)) return -2; }
    {
        const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)(R * S), (uint64_t)Cout};
        const uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)R * S * Cin * 2};
        const uint32_t box[3] = {64, 1, 64};
        if (v6_make_tmap_tiled_bf16(&tb, (uint64_t)w, 3, dims, strides, box, 1)) return -2;
    }
    if (v6_make_tmap_2d_bf16(&tc, (uint64_t)dx, M, Cin, (uint64_t)Cin * 2, 32, 64, 1)) return -2;
    return launch(ta, tb, tc, P, stream);
}

// dw[Cout,R,S,Cin] (fp32) += scale * dy[N,P,Q,Cout]^T . im2col(x[N,H,W,Cin])
extern "C" int v6_conv_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int Cin, int Cout, int R, int S, int stride,
                             int pad, float scale, int splits, int force_im2col, cudaStream_t stream) {
    using namespace igemm;
    if (Cin % 64 != 0 || Cout % 8 != 0 || R != S) return (int)cudaErrorInvalidValue;
    const int Pp = (H + 2 * pad - R) / stride + 1, Qq = (W + 2 * pad - S) / stride + 1;
    const long long pix = (long long)N * Pp * Qq;
    if (pix >= (1LL << 31)) return (int)cudaErrorInvalidValue;
    const bool plain = R == 1 && stride == 1 && pad == 0 && !force_im2col;
    Params P = {};
    P.mode = WGRAD; P.M = Cout; P.N = R * S * Cin; P.num_kb = (int)((pix + 63) / 64); P.block_n = 256;
    P.b_im2col = plain ? 0 : 1; P.PQ = Pp * Qq; P.Q = Qq; P.stride = stride; P.pad = pad; P.S = S; P.cblocks = Cin / 64; P.taps = R * S;
    P.dw = dw; P.ldw = (long long)R * S * Cin; P.out_scale = scale;
    const int tiles = ((Cout + 127) / 128) * ((P.N / 64 + 3) / 4);
    if (splits <= 0) {
        splits = (148 + tiles - 1) / tiles;                              // about one wave of work items
        const int max_splits = P.num_kb / 4 > 0 ? P.num_kb / 4 : 1;      // >= 4 k-blocks per item
        if (splits > max_splits) splits = max_splits;
    }
    if (splits > P.num_kb) splits = P.num_kb;
    P.kb_per_split = (P.num_kb + splits - 1) / splits;
    P.splits = (P.num_kb + P.kb_per_split - 1) / P.kb_per_split;
    alignas(64) CUtensorMap ta, tb, tc;
    if (v6_make_tmap_2d_bf16(&ta, (uint64_t)dy, (uint64_t)pix, Cout, (uint64_t)Cout * 2, 64, 64, 1)) return -2;
    if (plain) { if (v6_make_tmap_2d_bf16(&tb, (uint64_t)x, (uint64_t)pix, Cin, (uint64_t)Cin * 2, 64, 64, 1)) return -2; }
    else { ConvGeom g{N, H, W, Cin, R, S, stride, pad, Pp, Qq}; if (im2col_map(&tb, x, g, 64)) return -2; }
    tc = ta;
    return launch(ta, tb, tc, P, stream);
}
