// K8 on the tensor cores (default for bf16, F == 256; V6B200_GLM=cuda selects the CUDA-core kernel).
// Measured on B200, 1M x 256 bf16: 99.6 us (5.1 TB/s, 0.78 of the measured HBM copy bandwidth) vs 124.2 us for the
// CUDA-core kernel in the same process; federated GLM iteration (K8 + K3 + update) 115 us vs 141 us
// (profiles/README.md).
//
// The CUDA-core kernel (rope_glm.cu::glm_logistic_kernel) reads X once but is bound by issue slots: ~160
// instructions per lane and row for the two products z = X w and g = X^T r (0.50 of HBM bandwidth,
// profiles/kernel_bench_r1b.json).  Here both products are UMMA GEMVs on the SAME shared-memory tile of X:
//
//   per 128-row tile (TMA, 3-stage ring, 4 boxes of [128 rows x 64 features], SWIZZLE_128B):
//     GEMM1  z[128 x 16]  = X_tile[128 x 256] . Wb[16 x 256]^T      A K-major (as TMA wrote it), B K-major
//            Wb rows 0/1 = w split into hi/lo bf16 (z_hi + z_lo is fp32-accurate), rows 2..15 = 0
//     sigmoid warpgroup (thread == row): z from TMEM, r = sigmoid(z + b) - y, loss / sum(r) in registers,
//            r split hi/lo -> Rb[16 x 128] (K-major over the tile's rows, rows 2..15 stay 0)
//     GEMM2  g[256 x 16] += X_tile^T[256 x 128] . Rb[16 x 128]^T    A = the same tile read MN-major
//            (docs/ROUND2_PLAN.md appendix: K runs over the rows, SBO = 1024, LBO = box stride), two M=128 halves
//   g (hi + lo columns) accumulates in TMEM over all tiles of the persistent CTA; per-CTA partials
//   [g (256), sum r, loss] go through the same fold kernel as the CUDA-core path.
//
// F == 256, bf16 X only.  TMEM: z0 | z1 | g(features 0..127) | g(features 128..255), 16 columns each.
#include <cuda.h>
#include "common.cuh"
#include "api.h"

namespace glm_tc {

constexpr int F = 256;
constexpr int TILE_ROWS = 128;
constexpr int kStages = 3;
constexpr int kThreads = 256;
constexpr int BOX_BYTES = TILE_ROWS * 128;               // one [128 rows x 64 features] box
constexpr int STAGE_BYTES = 4 * BOX_BYTES;               // 64 KB
constexpr int W_OFF = kStages * STAGE_BYTES;             // Wb: 4 k-blocks x [16 x 64] = 8 KB
constexpr int W_BYTES = 4 * 2048;
constexpr int R_OFF = W_OFF + W_BYTES;                   // Rb: 2 buffers x 2 k-blocks x [16 x 64] = 8 KB
constexpr int R_BYTES = 2 * 2048;
constexpr int BAR_OFF = R_OFF + 2 * R_BYTES;
constexpr int SMEM_BYTES = BAR_OFF + 1024 + 256;
constexpr int kTmemCols = 64;
constexpr int Z_COL = 0, G_COL = 32;

// MN-major SWIZZLE_128B operand descriptor: 64 MN-elements contiguous, 8 K-rows 128 B apart,
// sbo = bytes between 8-row K groups, lbo = bytes between 64-element MN chunks
V6_DEVINL uint64_t make_smem_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}
V6_DEVINL constexpr uint32_t make_idesc_bf16_amn(int M, int N) {       // A MN-major, B K-major
    return make_idesc_bf16(M, N) | (1u << 15);
}
// byte offset of element (n, k) inside a K-major SWIZZLE_128B [16 x 64]-per-k-block operand
V6_DEVINL uint32_t kmajor16_off(int n, int k) {
    const int kb = k >> 6, kk = k & 63;
    return kb * 2048 + (n >> 3) * 1024 + (n & 7) * 128 + (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2;
}

struct Params {
    const float* y;
    const float* w;            // [F + 1]
    float* part;               // [grid][F + 2]
    int rows;
};

__global__ void __launch_bounds__(kThreads, 1)
glm_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-B aligned; derived by pointer arithmetic so that the compiler keeps the shared address space (LDS/STS, not generic LD/ST)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* x_full = bars;                 // [3]
    uint64_t* x_empty = bars + 3;            // [3]
    uint64_t* z_full = bars + 6;             // [2]
    uint64_t* z_empty = bars + 8;            // [2]
    uint64_t* r_full = bars + 10;            // [2]
    uint64_t* r_empty = bars + 12;           // [2]
    uint64_t* g_done = bars + 14;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
    float* red = reinterpret_cast<float*>(bars + 16);          // [2][4] warp partials of sum r / loss

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (P.rows + TILE_ROWS - 1) / TILE_ROWS;

    // operands written by threads: zero everything, then Wb rows 0/1 = hi/lo split of w
    for (int i = threadIdx.x; i < (W_BYTES + 2 * R_BYTES) / 16; i += kThreads)
        reinterpret_cast<uint4*>(smem + W_OFF)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int k = threadIdx.x; k < F; k += kThreads) {
        const float wv = P.w[k];
        const __nv_bfloat16 hi = __float2bfloat16(wv);
        const __nv_bfloat16 lo = __float2bfloat16(wv - __bfloat162float(hi));
        *reinterpret_cast<__nv_bfloat16*>(smem + W_OFF + kmajor16_off(0, k)) = hi;
        *reinterpret_cast<__nv_bfloat16*>(smem + W_OFF + kmajor16_off(1, k)) = lo;
    }
    fence_proxy_async_smem();

    if (warp == 4 && lane == 0) tma_prefetch_desc(&tmap_x);
    if (warp == 5 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&z_full[b], 1); mbar_init(&z_empty[b], 4);
            mbar_init(&r_full[b], 4); mbar_init(&r_empty[b], 1);
        }
        mbar_init(g_done, 1);
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
            int it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const int st = it % kStages;
                mbar_wait(&x_empty[st], ((it / kStages) & 1) ^ 1);
                mbar_expect_tx(&x_full[st], STAGE_BYTES);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
                    tma_load_2d(smem + st * STAGE_BYTES + kb * BOX_BYTES, &tmap_x, &x_full[st], kb * 64, t * TILE_ROWS);
            }
        }
    } else if (warp == 5) {
        // ================================ MMA issuer ==================================
        constexpr uint32_t idesc_z = make_idesc_bf16(128, 16);
        constexpr uint32_t idesc_g = make_idesc_bf16_amn(128, 16);
        const uint32_t w_base = smem_u32(smem + W_OFF);
        int n_local = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) ++n_local;
        auto issue_z = [&](int it) {
            const int st = it % kStages, b = it & 1;
            mbar_wait(&x_full[st], (it / kStages) & 1);
            mbar_wait(&z_empty[b], ((it >> 1) & 1) ^ 1);
            tcgen05_fence_after();
            if (lane == 0) {
                const uint32_t xa = smem_u32(smem + st * STAGE_BYTES);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_base + Z_COL + b * 16, make_smem_desc_sw128(xa + kb * BOX_BYTES + k * 32),
                                     make_smem_desc_sw128(w_base + kb * 2048 + k * 32), idesc_z, (kb | k) ? 1u : 0u);
                umma_commit(&z_full[b]);
            }
            __syncwarp();
        };
        if (n_local > 0) issue_z(0);
        for (int it = 0; it < n_local; ++it) {
            if (it + 1 < n_local) issue_z(it + 1);                    // overlaps the sigmoid pass of tile `it`
            const int st = it % kStages, b = it & 1;
            mbar_wait(&r_full[b], (it >> 1) & 1);
            tcgen05_fence_after();
            if (lane == 0) {
                const uint32_t xa = smem_u32(smem + st * STAGE_BYTES);
                const uint32_t rb = smem_u32(smem + R_OFF + b * R_BYTES);
#pragma unroll
                for (int h = 0; h < 2; ++h)                           // feature halves: boxes 2h, 2h+1
#pragma unroll
                    for (int ks = 0; ks < TILE_ROWS / 16; ++ks)       // K = the tile's rows, 16 per step
                        umma_bf16_ss(tmem_base + G_COL + h * 16,
                                     make_smem_desc_sw128_mn(xa + 2 * h * BOX_BYTES + ks * 2048, BOX_BYTES, 1024),
                                     make_smem_desc_sw128(rb + (ks >> 2) * 2048 + (ks & 3) * 32), idesc_g,
                                     (it > 0 || ks > 0) ? 1u : 0u);
                umma_commit(&x_empty[st]);
                umma_commit(&r_empty[b]);
                if (it == n_local - 1) umma_commit(g_done);
            }
            __syncwarp();
        }
    } else if (warp < 4) {
        // ================================ sigmoid / epilogue ==========================
        const int ew = warp;                                  // compute warps are 0-3: the role warps sit in the highest ids (issue priority)
        const int trow = ew * 32 + lane;                             // row of the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        const float bias = P.w[F];
        float gb = 0.f, loss = 0.f;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int b = it & 1;
            mbar_wait(&z_full[b], (it >> 1) & 1);
            tcgen05_fence_after();
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + Z_COL, v);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&z_empty[b]);
            const long long row = (long long)t * TILE_ROWS + trow;
            float r = 0.f;
            if (row < P.rows) {
                const float z = __uint_as_float(v[b * 16]) + __uint_as_float(v[b * 16 + 1]) + bias;
                const float yy = P.y[row];
                r = 1.f / (1.f + __expf(-z)) - yy;
                loss += fmaxf(z, 0.f) - z * yy + log1pf(__expf(-fabsf(z)));
                gb += r;
            }
            mbar_wait(&r_empty[b], ((it >> 1) & 1) ^ 1);              // GEMM2 of tile it-2 has read this buffer
            const __nv_bfloat16 hi = __float2bfloat16(r);
            const __nv_bfloat16 lo = __float2bfloat16(r - __bfloat162float(hi));
            uint8_t* rbuf = smem + R_OFF + b * R_BYTES;
            *reinterpret_cast<__nv_bfloat16*>(rbuf + kmajor16_off(0, trow)) = hi;
            *reinterpret_cast<__nv_bfloat16*>(rbuf + kmajor16_off(1, trow)) = lo;
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&r_full[b]);
        }
        // ---- per-CTA partial: g (hi + lo columns) from TMEM, sum r and loss through shared memory
        float* mine = P.part + (size_t)blockIdx.x * (F + 2);
        if (it > 0) {
            mbar_wait(g_done, 0);
            tcgen05_fence_after();
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + G_COL, v);
            tmem_ld_wait();
            mine[trow] = __uint_as_float(v[0]) + __uint_as_float(v[1]);
            mine[128 + trow] = __uint_as_float(v[16]) + __uint_as_float(v[17]);
        } else {
            mine[trow] = 0.f;
            mine[128 + trow] = 0.f;
        }
        gb = warp_sum(gb);
        loss = warp_sum(loss);
        if (lane == 0) { red[ew] = gb; red[4 + ew] = loss; }
        asm volatile("bar.sync 1, 128;" ::: "memory");                // the 4 epilogue warps only
        if (trow == 0) {
            mine[F] = red[0] + red[1] + red[2] + red[3];
            mine[F + 1] = red[4] + red[5] + red[6] + red[7];
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 6) { tcgen05_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}

}  // namespace glm_tc

// X: [rows, 256] bf16 row-major; y: [rows]; w: [257]; part: [>= grid][258] scratch.  Returns the grid size used
// (number of partials to fold) or a negative / CUDA error code.
extern "C" int v6_glm_logistic_grad_tc(const void* X, const float* y, const float* w, float* part, int max_parts, int rows,
                                       int F, cudaStream_t s) {
    using namespace glm_tc;
    if (F != glm_tc::F || rows < 1 || max_parts < 1) return -(int)cudaErrorInvalidValue;
    alignas(64) CUtensorMap tx;
    if (v6_make_tmap_2d_bf16(&tx, (uint64_t)X, (uint64_t)rows, (uint64_t)F, (uint64_t)F * 2, TILE_ROWS, 64, 1)) return -2;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(glm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
            return -(int)cudaErrorInvalidValue;
        attr_set = true;
    }
    const int ntiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
    int grid = ntiles < 148 ? ntiles : 148;
    if (grid > max_parts) grid = max_parts;
    Params P;
    P.y = y; P.w = w; P.part = part; P.rows = rows;
    glm_tc_kernel<<<grid, kThreads, SMEM_BYTES, s>>>(tx, P);
    if (cudaGetLastError() != cudaSuccess) return -(int)cudaErrorLaunchFailure;
    return grid;
}
