// Deterministic in-kernel tree reduction of per-column sums over the rows of an [R, C] tensor, shared by the
// BatchNorm kernels (bn.cu) and the bias / activation backward (act.cu).
#pragma once
#include "common.cuh"

namespace tree {

constexpr int THREADS = 256;

// ------------------------------------------------------------------ in-kernel tree reduction
// A reduce CTA covers a 64-channel slice (blockIdx.y) and
// a strided set of rows (blockIdx.x); its [2*SW] partial goes to global memory, then a two-level
// "last CTA to arrive folds" tree (groups of G1 row-CTAs, then the groups of the slice) produces the
// per-channel totals inside the same launch and the last CTA of each slice does the per-channel math.
// Every fold sums its inputs in a fixed order, so the result does not depend on which CTA happens to
// arrive last (deterministic, no floating-point atomics); the counters reset themselves.
// (v2 profile, profiles/launches_resnet50_fusedbn_r1.txt: finalize 12.5 us per launch, latency-bound,
// 15% of the forward; stats at 2 TB/s because only 2 CTAs/SM were resident.)
constexpr int G1 = 16;
constexpr int ONE_LEVEL_MAX = 128;
constexpr int MAX_CTAS = 148 * 8;                    // upper bound on reduce CTAs per launch
constexpr int L2_SLOTS = 256;                        // >= MAX_CTAS / G1 + slices
constexpr size_t SCR_L2 = (size_t)MAX_CTAS * 128;    // float offsets into the scratch buffer
constexpr size_t SCR_CNT1 = SCR_L2 + (size_t)L2_SLOTS * 128;
constexpr size_t SCR_CNT2 = SCR_CNT1 + L2_SLOTS;
constexpr size_t SCR_FLOATS = SCR_CNT2 + 64;

struct Red {
    float* l1;      // [slice][row-CTA][2*SW]
    float* l2;      // [slice][group][2*SW]
    int* cnt1;      // [slice][group]   (zero between launches)
    int* cnt2;      // [slice]
};
static inline Red make_red(float* scratch) {
    return Red{scratch, scratch + SCR_L2, reinterpret_cast<int*>(scratch + SCR_CNT1), reinterpret_cast<int*>(scratch + SCR_CNT2)};
}

// true in exactly one CTA: the last of `expected` to arrive at `counter` (which it resets to 0).
// CUTLASS-semaphore pattern: the CTA barrier orders every thread's partial writes before thread 0's
// acq_rel arrival (release is cumulative), and the acquire side orders the last CTA's reads after it.
V6_DEVINL bool arrive_last(int* counter, int expected, int* s_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        int old;
        asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
        const int last = old == expected - 1;
        if (last) *counter = 0;                 // every expected arrival has happened: plain reset for the next launch
        *s_flag = last;
    }
    __syncthreads();
    return *s_flag != 0;
}

// dst[col] = sum_p src[p][col], p < n; cols in {16..128}.  float4 columns: cols/4 threads per row, THREADS/(cols/4)
// row lanes, up to 8 independent 16 B loads in flight per thread; fixed summation order.
V6_DEVINL void fold_rows(const float* __restrict__ src, int n, int cols, float* red, float* dst) {
    const int c4 = cols >> 2, lanes = THREADS / c4;
    const int col4 = threadIdx.x % c4, ln = threadIdx.x / c4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p0 = ln; p0 < n; p0 += 8 * lanes) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int p = p0 + u * lanes;
            v[u] = p < n ? __ldcg(reinterpret_cast<const float4*>(src + (size_t)p * cols) + col4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    reinterpret_cast<float4*>(red)[threadIdx.x] = s;          // red: [lanes][cols]
    __syncthreads();
    if (threadIdx.x < cols) {
        float t = 0.f;
        for (int l = 0; l < lanes; ++l) t += red[l * cols + threadIdx.x];
        dst[threadIdx.x] = t;
    }
}

// Reduce the per-thread accumulators a[8], b[8] (thread = 8 channels of one row lane) over the whole
// slice.  Returns true in the one CTA per slice that ends up with the totals in tot[0..SW) (sum a) and
// tot[SW..2SW) (sum b).
V6_DEVINL bool slice_reduce(float (&a)[8], float (&b)[8], const Red& rd, int SW, float* tot) {
    __shared__ __align__(16) float red[8 * 128];
    __shared__ int s_flag;
    const int CGS = SW >> 3, cols = 2 * SW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int off = CGS; off < 32; off <<= 1) {           // lanes with the same channel group hold different rows
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a[k] += __shfl_xor_sync(0xffffffffu, a[k], off);
            b[k] += __shfl_xor_sync(0xffffffffu, b[k], off);
        }
    }
    if (lane < CGS) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { red[warp * 128 + lane * 8 + k] = a[k]; red[warp * 128 + SW + lane * 8 + k] = b[k]; }
    }
    __syncthreads();
    const int slice = blockIdx.y, rc = blockIdx.x, nrc = gridDim.x;
    const int g1 = nrc <= ONE_LEVEL_MAX ? nrc : G1;          // few row-CTAs: one fold does it all
    const int ngrp = (nrc + g1 - 1) / g1;
    if (threadIdx.x < cols) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) s += red[w * 128 + threadIdx.x];
        __stcg(rd.l1 + ((size_t)slice * nrc + rc) * cols + threadIdx.x, s);
    }
    const int grp = rc / g1;
    const int gsz = min(g1, nrc - grp * g1);
    if (!arrive_last(rd.cnt1 + slice * ngrp + grp, gsz, &s_flag)) return false;
    if (ngrp == 1) {
        fold_rows(rd.l1 + (size_t)slice * nrc * cols, gsz, cols, red, tot);
    } else {
        fold_rows(rd.l1 + ((size_t)slice * nrc + (size_t)grp * g1) * cols, gsz, cols, red, rd.l2 + ((size_t)slice * ngrp + grp) * cols);
        if (!arrive_last(rd.cnt2 + slice, ngrp, &s_flag)) return false;
        fold_rows(rd.l2 + (size_t)slice * ngrp * cols, ngrp, cols, red, tot);
    }
    __syncthreads();
    return true;
}


// one full wave of resident CTAs for a reduce kernel (occupancy queried once per kernel)
template <typename K>
static int wave_ctas(K kernel, int& cache) {
    if (cache == 0) {
        int dev = 0, sms = 148, occ = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, THREADS, 0) != cudaSuccess || occ < 1) occ = 1;
        cache = occ * sms > MAX_CTAS ? MAX_CTAS : occ * sms;
    }
    return cache;
}

// reduce grids: x = row-CTAs, y = 64-channel slices
static inline dim3 reduce_grid(long long R, int C, int target) {
    const int SW = C < 64 ? C : 64, slices = C / SW, RL = THREADS / (SW >> 3);
    long long nrc = (R + 4LL * RL - 1) / (4LL * RL);          // >= 4 rows per thread: the unrolled fast path, fewer partials
    const long long cap = target / slices > 0 ? target / slices : 1;
    if (nrc > cap) nrc = cap;
    return dim3((unsigned)(nrc < 1 ? 1 : nrc), (unsigned)slices, 1);
}


}  // namespace tree
