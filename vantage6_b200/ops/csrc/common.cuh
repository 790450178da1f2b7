// Shared device helpers for the vantage6_b200 sm_100a kernels.
//
// Everything here is plain CUDA C++ + inline PTX (no torch / CUTLASS headers) so that
// each .cu compiles in seconds with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define V6_MAX_PEERS 8
#define V6_DEVINL __device__ __forceinline__

// ----------------------------------------------------------------------------------------
// error handling for launchers (host side). Launchers return cudaError_t as int.
// ----------------------------------------------------------------------------------------
#define V6_CHECK_LAUNCH() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; } while (0)

// ----------------------------------------------------------------------------------------
// warp / block reductions
// ----------------------------------------------------------------------------------------
V6_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
V6_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum for blockDim.x <= 1024; `red` is >= 32 floats of shared memory.
V6_DEVINL float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();                  // protect `red` re-use between consecutive calls
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
    if (wid == 0) r = warp_sum(r);
    if (threadIdx.x == 0) red[0] = r;
    __syncthreads();
    return red[0];
}

// ----------------------------------------------------------------------------------------
// vectorised global access (16 B) with streaming hints
// ----------------------------------------------------------------------------------------
V6_DEVINL float4 ldg_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// plain (coherent) 16 B load: used for peer / symmetric memory that other GPUs write.
V6_DEVINL float4 ld_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
// volatile/relaxed system-scope load that always goes to the owner (never served from a
// stale L1 line): peer data that a remote GPU rewrites every round.
V6_DEVINL float4 ld_sys_f4(const float4* p) {
    float4 r;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
V6_DEVINL uint4 ld_sys_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
V6_DEVINL void st_f4(float4* p, float4 v) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
V6_DEVINL void st_stream_f4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
V6_DEVINL void st_u4(uint4* p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ----------------------------------------------------------------------------------------
// NVLS multicast (multimem.*) -- addresses are multicast VAs bound with cuMulticastBindMem
// ----------------------------------------------------------------------------------------
V6_DEVINL float4 multimem_ld_reduce_add_f4(const float4* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
    return r;
}
// 8 x bf16 summed in the switch with fp32 accumulation
V6_DEVINL uint4 multimem_ld_reduce_add_bf16x8(const uint4* mc) {
    uint4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc) : "memory");
    return r;
}
V6_DEVINL void multimem_st_f4(float4* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
V6_DEVINL void multimem_st_u4(uint4* mc, uint4 v) {
    // multimem.st has no .u32 vector form; bit-cast through f32
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
                    "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
}

// ----------------------------------------------------------------------------------------
// system-scope signalling (cross-GPU flags in symmetric memory)
// ----------------------------------------------------------------------------------------
V6_DEVINL void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
V6_DEVINL uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
V6_DEVINL uint32_t ld_relaxed_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
V6_DEVINL void red_release_sys_add_u32(uint32_t* p, uint32_t v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
V6_DEVINL void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

// Bounded spin: returns false on timeout (so a dead peer cannot hang the GPU).
// `abort_flag` (may be null) is a host-mapped / device word that aborts the wait when != 0.
V6_DEVINL bool spin_wait_ge(const uint32_t* flag, uint32_t target, long long timeout_cycles,
                            const uint32_t* abort_flag) {
    const long long t0 = clock64();
    uint32_t it = 0;
    while (true) {
        const uint32_t v = ld_acquire_sys_u32(flag);
        if ((int32_t)(v - target) >= 0) return true;
        if ((++it & 0x3ff) == 0) {
            if (clock64() - t0 > timeout_cycles) return false;
            if (abort_flag && ld_relaxed_sys_u32(abort_flag) != 0) return false;
        }
        __nanosleep(32);
    }
}

// ----------------------------------------------------------------------------------------
// bf16 pack helpers
// ----------------------------------------------------------------------------------------
V6_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
V6_DEVINL float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

// ----------------------------------------------------------------------------------------
// mbarrier / TMA / tcgen05 PTX wrappers (sm_100a)
// ----------------------------------------------------------------------------------------
V6_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

V6_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
V6_DEVINL void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
V6_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
V6_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
V6_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
V6_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: traps (kernel error, not a hang) if a pipeline deadlocks during bring-up.
V6_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0xfff) == 0 && clock64() - t0 > 4000000000LL) { asm volatile("trap;"); }   // ~2 s
    }
}

// 2-D tiled TMA load: global (via CUtensorMap) -> shared, completion on mbarrier.
V6_DEVINL void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
V6_DEVINL void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
V6_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
V6_DEVINL void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
V6_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
V6_DEVINL void tma_store_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }   // <= 1 store still reading
V6_DEVINL void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(tmap) : "memory");
}

V6_DEVINL bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(pred));
    return pred != 0;
}

// --- TMEM -------------------------------------------------------------------------------
template <int kCols>
V6_DEVINL void tmem_alloc(uint32_t* smem_result) {   // one full warp, .sync.aligned
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(smem_result)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
V6_DEVINL void tmem_dealloc(uint32_t taddr) {         // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(kCols) : "memory");
}
V6_DEVINL void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
V6_DEVINL void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16/fp16 inputs, fp32 accum.
V6_DEVINL void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                            uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// commit all prior tcgen05.mma of this thread; arrive (count 1) on the mbarrier when they finish
V6_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}
// TMEM -> registers: 32 lanes x 32 columns of fp32 (each thread: its lane, 32 consecutive cols)
V6_DEVINL void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
V6_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major bf16 tile stored by TMA with SWIZZLE_128B:
// rows of 64 bf16 (128 B), 8-row groups 1024 B apart (SBO=1024), LBO ignored for swizzled
// K-major, descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
V6_DEVINL uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // [0,14)  start address >> 4
    d |= (uint64_t)1 << 16;                              // [16,30) leading byte offset (unused) = 1
    d |= (uint64_t)(1024 >> 4) << 32;                    // [32,46) stride byte offset = 1024 B
    d |= (uint64_t)1 << 46;                              // [46,48) descriptor version = 1
    d |= (uint64_t)2 << 61;                              // [61,64) swizzle mode: 128B
    return d;
}
// Instruction descriptor for kind::f16: D=f32, A=B=bf16, both K-major, no negate/sparsity.
V6_DEVINL constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4)                      // [4,6)   D format: 1 = f32
         | (1u << 7)                      // [7,10)  A format: 1 = bf16
         | (1u << 10)                     // [10,13) B format: 1 = bf16
         | (0u << 15) | (0u << 16)        // A, B major: 0 = K-major
         | ((uint32_t)(N >> 3) << 17)     // [17,23) N >> 3
         | ((uint32_t)(M >> 4) << 24);    // [24,29) M >> 4
}

// --- TMEM stores and TMEM-sourced A operand (attention kernels) ------------------------------
V6_DEVINL void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        :: "r"(taddr),
           "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
           "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
           "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
           "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
V6_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]
V6_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

