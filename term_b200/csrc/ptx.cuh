// sm_100a inline-PTX helpers: mbarrier + TMA bulk (1-D) global->shared copies, streaming loads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// TMA bulk copy (non-tensor): global -> shared::cta, completion counted in bytes on an mbarrier.
// dst, src 16-byte aligned; bytes a multiple of 16. SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Bulk prefetch of a global range into L2 (no destination, no completion to wait for). 16-byte aligned, multiple of 16.
__device__ __forceinline__ void l2_prefetch(const void* gmem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}

// Streaming (read-once) 128-bit global load that does not allocate in L1.
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
    return __shfl_xor_sync(0xffffffffu, v, m);
}
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    return __shfl_xor_sync(0xffffffffu, (unsigned long long)v, m);
}

// up to 8 bytes starting at bytes[p] (len <= 8), little endian, without reading past p + len rounded up to 8
__device__ __forceinline__ uint64_t load_upto8(const uint8_t* bytes, int64_t p, int len) {
    const int64_t a = p & ~(int64_t)7;
    const int sh = (int)(p - a) * 8;
    uint64_t w = __ldg(reinterpret_cast<const unsigned long long*>(bytes + a)) >> sh;
    if (sh + len * 8 > 64) w |= __ldg(reinterpret_cast<const unsigned long long*>(bytes + a + 8)) << (64 - sh);
    return len >= 8 ? w : (w & ((1ull << (len * 8)) - 1ull));
}


}  // namespace tg
