// Shared helpers for libbhs (sm_100a).  Complex128 is an interleaved double2 (x = re, y = im).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/bhs.h"

// every kernel launch of the library is followed by one of the *_CHECK macros, which also count it
// (bhs_launch_count: evidence for bench.py's gpu_launches)
extern unsigned long long g_bhs_launches;
#define BHS_COUNT_LAUNCH() __sync_fetch_and_add(&g_bhs_launches, 1ULL)
#define BHS_CHECK_LAUNCH()                         \
    do {                                           \
        BHS_COUNT_LAUNCH();                        \
        cudaError_t e__ = cudaGetLastError();      \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

// number of SMs of the current device (queried once per device; grids of the persistent / capped kernels are sized from it)
int bhs_sm_count();

typedef double2 cplx;

__host__ __device__ __forceinline__ cplx cmake(double re, double im) { return make_double2(re, im); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return cmake(a.x * s, a.y * s); }
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) {  // a*b + c
    return cmake(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    // Smith's algorithm
    if (fabs(b.x) >= fabs(b.y)) {
        double r = b.y / b.x, den = b.x + b.y * r;
        return cmake((a.x + a.y * r) / den, (a.y - a.x * r) / den);
    } else {
        double r = b.x / b.y, den = b.x * r + b.y;
        return cmake((a.x * r + a.y) / den, (a.y * r - a.x) / den);
    }
}
__device__ __forceinline__ cplx crecip(cplx b) { return cdiv(cmake(1.0, 0.0), b); }
// i^k for any integer k
__host__ __device__ __forceinline__ cplx cipow(int k) {
    switch (k & 3) {
        case 0: return cmake(1.0, 0.0);
        case 1: return cmake(0.0, 1.0);
        case 2: return cmake(-1.0, 0.0);
        default: return cmake(0.0, -1.0);
    }
}
__device__ __forceinline__ cplx cmul_ipow(cplx a, int k) { return cmul(a, cipow(k)); }

// ---- mbarrier + 1-D bulk async copy (TMA, SASS: UBLKCP) ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes must be a multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// make generic-proxy smem writes visible to the async proxy / order reuse of a TMA buffer
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
