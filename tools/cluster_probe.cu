// Micro-benchmark (design aid, not product code): cost of one "everybody tells everybody" step inside a thread-block
// cluster, the primitive a cluster-resident LU panel needs once per pivot column.
//   mode 0: plain DSMEM stores + barrier.cluster (arrive.release / wait.acquire)
//   mode 1: st.async + remote mbarrier complete_tx (no cluster barrier)
// usage: cluster_probe [cluster_size] [threads] [msg_bytes]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

#define MAXC 16
#define MAXMSG 528

__global__ void probe_kernel(int mode, int steps, int msg_bytes, long long* out, double* sink) {
    __shared__ __align__(16) unsigned char slots[2][MAXC][MAXMSG];
    __shared__ __align__(8) uint64_t bar[2];
    const uint32_t rank = cluster_rank(), csz = cluster_size();
    const int tid = threadIdx.x;
    const int chunks = msg_bytes / 16;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();
    double acc = 0.0;
    long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
        const int buf = s & 1;
        if (mode == 1 && tid == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[buf])), "r"(csz * msg_bytes) : "memory");
        // thread t sends chunk (t % chunks) to peer (t / chunks)
        if (tid < (int)csz * chunks) {
            const uint32_t peer = tid / chunks, ch = tid % chunks;
            const uint32_t dst = mapa(smem_u32(&slots[buf][rank][ch * 16]), peer);
            const unsigned long long v0 = (unsigned long long)s * 131 + rank, v1 = ch;
            if (mode == 0) {
                asm volatile("st.shared::cluster.v2.u64 [%0], {%1, %2};" ::"r"(dst), "l"(v0), "l"(v1) : "memory");
            } else {
                const uint32_t rbar = mapa(smem_u32(&bar[buf]), peer);
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(dst), "l"(v0), "l"(v1), "r"(rbar) : "memory");
            }
        }
        if (mode == 0) {
            cluster_sync_all();
        } else {
            const uint32_t parity = (s >> 1) & 1;
            asm volatile(
                "{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DN;\nbra WL;\nDN:\n}\n" ::"r"(smem_u32(&bar[buf])), "r"(parity) : "memory");
        }
        // consume: everybody reads every slot's first word (as the arg-max over the cluster would)
        unsigned long long m = 0;
        for (uint32_t p = 0; p < csz; ++p) {
            unsigned long long v = *reinterpret_cast<volatile unsigned long long*>(&slots[buf][p][0]);
            m = v > m ? v : m;
        }
        acc += (double)m;
        __syncthreads();
    }
    long long t1 = clock64();
    cluster_sync_all();
    if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 1.2345) sink[0] = acc;
}

int main(int argc, char** argv) {
    int csz = argc > 1 ? atoi(argv[1]) : 16, threads = argc > 2 ? atoi(argv[2]) : 1024, msg = argc > 3 ? atoi(argv[3]) : 16;
    long long* d_out; double* d_sink;
    cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 8);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(csz); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = 0;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            const int steps = 2000;
            cudaError_t e = cudaLaunchKernelEx(&cfg, probe_kernel, mode, steps, msg, d_out, d_sink);
            cudaError_t e2 = cudaDeviceSynchronize();
            long long cyc = 0; cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
            printf("cluster %d x %d threads, msg %d B, mode %d (%s): launch %s / sync %s, %.1f cycles per step\n", csz, threads, msg, mode,
                   mode ? "st.async+mbarrier" : "st.cluster+barrier.cluster", cudaGetErrorString(e), cudaGetErrorString(e2), (double)cyc / steps);
        }
    }
    return 0;
}
