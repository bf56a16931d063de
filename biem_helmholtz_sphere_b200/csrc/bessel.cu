// K1: batch hyperspherical Bessel / Hankel functions, orders 0..n_max, complex128 output.
//
// One thread per argument runs the recurrences with its order sequence held in shared memory
// (order-major, thread-minor: conflict-free); the block then streams its contiguous
// [args, n_max+1] slab out with coalesced 16-byte stores.
// Replaces ultrasphere.shn1 / potential_coef radial factors (_biem.py:439,447,654-685,723-741,896-914).
#include "special.cuh"

struct StridedArr {
    double* base;
    int stride;
    __device__ __forceinline__ double& operator[](int n) const { return base[(size_t)n * stride]; }
};

// The order sequences sit in shared memory order-major with an ODD thread stride TS = T + 1: the recurrences (lane = argument,
// consecutive addresses) and the transposed read of the store phase (lane = order, stride TS) are both bank-conflict free.
__global__ void bessel_kernel(int d, int kind, int derivative, int n_max, int n_store, const double* __restrict__ x,
                              int64_t nx, cplx* __restrict__ out) {
    extern __shared__ __align__(16) double sm[];
    const int T = blockDim.x, TS = T + 1, L = n_max + 1;
    StridedArr aj{sm + threadIdx.x, TS};
    StridedArr ay{sm + (size_t)n_store * TS + threadIdx.x, TS};
    const bool want_j = kind != BHS_KIND_Y, want_y = kind != BHS_KIND_J;
    const double* s_re = want_j ? sm : sm + (size_t)n_store * TS;
    const double* s_im = sm + (size_t)n_store * TS;
    // (argument, order) of the first element this thread stores, and the step to its next one (element e += T)
    const int t_first = threadIdx.x / L, n_first = threadIdx.x % L, t_step = T / L, n_step = T % L;
    for (int64_t i0 = (int64_t)blockIdx.x * T; i0 < nx; i0 += (int64_t)gridDim.x * T) {
        int64_t i = i0 + threadIdx.x;
        double xv = (i < nx) ? x[i] : 1.0;
        radial_sequence(d, xv, n_max + 1, aj, ay, want_j, want_y);
        if (derivative) {
            for (int n = 0; n <= n_max; ++n) {
                if (want_j) aj[n] = radial_deriv(n, xv, aj[n], aj[n + 1]);
                if (want_y) ay[n] = radial_deriv(n, xv, ay[n], ay[n + 1]);
            }
        }
        __syncthreads();
        const int cnt = (int)((nx - i0 < T) ? nx - i0 : T);
        const int total = cnt * L;
        cplx* o = out + i0 * L;
        int t = t_first, n = n_first;
        for (int e = threadIdx.x; e < total; e += T) {
            const double re = s_re[n * TS + t];
            const double im = (kind == BHS_KIND_H1) ? s_im[n * TS + t] : 0.0;
            o[e] = cmake(re, im);
            t += t_step;
            n += n_step;
            if (n >= L) { n -= L; ++t; }
        }
        __syncthreads();
    }
}

extern "C" int bhs_bessel(int d, int kind, int derivative, int n_max, const double* d_x, int64_t nx, double* d_out,
                          void* stream) {
    if (d < 2 || kind < 0 || kind > 2 || n_max < 0 || nx < 0 || !d_x || !d_out) return BHS_ERR_INVALID;
    if (nx == 0) return BHS_OK;
    int shift = (d & 1) ? (d - 3) / 2 : d / 2 - 1;
    int n_store = n_max + 2 + shift + 1;
    int T = 128;
    while (T > 32 && (size_t)2 * n_store * (T + 1) * sizeof(double) > 160 * 1024) T >>= 1;
    size_t smem = (size_t)2 * n_store * (T + 1) * sizeof(double);
    if (smem > 200 * 1024) return BHS_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(bessel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = (nx + T - 1) / T;
    if (blocks > bhs_sm_count() * 8) blocks = bhs_sm_count() * 8;
    bessel_kernel<<<(unsigned)blocks, T, smem, (cudaStream_t)stream>>>(d, kind, derivative, n_max, n_store, d_x, nx,
                                                                       (cplx*)d_out);
    BHS_CHECK_LAUNCH();
    return BHS_OK;
}

// ---- FP64 peak micro-benchmarks ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double* sink) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll 4
        for (int u = 0; u < 4; ++u) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double* sink) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = 0.0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[2 * u]), "+d"(c[2 * u + 1])
                         : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) sink[0] = s;
}

// The larger PTX shapes (sm_90+).  On sm_100a ptxas lowers each of them to a sequence of DMMA.8x8x4 (2 / 4 / 8 per
// instruction; checked with cuobjdump, profiles/r02_fp64_mma_shapes.txt): there is one native FP64 tensor operation on this
// chip, so these measure the same pipe and only confirm that the wider shapes buy nothing.
template <int KK>
__global__ void __launch_bounds__(256) dmma16_peak_kernel(int iters, double* sink) {
    double a[KK / 2], b[KK / 4];
#pragma unroll
    for (int i = 0; i < KK / 2; ++i) a[i] = 1.0 + (threadIdx.x + i) * 1e-9;
#pragma unroll
    for (int i = 0; i < KK / 4; ++i) b[i] = 0.5 + i * 1e-9;
    double c[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[u][i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if constexpr (KK == 4)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3])
                             : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            else if constexpr (KK == 8)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            else
                asm volatile(
                    "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                    "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
                    : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3])
                    : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
                      "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += c[u][i];
    if (s == 123.456) sink[0] = s;
}

extern "C" int bhs_fp64_peak(int shape, int iters, double* tflops_out) {
    if (!tflops_out || iters <= 0 || shape < 0 || shape > 4) return BHS_ERR_INVALID;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* sink = nullptr;
    if (cudaMalloc(&sink, 8) != cudaSuccess) return BHS_ERR_ALLOC;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * 4, threads = 256;
    // flops per warp and loop iteration
    const double per_warp[5] = {32.0 * 32.0 * 2.0, 8.0 * 512.0, 4.0 * 1024.0, 4.0 * 2048.0, 4.0 * 4096.0};
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        switch (shape) {
            case 0: dfma_peak_kernel<<<blocks, threads>>>(iters, sink); break;
            case 1: dmma_peak_kernel<<<blocks, threads>>>(iters, sink); break;
            case 2: dmma16_peak_kernel<4><<<blocks, threads>>>(iters, sink); break;
            case 3: dmma16_peak_kernel<8><<<blocks, threads>>>(iters, sink); break;
            default: dmma16_peak_kernel<16><<<blocks, threads>>>(iters, sink); break;
        }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = (double)blocks * (threads / 32) * iters * per_warp[shape];
        double tf = flops / (ms * 1e-3) * 1e-12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    *tflops_out = best;
    return BHS_OK;
}

// ---- K1, complex argument ------------------------------------------------------------------------------------
// out[i, n] = z_n^{(d)}(x_i) (kind J or H1) or its derivative, x_i = xr[i] + i xi[i] != 0.  One thread per argument, order
// sequences in shared memory (order-major); used for the complex-wavenumber point source and by the parity tests.
struct StridedArrZ2 {
    cplx* base;
    int stride;
    __device__ __forceinline__ cplx& operator[](int n) const { return base[(size_t)n * stride]; }
};

__global__ void bessel_z_kernel(int d, int kind, int derivative, int n_max, int n_store, const double* __restrict__ xr,
                                const double* __restrict__ xi, int64_t nx, cplx* __restrict__ out) {
    extern __shared__ __align__(16) cplx smz[];
    const int T = blockDim.x;
    StridedArrZ2 arr{smz + threadIdx.x, T};
    for (int64_t i = (int64_t)blockIdx.x * T + threadIdx.x; i < nx; i += (int64_t)gridDim.x * T) {
        const cplx z = cmake(xr[i], xi[i]);
        const bool want_j = kind == BHS_KIND_J;
        radial_sequence_z(d, z, n_max + 1, arr, arr, want_j, !want_j);
        const cplx iz = crecip(z);
        for (int n = 0; n <= n_max; ++n)
            out[i * (n_max + 1) + n] = derivative ? radial_deriv_z(n, iz, arr[n], arr[n + 1]) : arr[n];
    }
}

extern "C" int bhs_bessel_z(int d, int kind, int derivative, int n_max, const double* d_x_re, const double* d_x_im,
                            int64_t nx, double* d_out, void* stream) {
    if (d < 2 || (kind != BHS_KIND_J && kind != BHS_KIND_H1) || n_max < 0 || nx < 0) return BHS_ERR_INVALID;
    if (nx == 0) return BHS_OK;
    if (!d_x_re || !d_x_im || !d_out) return BHS_ERR_INVALID;
    const int shift = (d & 1) ? (d - 3) / 2 : d / 2 - 1;
    const int n_store = n_max + 2 + shift + 1;
    int T = 64;
    while (T > 32 && (size_t)n_store * T * sizeof(cplx) > 160 * 1024) T >>= 1;
    const size_t smem = (size_t)n_store * T * sizeof(cplx);
    if (smem > 200 * 1024) return BHS_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(bessel_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = (nx + T - 1) / T;
    if (blocks > bhs_sm_count() * 8) blocks = bhs_sm_count() * 8;
    bessel_z_kernel<<<(unsigned)blocks, T, smem, (cudaStream_t)stream>>>(d, kind, derivative, n_max, n_store, d_x_re, d_x_im,
                                                                         nx, (cplx*)d_out);
    BHS_CHECK_LAUNCH();
    return BHS_OK;
}
