// Per-ball radial tables shared by assembly, the single-sphere shortcut and field evaluation:
//   rad[s][b][n] = ( j_n, j_n', y_n, y_n' )(k_s rho_b),   n = 0..L-1          (hyperspherical, d-dim)
// One thread per (system, ball); order sequences live in shared memory (order-major).
#include "radial.cuh"

#include "special.cuh"

struct StridedArr2 {
    double* base;
    int stride;
    __device__ __forceinline__ double& operator[](int n) const { return base[(size_t)n * stride]; }
};

__global__ void ball_radial_kernel(int d, int L, int n_store, int B, int nsys, const double* __restrict__ radii,
                                   const double* __restrict__ ks, double k_scalar, double4* __restrict__ out) {
    extern __shared__ __align__(16) double sm[];
    const int T = blockDim.x;
    StridedArr2 aj{sm + threadIdx.x, T};
    StridedArr2 ay{sm + (size_t)n_store * T + threadIdx.x, T};
    int64_t total = (int64_t)B * nsys;
    for (int64_t i = (int64_t)blockIdx.x * T + threadIdx.x; i < total; i += (int64_t)gridDim.x * T) {
        int s = (int)(i / B), b = (int)(i % B);
        double x = (ks ? ks[s] : k_scalar) * radii[b];
        radial_sequence(d, x, L, aj, ay, true, true);
        for (int n = 0; n < L; ++n) {
            double jn = aj[n], yn = ay[n];
            out[i * L + n] = make_double4(jn, radial_deriv(n, x, jn, aj[n + 1]), yn, radial_deriv(n, x, yn, ay[n + 1]));
        }
    }
}

int launch_ball_radial(int d, int L, int B, int nsys, const double* d_radii, const double* d_k, double k_scalar,
                       double4* d_out, cudaStream_t st) {
    int shift = (d & 1) ? (d - 3) / 2 : d / 2 - 1;
    int n_store = L + 2 + shift + 1;
    int T = 64;
    while (T > 32 && (size_t)2 * n_store * T * sizeof(double) > 160 * 1024) T >>= 1;
    size_t smem = (size_t)2 * n_store * T * sizeof(double);
    if (smem > 200 * 1024) return BHS_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(ball_radial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t total = (int64_t)B * nsys;
    int64_t blocks = (total + T - 1) / T;
    if (blocks > 148 * 8) blocks = 148 * 8;
    ball_radial_kernel<<<(unsigned)blocks, T, smem, st>>>(d, L, n_store, B, nsys, d_radii, d_k, k_scalar, d_out);
    BHS_CHECK_LAUNCH();
    return BHS_OK;
}
