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

// Order sequences live in shared memory (order-major, thread-minor); for very high orders (2-D runs with n_end in the
// thousands) they do not fit and the caller passes a global scratch of 2 * n_store * (gridDim.x * blockDim.x) doubles.
__global__ void ball_radial_kernel(int d, int L, int n_store, int B, int nsys, const double* __restrict__ radii,
                                   const double* __restrict__ ks, double k_scalar, double4* __restrict__ out,
                                   double* __restrict__ scratch) {
    extern __shared__ __align__(16) double sm[];
    const int T = scratch ? gridDim.x * blockDim.x : blockDim.x;
    const int t = scratch ? blockIdx.x * blockDim.x + threadIdx.x : threadIdx.x;
    double* base = scratch ? scratch : sm;
    StridedArr2 aj{base + t, T};
    StridedArr2 ay{base + (size_t)n_store * T + t, T};
    int64_t total = (int64_t)B * nsys;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int s = (int)(i / B), b = (int)(i % B);
        double x = (ks ? ks[s] : k_scalar) * radii[b];
        radial_sequence(d, x, L, aj, ay, true, true);
        for (int n = 0; n < L; ++n) {
            double jn = aj[n], yn = ay[n];
            out[i * L + n] = make_double4(jn, radial_deriv(n, x, jn, aj[n + 1]), yn, radial_deriv(n, x, yn, ay[n + 1]));
        }
    }
}

// Order sequences of very high orders (2-D, n_end in the hundreds or thousands) do not fit in shared memory: the kernel
// then keeps them in a global scratch that the CALLER provides (the C layer never allocates): this many bytes, 0 when the
// shared-memory path applies.
static void ball_radial_shape(int d, int L, int& n_store, int& T, size_t& smem) {
    int shift = (d & 1) ? (d - 3) / 2 : d / 2 - 1;
    n_store = L + 2 + shift + 1;
    T = 64;
    while (T > 32 && (size_t)2 * n_store * T * sizeof(double) > 160 * 1024) T >>= 1;
    smem = (size_t)2 * n_store * T * sizeof(double);
}
size_t ball_radial_scratch_bytes(int d, int L) {
    int n_store, T;
    size_t smem;
    ball_radial_shape(d, L, n_store, T, smem);
    return smem > 200 * 1024 ? (size_t)2 * n_store * 32 * T * sizeof(double) : 0;
}

int launch_ball_radial(int d, int L, int B, int nsys, const double* d_radii, const double* d_k, double k_scalar,
                       double4* d_out, double* d_scratch, cudaStream_t st) {
    int n_store, T;
    size_t smem;
    ball_radial_shape(d, L, n_store, T, smem);
    int64_t total = (int64_t)B * nsys;
    int64_t blocks = (total + T - 1) / T;
    if (blocks > bhs_sm_count() * 8) blocks = bhs_sm_count() * 8;
    if (smem > 200 * 1024) {
        // high orders: sequences in the caller's global scratch (rare path)
        if (!d_scratch) return BHS_ERR_INVALID;
        if (blocks > 32) blocks = 32;
        ball_radial_kernel<<<(unsigned)blocks, T, 0, st>>>(d, L, n_store, B, nsys, d_radii, d_k, k_scalar, d_out, d_scratch);
        BHS_CHECK_LAUNCH();
        return BHS_OK;
    }
    cudaFuncSetAttribute(ball_radial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ball_radial_kernel<<<(unsigned)blocks, T, smem, st>>>(d, L, n_store, B, nsys, d_radii, d_k, k_scalar, d_out, nullptr);
    BHS_CHECK_LAUNCH();
    return BHS_OK;
}

// ---- complex wavenumber -------------------------------------------------------------------
// radz[((s*B + b)*L + n)*4 + {0,1,2,3}] = ( j_n, j_n', h_n, h_n' )(k_s rho_b), all complex, k_s = kr[s] + i ki[s]
struct StridedArrZ {
    cplx* base;
    int stride;
    __device__ __forceinline__ cplx& operator[](int n) const { return base[(size_t)n * stride]; }
};

__global__ void ball_radial_z_kernel(int d, int L, int n_store, int B, int nsys, const double* __restrict__ radii,
                                     const double* __restrict__ kr, const double* __restrict__ ki, double kr_s, double ki_s,
                                     cplx* __restrict__ out) {
    extern __shared__ __align__(16) cplx smz[];
    const int T = blockDim.x;
    StridedArrZ aj{smz + threadIdx.x, T};
    StridedArrZ ah{smz + (size_t)n_store * T + threadIdx.x, T};
    int64_t total = (int64_t)B * nsys;
    for (int64_t i = (int64_t)blockIdx.x * T + threadIdx.x; i < total; i += (int64_t)gridDim.x * T) {
        int s = (int)(i / B), b = (int)(i % B);
        const double rho = radii[b];
        const cplx z = cmake((kr ? kr[s] : kr_s) * rho, (ki ? ki[s] : ki_s) * rho);
        radial_sequence_z(d, z, L, aj, ah, true, true);
        const cplx iz = crecip(z);
        for (int n = 0; n < L; ++n) {
            const cplx jn = aj[n], hn = ah[n];
            cplx* o = out + (i * L + n) * 4;
            o[0] = jn;
            o[1] = radial_deriv_z(n, iz, jn, aj[n + 1]);
            o[2] = hn;
            o[3] = radial_deriv_z(n, iz, hn, ah[n + 1]);
        }
    }
}

int launch_ball_radial_z(int d, int L, int B, int nsys, const double* d_radii, const double* d_kr, const double* d_ki,
                         double kr_s, double ki_s, cplx* d_out, cudaStream_t st) {
    const int shift = (d & 1) ? (d - 3) / 2 : d / 2 - 1;
    int n_store = L + 2 + shift;
    int T = 64;
    while (T > 32 && (size_t)2 * n_store * T * sizeof(cplx) > 160 * 1024) T >>= 1;
    size_t smem = (size_t)2 * n_store * T * sizeof(cplx);
    if (smem > 200 * 1024) return BHS_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(ball_radial_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t total = (int64_t)B * nsys;
    int64_t blocks = (total + T - 1) / T;
    if (blocks > bhs_sm_count() * 8) blocks = bhs_sm_count() * 8;
    ball_radial_z_kernel<<<(unsigned)blocks, T, smem, st>>>(d, L, n_store, B, nsys, d_radii, d_kr, d_ki, kr_s, ki_s, d_out);
    BHS_CHECK_LAUNCH();
    return BHS_OK;
}
