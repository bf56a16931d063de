// Per-ball radial tables (radial.cu).
#pragma once
#include "common.cuh"

// rad[(s*B + b)*L + n] = (j_n, j_n', y_n, y_n')(k_s * rho_b) for the d-dimensional hyperspherical functions
// (d_k == nullptr: every system uses k_scalar)
// d_scratch: ball_radial_scratch_bytes(d, L) bytes of caller-provided global memory (may be null when that is 0)
size_t ball_radial_scratch_bytes(int d, int L);
int launch_ball_radial(int d, int L, int B, int nsys, const double* d_radii, const double* d_k, double k_scalar,
                       double4* d_out, double* d_scratch, cudaStream_t st);

// SD_n(rho) = D - i eta S = rho^{d-1} k^{d-2} (eta j_n + i k j_n')      (_biem.py:742)
__device__ __forceinline__ cplx sd_coef(int d, double k, double eta, double rho, double jn, double jd) {
    double pr = 1.0;
    for (int i = 0; i < d - 1; ++i) pr *= rho;
    double pk = 1.0;
    for (int i = 0; i < d - 2; ++i) pk *= k;
    return cmake(pr * pk * eta * jn, pr * pk * k * jd);
}

// Complex wavenumber k = kr + i ki (3-D only): radz[((s*B + b)*L + n)*4 + q] = (j_n, j_n', h_n, h_n')[q](k_s rho_b), complex.
// d_kr / d_ki == nullptr: every system uses the scalars.
int launch_ball_radial_z(int d, int L, int B, int nsys, const double* d_radii, const double* d_kr, const double* d_ki,
                         double kr_s, double ki_s, cplx* d_out, cudaStream_t st);

// SD_n(rho) for complex k: rho^{d-1} k^{d-2} (eta j_n + i k j_n')
__device__ __forceinline__ cplx sd_coef_z(int d, cplx k, double eta, double rho, cplx jn, cplx jd) {
    double pr = 1.0;
    for (int i = 0; i < d - 1; ++i) pr *= rho;
    cplx pk = cmake(1.0, 0.0);
    for (int i = 0; i < d - 2; ++i) pk = cmul(pk, k);
    const cplx ikjd = cmul(cmake(-k.y, k.x), jd);  // i k j'
    return cmul(cscale(pk, pr), cadd(cscale(jn, eta), ikjd));
}
