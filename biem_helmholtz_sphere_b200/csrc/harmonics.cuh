// Device-side harmonic evaluation shared by the harmonics, assembly and generic field kernels.
#pragma once
#include "plan.h"

struct HarmTables {
    int d, L2;
    const double* all;  // [n_bnodes][L2]
    const double* c1;   // [n_bnodes][L2][L2]
    const double* c2;
};

static inline HarmTables harm_tables_of(const bhs_plan* p) {
    HarmTables t;
    t.d = p->d;
    t.L2 = p->L2;
    t.all = p->d_node_all;
    t.c1 = p->d_node_c1;
    t.c2 = p->d_node_c2;
    return t;
}

__host__ __device__ inline size_t harm_smem_bytes_per_warp(int d, int Lb) {
    size_t b = (size_t)(d - 2) * Lb * Lb * sizeof(double) + (size_t)Lb * sizeof(cplx);
    return (b + 15) & ~(size_t)15;
}
// per-warp scratch layout: E (complex, 16-byte aligned) first, then the node tables F
__device__ __forceinline__ void harm_smem_carve(unsigned char* base, int Lb, double*& F, cplx*& E) {
    E = reinterpret_cast<cplx*>(base);
    F = reinterpret_cast<double*>(base + (size_t)Lb * sizeof(cplx));
}

__device__ __forceinline__ double powi_d(double b, int e) {
    double r = 1.0;
    while (e) {
        if (e & 1) r *= b;
        b *= b;
        e >>= 1;
    }
    return r;
}

// Builds, cooperatively over one warp, F[i][n*Lb + l] (node functions, l <= n < Lb) for every b-node
// and E[m] = e^{i m phi}/sqrt(2 pi), m = 0..Lb-1, for the direction of the cartesian vector x[0..d).
__device__ inline void warp_harmonic_tables(const HarmTables& tb, int Lb, const double* x, double* F, cplx* E,
                                            int lane) {
    const int d = tb.d;
    // tails: tail[i] = sqrt(x_i^2 + ... + x_{d-1}^2)
    double tail[BHS_MAX_NODES + 3];
    double acc = 0.0;
    tail[d] = 0.0;
    for (int i = d - 1; i >= 0; --i) {
        acc += x[i] * x[i];
        tail[i] = sqrt(acc);
    }
    for (int i = 0; i < d - 2; ++i) {
        double ct = 1.0, st = 0.0;
        if (tail[i] > 0.0) {
            ct = x[i] / tail[i];
            st = tail[i + 1] / tail[i];
        }
        double* Fi = F + (size_t)i * Lb * Lb;
        const double* c1 = tb.c1 + (size_t)i * tb.L2 * tb.L2;
        const double* c2 = tb.c2 + (size_t)i * tb.L2 * tb.L2;
        for (int l = lane; l < Lb; l += 32) {
            double f0 = tb.all[(size_t)i * tb.L2 + l] * powi_d(st, l);
            Fi[(size_t)l * Lb + l] = f0;
            double fm2 = 0.0, fm1 = f0;
            for (int n = l + 1; n < Lb; ++n) {
                double f = c1[(size_t)n * tb.L2 + l] * ct * fm1 - c2[(size_t)n * tb.L2 + l] * fm2;
                Fi[(size_t)n * Lb + l] = f;
                fm2 = fm1;
                fm1 = f;
            }
        }
    }
    // azimuth
    double cp = 1.0, sp = 0.0;
    if (tail[d - 2] > 0.0) {
        cp = x[d - 2] / tail[d - 2];
        sp = x[d - 1] / tail[d - 2];
    }
    const double inv_s2pi = 0.39894228040143267794;
    for (int m = lane; m < Lb; m += 32) {
        // (cp + i sp)^m by binary powering
        cplx r = cmake(1.0, 0.0), b = cmake(cp, sp);
        int e = m;
        while (e) {
            if (e & 1) r = cmul(r, b);
            b = cmul(b, b);
            e >>= 1;
        }
        E[m] = cscale(r, inv_s2pi);
    }
}

// idx_row: (n_0, ..., n_{d-3}, m) of one flattened harmonic
__device__ __forceinline__ cplx harmonic_from_tables(const HarmTables& tb, int Lb, const int32_t* idx_row,
                                                     const double* F, const cplx* E) {
    const int s = tb.d - 1;
    int m = idx_row[s - 1];
    int am = m < 0 ? -m : m;
    double prod = 1.0;
    for (int i = 0; i < s - 1; ++i) {
        int n = idx_row[i];
        int l = (i + 1 < s - 1) ? idx_row[i + 1] : am;
        prod *= F[((size_t)i * Lb + n) * Lb + l];
    }
    cplx e = E[am];
    if (m < 0) e.y = -e.y;
    return cscale(e, prod);
}
