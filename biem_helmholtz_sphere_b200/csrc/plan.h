// bhs_plan: k-independent tables of one (d, n_end) pair.  Built on the host in long double
// (plan.cu), uploaded once, then read-only for every kernel.
#pragma once
#include <stdint.h>

#include <vector>

#include "common.cuh"

#define BHS_MAX_NODES 6  // chain depth supported by the struct (d <= 8)

// coupling-table tile geometry used by bhs_assemble (rows = harmonic h of the ROW ball b,
// columns = harmonic h' of the COLUMN ball b')
#define BHS_TILE_R 4
#define BHS_TILE_C 64
#define BHS_TILE_E (BHS_TILE_R * BHS_TILE_C)

struct bhs_tile_hdr {
    int32_t nt;        // term layers in this tile
    int32_t sy_lo;     // first SY index referenced (multiple of 1)
    int32_t sy_cnt;    // number of SY entries referenced (idx are relative to sy_lo)
    int32_t pad;
    int64_t coef_off;  // offset (doubles) into coef array : layout [nt][TILE_R][TILE_C]
    int64_t idx_off;   // offset (uint16)  into idx array  : layout [nt][TILE_R][TILE_C]
};

struct bhs_plan {
    int tree;  // BHS_TREE_CHAIN, or BHS_TREE_HOPF: right-hand-side quadrature of the 'caa' tree, no coupling table
    int d, s_ndim, n_end, L2;
    int H, H2, Q;
    int n_bnodes;  // d - 2
    // host copies
    std::vector<int32_t> h_idx;   // [H][s_ndim]
    std::vector<int32_t> h_idx2;  // [H2][s_ndim]
    std::vector<double> h_qdirs;  // [d][Q]
    std::vector<double> h_qw;     // [Q]
    int64_t coupling_terms, coupling_bytes;
    int tiles_r, tiles_c, max_nt;
    int max_sy_cnt;  // largest S window (hd.sy_cnt) referenced by one tile: sizes the shared-memory S buffer
    // device tables
    int32_t* d_idx;    // [H][s_ndim]
    int32_t* d_idx2;   // [H2][s_ndim]
    int32_t* d_deg;    // [H]
    int32_t* d_deg2;   // [H2]
    // b-node recurrence tables, band L2: node i (i < n_bnodes) has desc = d-2-i descendants
    double* d_node_all;  // [n_bnodes][L2]       A_{l,l}
    double* d_node_c1;   // [n_bnodes][L2][L2]   (n, l): f_n = c1*x*f_{n-1} - c2*f_{n-2}
    double* d_node_c2;   // [n_bnodes][L2][L2]
    double* d_qdirs;     // [d][Q]
    double* d_qw;        // [Q]
    cplx* d_WY;          // [Q][H]   w_q * conj(Y_h(y_q))   (filled by the harmonics kernel)
    // coupling table
    bhs_tile_hdr* d_tiles;  // [tiles_r * tiles_c]
    std::vector<bhs_tile_hdr> h_tiles;
    double* d_coef;
    uint16_t* d_cidx;
    // 3-D fast field-evaluation tables (monic Legendre recurrence), m-major order
    double* d_us_beta;  // [L(L+1)/2]  beta_{n,m} for the monic recurrence
    double* d_us_norm;  // [L(L+1)/2]  normalisation folded into the coefficients (incl. 1/sqrt(2pi))
    // 3-D planar field evaluation (n_end <= 32): rotation of the harmonic coefficients into the frame whose polar axis is the
    // normal x2 of the plane, degree blocks [(2n+1) x (2n+1)] at offset (4 n^3 - n) / 3: c'_{n,m'} = sum_m rot[n][m'][m] c_{n,m};
    // and the equator values K_h = Y_h(theta = pi/2, phi = 0) (real; zero when n + |m| is odd)
    cplx* d_us_rot;
    double* d_us_K;
};
