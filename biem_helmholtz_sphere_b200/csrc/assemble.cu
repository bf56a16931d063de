// K3 (right-hand side) and K4 (system assembly).
//
// A[(b,h),(b',h')] = SD_{n'}(rho_b') * ( b==b' ? delta_{hh'} (alpha_b h_n + beta_b k h_n')(k rho_b)
//                                                : (S|R)_{h',h}(c_b - c_b') (alpha_b j_n + beta_b k j_n')(k rho_b) )
// (S|R)_{h',h}(t) = sum_terms coef(h,h',t) * i^{n+n''-n'} h_{n''}(k|t|) Y_{h''}(t^)           
//
// The real coupling coefficients `coef` (Gaunt-type triple integrals, k-independent) live in the plan as
// 8x64 tiles in ELL form.  A CTA owns one tile: it stages the tile's coefficient/index layers in shared
// memory ONCE with 1-D TMA bulk copies and then loops over ball pairs, for each pair building the
// translation vector S_{h''}(t) = i^{n''} h_{n''}(k|t|) Y_{h''}(t^) in shared memory and contracting it
// against the resident tile (2 DFMA per term).  Output rows are written as 1 KB contiguous segments.
// The i^{n}, (-i)^{n'} phase split and the radial row/column factors are folded into two small vectors.
#include "harmonics.cuh"
#include "prof.h"
#include "radial.cuh"
#include "special.cuh"

#define ASM_THREADS 256
#define ASM_EPT (BHS_TILE_E / ASM_THREADS)  // tile entries per thread
#define ASM_LAYER_COEF (BHS_TILE_E * 8)
#define ASM_LAYER_IDX (BHS_TILE_E * 2)

// ---- pre-kernels --------------------------------------------------------------------------------------
// translation vectors t = c_b - c_b' for every ordered pair (b, b'):  tv[i][b*B+b'], dist[b*B+b']
__global__ void pair_vectors_kernel(int d, int B, const double* __restrict__ centers, double* __restrict__ tv,
                                    double* __restrict__ dist) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t np = (int64_t)B * B;
    if (i >= np) return;
    int b = (int)(i / B), bp = (int)(i % B);
    double r2 = 0.0;
    for (int a = 0; a < d; ++a) {
        double t = centers[(int64_t)b * d + a] - centers[(int64_t)bp * d + a];
        if (b == bp) t = (a == 0) ? 1.0 : 0.0;  // dummy direction for the unused diagonal pair
        tv[(int64_t)a * np + i] = t;
        r2 += t * t;
    }
    dist[i] = sqrt(r2);
}

// ---- translation-vector de-duplication --------------------------------------------------------------
// The (S|R)(t) block depends on the pair (b, b') only through t = c_b - c_b'.  Regular geometries repeat the
// same t many times (4x4 grid: 48 distinct t among 240 ordered pairs; 8x8 grid: 224 among 4032), so the
// contraction is done once per DISTINCT t and the result is scaled and written for every pair that shares it.
// rep[i] = smallest pair index with bit-identical t (exact comparison: irregular geometries simply get U = np - B).
// pm = 1 (register-resident kernel): pairs with t and -t share a group as well.  Every term of (S|R)_{h',h} has
// n'' = n + n' (mod 2) and S_{h''}(-t) = (-1)^{n''} S_{h''}(t), so (S|R)_{h',h}(-t) = (-1)^{n+n'} (S|R)_{h',h}(t): the block of
// (b', b) is the block of (b, b') up to that sign, for ANY geometry (c_b - c_b' and c_b' - c_b are exact negatives).  Members
// whose translation is the negative of the representative's carry bit 31 in their packed entry.
__global__ void pair_rep_kernel(int d, int B, int dedupe, int pm, const double* __restrict__ tv, int32_t* __restrict__ rep) {
    extern __shared__ __align__(16) double s_tv[];  // [d][np] when it fits (use_smem), else the scan reads global memory
    const int np = B * B;
    const bool use_smem = dedupe && (size_t)np * d * sizeof(double) <= 96 * 1024;
    if (use_smem) {
        for (int e = threadIdx.x; e < np * d; e += blockDim.x) s_tv[e] = tv[e];
        __syncthreads();
    }
    const double* T = use_smem ? s_tv : tv;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    if (i / B == i % B) { rep[i] = -1; return; }
    if (!dedupe) {  // no search: only the pair and its transpose share a group
        const int j = (i % B) * B + i / B;
        rep[i] = (pm && j < i) ? j : i;
        return;
    }
    double t[BHS_MAX_NODES + 2];
    for (int a = 0; a < d; ++a) t[a] = T[(int64_t)a * np + i];
    int r = i;
    for (int j = 0; j < i; ++j) {
        // first component differs (the diagonal pairs carry a dummy direction: checked below)
        const double tj = T[j];
        const bool eq = tj == t[0], ng = pm && tj == -t[0];
        if (!eq && !ng) continue;
        if (j / B == j % B) continue;
        bool same = eq, opp = ng;
        for (int a = 1; a < d; ++a) {
            const double v = T[(int64_t)a * np + j];
            same = same && (v == t[a]);
            opp = opp && (v == -t[a]);
        }
        if (same || opp) { r = j; break; }
    }
    rep[i] = r;
}
// Single CTA: number the distinct translations and bucket the pairs.
//   n_unique[0] = U;  grp_rep[u] = representative pair;  grp_start[u..u+1) -> members[] ((b << 16) | b')
// uid[] and cursor[] are scratch of np ints each.
__global__ void __launch_bounds__(1024) pair_group_kernel(int B, int d, int pm, const double* __restrict__ tv,
                                                          const int32_t* __restrict__ rep, int32_t* __restrict__ uid,
                                                          int32_t* __restrict__ cursor, int32_t* __restrict__ n_unique,
                                                          int32_t* __restrict__ grp_rep, int32_t* __restrict__ grp_start,
                                                          int32_t* __restrict__ members) {
    __shared__ int s_scan[1024];
    __shared__ int s_base;
    const int np = B * B, tid = threadIdx.x, T = blockDim.x;
    if (tid == 0) s_base = 0;
    __syncthreads();
    // pass 1: uid[i] = index of pair i among the representatives (exclusive scan of the flags), chunk by chunk
    for (int c0 = 0; c0 < np; c0 += T) {
        const int i = c0 + tid;
        const int flag = (i < np && rep[i] == i) ? 1 : 0;
        s_scan[tid] = flag;
        __syncthreads();
        for (int o = 1; o < T; o <<= 1) {
            int v = (tid >= o) ? s_scan[tid - o] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int incl = s_scan[tid], base = s_base;
        if (flag) {
            uid[i] = base + incl - 1;
            grp_rep[base + incl - 1] = i;
        }
        __syncthreads();
        if (tid == T - 1) s_base = base + incl;
        __syncthreads();
    }
    const int U = s_base;
    if (tid == 0) n_unique[0] = U;
    for (int u = tid; u <= U; u += T) { cursor[u] = 0; }
    __syncthreads();
    // pass 2: group sizes
    for (int i = tid; i < np; i += T)
        if (rep[i] >= 0) atomicAdd(&cursor[uid[rep[i]]], 1);
    __syncthreads();
    // pass 3: exclusive scan of the sizes -> grp_start (serial over chunks as above)
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < U; c0 += T) {
        const int u = c0 + tid;
        const int cnt = (u < U) ? cursor[u] : 0;
        s_scan[tid] = cnt;
        __syncthreads();
        for (int o = 1; o < T; o <<= 1) {
            int v = (tid >= o) ? s_scan[tid - o] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int incl = s_scan[tid], base = s_base;
        if (u < U) grp_start[u] = base + incl - cnt;
        __syncthreads();
        if (tid == T - 1) s_base = base + incl;
        __syncthreads();
    }
    if (tid == 0) grp_start[U] = s_base;
    for (int u = tid; u < U; u += T) cursor[u] = 0;
    __syncthreads();
    // pass 4: fill (order inside a group is irrelevant: every member gets the same block)
    for (int i = tid; i < np; i += T)
        if (rep[i] >= 0) {
            const int r = rep[i], u = uid[r];
            int neg = 0;  // translation of pair i is the negative of its representative's (first nonzero component decides)
            if (pm)
                for (int a = 0; a < d; ++a) {
                    const double ti = tv[(int64_t)a * np + i], tr = tv[(int64_t)a * np + r];
                    if (ti != tr) { neg = 1; break; }
                }
            members[grp_start[u] + atomicAdd(&cursor[u], 1)] = (neg << 31) | ((i / B) << 16) | (i % B);  // packed (b, b')
        }
}

// diagonal blocks A[(b,h),(b,h')] = delta_{hh'} diag[s][b][h]
__global__ void diag_blocks_kernel(int B, int H, int b_lo, const cplx* __restrict__ diag, cplx* __restrict__ A, int64_t ld,
                                   int64_t sys_stride) {
    const int sys = blockIdx.z, b = b_lo + blockIdx.y;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)H * H) return;
    const int h = (int)(e / H), hp = (int)(e % H);
    cplx v = cmake(0.0, 0.0);
    if (h == hp) v = diag[((int64_t)sys * B + b) * H + h];
    A[(int64_t)sys * sys_stride + ((int64_t)(b - b_lo) * H + h) * ld + (int64_t)b * H + hp] = v;
}

struct SmArrA {
    double* p;
    int stride;
    __device__ __forceinline__ double& operator[](int n) const { return p[(size_t)n * stride]; }
};
// hp[s][pair][n''] = i^{n''} h_{n''}(k_s |t_pair|), n'' < L2
__global__ void pair_radial_kernel(int d, int L2, int n_store, int B, int nsys, const double* __restrict__ ks,
                                   const double* __restrict__ dist, cplx* __restrict__ hp, double* __restrict__ scratch) {
    extern __shared__ __align__(16) double sm[];
    // order sequences in shared memory, or (very high orders) in a global scratch: see ball_radial_kernel
    const int T = scratch ? gridDim.x * blockDim.x : blockDim.x;
    const int t = scratch ? blockIdx.x * blockDim.x + threadIdx.x : threadIdx.x;
    double* base = scratch ? scratch : sm;
    SmArrA hr{base + t, T};
    SmArrA hi{base + (size_t)n_store * T + t, T};
    int64_t np = (int64_t)B * B, total = np * nsys;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int s = (int)(i / np);
        int64_t pr = i % np;
        if (pr / B == pr % B) {
            for (int n = 0; n < L2; ++n) hp[i * L2 + n] = cmake(0.0, 0.0);
            continue;
        }
        hankel_upward(d, ks[s] * dist[pr], L2 - 1, hr, hi);
        for (int n = 0; n < L2; ++n) hp[i * L2 + n] = cmul_ipow(cmake(hr[n], hi[n]), n);
    }
}

// per-system row / column / diagonal factors:
//   rowf[s][b][h] = i^n (alpha_b j_n + beta_b k j_n')      colf[s][b][h] = (-i)^n SD_n(rho_b)
//   diag[s][b][h] = SD_n (alpha_b h_n + beta_b k h_n')
__global__ void factors_kernel(int d, int L, int H, int B, int nsys, const double* __restrict__ radii,
                               const double* __restrict__ ks, const double* __restrict__ etas,
                               const cplx* __restrict__ alpha, const cplx* __restrict__ beta,
                               const double4* __restrict__ rad, const int32_t* __restrict__ deg,
                               cplx* __restrict__ rowf, cplx* __restrict__ colf, cplx* __restrict__ diag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)nsys * B * H;
    if (i >= total) return;
    int h = (int)(i % H);
    int b = (int)((i / H) % B);
    int s = (int)(i / ((int64_t)H * B));
    int n = deg[h];
    double k = ks[s], eta = etas ? etas[s] : 1.0;
    double4 r = rad[((int64_t)s * B + b) * L + n];
    cplx al = alpha ? alpha[b] : cmake(1.0, 0.0);
    cplx be = beta ? beta[b] : cmake(0.0, 0.0);
    cplx bek = cscale(be, k);
    cplx reg = cadd(cscale(al, r.x), cscale(bek, r.y));
    cplx sing = cadd(cmul(al, cmake(r.x, r.z)), cmul(bek, cmake(r.y, r.w)));
    cplx sd = sd_coef(d, k, eta, radii[b], r.x, r.y);
    if (rowf) rowf[i] = cmul_ipow(reg, n);
    if (colf) colf[i] = cmul_ipow(sd, -n);
    if (diag) diag[i] = cmul(sd, sing);
}

// ---- complex wavenumber variants: same outputs, complex arguments -----------------------------------------
struct SmArrZ {
    cplx* p;
    int stride;
    __device__ __forceinline__ cplx& operator[](int n) const { return p[(size_t)n * stride]; }
};
__global__ void pair_radial_z_kernel(int d, int L2, int n_store, int B, int nsys, const double* __restrict__ kr,
                                     const double* __restrict__ ki, const double* __restrict__ dist,
                                     cplx* __restrict__ hp) {
    extern __shared__ __align__(16) cplx smz[];
    const int T = blockDim.x;
    SmArrZ ah{smz + threadIdx.x, T};
    int64_t np = (int64_t)B * B, total = np * nsys;
    for (int64_t i = (int64_t)blockIdx.x * T + threadIdx.x; i < total; i += (int64_t)gridDim.x * T) {
        int s = (int)(i / np);
        int64_t pr = i % np;
        if (pr / B == pr % B) {
            for (int n = 0; n < L2; ++n) hp[i * L2 + n] = cmake(0.0, 0.0);
            continue;
        }
        const double r = dist[pr];
        radial_sequence_z(d, cmake(kr[s] * r, ki[s] * r), L2 - 1, ah, ah, false, true);
        for (int n = 0; n < L2; ++n) hp[i * L2 + n] = cmul_ipow(ah[n], n);
    }
    (void)n_store;
}

__global__ void factors_z_kernel(int d, int L, int H, int B, int nsys, const double* __restrict__ radii,
                                 const double* __restrict__ kr, const double* __restrict__ ki,
                                 const double* __restrict__ etas, const cplx* __restrict__ alpha,
                                 const cplx* __restrict__ beta, const cplx* __restrict__ radz,
                                 const int32_t* __restrict__ deg, cplx* __restrict__ rowf, cplx* __restrict__ colf,
                                 cplx* __restrict__ diag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)nsys * B * H;
    if (i >= total) return;
    int h = (int)(i % H);
    int b = (int)((i / H) % B);
    int s = (int)(i / ((int64_t)H * B));
    int n = deg[h];
    const cplx k = cmake(kr[s], ki[s]);
    const double eta = etas ? etas[s] : 1.0;
    const cplx* r = radz + (((int64_t)s * B + b) * L + n) * 4;  // j, j', h, h'
    cplx al = alpha ? alpha[b] : cmake(1.0, 0.0);
    cplx be = beta ? beta[b] : cmake(0.0, 0.0);
    cplx bek = cmul(be, k);
    cplx reg = cadd(cmul(al, r[0]), cmul(bek, r[1]));
    cplx sing = cadd(cmul(al, r[2]), cmul(bek, r[3]));
    cplx sd = sd_coef_z(d, k, eta, radii[b], r[0], r[1]);
    if (rowf) rowf[i] = cmul_ipow(reg, n);
    if (colf) colf[i] = cmul_ipow(sd, -n);
    if (diag) diag[i] = cmul(sd, sing);
}

// ---- the assembly kernel --------------------------------------------------------------------------------
struct AsmArgs {
    int B, H, H2, L2, nt_res;
    int b_lo, b_hi;            // block rows (row balls) written by this call; the strip starts at row ball b_lo
    int sy_global;             // 1: the S window does not fit in shared memory, gather from global (huge 2-D bands)
    const int32_t* n_unique;   // [1]   number of distinct translation vectors U
    const int32_t* grp_rep;    // [U]   representative pair of each
    const int32_t* grp_start;  // [U+1] member ranges
    const int32_t* members;    // pairs packed as (b << 16) | b', bucketed by translation
    const bhs_tile_hdr* tiles;
    const double* coef;
    const uint16_t* cidx;
    const int32_t* deg;   // [H]  degree of each harmonic
    const int32_t* deg2;
    const cplx* Y2;    // [B*B][H2]
    const cplx* hp;    // [nsys][B*B][L2]
    const cplx* Su;    // [nsys][ucap][H2]  S vectors of the distinct translations (register-resident kernel only)
    int ucap;
    const cplx* rowf;  // [nsys][B][H]
    const cplx* colf;
    const cplx* diag;
    cplx* A;
    int64_t ld, sys_stride;
};

__global__ void __launch_bounds__(ASM_THREADS) assemble_kernel(AsmArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const bhs_tile_hdr hd = a.tiles[blockIdx.x];
    const int tiles_c = (a.H + BHS_TILE_C - 1) / BHS_TILE_C;
    const int tr = blockIdx.x / tiles_c, tc = blockIdx.x % tiles_c;
    const int sys = blockIdx.z;
    const int nt = hd.nt, nt_res = a.nt_res;
    double* s_coef = reinterpret_cast<double*>(smem_raw);
    uint16_t* s_idx = reinterpret_cast<uint16_t*>(smem_raw + (size_t)nt_res * ASM_LAYER_COEF);
    cplx* s_sy = reinterpret_cast<cplx*>(smem_raw + (size_t)nt_res * (ASM_LAYER_COEF + ASM_LAYER_IDX));
    const bool resident = nt <= nt_res;
    uint32_t phase = 0;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (resident && nt > 0) {
        if (tid == 0) {
            mbar_expect_tx(&bar, (uint32_t)(nt * (ASM_LAYER_COEF + ASM_LAYER_IDX)));
            tma_load_1d(s_coef, a.coef + hd.coef_off, (uint32_t)(nt * ASM_LAYER_COEF), &bar);
            tma_load_1d(s_idx, a.cidx + hd.idx_off, (uint32_t)(nt * ASM_LAYER_IDX), &bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
    }
    // this thread's entries of the tile (ASM_EPT of them, ASM_THREADS apart: a warp covers 32 consecutive columns of a row)
    int h_[ASM_EPT], hp_[ASM_EPT];
    bool ok_[ASM_EPT];
#pragma unroll
    for (int q = 0; q < ASM_EPT; ++q) {
        const int e = tid + q * ASM_THREADS;
        h_[q] = tr * BHS_TILE_R + e / BHS_TILE_C;
        hp_[q] = tc * BHS_TILE_C + e % BHS_TILE_C;
        ok_[q] = h_[q] < a.H && hp_[q] < a.H;
    }
    const int64_t npairs = (int64_t)a.B * a.B;
    cplx* Asys = a.A + (int64_t)sys * a.sys_stride;
    const cplx* rowf = a.rowf + (int64_t)sys * a.B * a.H;
    const cplx* colf = a.colf + (int64_t)sys * a.B * a.H;
    const int U = a.n_unique[0];

    // distinct translations u = blockIdx.y, blockIdx.y + gridDim.y, ...: contract once, write for every member pair
    for (int u = blockIdx.y; u < U; u += gridDim.y) {
        const int64_t pr = a.grp_rep[u];
        // S_{h''}(t) for the index window this tile references
        const cplx* y2 = a.Y2 + pr * a.H2 + hd.sy_lo;
        const cplx* hpw = a.hp + ((int64_t)sys * npairs + pr) * a.L2;
        const int32_t* dg = a.deg2 + hd.sy_lo;
        if (!a.sy_global)
            for (int j = tid; j < hd.sy_cnt; j += ASM_THREADS) s_sy[j] = cmul(y2[j], hpw[dg[j]]);
        __syncthreads();
        double ar[ASM_EPT], ai[ASM_EPT];
#pragma unroll
        for (int q = 0; q < ASM_EPT; ++q) { ar[q] = 0.0; ai[q] = 0.0; }
        for (int t0 = 0; t0 < nt; t0 += nt_res) {
            const int tn = min(nt_res, nt - t0);
            if (!resident) {
                if (tid == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(&bar, (uint32_t)(tn * (ASM_LAYER_COEF + ASM_LAYER_IDX)));
                    tma_load_1d(s_coef, a.coef + hd.coef_off + (int64_t)t0 * BHS_TILE_E, (uint32_t)(tn * ASM_LAYER_COEF), &bar);
                    tma_load_1d(s_idx, a.cidx + hd.idx_off + (int64_t)t0 * BHS_TILE_E, (uint32_t)(tn * ASM_LAYER_IDX), &bar);
                }
                mbar_wait(&bar, phase);
                phase ^= 1;
            }
#pragma unroll 4
            for (int t = 0; t < tn; ++t) {
#pragma unroll
                for (int q = 0; q < ASM_EPT; ++q) {
                    const int e = tid + q * ASM_THREADS;
                    const double cf = s_coef[t * BHS_TILE_E + e];
                    const int ix = s_idx[t * BHS_TILE_E + e];
                    const cplx sv = a.sy_global ? cmul(y2[ix], hpw[dg[ix]]) : s_sy[ix];
                    ar[q] = fma(cf, sv.x, ar[q]);
                    ai[q] = fma(cf, sv.y, ai[q]);
                }
            }
            if (!resident) __syncthreads();  // everyone done with the chunk before it is overwritten
        }
        // Write phase: every pair that shares this translation gets the block, scaled by its own row / column factors.
        // Members are handled four at a time with all factor loads issued before the first store (the loop is
        // otherwise a chain of dependent global-memory latencies).
        const int q_end = a.grp_start[u + 1];
        for (int q0 = a.grp_start[u]; q0 < q_end; q0 += 4) {
            int bb[4], bq[4];
            cplx rf[4][ASM_EPT], cf[4][ASM_EPT];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int pk = (q0 + i < q_end) ? __ldg(a.members + q0 + i) : -1;
                bb[i] = pk >> 16;
                bq[i] = pk & 0xffff;
                if (bb[i] < a.b_lo || bb[i] >= a.b_hi) bb[i] = -1;  // row ball outside the strip of this call
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (bb[i] < 0) continue;
#pragma unroll
                for (int q = 0; q < ASM_EPT; ++q)
                    if (ok_[q]) {
                        rf[i][q] = __ldg(rowf + (int64_t)bb[i] * a.H + h_[q]);
                        cf[i][q] = __ldg(colf + (int64_t)bq[i] * a.H + hp_[q]);
                    }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (bb[i] < 0) continue;
#pragma unroll
                for (int q = 0; q < ASM_EPT; ++q)
                    if (ok_[q])
                        Asys[((int64_t)(bb[i] - a.b_lo) * a.H + h_[q]) * a.ld + (int64_t)bq[i] * a.H + hp_[q]] =
                            cmul(cmul(cmake(ar[q], ai[q]), rf[i][q]), cf[i][q]);
            }
        }
        __syncthreads();  // s_sy reuse
    }
}

// ---- the assembly kernel, register-resident form ------------------------------------------------------------
// Used whenever a tile has at most 32 term layers and two S windows fit in shared memory (every 3-D plan up to
// n_end = 32, small 2-D plans).  Same arithmetic, term order and roundings as assemble_kernel above (bit-identical
// matrices), different data movement:
//   * each thread keeps the coefficients and S indices of ITS tile entry in registers for the whole CTA lifetime
//     (NT doubles + NT/2 packed index words), loaded once with coalesced global loads: the contraction issues one
//     shared-memory load per term instead of three, and shared memory only holds S windows;
//   * the S vectors of the distinct translations are built once by s_vectors_kernel; a CTA streams the windows of its
//     translations through a two-stage TMA ring, two translations ahead of the contraction;
//   * the member list of the NEXT translation is fetched into shared memory while the current one is written, so the
//     write phase is  LDS -> factor loads (L1) -> stores  with no chain of dependent global loads in front of it;
//   * the diagonal blocks are written by the same CTAs (no separate launch).
#define ASM_MEMCAP 256  // members of one translation held in shared memory (more: read from global)

__global__ void s_vectors_kernel(int H2, int L2, int64_t np, int ucap, const int32_t* __restrict__ n_unique,
                                 const int32_t* __restrict__ grp_rep, const int32_t* __restrict__ deg2,
                                 const cplx* __restrict__ Y2, const cplx* __restrict__ hp, cplx* __restrict__ Su) {
    const int U = n_unique[0], sys = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= H2) return;
    const int dg = deg2[j];
    for (int u = blockIdx.y; u < U; u += gridDim.y) {
        const int64_t pr = grp_rep[u];
        Su[((int64_t)sys * ucap + u) * H2 + j] = cmul(Y2[pr * H2 + j], hp[((int64_t)sys * np + pr) * L2 + dg]);
    }
}

// One member pair of a translation, decoded once per CTA into the three offsets the write phase adds to its
// thread-constant bases.  (rs, cs): strides of the row / column factor tables in BYTES per ball -- H * 16 when the factors
// are read from global memory, TILE_R * 16 / TILE_C * 16 when the tile's factors are staged in shared memory.
struct __align__(16) AsmMember {
    int32_t row_off;  // b * rs   (-1: row ball outside the strip of this call, or padding)
    int32_t col_off;  // b' * cs, bit 31: the member's translation is the NEGATIVE of the group's (block sign (-1)^(n+n'))
    int64_t out_off;  // (b - b_lo) * H * ld + b' * H   (elements)
};
__device__ __forceinline__ AsmMember asm_member(int pk, int H, int rs, int cs, int b_lo, int b_hi, int64_t hld) {
    const int b = (pk >> 16) & 0x7fff, bq = pk & 0xffff;
    AsmMember m;
    m.row_off = (b >= b_lo && b < b_hi) ? b * rs : -1;
    m.col_off = (bq * cs) | (pk & (int)0x80000000);
    m.out_off = (int64_t)(b - b_lo) * hld + (int64_t)bq * H;
    return m;
}
// keeps a computed shared-window address in a register (ptxas otherwise re-derives it from %cta-id / %tid at every use
// when registers are tight: ~15 instructions per use)
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ cplx lds_cplx(uint32_t addr) {
    cplx v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ AsmMember lds_member(uint32_t addr) {
    AsmMember m;
    int32_t lo, hi;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(m.row_off), "=r"(m.col_off), "=r"(lo), "=r"(hi) : "r"(addr) : "memory");
    m.out_off = (int64_t)(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo);
    return m;
}

// FSM: the row / column factors of this tile position (B x 4 and B x 64 values) are staged in shared memory once per CTA,
// so the write phase has no global load at all (taken when they fit: B <= ASM_FSM_MAXB)
#define ASM_FSM_MAXB 36
template <int NT, bool FSM>
__global__ void __launch_bounds__(ASM_THREADS, 2) assemble_reg_kernel(AsmArgs a, int stage_bytes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[2];
    __shared__ AsmMember s_mem[2][ASM_MEMCAP];
    const int tid = threadIdx.x;
    const bhs_tile_hdr hd = a.tiles[blockIdx.x];
    const int tiles_c = (a.H + BHS_TILE_C - 1) / BHS_TILE_C;
    const int tr = blockIdx.x / tiles_c, tc = blockIdx.x % tiles_c;
    const int sys = blockIdx.z;
    const int nt = hd.nt;
    const int h = tr * BHS_TILE_R + tid / BHS_TILE_C, hp = tc * BHS_TILE_C + tid % BHS_TILE_C;
    const bool ok = h < a.H && hp < a.H;
    const int U = a.n_unique[0];
    const int ustep = gridDim.y;
    const int64_t hld = (int64_t)a.H * a.ld;
    const uint32_t win_bytes = (uint32_t)hd.sy_cnt * (uint32_t)sizeof(cplx);
    const cplx* su = a.Su + (int64_t)sys * a.ucap * a.H2 + hd.sy_lo;
    const uint32_t sa_stage = pin_u32(smem_u32(smem_raw));  // shared-window addresses: S stages, member lists, staged factors
    const uint32_t sa_mem = pin_u32(smem_u32(&s_mem[0][0]));
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int u = blockIdx.y + s * ustep;
            if (u < U) {
                mbar_expect_tx(&full[s], win_bytes);
                tma_load_1d(smem_raw + (size_t)s * stage_bytes, su + (int64_t)u * a.H2, win_bytes, &full[s]);
            }
        }
    }
    // member range of the first translation, and (below) its members
    int q0 = 0, cnt = 0;
    if ((int)blockIdx.y < U) {
        q0 = __ldg(a.grp_start + blockIdx.y);
        cnt = __ldg(a.grp_start + blockIdx.y + 1) - q0;
    }
    // this thread's tile entry: coefficients and packed S indices of every term layer
    double cf_[NT];
    uint32_t ix_[NT / 2];
    {
        const double* cp = a.coef + hd.coef_off + tid;
        const uint16_t* ip = a.cidx + hd.idx_off + tid;
#pragma unroll
        for (int t = 0; t < NT; ++t) cf_[t] = (t < nt) ? __ldg(cp + t * BHS_TILE_E) : 0.0;
#pragma unroll
        for (int t = 0; t < NT; t += 2) {
            const uint32_t lo = (t < nt) ? __ldg(ip + t * BHS_TILE_E) : 0u;
            const uint32_t hi = (t + 1 < nt) ? __ldg(ip + (t + 1) * BHS_TILE_E) : 0u;
            ix_[t / 2] = (lo | (hi << 16)) * (uint32_t)sizeof(cplx);  // byte offsets (windows are < 4096 entries)
        }
    }
    // thread-constant bases: everything a member adds is one of its three precomputed offsets
    const unsigned char* rowf_h = reinterpret_cast<const unsigned char*>(a.rowf + (int64_t)sys * a.B * a.H + h);
    const unsigned char* colf_hp = reinterpret_cast<const unsigned char*>(a.colf + (int64_t)sys * a.B * a.H + hp);
    const int rs = (FSM ? BHS_TILE_R : a.H) * (int)sizeof(cplx), cs = (FSM ? BHS_TILE_C : a.H) * (int)sizeof(cplx);
    uint32_t sa_rowf = 0, sa_colf = 0;
    if (FSM) {
        cplx* s_rowf = reinterpret_cast<cplx*>(smem_raw + (size_t)2 * stage_bytes);
        cplx* s_colf = s_rowf + (size_t)a.B * BHS_TILE_R;
        const cplx* gr = a.rowf + (int64_t)sys * a.B * a.H + tr * BHS_TILE_R;
        const cplx* gc = a.colf + (int64_t)sys * a.B * a.H + tc * BHS_TILE_C;
        for (int e = tid; e < a.B * BHS_TILE_R; e += ASM_THREADS) {
            const int b = e / BHS_TILE_R, r = e % BHS_TILE_R;
            s_rowf[e] = (tr * BHS_TILE_R + r < a.H) ? __ldg(gr + (int64_t)b * a.H + r) : cmake(0.0, 0.0);
        }
        for (int e = tid; e < a.B * BHS_TILE_C; e += ASM_THREADS) {
            const int b = e / BHS_TILE_C, c = e % BHS_TILE_C;
            s_colf[e] = (tc * BHS_TILE_C + c < a.H) ? __ldg(gc + (int64_t)b * a.H + c) : cmake(0.0, 0.0);
        }
        sa_rowf = pin_u32(smem_u32(s_rowf + tid / BHS_TILE_C));
        sa_colf = pin_u32(smem_u32(s_colf + tid % BHS_TILE_C));
    }
    cplx* out_base = a.A + (int64_t)sys * a.sys_stride + (int64_t)h * a.ld + hp;
    // block sign of a member whose translation is the negative of its group's: (-1)^(n + n')
    const double sflip = (ok && ((__ldg(a.deg + h) + __ldg(a.deg + hp)) & 1)) ? -1.0 : 1.0;
    // diagonal blocks of this tile position: balls b_lo + y, b_lo + y + gridDim.y, ...
    if (ok) {
        const cplx* dg = a.diag + (int64_t)sys * a.B * a.H + h;
        for (int b = a.b_lo + blockIdx.y; b < a.b_hi; b += ustep) {
            cplx v = cmake(0.0, 0.0);
            if (h == hp) v = __ldg(dg + (int64_t)b * a.H);
            out_base[(int64_t)(b - a.b_lo) * hld + (int64_t)b * a.H] = v;
        }
    }
    // member lists live in shared memory padded to a multiple of four with inactive entries (row_off = -1); the rare
    // members beyond ASM_MEMCAP (more spheres than that sharing one translation) are written by the tail loop below
    auto fetch_member = [&](int q_first, int n) -> int {  // the global load (issued early) ...
        return (tid < n && tid < ASM_MEMCAP) ? __ldg(a.members + q_first + tid) : 0;
    };
    auto store_member = [&](int buf, int pk, int n) {  // ... and the decoded entry (stored once the buffer is free)
        const int n4 = min((n + 3) & ~3, ASM_MEMCAP);
        if (tid < n4) {
            AsmMember m;
            m.row_off = -1; m.col_off = 0; m.out_off = 0;
            if (tid < n) m = asm_member(pk, a.H, rs, cs, a.b_lo, a.b_hi, hld);
            s_mem[buf][tid] = m;
        }
    };
    store_member(0, fetch_member(q0, cnt), cnt);
    __syncthreads();  // barrier inits, staged factors and s_mem[0] visible

    int it = 0;
    for (int u = blockIdx.y; u < U; u += ustep, ++it) {
        const int st = it & 1;
        // member range of the next translation (consumed after the contraction)
        const int un = u + ustep;
        int q0n = 0, cntn = 0;
        if (un < U) {
            q0n = __ldg(a.grp_start + un);
            cntn = __ldg(a.grp_start + un + 1) - q0n;
        }
        mbar_wait(&full[st], (uint32_t)((it >> 1) & 1));
        const uint32_t sa_sy = sa_stage + (uint32_t)st * (uint32_t)stage_bytes;
        double ar = 0.0, ai = 0.0;
        // four layers per (CTA-uniform) branch: their loads are issued together; a padded layer reads entry 0 of the
        // window and its FMAs are skipped, so the sums are those of a layer-by-layer loop.  (Issuing the next four loads
        // ahead of the FMAs was measured slower: the 16 extra registers spill.)
#pragma unroll
        for (int t0 = 0; t0 < NT; t0 += 4) {
            if (t0 < nt) {
                cplx sv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int t = t0 + i;
                    const uint32_t ix = (t & 1) ? (ix_[t / 2] >> 16) : (ix_[t / 2] & 0xffffu);
                    sv[i] = lds_cplx(sa_sy + ix);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool on = t0 + i < nt;
                    ar = on ? fma(cf_[t0 + i], sv[i].x, ar) : ar;
                    ai = on ? fma(cf_[t0 + i], sv[i].y, ai) : ai;
                }
            }
        }
        const int pk_next = fetch_member(q0n, cntn);
        __syncthreads();  // everyone is done with stage st (refill it) and with the write phase of the previous translation,
                          // i.e. with s_mem[st ^ 1]: only now may the next translation's members overwrite it (they are read
                          // after the next barrier)
        store_member(st ^ 1, pk_next, cntn);
        if (tid == 0) {
            const int u2 = u + 2 * ustep;
            if (u2 < U) {
                mbar_expect_tx(&full[st], win_bytes);
                tma_load_1d(smem_raw + (size_t)st * stage_bytes, su + (int64_t)u2 * a.H2, win_bytes, &full[st]);
            }
        }
        // write phase: every pair that shares this translation, four at a time, factor loads ahead of the stores
        if (ok) {
            const cplx v = cmake(ar, ai);
            const int n4 = min((cnt + 3) & ~3, ASM_MEMCAP);
            uint32_t sa_m = sa_mem + (uint32_t)st * (uint32_t)(ASM_MEMCAP * sizeof(AsmMember));
            for (int m0 = 0; m0 < n4; m0 += 4, sa_m += 4 * (uint32_t)sizeof(AsmMember)) {
                AsmMember mm[4];
                cplx rf[4], cf[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) mm[i] = lds_member(sa_m + i * (uint32_t)sizeof(AsmMember));
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (mm[i].row_off >= 0) {
                        const int co = mm[i].col_off & 0x7fffffff;
                        rf[i] = FSM ? lds_cplx(sa_rowf + mm[i].row_off) : __ldg(reinterpret_cast<const cplx*>(rowf_h + mm[i].row_off));
                        cf[i] = FSM ? lds_cplx(sa_colf + co) : __ldg(reinterpret_cast<const cplx*>(colf_hp + co));
                    }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (mm[i].row_off >= 0) {
                        const double sg = mm[i].col_off < 0 ? sflip : 1.0;
                        out_base[mm[i].out_off] = cmul(cmul(cmake(v.x * sg, v.y * sg), rf[i]), cf[i]);
                    }
            }
            for (int m = ASM_MEMCAP; m < cnt; ++m) {  // overflow of the shared-memory member list
                const AsmMember mo = asm_member(__ldg(a.members + q0 + m), a.H, a.H * (int)sizeof(cplx), a.H * (int)sizeof(cplx),
                                                a.b_lo, a.b_hi, hld);
                if (mo.row_off >= 0) {
                    const double sg = mo.col_off < 0 ? sflip : 1.0;
                    out_base[mo.out_off] = cmul(cmul(cmake(v.x * sg, v.y * sg), __ldg(reinterpret_cast<const cplx*>(rowf_h + mo.row_off))),
                                                __ldg(reinterpret_cast<const cplx*>(colf_hp + (mo.col_off & 0x7fffffff))));
                }
            }
        }
        q0 = q0n;
        cnt = cntn;
    }
}

// Which plans / problem sizes take the register-resident kernel; the S-vector scratch it needs (bytes, 0 if not taken).
static bool asm_reg_shape(const bhs_plan* p, int B, int nsys, int* stage_bytes, int64_t* su_bytes) {
    const int64_t np = (int64_t)B * B;
    const int64_t ucap = np - B;
    const int sb = (int)((((int64_t)p->max_sy_cnt * (int64_t)sizeof(cplx)) + 127) & ~(int64_t)127);
    const int64_t sub = (int64_t)nsys * ucap * p->H2 * (int64_t)sizeof(cplx);
    if (stage_bytes) *stage_bytes = sb;
    if (su_bytes) *su_bytes = 0;
    if (B < 2 || p->max_nt > 32 || (int64_t)p->max_sy_cnt * (int64_t)sizeof(cplx) > 48 * 1024 || sub > ((int64_t)1 << 30))
        return false;
    // the workspace layout depends on the shape only -- not on the A/B switch below, which may change between the
    // workspace query and the call (tests flip it while a caller keeps its workspace)
    if (su_bytes) *su_bytes = sub;
    return getenv("BHS_ASM_LEGACY") == nullptr;  // A/B switch (read per call): set = the shared-memory-resident kernel
}

// ---- host entries -------------------------------------------------------------------------------------------
static inline int64_t al256(int64_t v) { return (v + 255) & ~(int64_t)255; }

struct AsmWork {
    double4* rad;
    double* tv;
    double* dist;
    cplx* Y2;
    cplx* hp;
    cplx* rowf;
    cplx* colf;
    cplx* diag;
    int32_t *rep, *uid, *cursor, *n_unique, *grp_rep, *grp_start, *members;
    cplx* Su;         // S vectors of the distinct translations (register-resident kernel), else null
    double* scratch;  // global order-sequence scratch of the radial kernels (very high orders only), else null
    int64_t bytes;
};
// shape of pair_radial_kernel's order sequences (orders 0 .. L2 of the pair distances)
static void pair_radial_shape(const bhs_plan* p, int& n_store, int& T, size_t& smem) {
    const int d = p->d;
    int shift = (d & 1) ? (d - 3) / 2 : d / 2 - 1;
    n_store = p->L2 + 1 + shift;
    T = 64;
    while (T > 32 && (size_t)2 * n_store * T * sizeof(double) > 160 * 1024) T >>= 1;
    smem = (size_t)2 * n_store * T * sizeof(double);
}
static AsmWork carve(const bhs_plan* p, int B, int nsys, void* base) {
    AsmWork w;
    unsigned char* c = (unsigned char*)base;
    int64_t np = (int64_t)B * B, off = 0;
    auto take = [&](int64_t bytes) { unsigned char* r = c + off; off += al256(bytes); return r; };
    w.rad = (double4*)take((int64_t)nsys * B * p->n_end * 4 * sizeof(cplx));  // (j, j', y, y') real or (j, j', h, h') complex
    w.tv = (double*)take(np * p->d * sizeof(double));
    w.dist = (double*)take(np * sizeof(double));
    w.Y2 = (cplx*)take(np * p->H2 * sizeof(cplx));
    w.hp = (cplx*)take((int64_t)nsys * np * p->L2 * sizeof(cplx));
    w.rowf = (cplx*)take((int64_t)nsys * B * p->H * sizeof(cplx));
    w.colf = (cplx*)take((int64_t)nsys * B * p->H * sizeof(cplx));
    w.diag = (cplx*)take((int64_t)nsys * B * p->H * sizeof(cplx));
    w.rep = (int32_t*)take(np * 4);
    w.uid = (int32_t*)take(np * 4);
    w.cursor = (int32_t*)take((np + 1) * 4);
    w.n_unique = (int32_t*)take(256);
    w.grp_rep = (int32_t*)take(np * 4);
    w.grp_start = (int32_t*)take((np + 1) * 4);
    w.members = (int32_t*)take(np * 4);
    {
        int64_t su_bytes = 0;
        asm_reg_shape(p, B, nsys, nullptr, &su_bytes);
        w.Su = su_bytes > 0 ? (cplx*)take(su_bytes) : nullptr;
    }
    {
        int n_store, T;
        size_t smem;
        pair_radial_shape(p, n_store, T, smem);
        size_t scr = smem > 200 * 1024 ? (size_t)2 * n_store * 32 * T * sizeof(double) : 0;
        const size_t scr_ball = ball_radial_scratch_bytes(p->d, p->n_end);
        if (scr_ball > scr) scr = scr_ball;
        w.scratch = scr ? (double*)take((int64_t)scr) : nullptr;
    }
    w.bytes = off;
    return w;
}

extern "C" int64_t bhs_assemble_workspace(const bhs_plan_t* plan, int B, int nsys) {
    if (!plan || B <= 0 || nsys <= 0) return BHS_ERR_INVALID;
    return carve(plan, B, nsys, nullptr).bytes;
}

static int run_factors(const bhs_plan* p, int B, int nsys, const double* d_radii, const double* d_k,
                       const double* d_k_im, const double* d_eta, const double* d_alpha, const double* d_beta, AsmWork& w,
                       bool only_diag, cudaStream_t st) {
    int64_t tot = (int64_t)nsys * B * p->H;
    if (d_k_im) {
        int rc = launch_ball_radial_z(p->d, p->n_end, B, nsys, d_radii, d_k, d_k_im, 0.0, 0.0, (cplx*)w.rad, st);
        if (rc) return rc;
        factors_z_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(p->d, p->n_end, p->H, B, nsys, d_radii, d_k, d_k_im,
                                                                       d_eta, (const cplx*)d_alpha, (const cplx*)d_beta,
                                                                       (const cplx*)w.rad, p->d_deg,
                                                                       only_diag ? nullptr : w.rowf,
                                                                       only_diag ? nullptr : w.colf, w.diag);
        BHS_CHECK_LAUNCH();
        return BHS_OK;
    }
    int rc = launch_ball_radial(p->d, p->n_end, B, nsys, d_radii, d_k, 0.0, w.rad, w.scratch, st);
    if (rc) return rc;
    factors_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(p->d, p->n_end, p->H, B, nsys, d_radii, d_k, d_eta,
                                                                 (const cplx*)d_alpha, (const cplx*)d_beta, w.rad,
                                                                 p->d_deg, only_diag ? nullptr : w.rowf,
                                                                 only_diag ? nullptr : w.colf, w.diag);
    BHS_CHECK_LAUNCH();
    return BHS_OK;
}

int bhs_launch_harmonics_band2(const bhs_plan* plan, const double* d_xyz, int64_t npts, cplx* d_out, cudaStream_t st);

static int assemble_impl(const bhs_plan_t* plan, int B, int nsys, const double* d_centers, const double* d_radii,
                         const double* d_k, const double* d_k_im, const double* d_eta, const double* d_alpha,
                         const double* d_beta, int b_lo, int b_hi, double* d_A, int64_t ld, int64_t sys_stride, void* d_work,
                         void* stream) {
    if (!plan || B <= 0 || nsys <= 0 || !d_centers || !d_radii || !d_k || !d_A || !d_work) return BHS_ERR_INVALID;
    if (plan->tree != BHS_TREE_CHAIN) return BHS_ERR_INVALID;  // right-hand-side-only plan (no coupling table)
    if (b_lo < 0 || b_hi > B || b_lo >= b_hi) return BHS_ERR_INVALID;
    const int64_t N = (int64_t)B * plan->H;
    if (ld < N) return BHS_ERR_INVALID;
    if (B > 32767) return BHS_ERR_UNSUPPORTED;  // pairs are packed as (b << 16) | b' in 32 bits
    cudaStream_t st = (cudaStream_t)stream;
    AsmWork w = carve(plan, B, nsys, d_work);
    bhs_prof_begin(BHS_PROF_ASM_PRE, st);
    int rc = run_factors(plan, B, nsys, d_radii, d_k, d_k_im, d_eta, d_alpha, d_beta, w, false, st);
    if (rc) return rc;
    const int64_t np = (int64_t)B * B;
    pair_vectors_kernel<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(plan->d, B, d_centers, w.tv, w.dist);
    BHS_CHECK_LAUNCH();
    rc = bhs_harmonics(plan, 1, w.tv, np, (double*)w.Y2, stream);
    if (rc) return rc;
    {
        const int d = plan->d;
        const int shift = (d & 1) ? (d - 3) / 2 : d / 2 - 1;
        int n_store, T;
        size_t smem;
        pair_radial_shape(plan, n_store, T, smem);
        int64_t total = np * nsys, blocks = (total + T - 1) / T;
        if (blocks > bhs_sm_count() * 8) blocks = bhs_sm_count() * 8;
        if (d_k_im) {
            int Tz = 64;
            const int ns = plan->L2 + 1 + shift;
            while (Tz > 32 && (size_t)ns * Tz * sizeof(cplx) > 160 * 1024) Tz >>= 1;
            size_t smz = (size_t)ns * Tz * sizeof(cplx);
            if (smz > 200 * 1024) return BHS_ERR_UNSUPPORTED;
            cudaFuncSetAttribute(pair_radial_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smz);
            int64_t bz = (total + Tz - 1) / Tz;
            if (bz > bhs_sm_count() * 8) bz = bhs_sm_count() * 8;
            pair_radial_z_kernel<<<(unsigned)bz, Tz, smz, st>>>(d, plan->L2, ns, B, nsys, d_k, d_k_im, w.dist, w.hp);
        } else if (smem > 200 * 1024) {
            // very high orders: the sequences live in the workspace's global scratch
            if (blocks > 32) blocks = 32;
            pair_radial_kernel<<<(unsigned)blocks, T, 0, st>>>(d, plan->L2, n_store, B, nsys, d_k, w.dist, w.hp, w.scratch);
        } else {
            cudaFuncSetAttribute(pair_radial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            pair_radial_kernel<<<(unsigned)blocks, T, smem, st>>>(d, plan->L2, n_store, B, nsys, d_k, w.dist, w.hp, nullptr);
        }
        BHS_CHECK_LAUNCH();
    }
    // the search is O(np^2) in the worst case (no duplicates): bounded by skipping it for more than 128 spheres
    int pm = 0;  // 1: pairs with opposite translations share a group (register-resident kernel; BHS_ASM_NOPM=1 turns it off)
    {
        const int dedupe = B <= 128 ? 1 : 0;
        pm = (B > 1 && asm_reg_shape(plan, B, nsys, nullptr, nullptr) && getenv("BHS_ASM_NOPM") == nullptr) ? 1 : 0;
        size_t sm = (size_t)np * plan->d * sizeof(double);
        if (!dedupe || sm > 96 * 1024) sm = 0;
        if (sm > 48 * 1024) cudaFuncSetAttribute(pair_rep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        pair_rep_kernel<<<(unsigned)((np + 127) / 128), 128, sm, st>>>(plan->d, B, dedupe, pm, w.tv, w.rep);
    }
    BHS_CHECK_LAUNCH();
    pair_group_kernel<<<1, 1024, 0, st>>>(B, plan->d, pm, w.tv, w.rep, w.uid, w.cursor, w.n_unique, w.grp_rep, w.grp_start,
                                          w.members);
    BHS_CHECK_LAUNCH();
    AsmArgs a;
    a.B = B; a.H = plan->H; a.H2 = plan->H2; a.L2 = plan->L2;
    a.b_lo = b_lo; a.b_hi = b_hi;
    a.n_unique = w.n_unique; a.grp_rep = w.grp_rep; a.grp_start = w.grp_start; a.members = w.members;
    a.tiles = plan->d_tiles; a.coef = plan->d_coef; a.cidx = plan->d_cidx; a.deg = plan->d_deg; a.deg2 = plan->d_deg2;
    a.Y2 = w.Y2; a.hp = w.hp; a.rowf = w.rowf; a.colf = w.colf; a.diag = w.diag;
    a.A = (cplx*)d_A; a.ld = ld; a.sys_stride = sys_stride;
    a.Su = w.Su; a.ucap = (int)(np - B);
    const int ntiles = plan->tiles_r * plan->tiles_c;
    if (nsys > 65535) return BHS_ERR_UNSUPPORTED;
    int stage_bytes = 0;
    if (asm_reg_shape(plan, B, nsys, &stage_bytes, nullptr)) {
        // register-resident kernel: S vectors of the distinct translations first (U is only known on the device: the
        // y-dimension strides over them)
        {
            int64_t gy = np - B < 1024 ? np - B : 1024;
            dim3 sgrid((unsigned)((plan->H2 + 127) / 128), (unsigned)gy, (unsigned)nsys);
            s_vectors_kernel<<<sgrid, 128, 0, st>>>(plan->H2, plan->L2, np, a.ucap, w.n_unique, w.grp_rep, plan->d_deg2, w.Y2,
                                                   w.hp, w.Su);
            BHS_CHECK_LAUNCH();
        }
        bhs_prof_end(BHS_PROF_ASM_PRE, 0.0, st);
        bhs_prof_begin(BHS_PROF_ASM_MAIN, st);
        // Every CTA first loads its tile into registers (and stages its factors), so longer CTAs amortise that better,
        // shorter ones balance the last wave better: between four and nine waves of CTAs (two resident per SM), the count
        // whose last wave is fullest.
        int64_t chunks = 1;
        {
            const double slots = 2.0 * bhs_sm_count(), per = (double)ntiles * nsys;
            double best = -1.0;
            const int64_t cmax = np - B;
            for (int64_t c = 1; c <= cmax && c <= 65535; ++c) {
                const double w = per * (double)c / slots;
                if (w > 9.0 && best >= 0.0) break;
                const double eff = w / ceil(w) - (w < 4.0 ? (4.0 - w) : 0.0);  // fewer than four waves only if nothing else fits
                if (eff > best + 0.02) { best = eff; chunks = c; }
            }
        }
        dim3 grid((unsigned)ntiles, (unsigned)chunks, (unsigned)nsys);
        const bool fsm = B <= ASM_FSM_MAXB;
        const size_t smem = (size_t)2 * stage_bytes + (fsm ? (size_t)B * (BHS_TILE_R + BHS_TILE_C) * sizeof(cplx) : 0);
        const int nt_max = plan->max_nt;
#define BHS_ASM_REG_LAUNCH2(NT, F)                                                                                     \
    do {                                                                                                               \
        cudaFuncSetAttribute(assemble_reg_kernel<NT, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        assemble_reg_kernel<NT, F><<<grid, ASM_THREADS, smem, st>>>(a, stage_bytes);                                   \
    } while (0)
#define BHS_ASM_REG_LAUNCH(NT)                                                                                         \
    do {                                                                                                               \
        if (fsm) BHS_ASM_REG_LAUNCH2(NT, true);                                                                        \
        else BHS_ASM_REG_LAUNCH2(NT, false);                                                                           \
    } while (0)
        if (nt_max <= 8) BHS_ASM_REG_LAUNCH(8);
        else if (nt_max <= 16) BHS_ASM_REG_LAUNCH(16);
        else if (nt_max <= 24) BHS_ASM_REG_LAUNCH(24);
        else BHS_ASM_REG_LAUNCH(32);
#undef BHS_ASM_REG_LAUNCH2
#undef BHS_ASM_REG_LAUNCH
        BHS_CHECK_LAUNCH();
        bhs_prof_end(BHS_PROF_ASM_MAIN, 16.0 * (double)(b_hi - b_lo) * plan->H * (double)N * nsys, st);
        return BHS_OK;
    }
    // shared-memory budget: SY window (worst case H2 entries) + resident coefficient layers
    const size_t budget = 200 * 1024;
    size_t sy_bytes = ((size_t)plan->max_sy_cnt * sizeof(cplx) + 127) & ~(size_t)127;  // largest window of any tile
    a.sy_global = 0;
    if (sy_bytes + (ASM_LAYER_COEF + ASM_LAYER_IDX) > budget) {
        a.sy_global = 1;
        sy_bytes = 128;
    }
    int nt_cap = (int)((budget - sy_bytes) / (ASM_LAYER_COEF + ASM_LAYER_IDX));
    a.nt_res = plan->max_nt < nt_cap ? (plan->max_nt > 0 ? plan->max_nt : 1) : nt_cap;
    size_t smem = (size_t)a.nt_res * (ASM_LAYER_COEF + ASM_LAYER_IDX) + sy_bytes;
    cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // The number of distinct translations U is only known on the device: the y-dimension strides over them, sized
    // for about four waves of CTAs (4 resident per SM) and never more than the off-diagonal pair count.
    int64_t chunks = (4 * 4 * bhs_sm_count() + (int64_t)ntiles * nsys - 1) / ((int64_t)ntiles * nsys);
    if (chunks < 1) chunks = 1;
    if (chunks > np - B) chunks = np - B > 0 ? np - B : 1;
    if (chunks > 65535) return BHS_ERR_UNSUPPORTED;
    dim3 grid((unsigned)ntiles, (unsigned)chunks, (unsigned)nsys);
    bhs_prof_end(BHS_PROF_ASM_PRE, 0.0, st);
    bhs_prof_begin(BHS_PROF_ASM_MAIN, st);
    {
        const int64_t hh = (int64_t)plan->H * plan->H;
        dim3 dgrid((unsigned)((hh + 255) / 256), (unsigned)(b_hi - b_lo), (unsigned)nsys);
        diag_blocks_kernel<<<dgrid, 256, 0, st>>>(B, plan->H, b_lo, w.diag, (cplx*)d_A, ld, sys_stride);
        BHS_CHECK_LAUNCH();
    }
    if (B > 1) {
        assemble_kernel<<<grid, ASM_THREADS, smem, st>>>(a);
        BHS_CHECK_LAUNCH();
    }
    bhs_prof_end(BHS_PROF_ASM_MAIN, 16.0 * (double)(b_hi - b_lo) * plan->H * (double)N * nsys, st);
    return BHS_OK;
}

extern "C" int bhs_assemble(const bhs_plan_t* plan, int B, int nsys, const double* d_centers, const double* d_radii,
                            const double* d_k, const double* d_k_im, const double* d_eta, const double* d_alpha,
                            const double* d_beta, double* d_A, int64_t ld, int64_t sys_stride, void* d_work,
                            void* stream) {
    return assemble_impl(plan, B, nsys, d_centers, d_radii, d_k, d_k_im, d_eta, d_alpha, d_beta, 0, B, d_A, ld, sys_stride,
                         d_work, stream);
}

// Block rows b in [b_lo, b_hi) only, written to a strip [(b_hi - b_lo) * H, ld] whose first row is (b_lo, h = 0): the
// unit of the multi-GPU "assembly block-row" sharding (each rank builds its strip, strips are all-gathered onto the
// rank that factorises).  Same workspace as bhs_assemble.
extern "C" int bhs_assemble_rows(const bhs_plan_t* plan, int B, int nsys, const double* d_centers, const double* d_radii,
                                 const double* d_k, const double* d_k_im, const double* d_eta, const double* d_alpha,
                                 const double* d_beta, int b_lo, int b_hi, double* d_A, int64_t ld, int64_t sys_stride,
                                 void* d_work, void* stream) {
    return assemble_impl(plan, B, nsys, d_centers, d_radii, d_k, d_k_im, d_eta, d_alpha, d_beta, b_lo, b_hi, d_A, ld,
                         sys_stride, d_work, stream);
}

extern "C" int64_t bhs_diag_coef_workspace(const bhs_plan_t* plan, int B, int nsys) {
    if (!plan || B <= 0 || nsys <= 0) return BHS_ERR_INVALID;
    return al256((int64_t)nsys * B * plan->n_end * 4 * sizeof(cplx)) + al256((int64_t)ball_radial_scratch_bytes(plan->d, plan->n_end));
}

extern "C" int bhs_diag_coef(const bhs_plan_t* plan, int B, int nsys, const double* d_radii, const double* d_k,
                             const double* d_k_im, const double* d_eta, const double* d_alpha, const double* d_beta,
                             double* d_out, void* d_work, void* stream) {
    if (!plan || B <= 0 || nsys <= 0 || !d_radii || !d_k || !d_out || !d_work) return BHS_ERR_INVALID;
    // needs only the radial table (and, for very high orders, its scratch): both in the caller's workspace
    cudaStream_t st = (cudaStream_t)stream;
    AsmWork w;
    w.rad = (double4*)d_work;
    w.scratch = ball_radial_scratch_bytes(plan->d, plan->n_end)
                    ? (double*)((unsigned char*)d_work + al256((int64_t)nsys * B * plan->n_end * 4 * sizeof(cplx)))
                    : nullptr;
    w.rowf = nullptr; w.colf = nullptr; w.diag = (cplx*)d_out;
    return run_factors(plan, B, nsys, d_radii, d_k, d_k_im, d_eta, d_alpha, d_beta, w, true, st);
}

// ---- K3: right-hand side ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) rhs_expand_kernel(int d, int B, int H, int Q, const cplx* __restrict__ g,
                                                         const double* __restrict__ centers,
                                                         const double* __restrict__ radii,
                                                         const double* __restrict__ k_in,
                                                         const double* __restrict__ k_in_im,
                                                         const double* __restrict__ dir, const cplx* __restrict__ alpha,
                                                         const cplx* __restrict__ beta, const double* __restrict__ qdirs,
                                                         const cplx* __restrict__ WY, cplx* __restrict__ out) {
    extern __shared__ __align__(16) cplx s_g[];  // [Q]
    const int b = blockIdx.y, s = blockIdx.z;
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
        cplx v;
        if (g) {
            v = g[((int64_t)s * Q + q) * B + b];
        } else {
            double k = k_in[s], rho = radii[b];
            double dy = 0.0, dc = 0.0;
            for (int a = 0; a < d; ++a) {
                dy += dir[a] * qdirs[(int64_t)a * Q + q];
                dc += dir[a] * centers[(int64_t)b * d + a];
            }
            const double kim = k_in_im ? k_in_im[s] : 0.0;
            const double sarg = rho * dy + dc;
            double ph = k * sarg, sn, co;
            sincos(ph, &sn, &co);
            cplx u = cmake(co, sn);
            if (kim != 0.0) u = cscale(u, exp(-kim * sarg));  // exp(i (k + i kim) d.x)
            cplx al = alpha ? alpha[b] : cmake(1.0, 0.0);
            cplx be = beta ? beta[b] : cmake(0.0, 0.0);
            // -alpha u - beta (i k d.y) u
            cplx t = cadd(al, cmul(be, cmake(-kim * dy, k * dy)));
            v = cmul(cmake(-t.x, -t.y), u);
        }
        s_g[q] = v;
    }
    __syncthreads();
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    double ar = 0.0, ai = 0.0;
    for (int q = 0; q < Q; ++q) {
        cplx w = WY[(int64_t)q * H + h], gv = s_g[q];
        ar = fma(gv.x, w.x, ar); ar = fma(-gv.y, w.y, ar);
        ai = fma(gv.x, w.y, ai); ai = fma(gv.y, w.x, ai);
    }
    out[((int64_t)s * B + b) * H + h] = cmake(ar, ai);
}

extern "C" int bhs_rhs_expand(const bhs_plan_t* plan, int B, int nsys, const double* d_g, const double* d_centers,
                              const double* d_radii, const double* d_k_in, const double* d_k_in_im, const double* d_dir,
                              const double* d_alpha, const double* d_beta, double* d_out, void* stream) {
    if (!plan || B <= 0 || nsys <= 0 || !d_out) return BHS_ERR_INVALID;
    if (!d_g && (!d_centers || !d_radii || !d_k_in || !d_dir)) return BHS_ERR_INVALID;
    if (nsys > 65535 || B > 65535) return BHS_ERR_UNSUPPORTED;
    size_t smem = (size_t)plan->Q * sizeof(cplx);
    if (smem > 200 * 1024) return BHS_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(rhs_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((plan->H + 127) / 128, B, nsys);
    bhs_prof_begin(BHS_PROF_RHS_EXPAND, (cudaStream_t)stream);
    rhs_expand_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(plan->d, B, plan->H, plan->Q, (const cplx*)d_g,
                                                                 d_centers, d_radii, d_k_in, d_k_in_im, d_dir,
                                                                 (const cplx*)d_alpha, (const cplx*)d_beta,
                                                                 plan->d_qdirs, plan->d_WY, (cplx*)d_out);
    bhs_prof_end(BHS_PROF_RHS_EXPAND, 0.0, (cudaStream_t)stream);
    BHS_CHECK_LAUNCH();
    return BHS_OK;
}
