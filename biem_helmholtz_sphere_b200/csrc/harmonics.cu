// K2: orthonormal harmonics on S^{d-1} for chain coordinate trees, ultrasphere (Phase(0)) ordering.
//
//   Y_h = prod_i f^{(desc_i)}_{n_i, n_{i+1}}(theta_i) * e^{i m phi} / sqrt(2 pi)        
//
// One warp per direction: lanes build the node-function tables column-wise in shared memory
// (three-term recurrences, coefficients from the plan), then stream the flattened harmonics out with
// coalesced 16-byte stores.  Replaces ush.harmonics (_biem.py:922) and the harmonics inside
// ush.expand / ush.harmonics_translation_coef (_biem.py:627,697).
#include <vector>

#include "harmonics.cuh"

__global__ void __launch_bounds__(128) harmonics_kernel(HarmTables tb, int Lb, const int32_t* __restrict__ idx,
                                                        int Hb, const double* __restrict__ xyz, int64_t npts,
                                                        const double* __restrict__ scale, int conj_out,
                                                        cplx* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = harm_smem_bytes_per_warp(tb.d, Lb);
    double* F;
    cplx* E;
    harm_smem_carve(smem_raw + per_warp * warp, Lb, F, E);
    const int s = tb.d - 1;
    for (int64_t p = (int64_t)blockIdx.x * warps + warp; p < npts; p += (int64_t)gridDim.x * warps) {
        double x[BHS_MAX_NODES + 2];
        for (int i = 0; i < tb.d; ++i) x[i] = xyz[(int64_t)i * npts + p];
        warp_harmonic_tables(tb, Lb, x, F, E, lane);
        __syncwarp();
        double sc = scale ? scale[p] : 1.0;
        for (int h = lane; h < Hb; h += 32) {
            cplx y = harmonic_from_tables(tb, Lb, idx + (int64_t)h * s, F, E);
            if (conj_out) y.y = -y.y;
            out[p * Hb + h] = cscale(y, sc);
        }
        __syncwarp();
    }
}

// ---- 3-D specialisation ------------------------------------------------------------------------------------------
// Same tables, same recurrences and the same products as the generic kernel (so the values are identical), organised for
// throughput: a HALF-warp builds the tables of one direction (lane l runs the degree recurrence of the node functions f_{n,l},
// n >= l, and the power e^{i l phi}), i.e. a warp prepares two directions at once with every lane busy; the whole warp then
// streams each direction's H values out, 512 contiguous bytes per store instruction.  The (n, |m|, sign) of the harmonics a
// lane stores are direction-independent: they are unpacked once into NH registers per lane (NH = 0: read per use).
// The generic kernel spent ~1350 instructions per direction (run-time tree depth, 64-bit index arithmetic, local-memory
// coordinate arrays); this one ~300.
template <int NH>
__global__ void __launch_bounds__(128) harmonics3d_kernel(HarmTables tb, int Lb, const int32_t* __restrict__ idx, int Hb,
                                                          const double* __restrict__ xyz, int64_t npts,
                                                          const double* __restrict__ scale, int conj_out,
                                                          cplx* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = lane >> 4, hl = lane & 15;
    const size_t per_dir = harm_smem_bytes_per_warp(3, Lb);
    unsigned char* wbase = smem_raw + per_dir * 2 * warp;
    double* F;
    cplx* E;
    harm_smem_carve(wbase + per_dir * half, Lb, F, E);
    // the recurrence coefficients of the band, staged once per CTA ([n][l], l fastest: conflict-free for lane = l)
    double* s_c1 = reinterpret_cast<double*>(smem_raw + per_dir * 2 * warps);
    double* s_c2 = s_c1 + Lb * Lb;
    for (int e = threadIdx.x; e < Lb * Lb; e += blockDim.x) {
        const int n = e / Lb, l = e - n * Lb;
        s_c1[e] = tb.c1[(size_t)n * tb.L2 + l];
        s_c2[e] = tb.c2[(size_t)n * tb.L2 + l];
    }
    __syncthreads();
    // packed (F offset | |m| << 12 | negative-m << 20) of this lane's harmonics
    int pk[NH > 0 ? NH : 1];
    auto pack = [&](int h) {
        const int2 nm = *reinterpret_cast<const int2*>(idx + (int64_t)h * 2);
        const int am = nm.y < 0 ? -nm.y : nm.y;
        return (nm.x * Lb + am) | (am << 12) | ((nm.y < 0) << 20);
    };
    if (NH > 0) {
#pragma unroll
        for (int j = 0; j < NH; ++j) pk[j] = (lane + 32 * j < Hb) ? pack(lane + 32 * j) : 0;
    }
    const double inv_s2pi = 0.39894228040143267794;
    for (int64_t p0 = 2 * ((int64_t)blockIdx.x * warps + warp); p0 < npts; p0 += 2 * (int64_t)gridDim.x * warps) {
        const int64_t p = p0 + half;
        if (p < npts) {
            const double x0 = xyz[p], x1 = xyz[npts + p], x2 = xyz[2 * npts + p];
            // tails as in warp_harmonic_tables: accumulated from the last coordinate
            double acc = x2 * x2;
            acc += x1 * x1;
            const double t1 = sqrt(acc);
            acc += x0 * x0;
            const double t0 = sqrt(acc);
            double ct = 1.0, st = 0.0;
            if (t0 > 0.0) {
                ct = x0 / t0;
                st = t1 / t0;
            }
            double cp = 1.0, sp = 0.0;
            if (t1 > 0.0) {
                cp = x1 / t1;
                sp = x2 / t1;
            }
            for (int l = hl; l < Lb; l += 16) {
                const double f0 = tb.all[l] * powi_d(st, l);
                F[l * Lb + l] = f0;
                double fm2 = 0.0, fm1 = f0;
                for (int n = l + 1; n < Lb; ++n) {
                    const double f = s_c1[n * Lb + l] * ct * fm1 - s_c2[n * Lb + l] * fm2;
                    F[n * Lb + l] = f;
                    fm2 = fm1;
                    fm1 = f;
                }
                cplx r = cmake(1.0, 0.0), b = cmake(cp, sp);
                int e = l;
                while (e) {
                    if (e & 1) r = cmul(r, b);
                    b = cmul(b, b);
                    e >>= 1;
                }
                E[l] = cscale(r, inv_s2pi);
            }
        }
        __syncwarp();
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            const int64_t pq = p0 + q;
            if (pq >= npts) break;
            const double* Fq;
            const cplx* Eq;
            {
                double* f_;
                cplx* e_;
                harm_smem_carve(wbase + per_dir * q, Lb, f_, e_);
                Fq = f_;
                Eq = e_;
            }
            const double sc = scale ? scale[pq] : 1.0;
            cplx* o = out + pq * Hb;
            if (NH > 0) {
#pragma unroll
                for (int j = 0; j < NH; ++j) {
                    const int h = lane + 32 * j;
                    if (h < Hb) {
                        const int w = pk[j];
                        cplx e = Eq[(w >> 12) & 0xff];
                        if ((w >> 20) != conj_out) e.y = -e.y;  // conj for negative m, once more for a conjugated output
                        const double prod = 1.0 * Fq[w & 0xfff];
                        o[h] = cscale(cscale(e, prod), sc);
                    }
                }
            } else {
                for (int h = lane; h < Hb; h += 32) {
                    const int w = pack(h);
                    cplx e = Eq[(w >> 12) & 0xff];
                    if ((w >> 20) != conj_out) e.y = -e.y;
                    const double prod = 1.0 * Fq[w & 0xfff];
                    o[h] = cscale(cscale(e, prod), sc);
                }
            }
        }
        __syncwarp();
    }
}

static int launch_harmonics(const bhs_plan* plan, int Lb, const int32_t* d_idx, int Hb, const double* d_xyz,
                            int64_t npts, const double* d_scale, int conj_out, cplx* d_out, cudaStream_t st) {
    if (npts <= 0) return BHS_OK;
    HarmTables tb = harm_tables_of(plan);
    if (plan->d == 3 && Lb <= 64 && getenv("BHS_HARM_GENERIC") == nullptr) {  // F offsets are packed in 12 bits, |m| in 8
        int warps = 4;
        const size_t coef_bytes = (size_t)2 * Lb * Lb * sizeof(double);
        while (warps > 1 && harm_smem_bytes_per_warp(3, Lb) * 2 * warps + coef_bytes > 200 * 1024) warps >>= 1;
        const size_t smem = harm_smem_bytes_per_warp(3, Lb) * 2 * warps + coef_bytes;
        int64_t blocks = (npts + 2 * warps - 1) / (2 * warps);
        const int per_sm = (int)((200 * 1024) / (smem + 1024)) < 8 ? (int)((200 * 1024) / (smem + 1024)) : 8;
        if (blocks > (int64_t)bhs_sm_count() * per_sm) blocks = (int64_t)bhs_sm_count() * per_sm;
        if (smem <= 200 * 1024 && per_sm >= 1) {
#define BHS_H3_LAUNCH(NH)                                                                                       \
    do {                                                                                                        \
        cudaFuncSetAttribute(harmonics3d_kernel<NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        harmonics3d_kernel<NH><<<(unsigned)blocks, warps * 32, smem, st>>>(tb, Lb, d_idx, Hb, d_xyz, npts, d_scale, \
                                                                          conj_out, d_out);                     \
    } while (0)
            if (Hb <= 256) BHS_H3_LAUNCH(8);
            else if (Hb <= 1024) BHS_H3_LAUNCH(32);
            else BHS_H3_LAUNCH(0);
#undef BHS_H3_LAUNCH
            BHS_CHECK_LAUNCH();
            return BHS_OK;
        }
    }
    int warps = 4;
    while (warps > 1 && harm_smem_bytes_per_warp(plan->d, Lb) * warps > 200 * 1024) warps >>= 1;
    size_t smem = harm_smem_bytes_per_warp(plan->d, Lb) * warps;
    if (smem > 200 * 1024) return BHS_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(harmonics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = (npts + warps - 1) / warps;
    if (blocks > bhs_sm_count() * 16) blocks = bhs_sm_count() * 16;
    harmonics_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(tb, Lb, d_idx, Hb, d_xyz, npts, d_scale, conj_out,
                                                                 d_out);
    BHS_CHECK_LAUNCH();
    return BHS_OK;
}

extern "C" int bhs_harmonics(const bhs_plan_t* plan, int use_double_band, const double* d_xyz, int64_t npts,
                             double* d_out, void* stream) {
    if (!plan || !d_xyz || !d_out || npts < 0) return BHS_ERR_INVALID;
    if (use_double_band)
        return launch_harmonics(plan, plan->L2, plan->d_idx2, plan->H2, d_xyz, npts, nullptr, 0, (cplx*)d_out,
                                (cudaStream_t)stream);
    return launch_harmonics(plan, plan->n_end, plan->d_idx, plan->H, d_xyz, npts, nullptr, 0, (cplx*)d_out,
                            (cudaStream_t)stream);
}

// WY[q][h] = w_q conj(Y_h(y_q)) on the RHS quadrature nodes (called once from bhs_plan_create)
int bhs_fill_WY(bhs_plan* p) {
    return launch_harmonics(p, p->n_end, p->d_idx, p->H, p->d_qdirs, p->Q, p->d_qw, 1, p->d_WY, 0);
}

// ---- tables of the planar field kernel (3-D, n_end <= 32; called once from bhs_plan_create) ------------------------------
// In the frame (x', y', z') = (x0, x1, x2) read as (azimuth-cos, azimuth-sin, POLAR) every direction inside the plane
// x2 = const sits on the equator, where Y'_{n,m} = K_{n,m} e^{i m phi'} with K = 0 for n + |m| odd.  Y' is the chain harmonic
// evaluated at the permuted vector (x2, x0, x1), so the rotation between the two bases is a quadrature away: the plan's
// right-hand-side rule (n_end Gauss-Legendre x 2 n_end equispaced nodes) is exact for products of two harmonics of degree
// < n_end.   rot[n][m'][m] = conj( sum_q Y'_{n,m'}(y_q) w_q conj Y_{n,m}(y_q) ),   c'_{n,m'} = sum_m rot[n][m'][m] c_{n,m}.
__global__ void planar_rot_kernel(int L, int H, int Q, const cplx* __restrict__ Yp, const cplx* __restrict__ WY,
                                  cplx* __restrict__ rot) {
    const int n = blockIdx.x, w = 2 * n + 1;
    const int64_t off = ((int64_t)4 * n * n * n - n) / 3;
    for (int e = threadIdx.x; e < w * w; e += blockDim.x) {
        const int r = e / w, c = e % w;
        double sr = 0.0, si = 0.0;
        for (int q = 0; q < Q; ++q) {
            const cplx a = Yp[(int64_t)q * H + n * n + r], b = WY[(int64_t)q * H + n * n + c];
            sr += a.x * b.x - a.y * b.y;
            si += a.x * b.y + a.y * b.x;
        }
        rot[off + e] = cmake(sr, -si);
    }
    (void)L;
}
__global__ void planar_K_kernel(int H, const cplx* __restrict__ Yeq, double* __restrict__ K) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h < H) K[h] = Yeq[h].x;
}

int bhs_fill_planar_tables(bhs_plan* p) {
    p->d_us_rot = nullptr;
    p->d_us_K = nullptr;
    if (p->d != 3 || p->n_end > 32 || p->tree != BHS_TREE_CHAIN) return BHS_OK;
    const int L = p->n_end, H = p->H, Q = p->Q;
    // quadrature directions in the permuted frame (polar <- x2, azimuth-cos <- x0, azimuth-sin <- x1), then one equator point
    std::vector<double> dirs((size_t)3 * (Q + 1));
    const int64_t NP = Q + 1;
    for (int q = 0; q < Q; ++q) {
        dirs[0 * NP + q] = p->h_qdirs[(size_t)2 * Q + q];
        dirs[1 * NP + q] = p->h_qdirs[(size_t)0 * Q + q];
        dirs[2 * NP + q] = p->h_qdirs[(size_t)1 * Q + q];
    }
    dirs[0 * NP + Q] = 0.0; dirs[1 * NP + Q] = 1.0; dirs[2 * NP + Q] = 0.0;
    double* d_dirs = nullptr;
    cplx* d_Y = nullptr;
    const int64_t nrot = ((int64_t)4 * L * L * L - L) / 3;
    int rc = BHS_OK;
    if (cudaMalloc((void**)&d_dirs, dirs.size() * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&d_Y, (size_t)NP * H * sizeof(cplx)) != cudaSuccess ||
        cudaMalloc((void**)&p->d_us_rot, (size_t)nrot * sizeof(cplx)) != cudaSuccess ||
        cudaMalloc((void**)&p->d_us_K, (size_t)H * sizeof(double)) != cudaSuccess)
        rc = BHS_ERR_ALLOC;
    if (rc == BHS_OK && cudaMemcpy(d_dirs, dirs.data(), dirs.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
        rc = BHS_ERR_ALLOC;
    if (rc == BHS_OK) rc = launch_harmonics(p, L, p->d_idx, H, d_dirs, NP, nullptr, 0, d_Y, 0);
    if (rc == BHS_OK) {
        planar_rot_kernel<<<L, 256>>>(L, H, Q, d_Y, p->d_WY, p->d_us_rot);
        planar_K_kernel<<<(H + 127) / 128, 128>>>(H, d_Y + (int64_t)Q * H, p->d_us_K);
        if (cudaDeviceSynchronize() != cudaSuccess) rc = (int)cudaGetLastError();
    }
    cudaFree(d_dirs);
    cudaFree(d_Y);
    return rc;
}
