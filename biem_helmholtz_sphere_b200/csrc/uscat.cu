// K6: scattered-field evaluation
//
//   u_s(x) = sum_b sum_h  density[b,h] SD_n(rho_b)  h_n(k |x - c_b|)  Y_h((x - c_b)^)     (_biem.py:822-977)
//
// 3-D fast path (uscat3d_kernel): one thread per field point, balls streamed through shared memory with
// 1-D TMA bulk copies (double buffered, mbarrier completion).  Per (point, ball): Hankel orders by upward
// recurrence in registers, Legendre part by the monic recurrence q_{n+1} = x q_n - beta q_{n-1} with the
// normalisation folded into the pre-scaled coefficients, +-m handled together, azimuth by rotation.
// 12 FP64 instructions per (n, +-m) pair, nothing but 16 B/point of HBM traffic: DFMA-pipe bound.
//
// Generic path (uscat_generic_kernel): one warp per point, node tables in shared memory -- any chain
// type / any n_end (2-D, 4-D, 3-D beyond the register-resident limit).
#include "harmonics.cuh"
#include "prof.h"
#include "radial.cuh"
#include "special.cuh"

#define US3D_THREADS 128
#define US3D_CB 2  // balls per TMA stage

struct UscatArgs {
    int d, L, H, B, flags;
    double k, eta;
    const double* x;  // [d][P]
    int64_t P;
    const double* centers;  // [B][d]
    const double* radii;    // [B]
    const cplx* coefg;      // generic: [B][H]
    const double* rec;      // 3-D: [B][4 + 4*npair] doubles
    const double* beta;     // 3-D: [npair]
    const int32_t* idx;     // [H][s]
    const int32_t* deg;     // [H]
    cplx* out;
};

// ---- coefficient preparation ------------------------------------------------------------------------
// generic: coefg[b][h] = density[b][h] * SD_{deg h}(rho_b) * (far ? (-i)^deg : 1)
__global__ void uscat_coef_generic_kernel(int d, int L, int H, int B, double k, double eta, int far,
                                          const double* __restrict__ radii, const double4* __restrict__ rad,
                                          const int32_t* __restrict__ deg, const cplx* __restrict__ density,
                                          cplx* __restrict__ coefg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * H) return;
    int b = (int)(i / H), h = (int)(i % H);
    int n = deg[h];
    double4 r = rad[(int64_t)b * L + n];
    cplx sd = sd_coef(d, k, eta, radii[b], r.x, r.y);
    cplx c = cmul(density[i], sd);
    if (far) c = cmul_ipow(c, -n);
    coefg[i] = c;
}
// 3-D records: per ball [c0, c1, c2, rho] then for (m, n>=m): (c_{n,+m}, c_{n,-m}) * SD_n * norm_{n,m}
__global__ void uscat_coef3d_kernel(int L, int B, double k, double eta, int far, const double* __restrict__ centers,
                                    const double* __restrict__ radii, const double4* __restrict__ rad,
                                    const double* __restrict__ norm, const cplx* __restrict__ density,
                                    double* __restrict__ rec) {
    const int npair = L * (L + 1) / 2;
    const int64_t stride = 4 + 4 * (int64_t)npair;
    int b = blockIdx.x;
    double* rb = rec + stride * b;
    if (threadIdx.x < 4) rb[threadIdx.x] = threadIdx.x < 3 ? centers[b * 3 + threadIdx.x] : radii[b];
    const int H = L * L;
    for (int e = threadIdx.x; e < npair; e += blockDim.x) {
        // decode (m, n) from the m-major running index
        int m = 0, off = 0;
        while (off + (L - m) <= e) { off += L - m; ++m; }
        int n = m + (e - off);
        double4 r = rad[(int64_t)b * L + n];
        cplx sd = cscale(sd_coef(3, k, eta, radii[b], r.x, r.y), norm[e]);
        if (far) sd = cmul_ipow(sd, -n);
        cplx cp = cmul(density[(int64_t)b * H + n * n + m], sd);
        cplx cm = (m == 0) ? cmake(0.0, 0.0) : cmul(density[(int64_t)b * H + n * n + 2 * n + 1 - m], sd);
        reinterpret_cast<double4*>(rb + 4)[e] = make_double4(cp.x, cp.y, cm.x, cm.y);
    }
}

// ---- 3-D fast kernel --------------------------------------------------------------------------------
#define US_STEP(N)                                                              \
    case N: {                                                                   \
        if (N >= L) break;                                                      \
        const double4 cc = recs[idx];                                           \
        const double bt = sbeta[idx];                                           \
        const double gr = hr[N] * q1, gi = hi[N] * q1;                          \
        tpr = fma(gr, cc.x, tpr); tpi = fma(gr, cc.y, tpi);                     \
        tmr = fma(gr, cc.z, tmr); tmi = fma(gr, cc.w, tmi);                     \
        tpr = fma(-gi, cc.y, tpr); tpi = fma(gi, cc.x, tpi);                    \
        tmr = fma(-gi, cc.w, tmr); tmi = fma(gi, cc.z, tmi);                    \
        const double qn = fma(ct, q1, -(bt * q0));                              \
        q0 = q1; q1 = qn; ++idx;                                                \
    }

template <int LMAX>
__global__ void __launch_bounds__(US3D_THREADS) uscat3d_kernel(UscatArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int L = a.L, B = a.B;
    const int npair = L * (L + 1) / 2;
    const int rec_doubles = 4 + 4 * npair;
    const uint32_t stage_bytes = (uint32_t)(US3D_CB * rec_doubles * sizeof(double));
    double* stage0 = reinterpret_cast<double*>(smem_raw);
    double* stage1 = stage0 + US3D_CB * rec_doubles;
    double* sbeta = stage1 + US3D_CB * rec_doubles;
    __shared__ __align__(8) uint64_t full[2];

    const int tid = threadIdx.x;
    for (int e = tid; e < npair; e += US3D_THREADS) sbeta[e] = a.beta[e];
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int nchunks = (B + US3D_CB - 1) / US3D_CB;
    if (tid == 0) {
        int nb = min(US3D_CB, B);
        uint32_t bytes = (uint32_t)(nb * rec_doubles * sizeof(double));
        mbar_expect_tx(&full[0], bytes);
        tma_load_1d(stage0, a.rec, bytes, &full[0]);
    }
    (void)stage_bytes;

    const int64_t p = (int64_t)blockIdx.x * US3D_THREADS + tid;
    const bool active = p < a.P;
    const double x0 = active ? a.x[p] : 1e3, x1 = active ? a.x[a.P + p] : 1e3, x2 = active ? a.x[2 * a.P + p] : 1e3;
    const bool far = a.flags & BHS_FLAG_FAR_FIELD, inner = a.flags & BHS_FLAG_INNER, per_ball = a.flags & BHS_FLAG_PER_BALL;
    const double k = a.k;
    double accr = 0.0, acci = 0.0;
    bool bad = false;

    for (int c = 0; c < nchunks; ++c) {
        if (tid == 0 && c + 1 < nchunks) {
            int nb = min(US3D_CB, B - (c + 1) * US3D_CB);
            uint32_t bytes = (uint32_t)(nb * rec_doubles * sizeof(double));
            fence_proxy_async();
            mbar_expect_tx(&full[(c + 1) & 1], bytes);
            tma_load_1d(((c + 1) & 1) ? stage1 : stage0, a.rec + (int64_t)(c + 1) * US3D_CB * rec_doubles, bytes,
                        &full[(c + 1) & 1]);
        }
        mbar_wait(&full[c & 1], (c >> 1) & 1);
        const double* st = (c & 1) ? stage1 : stage0;
        const int nb = min(US3D_CB, B - c * US3D_CB);
        for (int bb = 0; bb < nb; ++bb) {
            const double* rb = st + bb * rec_doubles;
            const double4* recs = reinterpret_cast<const double4*>(rb + 4);
            const double dx0 = x0 - rb[0], dx1 = x1 - rb[1], dx2 = x2 - rb[2], rho = rb[3];
            const double rxy2 = dx1 * dx1 + dx2 * dx2;
            const double r = sqrt(dx0 * dx0 + rxy2);
            bad = bad || (inner ? (r > rho) : (r < rho));
            double ct = 1.0, sn = 0.0, cp = 1.0, sp = 0.0;
            const double ir = 1.0 / r;
            if (r > 0.0) {
                const double rxy = sqrt(rxy2);
                ct = dx0 * ir;
                sn = rxy * ir;
                if (rxy > 0.0) {
                    const double irxy = 1.0 / rxy;
                    cp = dx1 * irxy;
                    sp = dx2 * irxy;
                }
            }
            // radial part: h_n(kr) upward (far field: the (-i)^n is folded into the coefficients)
            double hr[LMAX], hi[LMAX];
            if (far) {
#pragma unroll
                for (int n = 0; n < LMAX; ++n) { hr[n] = 1.0; hi[n] = 0.0; }
            } else {
                const double z = k * r, iz = 1.0 / z;
                double s, co;
                sincos(z, &s, &co);
                hr[0] = s * iz; hi[0] = -co * iz;
                if (LMAX > 1) { hr[1] = (s * iz - co) * iz; hi[1] = (-co * iz - s) * iz; }
#pragma unroll
                for (int n = 1; n < LMAX - 1; ++n) {
                    if (n + 1 < L) {
                        const double cf = (2 * n + 1) * iz;
                        hr[n + 1] = fma(cf, hr[n], -hr[n - 1]);
                        hi[n + 1] = fma(cf, hi[n], -hi[n - 1]);
                    } else {
                        hr[n + 1] = 0.0; hi[n + 1] = 0.0;
                    }
                }
            }
            double br = 0.0, bi = 0.0;   // this ball's sum
            double qmm = 1.0, cm = 1.0, sm = 0.0;
            int idx = 0;
            for (int m = 0; m < L; ++m) {
                double q0 = 0.0, q1 = qmm;
                double tpr = 0.0, tpi = 0.0, tmr = 0.0, tmi = 0.0;
                switch (m) {
                    US_STEP(0) US_STEP(1) US_STEP(2) US_STEP(3) US_STEP(4) US_STEP(5) US_STEP(6) US_STEP(7)
#if 1
                    default: break;
#endif
                }
                if (LMAX > 8) {
                    switch (m < 8 ? 8 : m) {
                        US_STEP(8) US_STEP(9) US_STEP(10) US_STEP(11) US_STEP(12) US_STEP(13) US_STEP(14) US_STEP(15)
                        default: break;
                    }
                }
                if (LMAX > 16) {
                    switch (m < 16 ? 16 : m) {
                        US_STEP(16) US_STEP(17) US_STEP(18) US_STEP(19) US_STEP(20) US_STEP(21) US_STEP(22) US_STEP(23)
                        default: break;
                    }
                }
                if (LMAX > 24) {
                    switch (m < 24 ? 24 : m) {
                        US_STEP(24) US_STEP(25) US_STEP(26) US_STEP(27) US_STEP(28) US_STEP(29) US_STEP(30) US_STEP(31)
                        default: break;
                    }
                }
                // (T+ e^{i m phi} + T- e^{-i m phi})
                br += (tpr + tmr) * cm - (tpi - tmi) * sm;
                bi += (tpi + tmi) * cm + (tpr - tmr) * sm;
                const double cn = cm * cp - sm * sp;
                sm = sm * cp + cm * sp;
                cm = cn;
                qmm *= sn;
            }
            if (far) {
                // (ik)^{-1} exp(-i k x.c_b), x as given (_biem.py:931-944)
                double ph = -k * (x0 * rb[0] + x1 * rb[1] + x2 * rb[2]);
                double s, co;
                sincos(ph, &s, &co);
                // (br + i bi) * (co + i s) * (-i / k)
                double tr = br * co - bi * s, ti = br * s + bi * co;
                br = ti / k;
                bi = -tr / k;
            }
            if (per_ball) {
                if (active) a.out[p * B + (c * US3D_CB + bb)] = cmake(br, bi);
            } else {
                accr += br;
                acci += bi;
            }
        }
        __syncthreads();
    }
    if (!active) return;
    if (!far && bad) {
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        if (per_ball) {
            for (int b = 0; b < B; ++b) a.out[p * B + b] = cmake(qnan, 0.0);
        } else {
            a.out[p] = cmake(qnan, 0.0);
        }
        return;
    }
    if (!per_ball) a.out[p] = cmake(accr, acci);
}

// ---- generic kernel: one warp per point ---------------------------------------------------------------
struct SmArr {
    double* p;
    __device__ __forceinline__ double& operator[](int n) const { return p[n]; }
};

__global__ void __launch_bounds__(128) uscat_generic_kernel(UscatArgs a, HarmTables tb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = a.d, L = a.L, s = d - 1;
    const int shift = (d & 1) ? (d - 3) / 2 : d / 2 - 1;
    const size_t harm_bytes = harm_smem_bytes_per_warp(d, L);
    const size_t rad_bytes = (size_t)2 * (L + 2 + shift) * sizeof(double);
    unsigned char* base = smem_raw + (harm_bytes + rad_bytes) * warp;
    double* F;
    cplx* E;
    harm_smem_carve(base, L, F, E);
    double* Hr = reinterpret_cast<double*>(base + harm_bytes);
    double* Hi = Hr + (L + 2 + shift);
    const bool far = a.flags & BHS_FLAG_FAR_FIELD, inner = a.flags & BHS_FLAG_INNER, per_ball = a.flags & BHS_FLAG_PER_BALL;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t p = (int64_t)blockIdx.x * warps + warp; p < a.P; p += (int64_t)gridDim.x * warps) {
        double x[BHS_MAX_NODES + 2], dx[BHS_MAX_NODES + 2];
        for (int i = 0; i < d; ++i) x[i] = a.x[(int64_t)i * a.P + p];
        double accr = 0.0, acci = 0.0;
        bool bad = false;
        for (int b = 0; b < a.B; ++b) {
            double r2 = 0.0, dot = 0.0;
            for (int i = 0; i < d; ++i) {
                double c = a.centers[(int64_t)b * d + i];
                dx[i] = x[i] - c;
                r2 += dx[i] * dx[i];
                dot += x[i] * c;
            }
            double r = sqrt(r2), rho = a.radii[b];
            bad = bad || (inner ? (r > rho) : (r < rho));
            warp_harmonic_tables(tb, L, dx, F, E, lane);
            if (lane == 0) {
                if (far) {
                    for (int n = 0; n < L; ++n) { Hr[n] = 1.0; Hi[n] = 0.0; }
                } else {
                    hankel_upward(d, a.k * r, L - 1, SmArr{Hr}, SmArr{Hi});
                }
            }
            __syncwarp();
            double pr = 0.0, pi = 0.0;
            for (int h = lane; h < a.H; h += 32) {
                cplx y = harmonic_from_tables(tb, L, a.idx + (int64_t)h * s, F, E);
                int n = a.deg[h];
                cplx t = cmul(cmake(Hr[n], Hi[n]), y);
                cplx c = a.coefg[(int64_t)b * a.H + h];
                pr += t.x * c.x - t.y * c.y;
                pi += t.x * c.y + t.y * c.x;
            }
            if (far) {
                // (ik)^{-(d-1)/2} exp(-i k x.c_b): k^{-p} e^{-i pi p / 2}, p = (d-1)/2
                double pw = 0.5 * (d - 1);
                double ang = -a.k * dot - 1.57079632679489661923 * pw;
                double sn, co;
                sincos(ang, &sn, &co);
                double mag = pow(a.k, -pw);
                double tr = (pr * co - pi * sn) * mag, ti = (pr * sn + pi * co) * mag;
                pr = tr; pi = ti;
            }
            if (per_ball) {
                for (int o = 16; o > 0; o >>= 1) {
                    pr += __shfl_xor_sync(0xffffffffu, pr, o);
                    pi += __shfl_xor_sync(0xffffffffu, pi, o);
                }
                if (lane == 0) a.out[p * a.B + b] = cmake(pr, pi);
            } else {
                accr += pr;
                acci += pi;
            }
            __syncwarp();
        }
        if (!per_ball) {
            for (int o = 16; o > 0; o >>= 1) {
                accr += __shfl_xor_sync(0xffffffffu, accr, o);
                acci += __shfl_xor_sync(0xffffffffu, acci, o);
            }
        }
        if (!far && bad) {
            if (per_ball) {
                for (int b = lane; b < a.B; b += 32) a.out[p * a.B + b] = cmake(qnan, 0.0);
            } else if (lane == 0) {
                a.out[p] = cmake(qnan, 0.0);
            }
        } else if (!per_ball && lane == 0) {
            a.out[p] = cmake(accr, acci);
        }
    }
}

// ---- host entry ---------------------------------------------------------------------------------------
static inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }

extern "C" int64_t bhs_uscat_workspace(const bhs_plan_t* plan, int B) {
    if (!plan || B <= 0) return BHS_ERR_INVALID;
    int64_t L = plan->n_end, npair = L * (L + 1) / 2;
    int64_t rad = align256((int64_t)B * L * sizeof(double4));
    int64_t c3 = align256((int64_t)B * (4 + 4 * npair) * sizeof(double));
    int64_t cg = align256((int64_t)B * plan->H * sizeof(cplx));
    int64_t kbuf = 256;
    return rad + (c3 > cg ? c3 : cg) + kbuf;
}

template <int LMAX>
static int launch_uscat3d(const UscatArgs& a, cudaStream_t st) {
    const int npair = a.L * (a.L + 1) / 2;
    size_t smem = (size_t)(2 * US3D_CB * (4 + 4 * npair) + npair) * sizeof(double);
    cudaFuncSetAttribute(uscat3d_kernel<LMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = (a.P + US3D_THREADS - 1) / US3D_THREADS;
    bhs_prof_begin(BHS_PROF_USCAT, st);
    uscat3d_kernel<LMAX><<<(unsigned)blocks, US3D_THREADS, smem, st>>>(a);
    bhs_prof_end(BHS_PROF_USCAT, 8.0 * (double)a.P * a.B * a.H, st);
    BHS_CHECK_LAUNCH();
    return BHS_OK;
}

extern "C" int bhs_uscat(const bhs_plan_t* plan, int B, const double* d_centers, const double* d_radii, double k,
                         double eta, const double* d_density, const double* d_x, int64_t P, int flags, double* d_out,
                         void* d_work, void* stream) {
    if (!plan || B <= 0 || !d_centers || !d_radii || !d_density || !d_x || !d_out || !d_work || P < 0)
        return BHS_ERR_INVALID;
    if (!(k > 0.0)) return BHS_ERR_UNSUPPORTED;
    if (P == 0) return BHS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int d = plan->d, L = plan->n_end, H = plan->H;
    const int64_t npair = (int64_t)L * (L + 1) / 2;
    unsigned char* w = (unsigned char*)d_work;
    double4* d_rad = (double4*)(w + 256);
    unsigned char* d_coef = w + 256 + align256((int64_t)B * L * sizeof(double4));
    int rc = launch_ball_radial(d, L, B, 1, d_radii, nullptr, k, d_rad, st);
    if (rc) return rc;
    const int far = (flags & BHS_FLAG_FAR_FIELD) ? 1 : 0;
    UscatArgs a;
    a.d = d; a.L = L; a.H = H; a.B = B; a.flags = flags; a.k = k; a.eta = eta;
    a.x = d_x; a.P = P; a.centers = d_centers; a.radii = d_radii;
    a.coefg = nullptr; a.rec = nullptr; a.beta = plan->d_us_beta; a.idx = plan->d_idx; a.deg = plan->d_deg;
    a.out = (cplx*)d_out;
    if (d == 3 && L <= 32) {
        uscat_coef3d_kernel<<<B, 128, 0, st>>>(L, B, k, eta, far, d_centers, d_radii, d_rad, plan->d_us_norm,
                                               (const cplx*)d_density, (double*)d_coef);
        BHS_CHECK_LAUNCH();
        a.rec = (const double*)d_coef;
        (void)npair;
        if (L <= 8) return launch_uscat3d<8>(a, st);
        if (L <= 16) return launch_uscat3d<16>(a, st);
        if (L <= 24) return launch_uscat3d<24>(a, st);
        return launch_uscat3d<32>(a, st);
    }
    int64_t tot = (int64_t)B * H;
    uscat_coef_generic_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d, L, H, B, k, eta, far, d_radii, d_rad,
                                                                            plan->d_deg, (const cplx*)d_density,
                                                                            (cplx*)d_coef);
    BHS_CHECK_LAUNCH();
    a.coefg = (const cplx*)d_coef;
    const int warps = 4;
    const int shift = (d & 1) ? (d - 3) / 2 : d / 2 - 1;
    size_t smem = (harm_smem_bytes_per_warp(d, L) + (size_t)2 * (L + 2 + shift) * sizeof(double)) * warps;
    if (smem > 200 * 1024) return BHS_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(uscat_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = (P + warps - 1) / warps;
    if (blocks > 148 * 16) blocks = 148 * 16;
    bhs_prof_begin(BHS_PROF_USCAT, st);
    uscat_generic_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(a, harm_tables_of(plan));
    bhs_prof_end(BHS_PROF_USCAT, 8.0 * (double)a.P * a.B * a.H, st);
    BHS_CHECK_LAUNCH();
    return BHS_OK;
}
