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
    double k, k_im, eta;  // wavenumber k + i k_im (k_im != 0: 3-D only)
    const double* x;  // [d][P]
    int64_t P;
    const double* centers;  // [B][d]
    const double* radii;    // [B]
    const cplx* coefg;      // generic: [B][H]
    const double* rec;      // 3-D: [B][4 + 4*npair] doubles ([B][4 + 2*npair] for the planar variant)
    const int* planar;      // 3-D: device flag, 1 = every point and every centre share the same x2 (see uscat_planar_check_kernel)
    const double* rec3;     // 3-D planar kernel: [B][4 + 4 * planar_pairs(LMAX)] doubles
    const double* beta;     // 3-D: [npair]
    const int32_t* idx;     // [H][s]
    const int32_t* deg;     // [H]
    cplx* out;
};

// ---- coefficient preparation ------------------------------------------------------------------------
// generic: coefg[b][h] = density[b][h] * SD_{deg h}(rho_b) * (far ? (-i)^deg : 1)
// radz != nullptr: complex wavenumber, table (j, j', h, h') complex from launch_ball_radial_z
__global__ void uscat_coef_generic_kernel(int d, int L, int H, int B, double k, double k_im, double eta, int far,
                                          const double* __restrict__ radii, const double4* __restrict__ rad,
                                          const cplx* __restrict__ radz, const int32_t* __restrict__ deg,
                                          const cplx* __restrict__ density, cplx* __restrict__ coefg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * H) return;
    int b = (int)(i / H), h = (int)(i % H);
    int n = deg[h];
    cplx sd;
    if (radz) {
        const cplx* rz = radz + ((int64_t)b * L + n) * 4;
        sd = sd_coef_z(d, cmake(k, k_im), eta, radii[b], rz[0], rz[1]);
    } else {
        double4 r = rad[(int64_t)b * L + n];
        sd = sd_coef(d, k, eta, radii[b], r.x, r.y);
    }
    cplx c = cmul(density[i], sd);
    if (far) c = cmul_ipow(c, -n);
    coefg[i] = c;
}
// 3-D records, m-major over the PADDED band LMAX (multiple of 8, >= L): per ball [c0, c1, c2, rho] then for
// (m, n = m..LMAX-1): (A, B) = (c_{n,+m} + c_{n,-m}, i (c_{n,+m} - c_{n,-m})) * SD_n * norm_{n,m}, zero for n >= L.  Block 0 also writes the
// recurrence coefficients beta_{n,m} = (n^2 - m^2) / (4 n^2 - 1) in the same (m, n) order.
__global__ void uscat_coef3d_kernel(int L, int LMAX, int B, double k, double k_im, double eta, int far,
                                    const double* __restrict__ centers, const double* __restrict__ radii,
                                    const double4* __restrict__ rad, const cplx* __restrict__ radz,
                                    const double* __restrict__ norm,
                                    const cplx* __restrict__ density, double* __restrict__ rec,
                                    double* __restrict__ beta, double* __restrict__ rec2, int* __restrict__ planar) {
    const int npair = LMAX * (LMAX + 1) / 2;
    const int64_t stride = 4 + 4 * (int64_t)npair, stride2 = 4 + 2 * (int64_t)npair;
    int b = blockIdx.x;
    double* rb = rec + stride * b;
    double* rb2 = rec2 ? rec2 + stride2 * b : nullptr;
    if (threadIdx.x < 4) {
        rb[threadIdx.x] = threadIdx.x < 3 ? centers[b * 3 + threadIdx.x] : radii[b];
        if (rb2) rb2[threadIdx.x] = rb[threadIdx.x];
    }
    if (b == 0 && threadIdx.x == 0) planar[0] = 1;  // cleared by uscat_planar_check_kernel on the first mismatch
    const int H = L * L;
    for (int e = threadIdx.x; e < npair; e += blockDim.x) {
        // decode (m, n) from the m-major running index
        int m = 0, off = 0;
        while (off + (LMAX - m) <= e) { off += LMAX - m; ++m; }
        int n = m + (e - off);
        if (b == 0) beta[e] = (double)(n * n - m * m) / (double)(4 * n * n - 1);
        double4 out = make_double4(0.0, 0.0, 0.0, 0.0);
        if (n < L) {
            const int eL = m * L - m * (m - 1) / 2 + (n - m);
            cplx sd;
            if (radz) {
                const cplx* rz = radz + ((int64_t)b * L + n) * 4;
                sd = cscale(sd_coef_z(3, cmake(k, k_im), eta, radii[b], rz[0], rz[1]), norm[eL]);
            } else {
                double4 r = rad[(int64_t)b * L + n];
                sd = cscale(sd_coef(3, k, eta, radii[b], r.x, r.y), norm[eL]);
            }
            if (far) sd = cmul_ipow(sd, -n);
            cplx cp = cmul(density[(int64_t)b * H + n * n + m], sd);
            cplx cm = (m == 0) ? cmake(0.0, 0.0) : cmul(density[(int64_t)b * H + n * n + 2 * n + 1 - m], sd);
            // A = c+ + c-,  B = i (c+ - c-):  c+ e^{i m phi} + c- e^{-i m phi} = A cos m phi + B sin m phi
            out = make_double4(cp.x + cm.x, cp.y + cm.y, -(cp.y - cm.y), cp.x - cm.x);
        }
        reinterpret_cast<double4*>(rb + 4)[e] = out;
        // planar variant: e^{+-i m phi} are equal (phi = 0 or pi), so only c_{n,+m} + c_{n,-m} is needed
        if (rb2) reinterpret_cast<double2*>(rb2 + 4)[e] = make_double2(out.x, out.y);
    }
}

// planar[0] stays 1 only if every centre and every field point has the same x2: then x - c_b has no x2 component for any
// (point, ball), the azimuth of every pair is 0 or pi, and the +-m sums collapse (heat maps through coplanar centres,
// the reference's plot_biem use case: plot.py:63-82).
__global__ void uscat_planar_check_kernel(int64_t P, const double* __restrict__ x2, int B,
                                          const double* __restrict__ centers, int* __restrict__ planar) {
    const double c2 = centers[2];
    bool ok = true;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x)
        ok = ok && (x2[p] == c2);
    if (blockIdx.x == 0)
        for (int b = threadIdx.x; b < B; b += blockDim.x) ok = ok && (centers[b * 3 + 2] == c2);
    if (!ok) planar[0] = 0;
}

// ---- 3-D fast kernel --------------------------------------------------------------------------------
// u_b = sum_n h_n(k r) S_n,   S_n = sum_{m >= 0} P~_n^m(cos theta) ( A_{nm} cos m phi + B_{nm} sin m phi ),
// A = c_{n,+m} + c_{n,-m},  B = i (c_{n,+m} - c_{n,-m})  (normalisation and SD_n folded into A, B).
// One step n = N of the Legendre recurrence at fixed m (entered Duff-style at case m): 8 FP64 instructions -- the azimuthal
// combination w = A cm + B sm (4), S_N += q w (2), the recurrence (2).  The S_n are distinct registers, so the only dependent
// chain is the recurrence itself; h_n is generated once per (point, ball) AFTER the angular sums and never stored.
// The two recurrence registers swap roles with the parity of N so that no register moves are needed; every index is a
// compile-time immediate on top of the per-m base pointers.
#define US_ACC(Q)                                                               \
    {                                                                           \
        const double4 ab = recs_m[N_];                                          \
        const double wr = fma(ab.z, sm, ab.x * cm);                             \
        const double wi = fma(ab.w, sm, ab.y * cm);                             \
        Sr[N_] = fma((Q), wr, Sr[N_]);                                          \
        Si[N_] = fma((Q), wi, Si[N_]);                                          \
    }
#define US_STEP(N)                                                              \
    us_l##N:                                                                    \
        if ((N) < LMAX) {                                                       \
            constexpr int N_ = (N);                                             \
            const double bt = beta_m[N_];                                       \
            if ((N_ & 1) == 0) {                                                \
                US_ACC(qa)                                                      \
                qb = fma(ct, qa, -(bt * qb));                                   \
            } else {                                                            \
                US_ACC(qb)                                                      \
                qa = fma(ct, qb, -(bt * qa));                                   \
            }                                                                   \
        }

template <int LMAX, bool ZK>
__global__ void __launch_bounds__(US3D_THREADS) uscat3d_kernel(UscatArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // this kernel and the planar one are both launched; the device flag decides which one does the work
    if (a.planar && a.planar[0] != 0) return;
    const int L = a.L, B = a.B;
    constexpr int npair = LMAX * (LMAX + 1) / 2;
    constexpr int rec_doubles = 4 + 4 * npair;
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
            // angular sums S_n (registers, static indices)
            double Sr[LMAX], Si[LMAX];
#pragma unroll
            for (int n = 0; n < LMAX; ++n) { Sr[n] = 0.0; Si[n] = 0.0; }
            double qmm = 1.0, cm = 1.0, sm = 0.0;
            for (int m = 0; m < L; ++m) {
                // q_{m-1} = 0, q_m = sin^m; the register holding the current value depends on the parity of m
                double qa = (m & 1) ? 0.0 : qmm, qb = (m & 1) ? qmm : 0.0;
                const int off = m * LMAX - (m * (m - 1)) / 2 - m;
                const double4* recs_m = recs + off;  // (A, B) of (m, n >= m)
                const double* beta_m = sbeta + off;
                // binary dispatch to the Duff entry point n = m (a switch compiles to a linear compare chain here)
                if (m < 16) {
                    if (m < 8) {
                        if (m < 4) {
                            if (m < 2) {
                                if (m < 1) {
                                    goto us_l0;
                                } else {
                                    goto us_l1;
                                }
                            } else {
                                if (m < 3) {
                                    goto us_l2;
                                } else {
                                    goto us_l3;
                                }
                            }
                        } else {
                            if (m < 6) {
                                if (m < 5) {
                                    goto us_l4;
                                } else {
                                    goto us_l5;
                                }
                            } else {
                                if (m < 7) {
                                    goto us_l6;
                                } else {
                                    goto us_l7;
                                }
                            }
                        }
                    } else {
                        if (m < 12) {
                            if (m < 10) {
                                if (m < 9) {
                                    goto us_l8;
                                } else {
                                    goto us_l9;
                                }
                            } else {
                                if (m < 11) {
                                    goto us_l10;
                                } else {
                                    goto us_l11;
                                }
                            }
                        } else {
                            if (m < 14) {
                                if (m < 13) {
                                    goto us_l12;
                                } else {
                                    goto us_l13;
                                }
                            } else {
                                if (m < 15) {
                                    goto us_l14;
                                } else {
                                    goto us_l15;
                                }
                            }
                        }
                    }
                } else {
                    if (m < 24) {
                        if (m < 20) {
                            if (m < 18) {
                                if (m < 17) {
                                    goto us_l16;
                                } else {
                                    goto us_l17;
                                }
                            } else {
                                if (m < 19) {
                                    goto us_l18;
                                } else {
                                    goto us_l19;
                                }
                            }
                        } else {
                            if (m < 22) {
                                if (m < 21) {
                                    goto us_l20;
                                } else {
                                    goto us_l21;
                                }
                            } else {
                                if (m < 23) {
                                    goto us_l22;
                                } else {
                                    goto us_l23;
                                }
                            }
                        }
                    } else {
                        if (m < 28) {
                            if (m < 26) {
                                if (m < 25) {
                                    goto us_l24;
                                } else {
                                    goto us_l25;
                                }
                            } else {
                                if (m < 27) {
                                    goto us_l26;
                                } else {
                                    goto us_l27;
                                }
                            }
                        } else {
                            if (m < 30) {
                                if (m < 29) {
                                    goto us_l28;
                                } else {
                                    goto us_l29;
                                }
                            } else {
                                if (m < 31) {
                                    goto us_l30;
                                } else {
                                    goto us_l31;
                                }
                            }
                        }
                    }
                }
                US_STEP(0)
                US_STEP(1)
                US_STEP(2)
                US_STEP(3)
                US_STEP(4)
                US_STEP(5)
                US_STEP(6)
                US_STEP(7)
                US_STEP(8)
                US_STEP(9)
                US_STEP(10)
                US_STEP(11)
                US_STEP(12)
                US_STEP(13)
                US_STEP(14)
                US_STEP(15)
                US_STEP(16)
                US_STEP(17)
                US_STEP(18)
                US_STEP(19)
                US_STEP(20)
                US_STEP(21)
                US_STEP(22)
                US_STEP(23)
                US_STEP(24)
                US_STEP(25)
                US_STEP(26)
                US_STEP(27)
                US_STEP(28)
                US_STEP(29)
                US_STEP(30)
                US_STEP(31)
                {
                    const double cn = cm * cp - sm * sp;
                    sm = fma(sm, cp, cm * sp);
                    cm = cn;
                }
                qmm *= sn;
            }
            // radial part: u_b = sum_{n < L} h_n(k r) S_n, h_n upward (far field: h_n -> 1, the (-i)^n is in the coefficients)
            double br, bi;
            if (far) {
                br = 0.0; bi = 0.0;
#pragma unroll
                for (int n = 0; n < LMAX; ++n) {
                    if (n < L) { br += Sr[n]; bi += Si[n]; }
                }
            } else if (ZK) {
                const cplx z = cmake(k * r, a.k_im * r), iz = crecip(z), ex = cexp_i(z);
                cplx hm = cmul(cmake(ex.y, -ex.x), iz);                                       // h_0 = -i e^{iz} / z
                cplx hc = cmul(cmul(cmake(-z.x, -z.y - 1.0), ex), cmul(iz, iz));              // h_1 = -(z + i) e^{iz} / z^2
                br = hm.x * Sr[0] - hm.y * Si[0];
                bi = hm.x * Si[0] + hm.y * Sr[0];
#pragma unroll
                for (int n = 1; n < LMAX; ++n) {
                    if (n < L) {
                        br = fma(hc.x, Sr[n], fma(-hc.y, Si[n], br));
                        bi = fma(hc.x, Si[n], fma(hc.y, Sr[n], bi));
                        const cplx cf = cscale(iz, 2.0 * n + 1.0);
                        const cplx hn = cmake(fma(cf.x, hc.x, fma(-cf.y, hc.y, -hm.x)), fma(cf.x, hc.y, fma(cf.y, hc.x, -hm.y)));
                        hm = hc;
                        hc = hn;
                    }
                }
            } else {
                const double z = k * r, iz = 1.0 / z;
                double s, co;
                sincos(z, &s, &co);
                double hmr = s * iz, hmi = -co * iz;                          // h_0
                double hcr = (s * iz - co) * iz, hci = (-co * iz - s) * iz;   // h_1
                br = hmr * Sr[0] - hmi * Si[0];
                bi = hmr * Si[0] + hmi * Sr[0];
#pragma unroll
                for (int n = 1; n < LMAX; ++n) {
                    if (n < L) {
                        br = fma(hcr, Sr[n], fma(-hci, Si[n], br));
                        bi = fma(hcr, Si[n], fma(hci, Sr[n], bi));
                        const double cf = (2 * n + 1) * iz;
                        const double hnr = fma(cf, hcr, -hmr), hni = fma(cf, hci, -hmi);
                        hmr = hcr; hmi = hci;
                        hcr = hnr; hci = hni;
                    }
                }
            }
            if (far) {
                // (ik)^{-1} exp(-i k x.c_b), x as given (_biem.py:931-944)
                const double dotc = x0 * rb[0] + x1 * rb[1] + x2 * rb[2];
                double ph = -k * dotc;
                double s, co;
                sincos(ph, &s, &co);
                if (ZK) {
                    // exp(-i (k + i k_im) x.c) / (i (k + i k_im))
                    const double g = exp(a.k_im * dotc);
                    const cplx t = cmul(cmake(br, bi), cmake(g * co, g * s));
                    const cplx q = cdiv(t, cmake(-a.k_im, k));
                    br = q.x;
                    bi = q.y;
                } else {
                    // (br + i bi) * (co + i s) * (-i / k)
                    double tr = br * co - bi * s, ti = br * s + bi * co;
                    br = ti / k;
                    bi = -tr / k;
                }
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

// ---- 3-D planar kernel ------------------------------------------------------------------------------------------------
// Every field point and every centre share the same x2 (a heat map through coplanar spheres: the reference's plot_biem use
// case and config C5).  In the frame whose POLAR axis is the plane's normal every direction x - c_b lies on the equator:
//     Y'_{n,m}(pi/2, phi') = K_{n,m} e^{i m phi'},   K_{n,m} = 0 for n + |m| odd,   phi' = atan2(x1 - c1, x0 - c0),
// so with the coefficients rotated into that frame once per ball (plan tables d_us_rot / d_us_K) there is NO Legendre
// recurrence left and half of the (n, m) pairs vanish:
//     u_b = sum_n h_n(k r) S_n,   S_n = sum_{m >= 0, m = n mod 2} ( A_{nm} cos m phi' + B_{nm} sin m phi' ),
//     A = g_{n,+m} + g_{n,-m},  B = i (g_{n,+m} - g_{n,-m}),  g = K c'.
// Per (point, ball): 4 FMAs per surviving pair (LMAX (LMAX + 2) / 4 of them: 156 at n_end = 24) + 10 per degree, against
// 8 per (n, m) step (300 steps) of the polar-axis-in-plane form: ~2.8x fewer FP64 instructions, and no dependent chains --
// the S_n live in registers and consecutive FMAs touch different ones.
__host__ __device__ constexpr int planar_pairs(int LMAX) { return LMAX * (LMAX + 2) / 4; }  // LMAX even

// rec3[b] = [c0, c1, c2, rho | (A.re, A.im, B.re, B.im) per pair, m-major: m = 0 .. LMAX-1, n = m, m + 2, ... < LMAX]
__global__ void __launch_bounds__(128) uscat_coef_planar_kernel(int L, int LMAX, int B, double k, double k_im, double eta,
                                                                const double* __restrict__ centers,
                                                                const double* __restrict__ radii,
                                                                const double4* __restrict__ rad, const cplx* __restrict__ radz,
                                                                const cplx* __restrict__ rot, const double* __restrict__ K,
                                                                const cplx* __restrict__ density, double* __restrict__ rec3) {
    extern __shared__ __align__(16) cplx s_coef[];  // [L * L]: density * SD_n
    const int b = blockIdx.x, H = L * L;
    const int np = planar_pairs(LMAX);
    double* rb = rec3 + (int64_t)(4 + 4 * np) * b;
    if (threadIdx.x < 4) rb[threadIdx.x] = threadIdx.x < 3 ? centers[b * 3 + threadIdx.x] : radii[b];
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        int n = 0;
        while ((n + 1) * (n + 1) <= h) ++n;
        cplx sd;
        if (radz) {
            const cplx* rz = radz + ((int64_t)b * L + n) * 4;
            sd = sd_coef_z(3, cmake(k, k_im), eta, radii[b], rz[0], rz[1]);
        } else {
            const double4 r = rad[(int64_t)b * L + n];
            sd = sd_coef(3, k, eta, radii[b], r.x, r.y);
        }
        s_coef[h] = cmul(density[(int64_t)b * H + h], sd);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < np; e += blockDim.x) {
        int m = 0, off = 0;
        while (off + (LMAX - m + 1) / 2 <= e) { off += (LMAX - m + 1) / 2; ++m; }
        const int n = m + 2 * (e - off);
        double4 out = make_double4(0.0, 0.0, 0.0, 0.0);
        if (n < L) {
            const int w = 2 * n + 1;
            const cplx* blk = rot + ((int64_t)4 * n * n * n - n) / 3;
            const cplx* cf = s_coef + n * n;
            cplx gp = cmake(0.0, 0.0), gm = cmake(0.0, 0.0);
            const cplx* rp = blk + (int64_t)m * w;                    // row of +m
            const cplx* rm = blk + (int64_t)(m ? w - m : 0) * w;      // row of -m
            for (int c = 0; c < w; ++c) {
                gp = cfma(rp[c], cf[c], gp);
                if (m) gm = cfma(rm[c], cf[c], gm);
            }
            gp = cscale(gp, K[n * n + m]);
            gm = cscale(gm, m ? K[n * n + w - m] : 0.0);
            // A = g+ + g-,  B = i (g+ - g-)
            out = make_double4(gp.x + gm.x, gp.y + gm.y, -(gp.y - gm.y), gp.x - gm.x);
        }
        reinterpret_cast<double4*>(rb + 4)[e] = out;
    }
}

template <int LMAX, bool ZK>
__global__ void __launch_bounds__(US3D_THREADS, LMAX <= 24 ? 3 : 2) uscat3d_planar_kernel(UscatArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (a.planar[0] == 0) return;  // not coplanar: the general kernel (launched beside this one) does the work
    const int L = a.L, B = a.B;
    constexpr int np = planar_pairs(LMAX);
    constexpr int rec_doubles = 4 + 4 * np;
    double* stage0 = reinterpret_cast<double*>(smem_raw);
    double* stage1 = stage0 + US3D_CB * rec_doubles;
    __shared__ __align__(8) uint64_t full[2];
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int nchunks = (B + US3D_CB - 1) / US3D_CB;
    if (tid == 0) {
        const int nb = min(US3D_CB, B);
        const uint32_t bytes = (uint32_t)(nb * rec_doubles * sizeof(double));
        mbar_expect_tx(&full[0], bytes);
        tma_load_1d(stage0, a.rec3, bytes, &full[0]);
    }
    const int64_t p = (int64_t)blockIdx.x * US3D_THREADS + tid;
    const bool active = p < a.P;
    const double x0 = active ? a.x[p] : 1e3, x1 = active ? a.x[a.P + p] : 1e3;
    const bool inner = a.flags & BHS_FLAG_INNER, per_ball = a.flags & BHS_FLAG_PER_BALL;
    const double k = a.k;
    double accr = 0.0, acci = 0.0;
    bool bad = false;
    for (int c = 0; c < nchunks; ++c) {
        if (tid == 0 && c + 1 < nchunks) {
            const int nb = min(US3D_CB, B - (c + 1) * US3D_CB);
            const uint32_t bytes = (uint32_t)(nb * rec_doubles * sizeof(double));
            mbar_expect_tx(&full[(c + 1) & 1], bytes);
            tma_load_1d(((c + 1) & 1) ? stage1 : stage0, a.rec3 + (int64_t)(c + 1) * US3D_CB * rec_doubles, bytes,
                        &full[(c + 1) & 1]);
        }
        mbar_wait(&full[c & 1], (c >> 1) & 1);
        const double* st = (c & 1) ? stage1 : stage0;
        const int nb = min(US3D_CB, B - c * US3D_CB);
        for (int bb = 0; bb < nb; ++bb) {
            const double* rb = st + bb * rec_doubles;
            const double4* recs = reinterpret_cast<const double4*>(rb + 4);
            const double dx0 = x0 - rb[0], dx1 = x1 - rb[1], rho = rb[3];
            const double r = sqrt(dx0 * dx0 + dx1 * dx1);
            bad = bad || (inner ? (r > rho) : (r < rho));
            const double ir = 1.0 / r;
            const double cp = r > 0.0 ? dx0 * ir : 1.0, sp = r > 0.0 ? dx1 * ir : 0.0;
            // angular sums S_n: every register index below is a compile-time constant (full unroll)
            double Sr[LMAX], Si[LMAX];
#pragma unroll
            for (int n = 0; n < LMAX; ++n) { Sr[n] = 0.0; Si[n] = 0.0; }
            double cm = 1.0, sm = 0.0;
            int e = 0;
#pragma unroll
            for (int m = 0; m < LMAX; ++m) {
#pragma unroll
                for (int n = m; n < LMAX; n += 2) {
                    const double4 ab = recs[e];
                    ++e;
                    Sr[n] = fma(ab.x, cm, Sr[n]);
                    Si[n] = fma(ab.y, cm, Si[n]);
                    if (m > 0) {
                        Sr[n] = fma(ab.z, sm, Sr[n]);
                        Si[n] = fma(ab.w, sm, Si[n]);
                    }
                }
                const double cn = cm * cp - sm * sp;
                sm = fma(sm, cp, cm * sp);
                cm = cn;
            }
            // radial part: u_b = sum_{n < L} h_n(k r) S_n, h_n upward (orders >= L only ever meet zero sums: skipped)
            double br, bi;
            if (ZK) {
                const cplx z = cmake(k * r, a.k_im * r), iz = crecip(z), ex = cexp_i(z);
                cplx hm = cmul(cmake(ex.y, -ex.x), iz);                                       // h_0
                cplx hc = cmul(cmul(cmake(-z.x, -z.y - 1.0), ex), cmul(iz, iz));              // h_1
                br = hm.x * Sr[0] - hm.y * Si[0];
                bi = hm.x * Si[0] + hm.y * Sr[0];
#pragma unroll
                for (int n = 1; n < LMAX; ++n) {
                    if (n < L) {
                        br = fma(hc.x, Sr[n], fma(-hc.y, Si[n], br));
                        bi = fma(hc.x, Si[n], fma(hc.y, Sr[n], bi));
                        const cplx cf = cscale(iz, 2.0 * n + 1.0);
                        const cplx hn = cmake(fma(cf.x, hc.x, fma(-cf.y, hc.y, -hm.x)), fma(cf.x, hc.y, fma(cf.y, hc.x, -hm.y)));
                        hm = hc;
                        hc = hn;
                    }
                }
            } else {
                const double z = k * r, iz = 1.0 / z;
                double s, co;
                sincos(z, &s, &co);
                double hmr = s * iz, hmi = -co * iz;                          // h_0
                double hcr = (s * iz - co) * iz, hci = (-co * iz - s) * iz;   // h_1
                br = hmr * Sr[0] - hmi * Si[0];
                bi = hmr * Si[0] + hmi * Sr[0];
#pragma unroll
                for (int n = 1; n < LMAX; ++n) {
                    if (n < L) {
                        br = fma(hcr, Sr[n], fma(-hci, Si[n], br));
                        bi = fma(hcr, Si[n], fma(hci, Sr[n], bi));
                        const double cf = (2 * n + 1) * iz;
                        const double hnr = fma(cf, hcr, -hmr), hni = fma(cf, hci, -hmi);
                        hmr = hcr; hmi = hci;
                        hcr = hnr; hci = hni;
                    }
                }
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
    if (bad) {
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
struct SmArrC {
    cplx* p;
    __device__ __forceinline__ cplx& operator[](int n) const { return p[n]; }
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
    cplx* Hz = reinterpret_cast<cplx*>(Hr);  // complex-wavenumber view of the same scratch
    const bool zk = a.k_im != 0.0;
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
                    for (int n = 0; n < L; ++n) {
                        if (zk) Hz[n] = cmake(1.0, 0.0);
                        else { Hr[n] = 1.0; Hi[n] = 0.0; }
                    }
                } else if (zk) {
                    // complex wavenumber: the (Hr, Hi) scratch (2 (L + 2 + shift) doubles) holds the sequence as
                    // L + shift interleaved complex values Hz[n]
                    radial_sequence_z(d, cmake(a.k * r, a.k_im * r), L - 1, SmArrC{Hz}, SmArrC{Hz}, false, true);
                } else {
                    hankel_upward(d, a.k * r, L - 1, SmArr{Hr}, SmArr{Hi});
                }
            }
            __syncwarp();
            double pr = 0.0, pi = 0.0;
            for (int h = lane; h < a.H; h += 32) {
                cplx y = harmonic_from_tables(tb, L, a.idx + (int64_t)h * s, F, E);
                int n = a.deg[h];
                cplx t = cmul(zk ? Hz[n] : cmake(Hr[n], Hi[n]), y);
                cplx c = a.coefg[(int64_t)b * a.H + h];
                pr += t.x * c.x - t.y * c.y;
                pi += t.x * c.y + t.y * c.x;
            }
            if (far) {
                // (ik)^{-(d-1)/2} exp(-i k x.c_b): k^{-p} e^{-i pi p / 2}, p = (d-1)/2
                if (a.k_im != 0.0) {
                    // exp(-i (k + i k_im) x.c) * (i (k + i k_im))^{-(d-1)/2}   (principal branch, as numpy's power)
                    double sn, co;
                    sincos(-a.k * dot, &sn, &co);
                    const double g = exp(a.k_im * dot);
                    const cplx lg = clog_(cmake(-a.k_im, a.k));
                    const cplx pwz = cexp_(cscale(lg, -0.5 * (d - 1)));
                    const cplx q = cmul(cmul(cmake(pr, pi), cmake(g * co, g * sn)), pwz);
                    pr = q.x; pi = q.y;
                } else {
                    double pw = 0.5 * (d - 1);
                    double ang = -a.k * dot - 1.57079632679489661923 * pw;
                    double sn, co;
                    sincos(ang, &sn, &co);
                    double mag = pow(a.k, -pw);
                    double tr = (pr * co - pi * sn) * mag, ti = (pr * sn + pi * co) * mag;
                    pr = tr; pi = ti;
                }
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
    int64_t L = plan->n_end, LM = (L + 7) / 8 * 8, npair = LM * (LM + 1) / 2;
    int64_t rad = align256((int64_t)B * L * 4 * sizeof(cplx));  // real (j, j', y, y') or complex (j, j', h, h') table
    const int64_t LP = (L + 3) / 4 * 4;  // band of the planar kernel (multiple of 4)
    int64_t c3 = align256((int64_t)B * (4 + 4 * npair) * sizeof(double)) + align256(npair * sizeof(double)) +
                 align256((int64_t)B * (4 + 4 * planar_pairs((int)LP)) * sizeof(double)) + 256;  // records, beta, planar records, flag
    int64_t cg = align256((int64_t)B * plan->H * sizeof(cplx));
    int64_t kbuf = 256;
    // + the radial kernel's global order scratch (only for orders too high for shared memory; see radial.cuh)
    return rad + (c3 > cg ? c3 : cg) + kbuf + align256((int64_t)ball_radial_scratch_bytes(plan->d, (int)L));
}

template <int LMAX, bool ZK>
static void launch_uscat3d_one(const UscatArgs& a, cudaStream_t st) {
    const int npair = LMAX * (LMAX + 1) / 2;
    size_t smem = (size_t)(2 * US3D_CB * (4 + 4 * npair) + npair) * sizeof(double);
    cudaFuncSetAttribute(uscat3d_kernel<LMAX, ZK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = (a.P + US3D_THREADS - 1) / US3D_THREADS;
    uscat3d_kernel<LMAX, ZK><<<(unsigned)blocks, US3D_THREADS, smem, st>>>(a);
}

template <int LMAX, bool ZK>
static void launch_planar_one(const UscatArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)(2 * US3D_CB * (4 + 4 * planar_pairs(LMAX))) * sizeof(double);
    cudaFuncSetAttribute(uscat3d_planar_kernel<LMAX, ZK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t blocks = (a.P + US3D_THREADS - 1) / US3D_THREADS;
    uscat3d_planar_kernel<LMAX, ZK><<<(unsigned)blocks, US3D_THREADS, smem, st>>>(a);
}
template <int LMAX>
static void launch_planar(const UscatArgs& a, cudaStream_t st) {
    if (a.k_im != 0.0) launch_planar_one<LMAX, true>(a, st);
    else launch_planar_one<LMAX, false>(a, st);
}

// a.rec3 != nullptr: the planar kernel is launched as well; the device flag a.planar selects which of the two does the work
template <int LMAX>
static int launch_uscat3d(const UscatArgs& a, cudaStream_t st) {
    bhs_prof_begin(BHS_PROF_USCAT, st);
    if (a.k_im != 0.0) launch_uscat3d_one<LMAX, true>(a, st);
    else launch_uscat3d_one<LMAX, false>(a, st);
    BHS_COUNT_LAUNCH();
    if (a.rec3) {
        switch ((a.L + 3) / 4) {
            case 1: launch_planar<4>(a, st); break;
            case 2: launch_planar<8>(a, st); break;
            case 3: launch_planar<12>(a, st); break;
            case 4: launch_planar<16>(a, st); break;
            case 5: launch_planar<20>(a, st); break;
            case 6: launch_planar<24>(a, st); break;
            case 7: launch_planar<28>(a, st); break;
            default: launch_planar<32>(a, st); break;
        }
    }
    bhs_prof_end(BHS_PROF_USCAT, 8.0 * (double)a.P * a.B * a.H, st);
    BHS_CHECK_LAUNCH();
    return BHS_OK;
}

extern "C" int bhs_uscat(const bhs_plan_t* plan, int B, const double* d_centers, const double* d_radii, double k,
                         double k_im, double eta, const double* d_density, const double* d_x, int64_t P, int flags,
                         double* d_out, void* d_work, void* stream) {
    if (!plan || B <= 0 || P < 0) return BHS_ERR_INVALID;
    if (P == 0) return BHS_OK;  // empty point set: nothing to do (the buffers of empty arrays may be null)
    if (!d_centers || !d_radii || !d_density || !d_x || !d_out || !d_work) return BHS_ERR_INVALID;
    // real wavenumbers must be positive; with Im k > 0 (absorbing medium) any real part is a valid argument of h^{(1)}_n
    if (k_im == 0.0 ? !(k > 0.0) : !(k_im > 0.0 || k > 0.0)) return BHS_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int d = plan->d, L = plan->n_end, H = plan->H;
    const int64_t npair = (int64_t)L * (L + 1) / 2;
    unsigned char* w = (unsigned char*)d_work;
    double4* d_rad = (double4*)(w + 256);
    unsigned char* d_coef = w + 256 + align256((int64_t)B * L * 4 * sizeof(cplx));
    const cplx* d_radz = (k_im != 0.0) ? (const cplx*)d_rad : nullptr;
    const size_t scr = ball_radial_scratch_bytes(d, L);
    double* d_scratch = scr ? (double*)(w + bhs_uscat_workspace(plan, B) - align256((int64_t)scr)) : nullptr;
    int rc = d_radz ? launch_ball_radial_z(d, L, B, 1, d_radii, nullptr, nullptr, k, k_im, (cplx*)d_rad, st)
                    : launch_ball_radial(d, L, B, 1, d_radii, nullptr, k, d_rad, d_scratch, st);
    if (rc) return rc;
    const int far = (flags & BHS_FLAG_FAR_FIELD) ? 1 : 0;
    UscatArgs a;
    a.d = d; a.L = L; a.H = H; a.B = B; a.flags = flags; a.k = k; a.k_im = k_im; a.eta = eta;
    a.x = d_x; a.P = P; a.centers = d_centers; a.radii = d_radii;
    a.coefg = nullptr; a.rec = nullptr; a.rec3 = nullptr; a.planar = nullptr; a.beta = plan->d_us_beta; a.idx = plan->d_idx; a.deg = plan->d_deg;
    a.out = (cplx*)d_out;
    if (d == 3 && L <= 32) {
        const int LM = (L + 7) / 8 * 8;
        const int64_t npm = (int64_t)LM * (LM + 1) / 2;
        unsigned char* q = d_coef + align256((int64_t)B * (4 + 4 * npm) * sizeof(double));
        double* d_beta = (double*)q;
        q += align256(npm * sizeof(double));
        const int LP = (L + 3) / 4 * 4;
        double* d_rec3 = (double*)q;
        q += align256((int64_t)B * (4 + 4 * planar_pairs(LP)) * sizeof(double));
        int* d_planar = (int*)q;
        uscat_coef3d_kernel<<<B, 128, 0, st>>>(L, LM, B, k, k_im, eta, far, d_centers, d_radii, d_rad, d_radz,
                                               plan->d_us_norm, (const cplx*)d_density, (double*)d_coef, d_beta, nullptr,
                                               d_planar);
        BHS_CHECK_LAUNCH();
        a.rec = (const double*)d_coef;
        a.beta = d_beta;
        a.planar = nullptr;
        if (!far && plan->d_us_rot) {
            // coplanar points and centres (decided on the device, no host round trip): coefficients rotated into the frame
            // whose polar axis is the plane's normal -> no Legendre recurrence, half of the (n, m) pairs vanish
            int64_t cb = (P + 255) / 256;
            if (cb > bhs_sm_count() * 4) cb = bhs_sm_count() * 4;
            uscat_planar_check_kernel<<<(unsigned)cb, 256, 0, st>>>(P, d_x + 2 * P, B, d_centers, d_planar);
            BHS_CHECK_LAUNCH();
            uscat_coef_planar_kernel<<<B, 128, (size_t)H * sizeof(cplx), st>>>(L, LP, B, k, k_im, eta, d_centers, d_radii, d_rad,
                                                                           d_radz, plan->d_us_rot, plan->d_us_K,
                                                                           (const cplx*)d_density, d_rec3);
            BHS_CHECK_LAUNCH();
            a.planar = d_planar;
            a.rec3 = d_rec3;
        }
        (void)npair;
        if (L <= 8) return launch_uscat3d<8>(a, st);
        if (L <= 16) return launch_uscat3d<16>(a, st);
        if (L <= 24) return launch_uscat3d<24>(a, st);
        return launch_uscat3d<32>(a, st);
    }
    int64_t tot = (int64_t)B * H;
    uscat_coef_generic_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d, L, H, B, k, k_im, eta, far, d_radii,
                                                                            d_rad, d_radz, plan->d_deg,
                                                                            (const cplx*)d_density, (cplx*)d_coef);
    BHS_CHECK_LAUNCH();
    a.coefg = (const cplx*)d_coef;
    int warps = 4;
    const int shift = (d & 1) ? (d - 3) / 2 : d / 2 - 1;
    const size_t per_warp = harm_smem_bytes_per_warp(d, L) + (size_t)2 * (L + 2 + shift) * sizeof(double);
    while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
    size_t smem = per_warp * warps;
    if (smem > 200 * 1024) return BHS_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(uscat_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = (P + warps - 1) / warps;
    if (blocks > bhs_sm_count() * 16) blocks = bhs_sm_count() * 16;
    bhs_prof_begin(BHS_PROF_USCAT, st);
    uscat_generic_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(a, harm_tables_of(plan));
    bhs_prof_end(BHS_PROF_USCAT, 8.0 * (double)a.P * a.B * a.H, st);
    BHS_CHECK_LAUNCH();
    return BHS_OK;
}
