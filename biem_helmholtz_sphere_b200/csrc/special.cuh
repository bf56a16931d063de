// Hyperspherical Bessel / Hankel functions for real positive argument, FP64.
//
//   z_n^{(d)}(x) = sqrt(pi/2) Z_{n+d/2-1}(x) / x^{d/2-1}          (ultrasphere.shn1)
//
// Everything is derived from two base sequences:
//   even d : cylindrical  J_m, Y_m  (integer order),  z_n^{(d)} = sqrt(pi/2) Z_{n+s}(x)/x^s,   s = d/2-1
//   odd  d : spherical    j_n, y_n,                   z_n^{(d)} = z^{(3)}_{n+s}(x)/x^s,        s = (d-3)/2
// y (hence h = j + i y): upward three-term recurrence (dominant solution, stable).
// j: downward Miller recurrence normalised against the closed-form / Neumann-series start values
//    whenever n_max exceeds the argument; upward otherwise (oscillatory region, stable).
#pragma once
#include "common.cuh"

#define BHS_SQRT_PI_2 1.2533141373155002512
#define BHS_2_PI 0.63661977236758134308
#define BHS_EULER 0.57721566490153286061

// ---- cylindrical J0, J1, Y0, Y1 -------------------------------------------------------------------
// x >= 25: Hankel asymptotic series (error ~ e^{-2x}); x < 25: Miller + Neumann series.
static __device__ __noinline__ void cyl_jy01(double x, double& j0, double& j1, double& y0, double& y1) {
    if (x >= 25.0) {
        double s, c;
        sincos(x, &s, &c);
        const double r2 = 0.70710678118654752440;
        double pref = sqrt(BHS_2_PI / x);
        double inv8x = 0.125 / x;
#pragma unroll
        for (int nu = 0; nu < 2; ++nu) {
            double mu = 4.0 * nu * nu;
            double P = 1.0, Q = 0.0, term = 1.0;
            // term_k = term_{k-1} * (mu - (2k-1)^2) / (k * 8x); P takes even k (sign (-1)^{k/2}), Q odd k
            for (int kk = 1; kk < 60; ++kk) {
                double f = (mu - (2.0 * kk - 1.0) * (2.0 * kk - 1.0)) * inv8x / kk;
                double nt = term * f;
                if (fabs(nt) >= fabs(term) && kk > 2) break;
                term = nt;
                int h = kk >> 1;
                double sgn = (h & 1) ? -1.0 : 1.0;
                if (kk & 1) Q += sgn * term; else P += sgn * term;
                if (fabs(term) < 1e-18) break;
            }
            // chi = x - pi/4 - nu*pi/2
            double cc, ss;
            if (nu == 0) { cc = (c + s) * r2; ss = (s - c) * r2; }
            else         { cc = (s - c) * r2; ss = -(c + s) * r2; }
            double J = pref * (P * cc - Q * ss);
            double Y = pref * (P * ss + Q * cc);
            if (nu == 0) { j0 = J; y0 = Y; } else { j1 = J; y1 = Y; }
        }
        return;
    }
    // Miller backward recurrence from even order M
    int M = 2 * (int)(0.75 * x + 16.0);
    double jp1 = 0.0, jc = 1e-280;   // J_{m+1}, J_m (unnormalised)
    double sum = 0.0;                 // J0 + 2 sum J_{2k}
    double ysum = 0.0;                // sum_{k>=1} (-1)^k J_{2k}/k
    double y1sum = 0.0;               // sum_{k>=1} (-1)^k (J_{2k-1} - J_{2k+1})/k
    double twox = 2.0 / x;
    // at loop top: jc = J_m, jp1 = J_{m+1}
    for (int m = M; m >= 1; --m) {
        double jm1 = m * twox * jc - jp1;  // J_{m-1}
        if ((m & 1) == 0) {
            int kk = m >> 1;
            double sg = (kk & 1) ? -1.0 : 1.0;
            sum += 2.0 * jc;
            ysum += sg * jc / kk;
            y1sum += sg * (jm1 - jp1) / kk;
        }
        jp1 = jc;
        jc = jm1;
        if (fabs(jc) > 1e200) {
            const double sc = 1e-200;
            jc *= sc; jp1 *= sc; sum *= sc; ysum *= sc; y1sum *= sc;
        }
    }
    sum += jc;  // J0
    double inv = 1.0 / sum;
    j0 = jc * inv;
    j1 = jp1 * inv;
    ysum *= inv;
    y1sum *= inv;
    double lg = log(0.5 * x) + BHS_EULER;
    y0 = BHS_2_PI * (lg * j0 - 2.0 * ysum);
    y1 = BHS_2_PI * (-j0 / x + lg * j1 + y1sum);
}

// Miller start order so that the minimal solution is resolved to ~1e-17 at order n_top.
__device__ __forceinline__ int miller_start(int n_top, double x) {
    double a = fmax((double)n_top, x);
    return (int)(a + 24.0 + 6.5 * sqrt(a + 1.0));
}

// Base sequences.  out_j / out_y: arrays of length n_top+1 (either may be nullptr).
//   even_dim == 1: cylindrical J_n, Y_n ; even_dim == 0: spherical j_n, y_n.
template <typename ArrJ, typename ArrY>
__device__ void base_sequence(int even_dim, double x, int n_top, ArrJ out_j, ArrY out_y, bool want_j, bool want_y) {
    double f0, f1, g0, g1;  // j-type and y-type start values (orders 0, 1)
    if (even_dim) {
        cyl_jy01(x, f0, f1, g0, g1);
    } else {
        double s, c;
        sincos(x, &s, &c);
        double ix = 1.0 / x;
        f0 = s * ix;
        f1 = (s * ix - c) * ix;
        g0 = -c * ix;
        g1 = (-c * ix - s) * ix;
        if (x < 0.5) {
            // series for j1 to avoid cancellation: j1 = x/3 (1 - x^2/10 + x^4/280 - x^6/15120 + ...)
            double x2 = x * x;
            f1 = x / 3.0 * (1.0 - x2 / 10.0 * (1.0 - x2 / 28.0 * (1.0 - x2 / 54.0 * (1.0 - x2 / 88.0 * (1.0 - x2 / 130.0)))));
        }
    }
    // order step: cylindrical  Z_{m+1} = (2m/x) Z_m - Z_{m-1};  spherical z_{n+1} = ((2n+1)/x) z_n - z_{n-1}
    double ix = 1.0 / x;
    double off = even_dim ? 0.0 : 1.0;
    if (want_y) {
        double ym = g0, yc = g1;
        out_y[0] = g0;
        if (n_top >= 1) out_y[1] = g1;
        for (int n = 1; n < n_top; ++n) {
            double yn = (2.0 * n + off) * ix * yc - ym;
            ym = yc; yc = yn;
            out_y[n + 1] = yn;
        }
    }
    if (want_j) {
        if ((double)n_top + 1.0 <= x) {
            double jm = f0, jc = f1;
            out_j[0] = f0;
            if (n_top >= 1) out_j[1] = f1;
            for (int n = 1; n < n_top; ++n) {
                double jn = (2.0 * n + off) * ix * jc - jm;
                jm = jc; jc = jn;
                out_j[n + 1] = jn;
            }
        } else {
            int M = miller_start(n_top, x);
            double jp1 = 0.0, jc = 1e-280;  // orders M+1, M
            double scale_acc = 1.0;         // product of rescalings applied after entries were stored
            for (int m = M; m >= 1; --m) {
                double jm1 = (2.0 * m + off) * ix * jc - jp1;
                jp1 = jc; jc = jm1;  // jc = order m-1
                if (m - 1 <= n_top) out_j[m - 1] = jc;
                if (fabs(jc) > 1e200) {
                    const double sc = 1e-200;
                    jc *= sc; jp1 *= sc;
                    for (int q = m - 1; q <= n_top; ++q) out_j[q] *= sc;
                }
            }
            (void)scale_acc;
            // normalise on the larger of the two start values
            double nrm = (fabs(f0) >= fabs(f1)) ? f0 / out_j[0] : f1 / out_j[(n_top >= 1) ? 1 : 0];
            if (n_top == 0 && fabs(f0) < fabs(f1)) nrm = f1 / jp1;
            for (int q = 0; q <= n_top; ++q) out_j[q] *= nrm;
        }
    }
}

// z_n^{(d)} for n = 0..n_max(+1 internally for derivatives).  Arrays must hold n_max + 2 + shift entries,
// shift = d/2-1 (even) or (d-3)/2 (odd).  After the call arr[n] (n = 0..n_max+1) holds z_n^{(d)}.
template <typename ArrJ, typename ArrY>
__device__ void radial_sequence(int d, double x, int n_max_plus, ArrJ aj, ArrY ay, bool want_j, bool want_y) {
    int even_dim = (d & 1) == 0;
    int shift = even_dim ? d / 2 - 1 : (d - 3) / 2;
    base_sequence(even_dim, x, n_max_plus + shift, aj, ay, want_j, want_y);
    double sc = even_dim ? BHS_SQRT_PI_2 : 1.0;
    for (int s = 0; s < shift; ++s) sc /= x;
    if (shift > 0 || even_dim) {
        for (int n = 0; n <= n_max_plus; ++n) {
            if (want_j) aj[n] = aj[n + shift] * sc;
            if (want_y) ay[n] = ay[n + shift] * sc;
        }
    }
}
// derivative: z_n' = (n/x) z_n - z_{n+1}
__device__ __forceinline__ double radial_deriv(int n, double x, double zn, double znp1) { return (n / x) * zn - znp1; }

// h_n^{(d)}(x) = j + i y for n = 0..n_max by UPWARD recurrence of both parts.  The regular part loses
// relative accuracy once n > x, but its error is O(eps * |y_n|), i.e. O(eps) relative to |h_n|: exactly
// what field evaluation needs.  Arrays must hold n_max + 1 + shift entries.
template <typename ArrR, typename ArrI>
__device__ void hankel_upward(int d, double x, int n_max, ArrR hr, ArrI hi) {
    int even_dim = (d & 1) == 0;
    int shift = even_dim ? d / 2 - 1 : (d - 3) / 2;
    double f0, f1, g0, g1;
    double ix = 1.0 / x;
    if (even_dim) {
        cyl_jy01(x, f0, f1, g0, g1);
    } else {
        double s, c;
        sincos(x, &s, &c);
        f0 = s * ix; f1 = (s * ix - c) * ix; g0 = -c * ix; g1 = (-c * ix - s) * ix;
    }
    double off = even_dim ? 0.0 : 1.0;
    int n_top = n_max + shift;
    hr[0] = f0; hi[0] = g0;
    if (n_top >= 1) { hr[1] = f1; hi[1] = g1; }
    for (int n = 1; n < n_top; ++n) {
        double c = (2.0 * n + off) * ix;
        hr[n + 1] = c * hr[n] - hr[n - 1];
        hi[n + 1] = c * hi[n] - hi[n - 1];
    }
    double sc = even_dim ? BHS_SQRT_PI_2 : 1.0;
    for (int s = 0; s < shift; ++s) sc *= ix;
    if (shift > 0 || even_dim)
        for (int n = 0; n <= n_max; ++n) { hr[n] = hr[n + shift] * sc; hi[n] = hi[n + shift] * sc; }
}

// =====================================================================================================
// Complex argument (complex wavenumber, Im k != 0), spherical family only (odd d):
//   j_n(z) : downward Miller recurrence normalised on the closed forms j_0 / j_1 when n_top >= |z|, upward otherwise;
//   h_n(z) = h_n^{(1)}(z) : UPWARD from h_0 = -i e^{iz}/z, h_1 = -(z + i) e^{iz}/z^2.  h is never formed as j + i y:
//            for Im z > 0 it is exponentially smaller than j and y, and the sum would cancel.
// Even d (cylindrical family) is not implemented for complex arguments (needs K_nu-type algorithms for H^{(1)}).
// =====================================================================================================
__device__ __forceinline__ cplx cexp_i(cplx z) {  // e^{i z}
    double s, c;
    sincos(z.x, &s, &c);
    const double e = exp(-z.y);
    return cmake(e * c, e * s);
}
__device__ __forceinline__ double cabs1(cplx a) { return fabs(a.x) + fabs(a.y); }
__device__ __forceinline__ double cabs2(cplx a) { return hypot(a.x, a.y); }

// Arrays of length n_top + 1 (complex); either may be skipped.  z != 0.
template <typename ArrJ, typename ArrH>
__device__ void sph_sequence_z(cplx z, int n_top, ArrJ aj, ArrH ah, bool want_j, bool want_h) {
    const cplx iz = crecip(z);
    const cplx eiz = cexp_i(z);                            // e^{iz}
    const cplx emiz = cexp_i(cmake(-z.x, -z.y));           // e^{-iz}
    if (want_h) {
        // h0 = -i e^{iz} / z ; h1 = -(z + i) e^{iz} / z^2
        cplx h0 = cmul(cmake(eiz.y, -eiz.x), iz);
        cplx h1 = cmul(cmul(cmake(-z.x, -z.y - 1.0), eiz), cmul(iz, iz));
        ah[0] = h0;
        if (n_top >= 1) ah[1] = h1;
        cplx hm = h0, hc = h1;
        for (int n = 1; n < n_top; ++n) {
            cplx hn = csub(cmul(cscale(iz, 2.0 * n + 1.0), hc), hm);
            hm = hc; hc = hn;
            ah[n + 1] = hn;
        }
    }
    if (want_j) {
        // sin z = (e^{iz} - e^{-iz}) / 2i, cos z = (e^{iz} + e^{-iz}) / 2
        const cplx df = csub(eiz, emiz), sm = cadd(eiz, emiz);
        const cplx sn = cmake(0.5 * df.y, -0.5 * df.x), cs = cscale(sm, 0.5);
        cplx f0 = cmul(sn, iz);
        cplx f1 = cmul(csub(f0, cs), iz);
        const double az = cabs2(z);
        if (az < 0.5) {
            // series for j1 (cancellation in the closed form): z/3 (1 - z^2/10 (1 - z^2/28 (1 - z^2/54 (1 - z^2/88 (1 - z^2/130)))))
            const cplx z2 = cmul(z, z);
            cplx t = cmake(1.0, 0.0);
            const double den[5] = {130.0, 88.0, 54.0, 28.0, 10.0};
            for (int q = 0; q < 5; ++q) t = csub(cmake(1.0, 0.0), cmul(cscale(z2, 1.0 / den[q]), t));
            f1 = cmul(cscale(z, 1.0 / 3.0), t);
        }
        // Upward recurrence only in the oscillatory region close to the real axis: away from it j_n behaves like the
        // modified function i_n, which is the MINIMAL solution for every n (z = 40i lost 9 digits at n = 30 going upward).
        if ((double)n_top + 1.0 <= fabs(z.x) && fabs(z.y) <= 1.0) {
            aj[0] = f0;
            if (n_top >= 1) aj[1] = f1;
            cplx jm = f0, jc = f1;
            for (int n = 1; n < n_top; ++n) {
                cplx jn = csub(cmul(cscale(iz, 2.0 * n + 1.0), jc), jm);
                jm = jc; jc = jn;
                aj[n + 1] = jn;
            }
        } else {
            const int M = miller_start(n_top, az) + (int)(2.0 * fabs(z.y));
            cplx jp1 = cmake(0.0, 0.0), jc = cmake(1e-280, 0.0);
            for (int m = M; m >= 1; --m) {
                cplx jm1 = csub(cmul(cscale(iz, 2.0 * m + 1.0), jc), jp1);
                jp1 = jc; jc = jm1;  // jc = order m-1
                if (m - 1 <= n_top) aj[m - 1] = jc;
                if (cabs1(jc) > 1e200) {
                    const double sc = 1e-200;
                    jc = cscale(jc, sc); jp1 = cscale(jp1, sc);
                    for (int q = m - 1; q <= n_top; ++q) aj[q] = cscale(aj[q], sc);
                }
            }
            // normalise on the larger of the two closed-form start values
            cplx nrm;
            if (cabs1(f0) >= cabs1(f1) || n_top == 0) {
                nrm = (cabs1(f0) >= cabs1(f1)) ? cdiv(f0, jc) : cdiv(f1, jp1);
            } else {
                nrm = cdiv(f1, aj[1]);
            }
            for (int q = 0; q <= n_top; ++q) aj[q] = cmul(aj[q], nrm);
        }
    }
}
// derivative: z_n' = (n/z) z_n - z_{n+1}
__device__ __forceinline__ cplx radial_deriv_z(int n, cplx iz, cplx zn, cplx znp1) {
    return csub(cmul(cscale(iz, (double)n), zn), znp1);
}

// ---- cylindrical family, complex argument -----------------------------------------------------------------
__device__ __forceinline__ cplx csqrt_(cplx a) {
    const double m = hypot(a.x, a.y);
    if (m == 0.0) return cmake(0.0, 0.0);
    double sr = sqrt(0.5 * (m + fabs(a.x)));
    double si = 0.5 * a.y / sr;
    if (a.x >= 0.0) return cmake(sr, si);
    return cmake(fabs(si), a.y >= 0.0 ? sr : -sr);
}
__device__ __forceinline__ cplx clog_(cplx a) { return cmake(log(hypot(a.x, a.y)), atan2(a.y, a.x)); }
__device__ __forceinline__ cplx cexp_(cplx a) {
    double s, c;
    sincos(a.y, &s, &c);
    const double e = exp(a.x);
    return cmake(e * c, e * s);
}

// H^{(1)}_0(z), H^{(1)}_1(z) by the Hankel asymptotic series, |z| >= 18 (smallest term ~ e^{-2|z|} < 3e-16)
__device__ inline void cyl_h01_asym(cplx z, cplx& h0, cplx& h1) {
    const cplx iz = crecip(z);
    const cplx pref = csqrt_(cscale(iz, BHS_2_PI));  // sqrt(2 / (pi z))
    const cplx e = cexp_i(z);
    const double r2 = 0.70710678118654752440;
#pragma unroll
    for (int nu = 0; nu < 2; ++nu) {
        const double mu = 4.0 * nu * nu;
        cplx sum = cmake(1.0, 0.0), term = cmake(1.0, 0.0);
        const cplx i8z = cmul(cmake(0.0, 0.125), iz);  // i / (8 z)
        double last = 1.0;
        for (int kk = 1; kk < 80; ++kk) {
            const double f = (mu - (2.0 * kk - 1.0) * (2.0 * kk - 1.0)) / kk;
            const cplx nt = cmul(term, cscale(i8z, f));
            const double mag = cabs1(nt);
            if (mag >= last && kk > 2) break;
            term = nt;
            last = mag;
            sum = cadd(sum, term);
            if (mag < 1e-18) break;
        }
        // e^{i(z - nu pi/2 - pi/4)} = e^{iz} * e^{-i pi/4} * (-i)^nu
        cplx ph = cmul(e, cmake(r2, -r2));
        if (nu == 1) ph = cmake(ph.y, -ph.x);
        const cplx v = cmul(cmul(pref, ph), sum);
        if (nu == 0) h0 = v; else h1 = v;
    }
}

// K_0(w), K_1(w) for Re w > 0, |w| >= 2: Steed's continued fraction CF2 (Temme; Numerical Recipes `bessik`, mu = 0),
// in complex arithmetic.  Used through H^{(1)}_nu(z) = (2 / (pi i)) e^{-i nu pi / 2} K_nu(-i z).
__device__ inline void cyl_k01_cf2(cplx w, cplx& k0, cplx& k1) {
    cplx b = cscale(cadd(cmake(1.0, 0.0), w), 2.0);
    cplx d = crecip(b), h = d, delh = d;
    cplx q1 = cmake(0.0, 0.0), q2 = cmake(1.0, 0.0);
    const double a1 = 0.25;  // 1/4 - mu^2
    cplx q = cmake(a1, 0.0), c = q;
    double a = -a1;
    cplx s = cadd(cmake(1.0, 0.0), cmul(q, delh));
    for (int i = 2; i < 2000; ++i) {
        a -= 2.0 * (i - 1);
        c = cscale(c, -a / i);
        const cplx qnew = cscale(csub(q1, cmul(b, q2)), 1.0 / a);
        q1 = q2;
        q2 = qnew;
        q = cadd(q, cmul(c, qnew));
        b = cadd(b, cmake(2.0, 0.0));
        d = crecip(cadd(b, cscale(d, a)));
        delh = cmul(csub(cmul(b, d), cmake(1.0, 0.0)), delh);
        h = cadd(h, delh);
        const cplx dels = cmul(q, delh);
        s = cadd(s, dels);
        if (cabs1(dels) < 1e-17 * cabs1(s)) break;
    }
    h = cscale(h, a1);
    // K_0 = sqrt(pi / (2 w)) e^{-w} / s ;  K_1 = K_0 (w + 1/2 - h) / w
    const cplx iw = crecip(w);
    k0 = cmul(cmul(csqrt_(cscale(iw, 1.57079632679489661923)), cexp_(cmake(-w.x, -w.y))), crecip(s));
    k1 = cmul(cmul(k0, csub(cadd(w, cmake(0.5, 0.0)), h)), iw);
}

// J_0..J_{n_top} (Miller, normalised on e^{-+iz} = J_0 + 2 sum (-+i)^k J_k, which has no cancellation on the side of the
// real axis where it is used) and H^{(1)}_0..H^{(1)}_{n_top} (upward from the orders 0, 1).  z != 0.
template <typename ArrJ, typename ArrH>
__device__ void cyl_sequence_z(cplx z, int n_top, ArrJ aj, ArrH ah, bool want_j, bool want_h) {
    const double az = cabs2(z);
    const cplx iz = crecip(z);
    const bool asym = az >= 18.0;
    const bool via_k = !asym && z.y > 3.0 && az >= 2.0;  // H from K_nu(-iz); else (near the real axis) from J + i Y
    const bool need_miller = want_j || (want_h && !asym && !via_k);
    cplx j0 = cmake(0.0, 0.0), j1 = j0, ysum = j0, y1sum = j0;
    if (need_miller) {
        int M = miller_start(want_j ? n_top : 1, az) + (int)(2.0 * fabs(z.y));
        M += (M & 1);  // even
        const bool up = z.y >= 0.0;  // normalise on e^{-iz} (|.| = e^{Im z}) above the real axis, e^{+iz} below
        cplx jp1 = cmake(0.0, 0.0), jc = cmake(1e-280, 0.0);
        cplx nsum = cmake(0.0, 0.0);
        for (int m = M; m >= 1; --m) {
            // at loop top: jc = J_m, jp1 = J_{m+1}
            const cplx jm1 = csub(cmul(cscale(iz, 2.0 * m), jc), jp1);
            // 2 (-+i)^m J_m
            nsum = cadd(nsum, cscale(cmul_ipow(jc, up ? -m : m), 2.0));
            if ((m & 1) == 0) {
                const int kk = m >> 1;
                const double sg = (kk & 1) ? -1.0 : 1.0;
                ysum = cadd(ysum, cscale(jc, sg / kk));
                y1sum = cadd(y1sum, cscale(csub(jm1, jp1), sg / kk));
            }
            jp1 = jc;
            jc = jm1;  // order m-1
            if (want_j && m - 1 <= n_top) aj[m - 1] = jc;
            if (cabs1(jc) > 1e200) {
                const double sc = 1e-200;
                jc = cscale(jc, sc); jp1 = cscale(jp1, sc); nsum = cscale(nsum, sc);
                ysum = cscale(ysum, sc); y1sum = cscale(y1sum, sc);
                if (want_j) for (int q = m - 1; q <= n_top; ++q) aj[q] = cscale(aj[q], sc);
            }
        }
        nsum = cadd(nsum, jc);  // + J_0
        const cplx target = cexp_i(up ? cmake(-z.x, -z.y) : z);
        const cplx nrm = cdiv(target, nsum);
        j0 = cmul(jc, nrm);
        j1 = cmul(jp1, nrm);
        ysum = cmul(ysum, nrm);
        y1sum = cmul(y1sum, nrm);
        if (want_j) for (int q = 0; q <= n_top; ++q) aj[q] = cmul(aj[q], nrm);
    }
    if (want_h) {
        cplx h0, h1;
        if (asym) {
            cyl_h01_asym(z, h0, h1);
        } else if (via_k) {
            cplx k0, k1;
            cyl_k01_cf2(cmake(z.y, -z.x), k0, k1);  // w = -i z
            // H0 = (2 / (pi i)) K0 = -i (2/pi) K0 ;  H1 = -(2/pi) K1
            h0 = cscale(cmake(k0.y, -k0.x), BHS_2_PI);
            h1 = cscale(k1, -BHS_2_PI);
        } else {
            const cplx lg = cadd(clog_(cscale(z, 0.5)), cmake(BHS_EULER, 0.0));
            const cplx y0 = cscale(csub(cmul(lg, j0), cscale(ysum, 2.0)), BHS_2_PI);
            const cplx y1 = cscale(cadd(csub(cmul(lg, j1), cmul(j0, iz)), y1sum), BHS_2_PI);
            h0 = cadd(j0, cmake(-y0.y, y0.x));
            h1 = cadd(j1, cmake(-y1.y, y1.x));
        }
        ah[0] = h0;
        if (n_top >= 1) ah[1] = h1;
        cplx hm = h0, hc = h1;
        for (int n = 1; n < n_top; ++n) {
            const cplx hn = csub(cmul(cscale(iz, 2.0 * n), hc), hm);
            hm = hc; hc = hn;
            ah[n + 1] = hn;
        }
    }
}

// z_n^{(d)} for n = 0..n_max_plus (complex argument), any d >= 2: arrays must hold n_max_plus + 1 + shift entries.
template <typename ArrJ, typename ArrH>
__device__ void radial_sequence_z(int d, cplx z, int n_max_plus, ArrJ aj, ArrH ah, bool want_j, bool want_h) {
    const int even_dim = (d & 1) == 0;
    const int shift = even_dim ? d / 2 - 1 : (d - 3) / 2;
    if (even_dim) cyl_sequence_z(z, n_max_plus + shift, aj, ah, want_j, want_h);
    else sph_sequence_z(z, n_max_plus + shift, aj, ah, want_j, want_h);
    if (shift > 0 || even_dim) {
        cplx sc = cmake(even_dim ? BHS_SQRT_PI_2 : 1.0, 0.0);
        const cplx iz = crecip(z);
        for (int s = 0; s < shift; ++s) sc = cmul(sc, iz);
        for (int n = 0; n <= n_max_plus; ++n) {
            if (want_j) aj[n] = cmul(aj[n + shift], sc);
            if (want_h) ah[n] = cmul(ah[n + shift], sc);
        }
    }
}
