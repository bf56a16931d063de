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
        if ((double)n_top + 1.0 <= az) {
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
